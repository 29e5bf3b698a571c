python -m pytest tests -m gpu -x -q -k "wide or cfg5 or fp16 or coil" 2>&1 | tail -4
for pr in 0 1; do echo "== TRON_DEGRID_PAIR=$pr"; TRON_DEGRID_PAIR=$pr python profiles/forward_timing.py cfg5_fwd64 cfg5_fwd8; done
