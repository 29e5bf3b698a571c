python -m pytest tests -m gpu -x -q -k "cfg5 or fp16 or wide" 2>&1 | tail -3
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2g_cfg5_launches.csv python profiles/forward_timing.py cfg5_fwd64 cfg5_adj64 > /dev/null 2>&1
python profiles/summarize_launches.py gpurun_out/r2g_cfg5_launches.csv | grep tronb
python profiles/forward_timing.py cfg5_fwd64 cfg5_adj64
