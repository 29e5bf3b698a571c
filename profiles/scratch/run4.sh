python -m pytest tests -m gpu -x -q -k "wide or cfg5 or cfg3 or fp16 or coil or more_than_six" 2>&1 | tail -3
python bench.py --workload cfg3 --lean --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('cfg3 step',d['ms_per_step'],'grid',d['roofline']['ms_per_launch'],d['roofline']['frac'])"
python profiles/forward_timing.py cfg5_adj64 cfg5_adj8
