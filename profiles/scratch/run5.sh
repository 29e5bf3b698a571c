for hw in 0 2 3; do
echo "== TRON_WIDE_HALFWARP=$hw"
TRON_WIDE_HALFWARP=$hw python -m pytest tests -m gpu -x -q -k "wide_channel_gridding or cfg3 or more_than_six" 2>&1 | tail -1
TRON_WIDE_HALFWARP=$hw python bench.py --workload cfg3 --lean --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('cfg3 step',d['ms_per_step'],'grid',d['roofline']['ms_per_launch'],d['roofline']['frac'])"
TRON_WIDE_HALFWARP=$hw python profiles/forward_timing.py cfg5_adj64
done
