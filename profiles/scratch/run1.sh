for pf in 0 1 2; do
  echo "== TRON_WIDE_PREFETCH=$pf"
  TRON_WIDE_PREFETCH=$pf python bench.py --workload cfg3 --lean --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('cfg3 step',d['ms_per_step'],'grid',d['roofline']['ms_per_launch'],d['roofline']['frac'])"
  TRON_WIDE_PREFETCH=$pf python profiles/forward_timing.py cfg5_adj64 cfg5_adj8
done
mkdir -p gpurun_out; ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2d_cfg5_launches.csv python profiles/forward_timing.py cfg5_fwd64 cfg5_adj64 > /dev/null 2>&1
python profiles/summarize_launches.py gpurun_out/r2d_cfg5_launches.csv
