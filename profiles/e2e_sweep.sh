#!/bin/bash
# end-to-end (host buffers) and device-resident ms per cfg2 step for a list of environment settings
# usage: bash profiles/e2e_sweep.sh "A=1 B=2" "C=3" ...
for v in "$@"; do
  env $v python bench.py --lean --no-cpu --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('%-60s device %.3f ms  e2e %.3f ms  floor %.2f  grid %.4f ms/launch frac %.3f' % ('$v', d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['copy_floor_ms'], d['roofline']['ms_per_launch'], d['roofline']['frac']))
"
done
