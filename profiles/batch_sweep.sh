#!/bin/bash
# bench.py at several slices-per-launch values, with and without the gridding/FFT stream overlap
for b in "$@"; do
  for mode in overlap serial; do
    if [ $mode = serial ]; then export TRON_NO_OVERLAP=1; else unset TRON_NO_OVERLAP; fi
    TRON_BATCH=$b python bench.py --steps 3 --warmup 2 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$mode batch $b  device %.2f ms  e2e %.2f ms  grid frac %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac']))"
  done
done
