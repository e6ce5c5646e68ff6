#!/bin/bash
# sweep the phase-schedule knobs on the C2 workload (run under gpurun)
for R in 4 5 6; do for DT in 16 32 64; do
  echo "ratio_log2=$R dense_tiles=$DT"
  MMRS_RATIO_LOG2=$R MMRS_DENSE_TILES=$DT timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu --sweep 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  step16 ms', round(d['ms_per_step'],4), 'e2e ms', round(d['e2e']['ms_per_step'],4), 'kern', round(d['roofline']['avg_launch_ms'],4)); print('  sweep', [(s['batch'], round(s['ms'],3)) for s in d['sweep']])
"
done; done
