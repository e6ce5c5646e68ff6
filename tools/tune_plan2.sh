#!/bin/bash
for B in ${BATCHES:-32 128}; do for R in 3 4 5; do for DT in 16 32 64 128; do
  MMRS_RATIO_LOG2=$R MMRS_DENSE_TILES=$DT timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu --batch $B 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('batch $B ratio_log2=$R dense=$DT step ms', round(d['ms_per_step'],4), [(t['kernel'][:8], round(t['ms']*1000)) for t in d['kernel_timeline_ms']])
"
done; done; done
