#!/bin/bash
# per-batch step time and in-situ kernel timeline on the C2 workload (run under gpurun)
for B in ${BATCHES:-1 4 16 64 128 256}; do timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu --batch $B 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('batch', d['config']['queries_per_step'], 'step ms', round(d['ms_per_step'],4), 'e2e ms', round(d['e2e']['ms_per_step'],4), 'qps', int(d['value']), 'hbm_frac_step', round(d['roofline'].get('whole_step_frac', d['roofline'].get('frac')),3)); print('   ', [(t['kernel'], round(t['ms']*1000,1)) for t in d['kernel_timeline_ms']])
    elif 'Error' in l or 'error' in l: print(l.rstrip())
"; done
