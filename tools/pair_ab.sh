#!/bin/bash
# A/B of the CTA-pair K2 against the single-CTA kernel (run under gpurun)
show() { python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('$1', d['config']['workload'][:28], 'Q', d['config']['queries_per_step'], 'step ms', round(d['ms_per_step'],4), 'qps', int(d['value']), 'kernel', r.get('kernel'), 'avg ms', round(r['avg_launch_ms'],4), 'TF/s', round(r.get('tflops', r.get('achieved')),1), 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'])
    elif 'rror' in l: print(l.rstrip())
"; }
for NP in 0 1; do
  if [ $NP = 1 ]; then export MMRS_K2_NO_PAIR=1; else unset MMRS_K2_NO_PAIR; fi
  timeout 300 python bench.py --rows 4000000 --dim 768 --batch 4096 --steps 20 --warmup 3 --no-cpu 2>&1 | show "nopair=$NP"
  timeout 300 python bench.py --batch 256 --steps 200 --warmup 5 --no-cpu 2>&1 | show "nopair=$NP"
  timeout 300 python bench.py --batch 1024 --steps 100 --warmup 5 --no-cpu 2>&1 | show "nopair=$NP"
done
