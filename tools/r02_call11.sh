#!/bin/bash
# round 2, GPU call 11: two-phase plans for the 64-128-query regime (no mid phase)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
show() { python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('$1', 'Q', d['config']['queries_per_step'], 'dim', d['config']['dim'], 'step ms', round(d['ms_per_step'],4), 'blocking', round(d['blocking_call_ms'],4), 'kernel avg ms', round(r['avg_launch_ms'],4), 'whole_step_frac', round(r.get('whole_step_frac', 0),3))
        print('    ', [(t['kernel'], round(t['ms']*1000,1)) for t in d.get('kernel_timeline_ms', [])][:9])
    elif 'rror' in l: print(l.rstrip())
"; }
for CFG in "64 3" "64 5" "128 6" "128 3" "32 5"; do set -- $CFG
  for B in 64 128; do
    MMRS_DENSE_TILES=$1 MMRS_RATIO_LOG2=$2 timeout 300 python bench.py --rows 1000000 --dim 512 --batch $B --steps 300 --warmup 10 --no-cpu --legs none 2>&1 | show "dense=$1 ratio=$2"
  done
  MMRS_DENSE_TILES=$1 MMRS_RATIO_LOG2=$2 timeout 300 python bench.py --rows 1000000 --dim 768 --batch 128 --steps 300 --warmup 10 --no-cpu --legs none 2>&1 | show "dense=$1 ratio=$2"
done | tee gpurun_out/r02_plan_two_phase.log
