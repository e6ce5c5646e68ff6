#!/bin/bash
# round 2, GPU call 13: "small pair" K2 variant (pairs with half-SM footprint) A/B for 33..128 queries
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
show() { python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('$1', 'Q', d['config']['queries_per_step'], 'dim', d['config']['dim'], 'step ms', round(d['ms_per_step'],4), 'blocking', round(d['blocking_call_ms'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'kernel avg ms', round(r['avg_launch_ms'],4), 'whole_step_frac', round(r.get('whole_step_frac', 0),3))
    elif 'rror' in l: print(l.rstrip())
"; }
timeout 600 python -m pytest tests/test_search_gpu.py -m gpu -x -q -k "mma or pair or many or full_size or retained" > gpurun_out/r02_call13_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_call13_pytest.log
for SP in 0 128 0 128; do
  for B in 96 128; do
    MMRS_K2_SMALL_PAIR_MAX=$SP timeout 300 python bench.py --rows 1000000 --dim 512 --batch $B --steps 300 --warmup 10 --no-cpu --legs none 2>&1 | show "small_pair_max=$SP"
  done
  MMRS_K2_SMALL_PAIR_MAX=$SP timeout 300 python bench.py --rows 1000000 --dim 768 --batch 128 --steps 300 --warmup 10 --no-cpu --legs none 2>&1 | show "small_pair_max=$SP"
  MMRS_K2_SMALL_PAIR_MAX=$SP MMRS_K2_SMALL_MAX=32 timeout 300 python bench.py --rows 1000000 --dim 512 --batch 64 --steps 300 --warmup 10 --no-cpu --legs none 2>&1 | show "small_max=32 small_pair_max=$SP"
  MMRS_K2_SMALL_PAIR_MAX=$SP MMRS_K2_SMALL_MAX=32 timeout 300 python bench.py --rows 1000000 --dim 768 --batch 64 --steps 300 --warmup 10 --no-cpu --legs none 2>&1 | show "small_max=32 small_pair_max=$SP"
done | tee gpurun_out/r02_small_pair_ab.log
timeout 300 python bench.py --rows 1000000 --dim 512 --batch 64 --steps 300 --warmup 10 --no-cpu --legs none 2>&1 | show "default" | tee -a gpurun_out/r02_small_pair_ab.log
timeout 300 python bench.py --rows 1000000 --dim 768 --batch 64 --steps 300 --warmup 10 --no-cpu --legs none 2>&1 | show "default" | tee -a gpurun_out/r02_small_pair_ab.log
