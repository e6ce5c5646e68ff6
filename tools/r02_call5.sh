#!/bin/bash
# round 2, GPU call 5: fp32 pair+split path, HBM_PAIRS A/B for the mid-batch regime
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
show() { python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('$1', 'Q', d['config']['queries_per_step'], 'dim', d['config']['dim'], 'step ms', round(d['ms_per_step'],4), 'blocking', round(d['blocking_call_ms'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'kernel avg ms', round(r['avg_launch_ms'],4), 'whole_step_frac', round(r.get('whole_step_frac', 0),3), 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'])
    elif 'rror' in l: print(l.rstrip())
"; }
timeout 600 python -m pytest tests/test_search_gpu.py -m gpu -x -q -k "fp32 or c1 or pair" > gpurun_out/r02_call5_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_call5_pytest.log
timeout 300 python tools/bench_fp32.py 2>&1 | tee gpurun_out/r02_fp32_bench_after.log
for HP in 74 64 56; do
  for B in 96 128; do
    MMRS_K2_HBM_PAIRS=$HP timeout 300 python bench.py --rows 1000000 --dim 512 --batch $B --steps 200 --warmup 10 --no-cpu --legs none 2>&1 | show "hbm_pairs=$HP"
  done
  MMRS_K2_HBM_PAIRS=$HP timeout 300 python bench.py --rows 1000000 --dim 768 --batch 128 --steps 200 --warmup 10 --no-cpu --legs none 2>&1 | show "hbm_pairs=$HP"
  MMRS_K2_HBM_PAIRS=$HP MMRS_K2_SMALL_MAX=32 MMRS_K2_PAIR_MIN=32 timeout 300 python bench.py --rows 1000000 --dim 512 --batch 64 --steps 200 --warmup 10 --no-cpu --legs none 2>&1 | show "pair64 hbm_pairs=$HP"
done | tee gpurun_out/r02_hbm_pairs_ab.log
timeout 300 python bench.py --rows 1000000 --dim 512 --batch 64 --steps 200 --warmup 10 --no-cpu --legs none 2>&1 | show "default Q64"  | tee -a gpurun_out/r02_hbm_pairs_ab.log
timeout 300 python bench.py --rows 1000000 --dim 512 --batch 16 --steps 200 --warmup 10 --no-cpu --legs none 2>&1 | show "default Q16"  | tee -a gpurun_out/r02_hbm_pairs_ab.log
