#!/bin/bash
# round 2, GPU call 6: L2 prefetch A/B, ncu after-captures (Q = 128 on C2, last phase of the C4 shard)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
show() { python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('$1', 'Q', d['config']['queries_per_step'], 'rows', d['config']['global_rows'], 'dim', d['config']['dim'], 'step ms', round(d['ms_per_step'],4), 'blocking', round(d['blocking_call_ms'],4), 'kernel avg ms', round(r['avg_launch_ms'],4), 'frac', round(r['frac'],3), 'whole_step_frac', round(r.get('whole_step_frac', 0),3), 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'])
    elif 'rror' in l: print(l.rstrip())
"; }
timeout 600 python -m pytest tests/test_search_gpu.py -m gpu -x -q -k "mma or pair or c4 or many or full_size" > gpurun_out/r02_call6_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_call6_pytest.log
for PF in 0 1; do
  for B in 16 64 128 256; do
    MMRS_K2_PREFETCH=$PF timeout 300 python bench.py --rows 1000000 --dim 512 --batch $B --steps 200 --warmup 10 --no-cpu --legs none 2>&1 | show "prefetch=$PF"
  done
  MMRS_K2_PREFETCH=$PF timeout 300 python bench.py --rows 1000000 --dim 768 --batch 128 --steps 200 --warmup 10 --no-cpu --legs none 2>&1 | show "prefetch=$PF"
  MMRS_K2_PREFETCH=$PF timeout 300 python bench.py --rows 12500000 --dim 768 --batch 16 --steps 50 --warmup 5 --no-cpu --legs none 2>&1 | show "prefetch=$PF"
  MMRS_K2_PREFETCH=$PF timeout 300 python bench.py --rows 4000000 --dim 768 --batch 4096 --steps 10 --warmup 3 --no-cpu --legs none 2>&1 | show "prefetch=$PF"
done | tee gpurun_out/r02_prefetch_ab.log
timeout 300 python tools/prof_search.py --batch 128 --iters 3 > gpurun_out/plain_q128.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_mma_kernel -s 2 -c 1 \
    -o gpurun_out/r02_k2_q128_after -f python tools/prof_search.py --batch 128 --iters 3 > gpurun_out/ncu_q128.log 2>&1
echo "ncu q128 rc=$?"
timeout 300 python tools/prof_search.py --rows 12500000 --dim 768 --batch 16 --iters 2 > gpurun_out/plain_c4shard.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_mma_kernel -s 4 -c 1 \
    -o gpurun_out/r02_k2_c4shard_q16 -f python tools/prof_search.py --rows 12500000 --dim 768 --batch 16 --iters 2 > gpurun_out/ncu_c4shard.log 2>&1
echo "ncu c4shard rc=$?"
