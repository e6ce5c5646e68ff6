#!/bin/bash
# round 2, GPU call 8: plan-knob sweep for the mid-batch regime, ncu captures of the final kernels, full N = 1 bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
show() { python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('$1', 'Q', d['config']['queries_per_step'], 'dim', d['config']['dim'], 'step ms', round(d['ms_per_step'],4), 'blocking', round(d['blocking_call_ms'],4), 'kernel avg ms', round(r['avg_launch_ms'],4), 'whole_step_frac', round(r.get('whole_step_frac', 0),3))
    elif 'rror' in l: print(l.rstrip())
"; }
for DT in 32 64; do for RL in 3 4; do
  for B in 64 128; do
    MMRS_DENSE_TILES=$DT MMRS_RATIO_LOG2=$RL timeout 300 python bench.py --rows 1000000 --dim 512 --batch $B --steps 300 --warmup 10 --no-cpu --legs none 2>&1 | show "dense=$DT ratio=$RL"
  done
  MMRS_DENSE_TILES=$DT MMRS_RATIO_LOG2=$RL timeout 300 python bench.py --rows 1000000 --dim 768 --batch 128 --steps 300 --warmup 10 --no-cpu --legs none 2>&1 | show "dense=$DT ratio=$RL"
done; done | tee gpurun_out/r02_plan_sweep_midbatch.log
timeout 300 python tools/prof_search.py --rows 12500000 --dim 768 --batch 16 --iters 2 > gpurun_out/plain_c4shard.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_mma_kernel -s 4 -c 1 \
    -o gpurun_out/r02_k2_c4shard_q16 -f python tools/prof_search.py --rows 12500000 --dim 768 --batch 16 --iters 2 > gpurun_out/ncu_c4shard.log 2>&1
echo "ncu c4shard rc=$?"
MMRS_SJ_PAIR=0 timeout 300 python tools/prof_selfjoin.py 150000 > gpurun_out/plain_sj.log 2>&1 &&
MMRS_SJ_PAIR=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:selfjoin_mma_kernel -s 1 -c 1 \
    -o gpurun_out/r02_selfjoin_mma_single -f python tools/prof_selfjoin.py 150000 > gpurun_out/ncu_sj.log 2>&1
echo "ncu selfjoin rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_ref_n1.json 2>&1; echo "ref rc=$?"
