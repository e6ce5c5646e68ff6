#!/bin/bash
# round 2, GPU call 4: sticky chunk order A/B (C5-shaped), mid-batch timelines, fp32 timings, ncu after-capture
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
show() { python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('$1', 'Q', d['config']['queries_per_step'], 'rows', d['config']['global_rows'], 'dim', d['config']['dim'], 'step ms', round(d['ms_per_step'],4), 'blocking', round(d['blocking_call_ms'],4), 'kernel avg ms', round(r['avg_launch_ms'],4), 'TF/s', round(r.get('tflops', r.get('achieved')),1), 'GB/s', round(r.get('hbm_gbs', r.get('achieved')),0), 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'])
        print('    ', [(t['kernel'], round(t['ms']*1000,1)) for t in d.get('kernel_timeline_ms', [])])
    elif 'rror' in l: print(l.rstrip())
"; }
timeout 600 python -m pytest tests/test_search_gpu.py -m gpu -x -q > gpurun_out/r02_call4_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_call4_pytest.log
for ST in 0 1 0 1; do
  MMRS_K2_STICKY=$ST timeout 300 python bench.py --rows 4000000 --dim 768 --batch 4096 --steps 10 --warmup 3 --no-cpu --legs none 2>&1 | show "sticky=$ST"
done | tee gpurun_out/r02_sticky_ab.log
for B in 32 64 128 256; do
  timeout 300 python bench.py --rows 1000000 --dim 512 --batch $B --steps 100 --warmup 5 --no-cpu --legs none 2>&1 | show "1Mx512"
done | tee gpurun_out/r02_timeline_midbatch.log
timeout 300 python tools/bench_fp32.py 2>&1 | tee gpurun_out/r02_fp32_bench_before.log
timeout 300 python tools/prof_search.py --rows 2000000 --dim 768 --batch 1024 --iters 2 > gpurun_out/plain_c5like.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_mma_kernel -s 2 -c 1 \
    -o gpurun_out/r02_k2_c5like_sticky -f python tools/prof_search.py --rows 2000000 --dim 768 --batch 1024 --iters 2 > gpurun_out/ncu_c5like.log 2>&1
echo "ncu c5like rc=$?"
