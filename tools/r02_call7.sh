#!/bin/bash
# round 2, GPU call 7: carve-out A/B, SPLIT-as-template check on the C5 shape, fp32 timings
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
show() { python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('$1', 'Q', d['config']['queries_per_step'], 'rows', d['config']['global_rows'], 'dim', d['config']['dim'], 'step ms', round(d['ms_per_step'],4), 'blocking', round(d['blocking_call_ms'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'kernel avg ms', round(r['avg_launch_ms'],4), 'frac', round(r['frac'],3), 'whole_step_frac', round(r.get('whole_step_frac', 0),3), 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'])
        print('    ', [(t['kernel'], round(t['ms']*1000,1)) for t in d.get('kernel_timeline_ms', [])][:9])
    elif 'rror' in l: print(l.rstrip())
"; }
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_call7_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_call7_pytest.log
for NC in 1 0 1 0; do
  for B in 16 64 128; do
    MMRS_NO_CARVEOUT=$NC timeout 300 python bench.py --rows 1000000 --dim 512 --batch $B --steps 300 --warmup 10 --no-cpu --legs none 2>&1 | show "no_carveout=$NC"
  done
done | tee gpurun_out/r02_carveout_ab.log
timeout 300 python bench.py --rows 4000000 --dim 768 --batch 4096 --steps 10 --warmup 3 --no-cpu --legs none 2>&1 | show "c5like" | tee gpurun_out/r02_c5like_after.log
timeout 300 python tools/bench_fp32.py 2>&1 | tee gpurun_out/r02_fp32_bench_after2.log
