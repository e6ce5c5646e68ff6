#!/bin/bash
# round 2, GPU call 12: final tree -- full GPU suite, smoke, driver-like bench at N = 1, single-pass DRAM counters
# of the headline kernel at the 100M x 768 shape
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_final_smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err; echo "bench rc=$?"
timeout 300 python tools/prof_search.py --rows 100000000 --dim 768 --batch 16 --iters 1 > gpurun_out/plain_c4full.log 2>&1 &&
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:scan_mma_kernel -s 5 -c 1 --csv --log-file gpurun_out/r02_k2_c4full_q16_dram.csv \
    python tools/prof_search.py --rows 100000000 --dim 768 --batch 16 --iters 1 > gpurun_out/ncu_c4full.log 2>&1
echo "ncu c4full rc=$?"; tail -5 gpurun_out/r02_k2_c4full_q16_dram.csv
