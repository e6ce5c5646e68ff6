#!/bin/bash
# round 2, GPU call 1: sanity of the round-1 library under the new bench + "before" ncu captures
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r02_call1_smi.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_call1_pytest.log 2>&1; echo "pytest rc=$?"
timeout 900 python bench.py > gpurun_out/r02_bench_n1_first.json 2> gpurun_out/r02_bench_n1_first.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r02_bench_n1_first.err
# before-capture: Q = 128 pair-mode last-phase scan on C2 (1M x 512)
timeout 300 python tools/prof_search.py --batch 128 --iters 3 > gpurun_out/plain_q128.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_mma_kernel -s 2 -c 1 \
    -o gpurun_out/r02_k2_q128_before -f python tools/prof_search.py --batch 128 --iters 3 > gpurun_out/ncu_q128.log 2>&1
echo "ncu q128 rc=$?"
# headline-shape capture: last-phase scan at the per-GPU shape of the 8-GPU C4 run (12.5M x 768, Q = 16)
timeout 300 python tools/prof_search.py --rows 12500000 --dim 768 --batch 16 --iters 2 > gpurun_out/plain_c4shard.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_mma_kernel -s 2 -c 1 \
    -o gpurun_out/r02_k2_c4shard_q16 -f python tools/prof_search.py --rows 12500000 --dim 768 --batch 16 --iters 2 > gpurun_out/ncu_c4shard.log 2>&1
echo "ncu c4shard rc=$?"
