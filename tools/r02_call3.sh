#!/bin/bash
# round 2, GPU call 3 (2 GPUs): multi-GPU correctness test + the strong-scaling bench at N = 2
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_call3_gpus.txt
timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q -s > gpurun_out/r02_multigpu_n2_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r02_multigpu_n2_pytest.log
timeout 870 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench n2 rc=$?"
tail -c 1500 gpurun_out/r02_bench_n2.err
timeout 300 python bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_ref_n2.json 2>&1; echo "ref rc=$?"
