#!/bin/bash
# round 2, GPU call 9 (8 GPUs): multi-GPU correctness at world 8, strong-scaling bench at N = 8 and N = 4
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_call9_gpus.txt
timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q -s > gpurun_out/r02_multigpu_n8_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02_multigpu_n8_pytest.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; echo "bench n8 rc=$?"
tail -c 800 gpurun_out/r02_bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 \
    bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02_bench_n4.json 2> gpurun_out/r02_bench_n4.err; echo "bench n4 rc=$?"
tail -c 400 gpurun_out/r02_bench_n4.err
