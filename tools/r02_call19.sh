#!/bin/bash
# round 2, GPU call 19 (8 GPUs): the driver's SCALE command at N = 8 on the final tree
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29801 \
    bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_n8_final.json 2> gpurun_out/r02_bench_n8_final.err; echo "bench n8 rc=$?"
tail -c 300 gpurun_out/r02_bench_n8_final.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29802 \
    bench.py --impl reference --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_ref_n8_final.json 2> gpurun_out/r02_ref_n8_final.err; echo "ref n8 rc=$?"
