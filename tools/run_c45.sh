#!/bin/bash
# BASELINE configs C4 / C5 on N GPUs (run under gpurun --gpus N): 100M x 768 bf16 sharded row-wise
N=${1:-8}
ROWS=$((100000000 / N))
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --rows $ROWS --dim 768 --no-cpu "${@:2}" 2>&1 | grep "^{" ; }
echo "== C4 global batch 16";  run 29521 --global-batch 16 --steps 20 --warmup 3 | tee gpurun_out/c4_n${N}_q16.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline'].get('whole_step_frac'))"
echo "== C4 global batch 128"; run 29522 --global-batch 128 --steps 20 --warmup 3 | tee gpurun_out/c4_n${N}_q128.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline'].get('whole_step_frac'))"
echo "== C5 global batch 65536"; run 29523 --global-batch 65536 --steps 2 --warmup 1 | tee gpurun_out/c5_n${N}.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline'])"
