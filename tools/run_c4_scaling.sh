#!/bin/bash
# BASELINE config C4 strong scaling: 100M x 768 bf16 split over N GPUs, batch 16 (run under gpurun --gpus N)
N=${1:-1}
ROWS=$((100000000 / N))
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --gpus 1 --rows $ROWS --dim 768 --no-cpu --global-batch 16 --steps 20 --warmup 3 2>&1 | grep "^{\|Error\|error" | tail -2 > gpurun_out/c4_n${N}_q16.json
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --rows $ROWS --dim 768 --no-cpu --global-batch 16 --steps 20 --warmup 3 2>&1 | grep "^{\|Error\|error" | tail -2 > gpurun_out/c4_n${N}_q16.json
fi
python -c "
import json
d=json.loads(open('gpurun_out/c4_n${N}_q16.json').read().strip().splitlines()[-1])
print('C4 N=$N rows/gpu', d['config']['rows_per_gpu'], 'ms/step', d['ms_per_step'], 'q/s', d['value'], 'e2e', d['e2e']['value'], 'kernel frac', d['roofline']['frac'], 'step frac', d['roofline'].get('whole_step_frac'))
"
