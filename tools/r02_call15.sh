#!/bin/bash
# round 2, GPU call 15 (8 GPUs): fused NVLink gather vs NCCL variant, same box, small and large shards
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 "${@:2}" 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l)
        print(d['gather'][:5], 'rows', d['config']['global_rows'], 'dim', d['config']['dim'], 'Q', d['config']['queries_per_step'], 'step ms', round(d['ms_per_step'],4), 'e2e ms', round(d['e2e']['ms_per_step'],4), 'blocking', round(d['blocking_call_ms'],4), 'e2e blocking', round(d['e2e']['blocking_call_ms'],4), 'parity', d.get('parity_check',{}).get('ok'))
        print('    ', [(t['kernel'], round(t['ms']*1000,1)) for t in d.get('kernel_timeline_ms', [])][-5:])
"; }
P=29600
for rep in 1 2; do
  for F in "" "--no-fused"; do
    P=$((P+1)); run $P --rows 8000000 --dim 512 --batch 16 --steps 300 --warmup 10 --no-cpu --legs parity $F
    P=$((P+1)); run $P --rows 8000000 --dim 512 --batch 128 --steps 300 --warmup 10 --no-cpu --legs none $F
  done
done | tee gpurun_out/r02_fused_vs_nccl_n8.log
for F in "" "--no-fused"; do
  P=$((P+1)); run $P --steps 20 --warmup 5 --no-cpu --legs none $F
done | tee -a gpurun_out/r02_fused_vs_nccl_n8.log
