#!/bin/bash
# round 2, GPU call 16 (N GPUs visible): multi-GPU correctness after the cheaper publish sequence + fused/NCCL timing
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q > gpurun_out/r02_multigpu_n${N}_pytest2.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02_multigpu_n${N}_pytest2.log | cut -c1-200
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}" 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l)
        print(d['gather'][:5], 'N', d['n_gpus'], 'rows', d['config']['global_rows'], 'dim', d['config']['dim'], 'Q', d['config']['queries_per_step'], 'step ms', round(d['ms_per_step'],4), 'e2e ms', round(d['e2e']['ms_per_step'],4), 'blocking', round(d['blocking_call_ms'],4), 'e2e blocking', round(d['e2e']['blocking_call_ms'],4), 'parity', d.get('parity_check',{}).get('ok'))
        print('    ', [(t['kernel'], round(t['ms']*1000,1)) for t in d.get('kernel_timeline_ms', [])][-5:])
"; }
P=29700
for rep in $(seq 1 ${REPS:-2}); do
  for F in "" "--no-fused"; do
    P=$((P+1)); run $P --rows $((N*1000000)) --dim 512 --batch 16 --steps 300 --warmup 10 --no-cpu --legs parity $F
    P=$((P+1)); run $P --rows $((N*1000000)) --dim 512 --batch 128 --steps 300 --warmup 10 --no-cpu --legs none $F
  done
done | tee gpurun_out/r02_fused_vs_nccl_n${N}_v2.log
