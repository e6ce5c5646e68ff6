#!/bin/bash
# round 2, GPU call 17: K1 with software-pipelined row loads (R <= 4): fp32 timings + correctness
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_call17_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02_call17_pytest.log
timeout 300 python tools/bench_fp32.py 2>&1 | tee gpurun_out/r02_fp32_bench_after3.log
python - <<'PY' 2>&1 | tee -a gpurun_out/r02_fp32_bench_after3.log
import sys, time, json, torch
sys.path.insert(0, '.')
import mmrs_b200
dev = torch.device('cuda', 0)
gen = torch.Generator(device=dev).manual_seed(1)
g = torch.randn((1_000_000, 512), generator=gen, device=dev); g /= g.norm(dim=-1, keepdim=True)
gal = mmrs_b200.DeviceGallery(g, mode='fp32')
for nq in (2, 3, 4, 5, 6, 8):
    q = torch.randn((nq, 512), generator=gen, device=dev)
    for path in ('gemv',):
        for _ in range(5): mmrs_b200.search_topk(q, gal, 100, path=path)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(50): mmrs_b200.search_topk(q, gal, 100, path=path)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 50
        print(json.dumps({'mode': 'fp32 K1', 'queries': nq, 'ms': round(dt * 1e3, 4), 'frac_of_hbm': round(2.048e9 / dt / 6.5338e12, 3)}))
PY
