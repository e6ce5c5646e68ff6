"""Self-join throughput on one B200 (run under gpurun): pair-dots/s and TFLOP/s vs the measured bf16 peak."""
import sys, json, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import mmrs_b200
from mmrs_b200.dedup import selfjoin_tc_raw, selfjoin_raw
from oracle import oracle

dev = torch.device("cuda", 0)
peaks = json.loads((Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text())
def make(n, d, seed=0):
    gen = torch.Generator(device=dev).manual_seed(seed)
    x = torch.randn((n, d), generator=gen, device=dev)
    idx = torch.randperm(n, generator=gen, device=dev)
    m = n // 100
    x[idx[m:2 * m]] = x[idx[:m]] + 0.1 * torch.randn((m, d), generator=gen, device=dev)
    return x / x.norm(dim=-1, keepdim=True), m
for n, d, method in [(50_000, 512, "fp32"), (50_000, 512, "tc"), (200_000, 512, "tc"), (500_000, 512, "tc"), (500_000, 768, "tc")]:
    x, m = make(n, d)
    x16 = x.to(torch.bfloat16)
    fn = (lambda: selfjoin_tc_raw(x, 0.95, x16=x16)) if method == "tc" else (lambda: selfjoin_raw(x, 0.95))
    p = fn(); torch.cuda.synchronize()
    t0 = time.perf_counter(); p = fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    pairs = n * (n - 1) / 2
    print(json.dumps({"n": n, "d": d, "method": method, "seconds": round(dt, 4), "found": int(p.shape[0]), "planted": m,
                      "pair_dots_per_s": pairs / dt, "tflops": 2 * d * pairs / dt / 1e12,
                      "frac_of_measured_bf16_peak": 2 * d * pairs / dt / 1e12 / peaks["bf16_tflops"]}))
