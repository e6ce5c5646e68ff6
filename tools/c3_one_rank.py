"""One rank's share of BASELINE config C3 (10 M x 512 self-join, 8-way panel dealing) on ONE GPU:
the panels of rank R of 8 are independent of the other ranks', so this is the per-GPU time of the
8-GPU run without its pair gather.  Usage: python tools/c3_one_rank.py [N] [ranks...]"""
import sys, time, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import mmrs_b200
from mmrs_b200.dedup import selfjoin_tc_raw

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
ranks = [int(a) for a in sys.argv[2:]] or [0, 7]
d, world = 512, 8
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(0)
x = torch.empty((n, d), dtype=torch.float32, device=dev)
step = 1 << 20
for lo in range(0, n, step):
    x[lo:lo + step] = torch.randn((min(step, n - lo), d), generator=gen, device=dev)
m = n // 100
perm = torch.randperm(n, generator=gen, device=dev)
src, dst = perm[:m], perm[m:2 * m]
for lo in range(0, m, step):
    x[dst[lo:lo + step]] = x[src[lo:lo + step]] + 0.1 * torch.randn((min(step, m - lo), d), generator=gen, device=dev)
for lo in range(0, n, step):
    blk = x[lo:lo + step]
    blk /= blk.norm(dim=-1, keepdim=True)
x16 = x.to(torch.bfloat16)
selfjoin_tc_raw(x[:65536], 0.95, x16=x16[:65536]); torch.cuda.synchronize()   # warm-up
total_pairs = n * (n - 1) / 2
for r in ranks:
    t0 = time.perf_counter()
    mine = selfjoin_tc_raw(x, 0.95, r, world, x16=x16, capacity=max(4096, 2 * m))
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(json.dumps({"config": f"C3 share of rank {r}/{world}: {n} x {d}", "seconds": round(dt, 3), "pairs_found_by_rank": int(mine.shape[0]),
                      "tflops_this_gpu": 2 * d * total_pairs / world / dt / 1e12}))
