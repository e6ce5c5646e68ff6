"""fp32 (exact) mode timings: C1 shape (10k x 512, 100 queries, top-10) and 1M x 512 fp32 gallery."""
import sys, time, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import mmrs_b200
dev = torch.device("cuda", 0)
def timeit(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n
gen = torch.Generator(device=dev).manual_seed(1)
for rows, dim, nq, k in [(10_000, 512, 100, 10), (1_000_000, 512, 1, 100), (1_000_000, 512, 8, 100), (1_000_000, 512, 16, 100), (1_000_000, 512, 100, 10)]:
    g = torch.randn((rows, dim), generator=gen, device=dev); g /= g.norm(dim=-1, keepdim=True)
    gal = mmrs_b200.DeviceGallery(g, mode="fp32")
    q = torch.randn((nq, dim), generator=gen, device=dev)
    dt = timeit(lambda: mmrs_b200.search_topk(q, gal, k))
    print(json.dumps({"mode": "fp32", "rows": rows, "dim": dim, "queries": nq, "k": k, "ms": round(dt * 1e3, 4), "qps": round(nq / dt),
                      "gallery_GB_per_s": round(rows * dim * 4 * ((nq + 7) // 8) / dt / 1e9, 1)}))
