"""Small end-to-end pass over every kernel for compute-sanitizer --tool memcheck (run under gpurun)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import mmrs_b200
from oracle import oracle
g16 = oracle.synthetic_gallery(20_001, 72, seed=1, dtype=torch.bfloat16)      # ragged N, D with a K tail
g32 = g16.to(torch.float32)
q = oracle.synthetic_queries(37, 72, seed=2)
gal16, gal32 = mmrs_b200.DeviceGallery(g16), mmrs_b200.DeviceGallery(g32, mode="fp32")
for nq in (1, 2, 5, 37):
    v, i = mmrs_b200.search_topk(q[:nq], gal16, 10)                 # K1 / K2 small
    assert torch.equal(i, oracle.search_topk(q[:nq], g16, 10, mode="bf16")[1]) or nq > 2
v, i = mmrs_b200.search_topk(oracle.synthetic_queries(300, 72, seed=3), gal16, 10)   # K2 wide (16 epilogue warps) + second pass
v, i = mmrs_b200.search_topk(q[:3], gal32, 10)                      # K1 fp32
v, i = mmrs_b200.search_topk(q, gal32, 10)                          # K2 x 3 (fp32 on tensor cores)
s = mmrs_b200.full_scores(q[:5], gal16)
s = mmrs_b200.full_scores(q[:2], gal32)
x, planted = oracle.synthetic_dedup(3000, 64, dup_frac=0.05, seed=4)
assert [tuple(p) for p in mmrs_b200.find_duplicate_pairs(x, 0.95, method="tc").tolist()] == planted
assert [tuple(p) for p in mmrs_b200.find_duplicate_pairs(x[:600], 0.95, method="fp32").tolist()] == [p for p in planted if p[1] < 600]
sc = mmrs_b200.full_scores(q[:1].cuda(), gal32, scale=100.0)[0]
mmrs_b200.best_threshold_on_device(sc, np.arange(20_001) % 5, 2)
mmrs_b200.find_thresholds(sc.cpu().numpy()[:4000], sc.cpu().numpy()[4000:], "x")
torch.cuda.synchronize()
print("sanitize pass ok")
