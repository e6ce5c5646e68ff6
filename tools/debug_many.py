import sys, os
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import torch, numpy as np
import mmrs_b200
from oracle import oracle
from test_search_gpu import torch_gpu_topk
nq = int(os.environ.get("NQ", "2500")); n = int(os.environ.get("N", "120000")); d = int(os.environ.get("D", "128")); k = int(os.environ.get("K", "10"))
g = oracle.synthetic_gallery(n, d, seed=12, dtype=torch.bfloat16)
q = oracle.synthetic_queries(nq, d, seed=13)
gal = mmrs_b200.DeviceGallery(g)
wv, wi = torch_gpu_topk(q, g.cuda(), k)
for path in ("mma",):
    v, i = mmrs_b200.search_topk(q.cuda(), gal, k, path=path)
    v, i = v.cpu(), i.cpu()
    badq = ((v - wv).abs() > 2e-6).any(dim=1).nonzero().flatten().tolist()
    print(path, "bad queries", len(badq), badq[:40])
    for qq in badq[:3]:
        print("  q", qq, "got", i[qq].tolist(), "\n      want", wi[qq].tolist())
        missing = set(wi[qq].tolist()) - set(i[qq].tolist())
        print("      missing rows", sorted(missing), "tiles", [m // 128 for m in missing], "tile%16", [(m // 128) % 16 for m in missing])
