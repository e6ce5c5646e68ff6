"""Tiny driver for ncu: a few searches of one batch size on the C2 gallery (or a given shape)."""
import argparse, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import mmrs_b200
from bench import device_gallery_shard

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--rows", type=int, default=1_000_000)
ap.add_argument("--dim", type=int, default=512)
ap.add_argument("--k", type=int, default=100)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--path", default="auto")
a = ap.parse_args()
dev = torch.device("cuda", 0)
gal = mmrs_b200.DeviceGallery(device_gallery_shard(torch, a.rows, a.dim, 0, 0, dev))
q = torch.randn((a.batch, a.dim), generator=torch.Generator().manual_seed(1)).to(dev)
for _ in range(a.iters):
    v, i = mmrs_b200.search_topk(q, gal, a.k, path=a.path)
torch.cuda.synchronize()
print("ok", v[0, :3].tolist())
