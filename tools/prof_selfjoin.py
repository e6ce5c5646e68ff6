"""Tiny driver for ncu: one tensor-core self-join of N x 512 unit rows."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from mmrs_b200.dedup import selfjoin_tc_raw
dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 150_000
x = torch.randn((n, 512), generator=torch.Generator(device=dev).manual_seed(0), device=dev)
x[n // 2:n // 2 + 1000] = x[:1000] + 0.1 * torch.randn((1000, 512), device=dev)
x /= x.norm(dim=-1, keepdim=True)
x16 = x.to(torch.bfloat16)
for _ in range(2):
    p = selfjoin_tc_raw(x, 0.95, x16=x16)
torch.cuda.synchronize()
print("ok", p.shape[0])
