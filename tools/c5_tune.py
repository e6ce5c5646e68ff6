"""One GPU's share of C5 (12.5M x 768 shard, 65 536 queries): time per batch under plan knobs."""
import os, sys, time, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import mmrs_b200
from bench import device_gallery_shard
dev = torch.device("cuda", 0)
rows, dim, nq, k = int(os.environ.get("ROWS", 12_500_000)), 768, int(os.environ.get("NQ", 65536)), 100
gal = mmrs_b200.DeviceGallery(device_gallery_shard(torch, rows, dim, 0, 0, dev))
q = torch.randn(nq, dim, generator=torch.Generator().manual_seed(1)).to(dev)
mmrs_b200.search_topk(q[:2048], gal, k)
torch.cuda.synchronize()
for rep in range(2):
    t0 = time.perf_counter(); v, i = mmrs_b200.search_topk(q, gal, k); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(json.dumps({"ratio_log2": os.environ.get("MMRS_RATIO_LOG2"), "dense": os.environ.get("MMRS_DENSE_TILES"),
                  "seconds": round(dt, 4), "qps_8gpu_equiv": round(nq / dt), "tflops": round(2.0 * nq * rows * dim / dt / 1e12, 1)}))
