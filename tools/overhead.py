"""Where does a search step's time go besides the kernels?  (run under gpurun)"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import mmrs_b200
from mmrs_b200 import _cabi
from mmrs_b200.gallery import DeviceGallery
from bench import device_gallery_shard

dev = torch.device("cuda", 0)
lib = _cabi.lib
for rows in (2048, 1_000_000):
    gal = DeviceGallery(device_gallery_shard(torch, rows, 512, 0, 0, dev))
    for nq in (16,):
        q = torch.randn(nq, 512).to(dev)
        k = 100
        def py_call():
            return mmrs_b200.search_topk(q, gal, k)
        for _ in range(5): py_call()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(200): py_call()
        torch.cuda.synchronize(); t_py = (time.perf_counter() - t0) / 200
        ws_bytes = lib.mmrs_search_workspace_bytes(gal.n_rows, gal.padded_dim, gal.dtype_code, nq, k)
        ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device=dev)
        v = torch.empty((nq, k), dtype=torch.float32, device=dev); i = torch.empty((nq, k), dtype=torch.int64, device=dev)
        args = (gal.data.data_ptr(), gal.n_rows, gal.padded_dim, gal.data.stride(0), gal.dtype_code, q.data_ptr(), nq,
                q.stride(0), k, 1, 1.0, 0, 0, v.data_ptr(), i.data_ptr(), DeviceGallery.aligned_ptr(ws), ws_bytes,
                int(torch.cuda.current_stream().cuda_stream))
        for _ in range(5): lib.mmrs_search_topk(*args)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(200): lib.mmrs_search_topk(*args)
        torch.cuda.synchronize(); t_c = (time.perf_counter() - t0) / 200
        print(f"rows={rows} nq={nq}: python API {t_py*1e6:.1f} us/call, raw C ABI {t_c*1e6:.1f} us/call")
