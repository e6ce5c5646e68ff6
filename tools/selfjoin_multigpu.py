"""BASELINE config C3 under torchrun: N x 512 near-duplicate self-join, cos >= 0.95, sharded across the GPUs of
one box.  Every rank builds the same synthetic matrix on its device (seeded), 1 % of the rows are planted
near-duplicates (row_j = row_i + 0.1 * randn, cos ~ 0.995: nothing lies near tau), the tensor-core join
runs on this rank's column panels, pair lists are gathered and compared with the planted set."""
import os, sys, time, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch, torch.distributed as dist
import mmrs_b200

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n, d = int(os.environ.get("N", 10_000_000)), int(os.environ.get("D", 512))
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
planted = torch.stack([torch.minimum(src, dst), torch.maximum(src, dst)], 1)
planted = planted[torch.argsort(planted[:, 0] * n + planted[:, 1])]
sg = mmrs_b200.ShardedGallery(mmrs_b200.DeviceGallery(x[:128], mode="fp32"), n)   # only its rank/world are used
x16 = x.to(torch.bfloat16)
torch.cuda.synchronize(); dist.barrier()
from mmrs_b200.dedup import selfjoin_tc_raw
t0 = time.perf_counter()
mine = selfjoin_tc_raw(x, 0.95, rank, world, x16=x16, capacity=max(4096, 2 * m))
torch.cuda.synchronize(); t_local = time.perf_counter() - t0
dist.barrier(); t_join = time.perf_counter() - t0
pairs = sg.find_duplicate_pairs(x, 0.95)            # the public sharded call (join again + gather + sort)
torch.cuda.synchronize(); dist.barrier(); t_total = time.perf_counter() - t0 - t_join
ok = pairs.shape == planted.shape and bool(torch.equal(pairs, planted))
if rank == 0:
    total_pairs = n * (n - 1) / 2
    print(json.dumps({"config": f"C3: {n} x {d} self-join, cos >= 0.95, {world} GPUs", "pairs_found": int(pairs.shape[0]),
                      "planted": int(m), "exact_match_with_planted": ok, "join_seconds_max_rank": round(t_join, 3),
                      "public_call_seconds": round(t_total, 3), "pair_dots_per_s": total_pairs / t_join,
                      "tflops_aggregate": 2 * d * total_pairs / t_join / 1e12,
                      "frac_of_measured_bf16_peak": 2 * d * total_pairs / t_join / 1e12 / (world * 1623.3)}))
dist.destroy_process_group()
