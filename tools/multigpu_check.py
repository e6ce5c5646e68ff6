"""torchrun --nproc-per-node N tools/multigpu_check.py : sharded search + sharded self-join vs the oracle
(NCCL all-gather + CUDA merge).  Rank 0 prints 'multigpu ok'."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch, torch.distributed as dist
import mmrs_b200
from oracle import oracle

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n, d, k = 400_000, 512, 100
g = oracle.synthetic_gallery(n, d, seed=0, dtype=torch.bfloat16)
g[17] = g[n - 5]                                  # a tie across the first and last shard
sg = mmrs_b200.ShardedGallery.from_full(g, device=dev)
for nq in (3, 24):
    q = oracle.synthetic_queries(nq, d, seed=nq)
    v, i = sg.search_topk(q.to(dev), k)
    wv, wi = oracle.search_topk(q, g, k, mode="bf16")
    bad = (i.cpu() != wi).sum().item()
    assert bad <= 2, (rank, nq, bad)
    assert (v.cpu() - wv).abs().max().item() < 1e-5
assert sg._fused and sg._fused_ok, "the fused NVLink gather was not used"
# host queries in -> host results out; asynchronous handle
q = oracle.synthetic_queries(5, d, seed=9)
pend = sg.search_topk(q, k, sync=False)
hv, hi = pend.wait()
wv, wi = oracle.search_topk(q, g, k, mode="bf16")
assert not hv.is_cuda and (hi != wi).sum().item() <= 1
# a candidate-list overflow on ONE rank: every rank must fall back to the general path together
n2 = 140_032
g2 = oracle.synthetic_gallery(n2, 32, seed=5, dtype=torch.float32)
q2 = oracle.synthetic_queries(2, 32)
tiles = torch.arange(n2) // 128
hot = (tiles % 16 != 0) & (torch.arange(n2) < 70_016)          # rank 0's shard, outside its seed sample
g2[hot] = oracle.l2_normalize(g2[hot] + 2.0 * oracle.l2_normalize(q2)[0])
sg2 = mmrs_b200.ShardedGallery.from_full(g2, device=dev)
v2, i2 = sg2.search_topk(q2.to(dev), 10)
wv2, wi2 = oracle.search_topk(q2, g2, 10)
assert torch.equal(i2.cpu(), wi2), (rank, i2, wi2)
# zero-norm query: same error on every rank
qz = oracle.synthetic_queries(3, d); qz[1] = 0
try:
    sg.search_topk(qz.to(dev), k)
    raise SystemExit("zero-norm query was accepted")
except mmrs_b200._cabi.MmrsError as e:
    assert e.code == mmrs_b200._cabi.ERR_ZERO_NORM
x, planted = oracle.synthetic_dedup(30_000, 128, dup_frac=0.02, seed=5)
pairs = sg.find_duplicate_pairs(mmrs_b200.dedup._device_f32(x, dev), 0.95)
assert [tuple(p) for p in pairs.cpu().tolist()] == planted
# the NCCL packed-key path must agree with the fused one
os.environ["MMRS_NO_FUSED_GATHER"] = "1"
sg_nccl = mmrs_b200.ShardedGallery(sg.local, n)
q = oracle.synthetic_queries(7, d, seed=21).to(dev)
v_a, i_a = sg.search_topk(q, k)
v_b, i_b = sg_nccl.search_topk(q, k)
assert torch.equal(i_a, i_b) and torch.equal(v_a, v_b) and not sg_nccl._fused
dist.barrier()
if rank == 0:
    print(f"multigpu ok: world={world} (fused NVLink gather + NCCL path agree)")
dist.destroy_process_group()
