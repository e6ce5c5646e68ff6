"""torchrun --nproc-per-node N tools/multigpu_check.py : the sharded search (fused NVLink gather and NCCL
variant) and the sharded self-join against the oracle, on N >= 2 GPUs.  Rank 0 prints 'multigpu ok'.
Run by tests/test_multigpu_gpu.py (self-spawned, skipped on a single-GPU box) and by hand under
`gpurun --gpus N`; the logs of the runs are kept under profiles/."""
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
sg_nccl = mmrs_b200.ShardedGallery(sg.local, n, fused=False)


def unit_bf16(q):
    """Normalised and rounded to bf16 once, on the host: with more than ~100 queries the fused normalisation's
    last-ulp differences from torch flip a bf16 query element now and then (scores move by ~1e-4, inside
    north_star's 1e-2, outside the 1e-5 checked here) -- see tests/test_search_gpu.py::bf16_unit_queries."""
    return (q / q.norm(dim=-1, keepdim=True)).to(torch.bfloat16).to(torch.float32)


def near_equal(i, v, wi, wv, max_bad):
    bad = (i.cpu() != wi)
    assert int(bad.sum()) <= max_bad, (rank, int(bad.sum()))
    assert (v.cpu() - wv).abs().max().item() < 1e-5
    # a differing index must be a reference near-tie
    assert ((v.cpu() - wv).abs()[bad] < 2e-6).all()


for nq in (3, 24, 130):
    q = oracle.synthetic_queries(nq, d, seed=nq)
    norm = nq < 100
    if not norm:
        q = unit_bf16(q)
    v, i = sg.search_topk(q.to(dev), k, normalize_queries=norm)
    wv, wi = oracle.search_topk(q, g, k, mode="bf16", normalize_queries=norm)
    near_equal(i, v, wi, wv, 2 + nq // 16)
    vb, ib = sg_nccl.search_topk(q.to(dev), k, normalize_queries=norm)
    assert torch.equal(i, ib) and torch.equal(v, vb)          # fused and NCCL variants agree bit for bit
assert sg.fused_active and not sg_nccl.fused_active, "the fused NVLink gather was not used"

# host queries in -> host results out; asynchronous handles; caller-owned outputs; results retained
q = oracle.synthetic_queries(5, d, seed=9)
pend = sg.search_topk(q, k, sync=False)
hv, hi = pend.wait()
wv, wi = oracle.search_topk(q, g, k, mode="bf16")
assert not hv.is_cuda and (hi != wi).sum().item() <= 1
ov = torch.empty((5, k), dtype=torch.float32, device=dev); oi = torch.empty((5, k), dtype=torch.int64, device=dev)
rv, ri = sg.search_topk(q.to(dev), k, out=(ov, oi))
assert rv is ov and torch.equal(oi.cpu(), hi)
kept = [sg.search_topk(oracle.synthetic_queries(5, d, seed=100 + t).to(dev), k) for t in range(50)]
torch.cuda.synchronize()
for t in (0, 49):
    wv, wi = oracle.search_topk(oracle.synthetic_queries(5, d, seed=100 + t), g, k, mode="bf16")
    assert (kept[t][1].cpu() != wi).sum().item() <= 1

# two streams x several batches in flight with MORE queries than SMs: the merge selects of the two slots
# must not starve each other's scans across ranks (the waits run in one-warp kernels)
streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
qs = [unit_bf16(oracle.synthetic_queries(300, d, seed=40 + t)) for t in range(6)]
qd = [x.to(dev) for x in qs]
torch.cuda.synchronize()
pends = []
for t in range(6):
    streams[t % 2].wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(streams[t % 2]):
        pends.append(sg.search_topk(qd[t], k, sync=False, normalize_queries=False))
outs = [p.wait() for p in pends]
for t in (0, 5):
    wv, wi = oracle.search_topk(qs[t], g, k, mode="bf16", normalize_queries=False)
    near_equal(outs[t][1], outs[t][0], wi, wv, 40)
torch.cuda.synchronize()

# unequal shards: the last one just long enough for k (fused), and too short for it (NCCL variant, lists padded
# with key 0) -- in the second case most of the global top-k comes from the long shards, so a rank that
# contributed fewer than min(k, its rows) keys would be caught here
for last_rows, want_fused in ((110, True), (50, False), (10, False)):
    n_short = 128 * (world - 1) + last_rows
    gs = oracle.synthetic_gallery(n_short, 64, seed=7, dtype=torch.bfloat16)
    sgs = mmrs_b200.ShardedGallery.from_full(gs, device=dev)
    assert sgs.min_shard_rows == last_rows
    qq = oracle.synthetic_queries(4, 64, seed=3)
    v, i = sgs.search_topk(qq.to(dev), k)
    wv, wi = oracle.search_topk(qq, gs, k, mode="bf16")
    near_equal(i, v, wi, wv, 2)
    assert sgs.fused_active == want_fused, (n_short, sgs.fused_active, want_fused)

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
# the slot still works after the error
v, i = sg.search_topk(oracle.synthetic_queries(3, d, seed=3).to(dev), k)
wv, wi = oracle.search_topk(oracle.synthetic_queries(3, d, seed=3), g, k, mode="bf16")
near_equal(i, v, wi, wv, 2)
x, planted = oracle.synthetic_dedup(30_000, 128, dup_frac=0.02, seed=5)
pairs = sg.find_duplicate_pairs(mmrs_b200.dedup._device_f32(x, dev), 0.95)
assert [tuple(p) for p in pairs.cpu().tolist()] == planted
dist.barrier()
if rank == 0:
    print(f"multigpu ok: world={world} (fused NVLink gather + NCCL variant agree; short shards; 2 streams x 300 queries; "
          "overflow fallback; zero-norm; sharded self-join)")
dist.destroy_process_group()
