"""World-size-2 run of the sharded search / self-join protocol on CPU (gloo).  The CUDA kernels
are replaced by the oracle through ShardedGallery's injection points, so what is tested is the
host logic: shard bounds, global indices, gather layout, padding of short shards, tie rule."""
import os
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


class _HostShard:
    """Stand-in for DeviceGallery on CPU (tests only)."""
    def __init__(self, data, row_offset):
        self.data, self.row_offset = data, row_offset
    def __len__(self):
        return self.data.shape[0]


def _worker(rank, world, port, n_rows, k, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mmrs_b200
    from oracle import oracle
    g = oracle.synthetic_gallery(n_rows, 32, seed=4, dtype=torch.float32)
    if n_rows > 700:
        g[5] = g[n_rows - 3]          # an exact tie across the two shards
    q = oracle.synthetic_queries(6, 32)
    lo, hi = mmrs_b200.shard_bounds(n_rows, world)[rank]

    def local_search(queries, shard, kk, *, normalize_queries, scale, path):
        v, i = oracle.search_topk(queries, shard.data, kk, normalize_queries=normalize_queries, scale=scale)
        return v, i + shard.row_offset

    sg = mmrs_b200.ShardedGallery(_HostShard(g[lo:hi], lo), n_rows, local_search=local_search,
                                  merge=oracle.merge_topk)
    # the shortest shard is known on every rank (it decides whether the fused NVLink gather may be used:
    # only when every shard can contribute its full top-k)
    assert sg.min_shard_rows == min(b - a for a, b in mmrs_b200.shard_bounds(n_rows, world))
    assert not sg.fused_active
    v, i = sg.search_topk(q, k)
    want_v, want_i = oracle.search_topk(q, g, k)
    assert torch.equal(i, want_i), (rank, i, want_i)
    assert torch.equal(v, want_v)

    # self-join: triangular split, variable-length gather
    x, planted = oracle.synthetic_dedup(1500, 32, dup_frac=0.04, seed=6)

    def raw_join(emb, thr, r0, r1):
        p = oracle.dedup_pairs(emb, thr)
        return p[(p[:, 0] >= r0) & (p[:, 0] < r1)].flip(0).contiguous()   # unsorted on purpose

    pairs = sg.find_duplicate_pairs(x, 0.95, raw_join=raw_join)
    assert [tuple(p) for p in pairs.tolist()] == planted
    Path(out_dir, f"ok{rank}").write_text("ok")
    dist.destroy_process_group()


@pytest.mark.parametrize("n_rows,k", [(1000, 10), (200, 150)])   # second case: a shard shorter than k
def test_sharded_protocol_world2(tmp_path, n_rows, k):
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, n_rows, k, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
