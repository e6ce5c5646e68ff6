"""Row-sharded gallery over the GPUs of one box.

One process per GPU (`torch.distributed`, backend "nccl" over NVLink 5 / NVSwitch).  The gallery is
split into contiguous row blocks, queries are replicated, every rank runs the fused local top-k
on its shard with global row ids, ONE all-gather moves the `[Q, k]` (value, index) lists --
Q*k*12 bytes per rank -- and every rank merges the `world * k` candidates per query with the same
order rule (score desc, index asc).  There is no other exchange: the scan itself is embarrassingly
parallel, so scaling is weak in the gallery and the collective is latency-sized.

The reference has no distributed code at all; this is new.
"""
from __future__ import annotations

import math
from typing import Callable, Optional

import torch
import torch.distributed as dist

from . import _cabi
from .gallery import DeviceGallery


def shard_bounds(n_rows: int, world: int, align: int = 128) -> list[tuple[int, int]]:
    """Contiguous [lo, hi) row blocks, one per rank, block starts aligned to the 128-row tile
    (global index = lo + local index).  Earlier ranks hold lower indices, so "lower rank first"
    agrees with the index-ascending tie rule."""
    if world < 1:
        raise ValueError("world must be >= 1")
    per = math.ceil(n_rows / world / align) * align if n_rows > 0 else 0
    out = []
    for r in range(world):
        lo = min(n_rows, r * per)
        hi = min(n_rows, lo + per)
        out.append((lo, hi))
    return out


def triangle_bounds(n_rows: int, world: int, align: int = 128) -> list[tuple[int, int]]:
    """Row ranges of the i-side of the upper-triangular self-join with equal PAIR counts:
    rows [lo, hi) own pairs (i, j > i), i.e. area ~ integral of (n - i); boundary r solves
    1 - (1 - x)^2 = r / world."""
    bounds = [0]
    for r in range(1, world):
        x = 1.0 - math.sqrt(1.0 - r / world)
        b = int(round(x * n_rows / align)) * align
        bounds.append(min(max(b, bounds[-1]), n_rows))
    bounds.append(n_rows)
    return [(bounds[i], bounds[i + 1]) for i in range(world)]


def _all_gather(t: torch.Tensor, world: int, group) -> torch.Tensor:
    """[world, *t.shape] gathered copy of `t` (same shape on every rank); one collective."""
    t = t.contiguous()
    out = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
    if t.numel():
        dist.all_gather_into_tensor(out.view(-1), t.view(-1), group=group)
    return out


def _merge_cuda(values: torch.Tensor, indices: torch.Tensor, k: int):
    """values/indices [G, Q, k_in] on one device -> merged (values [Q, k], indices [Q, k])."""
    g, q, kin = values.shape
    dev = values.device
    lib = _cabi.lib
    out_v = torch.empty((q, k), dtype=torch.float32, device=dev)
    out_i = torch.empty((q, k), dtype=torch.int64, device=dev)
    if q == 0:
        return out_v, out_i
    with torch.cuda.device(dev):
        ws_bytes = lib.mmrs_topk_merge_workspace_bytes(g, q, kin)
        ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device=dev)
        _cabi.check(lib.mmrs_topk_merge(values.contiguous().data_ptr(), indices.contiguous().data_ptr(),
                                        g, q, kin, k, out_v.data_ptr(), out_i.data_ptr(),
                                        DeviceGallery.aligned_ptr(ws), ws_bytes,
                                        int(torch.cuda.current_stream(dev).cuda_stream)))
    return out_v, out_i


class _PendingSharded:
    def __init__(self, finish, general):
        self._finish, self._general, self._out = finish, general, None

    def wait(self):
        if self._out is None:
            out = self._finish()
            self._out = out if out is not None else self._general()
        return self._out


class ShardedGallery:
    """This rank's shard of a row-sharded gallery plus the merge step.

    `local` is the rank's DeviceGallery with `row_offset` = first global row.  `local_search` and
    `merge` exist so the host-side protocol (gather layout, index offsets, tie rule) can be
    exercised on CPU with the gloo backend in tests; the product defaults are the CUDA kernels.
    """

    def __init__(self, local, n_rows_global: int, group: Optional[dist.ProcessGroup] = None,
                 local_search: Optional[Callable] = None, merge: Optional[Callable] = None,
                 fused: Optional[bool] = None):
        self.local = local
        self.n_rows_global = int(n_rows_global)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._fast = local_search is None and merge is None   # product path: packed-key pipeline
        self._fused = {}
        import os
        self._fused_ok = os.environ.get("MMRS_NO_FUSED_GATHER", "0") != "1" if fused is None else bool(fused)
        if local_search is None:
            from .search import search_topk as local_search
        self._search = local_search
        self._merge = merge or _merge_cuda

    @property
    def fused_active(self) -> bool:
        """True once a search of this handle has gone through the fused NVLink gather."""
        return bool(self._fused) and self._fused_ok

    @classmethod
    def from_full(cls, features: torch.Tensor, mode: Optional[str] = None, device=None,
                  group: Optional[dist.ProcessGroup] = None) -> "ShardedGallery":
        """Every rank is handed the full host matrix and keeps only its block."""
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        lo, hi = shard_bounds(features.shape[0], world)[rank]
        local = DeviceGallery(features[lo:hi], mode=mode, device=device, row_offset=lo)
        return cls(local, features.shape[0], group)

    def _search_topk_keys(self, queries, k: int, normalize_queries: bool, scale: float, path: str):
        """CUDA fast path: local top-k as packed 64-bit keys -> ONE all-gather (Q*k*8 bytes per rank)
        -> merge kernel reading the gathered buffer in place.  Nothing synchronises with the host
        until the final status check."""
        from .search import _prep_queries
        gal = self.local
        q, on_host = _prep_queries(queries, gal)
        if on_host:
            q = q.to(gal.device, non_blocking=True)
        nq = int(q.shape[0])
        dev = gal.device
        lib = _cabi.lib
        if torch.cuda.current_device() != dev.index:
            torch.cuda.set_device(dev)
        k_local = min(k, gal.n_rows)
        n_keys = nq * k
        buf = torch.zeros(n_keys + 1, dtype=torch.int64, device=dev)      # key 0 = "no entry"; [-1] = status
        status = torch.zeros(2, dtype=torch.int32).pin_memory()
        stream = torch.cuda.current_stream(dev)
        ws_ptr, ws_bytes = gal.search_workspace(nq, k_local, False, stream.cuda_stream)
        if nq:
            dst = buf if k_local == k else torch.empty(nq * k_local + 1, dtype=torch.int64, device=dev)
            _cabi.check(lib.mmrs_search_topk_keys_async(
                gal.data.data_ptr(), gal.n_rows, gal.padded_dim, gal.data.stride(0), gal.dtype_code,
                q.data_ptr(), nq, q.stride(0), k_local, int(bool(normalize_queries)), float(scale),
                gal.row_offset, _cabi.PATHS[path], dst.data_ptr(), ws_ptr, ws_bytes, status.data_ptr(),
                stream.cuda_stream))
            if dst is not buf:                       # a shard shorter than k: pad its lists with key 0
                buf[:n_keys].view(nq, k)[:, :k_local] = dst[:-1].view(nq, k_local)
                buf[-1] = dst[-1]
        gathered = _all_gather(buf, self.world, self.group)               # [world, nq*k + 1]
        out_v = torch.empty((nq, k), dtype=torch.float32, device=dev)
        out_i = torch.empty((nq, k), dtype=torch.int64, device=dev)
        if nq:
            d_status = torch.empty(1, dtype=torch.int32, device=dev)
            _cabi.check(lib.mmrs_topk_merge_keys_async(gathered.data_ptr(), self.world, nq, k, n_keys + 1, k,
                                                       out_v.data_ptr(), out_i.data_ptr(), d_status.data_ptr(),
                                                       status[1:].data_ptr(), stream.cuda_stream))
            shard_status = gathered[:, -1].to("cpu", non_blocking=True)
        else:
            shard_status = None
        if on_host:
            hv = torch.empty((nq, k), dtype=torch.float32, pin_memory=True)
            hi = torch.empty((nq, k), dtype=torch.int64, pin_memory=True)
            hv.copy_(out_v, non_blocking=True)
            hi.copy_(out_i, non_blocking=True)
            out_v, out_i = hv, hi
        event = torch.cuda.Event()
        event.record(stream)

        def finish():
            event.synchronize()
            if shard_status is None:
                return out_v, out_i
            worst = int(shard_status.max().item())
            if worst == 1:
                return None                          # some rank overflowed: ALL ranks take the general path
            if worst != 0:                           # e.g. a zero-norm query: same error as single-GPU
                status[0] = worst
                _cabi.check(lib.mmrs_search_status(status.data_ptr()))
            _cabi.check(lib.mmrs_search_status(status[1:].data_ptr()))
            return out_v, out_i

        finish._keepalive = (q, buf, gathered)
        return finish

    # ---- search fused with its all-gather over NVLink peer memory -----------------------------------
    def _fused_state(self, nq: int, k_local: int, stream_id: int):
        """Symmetric (peer-mapped) gather buffer + flag array of one (shape, stream) slot; the
        rendezvous is collective and happens once per slot."""
        key = (nq, k_local, stream_id)
        st = self._fused.get(key)
        if st is None:
            import torch.distributed._symmetric_memory as symm_mem
            dev = self.local.device
            stride = nq * k_local + 1                       # one list + the rank's status word
            total = self.world * stride + self.world         # + 2 * world uint32 flags
            t = symm_mem.empty(total, dtype=torch.int64, device=dev)
            t.zero_()
            hdl = symm_mem.rendezvous(t, self.group if self.group is not None else dist.group.WORLD)
            torch.cuda.synchronize(dev)
            dist.barrier(group=self.group)                   # everyone zeroed before anyone stores
            ptrs = [int(x) for x in hdl.buffer_ptrs]
            st = {"t": t, "hdl": hdl, "stride": stride, "epoch": 0,
                  "bufs": torch.tensor(ptrs, dtype=torch.int64, device=dev),
                  "flags": torch.tensor([x + self.world * stride * 8 for x in ptrs], dtype=torch.int64, device=dev)}
            self._fused[key] = st
        return st

    def _search_topk_fused(self, queries, k: int, normalize_queries: bool, scale: float, path: str):
        """Local scan -> last select stores its keys into EVERY rank's buffer over NVLink and raises
        a ready flag -> merge select waits for all flags.  One library call, no NCCL, no host sync
        until the final status check."""
        from .search import _prep_queries
        gal = self.local
        q, on_host = _prep_queries(queries, gal)
        if on_host:
            q = q.to(gal.device, non_blocking=True)
        nq = int(q.shape[0])
        dev = gal.device
        lib = _cabi.lib
        if torch.cuda.current_device() != dev.index:
            torch.cuda.set_device(dev)
        k_local = min(k, gal.n_rows)
        stream = torch.cuda.current_stream(dev)
        st = self._fused_state(nq, k_local, stream.cuda_stream)
        st["epoch"] += 1
        ws_ptr, ws_bytes = gal.search_workspace(nq, k_local, False, stream.cuda_stream)
        out_v = torch.empty((nq, k), dtype=torch.float32, device=dev)
        out_i = torch.empty((nq, k), dtype=torch.int64, device=dev)
        status = torch.zeros(2 * self.world + 2, dtype=torch.int32).pin_memory()
        _cabi.check(lib.mmrs_search_topk_fused_gather_async(
            gal.data.data_ptr(), gal.n_rows, gal.padded_dim, gal.data.stride(0), gal.dtype_code,
            q.data_ptr(), nq, q.stride(0), k_local, k, int(bool(normalize_queries)), float(scale),
            gal.row_offset, _cabi.PATHS[path], st["bufs"].data_ptr(), st["flags"].data_ptr(),
            st["t"].data_ptr(), self.rank, self.world, st["stride"], st["epoch"],
            out_v.data_ptr(), out_i.data_ptr(), ws_ptr, ws_bytes, status.data_ptr(), stream.cuda_stream))
        if on_host:
            hv = torch.empty((nq, k), dtype=torch.float32, pin_memory=True)
            hi = torch.empty((nq, k), dtype=torch.int64, pin_memory=True)
            hv.copy_(out_v, non_blocking=True)
            hi.copy_(out_i, non_blocking=True)
            out_v, out_i = hv, hi
        event = torch.cuda.Event()
        event.record(stream)

        def finish():
            event.synchronize()
            rc = lib.mmrs_gather_status(status.data_ptr(), self.world)
            if rc == _cabi.ERR_RETRY:
                return None                  # the same verdict on every rank: all take the general path
            _cabi.check(rc)
            return out_v, out_i

        finish._keepalive = (q, status)
        return finish

    def search_topk(self, queries: torch.Tensor, k: int, *, normalize_queries: bool = True,
                    scale: float = 1.0, path: str = "auto", sync: bool = True):
        """Global top-k, identical on every rank.  `queries` must be the same on all ranks.
        `sync=False` returns an object whose `.wait()` yields `(values, indices)`, so that several
        batches (and their all-gathers) can be in flight."""
        if k > self.n_rows_global:
            raise RuntimeError("selected index k out of range")
        if self._fast and self.world > 1:
            nq = int(queries.shape[0]) if hasattr(queries, "shape") and len(queries.shape) == 2 else 1
            fused = self._fused_ok and 1 <= nq <= 1024
            if fused:
                try:
                    finish = self._search_topk_fused(queries, k, normalize_queries, scale, path)
                except (ImportError, AttributeError, RuntimeError) as e:   # no symmetric memory here
                    if isinstance(e, _cabi.MmrsError):
                        raise
                    self._fused_ok = fused = False
            if not fused:
                finish = self._search_topk_keys(queries, k, normalize_queries, scale, path)
            general = lambda: self.search_topk_general(queries, k, normalize_queries=normalize_queries,
                                                       scale=scale, path=path)
            pend = _PendingSharded(finish, general)
            return pend.wait() if sync else pend
        out = self.search_topk_general(queries, k, normalize_queries=normalize_queries, scale=scale, path=path)
        return out if sync else _PendingSharded(lambda: out, None)

    def search_topk_general(self, queries: torch.Tensor, k: int, *, normalize_queries: bool = True,
                            scale: float = 1.0, path: str = "auto"):
        """(values, indices) lists, two all-gathers, merge -- also the path every rank takes together
        when some rank's candidate list overflowed (None from the fast path on EVERY rank alike: the
        shard statuses were gathered with the keys)."""
        n_local = len(self.local)
        k_local = min(k, n_local)
        on_host = False
        if self._fast and self.world > 1 and not (isinstance(queries, torch.Tensor) and queries.is_cuda):
            on_host = True                    # NCCL gathers device tensors: search on the device, copy back
            queries = torch.as_tensor(queries).to(self.local.device)
        v, i = self._search(queries, self.local, k_local, normalize_queries=normalize_queries,
                            scale=scale, path=path)
        if self.world == 1:
            return v, i
        if k_local < k:
            # a shard smaller than k: pad with -inf / sentinel so all ranks gather equal shapes
            pad = k - k_local
            v = torch.cat([v, torch.full((v.shape[0], pad), float("-inf"), dtype=v.dtype, device=v.device)], 1)
            i = torch.cat([i, torch.full((i.shape[0], pad), 2 ** 32 - 1, dtype=i.dtype, device=i.device)], 1)
        gv = _all_gather(v, self.world, self.group)
        gi = _all_gather(i, self.world, self.group)
        mv, mi = self._merge(gv, gi, k)
        return (mv.cpu(), mi.cpu()) if on_host else (mv, mi)

    def find_duplicate_pairs(self, emb_full: torch.Tensor, threshold: float, raw_join: Optional[Callable] = None):
        """Self-join with the gallery replicated (10M x 512 fp32 = 20 GB fits every GPU) and the
        upper-triangular tile grid split by equal pair count; variable-length pair lists are
        gathered (counts first, then padded buffers) and sorted lexicographically."""
        from .dedup import selfjoin_tc_raw, sort_pairs
        n = int(emb_full.shape[0])
        if raw_join is None and emb_full.is_cuda and n >= 2048:
            # tensor-core path: column panels dealt round-robin to the ranks
            pairs = selfjoin_tc_raw(emb_full, threshold, self.rank, self.world)
        else:
            from .dedup import selfjoin_raw
            raw_join = raw_join or selfjoin_raw
            lo, hi = triangle_bounds(n, self.world)[self.rank]
            pairs = raw_join(emb_full, threshold, lo, hi) if hi > lo else emb_full.new_empty((0, 2), dtype=torch.int64)
        if self.world == 1:
            return sort_pairs(pairs, n)
        cnt = torch.tensor([pairs.shape[0]], dtype=torch.int64, device=pairs.device)
        cnts = _all_gather(cnt, self.world, self.group).view(-1)
        m = int(cnts.max().item())
        buf = torch.zeros((max(m, 1), 2), dtype=torch.int64, device=pairs.device)
        buf[:pairs.shape[0]] = pairs
        allbuf = _all_gather(buf, self.world, self.group)
        parts = [allbuf[r, :int(cnts[r].item())] for r in range(self.world)]
        return sort_pairs(torch.cat(parts, dim=0), n)
