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

import numpy as np
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
    `fused=False` forces the NCCL all-gather variant (also: MMRS_NO_FUSED_GATHER=1).
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
        self._fused_used = False
        if local_search is None:
            from .search import search_topk as local_search
        self._search = local_search
        self._merge = merge or _merge_cuda
        # The fused gather needs every rank to contribute exactly k keys per query (one buffer stride and one
        # merge shape for all ranks; a rank that contributed fewer than min(k, its rows) could drop rows of the
        # global top-k), so it is used only when the SHORTEST shard has at least k rows; otherwise the NCCL
        # variant pads short lists with key 0.  One small collective at construction.
        self.min_shard_rows = len(local)
        if self.world > 1:
            sizes = [None] * self.world
            dist.all_gather_object(sizes, len(local), group=group)
            self.min_shard_rows = int(min(sizes))

    @property
    def fused_active(self) -> bool:
        """True once a search of this handle has gone through the fused NVLink gather."""
        return self._fused_used and self._fused_ok

    @classmethod
    def from_full(cls, features: torch.Tensor, mode: Optional[str] = None, device=None,
                  group: Optional[dist.ProcessGroup] = None, **kw) -> "ShardedGallery":
        """Every rank is handed the full host matrix and keeps only its block."""
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        lo, hi = shard_bounds(features.shape[0], world)[rank]
        local = DeviceGallery(features[lo:hi], mode=mode, device=device, row_offset=lo)
        return cls(local, features.shape[0], group, **kw)

    # ---- per-slot buffers ---------------------------------------------------------------------------------
    def _stage_queries(self, slot, queries):
        """fp32 [Q, padded_dim] queries on the device, in a buffer the slot owns (so the library's graph
        of this slot keeps its query pointer); host queries are copied asynchronously."""
        from .search import _prep_queries
        gal = self.local
        q, on_host = _prep_queries(queries, gal)
        if not on_host:
            return q, False, None
        qd = slot.extra.get("q_dev")
        if qd is None or qd.shape != q.shape:
            qd = slot.extra["q_dev"] = torch.empty(q.shape, dtype=torch.float32, device=gal.device)
        if not q.is_pinned():
            q = q.pin_memory()
        qd.copy_(q, non_blocking=True)
        return qd, True, q

    @staticmethod
    def _results(slot, nq, k, dev, on_host, out):
        """(device result tensors the kernels write, host tensors handed back or None)."""
        if on_host:
            dv = slot.extra.get("out_v")
            if dv is None or dv.shape != (nq, k):
                dv = slot.extra["out_v"] = torch.empty((nq, k), dtype=torch.float32, device=dev)
                slot.extra["out_i"] = torch.empty((nq, k), dtype=torch.int64, device=dev)
            di = slot.extra["out_i"]
            if out is not None:
                hv, hi = out
            else:
                hv = torch.empty((nq, k), dtype=torch.float32, pin_memory=True)
                hi = torch.empty((nq, k), dtype=torch.int64, pin_memory=True)
            return dv, di, hv, hi
        if out is not None:
            return out[0], out[1], None, None
        return (torch.empty((nq, k), dtype=torch.float32, device=dev),
                torch.empty((nq, k), dtype=torch.int64, device=dev), None, None)

    def _search_topk_keys(self, queries, k: int, normalize_queries: bool, scale: float, path: str, out=None):
        """CUDA fast path: local top-k as packed 64-bit keys -> ONE all-gather (Q*k*8 bytes per rank)
        -> merge kernel reading the gathered buffer in place.  Nothing synchronises with the host
        until the final status check.  Shards shorter than k pad their lists with key 0."""
        gal = self.local
        dev = gal.device
        lib = _cabi.lib
        if torch.cuda.current_device() != dev.index:
            torch.cuda.set_device(dev)
        stream = torch.cuda.current_stream(dev)
        nq = int(queries.shape[0]) if hasattr(queries, "shape") and len(queries.shape) == 2 else 1
        k_local = min(k, gal.n_rows)
        slot = gal.search_slot(nq, k_local, False, stream.cuda_stream)
        q, on_host, keep = self._stage_queries(slot, queries)
        n_keys = nq * k
        bufs = slot.extra.get(("keys", k))
        if bufs is None:
            bufs = slot.extra[("keys", k)] = (
                torch.zeros(n_keys + 1, dtype=torch.int64, device=dev),                # key 0 = "no entry"; [-1] = status
                torch.zeros(nq * k_local + 1, dtype=torch.int64, device=dev) if k_local != k else None,
                torch.empty((self.world, n_keys + 1), dtype=torch.int64, device=dev),
                torch.empty(1, dtype=torch.int32, device=dev))
        buf, short, gathered, d_status = bufs
        dv, di, hv, hi = self._results(slot, nq, k, dev, on_host, out)
        ticket, status, event = slot.acquire()
        shard_status = None
        if nq:
            dst = buf if short is None else short
            st = lib.mmrs_search_topk_keys_async(
                gal.data.data_ptr(), gal.n_rows, gal.padded_dim, gal.data.stride(0), gal.dtype_code,
                q.data_ptr(), nq, q.stride(0), k_local, int(bool(normalize_queries)), float(scale),
                gal.row_offset, _cabi.PATHS[path], dst.data_ptr(), slot.ptr, slot.nbytes, status.data_ptr(),
                stream.cuda_stream)
            if st != _cabi.OK:
                slot.release(ticket)
                _cabi.check(st)
            if short is not None:                       # a shard shorter than k: the rest of its lists stays key 0
                buf[:n_keys].view(nq, k)[:, :k_local] = short[:-1].view(nq, k_local)
                buf[-1] = short[-1]
        dist.all_gather_into_tensor(gathered.view(-1), buf, group=self.group)      # [world, nq*k + 1]
        if nq:
            _cabi.check(lib.mmrs_topk_merge_keys_async(gathered.data_ptr(), self.world, nq, k, n_keys + 1, k,
                                                       dv.data_ptr(), di.data_ptr(), d_status.data_ptr(),
                                                       status[1:].data_ptr(), stream.cuda_stream))
            table = slot.extra.get("shard_status")          # one pinned row per status ticket of the slot
            if table is None or table.shape[0] <= ticket:
                table = slot.extra["shard_status"] = torch.zeros((max(ticket + 1, slot.N_STATUS), self.world),
                                                                 dtype=torch.int64).pin_memory()
            shard_status = table[ticket]
            shard_status.copy_(gathered[:, -1], non_blocking=True)
        if on_host:
            hv.copy_(dv, non_blocking=True)
            hi.copy_(di, non_blocking=True)
        event.record(stream)

        def finish():
            event.synchronize()
            res = (hv, hi) if on_host else (dv, di)
            if shard_status is None:
                slot.release(ticket)
                return res
            worst = int(shard_status.max().item())
            merge_rc = lib.mmrs_search_status(status[1:].data_ptr())
            if worst not in (0, 1):                  # e.g. a zero-norm query: same error as single-GPU
                status[0] = worst
            own_rc = lib.mmrs_search_status(status.data_ptr()) if worst not in (0, 1) else _cabi.OK
            slot.release(ticket)
            if worst == 1:
                return None                          # some rank overflowed: ALL ranks take the general path
            _cabi.check(own_rc)
            _cabi.check(merge_rc)
            return res

        finish._keepalive = (q, keep)
        return finish

    # ---- search fused with its all-gather over NVLink peer memory -----------------------------------
    def _fused_state(self, slot, nq: int, k_local: int):
        """Symmetric (peer-mapped) gather buffer + flag array of one (shape, stream) slot; the
        rendezvous is collective and happens once per slot."""
        st = slot.extra.get("fused")
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
            flag_off = self.world * stride * 8
            st = {"t": t, "hdl": hdl, "stride": stride,
                  "bufs": torch.tensor(ptrs, dtype=torch.int64, device=dev),
                  "flags": torch.tensor([x + flag_off for x in ptrs], dtype=torch.int64, device=dev),
                  "local_flags": t.data_ptr() + flag_off}
            slot.extra["fused"] = st
        return st

    def _search_topk_fused(self, queries, k: int, normalize_queries: bool, scale: float, path: str, out=None):
        """Local scan -> last select stores its keys into EVERY rank's buffer over NVLink and raises
        a ready flag -> a one-warp kernel waits for all flags -> merge select.  One library call, one
        graph replay, no NCCL, no host sync until the final status check; every buffer the call
        touches belongs to the (shape, stream) slot and was allocated once."""
        gal = self.local
        dev = gal.device
        lib = _cabi.lib
        if torch.cuda.current_device() != dev.index:
            torch.cuda.set_device(dev)
        stream = torch.cuda.current_stream(dev)
        nq = int(queries.shape[0])
        k_local = k                                      # every shard has >= k rows (checked by search_topk)
        slot = gal.search_slot(nq, k_local, False, stream.cuda_stream)
        st = self._fused_state(slot, nq, k_local)         # collective on first use of the slot
        q, on_host, keep = self._stage_queries(slot, queries)
        dv, di, hv, hi = self._results(slot, nq, k, dev, on_host, out)
        ticket, status, event = slot.acquire()
        rc = lib.mmrs_search_topk_fused_gather_async(
            gal.data.data_ptr(), gal.n_rows, gal.padded_dim, gal.data.stride(0), gal.dtype_code,
            q.data_ptr(), nq, q.stride(0), k_local, k, int(bool(normalize_queries)), float(scale),
            gal.row_offset, _cabi.PATHS[path], st["bufs"].data_ptr(), st["flags"].data_ptr(),
            st["t"].data_ptr(), st["local_flags"], self.rank, self.world, st["stride"],
            dv.data_ptr(), di.data_ptr(), slot.ptr, slot.nbytes, status.data_ptr(), stream.cuda_stream)
        if rc != _cabi.OK:
            slot.release(ticket)
            _cabi.check(rc)
        if on_host:
            hv.copy_(dv, non_blocking=True)
            hi.copy_(di, non_blocking=True)
        event.record(stream)
        self._fused_used = True

        def finish():
            event.synchronize()
            rc = lib.mmrs_gather_status(status.data_ptr(), self.world)
            slot.release(ticket)
            if rc == _cabi.ERR_RETRY:
                return None                  # the same verdict on every rank: all take the general path
            if rc == _cabi.ERR_TIMEOUT:
                self._fused_ok = False       # a peer is dead or diverged: later calls use NCCL (its own abort handling)
            _cabi.check(rc)
            return (hv, hi) if on_host else (dv, di)

        finish._keepalive = (q, keep)
        return finish

    def search_topk(self, queries: torch.Tensor, k: int, *, normalize_queries: bool = True,
                    scale: float = 1.0, path: str = "auto", sync: bool = True, out=None):
        """Global top-k, identical on every rank.  `queries` must be the same on all ranks.
        `sync=False` returns an object whose `.wait()` yields `(values, indices)`, so that several
        batches (and their all-gathers) can be in flight.  `out=(values, indices)`: caller-owned result
        tensors (pinned host tensors for host queries)."""
        if k > self.n_rows_global:
            raise RuntimeError("selected index k out of range")
        if self._fast and self.world > 1:
            if isinstance(queries, np.ndarray):
                queries = torch.from_numpy(queries)
            if queries.dim() == 1:
                queries = queries.unsqueeze(0)
            nq = int(queries.shape[0])
            if nq == 0:
                out_t = self.search_topk_general(queries, k, normalize_queries=normalize_queries, scale=scale, path=path)
                return out_t if sync else _PendingSharded(lambda: out_t, None)
            # fused NVLink gather: every rank contributes its full top-k (needs >= k rows in every shard)
            fused = self._fused_ok and 1 <= nq <= 1024 and self.min_shard_rows >= k
            finish = None
            if fused:
                try:
                    finish = self._search_topk_fused(queries, k, normalize_queries, scale, path, out)
                except (ImportError, AttributeError, RuntimeError) as e:   # no symmetric memory here
                    if isinstance(e, _cabi.MmrsError):
                        raise
                    self._fused_ok = False
            if finish is None:
                finish = self._search_topk_keys(queries, k, normalize_queries, scale, path, out)
            general = lambda: self.search_topk_general(queries, k, normalize_queries=normalize_queries,
                                                       scale=scale, path=path)
            pend = _PendingSharded(finish, general)
            return pend.wait() if sync else pend
        out_t = self.search_topk_general(queries, k, normalize_queries=normalize_queries, scale=scale, path=path)
        return out_t if sync else _PendingSharded(lambda: out_t, None)

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
