#!/usr/bin/env python
"""bench.py -- top-k queries/sec of the B200 search path on BASELINE.json's C2 workload.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
    python bench.py --impl reference ...                     (the reference's torch-CPU path)

Workload (config.workload = "C2"): per GPU a 1 000 000 x 512 bf16 unit-norm synthetic gallery
(BASELINE.json configs[1]), top-100, cosine.  A step is one search of a batch of queries:
  N = 1   16 queries against the 1M-row gallery (HBM-bound: one 1.024 GB gallery stream)
  N > 1   the gallery is row-sharded (1M rows PER GPU, global gallery N x 1M rows), the batch
          grows to 16 x N queries, every rank scans its shard for all of them, one NCCL
          all-gather moves the [Q, k] lists, every rank merges.  Per-GPU HBM bytes per step are
          fixed ("weak"), whole-job value = global queries / time.
`value`  : device-resident queries and results, CUDA-event timed, max over ranks.
`e2e`    : the same search through the public host API (pinned host queries -> H2D -> search ->
           D2H of values/indices), host clock around synchronised calls.
`roofline`: dominant kernel (the last-phase gallery scan) timed live by the library's CUDA-event
           hooks inside the timed region: algorithmic gallery bytes / duration vs the measured HBM peak.
`cpu_baseline`: the oracle (torch CPU restatement of search_image.py:107 + utils.py:17) on the
           host cores, bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

ROWS_PER_GPU = 1_000_000
DIM = 512
TOPK = 100
Q_PER_GPU = 16
L2_BYTES = 126 * 1024 * 1024


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=ROWS_PER_GPU)
    ap.add_argument("--dim", type=int, default=DIM)
    ap.add_argument("--k", type=int, default=TOPK)
    ap.add_argument("--batch", type=int, default=Q_PER_GPU, help="queries per GPU per step")
    ap.add_argument("--global-batch", type=int, default=0, help="total queries per step (overrides --batch x gpus)")
    ap.add_argument("--path", default="auto", choices=["auto", "gemv", "mma"])
    ap.add_argument("--sweep", action="store_true", help="also time batch sizes 1..256 (N=1, extra key)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--one-stream", action="store_true", help="keep the batches in flight on ONE stream")
    return ap.parse_args()


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), "measured"
    return 6650.0, 1590.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def device_gallery_shard(torch, rows, dim, seed, rank, device):
    """randn rows generated on the device in chunks, unit-normalised in fp32, cast to bf16
    (C2/C4 galleries are never materialised on the host)."""
    gen = torch.Generator(device=device).manual_seed(1000 * seed + rank)
    out = torch.empty((rows, dim), dtype=torch.bfloat16, device=device)
    step = 1 << 18
    for lo in range(0, rows, step):
        n = min(step, rows - lo)
        x = torch.randn((n, dim), generator=gen, device=device, dtype=torch.float32)
        out[lo:lo + n] = (x / x.norm(dim=-1, keepdim=True)).to(torch.bfloat16)
    return out


def cpu_reference_rate(torch, rows, dim, nq, k, budget_s=12.0, min_reps=2):
    """The reference's CPU path on this host: scores = q @ G.T (search_image.py:107, after the
    normalisation idiom :157), then output.topk(k, 1, True, True) (utils.py:17); all host threads."""
    from oracle import oracle
    torch.set_num_threads(os.cpu_count() or 1)
    g = oracle.synthetic_gallery(rows, dim, seed=0, dtype=torch.bfloat16).to(torch.float32)
    q = oracle.synthetic_queries(nq, dim, seed=1)

    def step():
        qq = oracle.l2_normalize(q)
        s = qq @ g.t()
        return s.topk(k, 1, True, True)

    step()  # warm-up
    times = []
    t_end = time.perf_counter() + budget_s
    while len(times) < min_reps or (time.perf_counter() < t_end and len(times) < 200):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    best = min(times)
    return nq / best, {"cores": torch.get_num_threads(), "reps": len(times), "best_s": best,
                       "sample": f"{nq} queries x {rows}x{dim} fp32-upcast gallery, top-{k}, best of {len(times)} reps"}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (the oracle port: the
    reference is a set of Python scripts that cannot be installed or imported)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import oracle
    world = args.gpus
    nq = args.global_batch or args.batch * world
    rows = args.rows     # bounded sample: one GPU's shard; the CPU rate is linear in rows
    torch.set_num_threads(os.cpu_count() or 1)
    g = oracle.synthetic_gallery(rows, args.dim, seed=0, dtype=torch.bfloat16).to(torch.float32)
    q = oracle.synthetic_queries(nq, args.dim, seed=1)

    def step():
        s = oracle.l2_normalize(q) @ g.t()
        return s.topk(args.k, 1, True, True)

    for _ in range(min(args.warmup, 2)):
        step()
    steps = max(1, min(args.steps, 20))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    # the global gallery is `world` shards: scale the measured one-shard time linearly in rows
    per_step = dt / steps * world
    value = nq / per_step
    sample = (f"{steps} steps of {nq} queries x one {rows}x{args.dim} shard (fp32 upcast), top-{args.k}; "
              f"time scaled x{world} for the {world * rows}-row global gallery")
    line = {
        "impl": "reference", "metric": "top-k queries/sec", "value": value, "unit": "queries/s", "n_gpus": world,
        "steps": steps, "warmup": min(args.warmup, 2), "ms_per_step": per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    name = "C2: 1M x 512 bf16 gallery per GPU, top-100 cosine" if (args.rows, args.dim) == (ROWS_PER_GPU, DIM) \
        else f"{args.rows} x {args.dim} bf16 gallery per GPU, top-{args.k} cosine"
    return {"workload": name, "rows_per_gpu": args.rows,
            "global_rows": args.rows * world, "dim": args.dim, "k": args.k, "queries_per_step": args.global_batch or args.batch * world,
            "sharding": f"rows x{world}" if world > 1 else "none",
            "l2": f"gallery shard {args.rows * args.dim * 2 / 1e9:.3f} GB streamed per step > 126 MB L2 (no flush needed)"}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist
    import mmrs_b200
    from mmrs_b200 import _cabi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    nq = args.global_batch or args.batch * world
    lib = _cabi.lib
    shard = device_gallery_shard(torch, args.rows, args.dim, seed=0, rank=rank, device=device)
    gal = mmrs_b200.DeviceGallery(shard, row_offset=rank * args.rows)
    sg = mmrs_b200.ShardedGallery(gal, args.rows * world) if world > 1 else None
    q_host = torch.randn((nq, args.dim), generator=torch.Generator().manual_seed(1)).pin_memory()
    q_dev = q_host.to(device)

    DEPTH = 2      # batches in flight: the host prepares step i+1 while the GPU runs step i

    def search_dev(sync=True):
        if sg is not None:
            return sg.search_topk(q_dev, args.k, path=args.path, sync=sync)
        return mmrs_b200.search_topk(q_dev, gal, args.k, path=args.path, sync=sync)

    def search_host(sync=True):
        if sg is not None:     # host queries in, host results out (H2D and D2H inside the call)
            return sg.search_topk(q_host, args.k, path=args.path, sync=sync)
        return mmrs_b200.search_topk(q_host, gal, args.k, path=args.path, sync=sync)

    streams = [torch.cuda.Stream(device=device) for _ in range(DEPTH)] if not args.one_stream else None

    def run_steps(fn, n):
        """n steps with up to DEPTH batches in flight, alternating over DEPTH streams (each with its
        own workspace) so that the short seed/mid kernels of one search overlap the long last-phase
        scan of the other; every batch is waited on and status-checked."""
        inflight = []
        out = None
        for i in range(n):
            if streams is not None:
                with torch.cuda.stream(streams[i % DEPTH]):
                    inflight.append(fn(sync=False))
            else:
                inflight.append(fn(sync=False))
            if len(inflight) >= DEPTH:
                out = inflight.pop(0).wait()
        for pnd in inflight:
            out = pnd.wait()
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident, CUDA events ---------------------------------------------------
    run_steps(search_dev, max(args.warmup, 3))
    barrier()
    launches0 = lib.mmrs_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        ev0.record()
        if streams is not None:
            for st in streams:
                st.wait_stream(torch.cuda.current_stream(device))       # timed region starts at ev0
        out = run_steps(search_dev, args.steps)
        if streams is not None:
            for st in streams:
                torch.cuda.current_stream(device).wait_stream(st)       # ... and ends when both streams drain
        ev1.record()
        barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = lib.mmrs_launch_count() - launches0
    # synchronous per-call latency of the same search (one batch in flight, host blocks on each)
    for _ in range(3):
        search_dev()                    # one-time set-up of the default-stream slot stays untimed
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        search_dev()
    torch.cuda.synchronize()
    sync_call_ms = (time.perf_counter() - t0) / args.steps * 1e3
    # dominant-kernel duration: the same steps again with the library's CUDA-event hooks on (the
    # hooks bracket every launch on the launching stream; while they are on the library issues
    # the launches one by one instead of replaying its CUDA graph -- the kernels are the same)
    lib.mmrs_profile_enable(1)
    for _ in range(args.steps):
        search_dev()
    torch.cuda.synchronize()
    lib.mmrs_profile_enable(0)
    import ctypes as C
    cap = (int(launches) // max(args.steps, 1) + 2) * args.steps + 64
    ms = (C.c_float * cap)(); kind = (C.c_int32 * cap)(); nbytes = (C.c_int64 * cap)(); flops = (C.c_int64 * cap)()
    nrec = lib.mmrs_profile_read(ms, kind, nbytes, flops, cap)
    t = torch.tensor([ms_total], device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = nq * args.steps / (ms_total / 1e3)

    # dominant kernel = the launches that stream the most gallery bytes (last phase of each search)
    recs = [(ms[i], kind[i], nbytes[i], flops[i]) for i in range(nrec) if ms[i] > 0]
    all_recs = recs
    recs = [r for r in all_recs if r[1] in (1, 2)]          # gallery scans only
    per_step = nrec // max(args.steps, 1)
    timeline = None
    if per_step and nrec == per_step * args.steps:            # same launch sequence every step
        names = {1: "scan_gemv", 2: "scan_mma", 3: "select", 4: "prep"}
        timeline = [{"kernel": names.get(kind[j], "?"),
                     "ms": sum(ms[s * per_step + j] for s in range(args.steps)) / args.steps}
                    for j in range(per_step)]
    hbm_peak, tc_peak, peak_src = measured_peaks()
    roofline = None
    if recs:
        big = max(r[2] for r in recs)
        dom = [r for r in recs if r[2] == big]
        avg_ms = sum(r[0] for r in dom) / len(dom)
        achieved = big / (avg_ms / 1e3) / 1e9
        scan_ms_per_step = sum(r[0] for r in recs) / args.steps
        kname = "scan_mma_kernel<filter>" if dom[0][1] == 2 else "scan_gemv_kernel<filter>"
        traffic = None
        tpath = ROOT / "profiles" / "traffic.json"      # dram__bytes_read+write of one ncu --set full capture
        if tpath.exists() and args.rows == ROWS_PER_GPU and args.dim == DIM:
            tj = json.loads(tpath.read_text()).get(kname)
            if tj:
                traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
        per_pass_q = min(nq, 256)
        if per_pass_q >= 224:     # past the ridge (Q* ~ 206-248): tensor-bound
            long_step = ms_total / args.steps > 50.0
            peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
            tpk = peaks.get("bf16_tflops_sustained", 1400.0) if long_step else tc_peak
            tfl = dom[0][3] / (avg_ms / 1e3) / 1e12
            roofline = {"bound": "tensor", "achieved": tfl, "peak": tpk, "unit": "TFLOP/s", "frac": tfl / tpk, "traffic": None,
                        "peak_source": f"{peak_src} (MEASURED_PEAKS.json bf16_tflops{'_sustained' if long_step else ''})",
                        "kernel": kname, "algorithmic_flops_per_launch": dom[0][3], "avg_launch_ms": avg_ms,
                        "launches_timed": len(dom), "hbm_gbs": achieved,
                        "whole_step_tflops": 2.0 * nq * args.rows * args.dim / (ms_total / args.steps / 1e3) / 1e12,
                        "how": "CUDA events around every launch of the kernel, on the launching stream, over the same "
                               f"{args.steps} steps re-run with the library's profiling hooks enabled"}
        else:
          roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "traffic": traffic, "peak_source": f"{peak_src} (MEASURED_PEAKS.json hbm_gbs, burst copy)",
                    "kernel": kname,
                    "algorithmic_bytes_per_launch": big, "avg_launch_ms": avg_ms, "launches_timed": len(dom),
                    "tflops": dom[0][3] / (avg_ms / 1e3) / 1e12,
                    "scan_kernels_ms_per_step": scan_ms_per_step, "share_of_step": scan_ms_per_step / (ms_total / args.steps),
                    "whole_step_frac": (args.rows * args.dim * 2) / (ms_total / args.steps / 1e3) / 1e9 / hbm_peak,
                    "how": "CUDA events around every launch of the kernel, on the launching stream, over the same "
                           f"{args.steps} steps re-run with the library's profiling hooks enabled"}

    # ---- e2e: host API, host buffers ---------------------------------------------------------------
    run_steps(search_host, 3)
    barrier()
    t0 = time.perf_counter()
    hv, hi = run_steps(search_host, args.steps)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], device=device)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    for _ in range(3):
        search_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        search_host()
    torch.cuda.synchronize()
    e2e_sync_ms = (time.perf_counter() - t0) / args.steps * 1e3
    e2e = {"value": nq * args.steps / float(tt.item()), "unit": "queries/s",
           "h2d_bytes_per_step": nq * args.dim * 4, "d2h_bytes_per_step": nq * args.k * 12,
           "ms_per_step": float(tt.item()) / args.steps * 1e3,
           "mode": f"host API search_topk(sync=False), {DEPTH} batches in flight, every batch waited on and status-checked"
                   if sg is None else f"ShardedGallery.search_topk(host queries, sync=False), {DEPTH} batches in flight",
           "blocking_call_ms": e2e_sync_ms}

    sweep = None
    if args.sweep and world == 1:
        sweep = []
        for b in (1, 2, 4, 8, 16, 32, 64, 128, 256):
            qd = torch.randn((b, args.dim), generator=torch.Generator().manual_seed(2)).to(device)
            for _ in range(3):
                mmrs_b200.search_topk(qd, gal, args.k, path=args.path)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                mmrs_b200.search_topk(qd, gal, args.k, path=args.path)
            e1.record()
            torch.cuda.synchronize()
            per = e0.elapsed_time(e1) / 20
            sweep.append({"batch": b, "ms": per, "qps": b / per * 1e3,
                          "hbm_frac": args.rows * args.dim * 2 / (per / 1e3) / 1e9 / hbm_peak,
                          "tflops": 2.0 * b * args.rows * args.dim / (per / 1e3) / 1e12})

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        rate, info = cpu_reference_rate(torch, args.rows, args.dim, nq, args.k)
        cpu = {"value": rate, "unit": "queries/s", "cores": info["cores"], "kind": "port", "sample": info["sample"]}

    if rank == 0:
        line = {
            "metric": "top-k queries/sec", "value": value, "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args, world), "clocks": clocks.summary(), "e2e": e2e,
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "pipelining": f"{DEPTH} batches in flight" + ("" if args.one_stream else f" on {DEPTH} streams"),
            "blocking_call_ms": sync_call_ms,
        }
        if sweep:
            line["sweep"] = sweep
            line["sweep_mode"] = "one blocking call at a time (no batches in flight), device-resident queries"
        if timeline:
            line["kernel_timeline_ms"] = timeline
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
