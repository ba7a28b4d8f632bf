#!/usr/bin/env python
"""bench.py — coupled columns/s of the per-column coupling step, and the slab-reduce roofline.

  python bench.py --gpus N --steps K --warmup W             (N>1: launched under torchrun)
  python bench.py --impl reference --gpus N --steps K --warmup W

One "step" = one pass of the hot path over every superparameterized column this job owns:
K2 gcm_to_les (forcings) -> K1 slab_reduce (slab means + cloud mask over the LES volumes) ->
K3 les_to_gcm (tendencies) [-> NCCL all_gather of the packed tendencies when N>1].
Workload (BASELINE.json configs[2], the one the north-star target is quoted on): 2048 columns
per GPU, LES 64x64x160, GCM L91, float32 storage, float64 arithmetic. Weak scaling: every rank owns
2048 columns (global columns = 2048*N); synthetic, seeded inputs.

`value`        device-resident throughput (inputs in HBM when the clock starts), CUDA events, max over ranks
`e2e`          the same step through the host-facing pipeline: GCM profiles H2D from pinned host memory
               every step, tendencies D2H to the GCM-owning rank every step (LES volumes are the
               GPU-resident LES state; see DESIGN.md "Measurement")
`roofline`     K1 slab_reduce: algorithmic bytes / CUDA-event time vs MEASURED_PEAKS.json hbm_gbs
`cpu_baseline` the numpy oracle (oracle/numpy_batched.py, kind "port") on a bounded sample, 1 core
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

CONFIGS = {
    # name: (ncol per GPU, nx, ny, nk, nlev, dtype)
    "c2": (128, 64, 64, 160, 91, "f64"),
    "c3": (2048, 64, 64, 160, 91, "f32"),
    "c4": (512, 256, 256, 160, 91, "f32"),
    "c5": (2048, 32, 32, 160, 137, "f32"),
}
METRIC = "coupled columns/s"
DT, F_LES, F_GCM = 900.0, 1.0, 1.0


def workload_name(cfg, ncol):
    _, nx, ny, nk, nlev, dt = CONFIGS[cfg]
    return "%s: %d SP columns/GPU, LES %dx%dx%d, GCM L%d, %s" % (cfg.upper(), ncol, nx, ny, nk, nlev, dt)


def alg_bytes_per_column(nx, ny, nk, esize):
    return 5 * nx * ny * nk * esize


# ------------------------------------------------------------------------------------ clocks
class ClockSampler(object):
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); pw.append(float(parts[2]))
            except ValueError:
                continue
            for nme, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                       samples=len(sm), power_w_max=float(max(pw)))
        return out


# ------------------------------------------------------------------------------------ CPU legs
def _cpu_sample_inputs(ncols, nx, ny, nk, nlev, np_dtype, seed, col0=0):
    from sp_coupler_b200 import synth
    zf, zh = synth.les_grid(nk)
    gcm = synth.make_gcm_columns(ncols, nlev, seed=seed, dtype=np_dtype, col0=col0)
    aux = synth.make_les_aux(ncols, nk, seed=seed, dtype=np_dtype, col0=col0)
    vols = synth.make_les_volumes(gcm, zf, nx, ny, seed=seed, dtype=np_dtype, col0=col0)
    return zf, zh, gcm, aux, vols


def _cpu_pass(inp, layout=0):
    """The reference's per-step work for a block of columns, as the numpy port does it: slab means
    + cloud count (DALES side) and set_les_forcings / set_gcm_tendencies (spcpl.py)."""
    from oracle import numpy_batched as nb
    zf, zh, gcm, aux, vols = inp
    return nb.coupling_step(gcm, zf, zh, vols, aux, aux["PS"], DT, F_LES, F_GCM, True, 0.0, layout, accumulate="native")


def _worker_init(cfg, ncols, seed):
    global _W_INP
    ncol, nx, ny, nk, nlev, dt = CONFIGS[cfg]
    _W_INP = _cpu_sample_inputs(ncols, nx, ny, nk, nlev, np.float32 if dt == "f32" else np.float64,
                                seed, col0=(os.getpid() % 1000) * ncols)
    return True


def _worker_ready(_):
    return os.getpid()


def _worker_pass(reps):
    t = time.perf_counter()
    for _ in range(reps):
        _cpu_pass(_W_INP)
    return time.perf_counter() - t


def run_reference(args):
    """--impl reference: the CPU implementation of the path (the numpy port of the reference; the
    reference itself is pure Python and cannot travel to the GPU box) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    ncol, nx, ny, nk, nlev, dt = CONFIGS[args.config]
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    procs = max(1, min(cores, args.ref_procs or cores))
    cols_w, reps = args.ref_cols, args.ref_reps
    ctx = mp.get_context("fork")
    with ctx.Pool(procs, initializer=_worker_init, initargs=(args.config, cols_w, 42)) as pool:
        pool.map(_worker_ready, range(procs * 4))
        times = []
        for s in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            pool.map(_worker_pass, [reps] * procs, chunksize=1)
            t = time.perf_counter() - t0
            if s >= args.warmup:
                times.append(t)
    t_step = float(np.mean(times))
    cols = procs * cols_w * reps
    val = cols / t_step
    sample = "%d processes x %d distinct columns x %d passes per step (%d column-steps/step)" % (procs, cols_w, reps, cols)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "columns/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.config, ncol), "sample": sample},
            "cpu_baseline": {"value": val, "unit": "columns/s", "cores": procs, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "columns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def cpu_baseline(args, vols_dev, gcm_host, aux_host, zf, zh, budget_s=12.0, ncols=32):
    """Single-core numpy port on a bounded sample of THIS run's inputs (first `ncols` columns)."""
    ncols = min(ncols, vols_dev[0].shape[0])
    from sp_coupler_b200.constants import LES_FIELDS
    vols = {f: v[:ncols].cpu().numpy() for f, v in zip(LES_FIELDS, vols_dev)}
    inp = (zf, zh, {k: v[:ncols] for k, v in gcm_host.items()}, {k: v[:ncols] for k, v in aux_host.items()}, vols)
    lay = 1 if args.layout == "ijk" else 0
    _cpu_pass(inp, lay)
    t0 = time.perf_counter()
    passes = 0
    while True:
        _cpu_pass(inp, lay)
        passes += 1
        el = time.perf_counter() - t0
        if el >= budget_s or passes >= 200:
            break
    return {"value": ncols * passes / el, "unit": "columns/s", "cores": 1, "kind": "port",
            "sample": "%d columns x %d passes of oracle/numpy_batched.coupling_step in %.1f s on 1 core "
                      "(host has %d)" % (ncols, passes, el, os.cpu_count() or 1)}


# ------------------------------------------------------------------------------------ GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from sp_coupler_b200 import synth
    from sp_coupler_b200.coupler import Coupler
    from sp_coupler_b200.pipeline import CouplingPipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus %d must be launched with torch.distributed.run --nproc-per-node %d"
                             % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        if args.bind:
            from sp_coupler_b200.pipeline import bind_host_thread_to_gpu
            cpus = bind_host_thread_to_gpu(dev)     # host pages and PCIe copies on the GPU's NUMA node
            sys.stderr.write("rank %d: %s\n" % (rank, "bound to %d CPUs local to GPU %d (%d-%d)" % (len(cpus), local, cpus[0], cpus[-1])
                                                 if cpus else "CPU binding unavailable"))

    ncol, nx, ny, nk, nlev, dts = CONFIGS[args.config]
    if args.ncol:
        ncol = args.ncol
    tdt = torch.float32 if dts == "f32" else torch.float64
    ndt = np.float32 if dts == "f32" else np.float64
    esize = 4 if dts == "f32" else 8
    ncol_total = ncol * world
    col0 = rank * ncol

    cpl = Coupler(dev)
    zf, zh = synth.les_grid(nk)
    gcm_host = synth.make_gcm_columns(ncol, nlev, seed=42 + 2, dtype=ndt, col0=col0, ncol_total=ncol_total)
    aux_host = synth.make_les_aux(ncol, nk, seed=42 + 2, dtype=ndt, col0=col0, ncol_total=ncol_total)
    gather = args.gather
    if world > 1 and gather != "nccl":
        # the fused gather needs NVLink symmetric memory; agree across ranks, else use the NCCL collective
        try:
            pipe = CouplingPipeline(cpl, zf, zh, ncol, nlev, tdt, couple_surface=True, gather=gather)
            ok = 1
        except Exception as e:          # noqa: BLE001
            sys.stderr.write("rank %d: symmetric-memory gather unavailable (%s); using NCCL all_gather\n" % (rank, e))
            ok = 0
        flag = torch.tensor([ok], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if not int(flag.item()):
            gather = "nccl"
    if world == 1 or gather == "nccl":
        pipe = CouplingPipeline(cpl, zf, zh, ncol, nlev, tdt, couple_surface=True, gather=gather)
    pipe.staging.fill_host(gcm_host)
    pipe.staging.upload()
    vols = synth.device_les_volumes(cpl, gcm_host, zf, nx, ny, seed=42 + 2, dtype=tdt, col0=col0)
    if args.layout == "ijk":     # the (itot, jtot, ktot) C-order view OMUSE hands to Python: k fastest
        for i in range(len(vols)):
            vols[i] = vols[i].permute(0, 3, 2, 1).contiguous()
        pipe.layout = "ijk"
    aux = {k: torch.from_numpy(v).to(dev) for k, v in aux_host.items()}
    pipe.attach_les(vols, aux)
    pipe.les_profiles()                      # first-step slab means (spcpl.py:302-308)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps

    sampler = ClockSampler(local) if rank == 0 else None
    # ---- device-resident leg (value + roofline) ----
    l0 = cpl.launches
    pipe.k1_events = None
    for _ in range(args.warmup):
        pipe.step_device(DT, F_LES, F_GCM)
    barrier()
    pipe.k1_events = []
    l0 = cpl.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        pipe.step_device(DT, F_LES, F_GCM)
    e1.record()
    barrier()
    launches = cpl.launches - l0          # our own kernels only (K2, K1, K3 per step); NCCL / barrier kernels not counted
    ms_t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_step = float(ms_t.item()) / args.steps
    k1_ms = float(np.mean([a.elapsed_time(b) for a, b in pipe.k1_events]))
    pipe.k1_events = None
    # ---- end-to-end leg: host GCM buffers in, host tendencies out, every step ----
    e2e_note = "GCM profiles H2D (pinned) + tendencies D2H every step; LES volumes are device-resident LES state"
    e2e_h2d = pipe.staging.nbytes
    e2e_d2h = pipe.tend_host.numel() * pipe.tend_host.element_size()
    exch = None
    if world > 1 and args.host_exchange:
        # the GCM lives in host memory of rank 0: every rank moves ITS columns' profiles / tendencies through one
        # pinned host buffer shared by all ranks (all PCIe links at once), no device gather (pipeline.HostExchange)
        from sp_coupler_b200.pipeline import HostExchange
        try:
            epipe = CouplingPipeline(cpl, zf, zh, ncol, nlev, tdt, couple_surface=True, gather=False, layout=args.layout)
            epipe.attach_les(vols, aux)
            epipe.les_profiles()
            ok = 1
        except Exception as e:          # noqa: BLE001
            sys.stderr.write("rank %d: e2e pipeline failed (%s)\n" % (rank, e))
            ok = 0
        flag = torch.tensor([ok], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()):
            try:
                exch = HostExchange(epipe.staging, world, rank, owner=0, tag="bench")
                ok = 1
            except Exception as e:      # noqa: BLE001
                sys.stderr.write("rank %d: shared pinned host buffer unavailable (%s)\n" % (rank, e))
                ok = 0
            flag = torch.tensor([ok], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if not int(flag.item()):
                exch = None
    if exch is not None:
        if rank == 0:      # the host GCM's profiles of ALL columns (identical to what each rank generated for itself)
            exch.fill_inputs(synth.make_gcm_columns(ncol_total, nlev, seed=42 + 2, dtype=ndt, col0=0, ncol_total=ncol_total))
        ms_e2e = timed(lambda: exch.step(epipe, DT, F_LES, F_GCM), args.steps, args.warmup)
        e2e_h2d, e2e_d2h = world * epipe.staging.nbytes, world * epipe.tend.numel() * epipe.tend.element_size()
        e2e_note = ("GCM profiles in / tendencies out of ONE pinned host buffer shared by all ranks (the host GCM's memory); "
                    "every rank copies its own columns over its own PCIe link, no device gather; bytes are totals over ranks; "
                    "LES volumes are device-resident LES state")
        if rank == 0:      # same numbers as the device-gathered block of the `value` leg
            torch.cuda.synchronize()
            same = torch.equal(exch.out, pipe.tend_all.cpu())
            e2e_note += "; tendencies identical to the device-gathered block: %s" % same
    else:
        ms_e2e = timed(lambda: pipe.step_host(DT, F_LES, F_GCM), args.steps, args.warmup)
    # single GPU, for the record: the same host-to-host step when only the levels that can be non-zero travel back
    e2e_compact = None
    if world == 1:
        ms_c = timed(lambda: pipe.step_host(DT, F_LES, F_GCM, compact=True), args.steps, args.warmup)
        first = pipe.first_live_level()
        e2e_compact = {"value": ncol_total / (ms_c * 1e-3), "unit": "columns/s", "ms_per_step": ms_c,
                       "h2d_bytes_per_step": pipe.staging.nbytes,
                       "d2h_bytes_per_step": ncol * 7 * (nlev - first) * pipe.tend.element_size(),
                       "note": "tendencies of GCM levels %d..%d only (levels above the LES top are exactly zero, known to the host "
                               "from its own heights); same values as e2e" % (first, nlev - 1)}
    clocks = sampler.stop() if sampler else None
    # ---- for the record: the same step if the LES volumes lived in HOST memory (they do not) ----
    host_vol = None
    if world == 1 and args.host_volume_cols > 0:
        hc = min(args.host_volume_cols, ncol)
        hpipe = CouplingPipeline(cpl, zf, zh, hc, nlev, tdt, couple_surface=True, layout=args.layout)
        hpipe.staging.fill_host({k: v[:hc] for k, v in gcm_host.items()})
        hvols_host = [torch.empty((hc,) + tuple(v.shape[1:]), dtype=tdt, pin_memory=True) for v in vols]
        for hv, v in zip(hvols_host, vols):
            hv.copy_(v[:hc])
        hvols_dev = [torch.empty_like(v[:hc]) for v in vols]
        hpipe.attach_les(hvols_dev, {k: v[:hc].contiguous() for k, v in aux.items()})
        hpipe.les_profiles()

        def host_volume_step():
            for dv, hv in zip(hvols_dev, hvols_host):
                dv.copy_(hv, non_blocking=True)
            hpipe.step_host(DT, F_LES, F_GCM)

        ms_hv = timed(host_volume_step, max(3, args.steps // 4), 2)
        host_vol = {"value": hc / (ms_hv * 1e-3), "unit": "columns/s", "columns": hc, "ms_per_step": ms_hv,
                    "h2d_bytes_per_step": int(sum(h.numel() * h.element_size() for h in hvols_host)) + hpipe.staging.nbytes,
                    "note": "NOT the design point: LES volumes copied from pinned host memory every step (PCIe-bound)"}
        del hvols_host, hvols_dev, hpipe

    # ---- optional: the same device step replayed from a CUDA graph (one launch per step) ----
    graph_leg = None
    if args.graph and world == 1:
        pipe.capture(DT, F_LES, F_GCM)
        ms_g = timed(pipe.step_graph, args.steps, args.warmup)
        graph_leg = {"value": ncol_total / (ms_g * 1e-3), "unit": "columns/s", "ms_per_step": ms_g,
                     "note": "device-resident step replayed from one CUDA graph (K2, K1, projection, K3 captured once)"}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    bpc = alg_bytes_per_column(nx, ny, nk, esize)
    achieved = bpc * ncol / (k1_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            if tj.get("config") == args.config and tj.get("ncol") == ncol:
                traffic = (tj if args.layout == "kji" else tj.get("ijk", {})).get("dram_bytes_per_launch")
        except Exception:
            pass
    line = {
        "metric": METRIC, "value": ncol_total / (ms_step * 1e-3), "unit": "columns/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.config, ncol), "ncol_total": ncol_total,
                   "storage_dtype": dts, "arithmetic": "f64", "layout": args.layout, "l2": "inputs larger than L2 (%.1f GB of LES volumes per GPU "
                   "streamed once per step; no reuse between steps)" % (bpc * ncol / 1e9),
                   "parallelism": ("columns sharded x%d, tendencies %s" % (world, "all_gather (NCCL)" if gather == "nccl" else "gathered by K3 itself (%s): NVLink peer stores into symmetric memory + device barrier" % gather)) if world > 1 else "1 GPU",
                   "step": "K2 gcm_to_les -> K1 slab_reduce -> K3 les_to_gcm"},
        "roofline": {"kernel": "slab_reduce_tma_kernel" if args.layout == "kji" else "slab_reduce_ijk_tma_kernel", "bound": "hbm", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "frac_of_nominal_8TBs": achieved / 8000.0, "k1_ms": k1_ms, "k1_share_of_step": k1_ms / ms_step,
                     "alg_bytes_per_launch": bpc * ncol},
        "e2e": {"value": ncol_total / (ms_e2e * 1e-3), "unit": "columns/s", "h2d_bytes_per_step": e2e_h2d,
                "d2h_bytes_per_step": e2e_d2h, "ms_per_step": ms_e2e, "note": e2e_note},
        "e2e_compact": e2e_compact,
        "e2e_host_volumes": host_vol,
        "cuda_graph": graph_leg,
        "gpu_launches": launches,
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(args, vols, gcm_host, aux_host, zf, zh, budget_s=args.cpu_budget)
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--ncol", type=int, default=0, help="override columns per GPU")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    ap.add_argument("--host-volume-cols", type=int, default=64,
                    help="columns of the extra 'volumes in host memory' measurement (0 = skip)")
    ap.add_argument("--ref-procs", type=int, default=0)
    ap.add_argument("--ref-cols", type=int, default=4, help="distinct columns per reference worker")
    ap.add_argument("--ref-reps", type=int, default=8, help="passes over them per step")
    ap.add_argument("--layout", default="kji", choices=["kji", "ijk"],
                    help="memory order of the LES volumes: kji = [ncol][nk][ny][nx] (DALES), ijk = [ncol][nx][ny][nk] (OMUSE view)")
    ap.add_argument("--no-bind", dest="bind", action="store_false", help="N>1: do not bind rank processes to their GPU's CPUs")
    ap.add_argument("--no-host-exchange", dest="host_exchange", action="store_false",
                    help="N>1: time e2e through the owner GPU (device gather + one D2H) instead of the shared pinned host buffer")
    ap.add_argument("--no-graph", dest="graph", action="store_false",
                    help="skip the extra leg that times the device step replayed from a CUDA graph")
    ap.add_argument("--gather", default="p2p", choices=["nccl", "p2p", "p2p-owner"],
                    help="multi-GPU tendency gather: NCCL all_gather, or fused into K3 (NVLink peer stores)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
