#!/usr/bin/env python
"""bench.py — coupled columns/s of the per-column coupling step, and the slab-reduce roofline.

  python bench.py --gpus N --steps K --warmup W             (N>1: launched under torchrun)
  python bench.py --impl reference --gpus N --steps K --warmup W

One "step" = one pass of the hot path over every superparameterized column of the job:
K2 gcm_to_les (forcings) -> K1 slab_reduce (slab means + cloud mask over the LES volumes) ->
K3 les_to_gcm (cloud projection + tendencies; its epilogue gathers the packed block on the GCM-owning rank over
NVLink when N>1, or writes it into the host GCM's pinned memory in the e2e leg).

Workload = BASELINE.json configs[2] AS WRITTEN (the one the north-star target is quoted on): 2048 SP columns IN
TOTAL, LES 64x64x160, GCM L91, float32 storage, float64 arithmetic, on N GPUs -> STRONG scaling: every rank owns
2048/N columns. (--scaling weak gives every rank the whole column count; at N>1 the default run also times that as
the `weak` sub-record so the round-1 curve stays comparable.) Synthetic, seeded inputs.

`value`        device-resident throughput (inputs in HBM when the clock starts): the step replayed from its CUDA graph
               (three kernels, no host synchronisation), CUDA events, max over ranks
`roofline`     K1 slab_reduce: algorithmic bytes / CUDA-event time (an eager pass of the same step in the same run)
               vs MEASURED_PEAKS.json hbm_gbs
`e2e`          the same step through the host-facing pipeline: per step the GCM profiles of the live level window go
               H2D from pinned host memory, the step graph is replayed, and the tendencies land in pinned host memory
               (one D2H copy on one GPU; with N>1 every rank's K3 stores its block straight into ONE host buffer shared
               by all ranks and raises a flag) - the LES volumes are the GPU-resident LES state (DESIGN.md
               "Measurement"); `e2e_host_volumes` is the same step with the volumes shipped from the host
`cpu_baseline` the UNMODIFIED reference (oracle/_ref via oracle/ref_driver.py, kind "reference") on 1 core, bounded
               sample; `cpu_baseline_port` = the numpy port (oracle/numpy_batched.py) on 1 core
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import faulthandler

faulthandler.enable()      # a rank that dies on a signal leaves a traceback instead of silence

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(1, os.path.join(ROOT, "tests"))      # synth_les: the synthetic LES volumes (data generation, not product)

import numpy as np  # noqa: E402

CONFIGS = {
    # name: (SP columns in total, nx, ny, nk, nlev, dtype)   BASELINE.json configs[1..4]
    "c2": (128, 64, 64, 160, 91, "f64"),
    "c3": (2048, 64, 64, 160, 91, "f32"),
    "c4": (512, 256, 256, 160, 91, "f32"),
    "c5": (16384, 32, 32, 160, 137, "f32"),
}
REF_COLS = {"c2": 32, "c3": 64, "c4": 8, "c5": 128}     # distinct columns per reference worker (BASELINE.md §2 rule)
METRIC = "coupled columns/s"
DT, F_LES, F_GCM = 900.0, 1.0, 1.0
SEED = 42 + 2


def workload_name(cfg, ncol_total, world, scaling):
    _, nx, ny, nk, nlev, dt = CONFIGS[cfg]
    return "%s: %d SP columns in total (%d per GPU on %d GPU%s, %s scaling), LES %dx%dx%d, GCM L%d, %s" % (
        cfg.upper(), ncol_total, ncol_total // world, world, "" if world == 1 else "s", scaling, nx, ny, nk, nlev, dt)


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
    import synth_les
    from sp_coupler_b200 import synth
    zf, zh = synth.les_grid(nk)
    gcm = synth.make_gcm_columns(ncols, nlev, seed=seed, dtype=np_dtype, col0=col0)
    aux = synth.make_les_aux(ncols, nk, seed=seed, dtype=np_dtype, col0=col0)
    vols = synth_les.make_les_volumes(gcm, zf, nx, ny, seed=seed, dtype=np_dtype, col0=col0)
    return zf, zh, gcm, aux, vols


def _port_pass(inp, layout=0):
    """The path's per-step work for a block of columns as the numpy PORT does it (vectorised over the block)."""
    from oracle import numpy_batched as nb
    zf, zh, gcm, aux, vols = inp
    return nb.coupling_step(gcm, zf, zh, vols, aux, aux["PS"], DT, F_LES, F_GCM, True, 0.0, layout, accumulate="native")


def _reference_pass(inp):
    """The same work as the UNMODIFIED reference does it: a serial loop over the columns (splib.py:317-332), numpy
    slab means + spcpl.set_les_forcings + spcpl.set_gcm_tendencies per column (oracle/ref_driver.py, BASELINE.md B-ref)."""
    from oracle import ref_driver
    zf, zh, gcm, aux, vols = inp
    return ref_driver.reference_column_steps(gcm, zf, zh, vols, aux, DT, F_LES, F_GCM, True, 0.0)


def reference_available():
    try:
        from oracle import ref_driver
        return ref_driver.available()
    except Exception:
        return False


def _worker_init(cfg, ncols, seed):
    global _W_INP
    _, nx, ny, nk, nlev, dt = CONFIGS[cfg]
    _W_INP = _cpu_sample_inputs(ncols, nx, ny, nk, nlev, np.float32 if dt == "f32" else np.float64,
                                seed, col0=(os.getpid() % 1000) * ncols)
    return True


def _worker_ready(_):
    return os.getpid()


def _worker_pass(kind):
    t = time.perf_counter()
    (_reference_pass if kind == "reference" else _port_pass)(_W_INP)
    return time.perf_counter() - t


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on all host cores - the UNMODIFIED
    spcpl functions from oracle/_ref (a verbatim, git-ignored copy of /root/reference/splib made by oracle/make_ref.py)
    driven per column under the unit shim, in a fork pool with one process per core and >= 64 distinct columns per
    process (BASELINE.md §2 sub-sampling rule); falls back to the numpy port (kind "port") only if that copy is
    missing. Rank 0 runs it; other ranks exit 0."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    ncol_total, nx, ny, nk, nlev, dt = CONFIGS[args.config]
    esize = 4 if dt == "f32" else 8
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    procs = max(1, min(cores, args.ref_procs or cores))
    cols_w = args.ref_cols or REF_COLS[args.config]
    try:        # stay well inside the host's memory: volumes + float64 temporaries of the slab means
        avail = int(open("/proc/meminfo").read().split("MemAvailable:")[1].split()[0]) * 1024
        per_col = alg_bytes_per_column(nx, ny, nk, esize) * 2.5
        cols_w = max(4, min(cols_w, int(0.5 * avail / procs / per_col)))
    except Exception:
        pass
    kind = "reference" if reference_available() and not args.ref_port else "port"
    ctx = mp.get_context("fork")
    with ctx.Pool(procs, initializer=_worker_init, initargs=(args.config, cols_w, SEED)) as pool:
        pool.map(_worker_ready, range(procs * 4))

        def timed(kind_, steps, warmup):
            times = []
            for s in range(warmup + steps):
                t0 = time.perf_counter()
                pool.map(_worker_pass, [kind_] * procs, chunksize=1)
                t = time.perf_counter() - t0
                if s >= warmup:
                    times.append(t)
            return float(np.mean(times))

        t_step = timed(kind, args.steps, args.warmup)
        t_port = timed("port", min(args.steps, 3), 1) if kind == "reference" else t_step
    cols = procs * cols_w
    val = cols / t_step
    what = ("unmodified reference spcpl.set_les_forcings + set_gcm_tendencies + numpy slab means, serial per column"
            if kind == "reference" else "numpy port oracle/numpy_batched.coupling_step")
    sample = "%d processes x %d distinct columns x 1 pass per step (%d column-steps/step, scaled linearly); %s" % (
        procs, cols_w, cols, what)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "columns/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.config, ncol_total, max(args.gpus, 1), args.scaling),
                       "ncol_total": ncol_total, "sample": sample},
            "cpu_baseline": {"value": val, "unit": "columns/s", "cores": procs, "kind": kind, "sample": sample},
            "cpu_baseline_port": {"value": cols / t_port, "unit": "columns/s", "cores": procs, "kind": "port",
                                  "sample": "same pool and columns, oracle/numpy_batched.coupling_step (vectorised numpy restatement)"},
            "e2e": {"value": val, "unit": "columns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def cpu_baseline(args, vols_dev, gcm_host, aux_host, zf, zh, budget_s=12.0, ncols=None):
    """Single-core CPU legs on a bounded sample of THIS run's inputs (the first `ncols` columns, copied back from the
    device): the unmodified reference (kind "reference") and the numpy port."""
    ncols = min(ncols or min(64, REF_COLS[args.config]), vols_dev[0].shape[0])
    from sp_coupler_b200.constants import LES_FIELDS
    vols = {f: v[:ncols].cpu().numpy() for f, v in zip(LES_FIELDS, vols_dev)}
    if args.layout == "ijk":        # the CPU legs read the slab-contiguous view; same values
        vols = {f: np.ascontiguousarray(np.transpose(v, (0, 3, 2, 1))) for f, v in vols.items()}
    inp = (zf, zh, {k: v[:ncols] for k, v in gcm_host.items()}, {k: v[:ncols] for k, v in aux_host.items()}, vols)

    def run(fn, budget):
        fn(inp)
        t0 = time.perf_counter()
        passes = 0
        while True:
            fn(inp)
            passes += 1
            el = time.perf_counter() - t0
            if el >= budget or passes >= 200:
                return passes, el

    ncpu = os.cpu_count() or 1
    passes, el = run(_port_pass, budget_s * 0.4)
    port = {"value": ncols * passes / el, "unit": "columns/s", "cores": 1, "kind": "port",
            "sample": "%d columns x %d passes of oracle/numpy_batched.coupling_step in %.1f s on 1 core (host has %d)"
                      % (ncols, passes, el, ncpu)}
    if not reference_available():
        return port, None
    passes, el = run(_reference_pass, budget_s)
    ref = {"value": ncols * passes / el, "unit": "columns/s", "cores": 1, "kind": "reference",
           "sample": "%d columns x %d passes of the unmodified reference (spcpl.set_les_forcings + set_gcm_tendencies per "
                     "column under the unit shim + numpy slab means, oracle/ref_driver.reference_column_steps) in %.1f s on "
                     "1 core (host has %d)" % (ncols, passes, el, ncpu)}
    return ref, port


# ------------------------------------------------------------------------------------ GPU arm
def k1_traffic(config, layout, ncol):
    """DRAM bytes per K1 launch from the committed ncu --set full captures (profiles/k1_traffic.json holds
    dram__bytes_read.sum + dram__bytes_write.sum PER COLUMN for every config / layout that was captured; traffic is
    linear in the column count). None when there is no capture for this config / layout."""
    tp = os.path.join(ROOT, "profiles", "k1_traffic.json")
    try:
        tj = json.load(open(tp))
        e = tj["per_column"].get("%s:%s" % (config, layout))
        return None if e is None else float(e["dram_bytes_per_column"]) * ncol
    except Exception:
        return None


def run_b200(args):
    import torch
    import torch.distributed as dist
    import synth_les
    from sp_coupler_b200 import synth
    from sp_coupler_b200.coupler import Coupler
    from sp_coupler_b200.pipeline import CouplingPipeline, HostExchange, bind_host_thread_to_gpu

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    args.direct = args.e2e_route == "direct" or (args.e2e_route == "auto" and world > 1)
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("bench.py --gpus %d must be launched with torch.distributed.run --nproc-per-node %d"
                         % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        if args.bind:
            cpus = bind_host_thread_to_gpu(dev)     # host pages and PCIe copies on the GPU's NUMA node
            sys.stderr.write("rank %d: %s\n" % (rank, "bound to %d CPUs local to GPU %d (%d-%d)" % (len(cpus), local, cpus[0], cpus[-1])
                                                 if cpus else "CPU binding unavailable"))

    ncol_cfg, nx, ny, nk, nlev, dts = CONFIGS[args.config]
    if args.ncol:
        ncol_cfg = args.ncol
    tdt = torch.float32 if dts == "f32" else torch.float64
    ndt = np.float32 if dts == "f32" else np.float64
    esize = 4 if dts == "f32" else 8
    cpl = Coupler(dev)
    zf, zh = synth.les_grid(nk)
    bpc = alg_bytes_per_column(nx, ny, nk, esize)

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

    def agree(ok):
        if world == 1:
            return bool(ok)
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        return bool(int(flag.item()))

    def make_job(ncol_total):
        """Volumes, GCM columns and the device-gather pipeline of this rank's share of `ncol_total` columns."""
        if ncol_total % world:
            raise SystemExit("%d columns do not divide over %d GPUs" % (ncol_total, world))
        ncol = ncol_total // world
        col0 = rank * ncol
        gcm_host = synth.make_gcm_columns(ncol, nlev, seed=SEED, dtype=ndt, col0=col0, ncol_total=ncol_total)
        aux_host = synth.make_les_aux(ncol, nk, seed=SEED, dtype=ndt, col0=col0, ncol_total=ncol_total)
        gather, pipe = (args.gather if world > 1 else False), None
        if world > 1 and gather != "nccl":
            # the fused gather needs NVLink symmetric memory; agree across ranks, else use the NCCL collective
            try:
                pipe = CouplingPipeline(cpl, zf, zh, ncol, nlev, tdt, couple_surface=True, gather=gather, layout=args.layout)
                ok = True
            except Exception as e:          # noqa: BLE001
                sys.stderr.write("rank %d: symmetric-memory gather unavailable (%s); using NCCL all_gather\n" % (rank, e))
                ok = False
            if not agree(ok):
                gather, pipe = "nccl", None
        if pipe is None:
            pipe = CouplingPipeline(cpl, zf, zh, ncol, nlev, tdt, couple_surface=True, gather=gather, layout=args.layout)
        # the live level window of the job (pipeline.py, "Level window"): the GCM levels above the first one over the LES top
        # carry zero tendencies and are not needed for the forcings, so the resident GCM columns are cut there; when the
        # columns are sharded every rank runs the job-wide window (the gathered block is [ncol_total][7][nlw])
        lev0 = pipe.first_live_level(gcm_host) if args.window else 0
        if world > 1:
            lv = torch.tensor([lev0], device=dev)
            dist.all_reduce(lv, op=dist.ReduceOp.MIN)
            lev0 = int(lv.item())
        pipe.stage_host(gcm_host, lev0=lev0)
        pipe.staging.upload()
        vols = synth_les.device_les_volumes(cpl, gcm_host, zf, nx, ny, seed=SEED, dtype=tdt, col0=col0)
        if args.layout == "ijk":     # the (itot, jtot, ktot) C-order view OMUSE hands to Python: k fastest
            for i in range(len(vols)):
                vols[i] = vols[i].permute(0, 3, 2, 1).contiguous()
        aux = {k: torch.from_numpy(v).to(dev) for k, v in aux_host.items()}
        pipe.attach_les(vols, aux)
        pipe.les_profiles()                      # first-step slab means (spcpl.py:302-308)
        torch.cuda.synchronize()
        return dict(ncol=ncol, ncol_total=ncol_total, col0=col0, gcm_host=gcm_host, aux_host=aux_host, gather=gather,
                    pipe=pipe, vols=vols, aux=aux, lev0=lev0)

    def capture(pipe, upload=False):
        """All ranks capture together or nobody does. upload: the H2D copy of the staged GCM columns is the graph's first node."""
        if not args.graph or (pipe.gather and pipe.gather_mode == "nccl"):
            return False
        try:
            pipe.capture(DT, F_LES, F_GCM, upload=upload)
            ok = True
        except Exception as e:              # noqa: BLE001
            sys.stderr.write("rank %d: CUDA graph capture failed (%s); eager launches\n" % (rank, e))
            ok = False
        if not agree(ok):
            pipe._graphs.clear()
            return False
        return True

    ncol_total = ncol_cfg * world if args.scaling == "weak" else ncol_cfg
    job = make_job(ncol_total)
    pipe, ncol = job["pipe"], job["ncol"]
    sampler = ClockSampler(local) if rank == 0 else None

    # ---- device-resident leg (value): the step replayed from its CUDA graph ----
    graphed = capture(pipe)
    for _ in range(args.warmup):
        pipe.step(DT, F_LES, F_GCM)
    barrier()
    l0 = cpl.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        pipe.step(DT, F_LES, F_GCM)
    e1.record()
    barrier()
    launches = cpl.launches - l0          # our own kernels only: K2, K1, K3 per step
    ms_t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_step = float(ms_t.item()) / args.steps
    value_tend = pipe.tend_all.clone() if rank == 0 else None      # the gathered block of the last timed step
    # the same step for at least half a second (the K-step region above is only K x ms_per_step long: 10 ms at 8 GPUs)
    n_long = max(args.steps, int(np.ceil(500.0 / ms_step)))
    ms_long = timed(lambda: pipe.step(DT, F_LES, F_GCM), n_long, 0)
    # ---- roofline leg: the same step launched eagerly with CUDA events around K1 ----
    pipe.k1_events = []
    ms_eager = timed(lambda: pipe.step_device(DT, F_LES, F_GCM), args.steps, args.warmup)
    k1_ms = float(np.mean([a.elapsed_time(b) for a, b in pipe.k1_events[-args.steps:]]))
    pipe.k1_events = None

    # ---- end-to-end leg: host GCM buffers in, host tendencies out, every step ----
    hp = CouplingPipeline(cpl, zf, zh, ncol, nlev, tdt, couple_surface=True, gather=False, layout=args.layout)
    hp.attach_les(job["vols"], job["aux"])
    hp.les_profiles()
    exch, e2e_ok = None, None
    if world > 1:
        # the GCM lives in host memory of rank 0: one pinned host buffer shared by all ranks; every rank uploads ITS
        # columns over its own PCIe link and its K3 stores ITS tendencies straight into the shared buffer
        try:
            exch = HostExchange(hp, world, rank, owner=0, tag="bench", window=args.window, direct=args.direct)
            ok = True
        except Exception as e:      # noqa: BLE001
            sys.stderr.write("rank %d: shared pinned host buffer unavailable (%s)\n" % (rank, e))
            ok = False
        if not agree(ok):
            raise SystemExit("bench.py: the shared pinned host buffer of the e2e leg could not be set up")
        if rank == 0:      # the host GCM's profiles of ALL columns (identical to what each rank generated for itself)
            exch.fill_inputs(synth.make_gcm_columns(ncol_total, nlev, seed=SEED, dtype=ndt, col0=0, ncol_total=ncol_total))
        exch.step(DT, F_LES, F_GCM)           # adopts the level window
        capture(hp, upload=True)
        ms_e2e = timed(lambda: exch.step(DT, F_LES, F_GCM), args.steps, args.warmup)
        lev0 = exch.lev0
        e2e_h2d = world * hp.staging.nbytes
        e2e_d2h = world * ncol * 7 * hp.nlw * esize
        e2e_note = ("GCM profiles in / tendencies out of ONE pinned host buffer shared by all ranks (the host GCM's memory): "
                    "every rank replays ONE graph (H2D of its own columns over its own PCIe link + the step's three kernels); " +
                    ("its K3 stores its tendency block into the shared buffer and raises a flag the owner polls (no D2H copy "
                     "call, no stream synchronisation); " if args.direct else
                     "it copies its tendency block into the shared buffer (copy engine), synchronises its stream and raises "
                     "a flag the owner polls; ") +
                    "no device gather; bytes are totals over ranks; LES volumes are device-resident LES state")
        if rank == 0:
            e2e_ok = bool(lev0 >= job["lev0"] and torch.equal(exch.out(), value_tend[:, :, lev0 - job["lev0"]:].cpu()))
    else:
        lev0 = hp.stage_host(job["gcm_host"], window=args.window)
        if args.direct:
            hp.bind_host_output()
        capture(hp, upload=True)
        ms_e2e = timed(lambda: hp.step_host(DT, F_LES, F_GCM), args.steps, args.warmup)
        e2e_h2d = hp.staging.nbytes
        e2e_d2h = ncol * 7 * hp.nlw * esize
        e2e_note = ("per step: GCM profiles H2D from pinned host memory (one copy) and the step's three kernels replayed as ONE CUDA graph, " +
                    ("K3 stores the tendencies into pinned host memory and raises a flag the host polls; " if args.direct else
                     "tendencies D2H into pinned host memory (one copy), stream synchronised; ") +
                    "LES volumes are device-resident LES state")
        e2e_ok = bool(lev0 >= job["lev0"] and torch.equal(hp.tend_host, value_tend[:, :, lev0 - job["lev0"]:].cpu()))
    e2e_note += "; GCM levels %d..%d of %d travel (the window up to the first level above the LES top; tendencies above are zero)" % (
        lev0, nlev - 1, nlev) if lev0 else "; all %d GCM levels travel" % nlev
    # for the record (1 GPU): the round-1 form of the leg - all levels both ways, D2H copy + stream synchronise
    e2e_full = None
    if world == 1 and args.window:
        fp = CouplingPipeline(cpl, zf, zh, ncol, nlev, tdt, couple_surface=True, gather=False, layout=args.layout)
        fp.attach_les(job["vols"], job["aux"])
        fp.les_profiles()
        fp.staging.fill_host(job["gcm_host"])
        fp.staging.upload()
        capture(fp, upload=True)
        ms_f = timed(lambda: fp.step_host(DT, F_LES, F_GCM), args.steps, args.warmup)
        e2e_full = {"value": ncol_total / (ms_f * 1e-3), "unit": "columns/s", "ms_per_step": ms_f,
                    "h2d_bytes_per_step": fp.staging.nbytes, "d2h_bytes_per_step": ncol * 7 * nlev * esize,
                    "note": "all %d levels H2D, device block copied back D2H, stream synchronised (no level window, no "
                            "direct stores)" % nlev}
        del fp
    clocks = sampler.stop() if sampler else None

    # ---- for the record: the same step if the LES volumes lived in HOST memory (they do not) ----
    host_vol = None
    if world == 1 and args.host_volume_cols > 0:
        hc = min(args.host_volume_cols, ncol)
        hpipe = CouplingPipeline(cpl, zf, zh, hc, nlev, tdt, couple_surface=True, layout=args.layout)
        hpipe.staging.fill_host({k: v[:hc] for k, v in job["gcm_host"].items()})
        hvols_host = [torch.empty((hc,) + tuple(v.shape[1:]), dtype=tdt, pin_memory=True) for v in job["vols"]]
        for hv, v in zip(hvols_host, job["vols"]):
            hv.copy_(v[:hc])
        hvols_dev = [torch.empty_like(v[:hc]) for v in job["vols"]]
        hpipe.attach_les(hvols_dev, {k: v[:hc].contiguous() for k, v in job["aux"].items()})
        hpipe.les_profiles()

        def host_volume_step():
            for dv, hv in zip(hvols_dev, hvols_host):
                dv.copy_(hv, non_blocking=True)
            hpipe.step_host(DT, F_LES, F_GCM)

        ms_hv = timed(host_volume_step, max(3, args.steps // 4), 2)
        host_vol = {"value": hc / (ms_hv * 1e-3), "unit": "columns/s", "columns": hc, "ms_per_step": ms_hv,
                    "h2d_bytes_per_step": int(sum(h.numel() * h.element_size() for h in hvols_host)) + hpipe.staging.nbytes,
                    "note": "NOT the design point: LES volumes copied from pinned host memory every step (PCIe-bound). The "
                            "speed-up of this path over the CPU requires the LES state to live in HBM"}
        del hvols_host, hvols_dev, hpipe

    # ---- N>1: the round-1 weak-scaling point (the whole column count on EVERY GPU) ----
    weak = None
    if world > 1 and args.scaling == "strong" and args.weak_leg:
        if exch is not None:
            torch.cuda.synchronize()
            dist.barrier()
            exch.close()
        del hp
        wjob = make_job(ncol_cfg * world)
        wp = wjob["pipe"]
        capture(wp)
        ms_w = timed(lambda: wp.step(DT, F_LES, F_GCM), args.steps, args.warmup)
        weak = {"value": ncol_cfg * world / (ms_w * 1e-3), "unit": "columns/s", "ms_per_step": ms_w, "scaling": "weak",
                "ncol_total": ncol_cfg * world, "ncol_per_gpu": ncol_cfg,
                "note": "every GPU owns the whole column count of the config (round-1 bench definition)"}
        # and its host-to-host form (shared pinned host buffer), as in the strong e2e leg
        try:
            whp = CouplingPipeline(cpl, zf, zh, ncol_cfg, nlev, tdt, couple_surface=True, gather=False, layout=args.layout)
            whp.attach_les(wjob["vols"], wjob["aux"])
            whp.les_profiles()
            wex = HostExchange(whp, world, rank, owner=0, tag="benchweak", window=args.window, direct=args.direct)
            ok = True
        except Exception as e:      # noqa: BLE001
            sys.stderr.write("rank %d: weak e2e leg unavailable (%s)\n" % (rank, e))
            ok = False
        if agree(ok):
            if rank == 0:
                wex.fill_inputs(synth.make_gcm_columns(ncol_cfg * world, nlev, seed=SEED, dtype=ndt, col0=0, ncol_total=ncol_cfg * world))
            wex.step(DT, F_LES, F_GCM)
            capture(whp, upload=True)
            ms_we = timed(lambda: wex.step(DT, F_LES, F_GCM), args.steps, args.warmup)
            weak["e2e"] = {"value": ncol_cfg * world / (ms_we * 1e-3), "unit": "columns/s", "ms_per_step": ms_we,
                           "h2d_bytes_per_step": world * whp.staging.nbytes, "d2h_bytes_per_step": world * ncol_cfg * 7 * whp.nlw * esize}
            torch.cuda.synchronize()
            dist.barrier()
            wex.close()
            del wex, whp
        del wjob, wp
        torch.cuda.empty_cache()

    # ---- N>1: the gathered block on the GCM owner against the owner's own single-GPU computation of EVERY column ----
    gather_check = None
    sync_err = pipe.sync_error()
    if world > 1:
        if rank == 0:
            same = True
            gcm_all = synth.make_gcm_columns(ncol_total, nlev, seed=SEED, dtype=ndt, col0=0, ncol_total=ncol_total)
            aux_all = synth.make_les_aux(ncol_total, nk, seed=SEED, dtype=ndt, col0=0, ncol_total=ncol_total)
            sp = CouplingPipeline(cpl, zf, zh, ncol, nlev, tdt, couple_surface=True, gather=False, layout=args.layout)
            for r in range(world):
                sl = slice(r * ncol, (r + 1) * ncol)
                g = {k: v[sl] for k, v in gcm_all.items()}
                v = synth_les.device_les_volumes(cpl, g, zf, nx, ny, seed=SEED, dtype=tdt, col0=r * ncol)
                if args.layout == "ijk":
                    v = [x.permute(0, 3, 2, 1).contiguous() for x in v]
                sp.stage_host(g, lev0=job["lev0"])
                sp.staging.upload()
                sp.attach_les(v, {k: torch.from_numpy(np.ascontiguousarray(a[sl])).to(dev) for k, a in aux_all.items()})
                sp.slab = None
                sp.les_profiles()
                sp.step_device(DT, F_LES, F_GCM)
                torch.cuda.synchronize()
                same = same and bool(torch.equal(sp.tend, value_tend[sl]))
                del v
            gather_check = bool(same and (e2e_ok is not False) and sync_err == 0)
        dist.barrier()

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
    achieved = bpc * ncol / (k1_ms * 1e-3) / 1e9
    gather = job["gather"]
    par = "1 GPU"
    if world > 1:
        par = "columns sharded x%d, tendencies %s" % (world, "all_gather (NCCL)" if gather == "nccl" else
              "gathered by K3 itself (%s): NVLink peer stores into symmetric memory + in-kernel barrier" % gather)
    line = {
        "metric": METRIC, "value": ncol_total / (ms_step * 1e-3), "unit": "columns/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.config, ncol_total, world, args.scaling), "ncol_total": ncol_total,
                   "ncol_per_gpu": ncol, "storage_dtype": dts, "arithmetic": "f64", "layout": args.layout,
                   "l2": "inputs larger than L2 (%.1f GB of LES volumes per GPU streamed once per step; no reuse between steps)" % (bpc * ncol / 1e9),
                   "parallelism": par,
                   "gcm_levels": "%d..%d of %d resident (live level window; tendencies above are zero)" % (job["lev0"], nlev - 1, nlev)
                                 if job["lev0"] else "all %d" % nlev,
                   "step": "K2 gcm_to_les -> K1 slab_reduce -> K3 les_to_gcm (cloud projection, tendencies, delivery)",
                   "launch": "one CUDA graph replay per step" if graphed else "eager: three C-ABI calls per step"},
        "roofline": {"kernel": "slab_reduce_tma_kernel" if args.layout == "kji" else "slab_reduce_ijk_tma_kernel", "bound": "hbm",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": k1_traffic(args.config, args.layout, ncol),
                     "peak_source": peak_src, "frac_of_nominal_8TBs": achieved / 8000.0, "k1_ms": k1_ms,
                     "k1_share_of_step": k1_ms / ms_eager, "ms_per_step_eager": ms_eager, "alg_bytes_per_launch": bpc * ncol,
                     "timing": "CUDA events around K1 in an eager pass of the same step, same run (events cannot bracket a kernel "
                               "inside a graph replay)"},
        "e2e": {"value": ncol_total / (ms_e2e * 1e-3), "unit": "columns/s", "h2d_bytes_per_step": e2e_h2d,
                "d2h_bytes_per_step": e2e_d2h, "ms_per_step": ms_e2e, "identical_to_value_leg": e2e_ok, "note": e2e_note},
        "sustained": {"value": ncol_total / (ms_long * 1e-3), "unit": "columns/s", "steps": n_long, "ms_per_step": ms_long,
                      "note": "the value leg's step repeated for >= 0.5 s (same graph, same inputs), max over ranks"},
        "e2e_full_levels": e2e_full,
        "e2e_host_volumes": host_vol,
        "weak": weak,
        "gather_check": gather_check,
        "gpu_launches": launches,
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu:
        ref, port = cpu_baseline(args, job["vols"], job["gcm_host"], job["aux_host"], zf, zh, budget_s=args.cpu_budget)
        line["cpu_baseline"] = ref
        if port is not None:
            line["cpu_baseline_port"] = port
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if gather_check is False or e2e_ok is False or sync_err:
        sys.stderr.write("bench.py: CHECK FAILED gather_check=%s e2e_identical=%s sync_error=%d\n" % (gather_check, e2e_ok, sync_err))
        return 3
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong (default, BASELINE.json as written): the config's columns in total over N GPUs; weak: on every GPU")
    ap.add_argument("--ncol", type=int, default=0, help="override the config's column count")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    ap.add_argument("--host-volume-cols", type=int, default=64,
                    help="columns of the extra 'volumes in host memory' measurement (0 = skip)")
    ap.add_argument("--ref-procs", type=int, default=0)
    ap.add_argument("--ref-cols", type=int, default=0, help="distinct columns per reference worker (default by config: 64 at C3)")
    ap.add_argument("--ref-port", action="store_true", help="--impl reference: time the numpy port instead of the unmodified reference")
    ap.add_argument("--layout", default="kji", choices=["kji", "ijk"],
                    help="memory order of the LES volumes: kji = [ncol][nk][ny][nx] (DALES), ijk = [ncol][nx][ny][nk] (OMUSE view)")
    ap.add_argument("--no-bind", dest="bind", action="store_false", help="N>1: do not bind rank processes to their GPU's CPUs")
    ap.add_argument("--no-window", dest="window", action="store_false", help="e2e: ship all GCM levels instead of the live window")
    ap.add_argument("--e2e-route", default="auto", choices=["auto", "copy", "direct"],
                    help="e2e: how the tendencies reach pinned host memory. copy = copy engine D2H + stream synchronise; direct = K3 "
                         "stores them there itself and raises a flag the host polls. auto = copy on one GPU (the copy engine moves "
                         "~54 GB/s, SM-issued PCIe stores ~33 GB/s), direct when the columns are sharded (no host call after the "
                         "launch on any rank; 0.6 %% faster at 8 GPUs)")
    ap.add_argument("--no-weak-leg", dest="weak_leg", action="store_false", help="N>1: skip the weak-scaling sub-record")
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="launch the steps eagerly instead of replaying CUDA graphs")
    ap.add_argument("--gather", default="p2p-owner", choices=["nccl", "p2p", "p2p-owner"],
                    help="multi-GPU tendency gather: fused into K3 (NVLink peer stores to the GCM owner / to every rank), or NCCL all_gather")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
