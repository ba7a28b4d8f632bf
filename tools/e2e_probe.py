"""Where the host-facing step spends its time (not the bench): GPU-side durations of the pieces with CUDA events over
back-to-back launches, and the host-synchronous step by wall clock.
    python tools/e2e_probe.py [config] [ncol]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from bench import CONFIGS
sys.path.insert(1, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import synth_les
from sp_coupler_b200 import synth
from sp_coupler_b200.coupler import Coupler
from sp_coupler_b200.pipeline import CouplingPipeline

cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
ncol, nx, ny, nk, nlev, dts = CONFIGS[cfg]
if len(sys.argv) > 2:
    ncol = int(sys.argv[2])
tdt, ndt = (torch.float32, np.float32) if dts == "f32" else (torch.float64, np.float64)
dev = torch.device("cuda:0")
cpl = Coupler(dev)
zf, zh = synth.les_grid(nk)
gcm = synth.make_gcm_columns(ncol, nlev, seed=44, dtype=ndt)
aux = {k: torch.from_numpy(v).to(dev) for k, v in synth.make_les_aux(ncol, nk, seed=44, dtype=ndt).items()}
vols = synth_les.device_les_volumes(cpl, gcm, zf, nx, ny, seed=44, dtype=tdt)


def gpu_ms(fn, n=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def wall_ms(fn, n=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


def make(window, host_out):
    p = CouplingPipeline(cpl, zf, zh, ncol, nlev, tdt, couple_surface=True, gather=False)
    p.attach_les(vols, aux)
    p.stage_host(gcm, window=window)
    if host_out:
        p.bind_host_output()
    p.staging.upload()
    p.les_profiles()
    p.capture(900.0, 1.0, 1.0)
    return p


dv = make(True, False)      # window, K3 -> device only
ho = make(True, True)       # window, K3 -> device + pinned host + flag
full = make(False, False)   # all levels, device only
print("%s ncol=%d: window %d of %d levels, staging %.2f MB, tendencies %.2f MB" % (cfg, ncol, dv.nlw, nlev, dv.staging.nbytes / 1e6,
      ncol * 7 * dv.nlw * dv.tend.element_size() / 1e6))
print("GPU-side, back to back:")
print("  H2D window staging            %.4f ms" % gpu_ms(dv.staging.upload))
print("  H2D full staging              %.4f ms" % gpu_ms(full.staging.upload))
print("  graph step, K3 -> device      %.4f ms" % gpu_ms(lambda: dv.step(900.0, 1.0, 1.0)))
print("  graph step, K3 -> host too    %.4f ms" % gpu_ms(lambda: ho.step(900.0, 1.0, 1.0)))
print("  graph step, all levels        %.4f ms" % gpu_ms(lambda: full.step(900.0, 1.0, 1.0)))
print("  H2D + step (K3 -> host)       %.4f ms" % gpu_ms(lambda: (ho.staging.upload(), ho.step(900.0, 1.0, 1.0))))
print("host-synchronous (wall clock):")
print("  step_host, window + host stores + flag poll   %.4f ms" % wall_ms(lambda: ho.step_host(900.0, 1.0, 1.0)))
print("  step_host, window, D2H copy + stream sync     %.4f ms" % wall_ms(lambda: dv.step_host(900.0, 1.0, 1.0)))
print("  step_host, all levels, D2H copy + stream sync %.4f ms" % wall_ms(lambda: full.step_host(900.0, 1.0, 1.0)))
print("  graph replay + stream sync only               %.4f ms" % wall_ms(lambda: (dv.step(900.0, 1.0, 1.0), torch.cuda.current_stream().synchronize())))
