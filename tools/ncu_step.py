"""Minimal driver for ncu captures: builds one config's pipeline on cuda:0 and runs a few eager steps (K2, K1, K3),
optionally through the host-output route. Not the bench (no CPU legs, no timing).

    ncu --set full --clock-control none --import-source on -k regex:slab_reduce -c 1 -o gpurun_out/k1_c3 \
        python tools/ncu_step.py c3 [kji|ijk] [ncol] [steps] [host]
"""
import os
import sys

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
layout = sys.argv[2] if len(sys.argv) > 2 else "kji"
ncol, nx, ny, nk, nlev, dts = CONFIGS[cfg]
if len(sys.argv) > 3 and int(sys.argv[3]) > 0:
    ncol = int(sys.argv[3])
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
host = len(sys.argv) > 5 and sys.argv[5] == "host"
tdt, ndt = (torch.float32, np.float32) if dts == "f32" else (torch.float64, np.float64)
dev = torch.device("cuda:0")
cpl = Coupler(dev)
zf, zh = synth.les_grid(nk)
gcm = synth.make_gcm_columns(ncol, nlev, seed=44, dtype=ndt)
aux = {k: torch.from_numpy(v).to(dev) for k, v in synth.make_les_aux(ncol, nk, seed=44, dtype=ndt).items()}
vols = synth_les.device_les_volumes(cpl, gcm, zf, nx, ny, seed=44, dtype=tdt)
if layout == "ijk":
    vols = [v.permute(0, 3, 2, 1).contiguous() for v in vols]
pipe = CouplingPipeline(cpl, zf, zh, ncol, nlev, tdt, couple_surface=True, gather=False, layout=layout)
pipe.attach_les(vols, aux)
if host:
    pipe.stage_host(gcm)
    pipe.bind_host_output()
else:
    pipe.staging.fill_host(gcm)
pipe.staging.upload()
pipe.les_profiles()
torch.cuda.synchronize()
for _ in range(steps):
    if host:
        pipe.step_host(900.0, 1.0, 1.0)
    else:
        pipe.step_device(900.0, 1.0, 1.0)
torch.cuda.synchronize()
print("ncu_step: %s %s ncol=%d steps=%d host=%s done" % (cfg, layout, ncol, steps, host))
