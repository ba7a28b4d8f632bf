"""K3 variants back to back (time them with `ncu --metrics gpu__time_duration.sum -k regex:les_to_gcm`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

sys.path.insert(1, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import synth_les
from sp_coupler_b200 import synth
from sp_coupler_b200.coupler import Coupler
from sp_coupler_b200.pipeline import CouplingPipeline

ncol, nx, nlev = 2048, int(sys.argv[1]) if len(sys.argv) > 1 else 64, 91
dev = torch.device("cuda:0")
cpl = Coupler(dev)
zf, zh = synth.les_grid(160)
gcm = synth.make_gcm_columns(ncol, nlev, dtype=np.float32)
aux = synth.make_les_aux(ncol, 160, dtype=np.float32)
pipe = CouplingPipeline(cpl, zf, zh, ncol, nlev, torch.float32)
pipe.staging.fill_host(gcm)
pipe.staging.upload()
pipe.attach_les(synth_les.device_les_volumes(cpl, gcm, zf, nx, nx), {k: torch.from_numpy(v).to(dev) for k, v in aux.items()})
slab = pipe.les_profiles()
frc = pipe.forcings(900.0, 1.0)
torch.cuda.synchronize()
A = torch.rand((ncol, nlev), device=dev)
nocnt = dict(slab, cnt=None)
for name, kw in (("mask+cnt", dict(slab=slab)), ("mask, no cnt skip", dict(slab=nocnt)), ("A given (no projection)", dict(slab={"prof": slab["prof"]}, A=A))):
    for _ in range(3):
        cpl.les_to_gcm(pipe.gcm, pipe.zf, pipe.zh, kw["slab"], pipe.aux, frc["slab_idx"], 900.0, 1.0, A=kw.get("A"), tend_out=pipe.tend)
    torch.cuda.synchronize()
    print(name)
