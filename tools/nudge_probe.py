"""Times spc_variability_nudge (K6) on a device-generated batch; not the bench.
Traffic model: qt read once + written once (+ thl/ql with --constT); the Brent iterations run on the
shared-memory copy of the slab."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from sp_coupler_b200.coupler import Coupler

ap = argparse.ArgumentParser()
ap.add_argument("--ncol", type=int, default=512)
ap.add_argument("--nx", type=int, default=64)
ap.add_argument("--nk", type=int, default=160)
ap.add_argument("--constT", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda:0")
cpl = Coupler(dev)
g = torch.Generator(device=dev).manual_seed(5)
ncol, nk, nx = a.ncol, a.nk, a.nx
qt_p = 0.008 * torch.exp(-torch.arange(nk, device=dev, dtype=torch.float64) / 8.0)
amp = 2.5e-5 * (0.05 + 3.95 * torch.rand((ncol, nk, 1, 1), device=dev, generator=g, dtype=torch.float64))
qt0 = (qt_p[None, :, None, None] + amp * (2 * torch.rand((ncol, nk, nx, nx), device=dev, generator=g) - 1)).float()
qsat_prof = (qt_p[None, :] + 2.5e-5 * (3 * torch.rand((ncol, nk), device=dev, generator=g, dtype=torch.float64) - 1.5)).float()
ql = (qt0.double() - qsat_prof.double()[:, :, None, None]).clamp_min(0).float()
thl0 = (290 + torch.randn((ncol, nk, nx, nx), device=dev, generator=g)).float()
prof = torch.zeros((5, ncol, nk), dtype=torch.float64, device=dev)
prof[1] = qt0.double().mean(dim=(2, 3))
prof[2] = ql.double().mean(dim=(2, 3))
mult = torch.tensor([0.0, 0.5, 1.5, 3.0, 40.0], device=dev, dtype=torch.float64)
ql_ref = (prof[2] * mult[torch.randint(0, 5, (ncol, nk), device=dev, generator=g)]).float()
presf = (1e5 * torch.exp(-torch.arange(nk, device=dev, dtype=torch.float64) * 25 / 7500.0))[None, :].repeat(ncol, 1).float().contiguous()
R = torch.randn((ncol, nx, nx), device=dev, generator=g, dtype=torch.float64)
R = (R - R.mean(dim=(1, 2), keepdim=True)).contiguous()
ts = []
for it in range(6):
    qt, thl = qt0.clone(), thl0.clone()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = cpl.variability_nudge(qt, prof, ql_ref, 900.0, qsat_prof=qsat_prof, R=R, constant_T=a.constT,
                                thl=thl if a.constT else None, ql=ql if a.constT else None, presf=presf if a.constT else None)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = float(np.median(ts[1:]))
st = out["status"]
nbytes = qt0.numel() * 4 * (2 + (3 if a.constT else 0))
print("nudge ncol=%d %dx%dx%d constT=%s: %.3f ms, %.0f columns/s, %.1f GB/s of qt(+thl,ql) traffic; levels: mult %d, unsat %d, additive %d, untouched %d"
      % (ncol, nx, nx, nk, a.constT, ms, ncol / ms * 1e3, nbytes / ms / 1e6, int((st & 1).ne(0).sum()), int((st & 2).ne(0).sum()),
         int((st & 4).ne(0).sum()), int(st.eq(0).sum())))
