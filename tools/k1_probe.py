"""Quick K1 bandwidth probe (not the bench): times spc_slab_reduce alone with CUDA events."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
# the sweep variants and the per-handle setter spc_tune_k1 exist only in the -DSPC_TUNING build (python -m sp_coupler_b200.build --tune)
os.environ.setdefault("SPCPL_B200_LIB", os.path.join(ROOT, "sp_coupler_b200", "lib", "libspcpl_b200_tune.so"))
import numpy as np
import torch

sys.path.insert(1, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import synth_les
from sp_coupler_b200 import synth
from sp_coupler_b200.coupler import Coupler

ap = argparse.ArgumentParser()
ap.add_argument("--ncol", type=int, default=512)
ap.add_argument("--nx", type=int, default=64)
ap.add_argument("--nk", type=int, default=160)
ap.add_argument("--dtype", default="f32")
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--layout", default="kji")
ap.add_argument("--nomask", action="store_true")
ap.add_argument("--variants", default="0", help="comma list of spc_tune_k1 ring variants to time")
ap.add_argument("--lib", default="", help="alternative libspcpl_b200 build to time (A/B on the same box)")
a = ap.parse_args()
if a.lib:
    from sp_coupler_b200 import _abi as _abi0
    _abi0.LIB_PATH = os.path.abspath(a.lib)
dev = torch.device("cuda:0")
cpl = Coupler(dev)
td = torch.float32 if a.dtype == "f32" else torch.float64
zf, zh = synth.les_grid(a.nk)
gcm = synth.make_gcm_columns(a.ncol, 91)
t0 = time.time()
vols = synth_les.device_les_volumes(cpl, gcm, zf, a.nx, a.nx, dtype=td)
torch.cuda.synchronize()
print("generated %.2f GB in %.2fs" % (sum(v.numel() * v.element_size() for v in vols) / 1e9, time.time() - t0))
if a.layout == "ijk":
    vols = [v.permute(0, 3, 2, 1).contiguous() for v in vols]
nbytes = sum(v.numel() * v.element_size() for v in vols)
from sp_coupler_b200 import _abi
for variant in [int(x) for x in a.variants.split(",")]:
    _abi.lib().spc_tune_k1(cpl._h, variant)
    for _ in range(3):
        s = cpl.slab_reduce(vols, layout=a.layout, want_mask=not a.nomask)
    torch.cuda.synchronize()
    ts = []
    for _ in range(a.iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s = cpl.slab_reduce(vols, layout=a.layout, want_mask=not a.nomask)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts = np.array(ts)
    print("K1 variant %d %s ncol=%d %dx%dx%d %s mask=%s: median %.3f ms min %.3f ms -> %.1f GB/s (median) %.1f GB/s (best); %.0f col/s"
          % (variant, a.layout, a.ncol, a.nx, a.nx, a.nk, a.dtype, not a.nomask, np.median(ts), ts.min(),
             nbytes / np.median(ts) / 1e6, nbytes / ts.min() / 1e6, a.ncol / np.median(ts) * 1e3))
_abi.lib().spc_tune_k1(cpl._h, 0)
_abi.lib().spc_tune_k1(cpl._h, 100)
# set_les_state write bandwidth
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
prof = torch.zeros((a.ncol, a.nk), dtype=torch.float64, device=dev)
e0.record()
cpl.set_les_state(prof, 0.1, 0, a.nx, a.nx, dtype=td, out=vols[0] if a.layout == "kji" else None)
e1.record()
torch.cuda.synchronize()
print("set_les_state: %.3f ms -> %.1f GB/s written" % (e0.elapsed_time(e1), vols[0].numel() * vols[0].element_size() / e0.elapsed_time(e1) / 1e6))
