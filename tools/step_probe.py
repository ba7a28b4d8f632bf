"""Per-kernel timing of one coupled step (K2, K1, K3) with CUDA events; not the bench."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 4:   # the thread sweep needs spc_tune_profiles: only in the -DSPC_TUNING build (python -m sp_coupler_b200.build --tune)
    os.environ.setdefault("SPCPL_B200_LIB", os.path.join(ROOT, "sp_coupler_b200", "lib", "libspcpl_b200_tune.so"))
import numpy as np
import torch

sys.path.insert(1, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import synth_les
from sp_coupler_b200 import synth
from sp_coupler_b200.coupler import Coupler
from sp_coupler_b200.pipeline import CouplingPipeline

ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
nx = int(sys.argv[2]) if len(sys.argv) > 2 else 64
nlev = int(sys.argv[3]) if len(sys.argv) > 3 else 91
dev = torch.device("cuda:0")
cpl = Coupler(dev)
zf, zh = synth.les_grid(160)
gcm = synth.make_gcm_columns(ncol, nlev, dtype=np.float32)
aux = synth.make_les_aux(ncol, 160, dtype=np.float32)
pipe = CouplingPipeline(cpl, zf, zh, ncol, nlev, torch.float32)
pipe.staging.fill_host(gcm)
pipe.staging.upload()
pipe.attach_les(synth_les.device_les_volumes(cpl, gcm, zf, nx, nx), {k: torch.from_numpy(v).to(dev) for k, v in aux.items()})
pipe.les_profiles()
torch.cuda.synchronize()


def t(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), r


from sp_coupler_b200 import _abi
frc = pipe.forcings(900.0, 1.0)
if len(sys.argv) > 4:      # sweep of threads per column CTA for K2 / K3 (spc_tune_profiles)
    for th in [int(x) for x in sys.argv[4].split(",")]:
        _abi.lib().spc_tune_profiles(pipe.cpl._h, 0, 0, th)
        a, _ = t(lambda: pipe.forcings(900.0, 1.0))
        b, _ = t(lambda: pipe.tendencies(frc, 900.0, 1.0))
        print("threads %d: K2 %.1f us  projection+K3 %.1f us" % (th, a * 1e3, b * 1e3))
    _abi.lib().spc_tune_profiles(pipe.cpl._h, 0, 0, 0)
k2, frc = t(lambda: pipe.forcings(900.0, 1.0))
k1, _ = t(lambda: pipe.les_profiles())
k3, _ = t(lambda: pipe.tendencies(frc, 900.0, 1.0))
st, _ = t(lambda: pipe.step_device())
sh, _ = t(lambda: pipe.step_host())
print("ncol=%d %dx%dx160 L%d: K2 %.3f ms  K1 %.3f ms  K3 %.3f ms  step_device %.3f ms  step_host %.3f ms" % (ncol, nx, nx, nlev, k2, k1, k3, st, sh))
