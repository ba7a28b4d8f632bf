"""torch.ops.spcpl_b200.* — the thin PyTorch C++ extension over the C ABI (csrc/torch_ext.cpp).

An alternative binding to the same kernels as coupler.Coupler (which uses ctypes): tensors in, torch's
current stream, no ctypes. `step()` runs one coupled step (K2 -> K1 -> K3) purely through the
registered ops. Fails loudly if the extension has not been built (`python -m sp_coupler_b200.build
--torch`); there is no fallback to any other path.
"""
import os

import torch

from . import _abi
from .constants import TENDENCIES

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "spcpl_b200_torch.so")
_loaded = False


def load():
    global _loaded
    if not _loaded:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("sp_coupler_b200: %s is missing; build it with "
                               "`python -m sp_coupler_b200.build --torch`" % LIB_PATH)
        torch.ops.load_library(LIB_PATH)
        _loaded = True
    return torch.ops.spcpl_b200


_DT = {torch.float32: _abi.SPC_F32, torch.float64: _abi.SPC_F64}


def slab_reduce(vols, layout=0, ql_thresh=0.0):
    ops = load()
    v0 = vols[0]
    if layout == 0:
        ncol, nk, ny, nx = v0.shape
    else:
        ncol, nx, ny, nk = v0.shape
    dev = v0.device
    prof = torch.empty((5, ncol, nk), dtype=torch.float64, device=dev)
    cnt = torch.empty((ncol, nk), dtype=torch.int32, device=dev)
    mw = ops.mask_words_per_column(_DT[v0.dtype], layout, nx, ny, nk)
    mask = torch.empty((ncol, mw), dtype=torch.int32, device=dev)
    ops.slab_reduce(list(vols), layout, float(ql_thresh), prof, cnt, mask)
    return dict(prof=prof, cnt=cnt, mask=mask, nx=nx, ny=ny, dtype=v0.dtype, layout=layout)


def step(gcm, zf, zh, vols, aux, les_prof_prev, dt=900.0, f_les=1.0, f_gcm=1.0, couple_surface=True, layout=0):
    """One coupled step through the registered ops: forcings from the previous slab means (K2), new
    slab means + cloud mask (K1), tendencies (K3). Returns (forcings dict, slab dict, tendencies dict)."""
    ops = load()
    T = gcm["T"]
    ncol, nlev = T.shape
    nk = zf.shape[0]
    dev, dtype = T.device, T.dtype
    e = lambda *s, dt_=dtype: torch.empty(s, dtype=dt_, device=dev)
    frc = {k: e(ncol, nk) for k in ("f_u", "f_v", "f_thl", "f_qt", "f_ql", "ql_ref")}
    frc.update({k: e(ncol) for k in ("f_ps", "ps", "z0m", "z0h", "wthl", "wqt")})
    frc["slab_idx"] = e(ncol, nlev, dt_=torch.int32)
    ops.gcm_to_les(gcm, zf, zh, les_prof_prev, aux["PS"], float(dt), float(f_les), bool(couple_surface), frc)
    slab = slab_reduce(vols, layout)
    tnd = {"tend": e(ncol, 7, nlev), "A_d": e(ncol, nlev), "cntslab": e(ncol, nlev, dt_=torch.int32),
           "start_index": e(ncol, dt_=torch.int32)}
    les = {"prof": slab["prof"], "QL_ice": aux["QL_ice"], "T": aux["T"], "mask": slab["mask"], "cnt": slab["cnt"],
           "slab_idx": frc["slab_idx"]}
    ops.les_to_gcm(gcm, zf, zh, les, slab["nx"], slab["ny"], layout, _DT[slab["dtype"]], float(dt), float(f_gcm), False, tnd)
    for i, name in enumerate(TENDENCIES):
        tnd[name] = tnd["tend"][:, i, :]
    return frc, slab, tnd
