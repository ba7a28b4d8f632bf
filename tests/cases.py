"""Seeded parity cases shared by the GPU tests, smoke() and bench.py: host inputs, the oracle's
answer, and the device upload."""
import numpy as np

from oracle import numpy_batched as nb
import synth_les
from sp_coupler_b200 import synth
from sp_coupler_b200.constants import LES_FIELDS


def host_case(ncol, nx, ny, nk, nlev, dtype=np.float32, seed=42, layout=0, dz=None):
    dz = dz if dz is not None else (25.0 if nk >= 100 else 200.0)
    zf, zh = synth.les_grid(nk, dz)
    gcm = synth.make_gcm_columns(ncol, nlev, seed=seed, dtype=dtype)
    aux = synth.make_les_aux(ncol, nk, seed=seed, dtype=dtype)
    vols = synth_les.make_les_volumes(gcm, zf, nx, ny, seed=seed, dtype=dtype)
    if layout == 1:
        vols = {f: np.ascontiguousarray(np.transpose(v, (0, 3, 2, 1))) for f, v in vols.items()}
    return dict(zf=zf, zh=zh, gcm=gcm, aux=aux, vols=vols, ncol=ncol, nx=nx, ny=ny, nk=nk, nlev=nlev,
                dtype=dtype, layout=layout)


def oracle_step(case, dt=900.0, f_les=1.0, f_gcm=1.0, ql_thresh=0.0):
    return nb.coupling_step(case["gcm"], case["zf"], case["zh"], case["vols"], case["aux"], case["aux"]["PS"],
                            dt, f_les, f_gcm, True, ql_thresh, case["layout"])


def to_device(case, device):
    import torch
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    return dict(zf=t(case["zf"]), zh=t(case["zh"]), gcm={k: t(v) for k, v in case["gcm"].items()},
                aux={k: t(v) for k, v in case["aux"].items()},
                vols=[t(case["vols"][f]) for f in LES_FIELDS])


def gpu_step(cpl, d, layout=0, dt=900.0, f_les=1.0, f_gcm=1.0, ql_thresh=0.0, diagnostics=True):
    """The device pipeline in the order the driver runs it: K1 -> K2 -> K3."""
    lay = "kji" if layout == 0 else "ijk"
    slab = cpl.slab_reduce(d["vols"], layout=lay, ql_thresh=ql_thresh, want_mask=True)
    frc = cpl.gcm_to_les(d["gcm"], d["zf"], d["zh"], slab["prof"], d["aux"]["PS"], dt, f_les, True,
                         diagnostics=diagnostics, want_state=diagnostics, want_bracket=diagnostics)
    tnd = cpl.les_to_gcm(d["gcm"], d["zf"], d["zh"], slab, d["aux"], frc["slab_idx"], dt, f_gcm,
                         diagnostics=diagnostics)
    return slab, frc, tnd
