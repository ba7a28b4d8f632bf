"""Synthetic LES volumes for tests, smoke() and bench.py - TEST / BENCH INFRASTRUCTURE, not product.

LES volumes are `profile[k] + amp * uniform(-1, 1)` exactly as the reference's ``set_les_state`` builds its initial
state (splib/spcpl.py:274-291: amplitudes 0.5 m/s, 0.1 K, 2.5e-5 kg/kg), with the noise drawn from a counter-based
Philox4x32-10 stream keyed by (seed, field) and counted by (element/4, column), so that

* the CUDA generator (sp_coupler_b200/csrc/les_state.cu) and the numpy generator below produce bit-identical volumes
  (parity-sized cases are generated on the host, the big bench configs on the device, and full-size tests check one
  against the other) - the numpy generator is the CHECKER of that kernel;
* a column's data depends only on its GLOBAL column index, so sharded and unsharded runs see identical inputs.

The volumes are centred on the GCM state interpolated to the LES levels; `les_target_profiles` computes that with
numpy (a restatement of spcpl.convert_profiles, spcpl.py:171-246) - which is why this module lives under tests/ and
not in the product package. Nothing here is on a timed path.
"""
import numpy as np

from sp_coupler_b200 import constants as C
from sp_coupler_b200.synth import NOISE_AMP, STREAM, cloud_offset

# ----------------------------------------------------------------------------- Philox4x32-10
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_LO = np.uint64(0xFFFFFFFF)


def philox4x32(c0, c1, c2, c3, k0, k1, rounds=10):
    """Vectorised Philox4x32 (Salmon et al. 2011). Counters are uint32 arrays, keys ints."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32).astype(np.uint64) for c in np.broadcast_arrays(c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(rounds):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _LO
        hi1, lo1 = p1 >> np.uint64(32), p1 & _LO
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def les_noise(stream, col, nelem, seed=42):
    """noise in [-1, 1) for elements 0..nelem-1 of (stream, global column `col`), float64."""
    ngrp = (nelem + 3) // 4
    g = np.arange(ngrp, dtype=np.uint64)
    r = philox4x32((g & _LO).astype(np.uint32), (g >> np.uint64(32)).astype(np.uint32),
                   np.uint32(col), np.uint32(stream), seed, 0x5BD1E995)
    x = np.stack(r, axis=1).reshape(-1)[:nelem]
    u = (x >> np.uint32(8)).astype(np.float64) * (1.0 / 16777216.0)
    return 2.0 * u - 1.0


def les_state_volume(prof, amp, stream, nx, ny, seed=42, col0=0, sub=None, clamp0=False,
                     dtype=np.float32):
    """profile -> volume broadcast with noise: out[c,k,j,i] = prof[c,k] + amp*noise - sub[c,k],
    clamped at 0 when `clamp0`. Layout [ncol][nk][ny][nx] (slab-contiguous).
    Bit-identical to csrc/les_state.cu (fp64 arithmetic, no FMA, then one rounding to dtype)."""
    prof = np.asarray(prof, dtype=np.float64)
    ncol, nk = prof.shape
    out = np.empty((ncol, nk, ny, nx), dtype=dtype)
    for c in range(ncol):
        n = les_noise(stream, col0 + c, nk * ny * nx, seed).reshape(nk, ny * nx)
        v = prof[c][:, None] + amp * n
        if sub is not None:
            v = v - np.asarray(sub, dtype=np.float64)[c][:, None]
        if clamp0:
            v = np.maximum(v, 0.0)
        out[c] = v.reshape(nk, ny, nx).astype(dtype)
    return out


def les_target_profiles(gcm, zf):
    """GCM state interpolated to the LES levels (what spcpl.convert_profiles returns,
    spcpl.py:171-246), used as the centre of the synthetic LES volumes. float64."""
    g = {k: np.asarray(v, dtype=np.float64) for k, v in gcm.items()}
    ncol = g["T"].shape[0]
    Zf = (g["Zgfull"] - g["Zghalf"][:, -1:]) / C.grav
    thl_ = (g["T"] - (C.rlv * (g["QL"] + g["QI"])) / C.cp) * (g["Pfull"] / C.pref0) ** (-C.rd / C.cp)
    qt_ = g["SH"] + g["QL"] + g["QI"]
    out = {n: np.empty((ncol, len(zf))) for n in ("THL", "QT", "U", "V")}
    for c in range(ncol):
        xp = Zf[c, ::-1]
        out["THL"][c] = np.interp(zf, xp, thl_[c, ::-1])
        out["QT"][c] = np.interp(zf, xp, qt_[c, ::-1])
        out["U"][c] = np.interp(zf, xp, g["U"][c, ::-1])
        out["V"][c] = np.interp(zf, xp, g["V"][c, ::-1])
    return out


def les_volume_plan(gcm, zf, col0=0):
    """Per field: (profile, amp, stream, sub, clamp0) describing the synthetic LES state.
    The LES mean state is the GCM target plus a small smooth drift so that the forcings
    (gcm - les)/dt are non-trivial."""
    tgt = les_target_profiles(gcm, zf)
    ncol, nk = tgt["THL"].shape
    z = np.asarray(zf)[None, :] / 4000.0
    ph = 2 * np.pi * (((col0 + np.arange(ncol)) * 0.6180339887498949) % 1.0)[:, None]
    drift = np.sin(2 * np.pi * z + ph)
    prof = {"THL": tgt["THL"] + 0.3 * drift, "QT": tgt["QT"] * (1 + 0.02 * drift),
            "U": tgt["U"] + 0.5 * drift, "V": tgt["V"] - 0.4 * drift}
    qsat = prof["QT"] + NOISE_AMP["QT"] * cloud_offset(nk)[None, :]
    plan = {
        "THL": (prof["THL"], NOISE_AMP["THL"], STREAM["THL"], None, False),
        "QT": (prof["QT"], NOISE_AMP["QT"], STREAM["QT"], None, False),
        "QL": (prof["QT"], NOISE_AMP["QT"], STREAM["QT"], qsat, True),
        "U": (prof["U"], NOISE_AMP["U"], STREAM["U"], None, False),
        "V": (prof["V"], NOISE_AMP["V"], STREAM["V"], None, False),
    }
    return plan


def make_les_volumes(gcm, zf, nx, ny, seed=42, dtype=np.float32, col0=0):
    """Host (numpy) LES volumes for parity-sized cases: dict field -> [ncol][nk][ny][nx]."""
    plan = les_volume_plan(gcm, zf, col0=col0)
    return {f: les_state_volume(p, amp, st, nx, ny, seed=seed, col0=col0, sub=sub, clamp0=cl, dtype=dtype)
            for f, (p, amp, st, sub, cl) in plan.items()}


def device_les_volumes(cpl, gcm, zf, nx, ny, seed=42, dtype=None, col0=0):
    """The same synthetic LES volumes as make_les_volumes(), generated on the device by the
    spc_set_les_state kernel (bit-identical; used for configs too large to build on the host).
    Returns the five [ncol][nk][ny][nx] tensors in LES_FIELDS order."""
    import torch
    dtype = dtype if dtype is not None else torch.float32
    plan = les_volume_plan(gcm, zf, col0=col0)
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(cpl.device)
    out = []
    for f in C.LES_FIELDS:
        prof, amp, stream, sub, clamp0 = plan[f]
        out.append(cpl.set_les_state(up(prof), amp, stream, nx, ny, seed=seed, col0=col0,
                                     sub=None if sub is None else up(sub), clamp0=clamp0, dtype=dtype))
    return out
