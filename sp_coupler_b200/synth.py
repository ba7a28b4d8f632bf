"""Seeded synthetic inputs for the coupling path.

GCM columns are generated on the host with numpy (they are ~1k values per column).
LES volumes are `profile[k] + amp * uniform(-1, 1)` exactly as the reference's
``set_les_state`` builds its initial state (splib/spcpl.py:274-291: amplitudes
0.5 m/s, 0.1 K, 2.5e-5 kg/kg), with the noise drawn from a counter-based
Philox4x32-10 stream keyed by (seed, field) and counted by (element/4, column),
so that

* the CUDA generator (csrc/les_state.cu) and the numpy generator below produce
  bit-identical volumes (parity-sized cases are generated on the host, the big
  bench configs on the device, and full-size tests spot-check one against the other);
* a column's data depends only on its GLOBAL column index, so sharded and
  unsharded runs see identical inputs.

Nothing here is on the timed path.
"""
import numpy as np

from . import constants as C

# hybrid A/B half-level coefficients of the OpenIFS T21 L19 test case
# (GRIB GDS of oifs-input/ICMGGTESTINIUA, listed in SURVEY.md §8d)
T21_A = np.array([0, 2000, 4000, 6046.109375, 8267.92578125, 10609.51171875, 12851.1015625,
                  14698.5, 15861.125, 16116.23828125, 15356.92578125, 13621.4609375, 11101.5625,
                  8127.14453125, 5125.140625, 2549.969482421875, 783.195068359375, 0, 0, 0],
                 dtype=np.float64)
T21_B = np.array([0, 0, 0, 0.0003389932680875063, 0.0033571866806596518, 0.013070043176412582,
                  0.03407714515924454, 0.07064980268478394, 0.12591665983200073,
                  0.20119541883468628, 0.2955196499824524, 0.40540921688079834,
                  0.5249322056770325, 0.6461079716682434, 0.7596983909606934,
                  0.8564375638961792, 0.9287469387054443, 0.9729852080345154,
                  0.9922814965248108, 1], dtype=np.float64)

NOISE_AMP = {"U": 0.5, "V": 0.5, "THL": 0.1, "QT": 2.5e-5}  # spcpl.py:285-287
SCALE_HEIGHT = 7500.0


def les_grid(nk=160, dz=25.0):
    """DALES vertical grid of dales-input/prof.inp.001: zf = dz/2 + dz*k, zh = dz*k."""
    k = np.arange(nk, dtype=np.float64)
    return dz * 0.5 + dz * k, dz * k


def make_gcm_columns(ncol, nlev, seed=42, dtype=np.float64, col0=0, ncol_total=None):
    """GCM state for columns [col0, col0+ncol) of a `ncol_total`-column run.

    Arrays are top -> bottom like OpenIFS (index 0 = model top, spcpl.py:197,241).
    Returns a dict with the reference's gcm_vars / surf_vars names (spcpl.py:32-33),
    cast to `dtype` (the oracle is always fed these already-rounded values).
    """
    ntot = ncol_total if ncol_total is not None else col0 + ncol
    rng = np.random.default_rng(seed)
    ps = rng.uniform(95e3, 103e3, ntot)
    zs = rng.uniform(0.0, 5000.0, ntot)           # orography geopotential, m^2/s^2
    if nlev == 19:
        ph = T21_A[None, :] + T21_B[None, :] * ps[:, None]
        zh = SCALE_HEIGHT * np.log(ps[:, None] / np.maximum(ph, 50.0))
    else:
        s = np.linspace(1.0, 0.0, nlev + 1)
        zh = np.broadcast_to(80e3 * (0.15 * s + 0.85 * s ** 3.5), (ntot, nlev + 1)).copy()
        ph = ps[:, None] * np.exp(-zh / SCALE_HEIGHT)
    zh[:, -1] = 0.0
    zf = 0.5 * (zh[:, 1:] + zh[:, :-1])
    pf = ps[:, None] * np.exp(-zf / SCALE_HEIGHT)
    g = {}
    g["Phalf"] = ph
    g["Pfull"] = pf
    g["Zghalf"] = C.grav * zh + zs[:, None]
    g["Zgfull"] = C.grav * zf + zs[:, None]
    g["T"] = np.maximum(288.15 - 6.5e-3 * zf, 216.65) + rng.normal(0.0, 1.0, (ntot, nlev))
    g["SH"] = 0.015 * (pf / ps[:, None]) ** 3
    g["QL"] = 1e-5 * rng.uniform(0.0, 1.0, (ntot, nlev))
    g["QI"] = 1e-6 * rng.uniform(0.0, 1.0, (ntot, nlev))
    g["U"] = 5.0 + rng.normal(0.0, 1.0, (ntot, nlev))
    g["V"] = rng.normal(0.0, 1.0, (ntot, nlev))
    g["A"] = rng.uniform(0.0, 1.0, (ntot, nlev))
    g["Z0M"] = rng.uniform(1e-4, 1.0, ntot)
    g["Z0H"] = rng.uniform(1e-5, 0.1, ntot)
    g["QLflux"] = -rng.uniform(0.0, 1e-6, ntot)
    g["QIflux"] = -rng.uniform(0.0, 1e-7, ntot)
    g["SHflux"] = -rng.uniform(0.0, 1e-4, ntot)
    g["TLflux"] = rng.uniform(-300.0, 20.0, ntot)
    g["TSflux"] = rng.uniform(-200.0, 50.0, ntot)
    sl = slice(col0, col0 + ncol)
    return {k: np.ascontiguousarray(v[sl]).astype(dtype) for k, v in g.items()}


def make_les_aux(ncol, nk, seed=42, dtype=np.float64, col0=0, ncol_total=None):
    """LES-internal profiles the reference fetches but does not compute on the path
    (spcpl.py:750-759): presf, Rhof, Rhobf, QL_ice, QR, T, PS, Rain."""
    ntot = ncol_total if ncol_total is not None else col0 + ncol
    rng = np.random.default_rng(seed + 1000003)
    zf, _ = les_grid(nk)
    ps = rng.uniform(95e3, 103e3, ntot)
    presf = ps[:, None] * np.exp(-zf[None, :] / SCALE_HEIGHT)
    t = 288.15 - 6.5e-3 * zf[None, :] + rng.normal(0.0, 0.5, (ntot, nk))
    rho = presf / (C.rd * t)
    a = {"presf": presf, "Rhof": rho, "Rhobf": rho * (1 + 1e-3 * rng.normal(0, 1, (ntot, nk))),
         "QL_ice": 2e-6 * rng.uniform(0, 1, (ntot, nk)), "QR": 1e-7 * rng.uniform(0, 1, (ntot, nk)),
         "T": t, "PS": ps, "Rain": rng.uniform(0, 1e-3, ntot)}
    sl = slice(col0, col0 + ncol)
    return {k: np.ascontiguousarray(v[sl]).astype(dtype) for k, v in a.items()}


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


def cloud_offset(nk):
    """qsat proxy offset s(k) in units of the qt noise amplitude: a cloud layer at k in
    [30, 80) where 10-30 % of the cells are saturated, no cloud elsewhere (exact zeros)."""
    k = np.arange(nk, dtype=np.float64)
    s = np.full(nk, 1.5)
    lo, hi = min(30, nk // 5), min(80, nk // 2)
    inside = (k >= lo) & (k < hi)
    s[inside] = 0.6 + 0.2 * np.cos(0.37 * k[inside])
    return s


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


# Philox stream ids of the five fields; QL re-uses QT's stream (same cell noise)
STREAM = {"THL": 0, "QT": 1, "U": 3, "V": 4}


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
