"""Seeded synthetic GCM columns and LES-internal profiles for the GPU stand-in models (spdummy.py) and the
benchmarks: plain data generation with numpy, ~1k values per column, a column's data depending only
on its GLOBAL column index so that sharded and unsharded runs see identical inputs. Nothing here is on the timed path
and nothing here restates the reference's arithmetic: the host-side twin of the set_les_state kernel and the synthetic
LES volumes built around converted GCM profiles are test infrastructure and live in tests/synth_les.py.
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


def cloud_offset(nk):
    """qsat proxy offset s(k) in units of the qt noise amplitude: a cloud layer at k in
    [30, 80) where 10-30 % of the cells are saturated, no cloud elsewhere (exact zeros)."""
    k = np.arange(nk, dtype=np.float64)
    s = np.full(nk, 1.5)
    lo, hi = min(30, nk // 5), min(80, nk // 2)
    inside = (k >= lo) & (k < hi)
    s[inside] = 0.6 + 0.2 * np.cos(0.37 * k[inside])
    return s


# Philox stream ids of the five fields; QL re-uses QT's stream (same cell noise)
STREAM = {"THL": 0, "QT": 1, "U": 3, "V": 4}
