"""GPU-backed mirror of the reference's coupling utilities (splib/sputils.py), same names.

Quantities are plain SI torch tensors on the device (the reference's AMUSE units are all
SI-coherent on this path, so no conversion factor is lost). Every function launches a hand-written
kernel through the C ABI; there is no CPU path. 1-D arguments are treated as a batch of one row.
"""
import math

import torch

from .constants import pref0, rd, rv, cp, rlv, grav  # noqa: F401  (sputils.py:14-19)
from .coupler import default_coupler


def _cpl(t):
    return default_coupler(t.device)


def rms(a):
    """Root mean square (sputils.py:23-24). Host-side convenience, not on the hot path."""
    return math.sqrt(float((a.double() ** 2).mean()))


def exner(p):
    """(p/pref0)^(rd/cp)  (sputils.py:28-29)."""
    p = torch.as_tensor(p)
    return _cpl(p).exner(p.contiguous(), inverse=False)


def iexner(p):
    """(p/pref0)^(-rd/cp)  (sputils.py:33-34)."""
    p = torch.as_tensor(p)
    return _cpl(p).exner(p.contiguous(), inverse=True)


def interp(x, xp, fp, **kwargs):
    """numpy.interp semantics (sputils.py:82-86): clamped outside [xp[0], xp[-1]], xp increasing.
    x: [nx] or [nb, nx]; xp, fp: [np] or [nb, np]."""
    if kwargs:
        raise TypeError("interp: left/right/period are not supported on the GPU path: %s" % sorted(kwargs))
    one = xp.dim() == 1
    xp2 = xp.unsqueeze(0).contiguous() if one else xp.contiguous()
    fp2 = fp.unsqueeze(0).contiguous() if one else fp.contiguous()
    out = _cpl(xp).interp(x.contiguous().to(xp.dtype), xp2, fp2.to(xp.dtype))
    return out[0] if one and x.dim() == 1 else out


def searchsorted(a, v, side="left", **kwargs):
    """numpy.searchsorted semantics (sputils.py:88-91). a: [na] or [nb, na]; v: [nv] or [nb, nv]."""
    if kwargs:
        raise TypeError("searchsorted: unsupported arguments %s" % sorted(kwargs))
    one = a.dim() == 1
    a2 = a.unsqueeze(0).contiguous() if one else a.contiguous()
    v = torch.as_tensor(v, dtype=a.dtype, device=a.device)
    scalar = v.dim() == 0
    v2 = v.reshape(1) if scalar else v.contiguous()
    out = _cpl(a).searchsorted(a2, v2, side=side)
    out = out[0] if one and v2.dim() == 1 else out
    return out[..., 0] if scalar else out


def _rows(t, dtype=None):
    t = t if dtype is None else t.to(dtype)
    return (t.unsqueeze(0) if t.dim() == 1 else t).contiguous()


def integral(a, b, z, q, w=None):
    """Integral from a to b of the piecewise-constant q(z) (q[i] on [z[i], z[i+1]]), optionally the w-weighted mean
    (sputils.py:94-161). a, b are host scalars; returns a 0-d tensor, or None (with the reference's message) when an
    end point is outside [z[0], z[-1]]."""
    cpl = _cpl(z)
    z64 = z.to(torch.float64).contiguous()
    a, b = float(a), float(b)
    z0, z1 = float(z64[0]), float(z64[-1])
    if len(z) != len(q) + 1:                                                 # sputils.py:111-112
        print("len(z) should be len(q) + 1. len(z)=%d, len(q) = %d", (len(z), len(q)))
    if a < z0 or a > z1 or b < z0 or b > z1:                                 # sputils.py:113-115
        print("integral: Interval end point outside range.")
        return None
    sign = 1.0
    if a > b:                                                                # sputils.py:117-120
        sign, a, b = -1.0, b, a
    Zh = torch.tensor([[b, a]], dtype=q.dtype, device=z.device)
    out = cpl.interp_c(Zh, z64, _rows(q), None if w is None else _rows(w, q.dtype),
                       mode=cpl.INT_PLAIN if w is None else cpl.INT_WEIGHTED)
    return out[0, 0] * sign


def interp_c(Zh, zh, q, rho):
    """Conservative (mass-weighted) coarsening from the LES cells to the GCM layers (sputils.py:173-189): Zh descending
    layer edges [nlev+1] (or [nb, nlev+1]), zh ascending cell edges, q and rho cell values [nq] (or [nb, nq]).
    Q[i] = integral(rho q) / integral(rho) over [Zh[i+1], Zh[i]], 0 where Zh[i] >= zh[-1]."""
    cpl = _cpl(zh)
    one = Zh.dim() == 1
    out = cpl.interp_c(_rows(Zh, q.dtype), zh.to(torch.float64).contiguous(), _rows(q), _rows(rho, q.dtype), mode=cpl.INT_C)
    return out[0] if one else out


def interp_rho(Zh, zh, rho):
    """Layer-mean density on the coarser grid (sputils.py:191-197)."""
    cpl = _cpl(zh)
    one = Zh.dim() == 1
    out = cpl.interp_c(_rows(Zh, rho.dtype), zh.to(torch.float64).contiguous(), None, _rows(rho), mode=cpl.INT_RHO)
    return out[0] if one else out
