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
