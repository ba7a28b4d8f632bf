"""qt-variability nudging on the GPU path (spcpl.variability_nudge, splib/spcpl.py:613-744), the
experimental `--qt_forcing variance` option: same call shape as the reference for one LES, plus the
batched form over a gpu_les_batch. The random field R of the additive branch (spcpl.py:620-621) is
drawn with torch's device generator (the reference uses numpy's global Mersenne Twister; only the
distribution matches) and made exactly zero-mean per column, as the reference does."""
import logging

import torch

from . import spio
from .constants import LES_FIELDS


ST_ADD_FAIL, ST_NOCONV = 16, 32


def check_status(out, strict=False):
    """scipy.optimize.brentq raises when a root is not bracketed (spcpl.py:708) or the iteration limit is hit, which
    stops the reference run; the kernel records those cases per (column, level) in `status` and leaves the slab
    un-nudged. strict=True raises like the reference; the default logs a warning. Returns the number of affected
    slabs (also stored as out["failed"]). One device->host read of a scalar."""
    bad = int(((out["status"] & (ST_ADD_FAIL | ST_NOCONV)) != 0).sum().item())
    out["failed"] = bad
    if bad:
        msg = ("variability_nudge: %d (column, level) slabs where the root search failed (status bits 16 = not "
               "bracketed / no R, 32 = not converged); scipy.optimize.brentq raises there" % bad)
        if strict:
            raise RuntimeError(msg)
        logging.getLogger(__name__).warning(msg)
    return bad


def zero_mean_normal(ncol, ny, nx, device, generator=None):
    """R = N(0,1) on the horizontal plane with its mean removed (spcpl.py:620-621), one per column."""
    R = torch.randn((ncol, ny, nx), dtype=torch.float64, device=device, generator=generator)
    return (R - R.sum(dim=(1, 2), keepdim=True) / (ny * nx)).contiguous()


def variability_nudge_all(batch, DT, constantT=False, R=None, generator=None, strict=False):
    """Nudge every LES of the batch in one launch. Uses the batch's last slab means (K1) for
    <qt>, <ql> and its ql_ref (K2); qsat is the stand-in's horizontally uniform profile."""
    cpl, pipe = batch.cpl, batch.pipe
    if pipe.slab is None:
        pipe.les_profiles()
    if R is None:
        R = zero_mean_normal(batch.ncol, batch.ny, batch.nx, cpl.device, generator)
    qt = batch.vols[LES_FIELDS.index("QT")]
    kw = {}
    if constantT:
        kw = dict(thl=batch.vols[LES_FIELDS.index("THL")], ql=batch.vols[LES_FIELDS.index("QL")],
                  presf=batch.aux["presf"])
    out = cpl.variability_nudge(qt, pipe.slab["prof"], batch.ql_ref.to(batch.dtype).contiguous(), float(DT),
                                qsat_prof=batch.qsat.to(batch.dtype).contiguous(), R=R, constant_T=constantT, **kw)
    batch.state_version = getattr(batch, "state_version", 0) + 1       # qt (and thl) changed in place
    check_status(out, strict)
    return out


def variability_nudge(les, DT, constantT=False, write=True, R=None, strict=False):
    """Reference signature (spcpl.py:613) for one LES of a gpu_les_batch."""
    b, i = les.batch, les.i
    cpl = b.cpl
    slab = cpl.slab_reduce([v[i:i + 1] for v in b.vols], want_cnt=False, want_mask=False)
    if R is None:
        R = zero_mean_normal(1, b.ny, b.nx, cpl.device)
    kw = {}
    if constantT:
        kw = dict(thl=b.vols[LES_FIELDS.index("THL")][i:i + 1], ql=b.vols[LES_FIELDS.index("QL")][i:i + 1],
                  presf=b.aux["presf"][i:i + 1].contiguous())
    ql_ref = torch.as_tensor(les.ql_ref, device=cpl.device).reshape(1, -1).to(b.dtype).contiguous()
    out = cpl.variability_nudge(b.vols[LES_FIELDS.index("QT")][i:i + 1], slab["prof"], ql_ref, float(DT),
                                qsat_prof=b.qsat[i:i + 1].to(b.dtype).contiguous(), R=R.reshape(1, b.ny, b.nx),
                                constant_T=constantT, **kw)
    b.state_version = getattr(b, "state_version", 0) + 1
    check_status(out, strict)
    if write:                                                    # spcpl.py:742-744
        spio.write_les_data(les, qt_alpha=out["alpha"][0])
        spio.write_les_data(les, qt_beta=out["beta"][0], qt_std=out["qt_std"][0])
    return out
