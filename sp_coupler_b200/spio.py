"""Diagnostics store with the reference's spifs.nc variable names (splib/spio.py:127-225).

netCDF4 is not available in this environment, so the same per-column groups / per-step records
are kept in memory and written as one .npz (`<group>/<variable>` keys, leading Time dimension), which
examples-style readers can open with numpy. Same call names as the reference module:
init_netcdf, update_time, write_les_data, write_netCDF_data, sync_root, close.
Values may be device tensors; they are copied to the host on write (diagnostics only, off the
timed path).
"""
import threading

import numpy as np

cdf_lock = threading.Lock()          # spio.py:24
_store = {}                          # group -> variable -> {step index: value}
_times = []
_file = None


def _host(v):
    try:
        import torch
        if isinstance(v, torch.Tensor):
            return v.detach().cpu().numpy()
    except ImportError:
        pass
    return np.asarray(v)


def init_netcdf(nc_name="spifs.npz", gcm=None, les_models=(), start_time=None, append=False, with_satdata=False):
    """spio.py:37-73: one group per LES column, keyed by its grid index."""
    global _file
    _file = nc_name
    if not append:
        _store.clear()
        del _times[:]
    for les in les_models:
        _store.setdefault(str(les.grid_index), {})
        les.cdf = _store[str(les.grid_index)]       # spio.py:57


def update_time(t):
    """spio.py:96-124: open a new record."""
    _times.append(float(_host(t)))


def _write(group, kwargs):
    step = max(len(_times) - 1, 0)
    g = _store.setdefault(group, {})
    for name, value in kwargs.items():
        g.setdefault(name, {})[step] = _host(value)


def write_les_data(les, lock=False, **kwargs):
    """spio.py:228-242."""
    if lock:
        with cdf_lock:
            _write(str(les.grid_index), kwargs)
    else:
        _write(str(les.grid_index), kwargs)


def write_netCDF_data(index, **kwargs):
    """spio.py:248-258 (extra GCM output columns)."""
    _write("col%s" % index, kwargs)


def get(group, name, step=-1):
    rec = _store[str(group)][name]
    return rec[sorted(rec)[step]]


def sync_root():
    """spio.py:76-84: flush to disk."""
    if _file is None:
        return
    with cdf_lock:
        out = {"Time": np.asarray(_times)}
        for g, vs in _store.items():
            for name, rec in vs.items():
                steps = sorted(rec)
                try:
                    out["%s/%s" % (g, name)] = np.stack([np.asarray(rec[s]) for s in steps])
                except ValueError:
                    out["%s/%s" % (g, name)] = np.asarray([rec[s] for s in steps], dtype=object)
        np.savez(_file, **out)


def close():
    sync_root()
