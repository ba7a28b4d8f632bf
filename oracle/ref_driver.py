"""Runs the UNMODIFIED reference coupler (splib/{spcpl,sputils}.py) per column under the unit shim in
oracle/stubs — TEST / BASELINE INFRASTRUCTURE ONLY.

The reference modules are imported from /root/reference where that tree exists (the build container), else from
the verbatim copy oracle/make_ref.py placed in the git-ignored oracle/_ref/ (which travels to the GPU box). Used to
(1) validate oracle/numpy_batched.py, (2) generate the golden vectors committed under tests/golden/
(oracle/make_golden.py), and (3) time the reference's own per-column coupling functions on the host cores
(bench.py --impl reference and its cpu_baseline leg: reference_column_steps below). The product never imports it.
"""
import contextlib
import io
import os
import sys
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_STUBS = os.path.join(_HERE, "stubs")


def _reference_root():
    for root in (os.environ.get("SPC_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "_ref")):
        if root and os.path.isfile(os.path.join(root, "splib", "spcpl.py")):
            return root
    return os.environ.get("SPC_REFERENCE_ROOT", "/root/reference")


REFERENCE_ROOT = _reference_root()


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "splib", "spcpl.py"))


def load_reference():
    """Import splib.{sputils,spcpl,spdummy,spio} from the reference tree, unmodified."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    for p in (REFERENCE_ROOT, _STUBS):
        if p not in sys.path:
            sys.path.insert(0, p)
    from splib import sputils, spcpl, spdummy, spio  # noqa
    return sputils, spcpl, spdummy, spio


class _Les(object):
    """Bare LES object with what spcpl reads/calls on the path."""

    def __init__(self, idx):
        self.grid_index = idx
        self.sent = {}

    def _setter(name):
        def f(self, v, return_request=False):
            self.sent[name] = np.array(v, dtype=np.float64)
            return None
        return f

    set_tendency_U = _setter("f_u")
    set_tendency_V = _setter("f_v")
    set_tendency_THL = _setter("f_thl")
    set_tendency_QT = _setter("f_qt")
    set_tendency_surface_pressure = _setter("f_ps")
    set_tendency_QL = _setter("f_ql")
    set_ref_profile_QL = _setter("ql_ref")
    set_z0m_surf = _setter("z0m")
    set_z0h_surf = _setter("z0h")
    set_wt_surf = _setter("wthl")
    set_wq_surf = _setter("wqt")


class _Gcm(object):
    def __init__(self):
        self.tend = {}

    def set_profile_tendency(self, name, idx, vals):
        self.tend.setdefault(idx, {})[name] = np.array(vals, dtype=np.float64)


def run_columns(gcm, zf, zh, les_prof, aux, A_les, dt, f_les, f_gcm, couple_surface=True,
                conservative=False):
    """Drive reference set_les_forcings + set_gcm_tendencies for every column.

    gcm: dict of [ncol, ...] float arrays (gcm_vars + surf_vars); les_prof: dict U,V,THL,QT,QL
    [ncol, nk]; aux: dict presf,Rhof,Rhobf,QL_ice,QR,T [ncol,nk], PS, Rain [ncol]; A_les: [ncol,
    nlev] cloud fraction in ascending slab order (profile["A"]). Returns dict of stacked outputs,
    including every diagnostic the reference hands to spio.write_les_data."""
    sputils, spcpl, spdummy, spio = load_reference()
    from _q import Q
    ncol = gcm["T"].shape[0]
    captured = {}
    spio.write_les_data = lambda les, **kw: captured.update(kw)
    rows = []
    for c in range(ncol):
        captured.clear()
        les = _Les(c)
        for v in spcpl.gcm_vars:
            setattr(les, v, Q(np.asarray(gcm[v][c], dtype=np.float64)))
        for v in spcpl.surf_vars:
            setattr(les, v, Q(np.float64(gcm[v][c])))
        les.zf_cache = Q(np.asarray(zf, dtype=np.float64))
        les.zh_cache = Q(np.asarray(zh, dtype=np.float64))
        les.rain = Q(0.0)
        prof = {k: Q(np.asarray(les_prof[k][c], dtype=np.float64)) for k in ("U", "V", "THL", "QT", "QL")}
        for k in ("presf", "Rhof", "Rhobf", "QL_ice", "QR", "T"):
            prof[k] = Q(np.asarray(aux[k][c], dtype=np.float64))
        prof["PS"] = Q(np.float64(aux["PS"][c]))
        prof["Rain"] = Q(np.float64(aux["Rain"][c]))
        prof["A"] = Q(np.asarray(A_les[c], dtype=np.float64))
        g = _Gcm()
        spcpl.set_les_forcings(les, None, False, False, prof, dt, f_les, couple_surface, write=True)
        # cloud-fraction slab mapping exactly as get_les_profiles builds it (spcpl.py:761-764)
        idx = sputils.searchsorted(les.zh_cache, les.gcm_Zh, side="right")[:-1:][::-1]
        # sputils.integral prints a length warning per call when len(zh) == len(q) (sputils.py:111)
        with contextlib.redirect_stdout(io.StringIO()):
            spcpl.set_gcm_tendencies(g, les, prof, dt, f_gcm, write=True, conservative=conservative)
        row = {k: np.array(v, dtype=np.float64) for k, v in captured.items()}
        row.update({k: v for k, v in les.sent.items()})
        row.update({"f_" + k: v for k, v in g.tend[c].items()})
        row["slab_idx"] = np.asarray(idx, dtype=np.int32)
        row["gcm_Zf"] = np.array(les.gcm_Zf, dtype=np.float64)
        row["gcm_Zh"] = np.array(les.gcm_Zh, dtype=np.float64)
        row["start_index"] = np.int32(sputils.searchsorted(-les.gcm_Zf, -les.zf_cache[-1]))  # spcpl.py:498
        rows.append(row)
    keys = rows[0].keys()
    return {k: np.stack([r[k] for r in rows]) for k in keys}


class _NudgeLes(object):
    """LES object with what spcpl.variability_nudge reads (spcpl.py:616-636, 657-658, 732-734)."""

    class _Fields(object):
        pass

    class _Domain(object):
        pass

    def __init__(self, fields, profiles, presf, ql_ref, Q):
        self._f, self._p, self._presf, self._Q = fields, profiles, presf, Q
        self.ql_ref = Q(ql_ref)
        self.fields = self._Fields()
        self.parameters_DOMAIN = self._Domain()
        self.parameters_DOMAIN.kmax = fields["QT"].shape[2]
        self.grid_index = 0

    def get_itot(self):
        return self._f["QT"].shape[0]

    def get_jtot(self):
        return self._f["QT"].shape[1]

    def get_field(self, name):
        return self._Q(self._f[name].copy())

    def get_profile(self, name):
        return self._Q(self._p[name].copy())

    def get_presf(self):
        return self._Q(self._presf.copy())


def run_variability_nudge(fields, profiles, presf, ql_ref, DT, constantT, seed):
    """Drive the UNMODIFIED reference spcpl.variability_nudge. fields: dict Qsat, QT, THL, QL of
    (itot, jtot, ktot) float64 arrays (the OMUSE view); profiles: dict QL, QT [ktot].
    numpy's global RNG is seeded with `seed` right before the call, so R (spcpl.py:620-621) can be
    regenerated by the caller with the same seed. Returns dict(qt, thl, beta, alpha, qt_std, R)."""
    sputils, spcpl, spdummy, spio = load_reference()
    from _q import Q
    captured = {}
    spio.write_les_data = lambda les, **kw: captured.update(kw)
    les = _NudgeLes(fields, profiles, presf, ql_ref, Q)
    itot, jtot = les.get_itot(), les.get_jtot()
    np.random.seed(seed)
    R = np.random.normal(size=(itot, jtot))
    R -= R.sum() / (itot * jtot)
    np.random.seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        spcpl.variability_nudge(les, Q(np.float64(DT)), constantT, write=True)
    out = dict(qt=np.array(les.fields.QT, dtype=np.float64), R=R,
               beta=np.array(captured["qt_beta"], dtype=np.float64),
               alpha=np.array(captured["qt_alpha"], dtype=np.float64),
               qt_std=np.array(captured["qt_std"], dtype=np.float64))
    if constantT:
        out["thl"] = np.array(les.fields.THL, dtype=np.float64)
    return out


# ------------------------------------------------------------------------------------------------ timing
def reference_column_steps(gcm, zf, zh, vols, aux, dt, f_les, f_gcm, couple_surface=True, ql_thresh=0.0):
    """BASELINE.md §2 "B-ref": the coupling step of every column of the block, serially, as splib.step runs it
    (splib.py:317-332) - numpy slab means of the five volumes + above-threshold ql count and projected cloud cover
    (the LES worker's part, spcpl.py:748-765), then the UNMODIFIED spcpl.set_les_forcings and
    spcpl.set_gcm_tendencies with write=False. vols: dict field -> [ncol][nk][ny][nx] in the storage dtype.
    Returns the packed tendencies [ncol][7][nlev] (so the work cannot be optimised away and can be checked)."""
    sputils, spcpl, spdummy, spio = load_reference()
    from _q import Q
    spio.write_les_data = lambda les, **kw: None          # convert_profiles is called with its default write=True
    ncol, nlev = gcm["T"].shape
    nk = len(zf)
    zfq, zhq = Q(np.asarray(zf, dtype=np.float64)), Q(np.asarray(zh, dtype=np.float64))
    out = np.empty((ncol, 7, nlev))
    names = ("T", "SH", "QL", "QI", "U", "V", "A")
    devnull = io.StringIO()
    for c in range(ncol):
        les = _Les(c)
        for v in spcpl.gcm_vars:                          # gather_gcm_data's per-LES attributes (spcpl.py:81-86)
            setattr(les, v, Q(np.asarray(gcm[v][c], dtype=np.float64)))
        for v in spcpl.surf_vars:
            setattr(les, v, Q(np.float64(gcm[v][c])))
        les.zf_cache, les.zh_cache = zfq, zhq
        les.rain = Q(0.0)
        prof = {}
        for f in ("U", "V", "THL", "QT", "QL"):           # les.get_profile_*: horizontal slab means
            prof[f] = Q(vols[f][c].reshape(nk, -1).mean(axis=1).astype(np.float64))
        cloudy = vols["QL"][c].reshape(nk, -1) > ql_thresh
        for k in ("presf", "Rhof", "Rhobf", "QL_ice", "QR", "T"):
            prof[k] = Q(np.asarray(aux[k][c], dtype=np.float64))
        prof["PS"] = Q(np.float64(aux["PS"][c]))
        prof["Rain"] = Q(np.float64(aux["Rain"][c]))
        g = _Gcm()
        spcpl.set_les_forcings(les, None, False, False, prof, dt, f_les, couple_surface, write=False)
        idx = np.asarray(sputils.searchsorted(les.zh_cache, les.gcm_Zh, side="right")[:-1:][::-1])   # spcpl.py:761-764
        A = np.zeros(nlev)                                # les.get_cloudfraction(indices): projected cover per slab
        k0 = 0
        for r in range(nlev):
            k1 = min(max(int(idx[r]), k0), nk)
            if k1 > k0:
                A[r] = np.count_nonzero(cloudy[k0:k1].any(axis=0)) / cloudy.shape[1]
            k0 = k1
        prof["A"] = Q(A)
        with contextlib.redirect_stdout(devnull):
            spcpl.set_gcm_tendencies(g, les, prof, dt, f_gcm, write=False)
        for n, name in enumerate(names):
            out[c, n] = g.tend[c][name]
    return out
