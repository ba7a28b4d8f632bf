"""CPU oracle of the per-column coupling step — TEST INFRASTRUCTURE, NOT PRODUCT.

A unit-free float64 numpy restatement, batched over [ncol, ...] arrays, of the
reference path splib/spcpl.py + splib/sputils.py (file:line cited per function).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package; the product (sp_coupler_b200/) never does.

Parity status
  * Profile math (convert_profiles, set_les_forcings, convert_surface_fluxes,
    set_gcm_tendencies, cloud-fraction index mapping, exner/iexner, integral / interp_c /
    interp_rho): PINNED against
    (a) the reference's own known-answer tests (splib/test/sputils_test.py:25-39,
    splib/test/spcpl_test.py:10-16) and (b) golden vectors produced by running the
    UNMODIFIED reference functions from /root/reference under oracle/stubs
    (oracle/make_golden.py -> tests/golden/ref_*.npz); see tests/test_oracle.py.
  * Slab averages and cloud-fraction COUNTS: the reference delegates these to the
    external DALES worker (call sites spcpl.py:748-766); nothing under /root/reference
    computes them and no reference test pins them -> "parity unpinned" for these two.
    The contract is BASELINE.json's north_star: horizontal mean; cloud fraction = count of
    ql > threshold.  Definitions are in slab_reduce() / cloud_project() below. (The only
    in-tree statement of a slab average, `.sum() / (itot * jtot)` over field[:, :, k] at
    spcpl.py:642,650, is the same definition.)

Orientation: GCM arrays run top -> bottom (index 0 = model top; half-level arrays have
nlev+1 entries ending at the ground); LES arrays bottom -> top.
"""
import numpy as np

pref0 = 1.0e5   # sputils.py:14
rd = 287.04     # sputils.py:15
rv = 461.5      # sputils.py:16
cp = 1004.0     # sputils.py:17
rlv = 2.53e6    # sputils.py:18
grav = 9.81     # sputils.py:19

gcm_vars = ["U", "V", "T", "SH", "QL", "QI", "Pfull", "Phalf", "A", "Zgfull", "Zghalf"]  # spcpl.py:32
surf_vars = ["Z0M", "Z0H", "QLflux", "QIflux", "SHflux", "TLflux", "TSflux"]               # spcpl.py:33


def exner(p):
    """sputils.py:28-29"""
    return (np.asarray(p, dtype=np.float64) / pref0) ** (rd / cp)


def iexner(p):
    """sputils.py:33-34"""
    return (np.asarray(p, dtype=np.float64) / pref0) ** (-rd / cp)


def bracket(x, xp):
    """Bracketing index used by numpy.interp (sputils.py:82-86 -> numpy.interp):
    j = upper_bound(xp, x) - 1 in [-1, n-1]; j == -1 -> left clamp, j >= n-1 -> right clamp."""
    return (np.searchsorted(xp, x, side="right") - 1).astype(np.int32)


def interp(x, xp, fp):
    """sputils.py:82-86 with units stripped."""
    return np.interp(x, xp, fp)


def _f64(d):
    return {k: np.asarray(v, dtype=np.float64) for k, v in d.items()}


def convert_profiles(gcm, zf):
    """spcpl.py:171-246, batched. `gcm`: dict of [ncol, nlev(+1)] arrays, `zf`: [nk] LES full
    levels (les.zf_cache). Returns dict with Tv, Zh, Zf, THL(=thl_), QT(=qt_) on GCM levels,
    thl, qt, ql, u, v on LES levels, ps, and the bracket indices of the 5 interpolations."""
    g = _f64(gcm)
    zf = np.asarray(zf, dtype=np.float64)
    c = rv / rd - 1                                                   # :175
    Tv = g["T"] * (1 + c * g["SH"] - (g["QL"] + g["QI"]))             # :176
    Zh = (g["Zghalf"] - g["Zghalf"][:, -1:]) / grav                   # :197
    Zf = (g["Zgfull"] - g["Zghalf"][:, -1:]) / grav                   # :198
    thl_ = (g["T"] - (rlv * (g["QL"] + g["QI"])) / cp) * iexner(g["Pfull"])  # :214
    qt_ = g["SH"] + g["QL"] + g["QI"]                                 # :215
    ncol, nk = Zf.shape[0], zf.shape[0]
    out = {n: np.empty((ncol, nk)) for n in ("thl", "qt", "ql", "u", "v")}
    br = np.empty((ncol, nk), dtype=np.int32)
    for i in range(ncol):
        xp = Zf[i, ::-1]                                              # :224-228 (reversed)
        out["thl"][i] = np.interp(zf, xp, thl_[i, ::-1])
        out["qt"][i] = np.interp(zf, xp, qt_[i, ::-1])
        out["ql"][i] = np.interp(zf, xp, g["QL"][i, ::-1])
        out["u"][i] = np.interp(zf, xp, g["U"][i, ::-1])
        out["v"][i] = np.interp(zf, xp, g["V"][i, ::-1])
        br[i] = bracket(zf, xp)
    out.update(Tv=Tv, Zh=Zh, Zf=Zf, THL=thl_, QT=qt_, ps=g["Phalf"][:, -1], bracket=br)  # :246
    return out


def convert_surface_fluxes(gcm):
    """spcpl.py:136-167."""
    g = _f64(gcm)
    ps = g["Phalf"][:, -1]
    rho = ps / (rd * g["T"][:, -1])                                   # :153
    wqt = -(g["QLflux"] + g["QIflux"] + g["SHflux"]) / rho            # :159
    wthl = -g["TSflux"] * iexner(ps) / (cp * rho)                     # :161
    return g["Z0M"], g["Z0H"], wthl, wqt                              # :167


def set_les_forcings(gcm, zf, les_prof, ps_les, dt_gcm, factor, couple_surface):
    """spcpl.py:299-385 (the arithmetic; RPC sets and rain bookkeeping excluded).
    `les_prof`: dict U,V,THL,QT,QL -> [ncol, nk] slab means; `ps_les` [ncol]."""
    cv = convert_profiles(gcm, zf)
    p = _f64(les_prof)
    out = dict(cv)
    out["f_u"] = factor * (cv["u"] - p["U"]) / dt_gcm                 # :328
    out["f_v"] = factor * (cv["v"] - p["V"]) / dt_gcm                 # :329
    out["f_thl"] = factor * (cv["thl"] - p["THL"]) / dt_gcm           # :330
    out["f_qt"] = factor * (cv["qt"] - p["QT"]) / dt_gcm              # :331
    out["f_ps"] = factor * (cv["ps"] - np.asarray(ps_les, dtype=np.float64)) / dt_gcm  # :332
    out["f_ql"] = factor * (cv["ql"] - p["QL"]) / dt_gcm              # :333
    out["ql_ref"] = cv["ql"]                                          # :347-348
    if couple_surface:
        out["z0m"], out["z0h"], out["wthl"], out["wqt"] = convert_surface_fluxes(gcm)  # :360
    return out


def slab_indices(zh, Zh):
    """Cloud-fraction slab mapping, spcpl.py:26 and :764:
    searchsorted(zh, Zh, side='right')[:-1][::-1], batched over columns -> int32 [ncol, nlev]."""
    zh = np.asarray(zh, dtype=np.float64)
    Zh = np.asarray(Zh, dtype=np.float64)
    return np.stack([np.searchsorted(zh, Zh[i], side="right")[:-1][::-1] for i in range(Zh.shape[0])]).astype(np.int32)


def integral(a, b, z, q, w):
    """sputils.py:94-161 (weighted branch, :151-161)."""
    if a < z[0] or a > z[-1] or b < z[0] or b > z[-1]:                # :113-115
        return None
    sign = 1
    if a > b:                                                         # :118-120
        sign = -1
        a, b = b, a
    ia = 0
    while z[ia + 1] < a:                                              # :123-124
        ia += 1
    ib = ia
    while z[ib + 1] < b:                                              # :126-127
        ib += 1
    S = (w[ia:ib + 1] * q[ia:ib + 1] * (z[ia + 1:ib + 2] - z[ia:ib + 1])).sum()   # :152
    Sa = w[ia] * q[ia] * (a - z[ia])                                  # :154
    Sb = w[ib] * q[ib] * (z[ib + 1] - b)                              # :155
    Sw = (w[ia:ib + 1] * (z[ia + 1:ib + 2] - z[ia:ib + 1])).sum()     # :157
    Swa = w[ia] * (a - z[ia])                                         # :159
    Swb = w[ib] * (z[ib + 1] - b)                                     # :160
    return (S - Sa - Sb) / (Sw - Swa - Swb) * sign                    # :161


def integral_plain(a, b, z, q):
    """sputils.py:94-148 (w is None)."""
    if a < z[0] or a > z[-1] or b < z[0] or b > z[-1]:                # :113-115
        return None
    sign = 1
    if a > b:                                                         # :118-120
        sign = -1
        a, b = b, a
    ia = 0
    while z[ia + 1] < a:                                              # :123-124
        ia += 1
    ib = ia
    while z[ib + 1] < b:                                              # :126-127
        ib += 1
    S = (q[ia:ib + 1] * (z[ia + 1:ib + 2] - z[ia:ib + 1])).sum()      # :142
    Sa = q[ia] * (a - z[ia])                                          # :145
    Sb = q[ib] * (z[ib + 1] - b)                                      # :146
    return (S - Sa - Sb) * sign                                       # :148


def interp_rho(Zh, zh, rho):
    """sputils.py:191-197."""
    RHO = np.zeros(len(Zh) - 1)
    for i in range(len(RHO)):
        if Zh[i] < zh[-1]:                                            # :195
            RHO[i] = integral_plain(Zh[i + 1], Zh[i], zh, rho) / (Zh[i] - Zh[i + 1])   # :196
    return RHO


def interp_c(Zh, zh, q, rho):
    """sputils.py:173-189. Zh descending [nlev+1]; zh ascending cell edges. The reference passes
    les.zh_cache, which has nk (not nk+1) entries, so the top LES cell is never integrated and
    GCM layers reaching above zh[-1] get 0 (sputils.py:111-112 only prints a warning)."""
    Q = np.zeros(len(Zh) - 1)
    for i in range(len(Q)):
        if Zh[i] < zh[-1]:                                            # :187
            Q[i] = integral(Zh[i + 1], Zh[i], zh, q, rho)             # :188
    return Q


def set_gcm_tendencies(gcm, zf, les_prof, A_les, dt_gcm, factor=1.0, conservative=False, zh=None):
    """spcpl.py:388-555 batched. `les_prof`: dict U,V,THL,QT,QL,QL_ice,T (+Rhobf when
    conservative) -> [ncol, nk]; `A_les` [ncol, nlev] is profile["A"] (ascending slab order, as
    returned by les.get_cloudfraction(indices), spcpl.py:765); it is reversed here (:404).
    Returns the 7 tendencies, start_index, the diagnostic t on LES levels and bracket indices."""
    g = _f64(gcm)
    p = _f64(les_prof)
    zf = np.asarray(zf, dtype=np.float64)
    Zf = (g["Zgfull"] - g["Zghalf"][:, -1:]) / grav                   # les.gcm_Zf, :198/:390
    Zh = (g["Zghalf"] - g["Zghalf"][:, -1:]) / grav
    ncol, nlev = Zf.shape
    nk = zf.shape[0]
    ql_water = p["QL"] - p["QL_ice"]                                  # :402
    A_d = np.asarray(A_les, dtype=np.float64)[:, ::-1]                # :404
    t = np.empty((ncol, nk))
    names = ("t_d", "qt_d", "ql_d", "ql_water_d", "ql_ice_d", "u_d", "v_d")
    src = (p["T"], p["QT"], p["QL"], ql_water, p["QL_ice"], p["U"], p["V"])   # :471-477
    d = {n: np.empty((ncol, nlev)) for n in names}
    br = np.empty((ncol, nlev), dtype=np.int32)
    brp = np.empty((ncol, nk), dtype=np.int32)
    start = np.empty(ncol, dtype=np.int32)
    for i in range(ncol):
        pf = np.interp(zf, Zf[i, ::-1], g["Pfull"][i, ::-1])          # :408
        brp[i] = bracket(zf, Zf[i, ::-1])
        t[i] = p["THL"][i] * exner(pf) + rlv * p["QL"][i] / cp        # :409
        for n, s in zip(names, src):
            if not conservative:
                d[n][i] = np.interp(Zf[i], zf, s[i])                  # :471-477
            else:
                d[n][i] = interp_c(Zh[i], zh, s[i], p["Rhobf"][i])  # :482-488
        br[i] = bracket(Zf[i], zf)
        start[i] = np.searchsorted(-Zf[i], -zf[-1])                   # :498
    ft = dt_gcm                                                       # :427
    out = {
        "f_T": factor * (d["t_d"] - g["T"]) / ft,                     # :518
        "f_SH": factor * ((d["qt_d"] - d["ql_d"]) - g["SH"]) / ft,    # :519
        "f_QL": factor * (d["ql_water_d"] - g["QL"]) / ft,            # :520
        "f_QI": factor * (d["ql_ice_d"] - g["QI"]) / ft,              # :521
        "f_U": factor * (d["u_d"] - g["U"]) / ft,                     # :524
        "f_V": factor * (d["v_d"] - g["V"]) / ft,                     # :525
        "f_A": factor * (A_d - g["A"]) / ft,                          # :526
    }
    lev = np.arange(nlev)[None, :]
    above = lev < start[:, None]
    for k in out:                                                     # :527-533
        out[k] = np.where(above, out[k] * 0, out[k])
    out.update(start_index=start, t=t, bracket=br, bracket_pf=brp, A_d=A_d, ql_water=ql_water)
    return out


# ---------------------------------------------------------------------------- slab part
def slab_reduce(vols, ql_thresh=0.0, layout=0, accumulate="f64"):
    """Slab averages + cloud count (north_star part 1; requested at spcpl.py:748-755,765;
    arithmetic external to the reference -> defined here, parity unpinned).

    vols: dict THL,QT,QL,U,V -> [ncol][nk][ny][nx] (layout 0) or [ncol][nx][ny][nk] (layout 1,
    the OMUSE view, spcpl.py:275,288). mean = float64 mean over the horizontal points;
    cnt[c,k] = #{(i,j): float64(ql) > ql_thresh} as int32.

    accumulate="native" sums in the volume's own dtype (what a plain `vol.mean(axis)` does; ~2x
    faster for float32, off by up to ~3e-6 relative). It is used ONLY by the timed CPU-baseline
    legs of bench.py, to be generous to the CPU; parity always uses the float64 definition."""
    ax = (2, 3) if layout == 0 else (1, 2)
    if accumulate == "native":
        prof = {f: np.asarray(v).mean(axis=ax).astype(np.float64) for f, v in vols.items()}
        cnt = np.count_nonzero(np.asarray(vols["QL"]) > ql_thresh, axis=ax).astype(np.int32)
        return prof, cnt
    def mean64(v):
        # float64 mean over the horizontal points with the points CONTIGUOUS, so that numpy's pairwise summation
        # applies: a reduction over strided axes is a plain running sum and drifts to ~1e-12 relative at 65536
        # points (measured against long double), an order of magnitude worse than the parity gate on the means
        v = np.asarray(v).astype(np.float64)
        if layout == 1:
            v = np.moveaxis(v, 3, 1)                 # [ncol][nk][ny][nx] view of the (i, j, k) volume
        ncol, nk = v.shape[:2]
        return np.ascontiguousarray(v).reshape(ncol, nk, -1).sum(axis=2) / float(v.shape[2] * v.shape[3])
    prof = {f: mean64(v) for f, v in vols.items()}
    cnt = np.count_nonzero(np.asarray(vols["QL"]).astype(np.float64) > ql_thresh, axis=ax).astype(np.int32)
    return prof, cnt


def cloud_project(ql, idx, ql_thresh=0.0, layout=0):
    """Projected cloud cover per GCM slab (les.get_cloudfraction(indices), spcpl.py:28,765):
    for ascending slab r the LES levels k in [idx[r-1], idx[r]) (idx[-1] := 0, clipped to nk);
    cntslab[c,r] = #{(i,j): any_k float64(ql[c,k,j,i]) > thr}; empty slab -> 0.
    Returns int32 [ncol, nlev] in ASCENDING slab order (what get_cloudfraction returns)."""
    ql = np.asarray(ql)
    if layout == 1:
        ql = np.transpose(ql, (0, 3, 2, 1))
    ncol, nk = ql.shape[:2]
    cloudy = ql.astype(np.float64) > ql_thresh
    out = np.zeros(idx.shape, dtype=np.int32)
    for c in range(ncol):
        k0 = 0
        for r in range(idx.shape[1]):
            k1 = min(max(int(idx[c, r]), k0), nk)
            if k1 > k0:
                out[c, r] = np.count_nonzero(cloudy[c, k0:k1].any(axis=0))
            k0 = k1
    return out


def coupling_step(gcm, zf, zh, vols, aux, ps_les, dt, f_les, f_gcm, couple_surface=True, ql_thresh=0.0, layout=0,
                  accumulate="f64"):
    """One pass of the whole path in the order the GPU pipeline runs it:
    slab_reduce -> set_les_forcings -> cloud fraction -> set_gcm_tendencies."""
    prof, cnt = slab_reduce(vols, ql_thresh, layout, accumulate)
    frc = set_les_forcings(gcm, zf, prof, ps_les, dt, f_les, couple_surface)
    idx = slab_indices(zh, frc["Zh"])
    cntslab = cloud_project(vols["QL"], idx, ql_thresh, layout)
    s = vols["QL"].shape
    npts = s[2] * s[3] if layout == 0 else s[1] * s[2]
    A = cntslab / float(npts)
    lp = dict(prof)
    lp.update({k: aux[k] for k in ("QL_ice", "T")})
    tnd = set_gcm_tendencies(gcm, zf, lp, A, dt, f_gcm)
    return dict(prof=prof, cnt=cnt, forcings=frc, slab_idx=idx, cntslab=cntslab, tendencies=tnd)
