"""GPU mirror of the reference coupling module (splib/spcpl.py): same entry-point names, argument
meaning and return structure, computed by the sm_100a kernels behind include/spcpl_b200.h.

Two routes to the same kernels:
  * per-column functions with the reference signatures (`set_les_forcings(les, gcm, ...)`,
    `set_gcm_tendencies(gcm, les, profile, ...)`, ...) — each call runs the kernels on a batch of
    ONE column; this is the compatibility route, so the reference's driver loop works unchanged;
  * `*_all` functions over a `gpu_les_batch` — what splib.step() calls on the GPU path: one launch
    per kernel for all columns.
Quantities are plain SI device tensors. Reference quirks deliberately NOT replicated: the
`if not any(cols)` test that treats the single column index 0 as "no columns" (spcpl.py:63,71) and
the dead duplicated tendency block (spcpl.py:501-515).
"""
import logging
import time

import numpy as np
import torch

from . import spio
from .constants import LES_FIELDS, gcm_vars, surf_vars
from .coupler import default_coupler

log = logging.getLogger(__name__)

var_to_netcdf_name = {"Z0M": "z0m", "Z0H": "z0h", "Phalf": "Ph", "Pfull": "Pf"}  # spcpl.py:47-51


def _cpl(les):
    return les.batch.cpl if hasattr(les, "batch") else default_coupler()


def _row(t):
    """[n] (or 0-dim) view -> contiguous [1, n] / [1]."""
    t = t if isinstance(t, torch.Tensor) else torch.as_tensor(t)
    return t.reshape(1, -1).contiguous() if t.dim() >= 1 else t.reshape(1).contiguous()


def _col_gcm(les, couple_surface):
    g = {v: _row(getattr(les, v)) for v in gcm_vars}
    if couple_surface:
        for v in surf_vars:
            g[v] = getattr(les, v).reshape(1).contiguous()
    return g


def _les_grid(les):
    zf = les.zf_cache if hasattr(les, "zf_cache") else les.get_zf()
    zh = les.zh_cache if hasattr(les, "zh_cache") else les.get_zh()
    return zf, zh


# ---------------------------------------------------------------------------------------------
def get_cloud_fraction(les):
    """spcpl.py:22-29: LES cloud fraction on the GCM levels (GCM order, top -> bottom)."""
    from . import sputils
    indices = sputils.searchsorted(les.zh_cache, les.gcm_Zh, side="right")[:-1].flip(0)   # spcpl.py:26
    return les.get_cloudfraction(indices).flip(0)                                           # spcpl.py:28


def gather_gcm_data(gcm, les_models, couple_surface, output_column_indices=None, write=True):
    """spcpl.py:55-131: fetch the GCM profiles / surface fields of every SP column and attach them to
    the LES objects. On the GPU path the data lands in ONE pinned staging buffer and goes to the
    device in ONE async copy; the per-LES attributes are views of the device arrays."""
    extra_cols = [] if output_column_indices is None else list(output_column_indices)
    cols = [les.grid_index for les in les_models] + extra_cols
    if len(cols) == 0:
        return
    start = time.time()
    profile_data = {v: gcm.get_profile_fields(v, cols) for v in gcm_vars}               # spcpl.py:62-67
    surface_data = {v: gcm.get_surface_field(v, cols) for v in surf_vars} if couple_surface else {}
    log.info("Fetching gcm data took %d s" % (time.time() - start))
    n = len(les_models)
    batch = getattr(les_models[0], "batch", None) if n else None
    if batch is not None and batch.pipe.world > 1:
        raise RuntimeError("columns are sharded over ranks: use gather_gcm_data_sharded")
    if batch is not None and n == batch.ncol and all(getattr(l, "batch", None) is batch and l.i == i
                                                     for i, l in enumerate(les_models)):
        st = batch.pipe.staging
        for v in gcm_vars:
            st.host[v].copy_(torch.as_tensor(np.ascontiguousarray(profile_data[v][:n])))
        for v in surface_data:
            st.host[v].copy_(torch.as_tensor(np.ascontiguousarray(surface_data[v][:n])))
        dev = st.upload()
        for i, les in enumerate(les_models):                                            # spcpl.py:81-86
            for v in gcm_vars:
                setattr(les, v, dev[v][i])
            for v in surface_data:
                setattr(les, v, dev[v][i])
    else:
        for i, les in enumerate(les_models):
            device = _cpl(les).device
            for v in gcm_vars:
                setattr(les, v, torch.as_tensor(np.ascontiguousarray(profile_data[v][i])).to(device))
            for v in surface_data:
                setattr(les, v, torch.as_tensor(np.asarray(surface_data[v][i])).to(device))
    # extra output columns (spcpl.py:88-129): converted profiles go to the diagnostics store
    if extra_cols and write:
        cpl = default_coupler()
        sl = slice(n, None)
        g = {v: torch.as_tensor(np.ascontiguousarray(profile_data[v][sl])).to(cpl.device) for v in gcm_vars}
        if couple_surface:
            g.update({v: torch.as_tensor(np.ascontiguousarray(surface_data[v][sl])).to(cpl.device) for v in surf_vars})
        zf = torch.zeros(1, dtype=torch.float64, device=cpl.device)
        d = cpl.gcm_to_les(g, zf, None, None, None, 1.0, 1.0, couple_surface, diagnostics=True)
        for j, col in enumerate(extra_cols):
            spio.write_netCDF_data(col, U=g["U"][j], V=g["V"][j], T=g["T"][j], SH=g["SH"][j], QL=g["QL"][j],
                                   QI=g["QI"][j], Pf=g["Pfull"][j], Ph=g["Phalf"][j][1:], Zf=d["Zf"][j],
                                   Zh=d["Zh"][j][1:], Psurf=g["Phalf"][j][-1], Tv=d["Tv"][j], THL=d["THL"][j],
                                   QT=d["QT"][j], A=g["A"][j])
            if couple_surface:
                spio.write_netCDF_data(col, z0m=d["z0m"][j], z0h=d["z0h"][j], wthl=d["wthl"][j], wqt=d["wqt"][j])


def gather_gcm_data_sharded(gcm, batch, couple_surface=True):
    """gather_gcm_data when the SP columns are sharded over GPUs (one process per GPU): the rank that
    owns the GCM fetches the profiles of ALL columns (spcpl.py:62-75), packs them per rank, uploads once,
    and a scatter over NVLink delivers every rank's block into its staging buffer; then the per-LES
    attributes are views of the device arrays as in the single-GPU path (spcpl.py:81-86)."""
    from .pipeline import GcmScatter, HostExchange
    pipe = batch.pipe
    if getattr(batch, "gather_mode", None) == "host":
        # the GCM's host memory is one pinned buffer shared by all ranks: every rank stages its own columns over
        # its own PCIe link (pipeline.HostExchange); the tendencies go back the same way (set_gcm_tendencies_all)
        if getattr(batch, "_exchange", None) is None:
            batch._exchange = HostExchange(pipe, pipe.world, pipe.rank, owner=0, group=pipe.group, tag="splib")
        if pipe.rank == 0:
            cols = batch.all_grid_indices
            data = {v: gcm.get_profile_fields(v, cols) for v in gcm_vars}
            data.update({v: gcm.get_surface_field(v, cols) for v in surf_vars})
            batch._exchange.fill_inputs(data)       # cut to the live level window (pipeline.py, "Level window")
        dev = batch._exchange.fetch_inputs()
        for i, les in enumerate(batch.models):
            for v in gcm_vars:
                setattr(les, v, dev[v][i])
            if couple_surface:
                for v in surf_vars:
                    setattr(les, v, dev[v][i])
        return dev
    if getattr(batch, "_scatter", None) is None:
        batch._scatter = GcmScatter(pipe.staging, pipe.world, pipe.rank, owner=0, group=pipe.group)
    if pipe.rank == 0:
        cols = batch.all_grid_indices
        data = {v: gcm.get_profile_fields(v, cols) for v in gcm_vars}
        data.update({v: gcm.get_surface_field(v, cols) for v in surf_vars})
        batch._scatter.fill_host(data)
    dev = batch._scatter.scatter()
    for i, les in enumerate(batch.models):
        for v in gcm_vars:
            setattr(les, v, dev[v][i])
        if couple_surface:
            for v in surf_vars:
                setattr(les, v, dev[v][i])
    return dev


def convert_surface_fluxes(les):
    """spcpl.py:136-167 -> (Z0M, Z0H, wthl, wqt)."""
    cpl = _cpl(les)
    zf, zh = _les_grid(les)
    d = cpl.gcm_to_les(_col_gcm(les, True), zf, None, None, None, 1.0, 1.0, True)
    return d["z0m"][0], d["z0h"][0], d["wthl"][0], d["wqt"][0]


def convert_profiles(les, write=True):
    """spcpl.py:171-246 -> (u, v, thl, qt, ps, ql) on the LES levels; stores les.gcm_Zf / gcm_Zh."""
    cpl = _cpl(les)
    zf, zh = _les_grid(les)
    d = cpl.gcm_to_les(_col_gcm(les, False), zf, zh, None, None, 1.0, 1.0, False, diagnostics=True, want_state=True)
    les.gcm_Zf, les.gcm_Zh = d["Zf"][0], d["Zh"][0]                                       # spcpl.py:200-201
    les.slab_idx = d["slab_idx"][0]
    if write:                                                                             # spcpl.py:230-244
        spio.write_les_data(les, U=les.U, V=les.V, T=les.T, SH=les.SH, QL=les.QL, QI=les.QI, Pf=les.Pfull,
                            Ph=les.Phalf[1:], Zf=d["Zf"][0], Zh=d["Zh"][0][1:], Psurf=les.Phalf[-1], Tv=d["Tv"][0],
                            THL=d["THL"][0], QT=d["QT"][0])
    return d["u"][0], d["v"][0], d["thl"][0], d["qt"][0], d["ps"][0], d["ql_ref"][0]


def output_column_conversion(profile):
    """spcpl.py:251-270: derived diagnostics of one extra output column (no LES attached), in place:
    Tv, Zh (without the top edge), Zf, Psurf, Ph (without the top edge), THL, QT from the GCM profiles
    T, SH, QL, QI, Pf, Ph, Zgfull, Zghalf (1-D device tensors, top -> bottom). Same kernel as convert_profiles."""
    T = profile['T']
    cpl = default_coupler(T.device)
    row = lambda t: t.reshape(1, -1).to(T.dtype).contiguous()
    zero = torch.zeros_like(T)
    g = {"U": row(profile.get('U', zero)), "V": row(profile.get('V', zero)), "T": row(T), "SH": row(profile['SH']),
         "QL": row(profile['QL']), "QI": row(profile['QI']), "Pfull": row(profile['Pf']), "A": row(profile.get('A', zero)),
         "Zgfull": row(profile['Zgfull']), "Phalf": row(profile['Ph']), "Zghalf": row(profile['Zghalf'])}
    d = cpl.gcm_to_les(g, torch.zeros(1, dtype=torch.float64, device=T.device), None, None, None, 1.0, 1.0, False,
                       diagnostics=True)
    profile['Tv'] = d["Tv"][0]                                                           # spcpl.py:253
    profile['Zh'] = d["Zh"][0][1:]                                                       # spcpl.py:258,261
    profile['Zf'] = d["Zf"][0]                                                           # spcpl.py:259,262
    profile['Psurf'] = profile['Ph'][-1]                                                 # spcpl.py:263
    profile['Ph'] = profile['Ph'][1:]                                                    # spcpl.py:264
    profile['THL'] = d["THL"][0]                                                         # spcpl.py:265-266
    profile['QT'] = d["QT"][0]                                                           # spcpl.py:267


def set_les_state(les, u, v, thl, qt, ps=None):
    """spcpl.py:274-294: broadcast the profiles to the 3-D fields with uniform noise of amplitude
    0.5 m/s, 0.1 K, 2.5e-5 (Philox stream instead of numpy's Mersenne Twister)."""
    from . import synth
    b, cpl = les.batch, les.batch.cpl
    for f, p in (("U", u), ("V", v), ("THL", thl), ("QT", qt)):
        cpl.set_les_state(p.reshape(1, -1).double().contiguous(), synth.NOISE_AMP[f], synth.STREAM[f], b.nx, b.ny,
                          seed=b.seed, col0=b.col0 + les.i, dtype=b.dtype,
                          out=b.vols[LES_FIELDS.index(f)][les.i:les.i + 1])
    b.state_version += 1
    if ps is not None:
        les.set_surface_pressure(ps)


def set_les_forcings(les, gcm, asynchronous, firststep, profile, dt_gcm, factor, couple_surface, qt_forcing='sp',
                     write=True, variability_nudge_constant_T=False):
    """spcpl.py:299-385 for one LES: forcings f = factor*(x_gcm->les - <x>_les)/dt_gcm handed to the
    LES through its set_tendency_* methods. Returns the dict of requests the reference returns."""
    cpl = _cpl(les)
    if firststep:                                                                         # spcpl.py:302-308
        u_d, v_d = les.get_profile_U(), les.get_profile_V()
        thl_d, qt_d, ql_d = les.get_profile_THL(), les.get_profile_QT(), les.get_profile_QL()
        ps_d = les.get_surface_pressure()
        rain = les.get_rain()
    else:                                                                                 # spcpl.py:310-315
        u_d, v_d, thl_d, qt_d, ql_d = (profile[k] for k in ("U", "V", "THL", "QT", "QL"))
        ps_d, rain = profile["PS"], profile["Rain"]
    rain_last = getattr(les, "rain", 0.0)                                                 # spcpl.py:316-319
    les.rain = rain
    rainrate = (rain - rain_last) / dt_gcm
    zf, zh = _les_grid(les)
    g = _col_gcm(les, couple_surface)
    dtype = g["T"].dtype
    lp = torch.stack([thl_d, qt_d, ql_d, u_d, v_d]).double().reshape(5, 1, -1).contiguous()
    d = cpl.gcm_to_les(g, zf, zh, lp, torch.as_tensor(ps_d, device=cpl.device).reshape(1).to(dtype), float(dt_gcm),
                       float(factor), couple_surface, diagnostics=True)
    les.gcm_Zf, les.gcm_Zh, les.slab_idx = d["Zf"][0], d["Zh"][0], d["slab_idx"][0]
    if write:                                                                             # convert_profiles' write
        spio.write_les_data(les, U=les.U, V=les.V, T=les.T, SH=les.SH, QL=les.QL, QI=les.QI, Pf=les.Pfull,
                            Ph=les.Phalf[1:], Zf=d["Zf"][0], Zh=d["Zh"][0][1:], Psurf=les.Phalf[-1], Tv=d["Tv"][0],
                            THL=d["THL"][0], QT=d["QT"][0])
    a = asynchronous
    u_t = les.set_tendency_U(d["f_u"][0], return_request=a)                               # spcpl.py:341-347
    v_t = les.set_tendency_V(d["f_v"][0], return_request=a)
    thl_t = les.set_tendency_THL(d["f_thl"][0], return_request=a)
    qt_t = les.set_tendency_QT(d["f_qt"][0], return_request=a)
    sp_t = les.set_tendency_surface_pressure(d["f_ps"][0], return_request=a)
    ql_t = les.set_tendency_QL(d["f_ql"][0], return_request=a)
    ql_p_t = les.set_ref_profile_QL(d["ql_ref"][0], return_request=a)
    les.ql_ref = d["ql_ref"][0]                                                           # spcpl.py:348
    if write:
        spio.write_les_data(les, f_u=d["f_u"][0], f_v=d["f_v"][0], f_thl=d["f_thl"][0], f_qt=d["f_qt"][0],
                            rain=rain, rainrate=rainrate * 3600)
    req = {"U": u_t, "V": v_t, "THL": thl_t, "QT": qt_t, "SP": sp_t, "QL": ql_t, "QLp": ql_p_t}
    if couple_surface:                                                                    # spcpl.py:359-376
        req["Z0M_surf"] = les.set_z0m_surf(d["z0m"][0], return_request=a)
        req["Z0H_surf"] = les.set_z0h_surf(d["z0h"][0], return_request=a)
        req["WT_surf"] = les.set_wt_surf(d["wthl"][0], return_request=a)
        req["WQ_surf"] = les.set_wq_surf(d["wqt"][0], return_request=a)
        if write:
            spio.write_les_data(les, z0m=d["z0m"][0], z0h=d["z0h"][0], wthl=d["wthl"][0], wqt=d["wqt"][0])
            spio.write_les_data(les, TLflux=les.TLflux, TSflux=les.TSflux, SHflux=les.SHflux, QLflux=les.QLflux,
                                QIflux=les.QIflux)
    if qt_forcing == 'variance':                                                          # spcpl.py:377-382
        if les.get_model_time() > 0:
            from .nudge import variability_nudge
            variability_nudge(les, dt_gcm, variability_nudge_constant_T, write=write)
    return req


def get_les_profiles(les, asynchronous):
    """spcpl.py:747-767: slab-averaged LES profiles + cloud fraction of one LES, as a dict of
    requests (asynchronous) or values."""
    cpl = _cpl(les)
    slab = cpl.slab_reduce(les.volumes(), want_mask=True)                                 # spcpl.py:748-755
    idx = les.slab_idx if hasattr(les, "slab_idx") else None
    if idx is None:
        from . import sputils
        idx = sputils.searchsorted(les.zh_cache, les.gcm_Zh, side="right")[:-1].flip(0)   # spcpl.py:761-764
    A, _ = cpl.cloud_fraction(slab, idx.reshape(1, -1).contiguous().to(torch.int32))      # spcpl.py:765
    p = slab["prof"]
    from .spdummy import Request
    w = (lambda x: Request(x)) if asynchronous else (lambda x: x)
    return {"U": w(p[LES_FIELDS.index("U"), 0]), "V": w(p[LES_FIELDS.index("V"), 0]),
            "presf": les.get_presf(return_request=asynchronous), "Rhof": les.get_rhof(return_request=asynchronous),
            "Rhobf": les.get_rhobf(return_request=asynchronous), "THL": w(p[LES_FIELDS.index("THL"), 0]),
            "QT": w(p[LES_FIELDS.index("QT"), 0]), "QL": w(p[LES_FIELDS.index("QL"), 0]),
            "QL_ice": les.get_profile_QL_ice(return_request=asynchronous),
            "QR": les.get_profile_QR(return_request=asynchronous),
            "PS": les.get_surface_pressure(return_request=asynchronous),
            "T": les.get_profile_T(return_request=asynchronous), "A": w(A[0]),
            "Rain": les.get_rain(return_request=asynchronous)}


def set_gcm_tendencies(gcm, les, profile, dt_gcm, factor=1, write=True, conservative=False):
    """spcpl.py:388-555 for one LES: tendencies f_X = factor*(<x>_les->gcm - X)/dt_gcm, zeroed above
    the LES top, handed to gcm.set_profile_tendency (7 calls, spcpl.py:535-542)."""
    cpl = _cpl(les)
    zf, zh = _les_grid(les)
    g = _col_gcm(les, False)
    dtype = g["T"].dtype
    lp = torch.stack([profile[k] for k in LES_FIELDS]).double().reshape(5, 1, -1).contiguous()
    aux = {"QL_ice": _row(profile["QL_ice"]).to(dtype), "T": _row(profile["T"]).to(dtype)}
    if conservative:
        aux["Rhobf"] = _row(profile["Rhobf"]).to(dtype)
    d = cpl.les_to_gcm(g, zf, zh, {"prof": lp}, aux, None, float(dt_gcm), float(factor), conservative=conservative,
                       A=_row(profile["A"]).to(dtype), diagnostics=write)
    if write:                                                                             # spcpl.py:412-425
        spio.write_les_data(les, u=profile["U"], v=profile["V"], presf=profile["presf"], rhof=profile["Rhof"],
                            rhobf=profile["Rhobf"], qt=profile["QT"], ql=profile["QL"], ql_ice=profile["QL_ice"],
                            ql_water=profile["QL"] - profile["QL_ice"], thl=profile["THL"], t=d["t"][0],
                            t_=profile["T"], qr=profile["QR"])
    for name in ("U", "V", "T", "SH", "QL", "QI", "A"):                                   # spcpl.py:535-542
        gcm.set_profile_tendency(name, les.grid_index, d["f_" + name][0])
    if write:                                                                             # spcpl.py:545-555
        spio.write_les_data(les, f_U=d["f_U"][0], f_V=d["f_V"][0], f_T=d["f_T"][0], f_SH=d["f_SH"][0], A=les.A,
                            A_d=d["A_d"][0], f_QL=d["f_QL"][0], f_QI=d["f_QI"][0], f_A=d["f_A"][0])
    return d


def write_les_profiles(les):
    """spcpl.py:574-609 (spin-up diagnostics)."""
    p = get_les_profiles(les, False)
    cpl = _cpl(les)
    zf, zh = _les_grid(les)
    g = _col_gcm(les, False)
    dtype = g["T"].dtype
    lp = torch.stack([p[k] for k in LES_FIELDS]).double().reshape(5, 1, -1).contiguous()
    aux = {"QL_ice": _row(p["QL_ice"]).to(dtype), "T": _row(p["T"]).to(dtype)}
    d = cpl.les_to_gcm(g, zf, zh, {"prof": lp}, aux, None, 1.0, 1.0, A=_row(p["A"]).to(dtype), diagnostics=True)
    spio.write_les_data(les, u=p["U"], v=p["V"], presf=p["presf"], qt=p["QT"], ql=p["QL"], ql_ice=p["QL_ice"],
                        ql_water=p["QL"] - p["QL_ice"], thl=p["THL"], t=d["t"][0], t_=p["T"], qr=p["QR"])


# --------------------------------------------------------------------------------- batched route
def set_les_forcings_all(batch, dt_gcm, factor, couple_surface=True, firststep=False, qt_forcing='sp',
                         variability_nudge_constant_T=False):
    """set_les_forcings for every LES of the batch in one K2 launch (first step: K1 first to get
    the slab means, spcpl.py:302-308). Forcings are written straight into the batch's tendency
    buffers (what the per-LES set_tendency_* calls do one by one). With qt_forcing == 'variance' the
    qt-variability nudge follows for all columns in one launch once the LES clocks have left 0, exactly
    where the reference runs it per LES (spcpl.py:377-382)."""
    pipe = batch.pipe
    if firststep or pipe.slab is None:
        pipe.les_profiles()
    frc = pipe.forcings(float(dt_gcm), float(factor))
    batch.tend["U"], batch.tend["V"] = frc["f_u"], frc["f_v"]
    batch.tend["THL"], batch.tend["QT"], batch.tend["QL"] = frc["f_thl"], frc["f_qt"], frc["f_ql"]
    batch.tend_ps, batch.ql_ref = frc["f_ps"], frc["ql_ref"]
    if couple_surface:
        batch.surf.update(z0m=frc["z0m"], z0h=frc["z0h"], wt=frc["wthl"], wq=frc["wqt"])
    batch.last_forcings = frc
    if qt_forcing == 'variance':                                                          # spcpl.py:377-382
        if batch.model_time > 0:
            from .nudge import variability_nudge_all
            batch.last_nudge = variability_nudge_all(batch, dt_gcm, variability_nudge_constant_T)
    elif qt_forcing != 'sp':
        raise ValueError("qt_forcing must be 'sp' or 'variance', got %r" % (qt_forcing,))
    return frc


def get_les_profiles_all(batch):
    """get_les_profiles for every LES in one K1 launch."""
    return batch.pipe.les_profiles()


def set_gcm_tendencies_all(gcm, batch, dt_gcm, factor=1, conservative=False, to_host=True):
    """set_gcm_tendencies for every LES in one K3 launch, which also gathers the packed [ncol][7][nlev] blocks on the
    GCM-owning rank when the columns are sharded (NVLink stores, or PCIe stores into the host GCM's memory with
    gather "host"; an NCCL all_gather with gather "nccl"); the block goes back to the host GCM in one copy."""
    pipe = batch.pipe
    res = pipe.tendencies(batch.last_forcings, float(dt_gcm), float(factor), conservative=conservative)
    if getattr(batch, "_exchange", None) is not None:          # sharded, host-resident GCM: no device gather
        out = batch._exchange.wait_tendencies()                # K3 stored every rank's block into the shared host buffer
        if to_host and gcm is not None and pipe.rank == 0:
            gcm.set_profile_tendencies(batch.all_grid_indices, out, lev0=batch._exchange.lev0)
        return res
    if to_host and gcm is not None and pipe.rank == 0:
        pipe.tend_host.copy_(pipe.tend_all, non_blocking=True)
        torch.cuda.current_stream(pipe.cpl.device).synchronize()
        if pipe.sync_error():       # K3's device-side barrier gave up on a peer (20 s): the gathered block is incomplete
            raise RuntimeError("set_gcm_tendencies_all: rank %d never signalled its tendency block (sync error %d)"
                               % (pipe.sync_error() - 1, pipe.sync_error()))
        cols = getattr(batch, "all_grid_indices", None)
        if cols is None:
            cols = [les.grid_index for les in batch.models]
        gcm.set_profile_tendencies(cols, pipe.tend_host)
    return res
