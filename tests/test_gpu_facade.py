"""The host-side mirror of the reference interface (sp_coupler_b200/{spcpl,sputils,splib,spdummy}.py):
per-LES reference-shaped calls and the batched route must give the same numbers as the oracle."""
import numpy as np

import synth_les
import pytest

from conftest import relerr
from oracle import numpy_batched as nb
from sp_coupler_b200.constants import LES_FIELDS, TENDENCIES, gcm_vars, surf_vars

pytestmark = pytest.mark.gpu


def n(t):
    return t.detach().cpu().numpy()


@pytest.fixture()
def world(cuda_device):
    """A 6-column GPU stand-in world driven through splib.initialize."""
    from sp_coupler_b200 import splib
    splib.initialize(dict(max_num_les=6, les_nx=16, les_ny=16, les_nk=160, gcm_nlev=91, dtype="f64",
                          cplsurf=True, per_column=False, write_diagnostics=False), device=cuda_device)
    return splib


def host_world(splib):
    b = splib.les_batch
    gcm = {k: splib.gcm_model.state[k][:b.ncol] for k in gcm_vars + surf_vars}
    vols = {f: n(v) for f, v in zip(LES_FIELDS, b.vols)}
    aux = {k: n(v) for k, v in b.aux.items()}
    return gcm, vols, aux


def test_batched_step_matches_oracle(world, cuda_device):
    splib = world
    b = splib.les_batch
    gcm, vols, aux = host_world(splib)
    ref = nb.coupling_step(gcm, b.zf_host, b.zh_host, vols, aux, aux["PS"], 900.0, 1.0, 1.0, True)
    from sp_coupler_b200 import spcpl
    spcpl.gather_gcm_data(splib.gcm_model, splib.les_models, True)
    frc = spcpl.set_les_forcings_all(b, 900.0, 1.0, True, firststep=True)
    spcpl.get_les_profiles_all(b)
    res = spcpl.set_gcm_tendencies_all(splib.gcm_model, b, 900.0, 1.0)
    for k in ("f_u", "f_v", "f_thl", "f_qt", "f_ql", "f_ps", "wthl", "wqt"):
        assert relerr(n(frc[k]), ref["forcings"][k]) <= 1e-6, k
    for k in TENDENCIES:
        assert relerr(n(res[k]), ref["tendencies"][k]) <= 1e-6, k
    # the GCM received the packed block
    idx, arr = splib.gcm_model.tendencies["T"]["batch"]
    assert relerr(arr, ref["tendencies"]["f_T"]) <= 1e-6


def test_per_column_calls_match_batched_and_oracle(world, cuda_device):
    splib = world
    from sp_coupler_b200 import spcpl, sputils
    b = splib.les_batch
    gcm, vols, aux = host_world(splib)
    ref = nb.coupling_step(gcm, b.zf_host, b.zh_host, vols, aux, aux["PS"], 900.0, 0.5, 2.0, True)
    spcpl.gather_gcm_data(splib.gcm_model, splib.les_models, True)
    for les in splib.les_models[:3]:
        i = les.i
        req = spcpl.set_les_forcings(les, splib.gcm_model, True, True, {}, 900.0, 0.5, True, write=False)
        assert set(req) == {"U", "V", "THL", "QT", "SP", "QL", "QLp", "Z0M_surf", "Z0H_surf", "WT_surf", "WQ_surf"}
        assert all(hasattr(r, "result") for r in req.values())
        assert relerr(n(b.tend["THL"][i]), ref["forcings"]["f_thl"][i]) <= 1e-6
        assert relerr(n(b.tend["U"][i]), ref["forcings"]["f_u"][i]) <= 1e-6
        assert relerr(n(b.ql_ref[i]), ref["forcings"]["ql_ref"][i]) <= 1e-6
        assert relerr(n(les.gcm_Zf), ref["forcings"]["Zf"][i]) == 0.0
        prof = {k: v.result() for k, v in spcpl.get_les_profiles(les, True).items()}
        assert set(prof) == {"U", "V", "presf", "Rhof", "Rhobf", "THL", "QT", "QL", "QL_ice", "QR", "PS", "T", "A", "Rain"}
        assert relerr(n(prof["A"]), ref["cntslab"][i] / 256.0) <= 1e-12
        assert relerr(n(prof["THL"]), ref["prof"]["THL"][i]) <= 1e-12
        # spcpl.get_cloud_fraction: same numbers, reversed to GCM order (spcpl.py:28)
        assert relerr(n(spcpl.get_cloud_fraction(les)), (ref["cntslab"][i] / 256.0)[::-1]) <= 1e-12
        spcpl.set_gcm_tendencies(splib.gcm_model, les, prof, 900.0, factor=2.0, write=False)
        for name, key in (("T", "f_T"), ("SH", "f_SH"), ("QL", "f_QL"), ("QI", "f_QI"), ("U", "f_U"), ("V", "f_V"), ("A", "f_A")):
            assert relerr(splib.gcm_model.tendencies[name][les.grid_index], ref["tendencies"][key][i]) <= 1e-6, name
        u, v, thl, qt, ps, ql = spcpl.convert_profiles(les, write=False)
        assert relerr(n(thl), ref["forcings"]["thl"][i]) <= 1e-12
        assert relerr(n(ps), ref["forcings"]["ps"][i]) == 0.0
        z0m, z0h, wthl, wqt = spcpl.convert_surface_fluxes(les)
        assert relerr(n(wthl), ref["forcings"]["wthl"][i]) <= 1e-12
    # sputils mirror
    p = b.pipe.gcm["Pfull"][0]
    assert relerr(n(sputils.exner(p) * sputils.iexner(p)), np.ones(91)) <= 1e-14
    x = b.pipe.zf
    Zf = splib.les_models[0].gcm_Zf.flip(0).contiguous()
    out = sputils.interp(x, Zf, b.pipe.gcm["U"][0].flip(0).contiguous())
    assert np.array_equal(n(out), np.interp(b.zf_host, n(Zf), n(b.pipe.gcm["U"][0])[::-1]))
    assert int(sputils.searchsorted(-splib.les_models[0].gcm_Zf, -x[-1])) == int(ref["tendencies"]["start_index"][0])


def test_driver_loop_runs_and_both_routes_agree(cuda_device):
    """splib.initialize/run/finalize with per-LES calls vs the batched route: identical GCM state."""
    from sp_coupler_b200 import splib
    cfg = dict(max_num_les=4, les_nx=16, les_ny=16, les_nk=160, gcm_nlev=19, dtype="f64", cplsurf=True,
               write_diagnostics=False)
    states = []
    for per_column in (False, True):
        splib.initialize(dict(cfg, per_column=per_column), device=cuda_device)
        splib.run(3)
        splib.finalize()
        assert len(splib.timing_rows) == 3 and splib.gcm_model.model_time == 3 * 900.0
        states.append({k: splib.gcm_model.state[k][:4].copy() for k in ("T", "SH", "U", "A")})
    for k in states[0]:
        assert relerr(states[1][k], states[0][k]) <= 1e-12, k
    # the coupling did something: relaxation pulled the LES means toward the GCM
    assert np.isfinite(states[0]["T"]).all()


def test_set_les_state_and_diagnostics_store(world, cuda_device):
    import torch
    from sp_coupler_b200 import spcpl, spio, synth
    splib = world
    les = splib.les_models[2]
    spcpl.gather_gcm_data(splib.gcm_model, splib.les_models, True)
    u, v, thl, qt, ps, ql = spcpl.convert_profiles(les, write=False)
    spcpl.set_les_state(les, u, v, thl, qt, ps)
    b = splib.les_batch
    vol = n(b.vols[LES_FIELDS.index("THL")][les.i])
    host = synth_les.les_state_volume(n(thl)[None, :], 0.1, synth.STREAM["THL"], b.nx, b.ny, seed=b.seed, col0=les.i,
                                  dtype=np.float64)[0]
    assert np.array_equal(vol, host)
    assert abs(vol.mean(axis=(1, 2)) - n(thl)).max() < 0.1 * 4 / np.sqrt(256)      # uniform noise, amplitude 0.1 K
    # diagnostics keep the reference's spifs.nc variable names
    spio.init_netcdf("unused.npz", splib.gcm_model, splib.les_models)
    spio.update_time(0.0)
    spcpl.set_les_forcings(les, splib.gcm_model, False, True, {}, 900.0, 1.0, True, write=True)
    prof = spcpl.get_les_profiles(les, False)
    spcpl.set_gcm_tendencies(splib.gcm_model, les, prof, 900.0, write=True)
    names = set(les.cdf)
    for want in ("U", "V", "T", "SH", "QL", "QI", "Pf", "Ph", "Zf", "Zh", "Psurf", "Tv", "THL", "QT", "f_u", "f_v", "f_thl",
                 "f_qt", "rain", "rainrate", "z0m", "z0h", "wthl", "wqt", "TLflux", "TSflux", "SHflux", "QLflux", "QIflux",
                 "u", "v", "presf", "rhof", "rhobf", "qt", "ql", "ql_ice", "ql_water", "thl", "t", "t_", "qr",
                 "f_U", "f_V", "f_T", "f_SH", "A", "A_d", "f_QL", "f_QI", "f_A"):
        assert want in names, want


def test_sputils_integrals_match_reference_golden(cuda_device):
    """sputils.integral / interp_c / interp_rho on the GPU path against the unmodified reference's outputs."""
    import os
    import torch
    from conftest import GOLDEN, relerr
    from sp_coupler_b200 import sputils
    z = np.load(os.path.join(GOLDEN, "ref_sputils.npz"))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda_device)
    for tag in ("a", "b"):
        Zh, q, rho = t(z[tag + "_Zh"]), t(z[tag + "_q"]), t(z[tag + "_rho"])
        for name in ("full", "short"):
            edges = t(z["%s_%s_z" % (tag, name)])
            qc = sputils.interp_c(Zh, edges, q, rho)                        # batched over the two columns
            assert relerr(qc.cpu().numpy(), z["%s_%s_interp_c" % (tag, name)]) <= 1e-13
            rc = sputils.interp_rho(Zh[1], edges, rho[1])                   # one column, reference call shape
            assert relerr(rc.cpu().numpy(), z["%s_%s_interp_rho" % (tag, name)][1]) <= 1e-13
            zero = z["%s_%s_interp_c" % (tag, name)] == 0                   # layers above the LES top stay exactly 0
            assert zero.any() and (qc.cpu().numpy()[zero] == 0).all()
        edges = t(z[tag + "_full_z"])
        scale = float(np.abs(z[tag + "_integral"][2, 0]))
        for (a, b), (plain, weighted) in zip(z[tag + "_ab"], z[tag + "_integral"]):
            assert abs(float(sputils.integral(a, b, edges, q[0])) - plain) <= 1e-13 * scale
            if not np.isnan(weighted):
                assert abs(float(sputils.integral(a, b, edges, q[0], rho[0])) - weighted) <= 1e-13 * abs(weighted)
        assert sputils.integral(-1.0, 10.0, edges, q[0]) is None
    f32 = sputils.interp_c(Zh.float(), edges, q.float(), rho.float())        # float32 storage, float64 arithmetic
    assert relerr(f32.cpu().numpy(), z["b_full_interp_c"]) <= 1e-6


def test_output_column_conversion(cuda_device):
    """spcpl.output_column_conversion (spcpl.py:251-270) against the same expressions in numpy."""
    import torch
    from sp_coupler_b200 import spcpl, synth
    from oracle import numpy_batched as nb
    g = synth.make_gcm_columns(1, 91, seed=12)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a[0])).to(cuda_device)
    prof = {"T": t(g["T"]), "SH": t(g["SH"]), "QL": t(g["QL"]), "QI": t(g["QI"]), "Pf": t(g["Pfull"]), "Ph": t(g["Phalf"]),
            "Zgfull": t(g["Zgfull"]), "Zghalf": t(g["Zghalf"])}
    spcpl.output_column_conversion(prof)
    T, SH, QL, QI = g["T"][0], g["SH"][0], g["QL"][0], g["QI"][0]
    c = nb.rv / nb.rd - 1
    want = {"Tv": T * (1 + c * SH - (QL + QI)), "Zh": ((g["Zghalf"][0] - g["Zghalf"][0][-1]) / nb.grav)[1:],
            "Zf": (g["Zgfull"][0] - g["Zghalf"][0][-1]) / nb.grav, "Psurf": g["Phalf"][0][-1], "Ph": g["Phalf"][0][1:],
            "THL": (T - (nb.rlv * (QL + QI)) / nb.cp) * nb.iexner(g["Pfull"][0]), "QT": SH + QL + QI}
    for k, v in want.items():
        got = prof[k].cpu().numpy()
        assert got.shape == np.shape(v), k
        assert np.max(np.abs(got - v)) <= 1e-12 * np.max(np.abs(v)), k


def test_spinup_then_coupled_steps(cuda_device):
    """splib.run_spinup / step_spinup (splib.py:233-251, 355-402): the LES relax towards the initial GCM profiles
    with the spin-up forcing factor, the GCM neither steps nor receives tendencies; the coupled steps that follow
    advance the LES clocks from the spin-up offset."""
    import torch
    from sp_coupler_b200 import splib
    cfg = dict(max_num_les=4, les_nx=16, les_ny=16, les_nk=160, gcm_nlev=19, dtype="f64", cplsurf=True)
    splib.initialize(dict(cfg, les_spinup=0), device=cuda_device)
    b0 = splib.les_batch
    thl0 = b0.vols[0].clone()
    gcm_T0 = splib.gcm_model.state["T"].copy()
    splib.initialize(dict(cfg, les_spinup=1800.0, les_spinup_steps=3, les_spinup_forcing_factor=0.5), device=cuda_device)
    b = splib.les_batch
    assert b.model_time == 1800.0 and len(splib.timing_rows) == 3
    assert all(r[1] == 0.0 and r[2] == 0.0 and r[5] == 0.0 for r in splib.timing_rows)   # no GCM work in a spin-up row
    assert np.array_equal(splib.gcm_model.state["T"], gcm_T0)                             # the GCM got no tendencies
    # same initial state (seeded), then nudged: the slab means moved towards the GCM profile on the LES levels
    target = b.cpl.gcm_to_les(b.pipe.gcm, b.pipe.zf, b.pipe.zh, None, None, 1.0, 1.0, False, want_state=True)["thl"].double()
    before = (thl0.double().mean(dim=(2, 3)) - target).abs().mean()
    after = (b.vols[0].double().mean(dim=(2, 3)) - target).abs().mean()
    assert float(after) < float(before)
    assert not splib.firststep
    splib.step()
    assert b.model_time == 1800.0 + splib.gcm_model.get_timestep()
    splib.finalize()
