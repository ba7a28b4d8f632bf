"""Pins oracle/numpy_batched.py: against the reference's own known-answer tests, against golden
vectors produced by the unmodified reference (tests/golden/, oracle/make_golden.py), and — where
/root/reference is present — against the reference run live."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, GOLDEN_CASES, load_golden, relerr
from oracle import numpy_batched as nb
from oracle import ref_driver

TOL = 1e-10  # tolerance of the reference's tests (splib/test/sputils_test.py:8)

FORCING_KEYS = ("f_u", "f_v", "f_thl", "f_qt", "f_ql", "f_ps", "ql_ref", "z0m", "z0h", "wthl", "wqt")
TEND_KEYS = ("f_T", "f_SH", "f_QL", "f_QI", "f_U", "f_V", "f_A")


def test_exner_log_identity():
    # splib/test/sputils_test.py:25-28
    a = 2.03947
    assert abs(np.log(nb.exner(a * nb.pref0)) - np.log(a) * nb.rd / nb.cp) < TOL


def test_exner_unity():
    # splib/test/sputils_test.py:31-33
    assert abs(nb.exner(nb.pref0) - 1) < TOL


def test_exner_iexner_inverse():
    # splib/test/sputils_test.py:36-39
    p = 12.03947 * nb.pref0
    assert abs(nb.exner(p) * nb.iexner(p) - 1) < TOL


def test_exner_values_from_reference():
    z = np.load(__import__("os").path.join(__import__("conftest").GOLDEN, "ref_kat.npz"))
    assert np.array_equal(nb.exner(z["p"]), z["exner"])
    assert np.array_equal(nb.iexner(z["p"]), z["iexner"])
    assert nb.exner(2e5) == 1.2191675545735878 and nb.iexner(2e5) == 0.8202318018131288


def test_cloud_fraction_index_mapping_kat():
    # splib/test/spcpl_test.py:10-16 (+ spdummy.py:220-222,319-321)
    z = np.load(__import__("os").path.join(__import__("conftest").GOLDEN, "ref_kat.npz"))
    idx = nb.slab_indices(z["cf_zh"], z["cf_Zh"][None, :])[0]
    assert idx.tolist() == [0, 0, 1, 5, 20] == z["cf_idx"].tolist()
    k = 20
    A = z["cf_Aprofile"][np.clip(idx, 0, k - 1)][::-1]           # what the dummy LES returns, reversed
    assert abs(A[0] - (0.5 + 0.2 * np.cos(6. * (1. - k) / k))) < TOL
    assert abs(A[-1] - (0.5 + 0.2)) < TOL
    assert np.allclose(A, z["cf_A"], rtol=0, atol=0)
    assert np.allclose(z["cf_A"], [0.66694256, 0.51414744, 0.6910673, 0.7, 0.7], atol=1e-8)


@pytest.mark.parametrize("name", GOLDEN_CASES + ["ref_L91_cons"])
def test_oracle_matches_reference_golden(name):
    c = load_golden(name)
    o = nb.set_les_forcings(c["gcm"], c["zf"], c["les"], c["aux"]["PS"], c["dt"], c["f_les"], True)
    for k in FORCING_KEYS:
        assert np.array_equal(o[k], c["out"][k]), k
    for k in ("Tv", "THL", "QT", "Zf"):
        assert np.array_equal(o[k], c["out"][k]), k
    assert np.array_equal(o["Zh"][:, 1:], c["out"]["Zh"])
    assert np.array_equal(o["Zh"], c["out"]["gcm_Zh"])
    assert np.array_equal(nb.slab_indices(c["zh"], o["Zh"]), c["out"]["slab_idx"])
    lp = dict(c["les"])
    lp.update(QL_ice=c["aux"]["QL_ice"], T=c["aux"]["T"], Rhobf=c["aux"]["Rhobf"])
    t = nb.set_gcm_tendencies(c["gcm"], c["zf"], lp, c["A_les"], c["dt"], c["f_gcm"],
                              conservative=c["conservative"], zh=c["zh"])
    for k in TEND_KEYS:
        assert np.array_equal(t[k], c["out"][k]), k
    assert np.array_equal(t["t"], c["out"]["t"])
    assert np.array_equal(t["start_index"], c["out"]["start_index"])
    assert np.array_equal(t["A_d"], c["out"]["A_d"])
    assert np.array_equal(t["ql_water"], c["out"]["ql_water"])
    # zeroing above the LES top (spcpl.py:527-533) and something left below it
    s = int(t["start_index"][0])
    assert np.all(t["f_T"][0, :s] == 0) and np.any(t["f_T"][0, s:] != 0)


def test_bracket_is_what_numpy_interp_uses():
    rng = np.random.default_rng(5)
    xp = np.sort(rng.uniform(0, 4000, 50))
    fp = rng.normal(size=50)
    x = np.concatenate([rng.uniform(-100, 4200, 300), xp[[0, 7, -1]]])
    j = nb.bracket(x, xp)
    ref = np.interp(x, xp, fp)
    mine = np.empty_like(x)
    for i, (xi, ji) in enumerate(zip(x, j)):
        if ji < 0:
            mine[i] = fp[0]
        elif ji >= len(xp) - 1:
            mine[i] = fp[-1]
        else:
            mine[i] = ((fp[ji + 1] - fp[ji]) / (xp[ji + 1] - xp[ji])) * (xi - xp[ji]) + fp[ji]
    assert np.array_equal(mine, ref)        # slope form is bit-exact


def test_searchsorted_semantics():
    # SURVEY.md Appendix B
    zh = 25.0 * np.arange(160)
    ss = lambda v, side: int(np.searchsorted(zh, v, side=side))
    assert [ss(0, "right"), ss(25, "right"), ss(24.999, "right"), ss(3975, "right")] == [1, 2, 1, 160]
    assert [ss(0, "left"), ss(25, "left"), ss(3975, "left"), ss(4000, "left")] == [0, 1, 159, 160]
    Zf = np.array([5000, 3987.5, 3987.4, 100.0])
    assert int(np.searchsorted(-Zf, -3987.5)) == 1


def test_slab_oracle_small():
    rng = np.random.default_rng(3)
    ncol, nk, ny, nx = 2, 5, 4, 6
    vols = {f: rng.normal(size=(ncol, nk, ny, nx)).astype(np.float32) for f in ("THL", "QT", "QL", "U", "V")}
    prof, cnt = nb.slab_reduce(vols, 0.25)
    for c in range(ncol):
        for k in range(nk):
            s = sum(float(vols["U"][c, k, j, i]) for j in range(ny) for i in range(nx))
            assert abs(prof["U"][c, k] - s / (nx * ny)) < 1e-12
            n = sum(1 for j in range(ny) for i in range(nx) if float(vols["QL"][c, k, j, i]) > 0.25)
            assert cnt[c, k] == n
    # layout 1 = same numbers from the transposed (OMUSE) view
    v1 = {f: np.ascontiguousarray(np.transpose(v, (0, 3, 2, 1))) for f, v in vols.items()}
    prof1, cnt1 = nb.slab_reduce(v1, 0.25, layout=1)
    assert np.array_equal(cnt, cnt1)
    assert all(np.allclose(prof[f], prof1[f], rtol=1e-14, atol=1e-15) for f in prof)
    idx = np.array([[0, 2, 2, 4, 9], [1, 1, 3, 5, 5]], dtype=np.int32)
    cs = nb.cloud_project(vols["QL"], idx, 0.25)
    for c in range(ncol):
        k0 = 0
        for r in range(idx.shape[1]):
            k1 = min(max(idx[c, r], k0), nk)
            n = sum(1 for j in range(ny) for i in range(nx)
                    if any(float(vols["QL"][c, k, j, i]) > 0.25 for k in range(k0, k1)))
            assert cs[c, r] == n
            k0 = k1
    assert np.array_equal(cs, nb.cloud_project(v1["QL"], idx, 0.25, layout=1))


@pytest.mark.skipif(not ref_driver.available(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("nlev,nk,seed,dt,f_les,f_gcm,cons", [
    (91, 160, 99, 900.0, 1.0, 1.0, False),
    (19, 160, 7, 900.0, 1.0, 1.0, False),        # T21 test-case levels
    (137, 160, 8, 600.0, 0.25, 2.0, False),      # non-unit factors, another time step
    (91, 160, 9, 900.0, 1.0, 1.0, True),         # conservative coarsening
    (137, 160, 10, 900.0, 0.5, 0.5, True),
    (19, 20, 11, 900.0, 1.0, 1.0, False),        # spdummy-sized LES, dz = 200 m
])
def test_oracle_matches_reference_live(nlev, nk, seed, dt, f_les, f_gcm, cons):
    """The restatement against the UNMODIFIED reference run here (oracle/ref_driver.py), beyond the committed golden
    cases: other seeds (orography, surface pressure, fluxes), level sets, factors, and the conservative option."""
    import synth_les
    from sp_coupler_b200 import synth
    ncol = 3
    zf, zh = synth.les_grid(nk, 25.0 if nk == 160 else 200.0)
    g = synth.make_gcm_columns(ncol, nlev, seed=seed)
    aux = synth.make_les_aux(ncol, nk, seed=seed)
    plan = synth_les.les_volume_plan(g, zf)
    rng = np.random.default_rng(seed)
    lp = {f: plan[f][0] + plan[f][1] * 0.02 * rng.normal(size=plan[f][0].shape) for f in ("THL", "QT", "U", "V")}
    lp["QL"] = np.maximum(4e-6 + 3e-6 * rng.normal(size=(ncol, nk)), 0.0)
    A = rng.integers(0, 65, (ncol, nlev)) / 64.0
    r = ref_driver.run_columns(g, zf, zh, lp, aux, A, dt, f_les, f_gcm, True, conservative=cons)
    o = nb.set_les_forcings(g, zf, lp, aux["PS"], dt, f_les, True)
    for k in FORCING_KEYS:
        assert relerr(o[k], r[k]) <= 1e-13, k
    lp2 = dict(lp, QL_ice=aux["QL_ice"], T=aux["T"], Rhobf=aux["Rhobf"])
    t = nb.set_gcm_tendencies(g, zf, lp2, A, dt, f_gcm, conservative=cons, zh=zh)
    for k in TEND_KEYS:
        assert relerr(t[k], r[k]) <= 1e-13, k
    assert np.array_equal(t["start_index"], r["start_index"])
    assert np.array_equal(nb.slab_indices(zh, r["gcm_Zh"]), r["slab_idx"])


def test_oracle_integrals_match_reference_golden():
    """sputils.integral / interp_c / interp_rho (sputils.py:94-197): the restatement against outputs of the unmodified
    reference (tests/golden/ref_sputils.npz, oracle/make_golden.py), with nk+1 cell edges and with the nk edges the
    coupler actually passes."""
    z = np.load(os.path.join(GOLDEN, "ref_sputils.npz"))
    for tag in ("a", "b"):
        Zh, q, rho = z[tag + "_Zh"], z[tag + "_q"], z[tag + "_rho"]
        for name in ("full", "short"):
            edges = z["%s_%s_z" % (tag, name)]
            for c in range(2):
                assert np.array_equal(nb.interp_c(Zh[c], edges, q[c], rho[c]), z["%s_%s_interp_c" % (tag, name)][c])
                assert np.array_equal(nb.interp_rho(Zh[c], edges, rho[c]), z["%s_%s_interp_rho" % (tag, name)][c])
        edges = z[tag + "_full_z"]
        for (a, b), (plain, weighted) in zip(z[tag + "_ab"], z[tag + "_integral"]):
            assert nb.integral_plain(a, b, edges, q[0]) == plain
            if not np.isnan(weighted):
                assert nb.integral(a, b, edges, q[0], rho[0]) == weighted
        assert nb.integral_plain(-1.0, 10.0, edges, q[0]) is None          # end point outside the range: sputils.py:113-115


def test_slab_mean_matches_the_reference_held_expression():
    """a1/a2 pin: the only slab average written inside the reference tree is
    `X[:, :, k].sum() / (itot * jtot)` over the (itot, jtot, ktot) view (spcpl.py:621,642,650). The fixture evaluates
    that expression literally (oracle/make_golden.py::make_slabmean_golden); the oracle's slab_reduce must agree in
    both memory layouts, for both storage types, and the counts must be exact."""
    z = np.load(os.path.join(GOLDEN, "ref_slabmean.npz"))
    for dt in (np.float64, np.float32):
        ijk = {f: z["vol_" + f].astype(dt)[None] for f in ("THL", "QT", "QL", "U", "V")}          # [1][nx][ny][nk]
        kji = {f: np.ascontiguousarray(np.transpose(v, (0, 3, 2, 1))) for f, v in ijk.items()}     # [1][nk][ny][nx]
        for layout, vols in ((1, ijk), (0, kji)):
            for thr in (0.0, 1e-6):
                prof, cnt = nb.slab_reduce(vols, thr, layout)
                assert np.array_equal(cnt[0], z["cnt_%g" % thr])
            for f in vols:
                assert relerr(prof[f][0], z["mean_" + f]) <= 1e-15, (f, layout, dt)


@pytest.mark.parametrize("nlev,nk,dz", [(19, 20, 200.0), (91, 160, 25.0), (137, 160, 25.0)])
def test_level_window_is_exact_in_the_oracle(nlev, nk, dz):
    """The level window of sp_coupler_b200/pipeline.py, checked on the CPU restatement of the reference alone: cutting
    the GCM columns off above the first level over the LES top (lev0 = min start_index - 1, the host-side bound
    `first_live_level`) changes no forcing, no cloud count and no tendency of the remaining levels - bit for bit - and
    everything above lev0 is zero in the full result (spcpl.py:494-533). This is what lets only 26 of 91 (38 of 137)
    levels cross PCIe / NVLink."""
    import synth_les
    from sp_coupler_b200 import synth
    from sp_coupler_b200.pipeline import GcmStaging, first_live_level, window_columns
    ncol, nx = 5, 8
    zf, zh = synth.les_grid(nk, dz)
    gcm = synth.make_gcm_columns(ncol, nlev, seed=31, dtype=np.float32)
    aux = synth.make_les_aux(ncol, nk, seed=31, dtype=np.float32)
    vols = synth_les.make_les_volumes(gcm, zf, nx, nx, seed=31, dtype=np.float32)
    full = nb.coupling_step(gcm, zf, zh, vols, aux, aux["PS"], 900.0, 1.0, 1.0, True)
    lev0 = first_live_level(gcm["Zgfull"], gcm["Zghalf"][:, -1], float(zf[-1]))
    assert lev0 == max(int(full["tendencies"]["start_index"].min()) - 1, 0) and lev0 > 0
    win = nb.coupling_step(window_columns(gcm, lev0), zf, zh, vols, aux, aux["PS"], 900.0, 1.0, 1.0, True)
    for k in FORCING_KEYS:
        assert np.array_equal(win["forcings"][k], full["forcings"][k]), k
    for k in TEND_KEYS:
        assert np.array_equal(win["tendencies"][k], full["tendencies"][k][:, lev0:]), k
        assert not full["tendencies"][k][:, :lev0].any(), k
    assert np.array_equal(win["tendencies"]["start_index"], full["tendencies"]["start_index"] - lev0)
    assert np.array_equal(win["cntslab"], full["cntslab"][:, :nlev - lev0])         # ascending slabs: the lowest ones
    assert np.array_equal(win["slab_idx"], full["slab_idx"][:, :nlev - lev0])
    assert np.array_equal(win["cnt"], full["cnt"])
    # the packed staging layout of a window is a contiguous prefix of the full-size buffers
    st = GcmStaging(ncol, nlev, __import__("torch").float32, "cpu", pin=False)
    st.set_levels(nlev - lev0)
    st.fill_host(window_columns(gcm, lev0))
    assert st.numel == GcmStaging.numel_for(ncol, nlev - lev0) < st.host_buf.numel()
    assert np.array_equal(st.host["T"].numpy(), gcm["T"][:, lev0:]) and st.host["Zghalf"].shape == (ncol, nlev - lev0 + 1)
