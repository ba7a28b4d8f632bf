"""Generates tests/golden/ref_*.npz by running the UNMODIFIED reference
(/root/reference/splib/spcpl.py, sputils.py, spdummy.py) under oracle/stubs.

Run in the build container only (`python -m oracle.make_golden`); the fixtures are committed,
so neither the CPU tests nor the GPU box need /root/reference.
"""
import contextlib
import io
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_driver  # noqa: E402
sys.path.insert(1, os.path.join(ROOT, "tests"))
import synth_les  # noqa: E402
from sp_coupler_b200 import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
CASES = [  # name, ncol, nlev, nk, seed, dt, f_les, f_gcm, conservative
    ("ref_L19", 2, 19, 160, 42, 900.0, 1.0, 1.0, False),      # config C1 shape (T21 test case)
    ("ref_L91", 4, 91, 160, 43, 900.0, 1.0, 1.0, False),
    ("ref_L137", 3, 137, 160, 46, 900.0, 0.5, 0.75, False),   # non-unit forcing factors
    ("ref_L19_nk20", 3, 19, 20, 47, 600.0, 1.0, 1.0, False),  # spdummy-sized LES (8x8x20, dz=200)
    ("ref_L91_cons", 3, 91, 160, 48, 900.0, 1.0, 1.0, True),  # --conservative_coarsening
]


def case_inputs(ncol, nlev, nk, seed):
    dz = 25.0 if nk == 160 else 200.0
    zf, zh = synth.les_grid(nk, dz)
    gcm = synth.make_gcm_columns(ncol, nlev, seed=seed)
    aux = synth.make_les_aux(ncol, nk, seed=seed)
    plan = synth_les.les_volume_plan(gcm, zf)
    rng = np.random.default_rng(seed + 7)
    lp = {f: plan[f][0] + plan[f][1] * 0.01 * rng.normal(size=plan[f][0].shape) for f in ("THL", "QT", "U", "V")}
    lp["QL"] = np.maximum(1e-5 * rng.normal(size=(ncol, nk)) + 5e-6, 0.0)
    A = rng.integers(0, 65, (ncol, nlev)) / 64.0
    return zf, zh, gcm, aux, lp, A


def nudge_inputs(seed, nx=16, ny=12, nk=24):
    """Synthetic LES state for spcpl.variability_nudge in the reference's (itot, jtot, ktot) view:
    levels with every regime (bracketed multiplicative root, unsaturated nudge, beta_max -> additive
    noise, untouched)."""
    rng = np.random.default_rng(seed)
    qt_p = 0.008 * np.exp(-np.arange(nk) / 8.0)
    qt = qt_p[None, None, :] + 2.5e-5 * rng.uniform(-1, 1, (nx, ny, nk)) * rng.uniform(0.05, 4, nk)[None, None, :]
    s = rng.uniform(-1.5, 1.5, nk)
    qsat = qt_p[None, None, :] + 2.5e-5 * s[None, None, :] + 1e-6 * rng.normal(size=(nx, ny, nk))
    ql = np.maximum(qt - qsat, 0)
    thl = 290 + rng.normal(size=(nx, ny, nk))
    qt_av, ql_av = qt.mean(axis=(0, 1)), ql.mean(axis=(0, 1))
    ql_ref = ql_av * rng.choice([0.0, 0.5, 1.5, 3.0, 40.0], nk) + rng.choice([0, 0, 2e-6], nk)
    presf = 1e5 * np.exp(-np.arange(nk) * 25 / 7500.)
    return dict(Qsat=qsat, QT=qt, THL=thl, QL=ql), dict(QT=qt_av, QL=ql_av), presf, ql_ref


def make_nudge_golden():
    kji = lambda a: np.ascontiguousarray(np.transpose(a, (2, 1, 0)))
    for name, seed, constT in (("ref_nudge", 8, False), ("ref_nudge_constT", 5, True)):
        f, p, presf, ql_ref = nudge_inputs(seed)
        r = ref_driver.run_variability_nudge(f, p, presf, ql_ref, 900.0, constT, seed=seed + 100)
        blob = dict(DT=900.0, constantT=constT, qt=kji(f["QT"]), qsat=kji(f["Qsat"]), thl=kji(f["THL"]), ql=kji(f["QL"]),
                    qt_av=p["QT"], ql_av=p["QL"], presf=presf, ql_ref=ql_ref, R=np.ascontiguousarray(r["R"].T),
                    out_qt=kji(r["qt"]), out_beta=r["beta"], out_alpha=r["alpha"], out_qt_std=r["qt_std"])
        if constT:
            blob["out_thl"] = kji(r["thl"])
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **blob)
        print(name, "-> beta", np.round(r["beta"], 3))


def make_sputils_golden():
    """sputils.integral / interp_c / interp_rho of the UNMODIFIED reference (sputils.py:94-197) on GCM layer edges
    (descending) over the LES cells, with nk+1 edges (the documented call) and with nk edges (what spcpl.py:479-488
    actually passes: zh_cache has nk entries)."""
    sputils, _, _, _ = ref_driver.load_reference()
    from omuse.units import units
    blob = {}
    for tag, nlev, nk, seed in (("a", 19, 20, 3), ("b", 91, 160, 4)):
        dz = 25.0 if nk == 160 else 200.0
        zf, zh = synth.les_grid(nk, dz)
        gcm = synth.make_gcm_columns(2, nlev, seed=seed)
        Zh = (gcm["Zghalf"] - gcm["Zghalf"][:, -1:]) / 9.81
        rng = np.random.default_rng(seed)
        q = 0.01 * np.exp(-zf / 2000.0)[None, :] * (1 + 0.1 * rng.normal(size=(2, nk)))
        rho = 1.2 * np.exp(-zf / 8000.0)[None, :] * (1 + 0.01 * rng.normal(size=(2, nk)))
        edges = np.append(zh, zh[-1] + dz)                               # nk+1 edges
        for name, z in (("full", edges), ("short", zh)):
            with contextlib.redirect_stdout(io.StringIO()):              # the len(z) != len(q)+1 message of :111-112
                Qc = np.stack([np.asarray(sputils.interp_c(Zh[c] | units.m, z | units.m, q[c] | units.mfu, rho[c] | units.kg))
                               for c in range(2)])
                Rc = np.stack([np.asarray(sputils.interp_rho(Zh[c] | units.m, z | units.m, rho[c] | units.kg))
                               for c in range(2)])
            blob.update({"%s_%s_z" % (tag, name): z, "%s_%s_interp_c" % (tag, name): Qc, "%s_%s_interp_rho" % (tag, name): Rc})
        ab = np.array([[30.0, 410.0], [410.0, 30.0], [0.0, float(zh[-1])], [125.0, 125.0], [12.5, 37.5]])
        I = np.array([[float(sputils.integral(a, b, edges, q[0])), float(sputils.integral(a, b, edges, q[0], rho[0]))]
                      for a, b in ab[:3]] + [[float(sputils.integral(a, b, edges, q[0])), np.nan] for a, b in ab[3:]])
        blob.update({tag + "_Zh": Zh, tag + "_q": q, tag + "_rho": rho, tag + "_ab": ab, tag + "_integral": I})
    np.savez_compressed(os.path.join(GOLDEN, "ref_sputils.npz"), **blob)
    print("ref_sputils ->", len(blob), "arrays")


def make_slabmean_golden():
    """The ONLY statement of a horizontal slab average inside the reference tree, evaluated literally: the nudging code
    forms `X[:, :, k].sum() / (itot * jtot)` over the (itot, jtot, ktot) view of a 3-D field (spcpl.py:621 for R,
    spcpl.py:642,650 for max(qt - qsat, 0)), and `qt.std(axis=(0, 1))` (spcpl.py:741). DALES' own profile routines are
    external, so this expression is what K1's slab means are pinned to; the above-threshold count is
    `(ql[:, :, k] > thr).sum()`. Volumes are stored in the (i, j, k) view, float64 values that are exactly
    representable in float32 (so the same fixture serves both storage types)."""
    rng = np.random.default_rng(77)
    itot, jtot, ktot = 16, 12, 32
    prof = {"THL": 290.0 + 0.01 * np.arange(ktot), "QT": 0.008 * np.exp(-np.arange(ktot) / 10.0),
            "U": 5.0 + 0.1 * np.arange(ktot), "V": -2.0 + 0.05 * np.arange(ktot)}
    amp = {"THL": 0.1, "QT": 2.5e-5, "U": 0.5, "V": 0.5}
    vol = {f: (prof[f][None, None, :] + amp[f] * rng.uniform(-1, 1, (itot, jtot, ktot))).astype(np.float32).astype(np.float64)
           for f in prof}
    qsat = prof["QT"][None, None, :] + 2.5e-5 * rng.uniform(-1.2, 1.2, ktot)[None, None, :]
    vol["QL"] = np.maximum(vol["QT"] - qsat, 0).astype(np.float32).astype(np.float64)
    blob = {}
    for f, v in vol.items():
        blob["vol_" + f] = v
        blob["mean_" + f] = np.array([v[:, :, k].sum() / (itot * jtot) for k in range(ktot)])      # spcpl.py:642 form
    blob["std_QT"] = vol["QT"].std(axis=(0, 1))                                                     # spcpl.py:741
    for thr in (0.0, 1e-6):
        blob["cnt_%g" % thr] = np.array([(vol["QL"][:, :, k] > thr).sum() for k in range(ktot)], dtype=np.int32)
    np.savez_compressed(os.path.join(GOLDEN, "ref_slabmean.npz"), **blob)
    print("ref_slabmean ->", len(blob), "arrays; cloudy cells per level", blob["cnt_0"])


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    make_slabmean_golden()
    make_sputils_golden()
    make_nudge_golden()
    for name, ncol, nlev, nk, seed, dt, fl, fg, cons in CASES:
        zf, zh, gcm, aux, lp, A = case_inputs(ncol, nlev, nk, seed)
        out = ref_driver.run_columns(gcm, zf, zh, lp, aux, A, dt, fl, fg, True, conservative=cons)
        blob = {"zf": zf, "zh": zh, "dt": dt, "f_les": fl, "f_gcm": fg, "conservative": cons, "A_les": A}
        blob.update({"gcm_" + k: v for k, v in gcm.items()})
        blob.update({"aux_" + k: v for k, v in aux.items()})
        blob.update({"les_" + k: v for k, v in lp.items()})
        blob.update({"out_" + k: v for k, v in out.items()})
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **blob)
        print(name, "->", len(blob), "arrays")

    # known-answer vectors of the reference's own tests, evaluated by the reference itself
    sputils, spcpl, spdummy, spio = ref_driver.load_reference()
    les = spdummy.dummy_les(1)
    les.commit_grid()
    les.zh_cache = les.get_zh()                       # as splib.initialize does (splib.py:152)
    les.gcm_Zh = np.array([100000., 1000., 100., 10., 1., 0.])   # splib/test/spcpl_test.py:13
    A = np.asarray(spcpl.get_cloud_fraction(les))
    idx = sputils.searchsorted(les.zh_cache, les.gcm_Zh, side="right")[:-1:][::-1]
    p = np.array([2e5, 2.03947e5, 12.03947e5, 1e5, 5e4, 101325.0])
    np.savez(os.path.join(GOLDEN, "ref_kat.npz"),
             cf_zh=np.asarray(les.zh_cache), cf_Zh=np.asarray(les.gcm_Zh), cf_idx=np.asarray(idx),
             cf_A=A, cf_Aprofile=np.asarray(les.get_profile_field("A")),
             p=p, exner=np.asarray(sputils.exner(p)), iexner=np.asarray(sputils.iexner(p)))
    print("ref_kat -> idx", idx, "A", A)


if __name__ == "__main__":
    main()
