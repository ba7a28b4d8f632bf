"""Pins oracle/nudge.py: brentq against scipy.optimize.brentq (bit-identical roots), and
variability_nudge against golden vectors from the unmodified reference function."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, relerr
from oracle import nudge


def test_brentq_is_scipys():
    sb = pytest.importorskip("scipy.optimize").brentq
    rng = np.random.default_rng(0)
    n = 0
    for t in range(400):
        d = rng.normal(size=50) * rng.uniform(0.1, 3)
        m, c = rng.normal(), rng.uniform(0, 1)
        f = lambda b: np.maximum(b * d + m, 0).sum() / 50 - c
        if f(0) > 0 or f(5) < 0:
            continue
        r, st = nudge.brentq(f, 0, 5)
        assert st == 0 and r == sb(f, 0, 5)
        n += 1
    assert n > 50
    assert nudge.brentq(lambda x: x + 1.0, 0, 5)[1] == -1          # no sign change (scipy: ValueError)
    assert nudge.brentq(lambda x: x, 0, 5) == (0.0, 0)             # f(a) == 0 returns a


@pytest.mark.parametrize("name", ["ref_nudge", "ref_nudge_constT"])
def test_nudge_oracle_matches_reference_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    cT = bool(z["constantT"])
    o = nudge.variability_nudge(z["qt"], z["qsat"], z["qt_av"], z["ql_av"], z["ql_ref"], z["R"], float(z["DT"]), cT,
                                thl=z["thl"], ql=z["ql"], presf=z["presf"])
    # identical algorithm; only numpy's summation order over the strided (i,j,k) view differs
    assert relerr(o["beta"], z["out_beta"]) <= 1e-13
    assert relerr(o["alpha"], z["out_alpha"]) <= 1e-12
    assert relerr(o["qt"] - z["qt"], z["out_qt"] - z["qt"]) <= 1e-12     # relative to the size of the nudge
    assert relerr(o["qt_std"], z["out_qt_std"]) <= 1e-12
    if cT:
        assert relerr(o["thl"] - z["thl"], z["out_thl"] - z["thl"]) <= 1e-12
    st = o["status"]
    assert (st & nudge.ST_MULT).any() and (st & nudge.ST_UNSAT).any() and (st & nudge.ST_ADD).any() and (st == 0).any()
