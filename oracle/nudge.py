"""CPU oracle of the qt-variability nudging (spcpl.variability_nudge, splib/spcpl.py:613-744) —
TEST INFRASTRUCTURE, NOT PRODUCT.

The root finder is scipy.optimize.brentq (spcpl.py:10,672,708), a third-party dependency that is not
vendored in the reference tree and not pinned there (README pip lists only name "scipy"); the
version installed here is scipy 1.18.1. `brentq` below restates its published algorithm
(scipy/optimize/Zeros/brentq.c; defaults xtol=2e-12, rtol=4*eps, maxiter=100 from
scipy/optimize/_zeros_py.py) and is pinned against scipy itself in tests/test_oracle_nudge.py
(bit-identical roots). `variability_nudge` is pinned against golden vectors produced by the
unmodified reference function (oracle/make_golden.py -> tests/golden/ref_nudge_*.npz).
"""
import numpy as np

XTOL = 2e-12
RTOL = 4 * np.finfo(float).eps
MAXITER = 100
rlv, cp, rd, pref0 = 2.53e6, 1004.0, 287.04, 1.0e5     # sputils.py:14-18


def brentq(f, xa, xb, xtol=XTOL, rtol=RTOL, maxiter=MAXITER):
    """scipy.optimize.brentq (Zeros/brentq.c). Returns (root, status): status 0 converged,
    -1 no sign change (scipy raises ValueError), -2 not converged in maxiter."""
    xpre, xcur = float(xa), float(xb)
    xblk = fblk = spre = scur = 0.0
    fpre, fcur = f(xpre), f(xcur)
    if fpre == 0:
        return xpre, 0
    if fcur == 0:
        return xcur, 0
    if np.signbit(fpre) == np.signbit(fcur):
        return 0.0, -1
    for _ in range(maxiter):
        if fpre != 0 and fcur != 0 and np.signbit(fpre) != np.signbit(fcur):
            xblk, fblk = xpre, fpre
            spre = scur = xcur - xpre
        if abs(fblk) < abs(fcur):
            xpre, xcur, xblk = xcur, xblk, xcur
            fpre, fcur, fblk = fcur, fblk, fcur
        delta = (xtol + rtol * abs(xcur)) / 2
        sbis = (xblk - xcur) / 2
        if fcur == 0 or abs(sbis) < delta:
            return xcur, 0
        if abs(spre) > delta and abs(fcur) < abs(fpre):
            if xpre == xblk:
                stry = -fcur * (xcur - xpre) / (fcur - fpre)                     # interpolate
            else:
                dpre = (fpre - fcur) / (xpre - xcur)                             # extrapolate
                dblk = (fblk - fcur) / (xblk - xcur)
                stry = -fcur * (fblk * dblk - fpre * dpre) / (dblk * dpre * (fblk - fpre))
            if 2 * abs(stry) < min(abs(spre), 3 * abs(sbis) - delta):
                spre, scur = scur, stry                                          # good short step
            else:
                spre = scur = sbis                                               # bisect
        else:
            spre = scur = sbis                                                   # bisect
        xpre, fpre = xcur, fcur
        if abs(scur) > delta:
            xcur += scur
        else:
            xcur += delta if sbis > 0 else -delta
        fcur = f(xcur)
    return xcur, -2


BETA_MIN, BETA_MAX = 0.0, 5.0      # spcpl.py:654-655
A_MIN, A_MAX = 0.0, 5.0            # spcpl.py:702-703

# status bits reported per (column, level)
ST_MULT = 1        # multiplicative nudge from a Brent root        (spcpl.py:672)
ST_UNSAT = 2       # nudged to barely unsaturated                  (spcpl.py:675-691)
ST_ADD = 4         # additive noise from a Brent root              (spcpl.py:707-714)
ST_NOBRACKET = 8   # multiplicative bracket failed -> beta_max     (spcpl.py:665-669)
ST_ADD_FAIL = 16   # additive root not bracketed (scipy would raise ValueError)


def variability_nudge(qt, qsat, qt_av, ql_av, ql_ref, R, DT, constantT=False, thl=None, ql=None, presf=None):
    """spcpl.py:613-744 for ONE LES, KJI layout: qt, qsat (thl, ql) are [nk][ny][nx] float arrays,
    qt_av, ql_av, ql_ref, presf [nk], R [ny][nx] with zero mean (spcpl.py:620-621).
    Returns dict(qt, thl, beta, alpha, qt_std, status); inputs are not modified."""
    qt = np.array(qt, dtype=np.float64)
    qsat = np.asarray(qsat, dtype=np.float64)
    nk = qt.shape[0]
    npts = qt.shape[1] * qt.shape[2]
    thl_o = None if thl is None else np.array(thl, dtype=np.float64)
    beta = np.ones(nk)                                                          # :657
    status = np.zeros(nk, dtype=np.int32)
    for k in range(nk):
        q, qs = qt[k], qsat[k]

        def ql_diff(b):                                                          # :641-643
            return np.maximum(b * (q - qt_av[k]) + qt_av[k] - qs, 0).sum() / npts - ql_ref[k]

        def ql_diff_additive(a):                                                 # :648-651
            return np.maximum(q + (a * R) - qs, 0).sum() / npts - ql_ref[k]

        if ql_ref[k] > 1e-9:                                                     # :661
            q_min, q_max = ql_diff(BETA_MIN), ql_diff(BETA_MAX)                  # :663-664
            if q_min > 0 or q_max < 0:                                           # :665
                beta[k] = BETA_MAX                                               # :669
                status[k] |= ST_NOBRACKET
            else:
                beta[k], _ = brentq(ql_diff, BETA_MIN, BETA_MAX)                 # :672
                status[k] |= ST_MULT
        elif ql_av[k] > ql_ref[k]:                                               # :675
            d = q - qs
            j, i = np.unravel_index(np.argmax(d.T), d.T.shape)[::-1]             # first max in (i, j) C order, :677
            beta[k] = (qs[j, i] - qt_av[k]) / (q[j, i] - qt_av[k])               # :678
            status[k] |= ST_UNSAT
            if beta[k] < 0:                                                      # :688-691
                beta[k] = 1
        else:
            continue                                                             # :693
        if beta[k] >= BETA_MAX:                                                  # :698
            if ql_ref[k] > ql_av[k]:                                             # :707
                a, st = brentq(ql_diff_additive, A_MIN, A_MAX)                   # :708
                if st == -1:
                    status[k] |= ST_ADD_FAIL
                else:
                    qt[k] = q + a * R                                            # :711-714
                    status[k] |= ST_ADD
            beta[k] = 1                                                          # :717
        else:
            qt[k] = q + (beta[k] - 1) * (q - qt_av[k])                           # :719-720
        if constantT:                                                            # :721-728
            ql_target = np.maximum(qt[k] - qs, 0)
            dQL = ql_target - np.asarray(ql[k], dtype=np.float64)
            dTHL = -rlv / (cp * (presf[k] / pref0) ** (rd / cp)) * dQL
            thl_o[k] = thl_o[k] + dTHL
    alpha = np.log(beta) / DT                                                    # :737
    qt_std = qt.reshape(nk, -1).std(axis=1)                                      # :741
    return dict(qt=qt, thl=thl_o, beta=beta, alpha=alpha, qt_std=qt_std, status=status)
