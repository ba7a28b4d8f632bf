"""GPU parity: the CUDA path, called through the C ABI, against the CPU oracle on the same seeded
inputs, and against the golden vectors produced by the unmodified reference.

Gates: counts, bracket indices, slab indices, start_index: exact integer equality.
Profiles / forcings / tendencies: max|a-b| / max|ref| <= 1e-6 (float64 mode), 1e-4 (float32 mode).
"""
import numpy as np
import pytest

import cases
from conftest import GOLDEN_CASES, load_golden, relerr
from oracle import numpy_batched as nb
from sp_coupler_b200.constants import LES_FIELDS, TENDENCIES

pytestmark = pytest.mark.gpu

RTOL = {np.float64: 1e-6, np.float32: 1e-4}
FORCINGS = ("f_u", "f_v", "f_thl", "f_qt", "f_ql", "f_ps", "ql_ref", "z0m", "z0h", "wthl", "wqt")


@pytest.fixture(scope="module")
def cpl(cuda_device):
    from sp_coupler_b200.coupler import Coupler
    return Coupler(cuda_device)


def n(t):
    return t.detach().cpu().numpy()


def check_step(case, ref, slab, frc, tnd, rtol, mask_expected=True):
    for i, f in enumerate(LES_FIELDS):
        assert relerr(n(slab["prof"][i]), ref["prof"][f]) <= min(rtol, 1e-12), f   # float64 accumulation
    assert np.array_equal(n(slab["cnt"]), ref["cnt"])
    assert np.array_equal(n(frc["slab_idx"]), ref["slab_idx"])
    assert np.array_equal(n(frc["bracket"]), ref["forcings"]["bracket"])
    for k in FORCINGS:
        assert relerr(n(frc[k]), ref["forcings"][k]) <= rtol, k
    for k, o in (("u", "u"), ("v", "v"), ("thl", "thl"), ("qt", "qt"), ("Tv", "Tv"), ("THL", "THL"), ("QT", "QT"),
                 ("Zf", "Zf"), ("Zh", "Zh"), ("ps", "ps")):
        assert relerr(n(frc[k]), ref["forcings"][o]) <= rtol, k
    t = ref["tendencies"]
    if mask_expected:
        assert np.array_equal(n(tnd["cntslab"]), ref["cntslab"])
    assert np.array_equal(n(tnd["start_index"]), t["start_index"])
    assert np.array_equal(n(tnd["bracket"]), t["bracket"])
    assert np.array_equal(n(tnd["bracket_pf"]), t["bracket_pf"])
    for k in TENDENCIES:
        assert relerr(n(tnd[k]), t[k]) <= rtol, k
    assert relerr(n(tnd["t"]), t["t"]) <= rtol
    assert relerr(n(tnd["A_d"]), t["A_d"]) <= rtol


@pytest.mark.parametrize("ncol,nx,ny,nk,nlev,dtype", [
    (2, 64, 64, 160, 19, np.float64),     # C1: T21 example shape
    (4, 64, 64, 160, 91, np.float64),     # C2 shape (fp64 equivalence), reduced ncol
    (4, 64, 64, 160, 91, np.float32),     # C3 shape, reduced ncol
    (3, 32, 32, 160, 137, np.float32),    # C5 shape
    (1, 256, 256, 20, 91, np.float32),    # C4 slab size (256 KB slabs span many TMA chunks)
    (3, 8, 8, 20, 19, np.float64),        # spdummy-sized LES: generic (non-TMA) path
    (2, 10, 10, 20, 19, np.float32),      # slab bytes not a multiple of 16: generic path
    (2, 40, 40, 24, 91, np.float32),      # ragged: slab = 1 full + 1 partial TMA chunk
    (2, 24, 20, 16, 19, np.float64),      # ragged float64 chunks
    (2, 48, 48, 40, 91, np.float32),      # three 4 KB sub-blocks per slab: 24 mask vectors per level (not a divisor of 32)
    (2, 32, 16, 160, 137, np.float32),    # 2 KB slabs: one partial sub-block, 8 mask vectors per level, four levels per warp pass
    (1, 32, 32, 21, 19, np.float32),      # 4 KB slabs, odd slab count per field: one slab per copy instead of pairs
    (5, 32, 16, 24, 19, np.float64),      # 4 KB float64 slabs in pairs
])
def test_coupling_step_matches_oracle(cpl, cuda_device, ncol, nx, ny, nk, nlev, dtype):
    case = cases.host_case(ncol, nx, ny, nk, nlev, dtype)
    ref = cases.oracle_step(case)
    assert ref["cnt"].sum() > 0 and (ref["cnt"] == 0).any()      # non-trivial counts, exact zeros
    d = cases.to_device(case, cuda_device)
    slab, frc, tnd = cases.gpu_step(cpl, d)
    check_step(case, ref, slab, frc, tnd, RTOL[dtype])
    if dtype == np.float64:
        # float64 mode is far tighter than the gate: only pow differs from numpy (<= 2 ulp)
        for k in TENDENCIES:
            assert relerr(n(tnd[k]), ref["tendencies"][k]) <= 1e-12, k
        for k in FORCINGS:
            assert relerr(n(frc[k]), ref["forcings"][k]) <= 1e-12, k


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape", [(2, 16, 12, 160, 19), (2, 6, 5, 21, 19), (3, 32, 32, 40, 91),
                                   (2, 4, 4, 288, 19)])     # nine mask words per point: generic projection kernel
def test_ijk_layout_matches_oracle(cpl, cuda_device, dtype, shape):
    """[ncol][nx][ny][nk] (OMUSE view, k fastest): whole step incl. the projected cloud cover from
    the per-point bit mask; odd nk exercises the scalar (non-vectorised) variant."""
    ncol, nx, ny, nk, nlev = shape
    case = cases.host_case(ncol, nx, ny, nk, nlev, dtype, layout=1)
    ref = cases.oracle_step(case)
    d = cases.to_device(case, cuda_device)
    slab, frc, tnd = cases.gpu_step(cpl, d, layout=1)
    check_step(case, ref, slab, frc, tnd, RTOL[dtype])
    A, cs = cpl.cloud_fraction(slab, frc["slab_idx"])
    assert np.array_equal(n(cs), ref["cntslab"])
    slab2 = cpl.slab_reduce(d["vols"], layout="ijk", want_mask=False)
    assert np.array_equal(n(slab2["cnt"]), ref["cnt"])


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape", [
    (40, 8, 8, 160, 19),     # 200 items > 148 CTAs: several items per CTA; 6 chunks per item < 12 warps
    (3, 16, 16, 128, 19),    # one slot per lane (float32), two (float64)
    (3, 16, 16, 96, 19),     # three-slot period
    (2, 16, 16, 64, 19),     # two horizontal points per period
    (2, 8, 8, 256, 19),      # two / four slots
    (2, 64, 64, 160, 91),    # C3 column shape
    (1, 4, 1, 160, 19),      # a single period per item: eleven warps have nothing to do
    (1, 256, 256, 32, 19),   # C4 slab size: 65536 points per item, 1024 chunks per item
])
def test_ijk_tma_path_matches_oracle(cpl, cuda_device, dtype, shape):
    """IJK shapes that take the per-warp TMA ring kernel (whole 16-byte vectors per point, period of at most
    5 x 32 vectors, nk % 32 == 0 for the mask): means, counts, per-point mask -> projected cloud cover, and the
    same bits from the CTA-per-item kernel it replaces."""
    import torch
    ncol, nx, ny, nk, nlev = shape
    case = cases.host_case(ncol, nx, ny, nk, nlev, dtype, layout=1, dz=25.0)
    ref = cases.oracle_step(case)
    d = cases.to_device(case, cuda_device)
    slab, frc, tnd = cases.gpu_step(cpl, d, layout=1)
    check_step(case, ref, slab, frc, tnd, RTOL[dtype])
    A, cs = cpl.cloud_fraction(slab, frc["slab_idx"])
    assert np.array_equal(n(cs), ref["cntslab"])
    # the CTA-per-item kernel serves whatever the TMA plan rejects, e.g. volumes that are not 16-byte aligned:
    # the same data one element into a larger buffer takes it
    shifted = []
    for v in d["vols"]:
        flat = torch.empty(v.numel() + 1, dtype=v.dtype, device=v.device)
        flat[1:].copy_(v.reshape(-1))
        shifted.append(flat[1:].view(v.shape))
    assert all(t.data_ptr() % 16 != 0 for t in shifted)
    old = cpl.slab_reduce(shifted, layout="ijk", want_mask=True)
    assert torch.equal(old["cnt"], slab["cnt"]) and torch.equal(old["mask"], slab["mask"])
    assert relerr(n(old["prof"]), n(slab["prof"])) <= 1e-14
    again = cpl.slab_reduce(d["vols"], layout="ijk", want_mask=True)
    for k in ("prof", "cnt", "mask"):
        assert torch.equal(again[k], slab[k]), k   # fixed summation order: bit-reproducible


@pytest.mark.parametrize("thr", [-1.0, 0.0, 1e-5])
def test_cloud_threshold_semantics(cpl, cuda_device, thr):
    """strict '>' on the float64 value; a negative threshold must not count chunk padding."""
    case = cases.host_case(2, 40, 40, 24, 19, np.float32)
    ref = cases.oracle_step(case, ql_thresh=thr)
    d = cases.to_device(case, cuda_device)
    slab, frc, tnd = cases.gpu_step(cpl, d, ql_thresh=thr)
    assert np.array_equal(n(slab["cnt"]), ref["cnt"])
    assert np.array_equal(n(tnd["cntslab"]), ref["cntslab"])
    if thr < 0:
        assert (ref["cnt"] == 1600).all()


@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("thr", [1e-5, 2.5e-4, float(np.float32(1e-5)), -3e-7])
def test_cloud_threshold_at_float32_neighbours(cpl, cuda_device, thr, layout):
    """float32 volumes are compared in float32 against the largest float32 <= thr; that must count exactly the
    cells with float64(ql) > thr, also when ql sits on the float32 values right around the threshold."""
    import torch
    rng = np.random.default_rng(7)
    ncol, nx, ny, nk = 2, 32, 32, 32
    t32 = np.float32(thr)
    cand = np.array([t32, np.nextafter(t32, np.float32(-np.inf)), np.nextafter(t32, np.float32(np.inf)),
                     np.float32(0.0), np.float32(2 * abs(thr))], dtype=np.float32)
    ql = cand[rng.integers(0, len(cand), size=(ncol, nk, ny, nx))]
    expect = np.count_nonzero(ql.astype(np.float64) > thr, axis=(2, 3)).astype(np.int32)
    vols = [torch.zeros((ncol, nk, ny, nx), dtype=torch.float32, device=cuda_device) for _ in range(5)]
    vols[2] = torch.from_numpy(ql).to(cuda_device)
    if layout == 1:
        vols = [v.permute(0, 3, 2, 1).contiguous() for v in vols]
    slab = cpl.slab_reduce(vols, layout="kji" if layout == 0 else "ijk", ql_thresh=thr)
    assert np.array_equal(n(slab["cnt"]), expect)
    idx = torch.full((ncol, 4), nk, dtype=torch.int32, device=cuda_device)      # one slab holding every level
    _, cs = cpl.cloud_fraction(slab, idx)
    assert np.array_equal(n(cs)[:, 0], np.count_nonzero((ql.astype(np.float64) > thr).any(axis=1), axis=(1, 2)))


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_profile_kernels_match_reference_golden(cpl, cuda_device, name):
    """K2/K3 against outputs of the UNMODIFIED reference (tests/golden, oracle/make_golden.py)."""
    import torch
    c = load_golden(name)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda_device)
    gcm = {k: t(v) for k, v in c["gcm"].items()}
    zf, zh = t(c["zf"]), t(c["zh"])
    prof = t(np.stack([c["les"][f] for f in LES_FIELDS]))
    frc = cpl.gcm_to_les(gcm, zf, zh, prof, t(c["aux"]["PS"]), c["dt"], c["f_les"], True, diagnostics=True)
    out = c["out"]
    for k in FORCINGS:
        assert relerr(n(frc[k]), out[k]) <= 1e-12, k
    for k in ("Tv", "THL", "QT", "Zf"):
        assert relerr(n(frc[k]), out[k]) <= 1e-12, k
    assert np.array_equal(n(frc["Zh"]), out["gcm_Zh"])          # heights are bit-exact (IEEE sub, div)
    assert np.array_equal(n(frc["Zf"]), out["gcm_Zf"])
    assert np.array_equal(n(frc["slab_idx"]), out["slab_idx"])
    aux = {k: t(v) for k, v in c["aux"].items()}
    tnd = cpl.les_to_gcm(gcm, zf, zh, {"prof": prof}, aux, None, c["dt"], c["f_gcm"], A=t(c["A_les"]), diagnostics=True)
    for k in TENDENCIES:
        assert relerr(n(tnd[k]), out[k]) <= 1e-12, k
    assert np.array_equal(n(tnd["start_index"]), out["start_index"])
    assert relerr(n(tnd["t"]), out["t"]) <= 1e-12
    assert np.array_equal(n(tnd["A_d"]), out["A_d"])


def test_conservative_coarsening_matches_reference_golden(cpl, cuda_device):
    """sputils.interp_c path (--conservative_coarsening), spcpl.py:479-488."""
    import torch
    c = load_golden("ref_L91_cons")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda_device)
    gcm = {k: t(v) for k, v in c["gcm"].items()}
    prof = t(np.stack([c["les"][f] for f in LES_FIELDS]))
    aux = {k: t(v) for k, v in c["aux"].items()}
    tnd = cpl.les_to_gcm(gcm, t(c["zf"]), t(c["zh"]), {"prof": prof}, aux, None, c["dt"], c["f_gcm"],
                         conservative=True, A=t(c["A_les"]))
    for k in TENDENCIES:
        assert relerr(n(tnd[k]), c["out"][k]) <= 1e-6, k
        assert relerr(n(tnd[k]), c["out"][k]) <= 1e-11, k      # only the summation order differs
    assert np.array_equal(n(tnd["start_index"]), c["out"]["start_index"])


def test_sputils_helpers(cpl, cuda_device):
    import torch
    rng = np.random.default_rng(11)
    nb_, np_, nx = 5, 37, 160
    xp = np.sort(rng.uniform(0, 4000, (nb_, np_)), axis=1)
    fp = rng.normal(size=(nb_, np_))
    x = np.concatenate([rng.uniform(-50, 4100, nx - 3), xp[0, [0, 5, -1]]])
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda_device)
    out, br = cpl.interp(t(x), t(xp), t(fp), want_bracket=True)
    ref = np.stack([np.interp(x, xp[i], fp[i]) for i in range(nb_)])
    assert np.array_equal(n(out), ref)                          # bit-exact in float64 (--fmad=false)
    assert np.array_equal(n(br), np.stack([nb.bracket(x, xp[i]) for i in range(nb_)]))
    for side in ("left", "right"):
        ss = cpl.searchsorted(t(xp), t(x), side=side)
        assert np.array_equal(n(ss), np.stack([np.searchsorted(xp[i], x, side=side) for i in range(nb_)]))
    p = rng.uniform(1e3, 1.05e5, 1000)
    assert relerr(n(cpl.exner(t(p))), nb.exner(p)) <= 1e-14
    assert relerr(n(cpl.exner(t(p), inverse=True)), nb.iexner(p)) <= 1e-14
    # the reference's own known-answer tests (splib/test/sputils_test.py:25-39)
    a = 2.03947
    e = n(cpl.exner(t(np.array([a * nb.pref0, nb.pref0, 12.03947 * nb.pref0]))))
    ie = n(cpl.exner(t(np.array([12.03947 * nb.pref0])), inverse=True))
    assert abs(np.log(e[0]) - np.log(a) * nb.rd / nb.cp) < 1e-10
    assert abs(e[1] - 1) < 1e-10
    assert abs(e[2] * ie[0] - 1) < 1e-10


def test_cloud_fraction_index_mapping_kat(cpl, cuda_device):
    """splib/test/spcpl_test.py:10-16: zh=(k+0.5)*200, Zh=[1e5,1e3,100,10,1,0] -> [0,0,1,5,20]."""
    import os
    import torch
    import conftest
    z = np.load(os.path.join(conftest.GOLDEN, "ref_kat.npz"))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda_device)
    ss = n(cpl.searchsorted(t(z["cf_zh"][None, :]), t(z["cf_Zh"]), side="right"))[0]
    assert ss[:-1][::-1].tolist() == [0, 0, 1, 5, 20]


def test_set_les_state_bit_identical_to_host_generator(cpl, cuda_device):
    import torch
    import synth_les
    from sp_coupler_b200 import synth
    rng = np.random.default_rng(2)
    for dtype, td in ((np.float32, torch.float32), (np.float64, torch.float64)):
        for (ncol, nk, ny, nx) in ((3, 7, 8, 12), (2, 5, 3, 5)):
            prof = rng.normal(300, 5, (ncol, nk))
            sub = prof + 0.05 * rng.normal(size=(ncol, nk))
            t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda_device)
            v = n(cpl.set_les_state(t(prof), 0.1, 3, nx, ny, seed=42, col0=5, dtype=td))
            assert np.array_equal(v, synth_les.les_state_volume(prof, 0.1, 3, nx, ny, seed=42, col0=5, dtype=dtype))
            v = n(cpl.set_les_state(t(prof), 0.1, 1, nx, ny, seed=7, col0=0, sub=t(sub), clamp0=True, dtype=td))
            h = synth_les.les_state_volume(prof, 0.1, 1, nx, ny, seed=7, col0=0, sub=sub, clamp0=True, dtype=dtype)
            assert np.array_equal(v, h) and (v == 0).any() and (v > 0).any()
            # every SUB / CLAMP instantiation, the reference amplitudes (spcpl.py:285-291), no noise, and an
            # amplitude too close to the subnormals for the slab kernel's exact-scaling form (generic kernel)
            for amp in (0.5, 2.5e-5, 0.0, 3e-305):
                for use_sub, clamp in ((False, False), (True, False), (False, True), (True, True)):
                    kw = dict(seed=11, col0=2, clamp0=clamp)
                    v = n(cpl.set_les_state(t(prof - 300.0), amp, 2, nx, ny, sub=t(sub - 300.0) if use_sub else None,
                                            dtype=td, **kw))
                    h = synth_les.les_state_volume(prof - 300.0, amp, 2, nx, ny, sub=(sub - 300.0) if use_sub else None,
                                               dtype=dtype, **kw)
                    assert np.array_equal(v, h), (amp, use_sub, clamp)


def test_errors_are_loud(cpl, cuda_device):
    import torch
    case = cases.host_case(1, 8, 8, 20, 19, np.float32)
    d = cases.to_device(case, cuda_device)
    with pytest.raises((RuntimeError, ValueError, TypeError)):
        cpl.slab_reduce([v.cpu() for v in d["vols"]])                       # CPU tensors are rejected
    with pytest.raises(RuntimeError):
        cpl.gcm_to_les(d["gcm"], d["zf"], d["zh"], None, d["aux"]["PS"], dt=0.0)   # dt == 0
    bad = dict(d["gcm"])
    bad["T"] = bad["T"].double()
    with pytest.raises((RuntimeError, ValueError, TypeError)):
        cpl.gcm_to_les(bad, d["zf"], d["zh"])


def test_empty_batch(cpl, cuda_device):
    import torch
    vols = [torch.empty((0, 20, 8, 8), device=cuda_device) for _ in range(5)]
    slab = cpl.slab_reduce(vols)
    assert slab["prof"].shape == (5, 0, 20)


@pytest.mark.parametrize("ncol,nx,ny,nk,dtype", [(3, 40, 40, 24, np.float32), (2, 10, 10, 20, np.float32),
                                                 (2, 24, 20, 16, np.float64), (2, 64, 64, 21, np.float32)])
def test_no_out_of_bounds_writes(cpl, cuda_device, ncol, nx, ny, nk, dtype):
    """compute-sanitizer is closed on this GPU pool, so output buffers are carved out of larger
    sentinel-filled allocations and the guard bands are checked after the kernels ran."""
    import torch
    case = cases.host_case(ncol, nx, ny, nk, 19, dtype)
    d = cases.to_device(case, cuda_device)
    td = torch.float32 if dtype == np.float32 else torch.float64
    G = 64

    def guarded(shape, dt, fill):
        numel = int(np.prod(shape))
        buf = torch.full((numel + 2 * G,), fill, dtype=dt, device=cuda_device)
        return buf, buf[G:G + numel].view(*shape)

    mw = cpl.mask_words_per_column(td, "kji", nx, ny, nk)
    bp, prof = guarded((5, ncol, nk), torch.float64, -7.0)
    bc, cnt = guarded((ncol, nk), torch.int32, -7)
    bm, mask = guarded((ncol, mw), torch.int32, -7)
    slab = cpl.slab_reduce(d["vols"], out=dict(prof=prof, cnt=cnt, mask=mask))
    frc = cpl.gcm_to_les(d["gcm"], d["zf"], d["zh"], slab["prof"], d["aux"]["PS"], 900.0, 1.0, True)
    bt, tend = guarded((ncol, 7, 19), td, -7.0)
    cpl.les_to_gcm(d["gcm"], d["zf"], d["zh"], slab, d["aux"], frc["slab_idx"], 900.0, 1.0, tend_out=tend)
    bv, vol = guarded((ncol, nk, ny, nx), td, -7.0)
    cpl.set_les_state(slab["prof"][0].contiguous(), 0.1, 0, nx, ny, dtype=td, out=vol)
    torch.cuda.synchronize()
    for buf, view in ((bp, prof), (bc, cnt), (bm, mask), (bt, tend), (bv, vol)):
        assert bool((buf[:G] == -7).all()) and bool((buf[-G:] == -7).all())
    assert not bool((prof == -7.0).any()) and not bool((tend == -7.0).any()) and not bool((vol == -7.0).any())
    ref = cases.oracle_step(case)
    assert np.array_equal(n(cnt), ref["cnt"])


def test_c_abi_status_codes(cuda_device):
    """Raw C-ABI calls (ctypes, no wrapper): status codes and spc_last_error() as documented in
    include/spcpl_b200.h: 0 ok, <0 invalid argument / alignment / unsupported, message set."""
    import ctypes as C
    import torch
    from sp_coupler_b200 import _abi
    L = _abi.lib()
    h = C.c_void_p()
    assert L.spc_create(C.byref(h), cuda_device.index) == 0
    assert L.spc_create(C.byref(C.c_void_p()), 9999) == -1 and b"out of range" in L.spc_last_error()
    nk, ny, nx = 4, 16, 16
    vols = [torch.zeros((1, nk, ny, nx), device=cuda_device) for _ in range(5)]
    prof = torch.zeros((5, 1, nk), dtype=torch.float64, device=cuda_device)
    arr = (C.c_void_p * 5)(*[v.data_ptr() for v in vols])
    call = lambda a, dtype=0, layout=0, handle=h, p=prof.data_ptr(): L.spc_slab_reduce(
        handle, a, dtype, layout, 1, nx, ny, nk, 0.0, C.c_void_p(p), None, None, None)
    assert call(arr) == 0
    assert call(arr, dtype=7) == _abi_code("SPC_ERR_ARG") and b"dtype" in L.spc_last_error()
    assert call(arr, layout=5) == -1
    assert call(arr, p=0) == -1 and b"NULL" in L.spc_last_error()
    assert call(arr, handle=None) == -4                                       # SPC_ERR_HANDLE
    bad = (C.c_void_p * 5)(*[v.data_ptr() + (4 if i == 2 else 0) for i, v in enumerate(vols)])
    assert call(bad) == -2 and b"16-byte aligned" in L.spc_last_error()       # SPC_ERR_ALIGN (TMA bulk path)
    nullv = (C.c_void_p * 5)(vols[0].data_ptr(), None, vols[2].data_ptr(), vols[3].data_ptr(), vols[4].data_ptr())
    assert call(nullv) == -1
    torch.cuda.synchronize()
    assert L.spc_destroy(h) == 0
    assert L.spc_destroy(None) == -4


def _abi_code(name):
    import re
    import os
    import conftest
    src = open(os.path.join(conftest.ROOT, "include", "spcpl_b200.h")).read()
    return int(re.search(name + r"\s*=\s*(-?\d+)", src).group(1))


@pytest.mark.parametrize("ncol", [2, 48])
def test_cuda_graph_step_is_bit_identical(cpl, cuda_device, ncol):
    """CouplingPipeline.capture(): the replayed graph (K2 -> K1 -> K3 with its fused projection) gives the bits of the
    eager step, also after the inputs changed in place (new GCM profiles uploaded, LES volumes rewritten)."""
    import torch
    import synth_les
    from sp_coupler_b200 import synth
    from sp_coupler_b200.pipeline import CouplingPipeline
    nx, nk, nlev = 16, 160, 91
    zf, zh = synth.les_grid(nk)
    gcm = synth.make_gcm_columns(ncol, nlev, seed=3, dtype=np.float32)
    aux = {k: torch.from_numpy(v).to(cuda_device) for k, v in synth.make_les_aux(ncol, nk, seed=3, dtype=np.float32).items()}
    pipes = []
    for _ in range(2):
        p = CouplingPipeline(cpl, zf, zh, ncol, nlev, torch.float32)
        p.staging.fill_host(gcm)
        p.staging.upload()
        p.attach_les(synth_les.device_les_volumes(cpl, gcm, zf, nx, nx, seed=3), aux)
        p.les_profiles()
        pipes.append(p)
    eager, graphed = pipes
    graphed.capture(900.0, 1.0, 0.5)
    l0 = cpl.launches
    for it in range(3):
        if it == 1:       # change the inputs in place: both pipelines must follow
            gcm2 = synth.make_gcm_columns(ncol, nlev, seed=9, dtype=np.float32)
            for p in pipes:
                p.staging.fill_host(gcm2)
                p.vols[1].mul_(1.001)
        f1, t1 = eager.step_host(900.0, 1.0, 0.5)
        f2, t2 = graphed.step_host(900.0, 1.0, 0.5)
        assert torch.equal(t1, t2)
        for k in ("f_u", "f_v", "f_thl", "f_qt", "f_ql", "f_ps", "slab_idx"):
            assert torch.equal(f1[k], f2[k]), k
        for k in ("prof", "cnt", "mask"):
            assert torch.equal(eager.slab[k], graphed.slab[k]), k
    assert cpl.launches - l0 == 2 * 3 * 3      # both paths count K2, K1, K3 per step


@pytest.mark.parametrize("nlev,graph", [(19, False), (91, False), (91, True), (137, True)])
def test_level_window_host_step(cpl, cuda_device, nlev, graph):
    """The host-facing step on the live level window (pipeline.py, "Level window"): the GCM columns are cut off above
    the first level over the LES top, only that window is uploaded, K3 itself stores the compact tendency block into
    pinned host memory and raises the completion flag. The block equals the full-level result from lev0 on (bit for
    bit), everything above lev0 is zero in the full result, the forcings are identical, and the host-side bound is
    one level looser than the kernel's start_index."""
    import torch
    import synth_les
    from sp_coupler_b200 import synth
    from sp_coupler_b200.pipeline import CouplingPipeline
    ncol, nx, nk = 12, 16, 160
    zf, zh = synth.les_grid(nk)
    gcm = synth.make_gcm_columns(ncol, nlev, seed=21, dtype=np.float32)
    aux = {k: torch.from_numpy(v).to(cuda_device) for k, v in synth.make_les_aux(ncol, nk, seed=21, dtype=np.float32).items()}
    vols = synth_les.device_les_volumes(cpl, gcm, zf, nx, nx, seed=21)
    full_pipe = CouplingPipeline(cpl, zf, zh, ncol, nlev, torch.float32)
    full_pipe.staging.fill_host(gcm)
    full_pipe.staging.upload()
    full_pipe.attach_les(vols, aux)
    full_pipe.les_profiles()
    win = CouplingPipeline(cpl, zf, zh, ncol, nlev, torch.float32)
    win.attach_les(vols, aux)
    win.les_profiles()
    lev0 = win.stage_host(gcm)                     # picks the window, packs the cut columns into pinned staging
    win.bind_host_output()                         # K3 -> pinned host memory + flag
    assert win.nlw == nlev - lev0 and win.staging.nbytes < full_pipe.staging.nbytes or lev0 == 0
    if graph:       # the host-facing step as ONE launch: H2D of the window + K2 -> K1 -> K3 in one graph
        win.capture(900.0, 1.0, 1.0, upload=True)
    for it in range(3):
        f_full, t_full = full_pipe.step_host(900.0, 1.0, 1.0)
        t_full = t_full.clone()
        f_win, t_win = win.step_host(900.0, 1.0, 1.0)
        assert t_win.shape == (ncol, 7, nlev - lev0) and t_win.is_pinned() and t_win.is_contiguous()
        assert torch.equal(t_win, t_full[:, :, lev0:]), it
        assert not t_full[:, :, :lev0].any()
        for k in ("f_u", "f_v", "f_thl", "f_qt", "f_ql", "f_ps", "ql_ref", "wthl", "wqt"):
            assert torch.equal(f_full[k], f_win[k]), k
        assert torch.equal(f_full["slab_idx"][:, :nlev - lev0], f_win["slab_idx"])
        assert win.sync_error() == 0
    frc = full_pipe.forcings(900.0, 1.0)
    tnd = cpl.les_to_gcm(full_pipe.gcm, full_pipe.zf, full_pipe.zh, full_pipe.slab, full_pipe.aux, frc["slab_idx"], 900.0, 1.0)
    assert lev0 == max(int(tnd["start_index"].min()) - 1, 0)
    assert t_full[:, :, lev0 + 1:].any()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_slab_reduce_matches_the_reference_held_expression(cpl, cuda_device, dtype):
    """a1/a2 pin: K1 against `X[:, :, k].sum() / (itot * jtot)` (the one slab average written in the reference tree,
    spcpl.py:621,642,650; fixture tests/golden/ref_slabmean.npz) in both memory layouts; counts exact for two thresholds."""
    import os
    import torch
    import conftest
    z = np.load(os.path.join(conftest.GOLDEN, "ref_slabmean.npz"))
    fields = ("THL", "QT", "QL", "U", "V")
    ijk = [torch.from_numpy(np.ascontiguousarray(z["vol_" + f].astype(dtype)[None])).to(cuda_device) for f in fields]
    kji = [v.permute(0, 3, 2, 1).contiguous() for v in ijk]
    for layout, vols in (("ijk", ijk), ("kji", kji)):
        for thr in (0.0, 1e-6):
            slab = cpl.slab_reduce(vols, layout=layout, ql_thresh=thr, want_mask=True)
            assert np.array_equal(n(slab["cnt"][0]), z["cnt_%g" % thr]), (layout, thr)
        for i, f in enumerate(fields):
            assert relerr(n(slab["prof"][i, 0]), z["mean_" + f]) <= 1e-12, (f, layout)
