"""The thin PyTorch C++ extension (torch.ops.spcpl_b200.*) must give bit-identical results to the
ctypes binding of the same C ABI, and the oracle's numbers."""
import numpy as np
import pytest

import cases
from conftest import relerr
from sp_coupler_b200.constants import TENDENCIES

pytestmark = pytest.mark.gpu


def test_torch_ops_step_equals_ctypes_binding(cuda_device):
    import torch
    from sp_coupler_b200 import torch_ops
    from sp_coupler_b200.coupler import Coupler
    cpl = Coupler(cuda_device)
    case = cases.host_case(3, 64, 64, 40, 91, np.float32)
    ref = cases.oracle_step(case)
    d = cases.to_device(case, cuda_device)
    slab0, frc0, tnd0 = cases.gpu_step(cpl, d, diagnostics=False)
    frc, slab, tnd = torch_ops.step(d["gcm"], d["zf"], d["zh"], d["vols"], d["aux"], slab0["prof"])
    torch.cuda.synchronize()
    for k in ("prof", "cnt", "mask"):
        assert torch.equal(slab[k], slab0[k]), k
    for k in ("f_u", "f_v", "f_thl", "f_qt", "f_ql", "f_ps", "ql_ref", "wthl", "wqt", "slab_idx"):
        assert torch.equal(frc[k], frc0[k]), k
    for k in ("tend", "A_d", "cntslab", "start_index"):
        assert torch.equal(tnd[k], tnd0[k]), k
    for k in TENDENCIES:
        assert relerr(tnd[k].cpu().numpy(), ref["tendencies"][k]) <= 1e-4, k
    assert np.array_equal(tnd["cntslab"].cpu().numpy(), ref["cntslab"])


def test_torch_ops_errors_are_loud(cuda_device):
    import torch
    from sp_coupler_b200 import torch_ops
    ops = torch_ops.load()
    vols = [torch.zeros((1, 4, 8, 8), device=cuda_device) for _ in range(4)]
    with pytest.raises(RuntimeError):
        ops.slab_reduce(vols, 0, 0.0, torch.zeros((5, 1, 4), dtype=torch.float64, device=cuda_device), None, None)


def test_torch_ops_reject_buffers_of_the_wrong_type_or_size(cuda_device):
    """The registered ops forward raw pointers to the C ABI, so every caller-supplied output is checked first: a
    float32 tendency block for float64 GCM columns, a short cloud mask or a short count buffer must raise instead of
    letting a kernel write out of bounds."""
    import torch
    from sp_coupler_b200 import torch_ops
    ops = torch_ops.load()
    dev = cuda_device
    ncol, nk, ny, nx, nlev = 2, 8, 8, 8, 5
    vols = [torch.zeros((ncol, nk, ny, nx), device=dev) for _ in range(5)]
    prof = torch.zeros((5, ncol, nk), dtype=torch.float64, device=dev)
    mw = ops.mask_words_per_column(0, 0, nx, ny, nk)
    good_cnt = torch.zeros((ncol, nk), dtype=torch.int32, device=dev)
    good_mask = torch.zeros((ncol, mw), dtype=torch.int32, device=dev)
    ops.slab_reduce(vols, 0, 0.0, prof, good_cnt, good_mask)
    with pytest.raises(RuntimeError):
        ops.slab_reduce(vols, 0, 0.0, prof, good_cnt, good_mask[:, :-1].contiguous())          # short mask
    with pytest.raises(RuntimeError):
        ops.slab_reduce(vols, 0, 0.0, prof, good_cnt[:1].contiguous(), good_mask)              # short counts
    with pytest.raises(RuntimeError):
        ops.slab_reduce(vols, 0, 0.0, prof, good_cnt.float(), good_mask)                       # wrong dtype
    g = {k: torch.ones((ncol, nlev), dtype=torch.float64, device=dev) for k in ("U", "V", "T", "SH", "QL", "QI", "Pfull", "A", "Zgfull")}
    g.update({k: torch.ones((ncol, nlev + 1), dtype=torch.float64, device=dev) for k in ("Phalf", "Zghalf")})
    zf = torch.linspace(10.0, 80.0, nk, dtype=torch.float64, device=dev)
    les = {"prof": prof, "QL_ice": torch.zeros((ncol, nk), dtype=torch.float64, device=dev),
           "T": torch.zeros((ncol, nk), dtype=torch.float64, device=dev), "A": torch.zeros((ncol, nlev), dtype=torch.float64, device=dev)}
    ok = {"tend": torch.zeros((ncol, 7, nlev), dtype=torch.float64, device=dev)}
    ops.les_to_gcm(g, zf, None, les, nx, ny, 0, 0, 900.0, 1.0, False, ok)
    with pytest.raises(RuntimeError):      # float32 block for float64 columns: half the bytes the kernel would write
        ops.les_to_gcm(g, zf, None, les, nx, ny, 0, 0, 900.0, 1.0, False, {"tend": ok["tend"].float()})
    with pytest.raises(RuntimeError):
        ops.les_to_gcm(g, zf, None, les, nx, ny, 0, 0, 900.0, 1.0, False, {"tend": ok["tend"][:, :6].contiguous()})
    with pytest.raises(RuntimeError):
        ops.gcm_to_les(g, zf, None, None, None, 900.0, 1.0, False, {"ql_ref": torch.zeros((ncol, nk), device=dev)})   # float32
    torch.cuda.synchronize()
