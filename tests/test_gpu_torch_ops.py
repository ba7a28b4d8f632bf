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
