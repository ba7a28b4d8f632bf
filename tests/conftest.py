import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    """Golden case written by oracle/make_golden.py from the unmodified reference."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    case = {"zf": z["zf"], "zh": z["zh"], "dt": float(z["dt"]), "f_les": float(z["f_les"]),
            "f_gcm": float(z["f_gcm"]), "conservative": bool(z["conservative"]), "A_les": z["A_les"],
            "gcm": {}, "aux": {}, "les": {}, "out": {}}
    for k in z.files:
        for pre in ("gcm", "aux", "les", "out"):
            if k.startswith(pre + "_"):
                case[pre][k[len(pre) + 1:]] = z[k]
    return case


GOLDEN_CASES = ["ref_L19", "ref_L91", "ref_L137", "ref_L19_nk20"]


def relerr(a, b):
    """max|a-b| / max|b| per profile variable (SURVEY.md §8d parity gate)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    d = float(np.max(np.abs(a - b))) if a.size else 0.0
    s = float(np.max(np.abs(b))) if b.size else 0.0
    return d / s if s > 0 else d


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
