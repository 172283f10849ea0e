import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


def load_golden(name):
    """Committed output of the unmodified reference (tests/golden/make_golden.py) as a dict of torch tensors."""
    with np.load(GOLDEN / f"{name}.npz") as z:
        return {k: torch.from_numpy(z[k]) for k in z.files}


@pytest.fixture(scope="session")
def golden():
    return load_golden


def rel_err(a, b):
    """Norm-wise relative error ||a-b|| / ||b|| (the 1e-3 tolerance of BASELINE.json is applied to this)."""
    a = a.double()
    b = b.double()
    d = (a - b).norm().item()
    n = b.norm().item()
    return d / n if n > 0 else d
