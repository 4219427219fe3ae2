import os
import sys
import warnings

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
warnings.filterwarnings("ignore", category=FutureWarning)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: torch.from_numpy(z[k]) if z[k].dtype != np.bool_ else torch.from_numpy(z[k].copy())
                for k in z.files}


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = load_golden(name)
        return cache[name]
    return get


def rel_err(a, b, floor=0.1):
    """Parity metric behind the "1e-4 rel fp32" gate of BASELINE.json:

        max_i |a_i - b_i| / max(|b_i|, floor * max_j |b_j|)

    With the default floor this is ``allclose(rtol=1e-4, atol=1e-5 * max|b|)`` when gated at
    1e-4: element-wise relative error for every element within 10x of the tensor's largest
    magnitude, and an absolute error (relative to the largest magnitude) for the small tail,
    where a purely relative figure measures fp32 cancellation (1 - exp(-x), s near 0) rather
    than the kernel."""
    a, b = a.double(), b.double()
    scale = b.abs().clamp(min=max(floor * float(b.abs().max()), 1e-30))
    return float(((a - b).abs() / scale).max())


def max_abs(a, b):
    return float((a.double() - b.double()).abs().max())
