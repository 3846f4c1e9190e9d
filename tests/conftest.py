import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


MODEL_NAMES = ["WDX4_rna004_v1_0", "WDX4b_rna004_v1_0", "WDX4c_rna004_v1_0", "WDX6_rna004_v1_0", "WDX10_rna004_v1_0",   # every shipped DTW_SVM
               "WDX12_rna002_v0_4_4"]   # + the largest deprecated rna002 model: gamma 1.2, 13 classes (KM1 = 16 kernel variant)


@pytest.fixture(scope="session")
def models():
    from warpdemux_b200 import model_io

    return {n: model_io.load_npz(os.path.join(GOLD, "models", n + ".npz")) for n in MODEL_NAMES}


@pytest.fixture(scope="session")
def golden_predict():
    out = {}
    for n in MODEL_NAMES:
        with np.load(os.path.join(GOLD, f"predict_{n}.npz")) as z:
            out[n] = {k: z[k] for k in z.files}
    return out


@pytest.fixture(scope="session")
def golden_fingerprint():
    with np.load(os.path.join(GOLD, "fingerprint_rna004.npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden_real():
    with np.load(os.path.join(GOLD, "real_rna004_WDX4.npz")) as z:
        return {k: z[k] for k in z.files}
