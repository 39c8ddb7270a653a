import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_logits():
    return np.load(os.path.join(GOLDEN, "model_logits.npz"))


@pytest.fixture(scope="session")
def golden_pre():
    return np.load(os.path.join(GOLDEN, "preprocess.npz"))


@pytest.fixture(scope="session")
def example_inputs():
    return np.load(os.path.join(GOLDEN, "example_inputs.npz"))


@pytest.fixture(scope="session")
def golden_batch(golden_logits, example_inputs):
    """The 39 shipped example alerts + the synthetic alerts the goldens were computed on."""
    from btsbot_b200 import synth
    nsyn = int(golden_logits["nsyn"])
    trip = np.concatenate([example_inputs["triplets"], synth.make_triplets(nsyn, start=1000)])
    meta = np.concatenate([example_inputs["metadata"], synth.make_metadata(nsyn, start=1000)])
    img = np.ascontiguousarray(trip.transpose(0, 3, 1, 2))
    return img, meta


@pytest.fixture(scope="session")
def cuda_dev():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from btsbot_b200 import _lib
    _lib.lib()          # fail loudly if the extension is missing on a GPU box
    return torch.device("cuda:0")
