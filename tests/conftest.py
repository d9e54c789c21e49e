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


class Fixture:
    """The rpn11_ubq fixture as arrays (tests/golden/rpn11_ubq_input.npz), so no test
    reads /root/reference at run time."""

    def __init__(self):
        from pycontact_b200.io.psf import Topology
        z = np.load(os.path.join(GOLDEN, "rpn11_ubq_input.npz"))
        self.coords = z["coords"]
        self.topology = Topology([str(s) for s in z["segids"]], z["resids"].astype(np.int64),
                                 [str(s) for s in z["resnames"]], [str(s) for s in z["names"]],
                                 [str(s) for s in z["types"]], np.zeros(len(z["names"])),
                                 np.ones(len(z["names"])), z["bonds"].astype(np.int64))
        self.backbone = z["backbone"].astype(np.int64)


@pytest.fixture(scope="session")
def rpn11():
    return Fixture()


@pytest.fixture(scope="session")
def golden_ab():
    return dict(np.load(os.path.join(GOLDEN, "rpn11_ubq_golden.npz")))


@pytest.fixture(scope="session")
def golden_self():
    return dict(np.load(os.path.join(GOLDEN, "rpn11_self_golden.npz")))


@pytest.fixture(scope="session")
def golden_synth():
    return dict(np.load(os.path.join(GOLDEN, "synth_small_golden.npz")))
