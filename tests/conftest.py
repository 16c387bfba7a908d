import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "links-3d-human-pose-estimation_b200")
for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
os.environ.setdefault("WANDB_MODE", "disabled")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with `-m gpu` under gpurun)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    d = os.path.join(ROOT, "tests", "golden")
    return {n[:-4]: np.load(os.path.join(d, n)) for n in os.listdir(d) if n.endswith(".npz")}


from hostsim_util import backend  # noqa: E402,F401  (fixture shared by the kernel tests)
