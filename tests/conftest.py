import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session", autouse=True)
def _built_native():
    """Make sure the CPU oracle library exists (the CUDA library is built by __graft_entry__.build)."""
    import fdtd_oracle
    fdtd_oracle.build_lib()
    so = os.path.join(ROOT, "py-fdtd_pic_b200", "libpyfdtd_b200.so")
    if not os.path.exists(so):
        import __graft_entry__
        __graft_entry__.build()
    yield


def load_golden(name):
    import json
    import numpy as np
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    for k in ("spec", "versions", "members"):
        if k in d:
            d[k] = json.loads(str(d[k]))
    return d
