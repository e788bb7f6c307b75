import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


REFERENCE_ROOT = "/root/reference"


@pytest.fixture(scope="session")
def have_reference():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "Functions"))


# tolerances stated by BASELINE.json's north_star: relative L2 error
TOL = {"single": 1e-5, "double": 1e-12}
