import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def fixture_nx4():
    from lyft3d_b200 import synth
    return synth.fixture_points_nx4()


@pytest.fixture(scope="session")
def fixture_4xn():
    from lyft3d_b200 import synth
    return synth.fixture_points_4xn()


@pytest.fixture(scope="session")
def cloud11():
    from lyft3d_b200 import synth
    return synth.multisweep_cloud(11)


@pytest.fixture(scope="session")
def cloud20():
    from lyft3d_b200 import synth
    return synth.multisweep_cloud(20)
