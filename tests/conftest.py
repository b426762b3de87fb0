import json
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def goldens():
    with open(os.path.join(GOLDEN, "reference_goldens.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def aniline_system():
    from paramol_b200.Utils.amber import system_from_prmtop
    return system_from_prmtop(os.path.join(GOLDEN, "aniline.prmtop"))


@pytest.fixture(scope="session")
def aniline_data():
    from paramol_b200.Utils.netcdf_lite import read_paramol_data
    return read_paramol_data(os.path.join(GOLDEN, "aniline_10_struct.nc"))


def gold(goldens, key):
    return np.asarray(goldens[key])
