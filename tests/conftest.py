import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def load_case(name):
    import json
    from openmoc_b200.trackfile import read_trackfile
    ft = read_trackfile(os.path.join(GOLDEN, name + ".b2trk"))
    ref = json.load(open(os.path.join(GOLDEN, name + ".json")))
    return ft, ref
