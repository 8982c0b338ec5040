import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_files():
    return sorted(f for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and f[0] == "n" and f[1].isdigit())


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN_DIR
