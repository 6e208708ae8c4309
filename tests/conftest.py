import importlib.util
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLD, "reference_outputs.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def small_clip():
    return np.load(os.path.join(GOLD, "small_clip.npz"))["clip"]


@pytest.fixture(scope="session")
def vqa():
    import rtvqa_b200
    return rtvqa_b200


@pytest.fixture(scope="session")
def synth():
    import rtvqa_b200
    return rtvqa_b200.synth

