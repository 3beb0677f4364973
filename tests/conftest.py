import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ckpt_sd():
    return torch.load(os.path.join(GOLDEN, "weights_20240121-172745.pt"), map_location="cpu")


@pytest.fixture(scope="session")
def rand_sd():
    from mind_b200 import synth
    shapes = json.load(open(os.path.join(GOLDEN, "shapes.json")))
    return synth.random_state_dict(0, like=shapes)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def rel_err(a, b):
    """max |a-b| / max |b| : the relative measure used for the 1e-3 bound of north_star"""
    a = torch.as_tensor(a, dtype=torch.float32).cpu()
    b = torch.as_tensor(b, dtype=torch.float32).cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()
