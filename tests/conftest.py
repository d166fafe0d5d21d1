import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session')
def goldens():
    with open(os.path.join(GOLDEN, 'reference_goldens.json')) as f:
        return json.load(f)


@pytest.fixture(scope='session')
def utd():
    """reference tests/unit_tests_data arrays (n=100)."""
    with np.load(os.path.join(GOLDEN, 'unit_tests_data.npz')) as d:
        return {k: d[k] for k in d.files}


@pytest.fixture(scope='session')
def lmmfix():
    with np.load(os.path.join(GOLDEN, 'lmm_fixture.npz')) as d:
        return {k: d[k] for k in d.files}


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name)) as d:
        return {k: d[k] for k in d.files}
