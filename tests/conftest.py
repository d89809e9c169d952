import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFERENCE = "/root/reference"  # only present in the authoring container, never on the GPU box


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def goldens():
    with open(os.path.join(ROOT, "tests", "golden", "reference_goldens.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def config5_golden():
    """Oracle results on the 128-reactor design over BASELINE config 5 (tools/make_config5_golden.py)."""
    import numpy as np
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "config5_oracle.npz")))


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def omech(oracle):
    cache = {}

    def get(name, constants=None):
        key = (name, constants)
        if key not in cache:
            cache[key] = oracle.Mechanism(name) if constants is None else oracle.Mechanism(name, constants=constants)
        return cache[key]
    return get


def _ensure_product_lib():
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(ROOT, "chemtools_b200", "libchemtools_b200.so")):
        g.build()


@pytest.fixture(scope="session")
def ctb():
    _ensure_product_lib()
    import chemtools_b200
    return chemtools_b200


@pytest.fixture(scope="session")
def chem(ctb):
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = ctb.ThermoKinetics(ctb.mechanism_path(name))
        return cache[name]
    return get


needs_reference = pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="/root/reference not present")
