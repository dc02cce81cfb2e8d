import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# small pipeline chunks so that multi-chunk staging is exercised by small test inputs
os.environ.setdefault("GT_CHUNK_BASES", "200000")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def gb():
    """The product package bound to cuda:0 (fails loudly if the CUDA library is missing)."""
    import goetia_b200
    goetia_b200.init(0)
    return goetia_b200


@pytest.fixture(scope="session")
def golden():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        return json.load(f)
