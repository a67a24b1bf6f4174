import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # GPU tests fail loudly (not skip) on a GPU box; on a CPU box they are deselected by -m "not gpu".
    pass


@pytest.fixture(scope="session")
def built():
    """make sure the in-tree libraries exist (CPU build of every artefact)"""
    import __graft_entry__ as g

    g.build()
    return True
