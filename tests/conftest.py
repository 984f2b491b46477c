import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import __graft_entry__ as graft  # noqa: E402

graft.import_zkl()  # registers the hyphenated package directory as module `zk_lisp_b200`


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def zkl():
    return graft.import_zkl()


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    oracle_lib.build()
    oracle_lib.load()
    return oracle_lib


@pytest.fixture(scope="session")
def ctx(zkl):
    """GPU context; fails loudly if the CUDA library or device is missing (no fallback)."""
    graft.build()
    return zkl.Context(0)
