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


# Bulk hashing modes every GPU parity test runs under (all three must be bit-identical to the oracle):
#   default   tensor-core sponge kernels from 1024 sponges up, 64-thread-group kernels (kernels_quad.cuh) below and for the coin
#   tc-forced tensor-core kernels for every bulk call site (threshold 1): small test sizes reach tcperm.cuh
#   int-pipe  the integer-pipe thread-per-sponge kernels for every bulk call site, 16-thread groups for the coin
#   coop16    like default with the 16-thread cooperative kernels instead of the 64-thread groups
HASH_MODES = {"default": (None, 0, 2), "tc-forced": (1, 0, 2), "int-pipe": (1, 1, 1), "coop16": (None, 0, 1)}


@pytest.fixture(scope="session", params=list(HASH_MODES))
def ctx(zkl, request):
    """GPU context; fails loudly if the CUDA library or device is missing (no fallback)."""
    graft.build()
    c = zkl.Context(0)
    thr, path, coop = HASH_MODES[request.param]
    if thr is not None:
        c._check(c.lib.zkl_set_bulk_threshold(c.h, thr))
    c._check(c.lib.zkl_set_hash_path(c.h, path))
    c._check(c.lib.zkl_set_coop_path(c.h, coop))
    c.hash_mode = request.param
    return c
