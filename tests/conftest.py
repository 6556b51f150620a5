import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def product_lib():
    """libbtbb.so.1 built in-tree; (re)built on demand.  No fallback: a missing nvcc fails the test."""
    from libbtbb_b200 import build
    build.build()
    from libbtbb_b200 import binding
    return binding.lib()


@pytest.fixture(scope="session")
def orc():
    import util
    return util.oracle()


@pytest.fixture(scope="session")
def gpu_ctx2(product_lib):
    """A library context with tables for up to 2 access-code errors on cuda:0."""
    from libbtbb_b200 import binding
    ctx = binding.Context(0, 2)
    yield ctx
    ctx.close()
