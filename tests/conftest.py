import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "gpu-provisioner_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """libb200sched.so (nvcc cross-compiles without a GPU) and the oracle are built once per session."""
    import __graft_entry__ as g
    g.build()
    yield


@pytest.fixture(scope="session")
def gpu_ctx():
    from b200sched import capi
    ctx = capi.Context(0)
    yield ctx
    ctx.close()
