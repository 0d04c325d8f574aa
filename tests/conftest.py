import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def ob():
    import oracle_binding
    oracle_binding.build()
    return oracle_binding


@pytest.fixture(scope="session")
def demo_index(ob):
    try:
        return ob.ensure_demo_index()
    except FileNotFoundError as e:
        pytest.skip(str(e))


@pytest.fixture(scope="session")
def oracle(ob, demo_index):
    o = ob.Oracle(demo_index)
    yield o
    o.close()


@pytest.fixture(scope="session")
def gpu(demo_index):
    """(module, Index, Context) on cuda:0 -- fails loudly (no skip, no fallback) when the library or the GPU is missing"""
    import desamba_b200 as dsb
    ix = dsb.Index(demo_index, 0)
    ctx = dsb.Context(ix)
    yield dsb, ix, ctx
    ctx.close()
    ix.close()
