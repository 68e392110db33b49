import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `-m gpu` under gpurun)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_py as op
    op.build(ref=True)
    return op


@pytest.fixture(scope="session")
def have_ref(oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/libref_alglib_qp.so not built (needs /root/reference)")
    return True


@pytest.fixture(scope="session")
def emu():
    from tests import util
    return util.Emu()


@pytest.fixture(scope="session")
def gpu_batch():
    """A WbcBatch on cuda:0.  Fails (not skips) when the CUDA library cannot be used: no CPU fallback."""
    from wbc_quadruped_dob_b200 import api
    b = api.WbcBatch(max_batch=70000, device=0)
    yield b
    b.close()
