import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
SHIPPED_F32 = os.path.join(GOLDEN, "jda_shipped_f32.model")
REF_SHIPPED_F64 = "/root/reference/model/jda_tmp_20160913-112648_stage_5_cart_540.model"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle.Oracle()


@pytest.fixture(scope="session")
def reflib():
    """The reference's own c/jda.c (oracle/_ref/libjda_ref.so); skip when it was never built."""
    from oracle import pyoracle
    if not os.path.exists(pyoracle.REF_SO):
        pytest.skip("oracle/_ref/libjda_ref.so not built (needs /root/reference)")
    return pyoracle.RefLib()


@pytest.fixture(scope="session")
def oracle_shipped(oracle):
    h = oracle.load(SHIPPED_F32, double=False)
    assert h
    yield h
    oracle.release(h)


@pytest.fixture(scope="session")
def gold():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "ref_outputs.npz"))
