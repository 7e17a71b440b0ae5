import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long-running")


def pytest_sessionfinish(session, exitstatus):
    """write the achieved-parity ledger of this session (tests/parity_report.py)"""
    import parity_report

    path = parity_report.write()
    if path:
        print(f"\nparity ledger: {path} ({len(parity_report.ENTRIES)} entries)")


@pytest.fixture(scope="session")
def oracle():
    """ctypes Backend over oracle/liboracle.so (the CPU restatement; checker only)."""
    import helpers

    return helpers.oracle_backend()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference extension built into oracle/_ref (skips when absent)."""
    import helpers

    mod = helpers.reference_module()
    if mod is None:
        pytest.skip("oracle/_ref/_cmtj*.so not built (needs /root/reference)")
    return mod


@pytest.fixture
def on_oracle(oracle):
    """Route the facade through the oracle for host-logic tests on a CPU-only machine.
    Only tests may do this; the product has no such fallback."""
    import helpers

    with helpers.facade_on(oracle):
        yield oracle


@pytest.fixture
def on_cuda():
    import helpers

    with helpers.facade_on(None) as be:
        yield be
