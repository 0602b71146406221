import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_report_header(config):
    """Says in every test log which oracle the parity tests compare with."""
    ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libb2ref.so"))
    return ("parity oracle: the reference itself (oracle/_ref/libb2ref.so, vanilla Box2D 2.3.2 compiled from /root/reference)"
            if ref else "parity oracle: oracle/_ref is MISSING — gpu parity tests will fail (or use the C restatement if "
                        "B2GPU_ALLOW_PORT_ORACLE=1)")


@pytest.fixture(scope="session")
def oracles():
    """Build (if needed) and return the oracle wrappers.  TEST INFRASTRUCTURE only."""
    import subprocess

    from oracle import worlds

    if not worlds.available("port"):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    return worlds
