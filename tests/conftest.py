import os
import sys

import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def root():
    return os.path.abspath(ROOT)


@pytest.fixture(scope="session", autouse=True)
def built():
    """The CUDA library and the oracle are built in-tree (graft build()); build on demand
    when a test session starts from a clean checkout."""
    lib = os.path.join(ROOT, "sdrreceiver_b200", "libsdrb200.so")
    orc = os.path.join(ROOT, "oracle", "libsdr_oracle.so")
    if not (os.path.exists(lib) and os.path.exists(orc)):
        import __graft_entry__ as g
        g.build()
    return True


PLANS = ["25E", "98W", "54W_all", "54W_288K", "CBAND_143E"]


def plan_path(name):
    return os.path.join(os.path.abspath(ROOT), "plans", name + ".ini")


def level_for(op):
    return 0.5 if any(s["late"] for s in op["subs"]) else 1.0
