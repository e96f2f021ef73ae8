import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# under pytest-xdist every worker would start its own OpenMP team of the oracle (libgomp spin-waits): 8 workers x
# 8 threads on 8 cores turn a 30 s suite into 25 minutes (measured).  One thread per worker, passive waiting.
if os.environ.get("PYTEST_XDIST_WORKER"):
    os.environ.setdefault("OMP_NUM_THREADS", "1")
os.environ.setdefault("OMP_WAIT_POLICY", "passive")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the oracle and the host-emulation library once (seconds); libpomgpu.so is built
    by __graft_entry__.build() and must already be in-tree for the -m gpu tests."""
    from oracle import pomo
    pomo.build()
    from tests import emu
    emu.build_emu()
