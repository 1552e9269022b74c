import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# The oracle runs OpenMP on every host core; under pytest-xdist each worker would do so at once and the tiny
# lattices of the fuzz tests then spend their time in oversubscribed barriers (a 4-worker soak once took 24 minutes
# instead of 30 s).  Give each xdist worker a small, fixed share instead.
if os.environ.get("PYTEST_XDIST_WORKER") and "OMP_NUM_THREADS" not in os.environ:
    os.environ["OMP_NUM_THREADS"] = "2"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the oracle (gcc) once per session; the CUDA library is built by __graft_entry__.build()."""
    from oracle import lbm_oracle
    lbm_oracle.build()
