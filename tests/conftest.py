import os
import sys

import pytest

# Several tests run >= 2 "ranks" (contexts + streams + host threads) inside ONE process, with kernels of one rank spinning
# on flags another rank's kernel sets.  With CUDA's default lazy module loading the FIRST launch of a kernel may need a
# context synchronisation, which can never complete while a peer's kernel is spinning on it — load everything eagerly.
# (One process per GPU, the production layout, is not affected: a process only waits on OTHER processes' kernels.)
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure): oracle/pyoracle.py over oracle/_build/liboracle.so."""
    from oracle import pyoracle
    pyoracle.lib()
    return pyoracle
