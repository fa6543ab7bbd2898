import os
import sys

import pytest

# several ranks of the multi-GPU session share device 0 in the GPU suite and wait for each other inside kernels:
# every stream needs its own hardware queue (must be set before CUDA is initialised)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def ctx():
    """One CUDA context for the whole GPU session; fails loudly if the extension is missing."""
    from rala_b200 import api
    c = api.Context(0)
    yield c
    c.close()
