import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def built_lib():
    """Builds (if needed) and loads the CUDA shared library.  No skip: a missing library is a failure."""
    from mlx_swift_audio_b200 import build, _lib
    build.build()
    return _lib.load()


@pytest.fixture(scope="session")
def ctx(built_lib):
    from mlx_swift_audio_b200 import api
    c = api.Context(0)
    yield c
    c.close()
