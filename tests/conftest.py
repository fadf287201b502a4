import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def golden():
    from tests import util
    return util.Golden()


@pytest.fixture(scope='session')
def native():
    """The built C-ABI library; GPU tests fail (not skip) when it is missing or has no device."""
    from autoencoder_based_image_compression_b200 import _native
    _native.lib()
    return _native
