import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope='session')
def oracle_pm():
    """The numpy oracle's ParticleMesh class (test infrastructure only)."""
    import oracle
    oracle.activate(shims=False)
    from pmesh.pm import ParticleMesh
    return ParticleMesh


@pytest.fixture(scope='session')
def device_pm():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from vmad_b200.pm import ParticleMesh
    return ParticleMesh


def rel_l2(a, b):
    import numpy
    a = numpy.asarray(a)
    b = numpy.asarray(b)
    den = numpy.sqrt((numpy.abs(b) ** 2).sum())
    num = numpy.sqrt((numpy.abs(a - b) ** 2).sum())
    return num / den if den > 0 else num
