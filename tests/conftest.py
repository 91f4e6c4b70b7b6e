import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')
    config.addinivalue_line('markers', 'timeout: per-test time limit (pytest-timeout; a hung kernel must not hang the run)')


@pytest.fixture(scope='session')
def golden_arrays():
    return dict(np.load(os.path.join(GOLDEN, 'ref_outputs.npz')))


@pytest.fixture(scope='session')
def golden_headers():
    with open(os.path.join(GOLDEN, 'ref_headers.json')) as f:
        return json.load(f)


@pytest.fixture(scope='session')
def bc_hst():
    """Jupiter from HST, 2005-01-01T00:00:00: the reference's main fixture
    (tests/test_body_xy.py:69-76)."""
    from planetmapper_b200.frame import BodyConstants

    with open(os.path.join(GOLDEN, 'jupiter_hst_2005.json')) as f:
        return BodyConstants.from_json_dict(json.load(f))


@pytest.fixture(scope='session')
def oracle():
    from oracle import oracle as O

    O.build()
    return O


# disc parameters of the golden FITS files (tests/test_observation.py:1017, :1084)
GOLDEN_DISC = dict(nx=7, ny=10, x0=2.5, y0=3.1, r0=3.9, rotation_radians=float(np.deg2rad(123.456)))
GOLDEN_ALT = 34567.8912


# Host instantiation of the kernels' per-pixel code (tests/host_check/host_check.cu): TEST INFRASTRUCTURE,
# never linked into or loaded by the product library
HOST_CHECK_SRC = os.path.join(ROOT, 'tests', 'host_check', 'host_check.cu')
HOST_CHECK_SO = os.path.join(ROOT, 'tests', 'host_check', '_build', 'libpm_hostcheck.so')


@pytest.fixture(scope='session')
def HC():
    import ctypes
    import shutil
    import subprocess

    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        pytest.skip('nvcc not available')
    csrc = os.path.join(ROOT, 'planetmapper_b200', 'csrc')
    deps = [HOST_CHECK_SRC] + [os.path.join(csrc, f) for f in ('pm_device.cuh', 'pm_math.cuh')]
    if not os.path.exists(HOST_CHECK_SO) or os.path.getmtime(HOST_CHECK_SO) < max(os.path.getmtime(d) for d in deps):
        os.makedirs(os.path.dirname(HOST_CHECK_SO), exist_ok=True)
        subprocess.run([nvcc, '-O2', '-std=c++17', '-shared', '-Xcompiler', '-fPIC', '-Wno-deprecated-gpu-targets',
                        '-o', HOST_CHECK_SO, HOST_CHECK_SRC], check=True, capture_output=True)
    return ctypes.CDLL(HOST_CHECK_SO)
