import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


# every finish() of a test run checks the running arrival records (kept by the kernels that write samples) against a fresh
# scan of the trajectories, the way check_solutions reads them; sub-processes of the sharding tests inherit it
os.environ.setdefault('DABRY_B200_VERIFY_ARRIVALS', '1')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (sm_100) and the built libdabry_b200.so')
    # a fresh checkout has no built artefacts (they are git-ignored): build them once when nvcc is there
    lib = os.path.join(ROOT, 'dabry_b200', 'libdabry_b200.so')
    if not os.path.exists(lib):
        import shutil
        if shutil.which(os.environ.get('NVCC', 'nvcc')):
            sys.path.insert(0, ROOT)
            import __graft_entry__
            __graft_entry__.build()


def _cuda_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture
def hostsim():
    """Bind the g++ build of the device headers (tests/hostsim) for the duration of a test - CPU tier only."""
    import parity_util as pu
    from dabry_b200 import _native
    _native.use_library(pu.build_hostsim())
    yield
    _native.use_library(None)


@pytest.fixture
def cuda_lib():
    """The product library on a real GPU.  Fails (not skips) when the extension is missing on a GPU box."""
    from dabry_b200 import _native
    _native.use_library(None)
    if not _cuda_available():
        pytest.skip('no CUDA device')
    assert os.path.exists(_native.DEFAULT_LIBRARY), 'libdabry_b200.so not built: run __graft_entry__.build()'
    assert _native.library().dabry_backend_name() == b'cuda-sm100a'
    yield
