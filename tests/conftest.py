import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_xdist_auto_num_workers(config):
    """Worker processes for `-n auto` (pytest.ini): none on a GPU box, a handful for the CPU-only suite."""
    env = os.environ.get("LDD_TEST_WORKERS")
    if env is not None:
        return max(int(env), 0)
    if _have_gpu():
        return 0
    return max(1, min(6, (os.cpu_count() or 2) - 1))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    if not hasattr(config, "workerinput") and not _have_gpu():
        # the controller builds the emulated library once, before the workers (which would race for its object files) start
        try:
            from emu_util import emu_backend
            emu_backend()
        except Exception:
            pass


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))
        return cache[name]

    return load


def _cuda_backend():
    from lddecode_b200._backend import CudaBackend
    return CudaBackend()


# Kernel tests are written once against the C ABI and run on two targets:
#   "cuda": the real library on the GPU (-m gpu; the parity tests proper)
#   "emu" : the same .cu sources compiled for the CPU by tests/emu (-m "not gpu"; small cases)
@pytest.fixture(scope="session", params=[pytest.param("emu"), pytest.param("cuda", marks=pytest.mark.gpu)])
def backend(request):
    if request.param == "cuda":
        return _cuda_backend()
    from emu_util import emu_backend
    return emu_backend()


@pytest.fixture(scope="session")
def cuda_backend():
    return _cuda_backend()
