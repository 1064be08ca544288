import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # The C-ABI library is a build artefact (git-ignored).  In a fresh checkout build it before collection:
    # nvcc cross-compiles sm_100a without a GPU.  (On the GPU box the prebuilt .so travels with the snapshot.)
    lib = os.path.join(ROOT, "rslo_b200", "_C", "librslo_b200.so")
    if not os.path.exists(lib):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "rslo_b200", "csrc"), "-j8"], stdout=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
