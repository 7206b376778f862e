import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    # the CPU oracle (eager PyTorch) collapses when over-subscribed on many-core hosts
    import torch

    torch.set_num_threads(min(16, os.cpu_count() or 1))
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a)")


@pytest.fixture(scope="session")
def cuda_dev():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import mtt_b200
    from mtt_b200 import lib

    l = lib.load()
    lib.check(l.mtt_device_check(), "mtt_device_check")
    return torch.device("cuda:0")
