import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def lib():
    from zoomearth_b200 import build, _lib
    build.build()
    return _lib.lib()


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from zoomearth_b200 import build
    build.build()
    return torch.device("cuda", 0)


@pytest.fixture(scope="session")
def full_sd():
    """Seeded random weights of the full 32-block tower (HF names), shared by the full-depth parity tests."""
    from oracle import tower as OT
    return OT.make_weights(0)


@pytest.fixture(scope="session")
def full_sd_cuda(cuda, full_sd):
    """The same weights resident on the GPU for the fp32 oracle run there (TF32 off, see oracle.tower.forward)."""
    return {k: v.to(cuda) for k, v in full_sd.items()}


@pytest.fixture(scope="session")
def full_visual(cuda, full_sd):
    """The shipped configuration: all 32 blocks, fp16 operands (the dtype bench.py times), fp32 embeddings out."""
    import torch
    from zoomearth_b200 import FusedVisual
    return FusedVisual(full_sd, device=cuda, dtype=torch.float32, operand_dtype=torch.float16)
