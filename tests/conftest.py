import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lib():
    """The built C-ABI library (built here if the .so is missing; nvcc cross-compiles without a GPU)."""
    from speech_transcript_embeddings_b200 import _lib, build
    if not _lib.LIB_PATH.exists():
        build.build()
    return _lib.load()


@pytest.fixture(scope="session")
def cuda_device(lib):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("test marked gpu but no CUDA device is visible (there is no CPU fallback)")
    return torch.device("cuda", 0)


def load_golden(name):
    return np.load(GOLDEN / name, allow_pickle=False)
