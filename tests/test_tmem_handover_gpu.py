"""GPU: kernels that allocate tensor memory launched back to back without host synchronisation (tests/scripts/tmem_handover.py),
under a timeout so that a blocked tcgen05.alloc fails the test instead of hanging the run."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_tmem_is_handed_over_between_back_to_back_kernels(cuda_device):
    try:
        r = subprocess.run([sys.executable, str(ROOT / "tests" / "scripts" / "tmem_handover.py")], capture_output=True,
                           text=True, timeout=120, cwd=ROOT)
    except subprocess.TimeoutExpired:
        pytest.fail("tests/scripts/tmem_handover.py did not finish in 120 s: a kernel is blocked (tcgen05.alloc?)")
    assert r.returncode == 0 and "tmem hand-over ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
