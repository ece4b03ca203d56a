"""GPU: recipe K fused with the encoder's input stage (stx_fbank_k_projection), checked by
tests/scripts/fused_projection_check.py under a timeout (a blocked kernel fails the test instead of hanging the run)."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_fused_front_end_and_projection(cuda_device):
    try:
        r = subprocess.run([sys.executable, str(ROOT / "tests" / "scripts" / "fused_projection_check.py"), "40"],
                           capture_output=True, text=True, timeout=180, cwd=ROOT)
    except subprocess.TimeoutExpired:
        pytest.fail("tests/scripts/fused_projection_check.py did not finish in 180 s")
    print(r.stdout[-1500:])
    assert r.returncode == 0 and "fused projection ok" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
