"""GPU: parity on signal classes the seeded generators do not cover (tests/scripts/adversarial_parity.py: DC offsets,
impulses, square waves, rumble, steps, huge / tiny amplitudes, integer-valued PCM, a lone high-frequency tone)."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_adversarial_signal_classes(cuda_device):
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "scripts" / "adversarial_parity.py")], capture_output=True,
                       text=True, timeout=600, cwd=ROOT)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
