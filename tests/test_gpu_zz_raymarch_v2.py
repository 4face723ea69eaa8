"""The opt-in second-generation lit ray march (raymarch_fast2_kernel: shared sampler taps for interior samples, exact leaping of
empty bricks; tbrm_options.reserved[1] = 3 or TBRM_RAYMARCH_V2=1) and the default kernel's 32-bit tap addressing (ADDR32) must give the frames and executed-step
counts of the first-generation fast kernel with 64-bit addressing and of the generic kernel, bit for bit: cfg2 (512^3 / 1080p / 512 steps), a clipped and a scaled + rotated world at
256^3, and a half-resolution light volume. The checks live in scripts/validate_raymarch_v2.py (also the round-1 validation run)."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def test_second_generation_raymarch_is_bit_identical():
    p = subprocess.run([sys.executable, str(ROOT / "scripts" / "validate_raymarch_v2.py")], cwd=str(ROOT), stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=300)
    assert p.returncode == 0 and "V2 OK" in p.stdout, p.stdout[-3000:]
