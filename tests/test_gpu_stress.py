"""Randomised parity sweep (GPU): random sizes / kernels / noise settings / batch sizes / append sequences against the
CPU oracle with conditioning-aware tolerances (tools/stress.py; 40 cases were clean in round 1, 12 run here)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_randomised_parity_sweep():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "stress.py"), "12", "1000"], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert "12 cases, 0 violations" in out.stdout
