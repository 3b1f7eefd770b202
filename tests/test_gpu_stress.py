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


def test_no_read_of_uninitialised_device_memory():
    """GPR_POISON=1 fills every fresh device allocation of the library with 0xFF bytes (NaN doubles): a parity sweep, the
    every-kernel-once workload and the large indefinite-tail check must be unaffected.  (A fresh process gets zeroed
    pages from cudaMalloc, a long-running one does not: two such bugs were found this way in round 1.)"""
    env = dict(os.environ, GPR_POISON="1")
    for args, needle in ((["stress.py", "10", "7000"], "10 cases, 0 violations"), (["sanitize_target.py"], "ok "),
                         (["tail_big_check.py"], "after failed update q=130")):
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", args[0])] + args[1:], capture_output=True, text=True, timeout=900, env=env)
        assert out.returncode == 0 and needle in out.stdout, out.stdout[-3000:] + out.stderr[-2000:]
    chk = [ln for ln in out.stdout.splitlines() if ln.startswith(("fresh", "after failed update"))]
    for ln in chk:                               # tail_big_check prints absolute / relative errors against a numpy solve
        vals = dict(zip(ln.split()[-10::2], ln.split()[-9::2]))
        assert float(vals["f"]) < 1e-8 and float(vals["v"]) < 1e-9 and float(vals["grad"]) < 1e-7, ln
