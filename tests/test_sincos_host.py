"""The branch-free Cody-Waite sincos of the basis phase (fastfp_b200/csrc/ffp_sincos.cuh), compiled
for the host and checked against long double sinl/cosl. CPU only."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_sincos_cw_accuracy(tmp_path):
    exe = tmp_path / "sincos_check"
    subprocess.run(
        ["g++", "-O2", "-mfma", "-o", str(exe), os.path.join(ROOT, "tests", "sincos_host_check.cpp")], check=True
    )
    out = subprocess.run([str(exe), "1500000"], check=True, capture_output=True, text=True).stdout
    m = re.search(r"max_abs_err (\S+) max_ulp_err (\S+)", out)
    assert m, out
    assert float(m.group(1)) < 2.3e-16  # absolute, over |x| <= 1e5 rad
    assert float(m.group(2)) < 2.0  # ulp, where |value| >= 1e-3
    m2 = re.search(r"lockstep_mismatch (\d+)", out)
    assert m2 and int(m2.group(1)) == 0, out  # sincos_cw_n<4> == sincos_cw bit for bit
