"""The fp32-accurate convolution mode (RFK_CONV_PRECISION=bf16x3, alias tf32: split-precision operands on the same tcgen05
kernels, include/rfk.h rfk_set_conv_split) under the 1e-3 gate of BASELINE.json: every module / ListGlow / ConvLSTM parity
test of test_gpu_modules.py and the full-depth configurations of test_gpu_fullsize.py are re-run in a child process (the
mode is a process-global switch read at import) with RFK_TEST_TOL=1e-3."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(files, select=None):
    env = dict(os.environ, RFK_CONV_PRECISION="bf16x3", RFK_TEST_TOL="1e-3")
    cmd = [sys.executable, "-m", "pytest", "-q", "-m", "gpu", "-x", "-p", "no:cacheprovider"] + [os.path.join(ROOT, "tests", f) for f in files]
    if select:
        cmd += ["-k", select]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    tail = r.stdout[-3000:] + r.stderr[-1500:]
    assert r.returncode == 0, tail
    return r.stdout


def test_modules_at_1e3_in_split_precision():
    out = _run(["test_gpu_modules.py"])
    assert " passed" in out and "failed" not in out
    print(out.strip().splitlines()[-1])


def test_full_depth_at_1e3_in_split_precision():
    out = _run(["test_gpu_fullsize.py"], "full_depth or cfg2_full_size")
    assert " passed" in out and "failed" not in out
    print(out.strip().splitlines()[-1])
