"""compute-sanitizer over the fused path (the reference has no race / memory checking at all; a CUDA
path needs it).  memcheck and racecheck run the 64x128 smoke invocation -- pixel stage, every cloud stage, both Open3D
filters, the answers, all checked against the oracle inside the script -- in a child process and must report no error."""
import os
import re
import shutil
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sanitizer():
    for cand in (shutil.which("compute-sanitizer"), "/usr/local/cuda/bin/compute-sanitizer"):
        if cand and os.path.exists(cand):
            return cand
    pytest.skip("compute-sanitizer is not installed")


@pytest.mark.parametrize("tool,summary", [("memcheck", r"ERROR SUMMARY: 0 errors"),
                                          ("racecheck", r"RACECHECK SUMMARY: 0 hazards displayed \(0 errors, 0 warnings\)")])
def test_fused_path_under_compute_sanitizer(cuda_device, tool, summary):
    exe = _sanitizer()
    cmd = [exe, "--tool", tool, "--error-exitcode", "9", sys.executable, os.path.join(ROOT, "tools", "gpu_smoke.py"), "64", "128"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-3000:]
    assert re.search(summary, out), out[-3000:]
    # the script itself compared every stage with the oracle
    assert "final src equal: True" in out and "final src equal: False" not in out, out[-3000:]
