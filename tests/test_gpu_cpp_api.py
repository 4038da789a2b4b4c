"""Runs the reference's own test cases, re-expressed against the C++ drop-in API (tests/cpp/reference_cases.cu), on
the GPU.  A fourth compilation of the same cases, next to the reference's impl / stdpar / lapack builds."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.gpu
def test_reference_cases_against_cpp_api():
    exe = os.path.join(HERE, "cpp", "reference_cases")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.join(HERE, "cpp"), "--no-print-directory"])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    print(r.stdout)
    print(r.stderr)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "0 failed" in r.stdout
