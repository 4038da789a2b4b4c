"""CPU tests of the header-only host engine (lambda_lanczos_b200/lambda_lanczos.hpp, exponentiator.hpp) against a TEST
DOUBLE of the C ABI (tests/cpp/mock_llz.cpp: host memory, one worker thread playing the CUDA stream): what is under
test is host logic only — results identical for every pipelining depth and for one or two host threads, the hand-over of
the DGKS refinement from the helper thread to the launch thread, the lock-step rule of row-sharded runs (every rank
enqueues exactly itern + depth iterations whatever its threads' timing), LLZ_ERR_OOM out of either thread without a
hang, the reference-verbatim constructors with a host mv_mul, and the coefficient scaling of the lazily normalised
Exponentiator against the exact exponential.  The second test runs the same cases under ThreadSanitizer."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def build():
    subprocess.check_call(["make", "-C", os.path.join(HERE, "cpp"), "--no-print-directory", "host_loop"])


def test_host_loop_cases():
    build()
    r = subprocess.run([os.path.join(HERE, "_build", "host_loop_cases")], capture_output=True, text=True, timeout=600)
    print(r.stdout)
    assert r.returncode == 0 and "0 failed" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


def test_host_loop_cases_under_thread_sanitizer():
    build()
    exe = os.path.join(HERE, "_build", "host_loop_cases_tsan")
    if not os.path.exists(exe):
        pytest.skip("this g++ cannot build with -fsanitize=thread")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900, env=dict(os.environ, TSAN_OPTIONS="halt_on_error=0"))
    if "FATAL: ThreadSanitizer" in r.stderr and "checks" not in r.stdout:
        pytest.skip("ThreadSanitizer cannot run in this sandbox: " + r.stderr.strip().splitlines()[0][:200])
    assert "WARNING: ThreadSanitizer" not in r.stderr, r.stderr[-4000:]
    assert r.returncode == 0 and "0 failed" in r.stdout, r.stdout[-3000:]
