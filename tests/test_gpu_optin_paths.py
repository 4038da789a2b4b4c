"""Every environment-selected code path of the library against the default path (test or delete).

Each switch is read when the library, the context or an operator is created, so each runs in a process of its own
(tests/optin_worker.py) on the same fixed set of applies and runs:

  LLZ_SELL_TMA=1     SELL SpMV with the matrix streamed by cp.async.bulk + mbarrier  -> y bit-identical (alpha to rounding)
  LLZ_SPMV=v         CSR lanes-per-row kernel instead of the stream kernel           -> y to rounding (shuffle tree)
  LLZ_BASIS_VMM=0    one cudaMalloc for the Krylov basis instead of the VMM store    -> everything bit-identical
  LLZ_FUSED_ORTH=0   project / reduce / update / scale as separate launches          -> runs to rounding, same counts
  LLZ_XXZ_KERNEL=state  one-thread-per-state XXZ kernel instead of the block kernel  -> y bit-identical
  LLZ_DIA=0          SELL instead of DIA storage for stencil operators               -> y bit-identical
  (LLZ_DIA=0 + LLZ_SELL_TMA=1: the TMA kernel on the stencil operators the default run stores as DIA)
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SWITCHES = ("LLZ_SELL_TMA", "LLZ_SPMV", "LLZ_BASIS_VMM", "LLZ_FUSED_ORTH", "LLZ_XXZ_KERNEL", "LLZ_XXZ_M", "LLZ_DIA")


def run_worker(extra):
    env = {k: v for k, v in os.environ.items() if k not in SWITCHES}
    env.update(extra)
    r = subprocess.run([sys.executable, os.path.join(HERE, "optin_worker.py")], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("OPTIN_RESULT ")][-1]
    return json.loads(line[len("OPTIN_RESULT "):])


@pytest.fixture(scope="module")
def default_result():
    return run_worker({})


def close_runs(a, b, same_vectors):
    if same_vectors:
        assert a["its"] == b["its"] and a["vec"] == b["vec"], (a["its"], b["its"])
    # another reduction tree of a scalar moves a run by rounding: the stopping test (a relative change of 2e-13 of the
    # Ritz values) may then fire an iteration earlier or later
    assert np.all(np.abs(np.atleast_1d(a["its"]) - np.atleast_1d(b["its"])) <= 2), (a["its"], b["its"])
    if "ev" in a:
        assert np.allclose(a["ev"], b["ev"], rtol=1e-10, atol=1e-12), (a["ev"], b["ev"])


@pytest.mark.gpu
@pytest.mark.parametrize("switch,value,bitwise", [("LLZ_SELL_TMA", "1", False), ("LLZ_SPMV", "v", False), ("LLZ_BASIS_VMM", "0", True),
                                                  ("LLZ_FUSED_ORTH", "0", False), ("LLZ_XXZ_KERNEL", "state", False), ("LLZ_DIA", "0", False),
                                                  ("LLZ_DIA+TMA", "0", False)])
def test_opt_in_path_matches_default(default_result, switch, value, bitwise):
    """`bitwise`: whole runs reproduce bit for bit (the switch changes no floating-point operation order at all).  The
    applies (y = A x) are compared bit for bit for every switch but LLZ_SPMV=v; a different alpha-dot reduction tree
    (XXZ kernels) or register tile (fused orthogonalisation at small n) moves a run by rounding only."""
    got = run_worker({"LLZ_DIA": "0", "LLZ_SELL_TMA": "1"} if switch == "LLZ_DIA+TMA" else {switch: value})
    assert got.keys() == default_result.keys()
    for key, ref in default_result.items():
        val = got[key]
        if key.startswith("csr_vals/"):
            assert np.allclose(val, ref, rtol=1e-13, atol=1e-14), key
        elif key.startswith("lanczos/") or key.startswith("expm/"):
            # (the small-n fused kernel keeps 16 instead of 8 columns per register tile: another summation tree)
            close_runs(val, ref, same_vectors=bitwise)
        elif key.startswith("csr/") and switch == "LLZ_SPMV":
            continue  # covered by csr_vals to rounding
        else:
            assert val == ref, (switch, key)
