"""Generates tests/golden/reference_runs.npz from the UNMODIFIED reference compiled here (oracle/_ref/libllz_ref.so).

Run in the build container (where /root/reference exists):   python tests/golden/make_golden.py
The fixture travels with the repository, so the GPU box (which has no /root/reference) and any later checkout can pin
both the CPU restatement and the CUDA path against outputs of the real reference.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def cases(wl):
    """name -> (csr, lanczos kwargs, dtype).  Start vector is always wl.start_vector(n, dtype) (seed 1)."""
    return {
        "random_sym_n3000_max1": (wl.random_symmetric_csr(3000), dict(find_max=True, num_eigs=1), np.float64),
        "random_sym_n2000_f32_max1": (wl.random_symmetric_csr(2000, dtype=np.float32), dict(find_max=True, num_eigs=1), np.float32),
        "random_sym_n3000_min3_offset": (wl.random_symmetric_csr(3000, seed=7), dict(find_max=False, num_eigs=3, offset=-3.0), np.float64),
        "laplacian_24_min4": (wl.laplacian2d_csr(24), dict(find_max=False, num_eigs=4), np.float64),
        "peierls_16x16_min2": (wl.peierls_csr(16, 16), dict(find_max=False, num_eigs=2), np.complex128),
        "xxz_L12_ground": (wl.xxz_csr(12), dict(find_max=False, num_eigs=1), np.float64),
        "xxz_L12_c128_ground": (wl.xxz_csr(12, dtype=np.complex128), dict(find_max=False, num_eigs=1), np.complex128),
    }


def expm_cases(wl):
    """name -> (csr, a, input, kwargs)."""
    n = 100
    ring = np.zeros((n, n))
    for i in range(n):
        ring[i, (i + 1) % n] = ring[(i + 1) % n, i] = -1.0
    x = np.zeros(n, complex)
    x[0] = 1 + 2j
    x[n - 1] = 1 + 2j
    x[n // 2] = 8 + 2j
    x /= np.linalg.norm(x)
    xxz = wl.xxz_csr(12, dtype=np.complex128)
    return {
        "expm_ring100_3i": (wl.dense_to_csr(ring.astype(complex)), 3j, x, dict()),
        "expm_ring100_zero_fullorth": (wl.dense_to_csr(ring.astype(complex)), 0j, x, dict(full_orth=True)),
        "expm_xxz12_neel": (xxz, -0.1j, wl.neel_state(12), dict()),
        "expm_xxz12_neel_fullorth": (xxz, -0.1j, wl.neel_state(12), dict(full_orth=True)),
        # real `a`: exp(aT) is not unitary, so the reference's overlap test (exponentiator.hpp:155) never fires and the
        # run ends at max_iteration — kept as a parity case for exactly that behaviour
        "expm_real_sym_cap30": (wl.random_symmetric_csr(500), -0.5, wl.start_vector(500), dict(max_iter=30)),
    }


def main():
    sys.path.insert(0, ROOT)
    import importlib

    import __graft_entry__ as entry
    import oracle

    entry.load_package()
    wl = importlib.import_module("lambda_lanczos_b200.workloads")
    oracle.build()
    ref = oracle.Reference()
    out = {}
    for name, (csr, kw, dt) in cases(wl).items():
        n = csr[0].size - 1
        r = ref.lanczos(*csr, init=wl.start_vector(n, dt), **kw)
        out[f"{name}/iters"] = np.array(r.iter_counts, dtype=np.int64)
        out[f"{name}/evals"] = r.eigenvalues
        out[f"{name}/evecs"] = r.eigenvectors
        print(name, r.iter_counts, r.eigenvalues)
    for name, (csr, a, x, kw) in expm_cases(wl).items():
        it, o = ref.expm(*csr, a, x, **kw)
        out[f"{name}/iters"] = np.array(it, dtype=np.int64)
        out[f"{name}/out"] = o
        print(name, it)
    np.savez_compressed(os.path.join(HERE, "reference_runs.npz"), **out)



def sample_outputs():
    """stdout of the reference's own sample programs (src/samples/sample{1,2,3,5}*.cpp, built against the reference's
    headers with g++ -O2) -> tests/golden/reference_sample_outputs.json.  tests/test_gpu_cpp_api.py builds the SAME,
    unmodified sources against this engine's compat headers and compares eigenvalues and eigenvectors."""
    import json
    import subprocess
    import tempfile

    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for s in ("1_simple", "2_sparse", "3_dynamic", "5_multiroot"):
            exe = os.path.join(tmp, "ref_sample" + s)
            subprocess.check_call(["g++", "-std=c++17", "-O2", "-I/root/reference/include/lambda_lanczos",
                                   f"/root/reference/src/samples/sample{s}.cpp", "-o", exe])
            out["sample" + s] = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_sample_outputs.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
    sample_outputs()
