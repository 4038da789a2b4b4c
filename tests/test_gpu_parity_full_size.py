"""Parity at (or near) BASELINE.json's named sizes, on the operator storage the benchmarks run.

The compiled reference cannot finish these workloads, but its first iterations are within reach: its public
run_iteration() (lambda_lanczos.hpp:216-322) is run for m iterations with an mv_mul spy (oracle/ref_shim.cpp) and
alpha_k, beta_k, the Lanczos vectors and the Ritz values are compared with the CUDA engine's on the same operator and
start vector.  Tolerances are the north star's (1e-10 relative, overlap 1 - 1e-9; 1e-5 / 1e-4 for float)."""
import math
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def step_level_parity(pkg, ctx, op, start, ref, m, tol=1e-10, overlap_tol=1e-9):
    n = start.size
    kry = pkg.Krylov(ctx, start.dtype, n, m + 2)
    kry.set_locked([])
    kry.begin(start)
    ab = []
    for k in range(1, m + 1):
        kry.step(op, 0.0, pkg.ORTH_FULL)
        ab.append(kry.fetch(k))
    alpha = np.array([a for a, _ in ab])
    beta = np.array([b for _, b in ab])
    assert np.max(np.abs(alpha - ref["alpha"][:m]) / np.maximum(np.abs(ref["alpha"][:m]), 1e-300)) < tol
    assert np.max(np.abs(beta[:m - 1] - ref["beta"][:m - 1]) / np.abs(ref["beta"][:m - 1])) < tol
    for j in range(min(m, len(ref["basis"]))):
        col = kry.column(j)
        assert abs(abs(np.vdot(ref["basis"][j], col)) - 1.0) < overlap_tol, j
    T = np.diag(alpha) + np.diag(beta[:-1], 1) + np.diag(beta[:-1], -1)
    kry.close()
    return alpha, beta, np.linalg.eigvalsh(T)


def reference_or_skip(oracle_mod):
    if not oracle_mod.have_reference():
        pytest.skip("needs the compiled reference (oracle/_ref/libllz_ref.so travels with the snapshot)")
    return oracle_mod.Reference()


def test_config2_full_size_sell_against_reference(pkg, ctx, wl, oracle_mod):
    """Config 2 at its named size (Laplacian 4096^2, n = 16 777 216) on the SELL-32-sigma operator bench.py times:
    16 Lanczos iterations against the reference, then the engine's capped run against the reference's capped run."""
    ref = reference_or_skip(oracle_mod)
    nx, m = 4096, 16
    csr = wl.laplacian2d_csr(nx)
    n = nx * nx
    start = wl.start_vector(n)
    r = ref.run_iteration(*csr, find_max=False, max_iter=m, init=start, mv_threads=ref.host_threads(), capture=m)
    op = pkg.Operator.sell(ctx, *csr)
    alpha, beta, ritz = step_level_parity(pkg, ctx, op, start, r, m)
    k = r["eigenvalues"].size
    assert np.max(np.abs(ritz[:k] - r["eigenvalues"]) / np.abs(r["eigenvalues"])) < 1e-10
    eng = pkg.LambdaLanczos(op, n, False, 1)
    eng.init_vector = start
    eng.max_iteration = m
    ev, vec = eng.run()
    assert eng.getIterationCounts() == [m]
    assert abs(ev[0] - r["eigenvalues"][0]) <= 1e-10 * abs(r["eigenvalues"][0])
    ctx.release_cache()


def test_config3_full_size_sell_complex_against_reference(pkg, ctx, wl, oracle_mod):
    """Config 3 at its named size (Peierls tight-binding 2896^2, n = 8 386 816, complex128), SELL storage."""
    ref = reference_or_skip(oracle_mod)
    lx, m = 2896, 12
    csr = wl.peierls_csr(lx, lx, flux=0.05, trap=0.02)
    n = lx * lx
    start = wl.start_vector(n, np.complex128)
    r = ref.run_iteration(*csr, find_max=False, max_iter=m, init=start, mv_threads=ref.host_threads(), capture=m)
    op = pkg.Operator.sell(ctx, *csr)
    step_level_parity(pkg, ctx, op, start, r, m)
    ctx.release_cache()


def test_config4_L24_matrix_free_against_reference(pkg, ctx, wl, oracle_mod):
    """Config 4 at L = 24 (2 704 156 states): the matrix-free operator against the reference driven by the explicit
    matrix, 24 iterations step by step; and the converged ground state at L = 22 (705 432 states), iteration counts equal."""
    ref = reference_or_skip(oracle_mod)
    L, m = 24, 24
    csr = wl.xxz_csr(L)
    n = csr[0].size - 1
    assert n == math.comb(L, L // 2)
    start = wl.start_vector(n)
    r = ref.run_iteration(*csr, find_max=False, max_iter=m, init=start, mv_threads=ref.host_threads(), capture=m)
    step_level_parity(pkg, ctx, pkg.Operator.xxz(ctx, L), start, r, m)
    del csr, r
    L = 22
    csr = wl.xxz_csr(L)
    n = csr[0].size - 1
    start = wl.start_vector(n)
    eng = pkg.LambdaLanczos(pkg.Operator.xxz(ctx, L), n, False, 1)
    eng.init_vector = start
    ev, vec = eng.run()
    rr = ref.lanczos(*csr, find_max=False, num_eigs=1, init=start, mv_threads=ref.host_threads())
    print(f"xxz L={L}: E0 ours {ev[0]!r} reference {rr.eigenvalues[0]!r} iterations ours {eng.getIterationCounts()} reference {rr.iter_counts}")
    assert abs(ev[0] - rr.eigenvalues[0]) <= 1e-10 * abs(rr.eigenvalues[0])
    assert 1 - abs(np.vdot(rr.eigenvectors[0], vec[0])) < 1e-9
    assert eng.getIterationCounts() == rr.iter_counts
    ctx.release_cache()


def test_config5_L20_twelve_steps_against_reference(pkg, ctx, wl, oracle_mod):
    """Config 5 at L = 20 (184 756 states, complex128): 12 time steps of exp(-i H 0.1) from the Neel state, every step's
    iteration count equal to the reference's and the state within 1e-10 relative L2 after every step."""
    ref = reference_or_skip(oracle_mod)
    L = 20
    csr = wl.xxz_csr(L, dtype=np.complex128)
    op = pkg.Operator.xxz(ctx, L, dtype=np.complex128)
    ex = pkg.Exponentiator(op, op.n)
    cur = wl.neel_state(L)
    cur_ref = cur.copy()
    for step in range(12):
        it, cur = ex.run(-0.1j, cur)
        it_ref, cur_ref = ref.expm(*csr, -0.1j, cur_ref)
        assert it == it_ref, (step, it, it_ref)
        assert np.linalg.norm(cur - cur_ref) <= 1e-10 * np.linalg.norm(cur_ref), step
    assert abs(np.linalg.norm(cur) - 1.0) < 1e-12


def test_config1_full_size_float_and_double_against_reference(pkg, ctx, wl, oracle_mod):
    """Config 1 at its named size (n = 100 000, ~17 non-zeros per row) in double AND float, to convergence, SELL
    storage: eigenvalue, eigenvector overlap, and the iteration count EQUAL to the reference's."""
    ref = reference_or_skip(oracle_mod)
    n = 100000
    for dtype, tol, ov_tol in ((np.float64, 1e-10, 1e-9), (np.float32, 1e-5, 1e-4)):
        csr = wl.random_symmetric_csr(n, dtype=dtype)
        start = wl.start_vector(n, dtype)
        eng = pkg.LambdaLanczos(pkg.Operator.sell(ctx, *csr), n, True, 1)
        eng.init_vector = start
        ev, vec = eng.run()
        rr = ref.lanczos(*csr, find_max=True, num_eigs=1, init=start, mv_threads=ref.host_threads())
        print(f"config1 {np.dtype(dtype).name}: ours {ev[0]!r} reference {rr.eigenvalues[0]!r} iterations ours {eng.getIterationCounts()} reference {rr.iter_counts}")
        assert abs(ev[0] - rr.eigenvalues[0]) <= tol * abs(rr.eigenvalues[0])
        assert 1 - abs(np.vdot(rr.eigenvectors[0].astype(np.float64), vec[0].astype(np.float64))) < ov_tol
        if dtype == np.float64:
            assert eng.getIterationCounts() == rr.iter_counts
        else:  # float: the stopping test compares Ritz values at 1e-4 relative, rounding decides the last iteration
            assert abs(eng.getIterationCounts()[0] - rr.iter_counts[0]) <= 2


@pytest.mark.parametrize("case", ["peierls", "random"])
def test_complex_float_engine_against_oracle(pkg, ctx, wl, oracle_mod, case):
    """std::complex<float> (LLZ_C64) end to end: the reference template covers it (lambda_lanczos.hpp:109,
    util/common.hpp:80-102), so does the engine; checked against the c64 instantiation of the oracle."""
    chk = oracle_mod.best()
    if case == "peierls":
        full = wl.peierls_csr(24, 20, flux=0.05, trap=0.3)
        csr = (full[0], full[1], full[2].astype(np.complex64))
        find_max, k = False, 2
    else:
        csr = wl.random_symmetric_csr(3000, 6, dtype=np.complex64)
        find_max, k = True, 1
    n = csr[0].size - 1
    start = wl.start_vector(n, np.complex64)
    for make in (pkg.Operator.csr, pkg.Operator.sell):
        eng = pkg.LambdaLanczos(make(ctx, *csr), n, find_max, k)
        eng.init_vector = start
        ev, vec = eng.run()
        rr = chk.lanczos(*csr, find_max=find_max, num_eigs=k, init=start)
        assert np.allclose(ev, rr.eigenvalues[:k], rtol=1e-5, atol=1e-6), (ev, rr.eigenvalues)
        for i in range(k):
            assert 1 - abs(np.vdot(rr.eigenvectors[i].astype(np.complex128), vec[i].astype(np.complex128))) < 1e-4, i
    x = start
    op = pkg.Operator.csr(ctx, *csr)
    it, out = pkg.Exponentiator(op, n).run(-0.05j, x / np.linalg.norm(x))
    it_ref, out_ref = chk.expm(*csr, -0.05j, (x / np.linalg.norm(x)).astype(np.complex64))
    assert abs(it - it_ref) <= 1 and np.linalg.norm(out - out_ref) <= 1e-4 * np.linalg.norm(out_ref)
