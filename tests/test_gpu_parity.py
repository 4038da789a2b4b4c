"""GPU parity tests: the CUDA path (through the C ABI, include/llz.h) against the CPU oracle on the same seeded inputs,
against the committed golden fixtures of the compiled reference, and — at BASELINE.json's full sizes — through
size-independent properties.

Tolerances are the north star's: eigenvalues 1e-10 relative (double) / 1e-5 (float); |<v_ref, v>| >= 1 - 1e-9;
residual no worse than the reference's (up to rounding noise, stated per test); Exponentiator 1e-10 relative L2.
"""
import math
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")

EVAL_TOL = {np.dtype(np.float32): 1e-5, np.dtype(np.float64): 1e-10, np.dtype(np.complex128): 1e-10}


def rnd(rs, n, dtype):
    v = rs.uniform(-1, 1, n)
    if np.dtype(dtype).kind == "c":
        v = v + 1j * rs.uniform(-1, 1, n)
    return v.astype(dtype)


def residual(wl, csr, lam, v):
    return np.linalg.norm(wl.csr_matvec(*csr, v.astype(np.complex128 if v.dtype.kind == "c" else np.float64)) - lam * v)


# ---- vector kernels (util/linear_algebra.hpp) ------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex64, np.complex128])
@pytest.mark.parametrize("n", [1, 3, 1000, 4099, 300001])
def test_dot_norm_axpy_scale(pkg, ctx, dtype, n):
    rs = np.random.RandomState(n)
    a, b = rnd(rs, n, dtype), rnd(rs, n, dtype)
    va, vb = pkg.Vector.from_host(ctx, a), pkg.Vector.from_host(ctx, b)
    exact = np.vdot(a.astype(np.complex128), b.astype(np.complex128))  # conjugates the first argument
    got = va.dot(vb)
    tol = (1e-5 if np.dtype(dtype).itemsize // (2 if np.dtype(dtype).kind == "c" else 1) == 4 else 1e-13) * max(1.0, math.sqrt(n))
    assert abs(complex(got) - exact) <= tol * max(1.0, abs(exact))
    assert abs(va.norm() - np.linalg.norm(a.astype(np.complex128))) <= tol * max(1.0, np.linalg.norm(a))
    va.axpy(0.5 - (0.25j if np.dtype(dtype).kind == "c" else 0), vb)
    expect = a + (0.5 - (0.25j if np.dtype(dtype).kind == "c" else 0)) * b
    assert np.allclose(va.download(), expect.astype(dtype), rtol=1e-5 if tol > 1e-8 else 1e-14, atol=1e-6 if tol > 1e-8 else 1e-14)
    nrm = va.normalize()
    assert abs(nrm - np.linalg.norm(expect)) <= tol * 10 * max(1.0, nrm)
    assert abs(va.norm() - 1.0) <= (1e-6 if tol > 1e-8 else 1e-14)


def test_inner_product_known_answer(pkg, ctx):  # lambda_lanczos_test.cpp:47-59
    v1 = pkg.Vector.from_host(ctx, np.array([3.0, 1 + 3j]))
    v2 = pkg.Vector.from_host(ctx, np.array([3.0, 2 + 4j]))
    assert v1.dot(v2) == complex(23.0, -2.0)


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex128])
@pytest.mark.parametrize("n,count", [(10, 5), (4099, 9), (100003, 40)])
def test_schmidt_orth_against_oracle(pkg, ctx, port, dtype, n, count):
    """CGS (GPU) and MGS (oracle) project onto the same complement when the basis is orthonormal."""
    rs = np.random.RandomState(count)
    basis = []
    for _ in range(count):  # orthonormal basis built by the ORACLE's own MGS + normalize
        u = rnd(rs, n, dtype)
        if basis:
            u = port.schmidt_orth(np.array(basis), u)
        basis.append((u / port.norm(u)).astype(dtype))
    w = rnd(rs, n, dtype)
    expect = port.schmidt_orth(np.array(basis), w)
    vw = pkg.Vector.from_host(ctx, w)
    vb = [pkg.Vector.from_host(ctx, u) for u in basis]
    vw.schmidt_orth(vb, passes=1)
    got = vw.download()
    single = np.dtype(dtype).itemsize // (2 if np.dtype(dtype).kind == "c" else 1) == 4
    tol = (2e-5 if single else 1e-13) * np.linalg.norm(w)
    assert np.linalg.norm(got - expect) <= tol
    for u in basis:  # lambda_lanczos_test.cpp:61-91: result is orthogonal to every basis vector
        assert abs(np.vdot(u, got)) <= (5e-5 if single else 1e-13) * np.linalg.norm(w)


# ---- operators --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex128])
def test_csr_operator_matches_numpy(pkg, ctx, wl, dtype):
    for csr in (wl.random_symmetric_csr(5003, dtype=np.float64), wl.laplacian2d_csr(37, 41), wl.peierls_csr(23, 19)):
        if np.dtype(dtype).kind != "c" and csr[2].dtype.kind == "c":
            continue
        csr = (csr[0], csr[1], csr[2].astype(dtype))
        n = csr[0].size - 1
        x = rnd(np.random.RandomState(1), n, dtype)
        op = pkg.Operator.csr(ctx, *csr)
        y = op.matvec(x)
        ref = wl.csr_matvec(*csr, x)
        assert np.allclose(y, ref, rtol=2e-5 if np.dtype(dtype).itemsize <= 8 and np.dtype(dtype) != np.float64 else 1e-13,
                           atol=1e-5 if np.dtype(dtype) == np.float32 else 1e-13)
        assert op.bytes() > 0


def ragged_csr(n, seed=0, dtype=np.float64):
    """Rows of wildly different lengths, including empty ones and one very long row."""
    import scipy.sparse as sp

    rs = np.random.RandomState(seed)
    lens = rs.choice([0, 1, 2, 3, 5, 9, 40], size=n, p=[0.1, 0.2, 0.2, 0.2, 0.15, 0.1, 0.05])
    lens[n // 2] = min(n, 700)
    rows = np.repeat(np.arange(n), lens)
    cols = rs.randint(0, n, size=rows.size)
    vals = rs.uniform(-1, 1, size=rows.size)
    a = sp.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsr()
    a.sum_duplicates()
    a.sort_indices()
    v = a.data.astype(dtype)
    if np.dtype(dtype).kind == "c":
        v = v * np.exp(1j * rs.uniform(0, 6, size=v.size)).astype(dtype)
    return a.indptr.astype(np.int64), a.indices.astype(np.int32), v


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex64, np.complex128])
@pytest.mark.parametrize("sigma", [0, 1, 32, 256, 1024])
def test_sell_operator_is_bit_identical_to_csr(pkg, ctx, wl, dtype, sigma):
    """SELL-C-sigma keeps the per-row summation order of CSR: y must match bit for bit, whatever the slicing/sorting,
    for regular stencils, random matrices and ragged ones (empty rows, one 700-entry row, n not a multiple of 32)."""
    cases = [wl.laplacian2d_csr(37, 41, dtype=np.float64), wl.random_symmetric_csr(5003), ragged_csr(3001, 1), ragged_csr(31, 2), ragged_csr(1, 3)]
    if np.dtype(dtype).kind == "c":
        cases.append(wl.peierls_csr(23, 19))
        cases.append(ragged_csr(2050, 4, np.complex128))
    for csr in cases:
        csr = (csr[0], csr[1], csr[2].astype(dtype))
        n = csr[0].size - 1
        x = rnd(np.random.RandomState(1), n, dtype)
        y_csr = pkg.Operator.csr(ctx, *csr).matvec(x)
        op = pkg.Operator.sell(ctx, *csr, sigma=sigma)
        y = op.matvec(x)
        assert np.array_equal(y, y_csr), (n, sigma)
        assert op.bytes() > 0
    # padding entries are skipped, never multiplied: an Inf in x only reaches the rows that reference it
    csr = ragged_csr(3001, 1, dtype)
    x = rnd(np.random.RandomState(2), 3001, dtype)
    x[17] = np.inf
    y = pkg.Operator.sell(ctx, *csr, sigma=sigma).matvec(x)
    touched = np.zeros(3001, bool)
    touched[np.repeat(np.arange(3001), np.diff(csr[0]))[csr[1] == 17]] = True
    touched[17] = True  # y_17 also holds sigma * x_17 (0 * inf)
    assert np.all(np.isfinite(y[~touched]))


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex128])
def test_gerschgorin_radius_matches_numpy(pkg, ctx, wl, dtype):
    """max abs row sum on the device (determine_eigenvalue_offset.cpp:12-29 of the reference) for CSR and SELL storage;
    used as eigenvalue_offset it turns the smallest eigenvalue into the dominant one."""
    for csr in (ragged_csr(3001, 1, dtype), wl.laplacian2d_csr(37, 41, dtype=dtype)):
        expect = np.abs(wl_dense_rowsum(csr))
        for make in (pkg.Operator.csr, pkg.Operator.sell):
            got = make(ctx, *csr).gerschgorin_radius()
            assert abs(got - expect) <= (1e-5 if np.dtype(dtype) == np.float32 else 1e-12) * expect
    op = pkg.Operator.xxz(ctx, 12)
    dense = np.abs(wl_dense_rowsum(wl.xxz_csr(12)))
    assert op.gerschgorin_radius() >= dense - 1e-12


def wl_dense_rowsum(csr):
    rowptr, colidx, vals = csr
    sums = np.add.reduceat(np.abs(vals.astype(np.complex128)), rowptr[:-1][np.diff(rowptr) > 0]) if vals.size else np.zeros(1)
    return float(sums.max())


def test_sell_rejects_bad_sigma(pkg, ctx, wl):
    with pytest.raises(pkg.LlzError) as e:
        pkg.Operator.sell(ctx, *wl.laplacian2d_csr(8), sigma=48)
    assert e.value.status == 1


def test_lanczos_on_sell_matches_csr_run(pkg, ctx, wl):
    """Same operator in both formats => identical Lanczos run (y is bit-identical; alpha's partial sums are grouped
    differently, so scalars agree to rounding and eigenpairs to the parity tolerance)."""
    csr = wl.laplacian2d_csr(48)
    n = 48 * 48
    out = []
    for make in (pkg.Operator.csr, pkg.Operator.sell):
        eng = pkg.LambdaLanczos(make(ctx, *csr), n, False, 4)
        eng.init_vector = wl.start_vector(n)
        ev, vec = eng.run()
        out.append((ev, vec, eng.getIterationCounts()))
    assert np.allclose(out[0][0], out[1][0], rtol=1e-10, atol=0)
    assert np.allclose(out[1][0], wl.laplacian2d_exact(48, count=4), rtol=1e-10, atol=0)
    assert abs(abs(np.vdot(out[0][1][0], out[1][1][0])) - 1) < 1e-9


def test_workspace_cache_reuse_is_invisible(pkg, ctx, wl):
    """The context revives the Krylov workspace and vector buffers of the previous run: results are bit-identical to a
    run on released caches, and different shapes in between do not confuse it."""
    csr = wl.random_symmetric_csr(20000)
    runs = []
    for i in range(3):
        if i == 1:
            ctx.release_cache()
        if i == 2:  # a run of another shape and dtype in between
            e2 = pkg.LambdaLanczos(pkg.Operator.csr(ctx, *wl.peierls_csr(12, 12)), 144, False, 2)
            e2.init_vector = wl.start_vector(144, np.complex128)
            e2.run()
        eng = pkg.LambdaLanczos(pkg.Operator.csr(ctx, *csr), 20000, True, 1)
        eng.init_vector = wl.start_vector(20000)
        ev, vec = eng.run()
        runs.append((ev, vec, eng.getIterationCounts()))
    for ev, vec, it in runs[1:]:
        assert np.array_equal(ev, runs[0][0]) and np.array_equal(vec, runs[0][1]) and it == runs[0][2]


@pytest.mark.parametrize("L,n_up,pbc", [(4, 2, True), (8, 4, True), (10, 3, False), (14, 7, True), (16, 8, True), (13, 6, True)])
@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_xxz_matrix_free_matches_explicit_matrix(pkg, ctx, wl, L, n_up, pbc, dtype):
    csr = wl.xxz_csr(L, n_up, jz=0.7, jxy=1.3, pbc=pbc, dtype=dtype)
    n = csr[0].size - 1
    op = pkg.Operator.xxz(ctx, L, n_up, jz=0.7, jxy=1.3, periodic=pbc, dtype=dtype)
    assert op.n == n == math.comb(L, n_up)
    x = rnd(np.random.RandomState(L), n, dtype)
    assert np.allclose(op.matvec(x), wl.csr_matvec(*csr, x), rtol=1e-13, atol=1e-13)
    assert op.bytes() < 64 * 1024 + 12 * 2 ** max(L - 6, 0)  # popcount-class tables and a block list: nothing per state


@pytest.mark.parametrize("L,n_up,pbc", [(2, 1, True), (3, 1, True), (5, 2, False), (9, 4, True), (12, 6, True), (15, 7, False),
                                        (18, 9, True), (20, 10, True), (20, 4, True), (17, 15, True), (6, 0, True), (6, 6, True)])
@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex64, np.complex128])
def test_xxz_block_kernel_is_bit_identical_to_per_state_kernel(pkg, ctx, wl, monkeypatch, L, n_up, pbc, dtype):
    """The block kernel (low-bit bonds through shared-memory neighbour lists, high-bit bonds as contiguous streams) adds
    the products of a row in the same bond order as the plain one-thread-per-state kernel: same bits, for every split of
    the bit string (no high bits at all, one high bit, the default, the largest supported low part)."""
    n = math.comb(L, n_up)
    x = rnd(np.random.RandomState(7 * L + n_up), n, dtype)
    monkeypatch.setenv("LLZ_XXZ_KERNEL", "state")
    y_ref = pkg.Operator.xxz(ctx, L, n_up, jz=0.7, jxy=1.3, periodic=pbc, dtype=dtype).matvec(x)
    monkeypatch.delenv("LLZ_XXZ_KERNEL")
    for m in sorted({0, L, L - 1, 1, 2, min(L, 14), max(1, L // 2)}):
        if m == 0:
            monkeypatch.delenv("LLZ_XXZ_M", raising=False)  # the default split
        else:
            monkeypatch.setenv("LLZ_XXZ_M", str(m))
        y = pkg.Operator.xxz(ctx, L, n_up, jz=0.7, jxy=1.3, periodic=pbc, dtype=dtype).matvec(x)
        assert np.array_equal(y.view(np.uint8), y_ref.view(np.uint8)), (m, np.abs(y - y_ref).max())


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex64, np.complex128])
def test_dia_storage_is_picked_for_stencils_and_is_bit_identical_to_csr(pkg, ctx, wl, dtype):
    """llz_op_create_sell(sigma = 0) re-stores operators whose non-zeros lie on <= 16 diagonals diagonal by diagonal (no
    column indices, contiguous shifted x reads); y must equal the CSR kernels' bit for bit (same products, same order),
    absent entries (grid boundaries) are skipped, and anything else falls back to SELL."""
    import scipy.sparse as sp

    cases = {"laplacian 37x41": wl.laplacian2d_csr(37, 41, dtype=dtype), "chain": wl.dense_to_csr(np.diag(np.ones(9), 1) + np.diag(np.ones(9), -1)),
             "laplacian 64x64": wl.laplacian2d_csr(64, dtype=dtype)}
    if np.dtype(dtype).kind == "c":
        full = wl.peierls_csr(33, 29, flux=0.07, trap=0.1)
        cases["peierls 33x29"] = (full[0], full[1], full[2].astype(dtype))
    for name, csr in cases.items():
        csr = (csr[0], csr[1], csr[2].astype(dtype))
        n = csr[0].size - 1
        op = pkg.Operator.sell(ctx, *csr)
        assert op.storage() == "DIA", (name, op.storage())
        x = rnd(np.random.RandomState(3), n, dtype)
        y_csr = pkg.Operator.csr(ctx, *csr).matvec(x)
        assert np.array_equal(op.matvec(x).view(np.uint8), y_csr.view(np.uint8)), name
        assert op.bytes() < pkg.Operator.sell(ctx, *csr, sigma=1).bytes()
        xi = x.copy()
        xi[n // 2] = np.inf  # an Inf in x reaches exactly the rows that reference it
        assert np.array_equal(np.isfinite(op.matvec(xi)), np.isfinite(pkg.Operator.csr(ctx, *csr).matvec(xi))), name
        assert abs(op.gerschgorin_radius() - pkg.Operator.csr(ctx, *csr).gerschgorin_radius()) < 1e-5
    # 17 diagonals, or non-zeros off the sampled diagonals: SELL
    n = 500
    many = sp.diags([np.ones(n - abs(k)) for k in range(-8, 9)], list(range(-8, 9)), format="csr")
    csr = (many.indptr.astype(np.int64), many.indices.astype(np.int32), many.data.astype(dtype))
    assert pkg.Operator.sell(ctx, *csr).storage().startswith("SELL")
    stray = sp.lil_matrix(sp.diags([np.ones(n - 1), np.ones(n - 1)], [-1, 1], format="csr").astype(np.float64))
    stray[301, 17] = stray[17, 301] = 0.5  # rows the sample of candidate diagonals does not visit... or visits: SELL either way
    stray = sp.csr_matrix(stray)
    stray.sort_indices()
    csr = (stray.indptr.astype(np.int64), stray.indices.astype(np.int32), stray.data.astype(dtype))
    op = pkg.Operator.sell(ctx, *csr)
    x = rnd(np.random.RandomState(4), n, dtype)
    assert np.array_equal(op.matvec(x).view(np.uint8), pkg.Operator.csr(ctx, *csr).matvec(x).view(np.uint8))


# ---- the Lanczos recurrence itself: alpha, beta and the basis, iteration by iteration -------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_lanczos_vectors_match_oracle_iteration_by_iteration(pkg, ctx, wl, oracle_mod, port, dtype):
    n, steps = 20011, 40
    csr = wl.random_symmetric_csr(n) if dtype == np.float64 else wl.peierls_csr(173, 113)
    n = csr[0].size - 1
    start = wl.start_vector(n, dtype)
    it, _, _, alpha, beta = port.run_iteration(*csr, find_max=True, max_iter=steps, nroot=5, init=start)
    assert it == steps
    op = pkg.Operator.csr(ctx, *csr)
    kry = pkg.Krylov(ctx, dtype, n, steps + 1)
    kry.begin(start)
    got_a, got_b = [], []
    for k in range(1, steps + 1):
        kry.step(op, 0.0, pkg.ORTH_FULL)
        a, b = kry.fetch(k)
        got_a.append(a)
        got_b.append(b)
    # beta[-1] is forced to 0 by the oracle after the loop (lambda_lanczos.hpp:314); compare the others
    assert np.allclose(got_a, alpha, rtol=1e-11, atol=1e-12)
    assert np.allclose(got_b[:-1], beta[:-1], rtol=1e-11, atol=1e-12)
    # the basis: orthonormal to rounding, and spanning the same Krylov sequence as the reference's
    V = np.array([kry.column(j) for j in range(steps + 1)])
    G = V.conj() @ V.T
    assert np.abs(G - np.eye(steps + 1)).max() < 1e-13
    if oracle_mod.have_reference():
        ref = oracle_mod.Reference().lanczos(*csr, find_max=True, num_eigs=1, max_iter=steps, init=start, capture=steps)
        for j in range(steps):
            assert abs(abs(np.vdot(ref.basis[j], V[j])) - 1.0) < 1e-9, j


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex128])
def test_lazy_recurrence_is_the_recurrence_without_the_normalisation_pass(pkg, ctx, wl, dtype):
    """LLZ_ORTH_RECURRENCE_LAZY (what the Exponentiator runs on one GPU): same alpha_k, beta_k to rounding, column k
    stored as beta_{k-1} u_k."""
    csr = wl.peierls_csr(30, 27) if np.dtype(dtype).kind == "c" else wl.random_symmetric_csr(6007, 7, dtype=dtype)
    csr = (csr[0], csr[1], csr[2].astype(dtype))
    n = csr[0].size - 1
    op = pkg.Operator.csr(ctx, *csr)
    start = wl.start_vector(n, dtype)
    steps = 12
    out = {}
    for mode in (pkg.ORTH_RECURRENCE, pkg.ORTH_RECURRENCE_LAZY):
        kry = pkg.Krylov(ctx, dtype, n, steps + 2)
        kry.begin(start)
        ab = []
        for k in range(1, steps + 1):
            kry.step(op, 0.25, mode)
            ab.append(kry.fetch(k))
        out[mode] = (np.array(ab), [kry.column(j) for j in range(steps + 1)])
        kry.close()
    tol = 2e-5 if np.dtype(dtype) == np.float32 else 1e-12
    ab0, cols0 = out[pkg.ORTH_RECURRENCE]
    ab1, cols1 = out[pkg.ORTH_RECURRENCE_LAZY]
    assert np.allclose(ab0, ab1, rtol=tol, atol=tol)
    for j in range(steps + 1):
        scale = 1.0 if j == 0 else ab1[j - 1, 1]
        assert np.allclose(cols1[j] / scale, cols0[j], rtol=0, atol=50 * tol), j


# ---- LambdaLanczos::run against the checker --------------------------------------------------------------------
def run_both(pkg, ctx, checker, csr, dtype, start, **kw):
    n = csr[0].size - 1
    op = pkg.Operator.csr(ctx, *csr)
    eng = pkg.LambdaLanczos(op, n, kw["find_max"], kw.get("num_eigs", 1))
    eng.init_vector = start
    eng.eigenvalue_offset = kw.get("offset", 0.0)
    if "eps" in kw:
        eng.eps = kw["eps"]
    if "max_iter" in kw:
        eng.max_iteration = kw["max_iter"]
    evals, evecs = eng.run()
    ref = checker.lanczos(*csr, init=start, **kw)
    return eng, evals, evecs, ref


def assert_eigenpairs(wl, csr, dtype, evals, evecs, ref, clusters=None, res_slack=50.0):
    tol = EVAL_TOL[np.dtype(dtype)]
    assert evals.size == ref.eigenvalues.size
    scale = max(1.0, np.abs(ref.eigenvalues).max())
    for i in range(evals.size):
        assert abs(evals[i] - ref.eigenvalues[i]) <= tol * max(abs(ref.eigenvalues[i]), 1e-300) or \
            abs(evals[i] - ref.eigenvalues[i]) <= 1e-13 * scale, (i, evals[i], ref.eigenvalues[i])
    clusters = clusters or [[i] for i in range(evals.size)]
    for cl in clusters:
        A = np.array([ref.eigenvectors[i] for i in cl])
        B = np.array([evecs[i] for i in cl])
        sv = np.linalg.svd(A.conj() @ B.T, compute_uv=False)  # cosines of the principal angles between the subspaces
        assert sv.min() >= 1 - (1e-9 if np.dtype(dtype) != np.float32 else 1e-4), (cl, sv)
        r_ref = max(residual(wl, csr, ref.eigenvalues[i], ref.eigenvectors[i]) for i in cl)
        r_got = max(residual(wl, csr, evals[i], evecs[i]) for i in cl)
        floor = (1e-4 if np.dtype(dtype) == np.float32 else 1e-12) * scale
        assert r_got <= max(res_slack * r_ref, floor), (cl, r_got, r_ref)


def test_simple_matrix_max_with_offset(pkg, ctx, wl, checker):  # lambda_lanczos_test.cpp:128-161
    a = np.array([[2.0, 1, 1], [1, 2, 1], [1, 1, 2]])
    csr = wl.dense_to_csr(a)
    eng, evals, evecs, ref = run_both(pkg, ctx, checker, csr, np.float64, wl.start_vector(3), find_max=True, num_eigs=1, offset=6.0)
    assert abs(evals[0] - 4.0) < 4.0 * eng.eps
    v = evecs[0] * np.sign(evecs[0][0])
    assert np.allclose(v, np.ones(3) / math.sqrt(3), atol=4.0 * eng.eps * 10)
    assert len(eng.getIterationCounts()) == 1
    assert_eigenpairs(wl, csr, np.float64, evals, evecs, ref)


def test_simple_matrix_float(pkg, ctx, wl, checker):  # :163-193
    a = np.array([[2.0, 1, 1], [1, 2, 1], [1, 1, 2]], dtype=np.float32)
    csr = wl.dense_to_csr(a)
    eng, evals, evecs, ref = run_both(pkg, ctx, checker, csr, np.float32, wl.start_vector(3, np.float32), find_max=True, num_eigs=1)
    assert abs(evals[0] - 4.0) < 4.0 * eng.eps
    assert_eigenpairs(wl, csr, np.float32, evals, evecs, ref)


def test_dynamic_matrix_min_with_negative_offset(pkg, ctx, wl, checker):  # :262-308
    n = 10
    a = np.zeros((n, n))
    for i in range(n - 1):
        a[i, i + 1] = a[i + 1, i] = -1.0
    csr = wl.dense_to_csr(a)
    eng, evals, evecs, ref = run_both(pkg, ctx, checker, csr, np.float64, wl.start_vector(n), find_max=False, num_eigs=1, offset=-10.0, eps=1e-14)
    lam = -2.0 * math.cos(math.pi / (n + 1))
    vec = np.sin(np.arange(1, n + 1) * math.pi / (n + 1))
    vec /= np.linalg.norm(vec)
    assert abs(evals[0] - lam) < abs(lam) * 1e-14 * 4
    assert np.allclose(evecs[0] * np.sign(evecs[0][0]), vec, atol=abs(lam) * 1e-13)
    assert_eigenpairs(wl, csr, np.float64, evals, evecs, ref)


def test_hermitian_matrix(pkg, ctx, wl, checker):  # :375-409
    h = np.array([[0, 1j, 1], [-1j, 0, 1j], [1, -1j, 0]])
    csr = wl.dense_to_csr(h)
    eng, evals, evecs, ref = run_both(pkg, ctx, checker, csr, np.complex128, wl.start_vector(3, np.complex128), find_max=False, num_eigs=1)
    assert abs(evals[0] + 2.0) < 2.0 * eng.eps
    v = evecs[0] * (np.conj(evecs[0][0]) / abs(evecs[0][0]))
    assert np.allclose(v, np.array([1, 1j, -1]) / math.sqrt(3), atol=2.0 * eng.eps * 10)
    assert_eigenpairs(wl, csr, np.complex128, evals, evecs, ref)


def test_single_element_matrix(pkg, ctx, wl):  # :411-440 (beta = 0 on the first iteration)
    op = pkg.Operator.csr(ctx, np.array([0, 1]), np.array([0]), np.array([2.0]))
    eng = pkg.LambdaLanczos(op, 1, True, 1)
    eng.init_vector = wl.start_vector(1)
    evals, evecs = eng.run()
    assert abs(evals[0] - 2.0) < 2.0 * eng.eps and abs(abs(evecs[0][0]) - 1.0) < 1e-12
    assert eng.getIterationCounts() == [1]


def test_multiple_eigenpairs_8x8(pkg, ctx, wl, checker):  # :442-488
    from test_oracle import EIGHT, EIGHT_VALS, EIGHT_VECS

    csr = wl.dense_to_csr(EIGHT)
    eng, evals, evecs, ref = run_both(pkg, ctx, checker, csr, np.float64, wl.start_vector(8), find_max=False, num_eigs=3, eps=1e-7)
    for i in range(3):
        assert abs(evals[i] - EIGHT_VALS[i]) < abs(EIGHT_VALS[i]) * 1e-7
        v = evecs[i] * np.sign(evecs[i][0])
        assert np.allclose(v, EIGHT_VECS[i] * np.sign(EIGHT_VECS[i][0]), atol=abs(EIGHT_VALS[i]) * 1e-6)


def test_multiple_degenerate_eigenpairs_ring50(pkg, ctx, wl, checker):  # :490-536 — 26 roots, 2-fold degeneracies
    n, num = 50, 26
    a = np.zeros((n, n))
    for i in range(n):
        a[i, (i + 1) % n] = a[(i + 1) % n, i] = -1.0
    csr = wl.dense_to_csr(a)
    eng, evals, evecs, ref = run_both(pkg, ctx, checker, csr, np.float64, wl.start_vector(n, seed=7), find_max=False, num_eigs=num, eps=1e-14)
    correct = np.sort(-2.0 * np.cos(2.0 * math.pi * np.arange(-num // 2, num - num // 2) / n))
    assert evals.size == num
    assert np.allclose(evals, correct, atol=1e-13)
    assert np.allclose(evals, ref.eigenvalues, atol=1e-13)
    print("iteration counts: ours", eng.getIterationCounts(), "reference", ref.iter_counts)


@pytest.mark.parametrize("dtype,n", [(np.float64, 100000), (np.float32, 30000)])
def test_config1_random_symmetric_max_eigenpair(pkg, ctx, wl, checker, dtype, n):
    """BASELINE.json config 1 (n = 100 000 for double): iteration counts side by side."""
    csr = wl.random_symmetric_csr(n, dtype=dtype)
    eng, evals, evecs, ref = run_both(pkg, ctx, checker, csr, dtype, wl.start_vector(n, dtype), find_max=True, num_eigs=1)
    print(f"config1 {np.dtype(dtype).name}: lambda ours {evals[0]!r} ref {ref.eigenvalues[0]!r}; iterations ours "
          f"{eng.getIterationCounts()} ref {ref.iter_counts}")
    assert_eigenpairs(wl, csr, dtype, evals, evecs, ref)
    if dtype == np.float64:
        assert abs(eng.getIterationCounts()[0] - ref.iter_counts[0]) <= 3


@pytest.mark.parametrize("nx", [32, 64])
def test_config2_laplacian_four_smallest_with_degenerate_pair(pkg, ctx, wl, checker, nx):
    csr = wl.laplacian2d_csr(nx)
    eng, evals, evecs, ref = run_both(pkg, ctx, checker, csr, np.float64, wl.start_vector(nx * nx), find_max=False, num_eigs=4)
    print(f"laplacian {nx}: iterations ours {eng.getIterationCounts()} ref {ref.iter_counts}")
    exact = wl.laplacian2d_exact(nx)
    assert np.allclose(evals, exact, rtol=1e-10)
    # (1,2)/(2,1) are exactly degenerate: compare that pair as a subspace
    assert_eigenpairs(wl, csr, np.float64, evals, evecs, ref, clusters=[[0], [1, 2], [3]], res_slack=1e3)


def test_config3_peierls_two_lowest_complex(pkg, ctx, wl, checker):
    csr = wl.peierls_csr(48, 48)
    n = csr[0].size - 1
    eng, evals, evecs, ref = run_both(pkg, ctx, checker, csr, np.complex128, wl.start_vector(n, np.complex128), find_max=False, num_eigs=2)
    print(f"peierls 48x48: evals {evals} iterations ours {eng.getIterationCounts()} ref {ref.iter_counts}")
    assert_eigenpairs(wl, csr, np.complex128, evals, evecs, ref, res_slack=1e3)


@pytest.mark.parametrize("L", [12, 16, 20])
def test_config4_xxz_ground_state_matrix_free(pkg, ctx, wl, checker, L):
    csr = wl.xxz_csr(L)
    n = csr[0].size - 1
    start = wl.start_vector(n)
    op = pkg.Operator.xxz(ctx, L)
    eng = pkg.LambdaLanczos(op, n, False, 1)
    eng.init_vector = start
    evals, evecs = eng.run()
    ref = checker.lanczos(*csr, find_max=False, num_eigs=1, init=start)
    print(f"xxz L={L}: E0 ours {evals[0]!r} ref {ref.eigenvalues[0]!r} iterations ours {eng.getIterationCounts()} ref {ref.iter_counts}")
    known = {16: -7.1422963606168, 20: -8.9043865298764}  # SURVEY.md §8d probe values
    if L in known:
        assert abs(evals[0] - known[L]) < 1e-10
    assert_eigenpairs(wl, csr, np.float64, evals, evecs, ref)


def test_golden_fixtures_of_the_compiled_reference(pkg, ctx, wl):
    import golden.make_golden as mg

    g = np.load(os.path.join(GOLDEN, "reference_runs.npz"))
    for name, (csr, kw, dt) in mg.cases(wl).items():
        n = csr[0].size - 1
        op = pkg.Operator.csr(ctx, *csr)
        eng = pkg.LambdaLanczos(op, n, kw["find_max"], kw["num_eigs"])
        eng.init_vector = wl.start_vector(n, dt)
        eng.eigenvalue_offset = kw.get("offset", 0.0)
        evals, evecs = eng.run()

        class Ref:
            eigenvalues = g[f"{name}/evals"]
            eigenvectors = g[f"{name}/evecs"]

        clusters = [[0], [1, 2], [3]] if name.startswith("laplacian") else None
        assert_eigenpairs(wl, csr, dt, evals, evecs, Ref, clusters=clusters, res_slack=1e3)
        print(name, "iterations ours", eng.getIterationCounts(), "reference", [int(x) for x in g[f"{name}/iters"]])
    for name, (csr, a, x, kw) in mg.expm_cases(wl).items():
        op = pkg.Operator.csr(ctx, *csr)
        ex = pkg.Exponentiator(op, op.n)
        ex.full_orthogonalize = kw.get("full_orth", False)
        if "max_iter" in kw:
            ex.max_iteration = kw["max_iter"]
        it, out = ex.run(a, x)
        ref_out = g[f"{name}/out"]
        assert it == int(g[f"{name}/iters"]), name
        assert np.linalg.norm(out - ref_out) <= 1e-10 * np.linalg.norm(ref_out), name


# ---- Exponentiator ------------------------------------------------------------------------------------------------
def test_exponentiate_real_3x3(pkg, ctx, wl, checker):  # exponentiator_test.cpp:31-81
    a = np.array([[2.0, 1, 1], [1, 2, 1], [1, 1, 2]])
    x = np.array([1.0, 0, 0])
    w, u = np.linalg.eigh(a)
    exact = u @ (np.exp(3.0 * w) * (u.T @ x))
    op = pkg.Operator.csr(ctx, *wl.dense_to_csr(a))
    ex = pkg.Exponentiator(op, 3)
    it, out = ex.run(3.0, x)
    ov = abs(np.vdot(exact, out)) / np.linalg.norm(exact) / np.linalg.norm(out)
    assert abs(1 - ov) <= ex.eps
    it_ref, out_ref = checker.expm(*wl.dense_to_csr(a), 3.0, x)
    assert it == it_ref == 3
    assert np.linalg.norm(out - out_ref) <= 1e-10 * np.linalg.norm(out_ref)
    it_t, out_t = ex.taylor_run(3.0, x)
    ov = abs(np.vdot(exact, out_t)) / np.linalg.norm(exact) / np.linalg.norm(out_t)
    assert abs(1 - ov) <= ex.eps
    assert it_t == checker.expm(*wl.dense_to_csr(a), 3.0, x, taylor=True)[0]


def test_exponentiate_ring_complex_and_zero_delta(pkg, ctx, wl, checker):  # exponentiator_test.cpp:106-222
    from test_oracle import ring, ring_input

    n = 100
    csr = wl.dense_to_csr(ring(n).astype(complex))
    x = ring_input(n)
    w, u = np.linalg.eigh(ring(n))
    exact = u @ (np.exp(3j * w) * (u.conj().T @ x))
    op = pkg.Operator.csr(ctx, *csr)
    ex = pkg.Exponentiator(op, n)
    it, out = ex.run(3j, x)
    it_ref, out_ref = checker.expm(*csr, 3j, x)
    assert it == it_ref == 19
    assert abs(1 - abs(np.vdot(exact, out)) / np.linalg.norm(exact) / np.linalg.norm(out)) <= ex.eps
    assert np.linalg.norm(out - out_ref) <= 1e-10 * np.linalg.norm(out_ref)
    it_t, out_t = ex.taylor_run(3j, x)
    assert it_t == 37 and np.linalg.norm(out_t - exact) <= 1e-12
    ex.full_orthogonalize = True
    it0, out0 = ex.run(0j, x)
    assert it0 == 2 and np.linalg.norm(out0 - x) <= 1e-14
    it0t, out0t = ex.taylor_run(0j, x)
    assert it0t == 1 and np.array_equal(out0t, x)


@pytest.mark.parametrize("L", [12, 16])
def test_config5_time_evolution_xxz_neel(pkg, ctx, wl, checker, L):
    """e^{-iH dt} on the Neel state, output fed back as input, matrix-free operator vs the checker's explicit CSR."""
    csr = wl.xxz_csr(L, dtype=np.complex128)
    op = pkg.Operator.xxz(ctx, L, dtype=np.complex128)
    ex = pkg.Exponentiator(op, op.n)
    x_gpu = wl.neel_state(L)
    x_ref = x_gpu.copy()
    steps = 10 if L <= 12 else 4
    for s in range(steps):
        it, x_gpu = ex.run(-0.1j, x_gpu)
        it_ref, x_ref = checker.expm(*csr, -0.1j, x_ref)
        assert it == it_ref, (s, it, it_ref)
        assert np.linalg.norm(x_gpu - x_ref) <= 1e-10 * np.linalg.norm(x_ref), s
        assert abs(np.linalg.norm(x_gpu) - 1.0) < 1e-12  # unitary evolution


# ---- engine behaviour at the edges ----------------------------------------------------------------------------------
def test_max_iteration_cap_and_speculation_do_not_change_results(pkg, ctx, wl):
    n = 50000
    csr = wl.random_symmetric_csr(n)
    op = pkg.Operator.csr(ctx, *csr)
    results = []
    for depth in (0, 1, 4):
        eng = pkg.LambdaLanczos(op, n, True, 2)
        eng.init_vector = wl.start_vector(n)
        eng.max_iteration = 60
        eng.pipeline_depth = depth
        evals, evecs = eng.run()
        results.append((evals, evecs, eng.getIterationCounts()))
    for r in results[1:]:
        assert r[2] == results[0][2]
        assert np.array_equal(r[0], results[0][0]) and np.array_equal(r[1], results[0][1])  # bit-for-bit reproducible
    assert all(c == 60 for c in results[0][2])


def test_ritz_solvers_agree_and_second_pass_is_harmless(pkg, ctx, wl):
    n = 20000
    csr = wl.random_symmetric_csr(n)
    op = pkg.Operator.csr(ctx, *csr)
    out = []
    for solver, orth in ((0, pkg.ORTH_FULL), (1, pkg.ORTH_FULL), (0, pkg.ORTH_FULL_TWICE)):
        eng = pkg.LambdaLanczos(op, n, False, 1)
        eng.init_vector = wl.start_vector(n)
        eng.ritz_solver = solver
        eng.orthogonalization = orth
        evals, evecs = eng.run()
        out.append((evals[0], evecs[0], eng.getIterationCounts()[0]))
    for lam, v, it in out[1:]:
        assert abs(lam - out[0][0]) <= 1e-12 * abs(out[0][0])
        assert abs(abs(np.vdot(v, out[0][1])) - 1) < 1e-10
        assert abs(it - out[0][2]) <= 2


def test_basis_capacity_error_is_reported(pkg, ctx, wl):
    n = 4000
    op = pkg.Operator.csr(ctx, *wl.random_symmetric_csr(n))
    kry = pkg.Krylov(ctx, np.float64, n, 4)
    kry.begin(wl.start_vector(n))
    for _ in range(3):
        kry.step(op)
    with pytest.raises(pkg.LlzError) as e:
        kry.step(op)
    assert e.value.status == 3  # LLZ_ERR_OOM: the store never silently drops a Lanczos vector


# ---- full-size properties (BASELINE.json config 2: n = 4096^2) ------------------------------------------------------
def test_config2_full_size_properties(pkg, ctx, wl):
    nx = 4096
    n = nx * nx
    csr = wl.laplacian2d_csr(nx)
    op = pkg.Operator.csr(ctx, *csr)
    steps = 24
    kry = pkg.Krylov(ctx, np.float64, n, steps + 1)
    start = wl.start_vector(n)
    nrm0 = kry.begin(start)
    assert abs(nrm0 - np.linalg.norm(start)) <= 1e-12 * nrm0
    alphas, betas = [], []
    for k in range(1, steps + 1):
        kry.step(op, 0.0, pkg.ORTH_FULL)
    for k in range(1, steps + 1):
        a, b = kry.fetch(k)
        alphas.append(a)
        betas.append(b)
    # (1) three-term recurrence holds column by column: A u_{k-1} = beta_{k-2} u_{k-2} + alpha_{k-1} u_{k-1} + beta_{k-1} u_k
    cols = {j: kry.column(j) for j in (0, 1, 2, steps - 2, steps - 1, steps)}
    for k in (2, steps):
        lhs = wl.csr_matvec(*csr, cols[k - 1])
        rhs = betas[k - 2] * cols[k - 2] + alphas[k - 1] * cols[k - 1] + betas[k - 1] * cols[k]
        assert np.linalg.norm(lhs - rhs) <= 1e-12 * np.linalg.norm(lhs)
    # (2) orthonormality of the sampled columns
    keys = sorted(cols)
    for i in keys:
        for j in keys:
            d = np.dot(cols[i], cols[j])
            assert abs(d - (1.0 if i == j else 0.0)) < 1e-12, (i, j, d)
    # (3) Ritz values stay inside the exact spectrum [lambda_min, lambda_max] of the Laplacian (interlacing)
    t = np.diag(alphas) + np.diag(betas[:-1], 1) + np.diag(betas[:-1], -1)
    ritz = np.linalg.eigvalsh(t)
    lo = 4 - 4 * math.cos(math.pi / (nx + 1))
    hi = 4 + 4 * math.cos(math.pi / (nx + 1))
    assert ritz[0] >= lo - 1e-12 and ritz[-1] <= hi + 1e-12
    # (4) linearity of the combine kernel: V (a y1 + b y2) = a V y1 + b V y2
    rs = np.random.RandomState(0)
    y1, y2 = rs.randn(steps), rs.randn(steps)
    o = kry.combine(np.array([y1, y2, 2.0 * y1 - 0.5 * y2]), normalize=False)
    v1, v2, v3 = (x.download() for x in o)
    assert np.linalg.norm(v3 - (2.0 * v1 - 0.5 * v2)) <= 1e-13 * np.linalg.norm(v3)
