"""CPU tests that PIN the oracle (oracle/llz_oracle.c, the plain-C restatement of the reference algorithm):

1. against the reference's own known-answer tests (literal matrices / closed-form spectra restated from
   /root/reference/test/lambda_lanczos_test.cpp and test/exponentiator_test.cpp — data, not code);
2. against the reference itself compiled here (oracle/_ref/libllz_ref.so) where it is present;
3. against fixtures in tests/golden/ generated from the compiled reference by tests/golden/make_golden.py.
"""
import math
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def fix_sign(v):
    return v * (np.conj(v[0]) / abs(v[0]))


# ---- unit level (lambda_lanczos_test.cpp:47-100) ----------------------------------------------------------------
def test_inner_product_conjugates_first_argument(port):
    v1 = np.array([3.0, 1 + 3j])
    v2 = np.array([3.0, 2 + 4j])
    assert port.inner_prod(v1, v2) == complex(23.0, -2.0)


def test_schmidt_orthogonalization(port):
    rs = np.random.RandomState(1)
    n = 10
    us = []
    for _ in range(n // 2):
        u = rs.uniform(-10, 10, n) + 1j * rs.uniform(-10, 10, n)
        if us:
            u = port.schmidt_orth(np.array(us), u)
        u = u / port.norm(u)
        us.append(u)
    v = rs.uniform(-10, 10, n) + 1j * rs.uniform(-10, 10, n)
    v = port.schmidt_orth(np.array(us), v)
    for u in us:
        ip = port.inner_prod(v, u)
        assert abs(ip.real) < 1e-15 * n * 10 and abs(ip.imag) < 1e-15 * n * 10


# ---- tridiagonal solver (lambda_lanczos_test.cpp:757-801) --------------------------------------------------------
def test_tridiagonal_implicit_shift_qr(port):
    ev, q, unc = port.tridiag(np.array([1.0, 2.0, 3.0]), np.array([2.0, 2.0]))
    assert np.allclose(ev, [-1, 2, 5], atol=1e-10)
    correct = np.array([[2, -2, 1], [2, 1, -2], [1, 2, 2]], dtype=float) / 3.0
    for i in range(3):
        s = np.sign(q[i][0]) * np.sign(correct[i][0])
        assert np.allclose(q[i], s * correct[i], atol=1e-10)
    assert unc == 0


def test_tridiagonal_null_eigenvalue_terminates(port):
    alpha = np.array([6.82333617e-03, 3.09398208e00, 1.89919458e00, 1.28531906e-16])
    beta = np.array([1.19582528e-01, -1.37689656e00, 6.16147405e-15])
    ev, q, _ = port.tridiag(alpha, beta)
    t = np.diag(alpha) + np.diag(beta, 1) + np.diag(beta, -1)
    assert np.allclose(np.sort(ev), np.linalg.eigvalsh(t), atol=1e-12)


# ---- eigen-engine known answers ------------------------------------------------------------------------------
def test_simple_matrix(port, wl):  # :128-161
    a = np.array([[2.0, 1, 1], [1, 2, 1], [1, 1, 2]])
    r = port.lanczos(*wl.dense_to_csr(a), find_max=True, num_eigs=1, offset=6.0, init=wl.start_vector(3))
    eps = np.finfo(float).eps * 1e3
    assert abs(r.eigenvalues[0] - 4.0) < 4.0 * eps
    assert np.allclose(fix_sign(r.eigenvectors[0]), np.ones(3) / math.sqrt(3), atol=4.0 * eps * 10)
    assert len(r.iter_counts) == 1


def test_simple_matrix_float(port, wl):  # :163-193
    a = np.array([[2.0, 1, 1], [1, 2, 1], [1, 1, 2]], dtype=np.float32)
    r = port.lanczos(*wl.dense_to_csr(a), find_max=True, num_eigs=1, offset=6.0, init=wl.start_vector(3, np.float32))
    eps = np.finfo(np.float32).eps * 1e3
    assert abs(r.eigenvalues[0] - 4.0) < 4.0 * eps


def test_dynamic_matrix(port, wl):  # :262-308
    n = 10
    a = np.zeros((n, n))
    for i in range(n - 1):
        a[i, i + 1] = a[i + 1, i] = -1.0
    r = port.lanczos(*wl.dense_to_csr(a), find_max=False, num_eigs=1, offset=-10.0, eps=1e-14, init=wl.start_vector(n))
    lam = -2.0 * math.cos(math.pi / (n + 1))
    vec = np.sin(np.arange(1, n + 1) * math.pi / (n + 1))
    vec /= np.linalg.norm(vec)
    assert abs(r.eigenvalues[0] - lam) < abs(lam) * 1e-14 * 4
    assert np.allclose(fix_sign(r.eigenvectors[0]), vec, atol=abs(lam) * 1e-13)


def test_hermitian_matrix(port, wl):  # :375-409
    h = np.array([[0, 1j, 1], [-1j, 0, 1j], [1, -1j, 0]])
    r = port.lanczos(*wl.dense_to_csr(h), find_max=False, num_eigs=1, init=wl.start_vector(3, np.complex128))
    eps = np.finfo(float).eps * 1e3
    assert abs(r.eigenvalues[0] + 2.0) < 2.0 * eps
    correct = np.array([1, 1j, -1]) / math.sqrt(3)
    assert np.allclose(fix_sign(r.eigenvectors[0]), correct, atol=2.0 * eps * 10)


def test_single_element_matrix(port, wl):  # :411-440
    r = port.lanczos(np.array([0, 1]), np.array([0]), np.array([2.0]), find_max=True, num_eigs=1, init=wl.start_vector(1))
    assert abs(r.eigenvalues[0] - 2.0) < 1e-12 and abs(abs(r.eigenvectors[0][0]) - 1.0) < 1e-12


EIGHT = np.array([[6, -3, -3, 0, -1, 1, -1, 1], [-3, -4, 2, 2, -1, -5, 0, -4], [-3, 2, 2, -3, 0, 0, -1, -1],
                  [0, 2, -3, 0, -3, 3, 2, 2], [-1, -1, 0, -3, -2, 0, -5, -4], [1, -5, 0, 3, 0, -4, 5, 0],
                  [-1, 0, -1, 2, -5, 5, -4, 4], [1, -4, -1, 2, -4, 0, 4, 2]], dtype=float)
EIGHT_VALS = [-13.21508597, -8.50033154, -4.26674892]
EIGHT_VECS = np.array([[0.02081752, -0.49222707, 0.13202088, 0.24048092, 0.15089223, -0.60850056, 0.48079787, -0.24043829],
                       [0.16645991, 0.51818471, -0.00646562, -0.09493495, 0.60595718, 0.02042567, 0.52346924, 0.23043415],
                       [0.03381669, -0.07999997, 0.32090331, 0.61650970, 0.41812886, -0.01782613, -0.45571810, 0.35575946]])


def test_multiple_eigenpairs(port, wl):  # :442-488
    r = port.lanczos(*wl.dense_to_csr(EIGHT), find_max=False, num_eigs=3, eps=1e-7, init=wl.start_vector(8))
    for i in range(3):
        assert abs(r.eigenvalues[i] - EIGHT_VALS[i]) < abs(EIGHT_VALS[i]) * 1e-7
        assert np.allclose(fix_sign(r.eigenvectors[i]), fix_sign(EIGHT_VECS[i]), atol=abs(EIGHT_VALS[i]) * 1e-6)


def test_multiple_degenerate_eigenpairs(port, wl):  # :490-536
    n, num = 50, 26
    a = np.zeros((n, n))
    for i in range(n):
        a[i, (i + 1) % n] = a[(i + 1) % n, i] = -1.0
    r = port.lanczos(*wl.dense_to_csr(a), find_max=False, num_eigs=num, eps=1e-14, init=wl.start_vector(n, seed=7))
    correct = np.sort(-2.0 * np.cos(2.0 * math.pi * np.arange(-num // 2, num - num // 2) / n))
    assert r.eigenvalues.size == num
    assert np.allclose(r.eigenvalues, correct, atol=1e-13)


def test_laplacian_exact_spectrum(port, wl):
    nx = 24
    r = port.lanczos(*wl.laplacian2d_csr(nx), find_max=False, num_eigs=4, init=wl.start_vector(nx * nx))
    assert np.allclose(r.eigenvalues, wl.laplacian2d_exact(nx), rtol=1e-10)


# ---- exponentiator known answers (exponentiator_test.cpp:31-222) ----------------------------------------------
def ring(n, t=-1.0):
    a = np.zeros((n, n))
    for i in range(n):
        a[i, (i + 1) % n] = a[(i + 1) % n, i] = t
    return a


def overlap(x, y):
    return abs(np.vdot(x, y)) / np.linalg.norm(x) / np.linalg.norm(y)


def test_exponentiate_real(port, wl):
    a = np.array([[2.0, 1, 1], [1, 2, 1], [1, 1, 2]])
    x = np.array([1.0, 0, 0])
    w, u = np.linalg.eigh(a)
    exact = u @ (np.exp(3.0 * w) * (u.T @ x))
    it, out = port.expm(*wl.dense_to_csr(a), 3.0, x)
    assert it == 3
    assert abs(1 - overlap(exact, out)) <= 2.3e-14
    it, out = port.expm(*wl.dense_to_csr(a), 3.0, x, taylor=True)
    assert abs(1 - overlap(exact, out)) <= 2.3e-14


def ring_input(n):
    x = np.zeros(n, complex)
    x[0] = 1 + 2j
    x[n - 1] = 1 + 2j
    x[n // 2] = 8 + 2j
    return x / np.linalg.norm(x)


def test_exponentiate_large_matrix(port, wl):
    n = 100
    a = ring(n)
    x = ring_input(n)
    w, u = np.linalg.eigh(a)
    exact = u @ (np.exp(3j * w) * (u.conj().T @ x))
    it, out = port.expm(*wl.dense_to_csr(a.astype(complex)), 3j, x)
    assert it == 19  # SURVEY.md §4 probe
    assert abs(1 - overlap(exact, out)) <= 2.3e-14
    it, out = port.expm(*wl.dense_to_csr(a.astype(complex)), 3j, x, taylor=True)
    assert it == 37
    assert abs(1 - overlap(exact, out)) <= 2.3e-14


def test_exponentiate_zero_delta(port, wl):
    n = 100
    x = ring_input(n)
    it, out = port.expm(*wl.dense_to_csr(ring(n).astype(complex)), 0j, x, full_orth=True)
    assert it == 2
    assert abs(1 - overlap(x, out)) <= 2.3e-14
    it, out = port.expm(*wl.dense_to_csr(ring(n).astype(complex)), 0j, x, taylor=True)
    assert it == 1 and np.array_equal(out, x)


# ---- restatement vs the compiled reference: bit-for-bit ---------------------------------------------------------
def _need_ref(oracle_mod):
    if not oracle_mod.have_reference():
        pytest.skip("oracle/_ref/libllz_ref.so not present (built only where /root/reference exists)")
    return oracle_mod.Reference()


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex64, np.complex128])
def test_restatement_matches_reference_units(oracle_mod, port, dtype):
    ref = _need_ref(oracle_mod)
    rs = np.random.RandomState(3)
    n = 1000

    def rnd():
        v = rs.uniform(-1, 1, n)
        if np.dtype(dtype).kind == "c":
            v = v + 1j * rs.uniform(-1, 1, n)
        return v.astype(dtype)

    a, b = rnd(), rnd()
    assert port.inner_prod(a, b) == ref.inner_prod(a, b)
    assert port.norm(a) == ref.norm(a)
    basis = np.array([rnd() for _ in range(6)])
    assert np.array_equal(port.schmidt_orth(basis, a), ref.schmidt_orth(basis, a))


def test_restatement_matches_reference_tridiagonal(oracle_mod, port):
    ref = _need_ref(oracle_mod)
    rs = np.random.RandomState(5)
    for m in (1, 2, 7, 60):
        alpha = rs.uniform(-2, 2, m)
        beta = rs.uniform(0.1, 1, max(m - 1, 0))
        e1, q1, _ = port.tridiag(alpha, beta)
        e2, q2, _ = ref.tridiag(alpha, beta)
        assert np.array_equal(e1, e2) and np.array_equal(q1, q2)


@pytest.mark.parametrize("case", ["random_sym", "laplacian", "peierls", "xxz", "random_sym_f32", "peierls_c64"])
def test_restatement_matches_reference_runs(oracle_mod, port, wl, case):
    ref = _need_ref(oracle_mod)
    if case == "random_sym":
        csr, kw, dt = wl.random_symmetric_csr(3000), dict(find_max=True, num_eigs=1), np.float64
    elif case == "random_sym_f32":
        csr, kw, dt = wl.random_symmetric_csr(2000, dtype=np.float32), dict(find_max=True, num_eigs=1), np.float32
    elif case == "laplacian":
        csr, kw, dt = wl.laplacian2d_csr(20), dict(find_max=False, num_eigs=4), np.float64
    elif case == "peierls":
        csr, kw, dt = wl.peierls_csr(16, 16), dict(find_max=False, num_eigs=2), np.complex128
    elif case == "peierls_c64":  # std::complex<float>: lambda_lanczos.hpp:109, util/common.hpp:80-102
        full = wl.peierls_csr(14, 12, flux=0.05, trap=0.3)
        csr, kw, dt = (full[0], full[1], full[2].astype(np.complex64)), dict(find_max=False, num_eigs=2), np.complex64
    else:
        csr, kw, dt = wl.xxz_csr(12), dict(find_max=False, num_eigs=1), np.float64
    n = csr[0].size - 1
    init = wl.start_vector(n, dt)
    r1 = port.lanczos(*csr, init=init, **kw)
    r2 = ref.lanczos(*csr, init=init, **kw)
    assert r1.iter_counts == r2.iter_counts
    assert np.array_equal(r1.eigenvalues, r2.eigenvalues)
    assert np.array_equal(r1.eigenvectors, r2.eigenvectors)


def test_restatement_matches_reference_expm(oracle_mod, port, wl):
    ref = _need_ref(oracle_mod)
    for dt in (np.complex128, np.complex64):
        csr = wl.xxz_csr(10, dtype=dt)
        x = wl.neel_state(10).astype(dt)
        for kw in (dict(), dict(full_orth=True), dict(taylor=True)):
            i1, o1 = port.expm(*csr, -0.1j, x, **kw)
            i2, o2 = ref.expm(*csr, -0.1j, x, **kw)
            assert i1 == i2 and np.array_equal(o1, o2), (dt, kw)


def test_run_iteration_spy_agrees_with_the_run_it_spies_on(oracle_mod, wl):
    """The shim's run_iteration spy (what bench.py's in-line parity block and the full-size GPU tests compare with): its
    alpha_k, beta_k and Ritz values are those of the reference's own run on the same input, and the captured Lanczos
    vectors are orthonormal and satisfy the three-term recurrence A u_k = beta_{k-1} u_{k-1} + alpha_k u_k + beta_k u_{k+1}."""
    ref = _need_ref(oracle_mod)
    csr = wl.laplacian2d_csr(24, 19)
    n = csr[0].size - 1
    start = wl.start_vector(n)
    m = 14
    r = ref.run_iteration(*csr, find_max=False, max_iter=m, init=start, capture=m)
    full = ref.lanczos(*csr, find_max=False, num_eigs=1, max_iter=m, init=start)
    assert full.iter_counts == [m] and r["iterations"] == m
    assert abs(r["eigenvalues"][0] - full.eigenvalues[0]) <= 1e-14 * abs(full.eigenvalues[0])
    V = np.array(r["basis"][:m])
    assert np.abs(V @ V.T - np.eye(m)).max() < 1e-13
    alpha, beta = r["alpha"][:m], r["beta"][:m - 1]
    for k in range(1, m - 1):
        lhs = wl.csr_matvec(*csr, V[k])
        rhs = beta[k - 1] * V[k - 1] + alpha[k] * V[k] + beta[k] * V[k + 1]
        assert np.linalg.norm(lhs - rhs) < 1e-12, k
    T = np.diag(alpha) + np.diag(beta, 1) + np.diag(beta, -1)
    assert abs(np.linalg.eigvalsh(T)[0] - r["eigenvalues"][0]) < 1e-13


# ---- restatement vs committed fixtures generated from the compiled reference -------------------------------------
def test_restatement_matches_golden_fixtures(port, wl):
    path = os.path.join(GOLDEN, "reference_runs.npz")
    if not os.path.exists(path):
        pytest.skip("golden fixtures not generated yet")
    g = np.load(path, allow_pickle=False)
    import golden.make_golden as mg  # the script that made them (tests/golden/make_golden.py)

    for name, (csr, kw, dt) in mg.cases(wl).items():
        n = csr[0].size - 1
        r = port.lanczos(*csr, init=wl.start_vector(n, dt), **kw)
        assert [int(x) for x in g[f"{name}/iters"]] == r.iter_counts, name
        assert np.array_equal(g[f"{name}/evals"], r.eigenvalues), name
        assert np.array_equal(g[f"{name}/evecs"], r.eigenvectors), name
    for name, (csr, a, x, kw) in mg.expm_cases(wl).items():
        it, out = port.expm(*csr, a, x, **kw)
        assert int(g[f"{name}/iters"]) == it, name
        assert np.array_equal(g[f"{name}/out"], out), name
