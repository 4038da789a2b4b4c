"""CPU tests of the host-side tridiagonal solvers (lambda_lanczos_b200/tridiagonal.hpp) — the part of the hot path that
stays on the host.  Checked against LAPACK (numpy/scipy) and against the oracle's restated implicit-shift QR."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def host():
    out_dir = os.path.join(HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libhost_shim.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"), "-I",
                           os.path.join(ROOT, "lambda-lanczos_b200", "include"), os.path.join(HERE, "host_shim.cpp"), "-o", so])
    lib = C.CDLL(so)
    lib.ht_ql.restype = C.c_int64
    lib.ht_sturm.restype = C.c_int64
    return lib


def p(a):
    return a.ctypes.data_as(C.c_void_p)


def lanczos_like(m, seed):
    """alpha/beta resembling a Lanczos recurrence of an operator with spectrum in [-2, 6]."""
    rs = np.random.RandomState(seed)
    return rs.uniform(-1, 5, m), rs.uniform(0.2, 1.5, m)


def dense(alpha, beta):
    m = alpha.size
    return np.diag(alpha) + np.diag(beta[: m - 1], 1) + np.diag(beta[: m - 1], -1)


@pytest.mark.parametrize("m", [1, 2, 3, 17, 200])
@pytest.mark.parametrize("find_max", [False, True])
def test_extreme_eigenvalues_match_lapack(host, m, find_max):
    alpha, beta = lanczos_like(m, m)
    exact = np.linalg.eigvalsh(dense(alpha, beta))
    nroot = min(5, m)
    out = np.zeros(nroot)
    host.ht_extreme(p(alpha), p(beta), C.c_int64(m), C.c_int64(nroot), C.c_int(int(find_max)), p(out))
    want = exact[::-1][:nroot] if find_max else exact[:nroot]
    assert np.allclose(out, want, rtol=0, atol=4e-15 * np.abs(exact).max())


@pytest.mark.parametrize("find_max", [False, True])
def test_extreme_sequence_warm_start_equals_cold(host, find_max):
    mmax, nroot = 120, 5
    alpha, beta = lanczos_like(mmax, 11)
    seq = np.zeros((mmax, nroot))
    host.ht_extreme_sequence(p(alpha), p(beta), C.c_int64(mmax), C.c_int64(nroot), C.c_int(int(find_max)), p(seq))
    for m in (1, 2, 5, 6, 50, 120):
        exact = np.linalg.eigvalsh(dense(alpha[:m], beta[:m]))
        k = min(nroot, m)
        want = exact[::-1][:k] if find_max else exact[:k]
        assert np.allclose(seq[m - 1, :k], want, rtol=0, atol=4e-15 * np.abs(exact).max()), m


def real_lanczos_recurrence(kind, m, seed):
    """alpha/beta of an actual (fully reorthogonalised) Lanczos run: converging extreme Ritz values, the regime the
    predicted brackets are built for — including a deflated operator with (near-)degenerate values and a breakdown."""
    rs = np.random.RandomState(seed)
    if kind == "clustered":  # spectrum with tight clusters at both ends
        n = 400
        lam = np.concatenate([[-3.0, -3.0 + 1e-9, -3.0 + 2e-9, -2.5], rs.uniform(-1, 1, n - 8), [2.5, 3.0 - 2e-10, 3.0 - 1e-10, 3.0]])
    elif kind == "lowrank":  # Krylov space exhausts after 12 steps: beta collapses to rounding level
        n = 300
        lam = np.repeat(np.linspace(-1, 2, 12), n // 12)
    else:  # 1-D Laplacian-like: many slowly converging values
        n = 600
        lam = 2 - 2 * np.cos(np.arange(1, n + 1) * np.pi / (n + 1))
    n = lam.size
    v = rs.uniform(-1, 1, n)
    v /= np.linalg.norm(v)
    V = [v]
    alpha, beta = [], []
    for k in range(m):
        w = lam * V[-1]
        alpha.append(float(w @ V[-1]))
        B = np.array(V)
        w = w - B.T @ (B @ w)
        w = w - B.T @ (B @ w)
        b = float(np.linalg.norm(w))
        beta.append(b)
        if b < 1e-13:
            break
        V.append(w / b)
    return np.array(alpha), np.array(beta)


@pytest.mark.parametrize("kind,m", [("clustered", 160), ("lowrank", 40), ("laplacian", 250)])
@pytest.mark.parametrize("find_max", [False, True])
@pytest.mark.parametrize("nroot", [1, 5, 8])
def test_predicted_brackets_never_lose_a_root(host, kind, m, find_max, nroot):
    """The warm-started solver (brackets predicted from the previous iteration, checked by Sturm counts) must return
    what a cold bisection of every prefix T_1..T_m returns, on recurrences of real Lanczos runs."""
    alpha, beta = real_lanczos_recurrence(kind, m, 3)
    mm = alpha.size
    seq = np.zeros((mm, nroot))
    host.ht_extreme_sequence(p(alpha), p(beta), C.c_int64(mm), C.c_int64(nroot), C.c_int(int(find_max)), p(seq))
    for k in range(1, mm + 1):
        cold = np.zeros(min(nroot, k))
        host.ht_extreme(p(alpha), p(beta), C.c_int64(k), C.c_int64(min(nroot, k)), C.c_int(int(find_max)), p(cold))
        scale = max(1.0, np.abs(alpha[:k]).max() + 2 * np.abs(beta[:k]).max())
        assert np.allclose(seq[k - 1, : cold.size], cold, rtol=0, atol=8e-16 * scale), (kind, k)
    exact = np.linalg.eigvalsh(dense(alpha, beta))
    want = exact[::-1][:nroot] if find_max else exact[:nroot]
    assert np.allclose(seq[mm - 1, : want.size], want[: min(nroot, mm)], rtol=0, atol=1e-13 * np.abs(exact).max())


def test_extreme_eigenvalues_clustered_and_tiny_couplings(host):
    # nearly decoupled blocks => nearly degenerate extreme values; Sturm counts must still separate them
    alpha = np.array([1.0, 1.0, 1.0 + 1e-13, 3.0, 3.0, -2.0])
    beta = np.array([1e-9, 1e-12, 0.5, 1e-15, 0.25, 0.0])
    exact = np.linalg.eigvalsh(dense(alpha, beta))
    out = np.zeros(5)
    host.ht_extreme(p(alpha), p(beta), C.c_int64(6), C.c_int64(5), C.c_int(0), p(out))
    assert np.allclose(out, exact[:5], rtol=0, atol=1e-14)


@pytest.mark.parametrize("m", [1, 2, 3, 40, 150])
def test_implicit_ql_decomposition(host, m, port):
    alpha, beta = lanczos_like(m, 100 + m)
    vals = np.zeros(m)
    vecs = np.zeros((m, m))
    fails = host.ht_ql(p(alpha), p(beta), C.c_int64(m), p(vals), p(vecs))
    assert fails == 0
    t = dense(alpha, beta)
    assert np.allclose(vals, np.linalg.eigvalsh(t), rtol=0, atol=1e-13 * max(1.0, np.abs(t).max()))
    assert np.allclose(vecs @ vecs.T, np.eye(m), atol=1e-12)          # rows are orthonormal eigenvectors
    assert np.allclose(vecs @ t @ vecs.T, np.diag(vals), atol=1e-12)  # and diagonalise T
    # same answer as the oracle's restatement of the reference's implicit-shift QR
    ev, _, _ = port.tridiag(alpha, beta[: max(m - 1, 0)])
    assert np.allclose(vals, ev, rtol=0, atol=1e-13 * max(1.0, np.abs(t).max()))


def test_reference_tridiagonal_known_answer(host):  # lambda_lanczos_test.cpp:757-784
    alpha, beta = np.array([1.0, 2.0, 3.0]), np.array([2.0, 2.0, 0.0])
    vals, vecs = np.zeros(3), np.zeros((3, 3))
    host.ht_ql(p(alpha), p(beta), C.c_int64(3), p(vals), p(vecs))
    assert np.allclose(vals, [-1, 2, 5], atol=1e-10)
    correct = np.array([[2, -2, 1], [2, 1, -2], [1, 2, 2]], float) / 3
    for i in range(3):
        assert np.allclose(vecs[i] * np.sign(vecs[i][0]), correct[i] * np.sign(correct[i][0]), atol=1e-10)


@pytest.mark.parametrize("m", [2, 3, 60, 300])
@pytest.mark.parametrize("find_max", [False, True])
def test_eigenvectors_for_extreme_values(host, m, find_max):
    alpha, beta = lanczos_like(m, 7 * m)
    beta[m - 1] = 0.0
    t = dense(alpha, beta)
    w, v = np.linalg.eigh(t)
    k = min(5, m)
    idx = (np.arange(m)[::-1] if find_max else np.arange(m))[:k]
    lam = w[idx].copy()
    out = np.zeros((k, m))
    host.ht_eigvecs(p(alpha), p(beta), C.c_int64(m), p(lam), C.c_int64(k), p(out))
    for r in range(k):
        assert abs(np.linalg.norm(out[r]) - 1) < 1e-13
        assert abs(abs(out[r] @ v[:, idx[r]]) - 1) < 1e-10, (m, r)
        assert np.linalg.norm(t @ out[r] - lam[r] * out[r]) < 1e-12 * max(1.0, np.abs(t).max())


def test_eigenvectors_for_cluster_are_orthonormal(host):
    # two weakly coupled identical blocks: the two lowest eigenvalues differ by ~1e-12
    blk_a, blk_b = np.array([0.5, 1.0, 2.0, 3.0]), np.array([0.7, 0.6, 0.9])
    alpha = np.concatenate([blk_a, blk_a])
    beta = np.concatenate([blk_b, [1e-12], blk_b, [0.0]])
    m = alpha.size
    t = dense(alpha, beta)
    w = np.linalg.eigvalsh(t)
    lam = w[:4].copy()
    out = np.zeros((4, m))
    host.ht_eigvecs(p(alpha), p(beta), C.c_int64(m), p(lam), C.c_int64(4), p(out))
    assert np.allclose(out @ out.T, np.eye(4), atol=1e-9)
    for r in range(4):
        assert np.linalg.norm(t @ out[r] - lam[r] * out[r]) < 1e-10
