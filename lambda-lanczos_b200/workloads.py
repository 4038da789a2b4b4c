"""Synthetic operators of the shapes BASELINE.json names, built with numpy so that the CUDA path, the
oracle and the compiled reference all receive bit-identical arrays.  Pure host-side data generation: no compute path.

Every builder returns ``(rowptr int64, colidx int32, vals)`` in CSR with sorted column indices.
"""
from __future__ import annotations

import math

import numpy as np


def start_vector(n: int, dtype=np.float64, seed: int = 1) -> np.ndarray:
    """Seeded start vector, uniform in [-1, 1] per component (real and imaginary parts for complex types).

    Plays the role of the reference tests' ``vector_initializer`` (test/lambda_lanczos_test.cpp:25-45): the same
    vector is handed out at the start of every Lanczos run.
    """
    rs = np.random.RandomState(seed)
    dt = np.dtype(dtype)
    if dt.kind == "c":
        raw = rs.uniform(-1.0, 1.0, size=2 * n)
        return (raw[0::2] + 1j * raw[1::2]).astype(dt)
    return rs.uniform(-1.0, 1.0, size=n).astype(dt)


def random_symmetric_csr(n: int, offdiag_per_row: int = 8, seed: int = 12345, dtype=np.float64):
    """Config 1: random symmetric sparse matrix, a diagonal plus ``offdiag_per_row`` mirrored entries per row
    (~2*offdiag_per_row+1 non-zeros per row), values uniform in [-1, 1]."""
    import scipy.sparse as sp

    rs = np.random.RandomState(seed)
    rows = np.repeat(np.arange(n, dtype=np.int64), offdiag_per_row)
    cols = rs.randint(0, n - 1, size=rows.size).astype(np.int64)
    cols = cols + (cols >= rows)  # j != i
    v = rs.uniform(-1.0, 1.0, size=rows.size)
    d = rs.uniform(-1.0, 1.0, size=n)
    ii = np.concatenate([rows, cols, np.arange(n)])
    jj = np.concatenate([cols, rows, np.arange(n)])
    vv = np.concatenate([v, v, d])
    a = sp.coo_matrix((vv, (ii, jj)), shape=(n, n)).tocsr()
    a.sum_duplicates()
    a.sort_indices()
    return a.indptr.astype(np.int64), a.indices.astype(np.int32), a.data.astype(dtype)


def laplacian2d_csr(nx: int, ny: int | None = None, dtype=np.float64):
    """Config 2: 5-point Laplacian 4*I - shifts on an nx x ny grid, Dirichlet boundaries.  Exact eigenvalues
    4 - 2cos(p*pi/(nx+1)) - 2cos(q*pi/(ny+1)); for nx == ny the (1,2)/(2,1) pair is degenerate."""
    ny = nx if ny is None else ny
    n = nx * ny
    idx = np.arange(n, dtype=np.int64)
    x = idx % nx
    y = idx // nx
    # candidates in increasing column order: (y-1), (x-1), diag, (x+1), (y+1)
    cand_col = np.stack([idx - nx, idx - 1, idx, idx + 1, idx + nx], axis=1)
    cand_ok = np.stack([y > 0, x > 0, np.ones(n, bool), x < nx - 1, y < ny - 1], axis=1)
    cand_val = np.array([-1.0, -1.0, 4.0, -1.0, -1.0])
    counts = cand_ok.sum(axis=1)
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(counts, out=rowptr[1:])
    colidx = cand_col[cand_ok].astype(np.int32)
    vals = np.broadcast_to(cand_val, (n, 5))[cand_ok].astype(dtype)
    return rowptr, colidx, vals


def partition(n: int, rank: int, nranks: int):
    """Row block of ``rank`` in a group of ``nranks``: (row0, n_local) with interior boundaries floor(n*r/nranks)
    rounded down to a multiple of 4 — the same arithmetic as ``llz_partition`` (csrc/llz_halo.cpp)."""
    def bound(r):
        return 0 if r <= 0 else n if r >= nranks else (n * r // nranks) & ~3

    a, b = bound(rank), bound(rank + 1)
    return a, b - a


def csr_row_block(rowptr, colidx, vals, row0: int, n_rows: int):
    """Rows [row0, row0 + n_rows) of a CSR matrix: local row pointers (starting at 0), GLOBAL column indices."""
    lo, hi = int(rowptr[row0]), int(rowptr[row0 + n_rows])
    return (rowptr[row0:row0 + n_rows + 1] - lo).astype(np.int64), colidx[lo:hi], vals[lo:hi]


def laplacian2d_csr_rows(nx: int, row0: int, n_rows: int, ny: int | None = None, dtype=np.float64):
    """Rows [row0, row0 + n_rows) of ``laplacian2d_csr(nx, ny)`` built directly (a rank of a row-sharded run never
    materialises the whole matrix); bit-identical to slicing the full matrix."""
    ny = nx if ny is None else ny
    idx = np.arange(row0, row0 + n_rows, dtype=np.int64)
    x = idx % nx
    y = idx // nx
    cand_col = np.stack([idx - nx, idx - 1, idx, idx + 1, idx + nx], axis=1)
    cand_ok = np.stack([y > 0, x > 0, np.ones(n_rows, bool), x < nx - 1, y < ny - 1], axis=1)
    cand_val = np.array([-1.0, -1.0, 4.0, -1.0, -1.0])
    counts = cand_ok.sum(axis=1)
    rowptr = np.zeros(n_rows + 1, dtype=np.int64)
    np.cumsum(counts, out=rowptr[1:])
    return rowptr, cand_col[cand_ok].astype(np.int32), np.broadcast_to(cand_val, (n_rows, 5))[cand_ok].astype(dtype)


def laplacian2d_exact(nx: int, ny: int | None = None, count: int = 4) -> np.ndarray:
    ny = nx if ny is None else ny
    p = np.arange(1, nx + 1)
    q = np.arange(1, ny + 1)
    lam = (4 - 2 * np.cos(p * math.pi / (nx + 1)))[:, None] - 2 * np.cos(q * math.pi / (ny + 1))[None, :]
    return np.sort(lam.ravel())[:count]


def peierls_csr(lx: int, ly: int, flux: float = 0.05, t: float = 1.0, trap: float = 0.02, dtype=np.complex128):
    """Config 3: complex Hermitian tight-binding square lattice, Landau gauge.  Hopping -t along x,
    -t*exp(+-i 2 pi flux x) along y, open boundaries, plus a harmonic on-site trap
    ``trap * ((x-cx)^2 + (y-cy)^2) / max(lx,ly)`` that lifts the Landau-level degeneracy so the two lowest levels
    are separated."""
    n = lx * ly
    idx = np.arange(n, dtype=np.int64)
    x = idx % lx
    y = idx // lx
    cx, cy = (lx - 1) / 2.0, (ly - 1) / 2.0
    onsite = trap * ((x - cx) ** 2 + (y - cy) ** 2) / max(lx, ly)
    phase = np.exp(2j * math.pi * flux * x)  # hop from (x,y) to (x,y+1) carries -t*phase(x); reverse carries conj
    cand_col = np.stack([idx - lx, idx - 1, idx, idx + 1, idx + lx], axis=1)
    cand_ok = np.stack([y > 0, x > 0, np.ones(n, bool), x < lx - 1, y < ly - 1], axis=1)
    cand_val = np.stack([-t * phase, np.full(n, -t, complex), onsite.astype(complex), np.full(n, -t, complex),
                         -t * np.conj(phase)], axis=1)
    # H[i, i+lx] = -t*conj(phase(x)), H[i, i-lx] = -t*phase(x)  =>  H[i+lx, i] = conj(H[i, i+lx]): Hermitian
    counts = cand_ok.sum(axis=1)
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(counts, out=rowptr[1:])
    return rowptr, cand_col[cand_ok].astype(np.int32), cand_val[cand_ok].astype(dtype)


def xxz_states(L: int, n_up: int) -> np.ndarray:
    """All L-bit integers with n_up bits set, in increasing order (the Sz sector basis of configs 4/5)."""
    from itertools import combinations

    if L <= 24:
        s = np.fromiter((sum(1 << b for b in c) for c in combinations(range(L), n_up)), dtype=np.int64)
        s.sort()
        return s
    out = np.arange(1 << L, dtype=np.int64)
    pop = np.zeros(out.size, dtype=np.int8)
    for b in range(L):
        pop += ((out >> b) & 1).astype(np.int8)
    return out[pop == n_up]


def xxz_csr(L: int, n_up: int | None = None, jz: float = 1.0, jxy: float = 1.0, pbc: bool = True, dtype=np.float64):
    """Configs 4/5 at oracle-feasible sizes: H = sum_b Jxy/2 (S+S- + h.c.) + Jz SzSz on a chain of L spins-1/2,
    restricted to the sector with n_up up spins (default L/2), states in increasing integer order."""
    n_up = L // 2 if n_up is None else n_up
    states = xxz_states(L, n_up)
    n = states.size
    bonds = [(i, (i + 1) % L) for i in range(L if pbc else L - 1)]
    diag = np.zeros(n)
    rows, cols, vals = [], [], []
    for (i, j) in bonds:
        bi = (states >> i) & 1
        bj = (states >> j) & 1
        diag += np.where(bi == bj, 0.25 * jz, -0.25 * jz)
        flip = bi != bj
        src = np.nonzero(flip)[0]
        tgt_state = states[src] ^ ((1 << i) | (1 << j))
        tgt = np.searchsorted(states, tgt_state)
        rows.append(src)
        cols.append(tgt)
        vals.append(np.full(src.size, 0.5 * jxy))
    import scipy.sparse as sp

    rows.append(np.arange(n))
    cols.append(np.arange(n))
    vals.append(diag)
    a = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n)).tocsr()
    a.sum_duplicates()
    a.sort_indices()
    return a.indptr.astype(np.int64), a.indices.astype(np.int32), a.data.astype(dtype)


def neel_state(L: int, n_up: int | None = None, dtype=np.complex128) -> np.ndarray:
    """|0101...> as a unit vector in the Sz-sector basis (start state of config 5)."""
    n_up = L // 2 if n_up is None else n_up
    states = xxz_states(L, n_up)
    neel = sum(1 << b for b in range(0, L, 2))
    v = np.zeros(states.size, dtype=dtype)
    v[np.searchsorted(states, neel)] = 1
    return v


def dense_to_csr(a: np.ndarray):
    """Literal dense test matrices (the reference's 3x3 / 8x8 cases) as CSR, keeping explicit zeros out."""
    a = np.asarray(a)
    n = a.shape[0]
    mask = a != 0
    counts = mask.sum(axis=1)
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(counts, out=rowptr[1:])
    cols = np.nonzero(mask)[1].astype(np.int32)
    return rowptr, cols, a[mask]


def csr_matvec(rowptr, colidx, vals, x):
    """Host numpy SpMV used by tests to compute residuals ||Av - lambda v|| (never on the product path)."""
    import scipy.sparse as sp

    n = rowptr.size - 1
    return sp.csr_matrix((vals, colidx, rowptr), shape=(n, n)) @ x
