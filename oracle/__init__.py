"""oracle — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes bindings for the two CPU checkers of the lambda-lanczos hot path:

* ``libllz_oracle.so`` (``oracle/llz_oracle.c``): this repo's plain-C restatement of the reference algorithm;
* ``_ref/libllz_ref.so`` (``oracle/ref_shim.cpp``): the UNMODIFIED reference compiled from ``/root/reference``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import
this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "libllz_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libllz_ref.so")            # strict IEEE build: the parity checker
REF_FAST_SO = os.path.join(_HERE, "_ref", "libllz_ref_fast.so")  # -O3 -march=x86-64-v3 build: timing legs of bench.py only

_SFX = {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64", np.dtype(np.complex128): "c128",
        np.dtype(np.complex64): "c64"}
_REAL = {"f32": np.float32, "f64": np.float64, "c128": np.float64, "c64": np.float32}
_ALL_SFX = ("f32", "f64", "c128", "c64")

i64 = C.c_int64
vp = C.c_void_p


def build(force: bool = False) -> None:
    """Compile the checkers (``make -C oracle``).  The reference library is only rebuilt where /root/reference exists."""
    if force or not os.path.exists(ORACLE_SO) or (os.path.isdir("/root/reference") and not os.path.exists(REF_SO)):
        subprocess.check_call(["make", "-C", _HERE, "--no-print-directory"] + (["-B"] if force else []))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(vp)


def _csr(rowptr, colidx, vals, dtype):
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
    colidx = np.ascontiguousarray(colidx, dtype=np.int32)
    vals = np.ascontiguousarray(vals, dtype=dtype)
    return rowptr, colidx, vals


@dataclass
class LanczosResult:
    eigenvalues: np.ndarray
    eigenvectors: np.ndarray  # (n_found, n)
    iter_counts: list
    seconds: float = 0.0
    mv_seconds: float = 0.0
    basis: np.ndarray | None = None  # Lanczos vectors of the first run, as handed to mv_mul
    extra: dict = field(default_factory=dict)


class _Lib:
    def __init__(self, path):
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle` (or __graft_entry__.build())")
        self.lib = C.CDLL(path)


class Restatement(_Lib):
    """The plain-C restatement (kind = "port")."""

    kind = "port"

    def __init__(self):
        super().__init__(ORACLE_SO)
        L = self.lib
        for s in _ALL_SFX:
            getattr(L, f"llzo_lanczos_run_{s}").restype = C.c_int
            getattr(L, f"llzo_expm_run_{s}").restype = i64
            getattr(L, f"llzo_expm_taylor_{s}").restype = i64
            getattr(L, f"llzo_run_iteration_{s}").restype = i64
        L.llzo_norm_f64.restype = C.c_double
        L.llzo_norm_c128.restype = C.c_double
        L.llzo_norm_f32.restype = C.c_float
        L.llzo_norm_c64.restype = C.c_float
        for s in ("f32", "f64"):
            getattr(L, f"llzo_tridiag_eigenpairs_{s}").restype = i64
            getattr(L, f"llzo_tridiag_eigenvalues_{s}").restype = i64

    @staticmethod
    def _scalar(sfx, x):
        if sfx == "f32":
            return C.c_float(float(np.real(x)))
        if sfx == "f64":
            return C.c_double(float(np.real(x)))

        # System V x86-64 passes double _Complex like two doubles in SSE registers, float _Complex like a struct of
        # two floats (one SSE eightbyte)
        ft = C.c_float if sfx == "c64" else C.c_double

        class _Cplx(C.Structure):
            _fields_ = [("re", ft), ("im", ft)]

        z = complex(x)
        return _Cplx(z.real, z.imag)

    def lanczos(self, rowptr, colidx, vals, *, find_max, num_eigs=1, offset=0.0, eps=-1.0, max_iter=0, nepi=0,
                init, max_runs=64):
        dt = np.dtype(vals.dtype)
        sfx = _SFX[dt]
        rt = _REAL[sfx]
        rowptr, colidx, vals = _csr(rowptr, colidx, vals, dt)
        n = rowptr.size - 1
        init = np.ascontiguousarray(init, dtype=dt)
        evals = np.zeros(num_eigs, dtype=rt)
        evecs = np.zeros((num_eigs, n), dtype=dt)
        iters = np.zeros(max_runs, dtype=np.int64)
        n_runs, n_found = i64(0), i64(0)
        real_c = C.c_float if sfx in ("f32", "c64") else C.c_double
        rc = getattr(self.lib, f"llzo_lanczos_run_{sfx}")(
            i64(n), _ptr(rowptr), _ptr(colidx), _ptr(vals), C.c_int(int(find_max)), i64(num_eigs), real_c(offset),
            real_c(eps), i64(max_iter), i64(nepi), _ptr(init), _ptr(evals), _ptr(evecs), _ptr(iters), i64(max_runs),
            C.byref(n_runs), C.byref(n_found))
        assert rc == 0
        nf = n_found.value
        return LanczosResult(evals[:nf].copy(), evecs[:nf].copy(), [int(x) for x in iters[: n_runs.value]])

    def run_iteration(self, rowptr, colidx, vals, *, find_max, offset=0.0, eps=-1.0, max_iter=0, nroot=5, init,
                      locked=None):
        """One Lanczos run; also returns the alpha/beta actually used (beta[-1] forced to 0 as the reference does)."""
        dt = np.dtype(vals.dtype)
        sfx = _SFX[dt]
        rt = _REAL[sfx]
        rowptr, colidx, vals = _csr(rowptr, colidx, vals, dt)
        n = rowptr.size - 1
        if eps <= 0:
            eps = float(np.finfo(rt).eps) * 1e3
        if max_iter <= 0:
            max_iter = n
        init = np.ascontiguousarray(init, dtype=dt)
        locked = np.zeros((0, n), dtype=dt) if locked is None else np.ascontiguousarray(locked, dtype=dt)
        nl = locked.shape[0]
        ptrs = (vp * max(nl, 1))(*[locked[i].ctypes.data for i in range(nl)])
        evals = np.zeros(nroot, dtype=rt)
        evecs = np.zeros((nroot, n), dtype=dt)
        alpha = np.zeros(max_iter + 1, dtype=rt)
        beta = np.zeros(max_iter + 1, dtype=rt)
        nvals = i64(0)
        real_c = C.c_float if sfx in ("f32", "c64") else C.c_double
        it = getattr(self.lib, f"llzo_run_iteration_{sfx}")(
            i64(n), _ptr(rowptr), _ptr(colidx), _ptr(vals), C.c_int(int(find_max)), real_c(offset), real_c(eps),
            i64(max_iter), i64(nroot), _ptr(init), i64(nl), ptrs, _ptr(evals), _ptr(evecs), C.byref(nvals),
            _ptr(alpha), _ptr(beta))
        nv = nvals.value
        return int(it), evals[:nv].copy(), evecs[:nv].copy(), alpha[:it].copy(), beta[:it].copy()

    def expm(self, rowptr, colidx, vals, a, x, *, eps=-1.0, full_orth=False, max_iter=0, taylor=False):
        dt = np.dtype(vals.dtype)
        sfx = _SFX[dt]
        rowptr, colidx, vals = _csr(rowptr, colidx, vals, dt)
        n = rowptr.size - 1
        x = np.ascontiguousarray(x, dtype=dt)
        out = np.zeros(n, dtype=dt)
        real_c = C.c_float if sfx in ("f32", "c64") else C.c_double
        if taylor:
            it = getattr(self.lib, f"llzo_expm_taylor_{sfx}")(
                i64(n), _ptr(rowptr), _ptr(colidx), _ptr(vals), self._scalar(sfx, a), _ptr(x), _ptr(out), real_c(eps))
        else:
            it = getattr(self.lib, f"llzo_expm_run_{sfx}")(
                i64(n), _ptr(rowptr), _ptr(colidx), _ptr(vals), self._scalar(sfx, a), _ptr(x), _ptr(out), real_c(eps),
                C.c_int(int(full_orth)), i64(max_iter))
        return int(it), out

    def inner_prod(self, a, b):
        dt = np.dtype(a.dtype)
        sfx = _SFX[dt]
        a = np.ascontiguousarray(a)
        b = np.ascontiguousarray(b, dtype=dt)
        f = getattr(self.lib, f"llzo_inner_prod_{sfx}")
        if sfx in ("c128", "c64"):
            ft = C.c_float if sfx == "c64" else C.c_double

            class _Cplx(C.Structure):
                _fields_ = [("re", ft), ("im", ft)]
            f.restype = _Cplx
            r = f(i64(a.size), _ptr(a), _ptr(b))
            return complex(r.re, r.im)
        f.restype = C.c_float if sfx == "f32" else C.c_double
        return f(i64(a.size), _ptr(a), _ptr(b))

    def norm(self, a):
        sfx = _SFX[np.dtype(a.dtype)]
        a = np.ascontiguousarray(a)
        return float(getattr(self.lib, f"llzo_norm_{sfx}")(i64(a.size), _ptr(a)))

    def schmidt_orth(self, basis, w):
        dt = np.dtype(w.dtype)
        sfx = _SFX[dt]
        basis = np.ascontiguousarray(basis, dtype=dt)
        w = np.array(w, dtype=dt, copy=True)
        getattr(self.lib, f"llzo_schmidt_orth_{sfx}")(i64(w.size), i64(basis.shape[0]), _ptr(basis), _ptr(w))
        return w

    def tridiag(self, alpha, beta, vectors=True):
        alpha = np.ascontiguousarray(alpha)
        dt = np.dtype(alpha.dtype)
        sfx = _SFX[dt]
        m = alpha.size
        beta = np.ascontiguousarray(np.concatenate([np.asarray(beta, dtype=dt), np.zeros(1, dtype=dt)]))
        ev = np.zeros(m, dtype=dt)
        if vectors:
            q = np.zeros((m, m), dtype=dt)
            unc = getattr(self.lib, f"llzo_tridiag_eigenpairs_{sfx}")(i64(m), _ptr(alpha), _ptr(beta), _ptr(ev), _ptr(q))
            return ev, q, int(unc)
        unc = getattr(self.lib, f"llzo_tridiag_eigenvalues_{sfx}")(i64(m), _ptr(alpha), _ptr(beta), _ptr(ev))
        return ev, None, int(unc)


class Reference(_Lib):
    """The unmodified reference compiled here from /root/reference (kind = "reference")."""

    kind = "reference"

    def __init__(self, fast: bool = False):
        """fast=True loads the -O3 build (timing only; parity always uses the strict IEEE build)."""
        self.fast = bool(fast) and os.path.exists(REF_FAST_SO)
        super().__init__(REF_FAST_SO if self.fast else REF_SO)
        self.build_flags = ("g++ -O3 -march=x86-64-v3 -fopenmp" if self.fast else "g++ -O2 -ffp-contract=off -fopenmp")
        L = self.lib
        for s in _ALL_SFX:
            getattr(L, f"ref_lanczos_run_{s}").restype = C.c_int
            getattr(L, f"ref_run_iteration_{s}").restype = i64
            getattr(L, f"ref_expm_run_{s}").restype = i64
            getattr(L, f"ref_norm_{s}").restype = C.c_double
        L.ref_tridiag_eigenpairs_f64.restype = i64
        L.ref_tridiag_eigenpairs_f32.restype = i64
        L.ref_host_threads.restype = C.c_int

    def host_threads(self):
        """Host cores this process may use for the mv_mul lambda.  (Not omp_get_max_threads(): torch.distributed.run
        exports OMP_NUM_THREADS=1, which must not silently shrink the CPU arm; the shim passes the count explicitly
        through an OpenMP num_threads clause.)"""
        try:
            return max(1, len(os.sched_getaffinity(0)))
        except Exception:
            return max(1, os.cpu_count() or 1)

    def run_iteration(self, rowptr, colidx, vals, *, find_max, offset=0.0, eps=-1.0, max_iter=0, nroot=5, init,
                      locked=None, locked_blocks=None, n_locked=0, mv_threads=1, capture=0, want_vectors=False,
                      want_beta=True):
        """One Lanczos run through the reference's public ``run_iteration`` (lambda_lanczos.hpp:216-322) with a spy on
        ``mv_mul``: returns a dict with iterations, eigenvalues, alpha[k], beta[k] (k < mv calls - 1), per-iteration
        wall seconds ``dt_iter`` (between consecutive mv_mul calls), the captured Lanczos vectors ``basis`` and the
        total seconds.  ``locked``: (q, n) array of vectors to deflate against; or ``locked_blocks`` + ``n_locked``:
        vector j = locked_blocks restricted to rows [floor(j n/q), floor((j+1) n/q))."""
        dt = np.dtype(vals.dtype)
        sfx = _SFX[dt]
        rt = _REAL[sfx]
        rowptr, colidx, vals = _csr(rowptr, colidx, vals, dt)
        n = rowptr.size - 1
        if max_iter <= 0:
            max_iter = n
        init = np.ascontiguousarray(init, dtype=dt)
        ptrs = None
        blocks = None
        if locked is not None:
            locked = np.ascontiguousarray(locked, dtype=dt)
            n_locked = locked.shape[0]
            ptrs = (vp * max(n_locked, 1))(*[locked[i].ctypes.data for i in range(n_locked)])
        elif locked_blocks is not None and n_locked > 0:
            blocks = np.ascontiguousarray(locked_blocks, dtype=dt)
            assert blocks.size == n
        else:
            n_locked = 0
        evals = np.zeros(nroot, dtype=rt)
        evecs = np.zeros((nroot, n), dtype=dt) if want_vectors else None
        alpha = np.zeros(max_iter + 1, dtype=np.float64)
        beta = np.zeros(max_iter + 1, dtype=np.float64) if want_beta else None
        t_mv = np.zeros(max_iter + 2, dtype=np.float64)
        cap = np.zeros((capture, n), dtype=dt) if capture > 0 else None
        nvals, calls, cap_count = i64(0), i64(0), i64(0)
        import time as _time

        t0 = _time.perf_counter()
        it = getattr(self.lib, f"ref_run_iteration_{sfx}")(
            i64(n), _ptr(rowptr), _ptr(colidx), _ptr(vals), C.c_int(mv_threads), C.c_int(int(find_max)),
            C.c_double(offset), C.c_double(eps), i64(max_iter), i64(nroot), _ptr(init), i64(n_locked), ptrs,
            _ptr(blocks), _ptr(evals), _ptr(evecs), C.byref(nvals), _ptr(alpha), _ptr(beta), _ptr(t_mv),
            C.byref(calls), i64(capture), _ptr(cap), C.byref(cap_count))
        seconds = _time.perf_counter() - t0
        c = calls.value
        return {"iterations": int(it), "eigenvalues": evals[: nvals.value].copy(),
                "eigenvectors": None if evecs is None else evecs[: nvals.value].copy(), "alpha": alpha[:c].copy(),
                "beta": None if beta is None else beta[: max(c - 1, 0)].copy(), "dt_iter": np.diff(t_mv[:c]),
                "setup_seconds": float(t_mv[0]) if c else 0.0,
                "basis": None if cap is None else cap[: cap_count.value].copy(), "seconds": seconds, "mv_calls": c}

    def lanczos(self, rowptr, colidx, vals, *, find_max, num_eigs=1, offset=0.0, eps=-1.0, max_iter=0, nepi=0,
                init=None, max_runs=64, mv_threads=1, capture=0, want_vectors=True):
        dt = np.dtype(vals.dtype)
        sfx = _SFX[dt]
        rt = _REAL[sfx]
        rowptr, colidx, vals = _csr(rowptr, colidx, vals, dt)
        n = rowptr.size - 1
        init = None if init is None else np.ascontiguousarray(init, dtype=dt)
        evals = np.zeros(num_eigs, dtype=rt)
        evecs = np.zeros((num_eigs, n), dtype=dt) if want_vectors else None
        iters = np.zeros(max_runs, dtype=np.int64)
        timing = np.zeros(2, dtype=np.float64)
        cap = np.zeros((capture, n), dtype=dt) if capture > 0 else None
        n_runs, n_found, cap_count = i64(0), i64(0), i64(0)
        rc = getattr(self.lib, f"ref_lanczos_run_{sfx}")(
            i64(n), _ptr(rowptr), _ptr(colidx), _ptr(vals), C.c_int(mv_threads), C.c_int(int(find_max)),
            i64(num_eigs), C.c_double(offset), C.c_double(eps), i64(max_iter), i64(nepi), _ptr(init), _ptr(evals),
            _ptr(evecs), _ptr(iters), i64(max_runs), C.byref(n_runs), C.byref(n_found), _ptr(timing), i64(capture),
            _ptr(cap), C.byref(cap_count))
        assert rc == 0
        nf = n_found.value
        return LanczosResult(evals[:nf].copy(), None if evecs is None else evecs[:nf].copy(),
                             [int(x) for x in iters[: n_runs.value]], float(timing[0]), float(timing[1]),
                             None if cap is None else cap[: cap_count.value].copy())

    def expm(self, rowptr, colidx, vals, a, x, *, eps=-1.0, full_orth=False, max_iter=0, taylor=False, mv_threads=1):
        dt = np.dtype(vals.dtype)
        sfx = _SFX[dt]
        rowptr, colidx, vals = _csr(rowptr, colidx, vals, dt)
        n = rowptr.size - 1
        x = np.ascontiguousarray(x, dtype=dt)
        out = np.zeros(n, dtype=dt)
        timing = np.zeros(2, dtype=np.float64)
        z = complex(a)
        it = getattr(self.lib, f"ref_expm_run_{sfx}")(
            i64(n), _ptr(rowptr), _ptr(colidx), _ptr(vals), C.c_int(mv_threads), C.c_double(z.real),
            C.c_double(z.imag), _ptr(x), _ptr(out), C.c_double(eps), C.c_int(int(full_orth)), i64(max_iter),
            C.c_int(int(taylor)), _ptr(timing))
        return int(it), out

    def inner_prod(self, a, b):
        dt = np.dtype(a.dtype)
        sfx = _SFX[dt]
        a = np.ascontiguousarray(a)
        b = np.ascontiguousarray(b, dtype=dt)
        out = np.zeros(1, dtype=dt)
        getattr(self.lib, f"ref_inner_prod_{sfx}")(i64(a.size), _ptr(a), _ptr(b), _ptr(out))
        return out[0]

    def norm(self, a):
        sfx = _SFX[np.dtype(a.dtype)]
        a = np.ascontiguousarray(a)
        return float(getattr(self.lib, f"ref_norm_{sfx}")(i64(a.size), _ptr(a)))

    def schmidt_orth(self, basis, w):
        dt = np.dtype(w.dtype)
        sfx = _SFX[dt]
        basis = np.ascontiguousarray(basis, dtype=dt)
        w = np.array(w, dtype=dt, copy=True)
        getattr(self.lib, f"ref_schmidt_orth_{sfx}")(i64(w.size), i64(basis.shape[0]), _ptr(basis), _ptr(w))
        return w

    def tridiag(self, alpha, beta, vectors=True):
        alpha = np.ascontiguousarray(alpha)
        dt = np.dtype(alpha.dtype)
        sfx = _SFX[dt]
        m = alpha.size
        beta = np.ascontiguousarray(beta, dtype=dt)
        ev = np.zeros(m, dtype=dt)
        q = np.zeros((m, m), dtype=dt) if vectors else None
        unc = getattr(self.lib, f"ref_tridiag_eigenpairs_{sfx}")(i64(m), _ptr(alpha), _ptr(beta), i64(beta.size),
                                                                 _ptr(ev), _ptr(q))
        return ev, q, int(unc)


def have_reference() -> bool:
    return os.path.exists(REF_SO)


def best(fast: bool = False):
    """The strongest checker available: the compiled reference if present, else the restatement.  ``fast`` selects the
    -O3 build of the reference shim and is for TIMING only."""
    return Reference(fast=fast) if have_reference() else Restatement()
