"""lambda_lanczos_b200 — Python face of the B200-native Lanczos engine.

The product is ``libllz.so`` (hand-written sm_100a kernels + C ABI, ``include/llz.h``) and the header-only C++ host
engine under ``include/lambda_lanczos_b200/``.  This module is only a thin ctypes binding over that C ABI so that the
parity tests and ``bench.py`` can drive it; it mirrors the reference's two engines:

* :class:`LambdaLanczos`  -> ``lambda_lanczos::LambdaLanczos<T>``   (reference lambda_lanczos.hpp:109-415)
* :class:`Exponentiator`  -> ``lambda_lanczos::Exponentiator<T>``   (reference exponentiator.hpp:24-211)

There is no CPU path: importing works anywhere, but creating a :class:`Context` without a CUDA device raises, and a
missing ``libllz.so`` raises at import of the library (``lib()``) with the build instruction.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libllz.so")

F32, F64, C64, C128 = 0, 1, 2, 3
_DTYPES = {np.dtype(np.float32): F32, np.dtype(np.float64): F64, np.dtype(np.complex64): C64, np.dtype(np.complex128): C128}
_NP = {F32: np.float32, F64: np.float64, C64: np.complex64, C128: np.complex128}
ORTH_RECURRENCE, ORTH_FULL, ORTH_FULL_TWICE, ORTH_RECURRENCE_LAZY = 0, 1, 2, 3

i64 = C.c_int64
vp = C.c_void_p


class LlzError(RuntimeError):
    def __init__(self, status, where, message):
        super().__init__(f"{where}: status {status} ({message})")
        self.status = status


class EigsParams(C.Structure):
    _fields_ = [("find_maximum", C.c_int), ("num_eigs", i64), ("eigenvalue_offset", C.c_double), ("eps", C.c_double),
                ("max_iteration", i64), ("num_eigs_per_iteration", i64), ("orth", C.c_int), ("pipeline_depth", C.c_int),
                ("ritz_solver", C.c_int)]


class RunStats(C.Structure):
    _fields_ = [("seconds_total", C.c_double), ("seconds_host", C.c_double), ("iterations", i64), ("runs", i64),
                ("basis_bytes", i64), ("kernel_launches", C.c_uint64)]


def build(force: bool = False) -> str:
    """Compile libllz.so in-tree for sm_100a (``make -C lambda-lanczos_b200``)."""
    if force or not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "-j8", "--no-print-directory"] + (["-B"] if force else []))
    return LIB_PATH


_lib = None


def lib():
    """The loaded C ABI.  Fails loudly when the CUDA library has not been built: there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing — build it with `make -C {_HERE}` (or __graft_entry__.build()); "
                              "this package has no CPU or PyTorch fallback")
        _lib = C.CDLL(LIB_PATH)
        _lib.llz_status_string.restype = C.c_char_p
        _lib.llz_last_error.restype = C.c_char_p
    return _lib


def _check(status, where):
    if status != 0:
        raise LlzError(status, where, (lib().llz_last_error() or b"").decode())


def _ptr(a):
    return None if a is None else a.ctypes.data_as(vp)


def dtype_code(dt) -> int:
    return _DTYPES[np.dtype(dt)]


class Context:
    """One GPU + one stream (``llz_ctx_t``)."""

    def __init__(self, device: int = 0):
        self.h = vp()
        _check(lib().llz_ctx_create(C.c_int(device), C.byref(self.h)), "llz_ctx_create")
        self.device = device
        self.rank, self.nranks = 0, 1

    @staticmethod
    def unique_id() -> bytes:
        """128-byte blob identifying a new row-sharded group; make it on rank 0 and hand it to every rank."""
        buf = (C.c_ubyte * 128)()
        _check(lib().llz_comm_unique_id(buf), "llz_comm_unique_id")
        return bytes(buf)

    def join(self, rank: int, nranks: int, comm_id: bytes):
        """Join a row-sharded group (one process per GPU, NCCL underneath): vectors become local row blocks."""
        buf = (C.c_ubyte * 128).from_buffer_copy(comm_id)
        _check(lib().llz_ctx_join(self.h, C.c_int(rank), C.c_int(nranks), buf), "llz_ctx_join")
        self.rank, self.nranks = rank, nranks

    def peer_channels(self) -> bool:
        """True when the per-iteration scalars travel through peer memory (NVLink stores from inside the kernels)."""
        e = C.c_int(0)
        _check(lib().llz_ctx_peer_channels(self.h, C.byref(e)), "llz_ctx_peer_channels")
        return bool(e.value)

    def synchronize(self):
        _check(lib().llz_ctx_synchronize(self.h), "llz_ctx_synchronize")

    def launch_count(self) -> int:
        c = C.c_uint64(0)
        _check(lib().llz_ctx_launch_count(self.h, C.byref(c)), "llz_ctx_launch_count")
        return int(c.value)

    def release_cache(self):
        """Give the cached device memory (vector pool, mapped Krylov basis of the last run) back to the driver."""
        _check(lib().llz_ctx_release_cache(self.h), "llz_ctx_release_cache")

    def profile(self, enable: bool):
        _check(lib().llz_ctx_profile(self.h, C.c_int(int(enable))), "llz_ctx_profile")

    def profile_read(self, name: str):
        """(milliseconds, launches, algorithmic bytes) of one kernel family since profiling was enabled."""
        ms, cnt, by = C.c_double(0), i64(0), C.c_double(0)
        _check(lib().llz_ctx_profile_read(self.h, name.encode(), C.byref(ms), C.byref(cnt), C.byref(by)), "llz_ctx_profile_read")
        return float(ms.value), int(cnt.value), float(by.value)

    def stream(self) -> int:
        s = vp()
        _check(lib().llz_ctx_stream(self.h, C.byref(s)), "llz_ctx_stream")
        return int(s.value or 0)

    def close(self):
        if self.h:
            lib().llz_ctx_destroy(self.h)
            self.h = vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Operator:
    """Device operator (``llz_op_t``): the replacement of the reference's ``mv_mul`` std::function."""

    def __init__(self, ctx: Context, handle, dtype, n=None):
        self.ctx, self.h, self.dtype = ctx, handle, np.dtype(dtype)
        nl, ng, r0 = i64(0), i64(0), i64(0)
        _check(lib().llz_op_shape(handle, C.byref(nl), C.byref(ng), C.byref(r0)), "llz_op_shape")
        self.n, self.n_global, self.row0 = int(nl.value), int(ng.value), int(r0.value)  # n = rows of the LOCAL block

    @classmethod
    def csr(cls, ctx: Context, rowptr, colidx, vals, row0: int = 0, n_cols: int | None = None):
        """CSR operator.  In a joined context pass the local row block: ``rowptr`` local (starts at 0), ``colidx``
        GLOBAL, ``row0`` the first global row and ``n_cols`` the global dimension."""
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
        colidx = np.ascontiguousarray(colidx, dtype=np.int32)
        vals = np.ascontiguousarray(vals)
        n = rowptr.size - 1
        h = vp()
        _check(lib().llz_op_create_csr(ctx.h, C.c_int(dtype_code(vals.dtype)), i64(n), i64(n if n_cols is None else n_cols),
                                       i64(row0), _ptr(rowptr), _ptr(colidx), _ptr(vals), C.c_int(1), C.byref(h)),
               "llz_op_create_csr")
        return cls(ctx, h, vals.dtype, n)

    @classmethod
    def sell(cls, ctx: Context, rowptr, colidx, vals, sigma: int = 0, row0: int = 0, n_cols: int | None = None):
        """SELL-32-sigma operator built on the device from CSR arrays (same arguments as :meth:`csr`); ``sigma`` = 1
        keeps the row order, a multiple of 32 sorts windows of that many rows by length, 0 picks."""
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
        colidx = np.ascontiguousarray(colidx, dtype=np.int32)
        vals = np.ascontiguousarray(vals)
        n = rowptr.size - 1
        h = vp()
        _check(lib().llz_op_create_sell(ctx.h, C.c_int(dtype_code(vals.dtype)), i64(n), i64(n if n_cols is None else n_cols),
                                        i64(row0), _ptr(rowptr), _ptr(colidx), _ptr(vals), C.c_int(1), C.c_int(sigma), C.byref(h)),
               "llz_op_create_sell")
        return cls(ctx, h, vals.dtype, n)

    @classmethod
    def xxz(cls, ctx: Context, L, n_up=None, jz=1.0, jxy=1.0, periodic=True, dtype=np.float64):
        n_up = L // 2 if n_up is None else n_up
        h = vp()
        _check(lib().llz_op_create_xxz(ctx.h, C.c_int(dtype_code(dtype)), C.c_int(L), C.c_int(n_up), C.c_double(jz),
                                       C.c_double(jxy), C.c_int(int(periodic)), C.byref(h)), "llz_op_create_xxz")
        n = i64(0)
        _check(lib().llz_op_rows(h, C.byref(n)), "llz_op_rows")
        return cls(ctx, h, dtype, n.value)

    def bytes(self) -> int:
        b = i64(0)
        _check(lib().llz_op_bytes(self.h, C.byref(b)), "llz_op_bytes")
        return int(b.value)

    def gerschgorin_radius(self) -> float:
        """max_i sum_j |a_ij| (every eigenvalue lies in [-radius, radius]): the value to choose eigenvalue_offset from."""
        r = C.c_double(0)
        _check(lib().llz_op_gerschgorin_radius(self.h, C.byref(r)), "llz_op_gerschgorin_radius")
        return float(r.value)

    def storage(self) -> str:
        """How the operator is held on the device (llz_op_storage): 'DIA', 'SELL-32', 'CSR (stream kernel)', ..."""
        f = lib().llz_op_storage
        f.restype = C.c_char_p
        return f(self.h).decode()

    def apply(self, x: "Vector", y: "Vector"):
        _check(lib().llz_op_apply(self.h, x.h, y.h), "llz_op_apply")

    def matvec(self, x: np.ndarray) -> np.ndarray:
        vx, vy = Vector(self.ctx, self.dtype, self.n), Vector(self.ctx, self.dtype, self.n)
        vx.upload(x)
        self.apply(vx, vy)
        return vy.download()

    def close(self):
        if self.h:
            lib().llz_op_destroy(self.h)
            self.h = vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Vector:
    """Device vector (``llz_vec_t``) with the util:: vector kernels of the reference (util/linear_algebra.hpp)."""

    def __init__(self, ctx: Context, dtype, n: int):
        self.ctx, self.dtype, self.n = ctx, np.dtype(dtype), int(n)
        self.h = vp()
        _check(lib().llz_vec_create(ctx.h, C.c_int(dtype_code(dtype)), i64(n), C.byref(self.h)), "llz_vec_create")

    @classmethod
    def from_host(cls, ctx, a: np.ndarray):
        v = cls(ctx, a.dtype, a.size)
        v.upload(a)
        return v

    def upload(self, a):
        a = np.ascontiguousarray(a, dtype=self.dtype)
        assert a.size == self.n
        _check(lib().llz_vec_upload(self.h, _ptr(a)), "llz_vec_upload")

    def download(self) -> np.ndarray:
        out = np.empty(self.n, dtype=self.dtype)
        _check(lib().llz_vec_download(self.h, _ptr(out)), "llz_vec_download")
        return out

    def dot(self, other: "Vector"):
        out = (C.c_double * 2)()
        _check(lib().llz_vec_dot(self.h, other.h, out), "llz_vec_dot")
        return complex(out[0], out[1]) if self.dtype.kind == "c" else float(out[0])

    def norm(self) -> float:
        out = C.c_double(0)
        _check(lib().llz_vec_norm(self.h, C.byref(out)), "llz_vec_norm")
        return float(out.value)

    def m_norm(self) -> float:
        """util::m_norm: sum |Re v_i| + |Im v_i| (util/linear_algebra.hpp:83-125)."""
        out = C.c_double(0)
        _check(lib().llz_vec_m_norm(self.h, C.byref(out)), "llz_vec_m_norm")
        return float(out.value)

    def scale(self, a):
        z = complex(a)
        _check(lib().llz_vec_scale(self.h, (C.c_double * 2)(z.real, z.imag)), "llz_vec_scale")

    def normalize(self) -> float:
        out = C.c_double(0)
        _check(lib().llz_vec_normalize(self.h, C.byref(out)), "llz_vec_normalize")
        return float(out.value)

    def axpy(self, a, x: "Vector"):
        z = complex(a)
        _check(lib().llz_vec_axpy(self.h, (C.c_double * 2)(z.real, z.imag), x.h), "llz_vec_axpy")

    def schmidt_orth(self, basis, passes: int = 1):
        arr = (vp * max(len(basis), 1))(*[b.h for b in basis])
        _check(lib().llz_vec_schmidt_orth(self.h, arr, i64(len(basis)), C.c_int(passes)), "llz_vec_schmidt_orth")

    def close(self):
        if self.h:
            lib().llz_vec_destroy(self.h)
            self.h = vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Krylov:
    """Device-resident Krylov workspace (``llz_krylov_t``): step-level access for tests and profiling."""

    def __init__(self, ctx: Context, dtype, n: int, max_cols: int):
        self.ctx, self.dtype, self.n = ctx, np.dtype(dtype), int(n)
        self.h = vp()
        _check(lib().llz_krylov_create(ctx.h, C.c_int(dtype_code(dtype)), i64(n), i64(max_cols), C.byref(self.h)),
               "llz_krylov_create")

    def capacity(self) -> int:
        c = i64(0)
        _check(lib().llz_krylov_capacity(self.h, C.byref(c)), "llz_krylov_capacity")
        return int(c.value)

    def set_locked(self, vecs):
        arr = (vp * max(len(vecs), 1))(*[v.h for v in vecs])
        self._locked = list(vecs)
        _check(lib().llz_krylov_set_locked(self.h, arr, i64(len(vecs))), "llz_krylov_set_locked")

    def begin(self, start: np.ndarray) -> float:
        start = np.ascontiguousarray(start, dtype=self.dtype)
        nrm = C.c_double(0)
        _check(lib().llz_krylov_begin(self.h, _ptr(start), C.c_int(1), C.byref(nrm)), "llz_krylov_begin")
        return float(nrm.value)

    def step(self, op: Operator, sigma: float = 0.0, orth: int = ORTH_FULL):
        _check(lib().llz_krylov_step(self.h, op.h, C.c_double(sigma), C.c_int(orth)), "llz_krylov_step")

    def fetch(self, k: int, with_wnorm: bool = False):
        a, b, w = C.c_double(0), C.c_double(0), C.c_double(0)
        _check(lib().llz_krylov_fetch(self.h, i64(k), C.byref(a), C.byref(b), C.byref(w)), "llz_krylov_fetch")
        return (float(a.value), float(b.value), float(w.value)) if with_wnorm else (float(a.value), float(b.value))

    def refine(self, k: int) -> float:
        s = C.c_double(0)
        _check(lib().llz_krylov_refine(self.h, i64(k), C.byref(s)), "llz_krylov_refine")
        return float(s.value)

    def column(self, j: int) -> np.ndarray:
        out = np.empty(self.n, dtype=self.dtype)
        _check(lib().llz_krylov_download_column(self.h, i64(j), _ptr(out)), "llz_krylov_download_column")
        return out

    def combine(self, coeff: np.ndarray, normalize: bool = True):
        coeff = np.ascontiguousarray(np.atleast_2d(coeff), dtype=self.dtype)
        nvec, m = coeff.shape
        outs = [Vector(self.ctx, self.dtype, self.n) for _ in range(nvec)]
        arr = (vp * nvec)(*[v.h for v in outs])
        _check(lib().llz_krylov_combine(self.h, i64(m), i64(nvec), _ptr(coeff), C.c_int(int(normalize)), arr),
               "llz_krylov_combine")
        return outs

    def close(self):
        if self.h:
            lib().llz_krylov_destroy(self.h)
            self.h = vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class LambdaLanczos:
    """``LambdaLanczos(mv_mul, n, find_maximum, num_eigs)`` with the reference's public fields
    (lambda_lanczos.hpp:126-181,200); ``run()`` returns (eigenvalues, eigenvectors) like lambda_lanczos.hpp:376."""

    def __init__(self, mv_mul: Operator, matrix_size: int, find_maximum: bool, num_eigs: int = 1):
        assert mv_mul.n_global == matrix_size  # row-sharded: matrix_size is the global dimension
        self.mv_mul = mv_mul
        self.matrix_size = matrix_size
        self.max_iteration = matrix_size
        self.eps = float(np.finfo(np.zeros(1, mv_mul.dtype).real.dtype).eps) * 1e3
        self.find_maximum = find_maximum
        self.num_eigs = num_eigs
        self.eigenvalue_offset = 0.0
        self.num_eigs_per_iteration = 5
        # host array handed out at the start of every Lanczos run (row-sharded: the LOCAL block); None = seeded default
        self.init_vector = None
        self.orthogonalization = ORTH_FULL
        self.pipeline_depth = -1  # auto: deeper for small vectors (launch-latency bound), see auto_pipeline_depth
        self.ritz_solver = 0
        self.want_eigenvectors = True
        self._iter_counts = []
        self.stats = None

    def run(self, out=None):
        """``out``: optional (num_eigs, n_local) host array receiving the eigenvectors (e.g. a view of pinned memory)."""
        op = self.mv_mul
        n = op.n  # eigenvectors come back as local row blocks
        p = EigsParams(int(self.find_maximum), self.num_eigs, float(self.eigenvalue_offset), float(self.eps),
                       int(self.max_iteration), int(self.num_eigs_per_iteration), int(self.orthogonalization),
                       int(self.pipeline_depth), int(self.ritz_solver))
        start = None if self.init_vector is None else np.ascontiguousarray(self.init_vector, dtype=op.dtype)
        assert start is None or start.size == n, "init_vector must have the operator's (local) row count"
        evals = np.zeros(self.num_eigs, dtype=np.float64)
        evecs = None
        if self.want_eigenvectors:
            evecs = np.empty((self.num_eigs, n), dtype=op.dtype) if out is None else out
            assert evecs.shape == (self.num_eigs, n) and evecs.dtype == op.dtype and evecs.flags.c_contiguous
        iters = np.zeros(256, dtype=np.int64)
        n_found, n_runs = i64(0), i64(0)
        stats = RunStats()
        _check(lib().llz_eigs_run(op.ctx.h, op.h, C.c_int(dtype_code(op.dtype)), C.byref(p), _ptr(start), _ptr(evals),
                                  _ptr(evecs), C.byref(n_found), _ptr(iters), i64(iters.size), C.byref(n_runs),
                                  C.byref(stats)), "llz_eigs_run")
        self._iter_counts = [int(x) for x in iters[: n_runs.value]]
        self.stats = stats
        nf = n_found.value
        return evals[:nf].copy(), (None if evecs is None else evecs[:nf])

    def getIterationCounts(self):
        return list(self._iter_counts)


class Exponentiator:
    """``Exponentiator(mv_mul, n).run(a, input)`` -> (iterations, output)  (exponentiator.hpp:80,87-173)."""

    def __init__(self, mv_mul: Operator, matrix_size: int):
        assert mv_mul.n_global == matrix_size
        self.mv_mul = mv_mul
        self.matrix_size = matrix_size
        self.max_iteration = matrix_size
        self.eps = float(np.finfo(np.zeros(1, mv_mul.dtype).real.dtype).eps) * 1e2
        self.full_orthogonalize = False

    def _run(self, a, x, taylor):
        op = self.mv_mul
        x = np.ascontiguousarray(x, dtype=op.dtype)
        assert x.size == op.n
        out = np.empty(op.n, dtype=op.dtype)
        z = complex(a)
        it = i64(0)
        _check(lib().llz_expm_run(op.ctx.h, op.h, C.c_int(dtype_code(op.dtype)), (C.c_double * 2)(z.real, z.imag), _ptr(x),
                                  _ptr(out), C.c_int(1), C.c_double(self.eps), C.c_int(int(self.full_orthogonalize)),
                                  i64(self.max_iteration), C.c_int(int(taylor)), C.byref(it)), "llz_expm_run")
        return int(it.value), out

    def run(self, a, x):
        return self._run(a, x, False)

    def taylor_run(self, a, x):
        return self._run(a, x, True)
