/* llz.h — C ABI of the B200-native Lanczos engine (libllz.so).
 *
 * This is the drop-in boundary for the ONE hot path of mrcdr/lambda-lanczos: the Krylov iteration inside
 * LambdaLanczos<T>::run and Exponentiator<T>::run.  Plain pointers and sizes only; no C++ or torch types.
 * Every entry point names the reference interface (file:line under /root/reference/include/lambda_lanczos/) that it
 * replaces.  The reference has no FFI of its own (it is a header-only C++ library); the C++ header
 * lambda-lanczos_b200/include/lambda_lanczos_b200/lambda_lanczos.hpp rebuilds the reference's classes on top of
 * exactly these calls, and INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns an llz_status_t (0 = OK); llz_last_error() gives the message for the calling thread;
 *   - dtype selects the scalar type T of vectors and operators; alpha/beta/eigenvalues/eps/offset are always passed
 *     as double across this ABI (real_t<T> of the reference, util/common.hpp:80-102);
 *   - complex scalars are interleaved (re, im) pairs, layout-compatible with std::complex<> and C99 _Complex;
 *   - "device pointer" arguments must be 16-byte aligned device memory on the context's GPU;
 *   - one context = one GPU + one stream (+ one rank of a row-sharded group); a context is not thread-safe, the
 *     reference engine objects are not either (lambda_lanczos.hpp:331,395-403).
 *   - there is NO CPU fallback: without a usable CUDA device llz_ctx_create fails with LLZ_ERR_NO_DEVICE.
 */
#ifndef LLZ_H_
#define LLZ_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LLZ_VERSION 100 /* 0.1.0 */

typedef struct llz_ctx_s* llz_ctx_t;
typedef struct llz_op_s* llz_op_t;
typedef struct llz_vec_s* llz_vec_t;
typedef struct llz_krylov_s* llz_krylov_t;

typedef enum {
  LLZ_F32 = 0,  /* float                */
  LLZ_F64 = 1,  /* double               */
  LLZ_C64 = 2,  /* std::complex<float>  */
  LLZ_C128 = 3  /* std::complex<double> */
} llz_dtype_t;

typedef enum {
  LLZ_OK = 0,
  LLZ_ERR_INVALID = 1,     /* bad argument / shape / dtype */
  LLZ_ERR_CUDA = 2,        /* CUDA runtime or driver error */
  LLZ_ERR_OOM = 3,         /* device memory exhausted (e.g. the Krylov basis cannot grow further) */
  LLZ_ERR_COMM = 4,        /* inter-GPU communication error */
  LLZ_ERR_UNSUPPORTED = 5, /* valid request this build cannot serve */
  LLZ_ERR_NO_DEVICE = 6,   /* no CUDA device: the engine has no CPU path */
  LLZ_ERR_USER = 7         /* a user callback reported failure */
} llz_status_t;

/* How llz_krylov_step orthogonalises the new Lanczos vector. */
typedef enum {
  LLZ_ORTH_RECURRENCE = 0, /* three-term recurrence only          (exponentiator.hpp:112-118, full_orthogonalize=false) */
  LLZ_ORTH_FULL = 1,       /* recurrence + one classical Gram-Schmidt pass against [locked, basis]
                              (replaces the MGS sweeps of lambda_lanczos.hpp:259-260, exponentiator.hpp:120-122) */
  LLZ_ORTH_FULL_TWICE = 2, /* as FULL, followed by a second projection pass (CGS2) */
  LLZ_ORTH_RECURRENCE_LAZY = 3 /* RECURRENCE without the normalisation pass (2 n s bytes per iteration less): column k >= 1
                              is stored as beta_{k-1} u_k, the recurrence of the next iterations absorbs the factors
                              (w = A(beta u)/beta ...), alpha / beta are reported as usual; the caller divides the
                              coefficient of column k by beta_{k-1} in llz_krylov_combine.  Single rank only; every
                              iteration of a run must use it (what Exponentiator<T>::run does, exponentiator.hpp:106-160) */
} llz_orth_t;

int llz_version(void);
const char* llz_status_string(int status);
const char* llz_last_error(void);

/* ------------------------------------------------------------------------------------------------------------------
 * Context
 * ---------------------------------------------------------------------------------------------------------------- */
int llz_ctx_create(int device, llz_ctx_t* ctx);
/* Same, but launch everything on a caller-owned cudaStream_t (passed as void*). */
int llz_ctx_create_on_stream(int device, void* cuda_stream, llz_ctx_t* ctx);
int llz_ctx_destroy(llz_ctx_t ctx);
int llz_ctx_synchronize(llz_ctx_t ctx);
int llz_ctx_stream(llz_ctx_t ctx, void** cuda_stream);
/* Copy between host and device memory in the context's stream order and wait for it (to_device != 0: host -> device).
 * What a HOST-side mv_mul adapter needs (lambda_lanczos.hpp:120-126: the reference's mv_mul reads and writes host
 * std::vectors): DeviceOperator<T>::host_function stages x out and y back in around the user's callable. */
int llz_ctx_memcpy(llz_ctx_t ctx, void* dst, const void* src, size_t bytes, int to_device);
/* A context keeps the device memory of destroyed vectors and of the last destroyed Krylov workspace (the mapped basis
 * slab) for reuse by the next run — mapping tens of GB of HBM costs as much as hundreds of Lanczos iterations.
 * This call returns all of it to the driver. */
int llz_ctx_release_cache(llz_ctx_t ctx);
/* Number of kernels this context has launched so far (the bench's "gpu_launches" evidence). */
int llz_ctx_launch_count(llz_ctx_t ctx, uint64_t* count);
/* Per-kernel-family device time (CUDA events on the context's stream) accumulated since profiling was last switched
 * on.  Families: "spmv", "dot", "project", "reduce", "update", "recurrence", "scale", "combine".  `bytes` is the
 * algorithmic traffic of the timed launches (SURVEY.md §8d accounting), so bytes/ms is the achieved bandwidth. */
int llz_ctx_profile(llz_ctx_t ctx, int enable);
int llz_ctx_profile_read(llz_ctx_t ctx, const char* name, double* ms, int64_t* launches, double* bytes);
/* Join a row-sharded group of `nranks` contexts (one per GPU, one process each).  `comm_id` is the 128-byte blob
 * produced by llz_comm_unique_id on rank 0 and distributed by the caller (e.g. torch.distributed / MPI / a file).
 * Afterwards every vector is the local row block of a global vector and all reductions are group-wide. */
int llz_comm_unique_id(void* id128);
int llz_ctx_join(llz_ctx_t ctx, int rank, int nranks, const void* id128);
int llz_ctx_rank(llz_ctx_t ctx, int* rank, int* nranks);
/* *enabled = 1 when the joined group exchanges the per-iteration scalars (alpha, the projection coefficients, beta^2)
 * through CUDA-IPC mapped peer memory written from inside the producing kernels (csrc/llz_peer.cuh), 0 when it uses
 * NCCL all-reduces (single rank, IPC unavailable, or LLZ_P2P=0 in the environment). */
int llz_ctx_peer_channels(llz_ctx_t ctx, int* enabled);
/* The row partition every built-in operator and the bench use: rank r owns [b(r), b(r+1)) with b(r) = floor(n*r/G)
 * rounded down to a multiple of 4 (b(0) = 0, b(G) = n), so row blocks start 16-byte aligned.  Pure host arithmetic. */
int llz_partition(int64_t n_global, int rank, int nranks, int64_t* row0, int64_t* n_local);
/* Host-side halo planning for the local row block of a CSR matrix with GLOBAL column indices (no GPU involved):
 * boundaries[0..nranks] are the row-block boundaries of the group.  Outputs (each may be NULL): the column indices in
 * the local extended numbering ([0,n_rows) own block, n_rows + h = h-th halo entry), the sorted global columns of the
 * halo entries (capacity halo_capacity), their number, and how many of them each rank owns. */
int llz_halo_plan(int64_t n_rows, int64_t row0, const int64_t* rowptr, const int32_t* colidx, int nranks,
                  const int64_t* boundaries, int32_t* colidx_local, int64_t* halo_cols, int64_t halo_capacity,
                  int64_t* n_halo, int64_t* per_owner);

/* ------------------------------------------------------------------------------------------------------------------
 * Operators — the device-side replacement of the `mv_mul` std::function
 * (lambda_lanczos.hpp:120-126, exponentiator.hpp:35-41).  Contract kept: one apply per iteration, never concurrent.
 * Built-in operators OVERWRITE y (the reference pre-zeroes `out`, lambda_lanczos.hpp:242, and lets the callee
 * accumulate; y = 0 + A x is the same result without the extra pass).
 * ---------------------------------------------------------------------------------------------------------------- */
/* CSR with 32-bit column indices.  Arrays are copied to the device (host_arrays != 0) or adopted by copy from device
 * memory (host_arrays == 0).  In a joined context `n_rows` is the local row block starting at global row `row0`, and
 * column indices are global. */
int llz_op_create_csr(llz_ctx_t ctx, int dtype, int64_t n_rows, int64_t n_cols, int64_t row0, const int64_t* rowptr,
                      const int32_t* colidx, const void* vals, int host_arrays, llz_op_t* op);
/* SELL-C-sigma with C = 32 (one warp per slice): takes the same CSR arrays as llz_op_create_csr and re-stores them on
 * the device in sliced-ELL form — coalesced, barrier-free SpMV for short-row matrices (stencils, lattice
 * Hamiltonians).  sigma: 1 keeps the row order, a multiple of 32 (<= 1024) sorts windows of sigma rows by length to cut
 * padding, 0 picks.  y is bit-identical to the CSR operator's (same per-row summation order). */
int llz_op_create_sell(llz_ctx_t ctx, int dtype, int64_t n_rows, int64_t n_cols, int64_t row0, const int64_t* rowptr,
                       const int32_t* colidx, const void* vals, int host_arrays, int sigma, llz_op_t* op);
/* Matrix-free spin-1/2 XXZ chain  H = sum_b Jxy/2 (S+S- + h.c.) + Jz SzSz  on L sites in the sector with n_up up
 * spins, basis states in increasing integer order (BASELINE.json configs 4 and 5). */
int llz_op_create_xxz(llz_ctx_t ctx, int dtype, int L, int n_up, double jz, double jxy, int periodic, llz_op_t* op);
/* User operator.  `apply` must enqueue y (+)= A x on `cuda_stream` and return 0.  If overwrites_y == 0 the engine
 * zero-fills y first, as the reference promises its callee (lambda_lanczos.hpp:124). */
typedef int (*llz_apply_fn)(void* user, const void* x_dev, void* y_dev, int64_t n_local, void* cuda_stream);
int llz_op_create_callback(llz_ctx_t ctx, int dtype, int64_t n_local, llz_apply_fn apply, void* user,
                           int overwrites_y, llz_op_t* op);
int llz_op_destroy(llz_op_t op);
int llz_op_rows(llz_op_t op, int64_t* n_local);
/* Local rows, rows of the whole operator and first global row of the local block (n_global = n_local, row0 = 0 for a
 * single rank).  The reference's `matrix_size` (lambda_lanczos.hpp:136) is n_global. */
int llz_op_shape(llz_op_t op, int64_t* n_local, int64_t* n_global, int64_t* row0);
/* How the operator is held on the device: "DIA", "SELL-32", "SELL-32-sigma (sorted windows)", "CSR (stream kernel)",
 * "XXZ matrix-free (block kernel)", "user callback" ... (static string; "" for a null handle).  llz_op_create_sell with
 * sigma = 0 picks DIA for operators whose non-zeros lie on at most 16 diagonals, else SELL. */
const char* llz_op_storage(llz_op_t op);
/* Algorithmic bytes one apply has to move for the operator itself (A_bytes of SURVEY.md §8d; 0 for matrix-free). */
int llz_op_bytes(llz_op_t op, int64_t* bytes);
/* Gerschgorin radius max_i sum_j |a_ij| of the whole operator (group-wide for row-sharded operators): every eigenvalue
 * lies in [-radius, radius], so eigenvalue_offset = -radius (find_maximum) or +radius (minimum) makes the wanted
 * end of the spectrum the dominant one — the job of the reference's stand-alone helper
 * src/determine_eigenvalue_offset/determine_eigenvalue_offset.cpp:12-49, on the device.  CSR / SELL: exact;
 * XXZ: the analytic bound; user callbacks: LLZ_ERR_UNSUPPORTED. */
int llz_op_gerschgorin_radius(llz_op_t op, double* radius);
/* y = A x on device vectors (stand-alone use and Exponentiator::taylor_run, exponentiator.hpp:191). */
int llz_op_apply(llz_op_t op, llz_vec_t x, llz_vec_t y);

/* ------------------------------------------------------------------------------------------------------------------
 * Device vectors and the util:: vector kernels (util/linear_algebra.hpp)
 * ---------------------------------------------------------------------------------------------------------------- */
int llz_vec_create(llz_ctx_t ctx, int dtype, int64_t n_local, llz_vec_t* v);
int llz_vec_destroy(llz_vec_t v);
int llz_vec_upload(llz_vec_t v, const void* host);  /* H2D, n_local elements */
int llz_vec_download(llz_vec_t v, void* host);      /* D2H (synchronises the stream) */
int llz_vec_device_ptr(llz_vec_t v, void** dev);
int llz_vec_copy(llz_vec_t dst, llz_vec_t src);
int llz_vec_fill_zero(llz_vec_t v);
/* out[0..1] = <a,b> = sum conj(a_i) b_i   (util::inner_prod, linear_algebra.hpp:30-51; conjugates the FIRST argument) */
int llz_vec_dot(llz_vec_t a, llz_vec_t b, double out[2]);
/* *out = sqrt(Re<v,v>)                     (util::norm, linear_algebra.hpp:57-60) */
int llz_vec_norm(llz_vec_t v, double* out);
/* *out = sum_i |Re v_i| + |Im v_i|         (util::m_norm, linear_algebra.hpp:83-125: the _ASUM definition) */
int llz_vec_m_norm(llz_vec_t v, double* out);
/* v *= a                                   (util::scalar_mul, linear_algebra.hpp:66-72) */
int llz_vec_scale(llz_vec_t v, const double a[2]);
/* v *= 1/norm(v), *norm_out = norm before  (util::normalize, linear_algebra.hpp:78-80) */
int llz_vec_normalize(llz_vec_t v, double* norm_out);
/* y += a x */
int llz_vec_axpy(llz_vec_t y, const double a[2], llz_vec_t x);
/* w -= sum_j <u_j,w> u_j over `count` orthonormal vectors (util::schmidt_orth, linear_algebra.hpp:133-144), done as
 * `passes` classical Gram-Schmidt passes (projection V^H w, then update w - V h) instead of the reference's
 * vector-by-vector modified Gram-Schmidt. */
int llz_vec_schmidt_orth(llz_vec_t w, const llz_vec_t* basis, int64_t count, int passes);

/* ------------------------------------------------------------------------------------------------------------------
 * Krylov workspace — the device-resident Lanczos basis (`u`, lambda_lanczos.hpp:221; exponentiator.hpp:90) plus the
 * alpha/beta recurrences (:222-223).  Column-major, one allocation, columns 256-byte aligned, grown on demand.
 * ---------------------------------------------------------------------------------------------------------------- */
/* max_cols: upper bound on stored Lanczos vectors (max_iteration + 1); the store reserves address space for
 * min(max_cols, what fits in device memory) columns and maps physical memory as the iteration advances. */
int llz_krylov_create(llz_ctx_t ctx, int dtype, int64_t n_local, int64_t max_cols, llz_krylov_t* kry);
int llz_krylov_destroy(llz_krylov_t kry);
/* Largest number of columns this workspace can ever hold (address space and device memory permitting). */
int llz_krylov_capacity(llz_krylov_t kry, int64_t* max_cols);
/* Vectors every new Lanczos vector is additionally orthogonalised against (`orthogonalizeTo`,
 * lambda_lanczos.hpp:220,233,259): the eigenvectors kept by earlier runs.  The handles must stay alive. */
int llz_krylov_set_locked(llz_krylov_t kry, const llz_vec_t* locked, int64_t count);
/* Start a run: column 0 := start (host pointer if host != 0, else device pointer), orthogonalised against the locked
 * vectors, normalised (lambda_lanczos.hpp:231-234; exponentiator.hpp:100-101).  *norm_out = norm before normalising. */
int llz_krylov_begin(llz_krylov_t kry, const void* start, int host, double* norm_out);
/* Enqueue Lanczos iteration k = ncols (1-based as in lambda_lanczos.hpp:240):
 *   w = (A + sigma I) u_{k-1}; alpha_{k-1} = Re<u_{k-1}, w>                  (:242-248, exponentiator.hpp:107-110)
 *   u_k = w - alpha u_{k-1} - beta_{k-2} u_{k-2}, orthogonalised per `orth`  (:250-260, exponentiator.hpp:112-122)
 *   beta_{k-1} = ||u_k||; u_k /= beta_{k-1}                                  (:262,285, exponentiator.hpp:145,160)
 * Asynchronous: returns once the work is queued; the scalars arrive through llz_krylov_fetch. */
int llz_krylov_step(llz_krylov_t kry, llz_op_t op, double sigma, int orth);
/* Block until iteration k (1-based) has finished and return alpha_{k-1}, beta_{k-1} and ||w'||, the norm of the new
 * vector after the three-term recurrence but BEFORE the Gram-Schmidt pass (= beta for LLZ_ORTH_RECURRENCE).
 * beta << ||w'|| means the pass cancelled most of the vector and one pass was not enough (DGKS criterion). */
int llz_krylov_fetch(llz_krylov_t kry, int64_t k, double* alpha, double* beta, double* wnorm);
/* Repeat the Gram-Schmidt pass on the (normalised) vector of iteration k and renormalise it; beta_{k-1} is multiplied
 * by the shrink factor returned in *shrink.  Iterations already enqueued beyond k are discarded (steps() == k
 * afterwards).  Synchronous; meant for the rare near-breakdown iterations. */
int llz_krylov_refine(llz_krylov_t kry, int64_t k, double* shrink);
/* Number of iterations enqueued so far (stored columns = this + 1). */
int llz_krylov_steps(llz_krylov_t kry, int64_t* k);
/* out_r = sum_{j<m} coeff[r*m + j] u_j for r < nvec, optionally normalised — the eigenvector assembly of
 * compute_eigenvectors (lambda_lanczos.hpp:51-58) and the Exponentiator's output sum (exponentiator.hpp:163-170) in
 * ONE pass over the basis.  coeff is a host array of T (interleaved complex), out are device vectors. */
int llz_krylov_combine(llz_krylov_t kry, int64_t m, int64_t nvec, const void* coeff, int normalize,
                       const llz_vec_t* out);
/* Device pointer of column j (for tests: the oracle's Lanczos vectors are observable through its mv_mul spy). */
int llz_krylov_column_ptr(llz_krylov_t kry, int64_t j, void** dev);
int llz_krylov_download_column(llz_krylov_t kry, int64_t j, void* host);
/* ------------------------------------------------------------------------------------------------------------------
 * Whole-engine entry points (what a non-C++ caller binds): the host control loop of the reference, restated in
 * lambda_lanczos_b200/lambda_lanczos.hpp and exponentiator.hpp, instantiated for the four dtypes.
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct {
  int find_maximum;               /* lambda_lanczos.hpp:153 */
  int64_t num_eigs;               /* :156 */
  double eigenvalue_offset;       /* :165 */
  double eps;                     /* :150; <= 0 selects the reference default 1e3*machine-eps of real_t<T> */
  int64_t max_iteration;          /* :138; <= 0 selects matrix_size */
  int64_t num_eigs_per_iteration; /* :173; <= 0 selects 5 */
  int orth;                       /* llz_orth_t; the reference behaviour is LLZ_ORTH_FULL */
  int pipeline_depth;             /* iterations the GPU may run ahead of the host convergence test (0 = lock-step,
                                     < 0 = chosen from the vector size: 4 below 4 MB, 2 below 32 MB, else 1) */
  int ritz_solver;                /* 0 = bisection on the nroot extreme Ritz values, 1 = full implicit QL every step */
} llz_eigs_params_t;

typedef struct {
  double seconds_total;    /* wall time of the run, host clock */
  double seconds_host;     /* host tridiagonal solves + convergence logic */
  int64_t iterations;      /* Lanczos iterations over all runs */
  int64_t runs;            /* Lanczos runs (lambda_lanczos.hpp:334 loop) */
  int64_t basis_bytes;     /* peak bytes of the mapped Krylov basis */
  uint64_t kernel_launches;
} llz_run_stats_t;

/* LambdaLanczos<T>(op, n, find_maximum, num_eigs).run(eigenvalues, eigenvectors)  (lambda_lanczos.hpp:200,330-366).
 * start: host vector handed out by `init_vector` at the start of every Lanczos run (:133,232); NULL selects a seeded
 * uniform[-1,1] vector.  eigenvalues_out: num_eigs doubles.  eigenvectors_out: host, num_eigs x n_local elements of T
 * (may be NULL).  iter_counts: getIterationCounts() (:412), at most max_runs entries. */
int llz_eigs_run(llz_ctx_t ctx, llz_op_t op, int dtype, const llz_eigs_params_t* params, const void* start,
                 double* eigenvalues_out, void* eigenvectors_out, int64_t* n_found, int64_t* iter_counts,
                 int64_t max_runs, int64_t* n_runs, llz_run_stats_t* stats);

/* Exponentiator<T>(op, n).run(a, input, output)  (exponentiator.hpp:80,87-173): output = exp(a A) input.
 * a is (re, im); im is ignored for real dtypes.  input/output are host vectors (host != 0) or device pointers.
 * eps <= 0 selects 1e2*machine-eps (:58); max_iteration <= 0 selects n (:81).  taylor != 0 runs taylor_run (:175). */
int llz_expm_run(llz_ctx_t ctx, llz_op_t op, int dtype, const double a[2], const void* input, void* output, int host,
                 double eps, int full_orthogonalize, int64_t max_iteration, int taylor, int64_t* iterations);

#ifdef __cplusplus
}
#endif
#endif /* LLZ_H_ */
