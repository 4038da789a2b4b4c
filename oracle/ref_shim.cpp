// oracle/ref_shim.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// C-ABI shim around the UNMODIFIED reference (mrcdr/lambda-lanczos).  The reference is header-only C++;
// this file is compiled against the headers WHERE THEY LIE (-I/root/reference/include/lambda_lanczos, see
// oracle/Makefile) into oracle/_ref/libllz_ref.so, so the reference's own machine code is what answers.  Nothing
// of the reference is copied into this repository.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load the resulting library.
//
// What is exposed (each wrapper just forwards to the reference symbol named in its comment):
//   ref_lanczos_run_<sfx>      -> lambda_lanczos::LambdaLanczos<T>::run            (lambda_lanczos.hpp:330-366)
//   ref_run_iteration_<sfx>    -> lambda_lanczos::LambdaLanczos<T>::run_iteration  (lambda_lanczos.hpp:216-322), with a spy
//                                 on mv_mul that records alpha_k, beta_k, the Lanczos vectors and one timestamp per call
//   ref_expm_run_<sfx>         -> lambda_lanczos::Exponentiator<T>::run/taylor_run (exponentiator.hpp:87-210)
//   ref_inner_prod_<sfx>       -> util::inner_prod                                 (util/linear_algebra.hpp:30-51)
//   ref_norm_<sfx>             -> util::norm                                       (util/linear_algebra.hpp:57-60)
//   ref_schmidt_orth_<sfx>     -> util::schmidt_orth                               (util/linear_algebra.hpp:133-144)
//   ref_tridiag_eigenpairs_<r> -> tridiagonal_impl::tridiagonal_eigenpairs         (lambda_lanczos_tridiagonal_impl.hpp:291-343)
// The operator handed to the reference is a "sample-style" CSR mv_mul lambda (cf. src/samples/sample2_sparse.cpp:43-47):
// out += A*in, rows optionally spread over OpenMP threads (the reference's own vector kernels stay single-threaded).
#include <chrono>
#include <complex>
#include <cstdint>
#include <cstring>
#include <vector>

#include <exponentiator.hpp>
#include <lambda_lanczos.hpp>
#include <lambda_lanczos_tridiagonal_impl.hpp>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

using clk = std::chrono::steady_clock;

template <typename T>
struct CsrView {
  int64_t n;
  const int64_t* rowptr;
  const int32_t* colidx;
  const T* vals;
  int threads;
};

template <typename T>
inline void csr_accumulate(const CsrView<T>& A, const std::vector<T>& in, std::vector<T>& out) {
  const int64_t n = A.n;
#pragma omp parallel for schedule(static) num_threads(A.threads) if (A.threads > 1)
  for (int64_t i = 0; i < n; ++i) {
    T s = T();
    for (int64_t p = A.rowptr[i]; p < A.rowptr[i + 1]; ++p) s += A.vals[p] * in[A.colidx[p]];
    out[i] += s;
  }
}

template <typename T>
using real_of = lambda_lanczos::util::real_t<T>;

template <typename T>
int lanczos_run(int64_t n, const int64_t* rowptr, const int32_t* colidx, const T* vals, int mv_threads,
                int find_max, int64_t num_eigs, real_of<T> offset, real_of<T> eps, int64_t max_iter, int64_t nepi,
                const T* init, real_of<T>* evals_out, T* evecs_out, int64_t* iter_counts, int64_t max_runs,
                int64_t* n_runs, int64_t* n_found, double* timing, int64_t cap_k, T* cap_basis,
                int64_t* cap_count) {
  CsrView<T> A{n, rowptr, colidx, vals, mv_threads < 1 ? 1 : mv_threads};
  double mv_seconds = 0.0;
  int64_t seen = 0;      // mv_mul calls in the current Lanczos run
  int64_t captured = 0;  // vectors stored from the FIRST Lanczos run
  int64_t run_index = -1;

  auto mv = [&](const std::vector<T>& in, std::vector<T>& out) {
    if (cap_basis && run_index == 0 && seen < cap_k) {
      std::memcpy(cap_basis + seen * n, in.data(), sizeof(T) * (size_t)n);
      captured = seen + 1;
    }
    ++seen;
    auto t0 = clk::now();
    csr_accumulate(A, in, out);
    mv_seconds += std::chrono::duration<double>(clk::now() - t0).count();
  };

  lambda_lanczos::LambdaLanczos<T> engine(mv, (size_t)n, find_max != 0, (size_t)num_eigs);
  if (init) {
    engine.init_vector = [&](std::vector<T>& v) {
      ++run_index;  // the reference asks for a fresh start vector once per Lanczos run
      seen = 0;
      std::memcpy(v.data(), init, sizeof(T) * (size_t)n);
    };
  } else {
    auto dflt = engine.init_vector;
    engine.init_vector = [&, dflt](std::vector<T>& v) {
      ++run_index;
      seen = 0;
      dflt(v);
    };
  }
  engine.eigenvalue_offset = offset;
  if (eps > 0) engine.eps = eps;
  if (max_iter > 0) engine.max_iteration = (size_t)max_iter;
  if (nepi > 0) engine.num_eigs_per_iteration = (size_t)nepi;

  std::vector<real_of<T>> evals;
  std::vector<std::vector<T>> evecs;
  auto t0 = clk::now();
  engine.run(evals, evecs);
  double total = std::chrono::duration<double>(clk::now() - t0).count();

  const auto& counts = engine.getIterationCounts();
  if (n_runs) *n_runs = (int64_t)counts.size();
  for (size_t i = 0; i < counts.size() && (int64_t)i < max_runs; ++i) iter_counts[i] = (int64_t)counts[i];
  if (n_found) *n_found = (int64_t)evals.size();
  for (size_t i = 0; i < evals.size() && (int64_t)i < num_eigs; ++i) {
    evals_out[i] = evals[i];
    if (evecs_out) std::memcpy(evecs_out + i * (size_t)n, evecs[i].data(), sizeof(T) * (size_t)n);
  }
  if (timing) {
    timing[0] = total;
    timing[1] = mv_seconds;
  }
  if (cap_count) *cap_count = captured;
  return 0;
}

// A by-value `Iterable` (lambda_lanczos.hpp:216-220 copies its argument) that only refers to the locked vectors.
template <typename T>
struct LockedView {
  const std::vector<std::vector<T>>* v;
  typename std::vector<std::vector<T>>::const_iterator cbegin() const { return v->cbegin(); }
  typename std::vector<std::vector<T>>::const_iterator cend() const { return v->cend(); }
};

// One Lanczos run through the reference's public run_iteration with `n_locked` vectors to deflate against.
//   locked_ptrs != null : n_locked host vectors of n elements (copied into std::vectors, the type the reference takes)
//   locked_ptrs == null : vector j = `locked_blocks` restricted to rows [floor(j n / q), floor((j+1) n / q)), zero
//                         elsewhere (disjoint supports => exactly orthogonal; the caller normalises each block) — a
//                         deflation set of any size q for n extra values of input.
// Spy outputs (each may be null): alpha_out[i] = Re<in_i, A in_i> + offset (what lambda_lanczos.hpp:248 pushes, bit for
// bit when offset == 0), beta_out[i] = Re<in_{i+1}, A in_i> (= ||u_{i+1}|| before normalisation, :262, up to rounding),
// t_mv[i] = seconds since the call began at the entry of the i-th mv_mul call, t_mv[calls] = at return of the last one;
// so t_mv[i+1] - t_mv[i] is the wall time of one complete Lanczos iteration.  Returns the iteration count.
template <typename T>
int64_t run_iteration_spy(int64_t n, const int64_t* rowptr, const int32_t* colidx, const T* vals, int mv_threads,
                          int find_max, real_of<T> offset, real_of<T> eps, int64_t max_iter, int64_t nroot,
                          const T* init, int64_t n_locked, const T* const* locked_ptrs, const T* locked_blocks,
                          real_of<T>* evals_out, T* evecs_out, int64_t* n_vals, double* alpha_out, double* beta_out,
                          double* t_mv, int64_t* mv_calls, int64_t cap_k, T* cap_basis, int64_t* cap_count) {
  CsrView<T> A{n, rowptr, colidx, vals, mv_threads < 1 ? 1 : mv_threads};
  std::vector<std::vector<T>> locked;
  locked.reserve((size_t)n_locked);
  for (int64_t j = 0; j < n_locked; ++j) {
    if (locked_ptrs) {
      locked.emplace_back(locked_ptrs[j], locked_ptrs[j] + n);
    } else {
      locked.emplace_back((size_t)n);
      const int64_t lo = (int64_t)((__int128)j * n / n_locked), hi = (int64_t)((__int128)(j + 1) * n / n_locked);
      std::memcpy(locked.back().data() + lo, locked_blocks + lo, sizeof(T) * (size_t)(hi - lo));
    }
  }
  int64_t calls = 0;
  std::vector<T> prev_au;
  const auto t_begin = clk::now();
  auto since = [&]() { return std::chrono::duration<double>(clk::now() - t_begin).count(); };
  auto mv = [&](const std::vector<T>& in, std::vector<T>& out) {
    if (t_mv) t_mv[calls] = since();
    if (cap_basis && calls < cap_k) std::memcpy(cap_basis + calls * n, in.data(), sizeof(T) * (size_t)n);
    if (beta_out && calls > 0) beta_out[calls - 1] = (double)std::real(lambda_lanczos::util::inner_prod(in, prev_au));
    csr_accumulate(A, in, out);
    if (alpha_out) alpha_out[calls] = (double)std::real(lambda_lanczos::util::inner_prod(in, out)) + (double)offset;
    if (beta_out) prev_au = out;
    ++calls;
    if (t_mv) t_mv[calls] = since();
  };
  lambda_lanczos::LambdaLanczos<T> engine(mv, (size_t)n, find_max != 0, (size_t)1);
  engine.init_vector = [&](std::vector<T>& v) { std::memcpy(v.data(), init, sizeof(T) * (size_t)n); };
  engine.eigenvalue_offset = offset;
  if (eps > 0) engine.eps = eps;
  if (max_iter > 0) engine.max_iteration = (size_t)max_iter;
  std::vector<real_of<T>> evals;
  std::vector<std::vector<T>> evecs;
  const size_t it = engine.run_iteration(evals, evecs, (size_t)nroot, LockedView<T>{&locked});
  if (n_vals) *n_vals = (int64_t)evals.size();
  for (size_t i = 0; i < evals.size(); ++i) {
    if (evals_out) evals_out[i] = evals[i];
    if (evecs_out) std::memcpy(evecs_out + i * (size_t)n, evecs[i].data(), sizeof(T) * (size_t)n);
  }
  if (mv_calls) *mv_calls = calls;
  if (cap_count) *cap_count = calls < cap_k ? calls : cap_k;
  return (int64_t)it;
}

template <typename T>
int64_t expm_run(int64_t n, const int64_t* rowptr, const int32_t* colidx, const T* vals, int mv_threads, T a,
                 const T* input, T* output, real_of<T> eps, int full_orth, int64_t max_iter, int taylor,
                 double* timing) {
  CsrView<T> A{n, rowptr, colidx, vals, mv_threads < 1 ? 1 : mv_threads};
  double mv_seconds = 0.0;
  auto mv = [&](const std::vector<T>& in, std::vector<T>& out) {
    auto t0 = clk::now();
    csr_accumulate(A, in, out);
    mv_seconds += std::chrono::duration<double>(clk::now() - t0).count();
  };
  lambda_lanczos::Exponentiator<T> ex(mv, (size_t)n);
  if (eps > 0) ex.eps = eps;
  if (max_iter > 0) ex.max_iteration = (size_t)max_iter;
  ex.full_orthogonalize = full_orth != 0;
  std::vector<T> in(input, input + n);
  std::vector<T> out;  // left unsized on purpose, as exponentiator_test.cpp:127 does
  auto t0 = clk::now();
  size_t it = taylor ? ex.taylor_run(a, in, out) : ex.run(a, in, out);
  double total = std::chrono::duration<double>(clk::now() - t0).count();
  std::memcpy(output, out.data(), sizeof(T) * (size_t)n);
  if (timing) {
    timing[0] = total;
    timing[1] = mv_seconds;
  }
  return (int64_t)it;
}

template <typename T>
void schmidt(int64_t n, int64_t nvec, const T* basis, T* uorth) {
  std::vector<std::vector<T>> us;
  for (int64_t k = 0; k < nvec; ++k) us.emplace_back(basis + k * n, basis + (k + 1) * n);
  std::vector<T> v(uorth, uorth + n);
  lambda_lanczos::util::schmidt_orth(v, us.begin(), us.end());
  std::memcpy(uorth, v.data(), sizeof(T) * (size_t)n);
}

template <typename R>
int64_t tridiag(int64_t m, const R* alpha, const R* beta, int64_t nbeta, R* evals, R* evecs) {
  std::vector<R> a(alpha, alpha + m), b(beta, beta + nbeta), ev;
  std::vector<std::vector<R>> q;
  size_t unconv = lambda_lanczos::tridiagonal_impl::tridiagonal_eigenpairs(a, b, ev, q, evecs != nullptr);
  for (int64_t i = 0; i < m; ++i) evals[i] = ev[i];
  if (evecs)
    for (int64_t i = 0; i < m; ++i)
      for (int64_t j = 0; j < m; ++j) evecs[i * m + j] = q[i][j];
  return (int64_t)unconv;
}

using cd = std::complex<double>;
using cf = std::complex<float>;

}  // namespace

extern "C" {

int ref_host_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

#define LLZ_REF_LANCZOS(SFX, T, R)                                                                                 \
  int ref_lanczos_run_##SFX(int64_t n, const int64_t* rowptr, const int32_t* colidx, const void* vals,              \
                            int mv_threads, int find_max, int64_t num_eigs, double offset, double eps,             \
                            int64_t max_iter, int64_t nepi, const void* init, void* evals_out, void* evecs_out,    \
                            int64_t* iter_counts, int64_t max_runs, int64_t* n_runs, int64_t* n_found,             \
                            double* timing, int64_t cap_k, void* cap_basis, int64_t* cap_count) {                  \
    return lanczos_run<T>(n, rowptr, colidx, (const T*)vals, mv_threads, find_max, num_eigs, (R)offset, (R)eps,    \
                          max_iter, nepi, (const T*)init, (R*)evals_out, (T*)evecs_out, iter_counts, max_runs,     \
                          n_runs, n_found, timing, cap_k, (T*)cap_basis, cap_count);                               \
  }

LLZ_REF_LANCZOS(f32, float, float)
LLZ_REF_LANCZOS(f64, double, double)
LLZ_REF_LANCZOS(c128, cd, double)
LLZ_REF_LANCZOS(c64, cf, float)

#define LLZ_REF_RUN_ITERATION(SFX, T, R)                                                                            \
  int64_t ref_run_iteration_##SFX(int64_t n, const int64_t* rowptr, const int32_t* colidx, const void* vals,        \
                                  int mv_threads, int find_max, double offset, double eps, int64_t max_iter,        \
                                  int64_t nroot, const void* init, int64_t n_locked, const void* const* locked_ptrs, \
                                  const void* locked_blocks, void* evals_out, void* evecs_out, int64_t* n_vals,     \
                                  double* alpha_out, double* beta_out, double* t_mv, int64_t* mv_calls,             \
                                  int64_t cap_k, void* cap_basis, int64_t* cap_count) {                             \
    return run_iteration_spy<T>(n, rowptr, colidx, (const T*)vals, mv_threads, find_max, (R)offset, (R)eps,         \
                                max_iter, nroot, (const T*)init, n_locked, (const T* const*)locked_ptrs,            \
                                (const T*)locked_blocks, (R*)evals_out, (T*)evecs_out, n_vals, alpha_out, beta_out, \
                                t_mv, mv_calls, cap_k, (T*)cap_basis, cap_count);                                   \
  }

LLZ_REF_RUN_ITERATION(f32, float, float)
LLZ_REF_RUN_ITERATION(f64, double, double)
LLZ_REF_RUN_ITERATION(c128, cd, double)
LLZ_REF_RUN_ITERATION(c64, cf, float)

int64_t ref_expm_run_f32(int64_t n, const int64_t* rowptr, const int32_t* colidx, const void* vals, int mv_threads,
                         double a_re, double a_im, const void* input, void* output, double eps, int full_orth,
                         int64_t max_iter, int taylor, double* timing) {
  (void)a_im;
  return expm_run<float>(n, rowptr, colidx, (const float*)vals, mv_threads, (float)a_re, (const float*)input,
                         (float*)output, (float)eps, full_orth, max_iter, taylor, timing);
}
int64_t ref_expm_run_f64(int64_t n, const int64_t* rowptr, const int32_t* colidx, const void* vals, int mv_threads,
                         double a_re, double a_im, const void* input, void* output, double eps, int full_orth,
                         int64_t max_iter, int taylor, double* timing) {
  (void)a_im;
  return expm_run<double>(n, rowptr, colidx, (const double*)vals, mv_threads, a_re, (const double*)input,
                          (double*)output, eps, full_orth, max_iter, taylor, timing);
}
int64_t ref_expm_run_c128(int64_t n, const int64_t* rowptr, const int32_t* colidx, const void* vals, int mv_threads,
                          double a_re, double a_im, const void* input, void* output, double eps, int full_orth,
                          int64_t max_iter, int taylor, double* timing) {
  return expm_run<cd>(n, rowptr, colidx, (const cd*)vals, mv_threads, cd(a_re, a_im), (const cd*)input, (cd*)output,
                      eps, full_orth, max_iter, taylor, timing);
}

int64_t ref_expm_run_c64(int64_t n, const int64_t* rowptr, const int32_t* colidx, const void* vals, int mv_threads,
                         double a_re, double a_im, const void* input, void* output, double eps, int full_orth,
                         int64_t max_iter, int taylor, double* timing) {
  return expm_run<cf>(n, rowptr, colidx, (const cf*)vals, mv_threads, cf((float)a_re, (float)a_im), (const cf*)input,
                      (cf*)output, (float)eps, full_orth, max_iter, taylor, timing);
}

#define LLZ_REF_BLAS1(SFX, T, R)                                                                     \
  void ref_inner_prod_##SFX(int64_t n, const void* a, const void* b, void* out) {                     \
    std::vector<T> va((const T*)a, (const T*)a + n), vb((const T*)b, (const T*)b + n);                \
    *(T*)out = lambda_lanczos::util::inner_prod(va, vb);                                              \
  }                                                                                                  \
  double ref_norm_##SFX(int64_t n, const void* a) {                                                  \
    std::vector<T> va((const T*)a, (const T*)a + n);                                                  \
    return (double)lambda_lanczos::util::norm(va);                                                    \
  }                                                                                                  \
  void ref_normalize_##SFX(int64_t n, void* a) {                                                     \
    std::vector<T> va((T*)a, (T*)a + n);                                                              \
    lambda_lanczos::util::normalize(va);                                                              \
    std::memcpy(a, va.data(), sizeof(T) * (size_t)n);                                                 \
  }                                                                                                  \
  void ref_schmidt_orth_##SFX(int64_t n, int64_t nvec, const void* basis, void* uorth) {             \
    schmidt<T>(n, nvec, (const T*)basis, (T*)uorth);                                                  \
  }

LLZ_REF_BLAS1(f32, float, float)
LLZ_REF_BLAS1(f64, double, double)
LLZ_REF_BLAS1(c128, cd, double)
LLZ_REF_BLAS1(c64, cf, float)

int64_t ref_tridiag_eigenpairs_f64(int64_t m, const double* alpha, const double* beta, int64_t nbeta, double* evals,
                                   double* evecs) {
  return tridiag<double>(m, alpha, beta, nbeta, evals, evecs);
}
int64_t ref_tridiag_eigenpairs_f32(int64_t m, const float* alpha, const float* beta, int64_t nbeta, float* evals,
                                   float* evecs) {
  return tridiag<float>(m, alpha, beta, nbeta, evals, evecs);
}

}  // extern "C"
