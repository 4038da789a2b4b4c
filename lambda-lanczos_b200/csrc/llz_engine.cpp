// llz_engine.cpp — whole-engine C entry points (llz_eigs_run, llz_expm_run): the header-only host engine
// (lambda_lanczos_b200/lambda_lanczos.hpp, exponentiator.hpp) instantiated for the four scalar types, for callers
// that cannot include C++ templates (ctypes, cgo, JNI ...).  Plain host C++: no kernels in this file.
#include <complex>
#include <cstring>
#include <exception>

#include "lambda_lanczos_b200/exponentiator.hpp"
#include "lambda_lanczos_b200/lambda_lanczos.hpp"
#include "llz_internal.hpp"

namespace {

using namespace lambda_lanczos_b200;

template <typename T>
int eigs_run(llz_ctx_t ctx, llz_op_t op, const llz_eigs_params_t* p, const void* start, double* evals_out, void* evecs_out,
             int64_t* n_found, int64_t* iter_counts, int64_t max_runs, int64_t* n_runs, llz_run_stats_t* stats) {
  using R = util::real_t<T>;
  Context c = Context::borrow(ctx);
  DeviceOperator<T> A = DeviceOperator<T>::borrow(c, op);
  const size_t n = A.rows();  // local block
  LambdaLanczos<T> engine(A, A.global_rows(), p->find_maximum != 0, (size_t)p->num_eigs);
  engine.eigenvalue_offset = (R)p->eigenvalue_offset;
  if (p->eps > 0) engine.eps = (R)p->eps;
  if (p->max_iteration > 0) engine.max_iteration = (size_t)p->max_iteration;
  if (p->num_eigs_per_iteration > 0) engine.num_eigs_per_iteration = (size_t)p->num_eigs_per_iteration;
  engine.orthogonalization = p->orth;
  engine.pipeline_depth = p->pipeline_depth;
  engine.ritz_solver = p->ritz_solver;
  if (start) {
    engine.start_local = static_cast<const T*>(start);
  } else {
    engine.init_vector = [](std::vector<T>& v) {  // seeded, so that C-ABI callers get reproducible runs by default
      std::mt19937 gen(1);
      std::uniform_real_distribution<R> dist(R(-1), R(1));
      R* raw = reinterpret_cast<R*>(v.data());
      for (size_t i = 0; i < v.size() * (sizeof(T) / sizeof(R)); ++i) raw[i] = dist(gen);
    };
  }
  std::vector<R> evals;
  std::vector<DeviceVector<T>> evecs;
  engine.run_device(evals, evecs);
  if (n_found) *n_found = (int64_t)evals.size();
  for (size_t i = 0; i < evals.size() && (int64_t)i < p->num_eigs; ++i) {
    evals_out[i] = (double)evals[i];
    if (evecs_out) evecs[i].download(static_cast<T*>(evecs_out) + i * n);
  }
  const auto& counts = engine.getIterationCounts();
  if (n_runs) *n_runs = (int64_t)counts.size();
  for (size_t i = 0; i < counts.size() && (int64_t)i < max_runs; ++i) iter_counts[i] = (int64_t)counts[i];
  if (stats) {
    const RunStatistics& s = engine.statistics();
    stats->seconds_total = s.seconds_total;
    stats->seconds_host = s.seconds_host;
    stats->iterations = (int64_t)s.iterations;
    stats->runs = (int64_t)s.runs;
    stats->basis_bytes = 0;
    stats->kernel_launches = s.kernel_launches;
  }
  return LLZ_OK;
}

template <typename T>
int expm_run(llz_ctx_t ctx, llz_op_t op, const double a[2], const void* input, void* output, int host, double eps,
             int full_orth, int64_t max_iteration, int taylor, int64_t* iterations) {
  using R = util::real_t<T>;
  Context c = Context::borrow(ctx);
  DeviceOperator<T> A = DeviceOperator<T>::borrow(c, op);
  const size_t n = A.rows();  // local block
  Exponentiator<T> ex(A, A.global_rows());
  if (eps > 0) ex.eps = (R)eps;
  if (max_iteration > 0) ex.max_iteration = (size_t)max_iteration;
  ex.full_orthogonalize = full_orth != 0;
  T av;
  {
    R parts[2] = {(R)a[0], (R)a[1]};
    std::memcpy(&av, parts, sizeof(T));  // T is R or std::complex<R>
  }
  DeviceVector<T> in(c, n), out(c, n);
  if (host) {
    in.upload(static_cast<const T*>(input));
  } else {
    if (cudaMemcpyAsync(in.device_ptr(), input, sizeof(T) * n, cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess)
      return llz::fail(LLZ_ERR_CUDA, "expm: device input copy failed");
  }
  const size_t it = taylor ? ex.taylor_run_device(av, in, out) : ex.run_device(av, in, out);
  if (host) {
    out.download(static_cast<T*>(output));
  } else {
    if (cudaMemcpyAsync(output, out.device_ptr(), sizeof(T) * n, cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess)
      return llz::fail(LLZ_ERR_CUDA, "expm: device output copy failed");
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return llz::fail(LLZ_ERR_CUDA, "expm: sync failed");
  }
  if (iterations) *iterations = (int64_t)it;
  return LLZ_OK;
}

template <typename F>
int guarded(F&& f) {
  try {
    return f();
  } catch (const Error& e) {
    return llz::fail(e.status(), "%s", e.what());
  } catch (const std::bad_alloc&) {
    return llz::fail(LLZ_ERR_OOM, "host allocation failed");
  } catch (const std::exception& e) {
    return llz::fail(LLZ_ERR_INVALID, "%s", e.what());
  }
}

}  // namespace

extern "C" {

int llz_eigs_run(llz_ctx_t ctx, llz_op_t op, int dtype, const llz_eigs_params_t* params, const void* start,
                 double* eigenvalues_out, void* eigenvectors_out, int64_t* n_found, int64_t* iter_counts,
                 int64_t max_runs, int64_t* n_runs, llz_run_stats_t* stats) {
  if (!ctx || !op || !params || !eigenvalues_out || params->num_eigs < 1 || (max_runs > 0 && !iter_counts))
    return llz::fail(LLZ_ERR_INVALID, "llz_eigs_run: bad argument");
  if (op->impl->dtype != dtype) return llz::fail(LLZ_ERR_INVALID, "llz_eigs_run: operator dtype %d != %d", op->impl->dtype, dtype);
  return guarded([&]() -> int {
    switch (dtype) {
      case LLZ_F32:
        return eigs_run<float>(ctx, op, params, start, eigenvalues_out, eigenvectors_out, n_found, iter_counts, max_runs, n_runs, stats);
      case LLZ_F64:
        return eigs_run<double>(ctx, op, params, start, eigenvalues_out, eigenvectors_out, n_found, iter_counts, max_runs, n_runs, stats);
      case LLZ_C64:
        return eigs_run<std::complex<float>>(ctx, op, params, start, eigenvalues_out, eigenvectors_out, n_found, iter_counts, max_runs, n_runs, stats);
      case LLZ_C128:
        return eigs_run<std::complex<double>>(ctx, op, params, start, eigenvalues_out, eigenvectors_out, n_found, iter_counts, max_runs, n_runs, stats);
    }
    return llz::fail(LLZ_ERR_INVALID, "unknown dtype %d", dtype);
  });
}

int llz_expm_run(llz_ctx_t ctx, llz_op_t op, int dtype, const double a[2], const void* input, void* output, int host,
                 double eps, int full_orthogonalize, int64_t max_iteration, int taylor, int64_t* iterations) {
  if (!ctx || !op || !a || !input || !output) return llz::fail(LLZ_ERR_INVALID, "llz_expm_run: bad argument");
  if (op->impl->dtype != dtype) return llz::fail(LLZ_ERR_INVALID, "llz_expm_run: operator dtype %d != %d", op->impl->dtype, dtype);
  return guarded([&]() -> int {
    switch (dtype) {
      case LLZ_F32: return expm_run<float>(ctx, op, a, input, output, host, eps, full_orthogonalize, max_iteration, taylor, iterations);
      case LLZ_F64: return expm_run<double>(ctx, op, a, input, output, host, eps, full_orthogonalize, max_iteration, taylor, iterations);
      case LLZ_C64:
        return expm_run<std::complex<float>>(ctx, op, a, input, output, host, eps, full_orthogonalize, max_iteration, taylor, iterations);
      case LLZ_C128:
        return expm_run<std::complex<double>>(ctx, op, a, input, output, host, eps, full_orthogonalize, max_iteration, taylor, iterations);
    }
    return llz::fail(LLZ_ERR_INVALID, "unknown dtype %d", dtype);
  });
}

}  // extern "C"
