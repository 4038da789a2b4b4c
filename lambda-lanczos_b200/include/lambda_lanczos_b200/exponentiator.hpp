// exponentiator.hpp — Exponentiator<T>: the reference's Krylov exponentiator (exponentiator.hpp:24-211) rebuilt for
// B200.  output = exp(a A) input through the Lanczos basis of K_k(A, input); the basis, the matvec, the recurrence and
// the final linear combination run on the GPU, exp(a T_k) on the host (k is a few dozen).
#pragma once
#include <algorithm>
#include <cmath>
#include <complex>
#include <limits>
#include <vector>

#include "common.hpp"
#include "device_operator.hpp"
#include "lambda_lanczos.hpp"
#include "tridiagonal.hpp"

namespace lambda_lanczos_b200 {

template <typename T>
class Exponentiator {
 private:
  template <typename n_type>
  using real_t = util::real_t<n_type>;
  using R = real_t<T>;

 public:
  // ---- the reference's public surface (exponentiator.hpp:41-71) ----
  DeviceOperator<T> mv_mul;
  size_t matrix_size;
  size_t max_iteration;
  real_t<T> eps = std::numeric_limits<real_t<T>>::epsilon() * real_t<T>(1e2);
  bool full_orthogonalize = false;
  size_t initial_vector_size = 200;

  // ---- engine knob: iterations the GPU may run ahead of the host overlap test (< 0: auto_pipeline_depth) ----
  int pipeline_depth = -1;

  Exponentiator(DeviceOperator<T> mv_mul, size_t matrix_size)
      : mv_mul(std::move(mv_mul)), matrix_size(matrix_size), max_iteration(matrix_size) {}
  // The reference's constructor, verbatim (exponentiator.hpp:80-82): a HOST mv_mul, see LambdaLanczos.
  Exponentiator(std::function<void(const std::vector<T>&, std::vector<T>&)> host_mv_mul, size_t matrix_size)
      : Exponentiator(DeviceOperator<T>::host_function(Context::default_context(), matrix_size, std::move(host_mv_mul)), matrix_size) {}

  // exponentiator.hpp:87-173 with device-resident input/output (no PCIe traffic per call: what a time-evolution loop
  // that feeds output back as input wants).  `output` may alias `input`.
  size_t run_device(const T& a, const DeviceVector<T>& input, DeviceVector<T>& output) const {
    const Context& ctx = mv_mul.context();
    const size_t n = mv_mul.rows();  // local block of a row-sharded run; == matrix_size on a single GPU
    if (mv_mul.global_rows() != matrix_size) throw Error(LLZ_ERR_INVALID, "Exponentiator: mv_mul does not match matrix_size");
    if (input.size() != n) throw Error(LLZ_ERR_INVALID, "Exponentiator: input size differs from the operator's (local) rows");
    if (!output.valid() || output.size() != n) output = DeviceVector<T>(ctx, n);
    const size_t want_cols = std::max<size_t>(2, max_iteration + 1);
    if (!work_.matches(util::dtype_of<T>::value, n, want_cols)) work_ = KrylovWorkspace(ctx, util::dtype_of<T>::value, n, want_cols);
    llz_krylov_t kry = work_.get();
    const size_t capacity = work_.capacity();
    check(llz_krylov_set_locked(kry, nullptr, 0), "llz_krylov_set_locked");

    double input_norm = 0.0;  // :100-101 and :165
    check(llz_krylov_begin(kry, input.device_ptr(), 0, &input_norm), "llz_krylov_begin");

    std::vector<double> alpha, beta;
    std::vector<T> coeff_prev;
    // Without reorthogonalisation nothing but the next two iterations reads a Lanczos vector before the final sum, so on
    // one GPU the vectors stay un-normalised (column l >= 1 holds beta_{l-1} u_l: no normalisation pass, 2 n s bytes per
    // iteration less) and the sum below divides its coefficients by those factors.
    const bool lazy = !full_orthogonalize && ctx.nranks() == 1;
    const int orth = full_orthogonalize ? LLZ_ORTH_FULL : (lazy ? LLZ_ORTH_RECURRENCE_LAZY : LLZ_ORTH_RECURRENCE);
    const double beta_threshold = (double)std::numeric_limits<R>::epsilon();  // :154
    size_t itern = max_iteration;
    size_t enqueued = 0;
    const size_t depth = pipeline_depth < 0 ? (size_t)auto_pipeline_depth(matrix_size * sizeof(T) / (size_t)ctx.nranks(), false) : (size_t)pipeline_depth;
    for (size_t k = 1; k <= max_iteration; ++k) {
      const size_t ahead = std::min(max_iteration, k + depth);
      while (enqueued < ahead) {
        if (enqueued + 2 > capacity) {
          if (enqueued >= k) break;
          throw Error(LLZ_ERR_OOM, "Exponentiator: the Krylov basis is full; lower max_iteration");
        }
        check(llz_krylov_step(kry, mv_mul.get(), 0.0, orth), "llz_krylov_step");
        ++enqueued;
      }
      double ak = 0, bk = 0, wn = 0;
      check(llz_krylov_fetch(kry, (int64_t)k, &ak, &bk, &wn), "llz_krylov_fetch");
      for (int pass = 0; pass < 3 && full_orthogonalize && bk >= beta_threshold && bk < 0.5 * wn; ++pass) {  // DGKS, see LambdaLanczos
        double shrink = 1.0;
        check(llz_krylov_refine(kry, (int64_t)k, &shrink), "llz_krylov_refine");
        enqueued = k;
        wn = bk;
        bk *= shrink;
      }
      alpha.push_back(ak);  // :110

      // :124-133 — T_k = P diag(ev) P^T with the k-1 couplings known so far; coeff = exp(a T_k) e_0
      const size_t m = alpha.size();
      std::vector<double> ev, p;
      tridiagonal::implicit_ql(alpha.data(), beta.data(), m, ev, &p);  // p[j*m + i] = component i of vector j
      std::vector<T> coeff(m, T(0));
      for (size_t j = 0; j < m; ++j) {
        const T w = std::exp(a * T((R)ev[j])) * T((R)p[j * m + 0]);
        for (size_t i = 0; i < m; ++i) coeff[i] += T((R)p[j * m + i]) * w;
      }
      beta.push_back(bk);  // :145

      T overlap = T(0);  // :147-152
      for (size_t i = 0; i < coeff_prev.size(); ++i) overlap += util::typed_conj(coeff_prev[i]) * coeff[i];
      coeff_prev = std::move(coeff);

      if (std::abs(R(1) - std::abs(overlap)) < eps || beta.back() < beta_threshold) {  // :154-158
        itern = k;
        break;
      }
    }

    // :163-170 — output = ||input|| * sum_l coeff_prev[l] u_l
    std::vector<T> scaled(coeff_prev.size());
    for (size_t l = 0; l < coeff_prev.size(); ++l) {
      scaled[l] = T((R)input_norm) * coeff_prev[l];
      if (lazy && l >= 1) scaled[l] /= T((R)beta[l - 1]);
    }
    llz_vec_t out = output.get();
    check(llz_krylov_combine(kry, (int64_t)scaled.size(), 1, scaled.data(), 0, &out), "llz_krylov_combine");
    last_iterations_ = itern;
    return itern;
  }

  // exponentiator.hpp:87 — host vectors in, host vector out (resized like the reference does, :163)
  size_t run(const T& a, const std::vector<T>& input, std::vector<T>& output) const {
    // (row-sharded runs: `input` / `output` are this rank's row block)
    const Context& ctx = mv_mul.context();
    const size_t n = mv_mul.rows();
    if (input.size() != n) throw Error(LLZ_ERR_INVALID, "Exponentiator: input size differs from matrix_size");
    DeviceVector<T> in(ctx, n), out(ctx, n);
    in.upload(input);
    const size_t it = run_device(a, in, out);
    output.resize(n);
    out.download(output.data());
    return it;
  }

  // exponentiator.hpp:175-210 — plain Taylor series summed backwards (kept as the reference keeps it: a cross-check)
  size_t taylor_run_device(const T& a, const DeviceVector<T>& input, DeviceVector<T>& output) {
    const Context& ctx = mv_mul.context();
    const size_t n = mv_mul.rows();
    if (!output.valid() || output.size() != n) output = DeviceVector<T>(ctx, n);
    if (a == T()) {  // :179-182
      check(llz_vec_copy(output.get(), input.get()), "llz_vec_copy");
      return 1;
    }
    std::vector<DeviceVector<T>> terms;
    terms.push_back(input);
    T factor = T(1);
    for (size_t k = 1;; ++k) {  // :188-196
      factor *= a / T((R)k);
      terms.emplace_back(ctx, n);
      check(llz_op_apply(mv_mul.get(), terms[k - 1].get(), terms[k].get()), "llz_op_apply");
      double nrm = 0.0;
      check(llz_vec_norm(terms[k].get(), &nrm), "llz_vec_norm");
      if ((R)nrm * std::abs(factor) < eps) break;
    }
    DeviceVector<T> acc(ctx, n);
    check(llz_vec_fill_zero(acc.get()), "llz_vec_fill_zero");
    for (size_t k = terms.size(); k-- > 0;) {  // :199-207
      double f[2];
      util::to_pair(factor, f);
      check(llz_vec_axpy(acc.get(), f, terms[k].get()), "llz_vec_axpy");
      factor *= T((R)k) / a;
    }
    check(llz_vec_copy(output.get(), acc.get()), "llz_vec_copy");
    return terms.size();
  }

  size_t taylor_run(const T& a, const std::vector<T>& input, std::vector<T>& output) {
    const Context& ctx = mv_mul.context();
    const size_t n = mv_mul.rows();
    DeviceVector<T> in(ctx, n), out(ctx, n);
    in.upload(input);
    const size_t it = taylor_run_device(a, in, out);
    output.resize(n);
    out.download(output.data());
    return it;
  }

 private:
  mutable KrylovWorkspace work_;
  mutable size_t last_iterations_ = 0;
};

}  // namespace lambda_lanczos_b200
