// lambda_lanczos.hpp — LambdaLanczos<T>: the reference's Lanczos eigen-engine (lambda_lanczos.hpp:109-415) rebuilt for
// B200.  Same constructor, public fields and run() overloads; `mv_mul` is a DeviceOperator<T>, the Lanczos basis and
// every vector kernel live on the GPU behind the C ABI (include/llz.h), and this header keeps what the north star
// leaves on the host: the control flow, the tridiagonal solves, the convergence tests and the eigenpair bookkeeping.
//
// What is deliberately different from the reference (all result-preserving within the stated tolerances):
//   * reorthogonalisation is one classical Gram-Schmidt pass over [locked, basis] after the three-term recurrence
//     (two streaming passes over HBM) instead of vector-by-vector modified Gram-Schmidt;
//   * only the `nroot` extreme Ritz values are computed per iteration (Sturm bisection) instead of the full spectrum;
//   * the GPU runs `pipeline_depth` iterations ahead of the host convergence test; basis vectors past the converged
//     iteration are simply not used (compute_eigenvectors only touches u[0..m-1], lambda_lanczos.hpp:53);
//   * eigenvectors of all requested roots are assembled in ONE pass over the basis and stay on the device; kept
//     eigenvectors ("locked" vectors for deflation) never leave HBM between Lanczos runs.
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <exception>
#include <thread>
#include <cmath>
#include <functional>
#include <limits>
#include <map>
#include <random>
#include <tuple>
#include <utility>
#include <vector>

#include "common.hpp"
#include "device_operator.hpp"
#include "tridiagonal.hpp"

namespace lambda_lanczos_b200 {

// Default start vector: uniform in [-1, 1] per real component, seeded from std::random_device — the behaviour of the
// reference's VectorRandomInitializer (lambda_lanczos.hpp:70-104).  Inject `init_vector` for reproducible runs.
template <typename T>
struct VectorRandomInitializer {
  static void init(std::vector<T>& v) {
    std::mt19937 gen{std::random_device{}()};
    std::uniform_real_distribution<T> dist(T(-1), T(1));
    for (auto& x : v) x = dist(gen);
  }
};
template <typename R>
struct VectorRandomInitializer<std::complex<R>> {
  static void init(std::vector<std::complex<R>>& v) {
    std::mt19937 gen{std::random_device{}()};
    std::uniform_real_distribution<R> dist(R(-1), R(1));
    for (auto& x : v) {
      const R re = dist(gen);
      const R im = dist(gen);
      x = std::complex<R>(re, im);
    }
  }
};

// RAII handle of the device-resident Krylov workspace.
class KrylovWorkspace {
 public:
  KrylovWorkspace() {}
  KrylovWorkspace(const Context& ctx, int dtype, size_t n, size_t max_cols) : ctx_(ctx), n_(n), max_cols_(max_cols), dtype_(dtype) {
    llz_krylov_t k = nullptr;
    check(llz_krylov_create(ctx.get(), dtype, (int64_t)n, (int64_t)max_cols, &k), "llz_krylov_create");
    h_.reset(k, [](llz_krylov_t p) { llz_krylov_destroy(p); });
  }
  bool matches(int dtype, size_t n, size_t max_cols) const { return h_ && dtype_ == dtype && n_ == n && max_cols_ == max_cols; }
  llz_krylov_t get() const { return h_.get(); }
  size_t capacity() const {
    int64_t c = 0;
    check(llz_krylov_capacity(h_.get(), &c), "llz_krylov_capacity");
    return (size_t)c;
  }

 private:
  Context ctx_ = Context::none();  // the workspace's memory is owned by the context: keep it alive (declared first: dies last)
  std::shared_ptr<llz_krylov_s> h_;
  size_t n_ = 0, max_cols_ = 0;
  int dtype_ = -1;
};

// Keeps the best `num_eigs` eigenpairs seen so far and reports whether a batch changed the kept set — the semantics of
// the reference's EigenPairManager (eigenpair_manager.hpp:21-79), with device-resident eigenvectors.
template <typename T, typename Vec>
class EigenPairManager {
  using R = util::real_t<T>;

 public:
  EigenPairManager(bool find_maximum, size_t num_eigs)
      : num_eigs_(num_eigs), pairs_(find_maximum ? Compare([](R a, R b) { return a > b; }) : Compare([](R a, R b) { return a < b; })) {}
  size_t size() const { return pairs_.size(); }
  // Returns true when NOTHING of the batch survived (=> no further Lanczos run is needed).
  bool insertEigenpairs(std::vector<R>& values, std::vector<Vec>& vectors) {
    bool nothing_added = true;
    for (size_t i = 0; i < values.size(); ++i) {
      auto pos = pairs_.emplace(values[i], std::move(vectors[i]));
      if (pairs_.size() > num_eigs_) {
        auto worst = std::prev(pairs_.end());
        if (pos != worst) nothing_added = false;
        pairs_.erase(worst);
      } else {
        nothing_added = false;
      }
    }
    return nothing_added;
  }
  std::vector<Vec> getEigenvectors() const {
    std::vector<Vec> out;
    out.reserve(pairs_.size());
    for (const auto& p : pairs_) out.push_back(p.second);
    return out;
  }
  using Compare = std::function<bool(R, R)>;
  std::multimap<R, Vec, Compare>& getEigenpairs() { return pairs_; }

 private:
  size_t num_eigs_;
  std::multimap<R, Vec, Compare> pairs_;
};

struct RunStatistics {
  double seconds_total = 0.0;
  double seconds_host = 0.0;  // tridiagonal solves + convergence logic
  size_t iterations = 0;
  size_t runs = 0;
  uint64_t kernel_launches = 0;
  size_t refinements = 0;  // Gram-Schmidt passes repeated because of cancellation
};

template <typename T>
class LambdaLanczos {
 private:
  template <typename n_type>
  using real_t = util::real_t<n_type>;
  using R = real_t<T>;

 public:
  // ---- the reference's public surface (lambda_lanczos.hpp:120-181) ----
  DeviceOperator<T> mv_mul;
  std::function<void(std::vector<T>& vec)> init_vector = VectorRandomInitializer<T>::init;
  size_t matrix_size;
  size_t max_iteration;
  real_t<T> eps = std::numeric_limits<real_t<T>>::epsilon() * real_t<T>(1e3);
  bool find_maximum;
  size_t num_eigs = 1;
  real_t<T> eigenvalue_offset = 0.0;
  size_t num_eigs_per_iteration = 5;
  size_t initial_vector_size = 200;  // kept for source compatibility; the device store grows by itself

  // ---- engine knobs without a reference counterpart ----
  int orthogonalization = LLZ_ORTH_FULL;  // llz_orth_t
  int pipeline_depth = -1;                // iterations the GPU may run ahead of the host convergence test (< 0: auto_pipeline_depth)
  int ritz_solver = 0;                    // 0: bisection on the extreme values, 1: full implicit QL every iteration
  int host_threads = 0;                   // 1: never start the helper thread that runs the Ritz solves of short iterations
  double reorth_eta = 0.5;                // repeat the Gram-Schmidt pass when beta < reorth_eta * ||w'|| (DGKS)
  // Row-sharded runs (the context joined a group): `matrix_size` stays the GLOBAL dimension and `init_vector` is asked
  // for the whole start vector, of which this rank keeps rows [row_offset, row_offset + rows) — so a seeded
  // initializer gives the same start vector for any number of GPUs.  Set `init_vector_local` to produce only the
  // local block instead.  Either way, `start_local` (host pointer to the local block, n_local elements), when
  // non-null, replaces both and is uploaded once per run().
  std::function<void(std::vector<T>& vec)> init_vector_local;
  const T* start_local = nullptr;

  LambdaLanczos(DeviceOperator<T> mv_mul, size_t matrix_size, bool find_maximum, size_t num_eigs)
      : mv_mul(std::move(mv_mul)), matrix_size(matrix_size), max_iteration(matrix_size), find_maximum(find_maximum), num_eigs(num_eigs) {}
  // The reference's constructor, verbatim (lambda_lanczos.hpp:200-206): mv_mul is a HOST callable (out += A*in on
  // std::vectors).  It is wrapped in DeviceOperator<T>::host_function on the process-wide default context.
  LambdaLanczos(std::function<void(const std::vector<T>&, std::vector<T>&)> host_mv_mul, size_t matrix_size, bool find_maximum, size_t num_eigs)
      : LambdaLanczos(DeviceOperator<T>::host_function(Context::default_context(), matrix_size, std::move(host_mv_mul)), matrix_size,
                      find_maximum, num_eigs) {}

  // One complete Lanczos run (lambda_lanczos.hpp:217-322) with device-resident inputs and outputs.
  size_t run_iteration(std::vector<real_t<T>>& eigvalues, std::vector<DeviceVector<T>>& eigvecs, size_t nroot,
                       const std::vector<DeviceVector<T>>& orthogonalizeTo) {
    using clock = std::chrono::steady_clock;
    const Context& ctx = mv_mul.context();
    if (!mv_mul.valid() || mv_mul.global_rows() != matrix_size) throw Error(LLZ_ERR_INVALID, "LambdaLanczos: mv_mul does not match matrix_size");
    const size_t n = mv_mul.rows();  // local block; == matrix_size for a single GPU
    const size_t want_cols = std::max<size_t>(2, max_iteration + 1);
    if (!work_.matches(util::dtype_of<T>::value, n, want_cols)) work_ = KrylovWorkspace(ctx, util::dtype_of<T>::value, n, want_cols);
    llz_krylov_t kry = work_.get();
    const size_t capacity = work_.capacity();

    std::vector<llz_vec_t> locked;
    for (const auto& v : orthogonalizeTo) locked.push_back(v.get());
    check(llz_krylov_set_locked(kry, locked.data(), (int64_t)locked.size()), "llz_krylov_set_locked");

    if (start_local) {  // :232 — the start vector is the same for every Lanczos run of this run(): keep it in HBM
      if (!start_dev_.valid() || start_dev_.size() != n || start_dev_src_ != start_local) {
        start_dev_ = DeviceVector<T>(ctx, n);
        start_dev_.upload(start_local);
        start_dev_src_ = start_local;
      }
      check(llz_krylov_begin(kry, start_dev_.device_ptr(), 0, nullptr), "llz_krylov_begin");
    } else if (init_vector_local) {
      std::vector<T> start(n);
      init_vector_local(start);
      check(llz_krylov_begin(kry, start.data(), 1, nullptr), "llz_krylov_begin");
    } else {
      std::vector<T> start(matrix_size);
      init_vector(start);
      check(llz_krylov_begin(kry, start.data() + mv_mul.row_offset(), 1, nullptr), "llz_krylov_begin");
    }

    std::vector<double> alpha, beta;
    std::vector<double> evs, pevs;
    tridiagonal::ExtremeState<double> warm;
    const double zero_threshold = (double)std::numeric_limits<R>::epsilon() * 1e1;  // :279
    size_t itern = max_iteration;
    const size_t depth = pipeline_depth < 0 ? (size_t)auto_pipeline_depth(matrix_size * sizeof(T) / (size_t)ctx.nranks(), true) : (size_t)pipeline_depth;

    // The host side of iteration k once its scalars are known: Ritz values of T_k, then the reference's stopping rules.
    auto host_step = [&](size_t k, double a, double b) -> bool {
      const auto t0 = clock::now();
      alpha.push_back(a);  // :248
      beta.push_back(b);   // :262
      const size_t ncalc = std::min(nroot, alpha.size());  // :264
      if (ritz_solver == 1) {
        std::vector<double> all;
        tridiagonal::implicit_ql(alpha.data(), beta.data(), alpha.size(), all, static_cast<std::vector<double>*>(nullptr));
        evs.assign(ncalc, 0.0);
        for (size_t i = 0; i < ncalc; ++i) evs[i] = find_maximum ? all[all.size() - 1 - i] : all[i];
      } else {
        tridiagonal::extreme_eigenvalues(alpha.data(), beta.data(), alpha.size(), ncalc, find_maximum, evs, &warm);
      }
      bool stop = false;
      if (beta.back() < zero_threshold) {  // :279-283 (the new vector is left un-normalised and never used)
        itern = k;
        stop = true;
      } else {
        bool converged = pevs.size() == evs.size();  // :290-302, on the SHIFTED values
        for (size_t r = 0; converged && r < nroot; ++r) {
          const double ev = evs[r], pev = pevs[r];
          if (std::abs(ev - pev) >= std::min(std::abs(ev), std::abs(pev)) * (double)eps) converged = false;
        }
        if (converged) {
          itern = k;
          stop = true;
        } else {
          pevs = evs;
        }
      }
      stats_.seconds_host += std::chrono::duration<double>(clock::now() - t0).count();
      return stop;
    };
    // DGKS test: if the Gram-Schmidt pass removed most of the vector (near breakdown, e.g. the Krylov space of a
    // deflated run is exhausted), one classical pass leaves it non-orthogonal: repeat the pass ("twice is enough").
    // The reference's modified Gram-Schmidt does not need this; it keeps the same vector to rounding.
    auto needs_refinement = [&](double b, double wn) {
      return orthogonalization != LLZ_ORTH_RECURRENCE && b >= zero_threshold && b < reorth_eta * wn;
    };
    auto enqueue_step = [&]() { check(llz_krylov_step(kry, mv_mul.get(), (double)eigenvalue_offset, orthogonalization), "llz_krylov_step"); };
    auto basis_full = [&](size_t enqueued) {
      return Error(LLZ_ERR_OOM, "LambdaLanczos: the Krylov basis is full after " + std::to_string(enqueued) +
                                    " iterations; lower max_iteration (the algorithm stores every Lanczos vector)");
    };

    if (depth < 2 || host_threads == 1) {
      // ---- one host thread: enqueue up to `depth` iterations ahead, then wait for iteration k and test it ----
      size_t enqueued = 0;
      for (size_t k = 1; k <= max_iteration; ++k) {
        const size_t ahead = std::min(max_iteration, k + depth);
        while (enqueued < ahead) {
          if (enqueued + 2 > capacity) {
            if (enqueued >= k) break;  // only speculative work would not fit
            throw basis_full(enqueued);
          }
          enqueue_step();
          ++enqueued;
        }
        double a = 0, b = 0, wn = 0;
        check(llz_krylov_fetch(kry, (int64_t)k, &a, &b, &wn), "llz_krylov_fetch");
        for (int pass = 0; pass < 3 && needs_refinement(b, wn); ++pass) {
          double shrink = 1.0;
          check(llz_krylov_refine(kry, (int64_t)k, &shrink), "llz_krylov_refine");
          enqueued = k;
          wn = b;
          b *= shrink;
          ++stats_.refinements;
        }
        if (host_step(k, a, b)) break;
      }
    } else {
      // ---- two host threads (short iterations: the Ritz solve of T_k costs as much as the GPU's iteration).  This
      //      thread only launches, up to `depth` iterations ahead of the last TESTED one; a helper waits for each
      //      iteration's scalars (it spins on the pinned flag the GPU writes), runs the Ritz solve and the stopping
      //      rules, and asks this thread for the rare DGKS refinement, the only other call that touches the workspace.
      std::atomic<size_t> enqueued{0}, tested{0}, refine_request{0};
      std::atomic<bool> refine_done{false}, finished{false}, abort_helper{false};
      double refine_shrink = 1.0;
      std::exception_ptr helper_error;
      std::thread helper([&]() {
        try {
          for (size_t k = 1; k <= max_iteration; ++k) {
            while (enqueued.load(std::memory_order_acquire) < k) {
              if (abort_helper.load(std::memory_order_acquire)) return;
              detail::cpu_relax();
            }
            double a = 0, b = 0, wn = 0;
            check(llz_krylov_fetch(kry, (int64_t)k, &a, &b, &wn), "llz_krylov_fetch");
            for (int pass = 0; pass < 3 && needs_refinement(b, wn); ++pass) {
              refine_done.store(false, std::memory_order_relaxed);
              refine_request.store(k, std::memory_order_release);
              while (!refine_done.load(std::memory_order_acquire)) {
                if (abort_helper.load(std::memory_order_acquire)) return;
                detail::cpu_relax();
              }
              wn = b;
              b *= refine_shrink;
            }
            if (host_step(k, a, b)) break;  // (not published as tested: the launcher must not run further ahead)
            tested.store(k, std::memory_order_release);
          }
        } catch (...) {
          helper_error = std::current_exception();
        }
        finished.store(true, std::memory_order_release);
      });
      std::exception_ptr launch_error;
      // Row-sharded runs replicate this control flow on every rank, and the kernels of an enqueued iteration wait for
      // the peers' messages of that iteration: wherever the ranks take a decision together (refinement, end of the
      // run) every rank must have enqueued the SAME iterations, whatever its threads' timing was — exactly `depth`
      // beyond the iteration the decision is about.
      const bool lockstep = ctx.nranks() > 1;
      try {
        size_t enq = 0;
        auto top_up = [&](size_t decided) {
          const size_t target = std::min(max_iteration, decided + depth);
          while (lockstep && enq < target && enq + 2 <= capacity) {
            enqueue_step();
            enqueued.store(++enq, std::memory_order_release);
          }
        };
        while (!finished.load(std::memory_order_acquire)) {
          const size_t rq = refine_request.load(std::memory_order_acquire);
          if (rq != 0) {
            top_up(rq);
            check(llz_krylov_refine(kry, (int64_t)rq, &refine_shrink), "llz_krylov_refine");
            ++stats_.refinements;
            enq = rq;  // the iterations enqueued beyond rq used the un-refined vector and were dropped
            enqueued.store(enq, std::memory_order_release);
            refine_request.store(0, std::memory_order_relaxed);
            refine_done.store(true, std::memory_order_release);
            continue;
          }
          const size_t next = tested.load(std::memory_order_acquire) + 1;  // first iteration not tested yet
          const size_t ahead = std::min(max_iteration, next + depth);
          if (enq < ahead && enq + 2 <= capacity) {
            enqueue_step();
            enqueued.store(++enq, std::memory_order_release);
          } else if (enq < next && enq < max_iteration) {
            throw basis_full(enq);  // a needed iteration does not fit
          } else {
            detail::cpu_relax();
          }
        }
        if (!helper_error) top_up(itern);
      } catch (...) {
        launch_error = std::current_exception();
        abort_helper.store(true, std::memory_order_release);
      }
      helper.join();
      if (launch_error) std::rethrow_exception(launch_error);
      if (helper_error) std::rethrow_exception(helper_error);
    }

    // :312-319 — Ritz vectors from T_m with the last coupling dropped
    const auto t1 = clock::now();
    const size_t m = alpha.size();
    std::vector<double> tri;  // tri[r*m + j]
    if (ritz_solver == 1) {
      std::vector<double> all, vecs;
      std::vector<double> b0(beta);
      if (!b0.empty()) b0.back() = 0.0;
      tridiagonal::implicit_ql(alpha.data(), b0.data(), m, all, &vecs);
      tri.assign(evs.size() * m, 0.0);
      for (size_t r = 0; r < evs.size(); ++r) {
        const size_t idx = find_maximum ? m - 1 - r : r;
        std::copy(vecs.begin() + idx * m, vecs.begin() + (idx + 1) * m, tri.begin() + r * m);
      }
    } else {
      tridiagonal::eigenvectors_for(alpha.data(), beta.data(), m, evs, tri);
    }
    std::vector<T> coeff(tri.size());
    for (size_t i = 0; i < tri.size(); ++i) coeff[i] = T((R)tri[i]);
    stats_.seconds_host += std::chrono::duration<double>(clock::now() - t1).count();

    eigvecs.clear();
    std::vector<llz_vec_t> outs;
    for (size_t r = 0; r < evs.size(); ++r) {
      eigvecs.emplace_back(ctx, n);
      outs.push_back(eigvecs.back().get());
    }
    if (!outs.empty())
      check(llz_krylov_combine(kry, (int64_t)m, (int64_t)outs.size(), coeff.data(), 1, outs.data()), "llz_krylov_combine");
    eigvalues.resize(evs.size());
    for (size_t r = 0; r < evs.size(); ++r) eigvalues[r] = (R)(evs[r] - (double)eigenvalue_offset);
    stats_.iterations += itern;
    last_alpha_ = alpha;
    last_beta_ = beta;
    return itern;
  }

  // Host-vector form of run_iteration (the reference's signature, lambda_lanczos.hpp:216-220).
  template <typename Iterable>
  size_t run_iteration(std::vector<real_t<T>>& eigvalues, std::vector<std::vector<T>>& eigvecs, size_t nroot, Iterable orthogonalizeTo) {
    std::vector<DeviceVector<T>> locked, out;
    for (auto it = orthogonalizeTo.cbegin(); it != orthogonalizeTo.cend(); ++it) {
      locked.emplace_back(mv_mul.context(), mv_mul.rows());
      locked.back().upload(it->data() + (it->size() == matrix_size ? mv_mul.row_offset() : 0));
    }
    const size_t it = run_iteration(eigvalues, out, nroot, locked);
    eigvecs.clear();
    for (auto& v : out) eigvecs.push_back(v.to_host());
    return it;
  }

  // run() with the eigenvectors left in device memory.
  void run_device(std::vector<real_t<T>>& eigenvalues, std::vector<DeviceVector<T>>& eigenvectors) {
    using clock = std::chrono::steady_clock;
    const auto t0 = clock::now();
    stats_ = RunStatistics();
    const uint64_t launches0 = mv_mul.context().launch_count();
    iter_counts_.clear();
    EigenPairManager<T, DeviceVector<T>> ep_manager(find_maximum, num_eigs);
    while (true) {  // lambda_lanczos.hpp:334-354
      std::vector<R> values_current;
      std::vector<DeviceVector<T>> vectors_current;
      const size_t nroot = std::min(num_eigs_per_iteration, matrix_size - ep_manager.size());
      const size_t count = run_iteration(values_current, vectors_current, nroot, ep_manager.getEigenvectors());
      iter_counts_.push_back(count);
      const bool nothing_added = ep_manager.insertEigenpairs(values_current, vectors_current);
      if (nothing_added) break;
      if (num_eigs == 1) break;
    }
    eigenvalues.clear();
    eigenvectors.clear();
    for (auto& p : ep_manager.getEigenpairs()) {
      eigenvalues.push_back(p.first);
      eigenvectors.push_back(p.second);
    }
    mv_mul.context().synchronize();
    start_dev_ = DeviceVector<T>();
    start_dev_src_ = nullptr;
    stats_.runs = iter_counts_.size();
    stats_.kernel_launches = mv_mul.context().launch_count() - launches0;
    stats_.seconds_total = std::chrono::duration<double>(clock::now() - t0).count();
  }

  // lambda_lanczos.hpp:330-366
  void run(std::vector<real_t<T>>& eigenvalues, std::vector<std::vector<T>>& eigenvectors) {
    std::vector<DeviceVector<T>> dev;
    run_device(eigenvalues, dev);
    eigenvectors.clear();
    eigenvectors.reserve(dev.size());
    for (auto& v : dev) eigenvectors.push_back(v.to_host());
  }

  // lambda_lanczos.hpp:376-386
  std::tuple<std::vector<real_t<T>>, std::vector<std::vector<T>>> run() {
    std::vector<real_t<T>> eigenvalues;
    std::vector<std::vector<T>> eigenvectors;
    run(eigenvalues, eigenvectors);
    return std::make_tuple(std::move(eigenvalues), std::move(eigenvectors));
  }

  // lambda_lanczos.hpp:394-407 — one eigenpair regardless of num_eigs
  void run(real_t<T>& eigenvalue, std::vector<T>& eigenvector) {
    const size_t keep = num_eigs;
    num_eigs = 1;
    std::vector<real_t<T>> eigenvalues;
    std::vector<std::vector<T>> eigenvectors;
    try {
      run(eigenvalues, eigenvectors);
    } catch (...) {
      num_eigs = keep;
      throw;
    }
    num_eigs = keep;
    eigenvalue = eigenvalues[0];
    eigenvector = std::move(eigenvectors[0]);
  }

  // lambda_lanczos.hpp:412-414
  const std::vector<size_t>& getIterationCounts() const { return iter_counts_; }

  const RunStatistics& statistics() const { return stats_; }
  // alpha/beta of the most recent Lanczos run (beta as measured, i.e. before the final coupling is dropped)
  const std::vector<double>& last_alpha() const { return last_alpha_; }
  const std::vector<double>& last_beta() const { return last_beta_; }
  // the device workspace of the most recent run (tests read Lanczos vectors through it)
  llz_krylov_t workspace() const { return work_.get(); }

 private:
  std::vector<size_t> iter_counts_;
  KrylovWorkspace work_;
  RunStatistics stats_;
  std::vector<double> last_alpha_, last_beta_;
  DeviceVector<T> start_dev_;
  const T* start_dev_src_ = nullptr;
};

}  // namespace lambda_lanczos_b200
