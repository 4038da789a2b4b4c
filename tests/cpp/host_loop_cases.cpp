// tests/cpp/host_loop_cases.cpp — CPU tests of the header-only host engine against the TEST DOUBLE of the C ABI
// (mock_llz.cpp: one worker thread plays the CUDA stream).  What is under test is host code only: the pipelined /
// two-thread iteration loop of LambdaLanczos<T>::run_iteration, the hand-over of the DGKS refinement, the lock-step
// rule of row-sharded runs, error propagation out of the helper thread, the reference-verbatim constructors, and the
// coefficient scaling of the lazily normalised Exponentiator.  Run by tests/test_host_loop.py (no GPU).
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "lambda_lanczos_b200/exponentiator.hpp"
#include "lambda_lanczos_b200/lambda_lanczos.hpp"

namespace ll = lambda_lanczos_b200;
using ll::Exponentiator;
using ll::LambdaLanczos;
template <typename T> using vector = std::vector<T>;
extern "C" int64_t mock_runs(void);
extern "C" int64_t mock_steps_of_run(int64_t);
extern "C" void mock_reset_stats(void);

static int g_failed = 0, g_checks = 0;
#define CHECK(cond)                                                     \
  do {                                                                  \
    ++g_checks;                                                         \
    if (!(cond)) {                                                      \
      ++g_failed;                                                       \
      std::printf("  FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond);     \
    }                                                                   \
  } while (0)

static vector<vector<double>> random_symmetric(size_t n, unsigned seed) {
  std::mt19937 mt(seed);
  std::uniform_real_distribution<double> d(-1.0, 1.0);
  vector<vector<double>> a(n, vector<double>(n));
  for (size_t i = 0; i < n; ++i)
    for (size_t j = i; j < n; ++j) a[i][j] = a[j][i] = d(mt);
  return a;
}
static void seeded(vector<double>& v) {
  std::mt19937 mt(1);
  std::uniform_real_distribution<double> d(-1.0, 1.0);
  for (auto& x : v) x = d(mt);
}

struct Outcome {
  vector<double> values;
  vector<vector<double>> vectors;
  vector<size_t> counts;
  size_t refinements;
};

template <class F> Outcome solve(F mv_mul, size_t n, bool find_max, size_t num_eigs, int depth, int host_threads, size_t max_iteration = 0) {
  LambdaLanczos<double> engine(mv_mul, n, find_max, num_eigs);
  engine.init_vector = seeded;
  engine.pipeline_depth = depth;
  engine.host_threads = host_threads;
  if (max_iteration) engine.max_iteration = max_iteration;
  Outcome o;
  engine.run(o.values, o.vectors);
  o.counts = engine.getIterationCounts();
  o.refinements = engine.statistics().refinements;
  return o;
}

static bool same(const Outcome& a, const Outcome& b) { return a.values == b.values && a.vectors == b.vectors && a.counts == b.counts; }

// every depth, one or two host threads: the very same bits (the GPU's work does not depend on who waits for it)
void DEPTH_AND_THREADS_DO_NOT_CHANGE_RESULTS() {
  const size_t n = 120;
  const auto a = random_symmetric(n, 7);
  auto mv = [&](const vector<double>& in, vector<double>& out) {
    for (size_t i = 0; i < n; ++i)
      for (size_t j = 0; j < n; ++j) out[i] += a[i][j] * in[j];
  };
  const Outcome ref = solve(mv, n, true, 3, 0, 1);
  CHECK(ref.values.size() == 3 && ref.values[0] > ref.values[1] && ref.values[1] > ref.values[2]);
  for (int depth : {1, 2, 4, 9, -1})
    for (int threads : {0, 1}) CHECK(same(ref, solve(mv, n, true, 3, depth, threads)));
  // residual of the largest eigenpair
  double res = 0;
  for (size_t i = 0; i < n; ++i) {
    double s = 0;
    for (size_t j = 0; j < n; ++j) s += a[i][j] * ref.vectors[0][j];
    res += (s - ref.values[0] * ref.vectors[0][i]) * (s - ref.values[0] * ref.vectors[0][i]);
  }
  CHECK(std::sqrt(res) < 1e-9);
}

// A diagonal matrix with seven distinct levels, twelve eigenpairs wanted: every Lanczos run exhausts its (at most
// seven-dimensional) Krylov space, the Gram-Schmidt pass then cancels most of the vector and the DGKS test asks for a
// second pass — which the helper thread must hand to the launch thread.  Eight-fold degenerate levels come out through
// the deflation of the outer loop (lambda_lanczos.hpp:334-354).
void REFINEMENT_IS_HANDED_TO_THE_LAUNCH_THREAD() {
  const size_t n = 60;
  auto mv = [&](const vector<double>& in, vector<double>& out) {
    for (size_t i = 0; i < n; ++i) out[i] += (double)(i % 7) * in[i];
  };
  const Outcome one = solve(mv, n, true, 12, 1, 1);
  const Outcome two = solve(mv, n, true, 12, 4, 0);
  CHECK(one.values.size() == 12 && two.values.size() == 12);
  CHECK(one.refinements > 0 && two.refinements > 0);
  CHECK(one.counts == two.counts);
  for (size_t i = 0; i < one.values.size() && i < two.values.size(); ++i) {
    CHECK(std::abs(one.values[i] - (i < 8 ? 6.0 : 5.0)) < 1e-10);  // i % 7 == 6 for 8 rows, == 5 for 9
    CHECK(std::abs(two.values[i] - one.values[i]) < 1e-12);
  }
  // the eigenvectors of a level span the rows of that level
  for (size_t r = 0; r < two.vectors.size(); ++r) {
    double inside = 0;
    for (size_t i = 0; i < n; ++i)
      if ((double)(i % 7) == (r < 8 ? 6.0 : 5.0)) inside += two.vectors[r][i] * two.vectors[r][i];
    CHECK(std::abs(inside - 1.0) < 1e-10);
  }
}

// row-sharded runs (the mock reports two ranks): whatever the threads' timing, a run enqueues exactly
// min(max_iteration, itern + depth) iterations — what every other rank enqueues — and does so on every repetition
void LOCKSTEP_ENQUEUES_THE_SAME_ITERATIONS_EVERY_TIME() {
  setenv("MOCK_NRANKS", "2", 1);
  setenv("MOCK_DELAY_US", "60", 1);
  const size_t n = 90;
  const auto a = random_symmetric(n, 11);
  auto mv = [&](const vector<double>& in, vector<double>& out) {
    for (size_t i = 0; i < n; ++i)
      for (size_t j = 0; j < n; ++j) out[i] += a[i][j] * in[j];
  };
  for (int depth : {2, 5}) {
    vector<vector<int64_t>> seen;
    for (int rep = 0; rep < 4; ++rep) {
      mock_reset_stats();
      ll::Context ctx(0);
      LambdaLanczos<double> engine(ll::DeviceOperator<double>::host_function(ctx, n, mv), n, true, 2);
      engine.init_vector = seeded;
      engine.pipeline_depth = depth;
      engine.max_iteration = 60;
      vector<double> values;
      vector<vector<double>> vectors;
      engine.run(values, vectors);
      const auto counts = engine.getIterationCounts();
      vector<int64_t> steps;
      CHECK((size_t)mock_runs() == counts.size());
      for (size_t r = 0; r < counts.size() && r < (size_t)mock_runs(); ++r) {
        steps.push_back(mock_steps_of_run((int64_t)r));
        CHECK(steps.back() == (int64_t)std::min<size_t>(60, counts[r] + (size_t)depth));
      }
      seen.push_back(steps);
    }
    for (size_t r = 1; r < seen.size(); ++r) CHECK(seen[r] == seen[0]);
  }
  unsetenv("MOCK_NRANKS");
  unsetenv("MOCK_DELAY_US");
}

// a basis that cannot hold the run: LLZ_ERR_OOM reaches the caller from either thread, nothing hangs
void A_FULL_BASIS_IS_AN_ERROR_NOT_A_HANG() {
  setenv("MOCK_BASIS_CAPACITY", "12", 1);
  const size_t n = 80;
  const auto a = random_symmetric(n, 3);
  auto mv = [&](const vector<double>& in, vector<double>& out) {
    for (size_t i = 0; i < n; ++i)
      for (size_t j = 0; j < n; ++j) out[i] += a[i][j] * in[j];
  };
  for (int depth : {1, 4}) {
    bool thrown = false;
    try {
      solve(mv, n, true, 1, depth, 0);
    } catch (const ll::Error& e) {
      thrown = e.status() == LLZ_ERR_OOM;
    }
    CHECK(thrown);
  }
  // ... while a run that stops inside the capacity is fine, speculation included
  CHECK(solve(mv, n, true, 1, 4, 0, 10).counts == vector<size_t>{10});
  unsetenv("MOCK_BASIS_CAPACITY");
}

// a user operator that throws: the exception type is lost in the C callback (status LLZ_ERR_USER), not the error
void A_THROWING_OPERATOR_SURFACES_AS_AN_ERROR() {
  const size_t n = 30;
  int calls = 0;
  auto mv = [&](const vector<double>& in, vector<double>& out) {
    if (++calls > 5) throw std::runtime_error("user operator failed");
    for (size_t i = 0; i < n; ++i) out[i] += (double)(i + 1) * in[i];
  };
  // (the mock's worker thread calls the operator; a real context calls it on the launching thread — either way the run
  //  must end; here the trampoline turns the exception into a non-zero return that the mock ignores, so the run simply
  //  proceeds on a zero vector and stops at the breakdown test)
  bool ended = false;
  try {
    solve(mv, n, true, 1, 2, 0);
    ended = true;
  } catch (const ll::Error&) {
    ended = true;
  }
  CHECK(ended);
}

// Exponentiator: exact exp(aA)v for a small Hermitian matrix, complex, with the lazily normalised recurrence the engine
// selects on one rank (un-normalised Lanczos vectors, coefficients divided by beta in the final sum) and with the
// normalised one it selects when row-sharded — both must give the exact vector, with the same iteration count
void LAZY_EXPONENTIATOR_MATCHES_THE_EXACT_EXPONENTIAL() {
  using cd = std::complex<double>;
  const size_t n = 40;
  const double t = -1.0;
  auto mv = [&](const vector<cd>& in, vector<cd>& out) {  // periodic hopping chain (test/exponentiator_test.cpp:106-222)
    for (size_t i = 0; i < n; ++i) out[i] += t * (in[(i + 1) % n] + in[(i + n - 1) % n]);
  };
  vector<cd> input(n);
  input[0] = cd(1, 2);
  input[n - 1] = cd(1, 2);
  input[n / 2] = cd(8, 2);
  const cd a(0.0, 1.5);
  vector<cd> exact(n, cd(0));
  for (size_t q = 0; q < n; ++q) {
    const double k = 2 * M_PI / n * q;
    cd proj(0);
    for (size_t j = 0; j < n; ++j) proj += std::exp(cd(0, -k * (double)j)) * input[j];
    const cd w = std::exp(a * (2 * t * std::cos(k))) * proj / (double)n;
    for (size_t i = 0; i < n; ++i) exact[i] += w * std::exp(cd(0, k * (double)i));
  }
  size_t its[2] = {0, 0};
  vector<cd> outs[2];
  for (int sharded = 0; sharded < 2; ++sharded) {
    if (sharded) setenv("MOCK_NRANKS", "2", 1);
    ll::Context ctx(0);
    Exponentiator<cd> ex(ll::DeviceOperator<cd>::host_function(ctx, n, mv), n);
    its[sharded] = ex.run(a, input, outs[sharded]);
    cd ip(0);
    double na = 0, nb = 0, err = 0;
    for (size_t i = 0; i < n; ++i) {
      ip += std::conj(exact[i]) * outs[sharded][i];
      na += std::norm(exact[i]);
      nb += std::norm(outs[sharded][i]);
      err += std::norm(outs[sharded][i] - exact[i]);
    }
    CHECK(std::abs(1.0 - std::abs(ip) / std::sqrt(na * nb)) < ex.eps * 2);  // the reference's own criterion (:218-221)
    CHECK(std::sqrt(err / na) < 1e-7);
    CHECK(std::abs(std::sqrt(nb / na) - 1.0) < 1e-12);  // unitary evolution: the norm is kept
    if (sharded) unsetenv("MOCK_NRANKS");
  }
  CHECK(its[0] == its[1] && its[0] > 5);
  double diff = 0, nrm = 0;  // lazily normalised vs normalised Lanczos vectors: the same Krylov process
  for (size_t i = 0; i < n; ++i) {
    diff += std::norm(outs[0][i] - outs[1][i]);
    nrm += std::norm(outs[1][i]);
  }
  CHECK(std::sqrt(diff / nrm) < 1e-13);
}

// Exponentiator with full reorthogonalisation: same iteration count and output for every depth (the speculative
// iterations past the stopping test are never read)
void EXPONENTIATOR_DEPTH_DOES_NOT_CHANGE_RESULTS() {
  const size_t n = 64;
  const auto a = random_symmetric(n, 5);
  auto mv = [&](const vector<double>& in, vector<double>& out) {
    for (size_t i = 0; i < n; ++i)
      for (size_t j = 0; j < n; ++j) out[i] += 0.1 * a[i][j] * in[j];
  };
  vector<double> input(n);
  seeded(input);
  vector<double> first;
  size_t it_first = 0;
  for (int full = 0; full < 2; ++full)
    for (int depth : {0, 1, 3, -1}) {
      Exponentiator<double> ex(mv, n);
      ex.full_orthogonalize = full != 0;
      ex.pipeline_depth = depth;
      vector<double> output;
      const size_t it = ex.run(-0.7, input, output);
      if (first.empty()) {
        first = output;
        it_first = it;
      }
      double diff = 0, nrm = 0;
      for (size_t i = 0; i < n; ++i) {
        diff += (output[i] - first[i]) * (output[i] - first[i]);
        nrm += first[i] * first[i];
      }
      CHECK(std::sqrt(diff / nrm) < (full ? 1e-10 : 1e-13));  // (with / without reorthogonalisation: rounding apart)
      CHECK(it == it_first || full);
    }
}

// The reference's own Exponentiator cases (test/exponentiator_test.cpp:106-222, periodic hopping chain n = 100, a = 3i):
// the iteration counts are decided by host logic alone — 19 Krylov iterations, 37 Taylor terms, 2 / 1 for a = 0.
void REFERENCE_EXPONENTIATOR_ITERATION_COUNTS() {
  using cd = std::complex<double>;
  const size_t n = 100;
  const double t = -1.0;
  auto mv = [&](const vector<cd>& in, vector<cd>& out) {
    for (size_t i = 0; i < n; ++i) out[i] += t * (in[(i + 1) % n] + in[(i + n - 1) % n]);
  };
  vector<cd> input(n);
  input[0] = cd(1, 2);
  input[n - 1] = cd(1, 2);
  input[n / 2] = cd(8, 2);
  double nrm = 0;
  for (auto& x : input) nrm += std::norm(x);
  for (auto& x : input) x /= std::sqrt(nrm);
  Exponentiator<cd> ex(mv, n);
  vector<cd> krylov, taylor;
  CHECK(ex.run(cd(0.0, 3.0), input, krylov) == 19);
  CHECK(ex.taylor_run(cd(0.0, 3.0), input, taylor) == 37);
  double diff = 0;
  for (size_t i = 0; i < n; ++i) diff += std::norm(krylov[i] - taylor[i]);
  CHECK(std::sqrt(diff) < 1e-7);
  ex.full_orthogonalize = true;
  vector<cd> same;
  CHECK(ex.run(cd(0, 0), input, same) == 2);
  diff = 0;
  for (size_t i = 0; i < n; ++i) diff += std::norm(same[i] - input[i]);
  CHECK(std::sqrt(diff) < 1e-14);
  CHECK(ex.taylor_run(cd(0, 0), input, same) == 1);
}

void AUTO_DEPTH() {
  CHECK(ll::auto_pipeline_depth(800000, true) == 4);                    // config 1: 100 k doubles
  CHECK(ll::auto_pipeline_depth((size_t)16 << 20, true) == 2);          // config 2 on 8 GPUs
  CHECK(ll::auto_pipeline_depth((size_t)134217728, true) == 1);         // config 2 on 1 GPU
  CHECK(ll::auto_pipeline_depth((size_t)641865600, false) == 0);        // config 5: lock-step
  CHECK(ll::auto_pipeline_depth(1600, false) == 4);
}

#define RUN(name)                       \
  do {                                  \
    std::printf("[ RUN ] %s\n", #name); \
    name();                             \
  } while (0)

int main() {
  try {
    RUN(DEPTH_AND_THREADS_DO_NOT_CHANGE_RESULTS);
    RUN(REFINEMENT_IS_HANDED_TO_THE_LAUNCH_THREAD);
    RUN(LOCKSTEP_ENQUEUES_THE_SAME_ITERATIONS_EVERY_TIME);
    RUN(A_FULL_BASIS_IS_AN_ERROR_NOT_A_HANG);
    RUN(A_THROWING_OPERATOR_SURFACES_AS_AN_ERROR);
    RUN(LAZY_EXPONENTIATOR_MATCHES_THE_EXACT_EXPONENTIAL);
    RUN(EXPONENTIATOR_DEPTH_DOES_NOT_CHANGE_RESULTS);
    RUN(REFERENCE_EXPONENTIATOR_ITERATION_COUNTS);
    RUN(AUTO_DEPTH);
  } catch (const std::exception& e) {
    std::printf("EXCEPTION: %s\n", e.what());
    return 2;
  }
  std::printf("%d checks, %d failed\n", g_checks, g_failed);
  return g_failed ? 1 : 0;
}
