// tests/cpp/reference_cases.cu — the reference's own test cases (test/lambda_lanczos_test.cpp, test/exponentiator_test.cpp)
// re-expressed against the C++ drop-in API of this repository: same class names, constructor arguments, public
// fields and run() overloads; the only change a user makes is that `mv_mul` is a DeviceOperator (here: COO/CSR
// built-ins and user CUDA kernels plugged in through DeviceOperator::callback — the "user __device__ functor" path).
// Built by tests/cpp/Makefile, run by tests/test_gpu_cpp_api.py on the GPU box.
#include <cmath>
#include <complex>
#include <cstdio>
#include <numeric>
#include <random>
#include <vector>

#include <cuda_runtime.h>

#include "lambda_lanczos_b200/exponentiator.hpp"
#include "lambda_lanczos_b200/functor_operator.cuh"
#include "lambda_lanczos_b200/lambda_lanczos.hpp"

namespace ll = lambda_lanczos_b200;
using ll::DeviceOperator;
using ll::Exponentiator;
using ll::LambdaLanczos;
template <typename T> using vector = std::vector<T>;
template <typename T> using complex = std::complex<T>;

static int g_failed = 0, g_checks = 0;
#define EXPECT_NEAR(expected, actual, tol)                                                                         \
  do {                                                                                                             \
    ++g_checks;                                                                                                    \
    const double e__ = (double)(expected), a__ = (double)(actual), t__ = (double)(tol);                            \
    if (!(std::abs(e__ - a__) <= t__)) {                                                                           \
      ++g_failed;                                                                                                  \
      std::printf("  FAIL %s:%d  expected %.17g got %.17g (tol %.3g)\n", __FILE__, __LINE__, e__, a__, t__);       \
    }                                                                                                              \
  } while (0)
#define EXPECT_EQ(expected, actual)                                                                  \
  do {                                                                                               \
    ++g_checks;                                                                                      \
    if (!((expected) == (actual))) {                                                                 \
      ++g_failed;                                                                                    \
      std::printf("  FAIL %s:%d  %s != %s\n", __FILE__, __LINE__, #expected, #actual);               \
    }                                                                                                \
  } while (0)
#define RUN(name)                       \
  do {                                  \
    std::printf("[ RUN ] %s\n", #name); \
    name();                             \
  } while (0)

// seeded start vectors as in the reference's tests (test/lambda_lanczos_test.cpp:25-45)
template <typename T> void vector_initializer(vector<T>& v);
template <> void vector_initializer(vector<double>& v) {
  std::mt19937 mt(1);
  std::uniform_real_distribution<double> rand(-1.0, 1.0);
  for (auto& x : v) x = rand(mt);
}
template <> void vector_initializer(vector<complex<double>>& v) {
  std::mt19937 mt(1);
  std::uniform_real_distribution<double> rand(-1.0, 1.0);
  for (auto& x : v) {
    const double re = rand(mt), im = rand(mt);
    x = complex<double>(re, im);
  }
}

// dense n x n host matrix -> COO triplets -> device operator
template <typename T, typename M> DeviceOperator<T> dense_operator(const ll::Context& ctx, const M& matrix, size_t n) {
  vector<size_t> rows, cols;
  vector<T> vals;
  for (size_t i = 0; i < n; ++i)
    for (size_t j = 0; j < n; ++j)
      if (matrix[i][j] != T()) {
        rows.push_back(i);
        cols.push_back(j);
        vals.push_back(matrix[i][j]);
      }
  return DeviceOperator<T>::coo(ctx, n, rows, cols, vals);
}

// ---- user device code: the matrix-free 1-D hopping chain of DYNAMIC_MATRIX / MULTIPLE_DEGENERATE / the exponentiator
//      tests, written the way the reference's lambda is (out += A in; out arrives zero-filled) ----
template <typename T> __global__ void chain_kernel(const T* in, T* out, size_t n, double t, bool periodic);
template <> __global__ void chain_kernel<double>(const double* in, double* out, size_t n, double t, bool periodic) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  if (i + 1 < n) s += t * in[i + 1];
  if (i > 0) s += t * in[i - 1];
  if (periodic && i == 0) s += t * in[n - 1];
  if (periodic && i == n - 1) s += t * in[0];
  out[i] += s;
}
template <> __global__ void chain_kernel<double2>(const double2* in, double2* out, size_t n, double t, bool periodic) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double sx = 0.0, sy = 0.0;
  auto add = [&](size_t j) {
    sx += t * in[j].x;
    sy += t * in[j].y;
  };
  if (i + 1 < n) add(i + 1);
  if (i > 0) add(i - 1);
  if (periodic && i == 0) add(n - 1);
  if (periodic && i == n - 1) add(0);
  out[i].x += sx;
  out[i].y += sy;
}
template <typename T, typename D> DeviceOperator<T> chain_operator(const ll::Context& ctx, size_t n, double t, bool periodic) {
  return DeviceOperator<T>::callback(
      ctx, n,
      [t, periodic](const T* x, T* y, size_t len, void* stream) {
        chain_kernel<D><<<(unsigned)((len + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const D*>(x), reinterpret_cast<D*>(y), len, t, periodic);
      },
      /*overwrites=*/false);
}

static ll::Context* g_ctx = nullptr;

void SIMPLE_MATRIX() {  // test/lambda_lanczos_test.cpp:128-161
  const size_t n = 3;
  double matrix[n][n] = {{2.0, 1.0, 1.0}, {1.0, 2.0, 1.0}, {1.0, 1.0, 2.0}};
  LambdaLanczos<double> engine(dense_operator<double>(*g_ctx, matrix, n), n, true, 1);
  engine.init_vector = vector_initializer<double>;
  engine.eigenvalue_offset = 6.0;
  double eigvalue;
  vector<double> eigvec(1);
  engine.run(eigvalue, eigvec);
  const double sign = eigvec[0] / std::abs(eigvec[0]);
  EXPECT_NEAR(4.0, eigvalue, std::abs(4.0 * engine.eps));
  for (size_t i = 0; i < n; ++i) EXPECT_NEAR(sign / std::sqrt(3.0), eigvec[i], std::abs(4.0 * engine.eps * 10));
  EXPECT_EQ(engine.getIterationCounts().size(), (size_t)1);
}

void SIMPLE_MATRIX_FLOAT() {  // :163-193 (default random start vector)
  const size_t n = 3;
  float matrix[n][n] = {{2.0f, 1.0f, 1.0f}, {1.0f, 2.0f, 1.0f}, {1.0f, 1.0f, 2.0f}};
  LambdaLanczos<float> engine(dense_operator<float>(*g_ctx, matrix, n), n, true, 1);
  float eigvalue;
  vector<float> eigvec(1);
  engine.run(eigvalue, eigvec);
  const float sign = eigvec[0] / std::abs(eigvec[0]);
  EXPECT_NEAR(4.0f, eigvalue, std::abs(4.0f * engine.eps));
  for (size_t i = 0; i < n; ++i) EXPECT_NEAR(sign / std::sqrt(3.0f), eigvec[i], std::abs(4.0f * engine.eps * 10));
}

void MULTIPLE_VALUE_RETURN_FEATURE() {  // :195-227
  const size_t n = 3;
  double matrix[n][n] = {{2.0, 1.0, 1.0}, {1.0, 2.0, 1.0}, {1.0, 1.0, 2.0}};
  LambdaLanczos<double> engine(dense_operator<double>(*g_ctx, matrix, n), n, true, 1);
  engine.init_vector = vector_initializer<double>;
  auto result = engine.run();
  const auto& eigvalues = std::get<0>(result);
  const auto& eigvecs = std::get<1>(result);
  EXPECT_EQ(eigvalues.size(), (size_t)1);
  const double sign = eigvecs[0][0] / std::abs(eigvecs[0][0]);
  EXPECT_NEAR(4.0, eigvalues[0], std::abs(4.0 * engine.eps));
  for (size_t i = 0; i < n; ++i) EXPECT_NEAR(sign / std::sqrt(3.0), eigvecs[0][i], std::abs(4.0 * engine.eps * 10));
}

void DYNAMIC_MATRIX() {  // :262-308 — matrix-free user kernel, smallest eigenvalue -2cos(pi/(n+1))
  const size_t n = 10;
  LambdaLanczos<double> engine(chain_operator<double, double>(*g_ctx, n, -1.0, false), n, false, 1);
  engine.init_vector = vector_initializer<double>;
  engine.eps = 1e-14;
  engine.eigenvalue_offset = -10.0;
  double eigvalue;
  vector<double> eigvec(n);
  engine.run(eigvalue, eigvec);
  const double correct = -2.0 * std::cos(M_PI / (n + 1));
  const double sign = eigvec[0] / std::abs(eigvec[0]);
  vector<double> v(n);
  double nrm = 0;
  for (size_t i = 0; i < n; ++i) {
    v[i] = std::sin((i + 1) * M_PI / (n + 1));
    nrm += v[i] * v[i];
  }
  EXPECT_NEAR(correct, eigvalue, std::abs(correct * engine.eps) * 4);
  for (size_t i = 0; i < n; ++i) EXPECT_NEAR(sign * v[i] / std::sqrt(nrm), eigvec[i], std::abs(correct * engine.eps * 10) * 4);
}

// The same case with the operator written as a user __device__ functor (functor_operator.cuh) — the GPU counterpart of
// the reference's matrix-free lambda (src/samples/sample3_dynamic.cpp:17-22).
struct HopFunctor {
  size_t n;
  __device__ double operator()(size_t i, const double* x) const { return -(i > 0 ? x[i - 1] : 0.0) - (i + 1 < n ? x[i + 1] : 0.0); }
};
struct HopFunctorComplex {
  size_t n;
  __device__ double2 operator()(size_t i, const double2* x) const {
    double2 s = make_double2(0.0, 0.0);
    if (i > 0) { s.x -= x[i - 1].x; s.y -= x[i - 1].y; }
    if (i + 1 < n) { s.x -= x[i + 1].x; s.y -= x[i + 1].y; }
    return s;
  }
};
void DYNAMIC_MATRIX_FUNCTOR() {
  const size_t n = 1000;
  const double correct = -2.0 * std::cos(M_PI / (n + 1));
  {
    LambdaLanczos<double> engine(ll::make_functor_operator<double>(*g_ctx, n, HopFunctor{n}), n, false, 1);
    engine.init_vector = vector_initializer<double>;
    double eigvalue;
    vector<double> eigvec(n);
    engine.run(eigvalue, eigvec);
    EXPECT_NEAR(correct, eigvalue, std::abs(correct) * 1e-10);
    const double sign = eigvec[0] / std::abs(eigvec[0]);
    double nrm = 0, err = 0;
    for (size_t i = 0; i < n; ++i) nrm += std::pow(std::sin((i + 1) * M_PI / (n + 1)), 2);
    for (size_t i = 0; i < n; ++i) err = std::max(err, std::abs(sign * std::sin((i + 1) * M_PI / (n + 1)) / std::sqrt(nrm) - eigvec[i]));
    EXPECT_NEAR(0.0, err, 1e-7);
  }
  {
    LambdaLanczos<complex<double>> engine(ll::make_functor_operator<complex<double>>(*g_ctx, n, HopFunctorComplex{n}), n, false, 1);
    engine.init_vector = vector_initializer<complex<double>>;
    double eigvalue;
    vector<complex<double>> eigvec(n);
    engine.run(eigvalue, eigvec);
    EXPECT_NEAR(correct, eigvalue, std::abs(correct) * 1e-10);
  }
}

void HERMITIAN_MATRIX() {  // :375-409
  const size_t n = 3;
  const complex<double> I_(0.0, 1.0);
  complex<double> matrix[n][n] = {{0.0, I_, 1.0}, {-I_, 0.0, I_}, {1.0, -I_, 0.0}};
  LambdaLanczos<complex<double>> engine(dense_operator<complex<double>>(*g_ctx, matrix, n), n, false, 1);
  engine.init_vector = vector_initializer<complex<double>>;
  double eigvalue;
  vector<complex<double>> eigvec(n);
  engine.run(eigvalue, eigvec);
  vector<complex<double>> correct{1.0, I_, -1.0};
  const complex<double> phase = std::exp(I_ * std::arg(eigvec[0]));
  EXPECT_NEAR(-2.0, eigvalue, std::abs(2.0 * engine.eps));
  for (size_t i = 0; i < n; ++i) {
    const complex<double> c = correct[i] / std::sqrt(3.0) * phase;
    EXPECT_NEAR(c.real(), eigvec[i].real(), std::abs(2.0 * engine.eps * 10));
    EXPECT_NEAR(c.imag(), eigvec[i].imag(), std::abs(2.0 * engine.eps * 10));
  }
}

void SINGLE_ELEMENT_MATRIX() {  // :411-440
  const size_t n = 1;
  double matrix[n][n] = {{2.0}};
  LambdaLanczos<double> engine(dense_operator<double>(*g_ctx, matrix, n), n, true, 1);
  engine.init_vector = vector_initializer<double>;
  double eigvalue;
  vector<double> eigvec(1);
  engine.run(eigvalue, eigvec);
  EXPECT_NEAR(2.0, eigvalue, std::abs(2.0 * engine.eps));
  EXPECT_NEAR(1.0, std::abs(eigvec[0]), 1e-12);
}

void MULTIPLE_EIGENPAIRS() {  // :442-488 (default random start)
  const int n = 8;
  const size_t nroot = 3;
  double matrix[n][n] = {{6, -3, -3, 0, -1, 1, -1, 1},  {-3, -4, 2, 2, -1, -5, 0, -4}, {-3, 2, 2, -3, 0, 0, -1, -1},
                         {0, 2, -3, 0, -3, 3, 2, 2},    {-1, -1, 0, -3, -2, 0, -5, -4}, {1, -5, 0, 3, 0, -4, 5, 0},
                         {-1, 0, -1, 2, -5, 5, -4, 4},  {1, -4, -1, 2, -4, 0, 4, 2}};
  LambdaLanczos<double> engine(dense_operator<double>(*g_ctx, matrix, n), n, false, 1);
  engine.num_eigs = nroot;
  engine.eps = 1e-7;
  vector<double> eigenvalues;
  vector<vector<double>> eigenvectors;
  engine.run(eigenvalues, eigenvectors);
  const double vals[3] = {-13.21508597, -8.50033154, -4.26674892};
  double vecs[3][8] = {{0.02081752, -0.49222707, 0.13202088, 0.24048092, 0.15089223, -0.60850056, 0.48079787, -0.24043829},
                       {0.16645991, 0.51818471, -0.00646562, -0.09493495, 0.60595718, 0.02042567, 0.52346924, 0.23043415},
                       {0.03381669, -0.07999997, 0.32090331, 0.61650970, 0.41812886, -0.01782613, -0.45571810, 0.35575946}};
  EXPECT_EQ(eigenvalues.size(), nroot);
  for (size_t r = 0; r < nroot && r < eigenvalues.size(); ++r) {
    EXPECT_NEAR(vals[r], eigenvalues[r], std::abs(vals[r] * engine.eps));
    const double sign = eigenvectors[r][0] / std::abs(eigenvectors[r][0]);
    for (int i = 0; i < n; ++i) EXPECT_NEAR(vecs[r][i] * sign, eigenvectors[r][i], std::abs(vals[r] * engine.eps * 10));
  }
}

void MULTIPLE_DEGENERATE_EIGENPAIRS() {  // :490-536 — 26 smallest of the periodic chain, 2-fold degeneracies
  const size_t n = 50;
  const int num_eigs = 26;
  LambdaLanczos<double> engine(chain_operator<double, double>(*g_ctx, n, -1.0, true), n, false, 1);
  engine.num_eigs = num_eigs;
  engine.eps = 1e-14;
  vector<double> eigvals;
  vector<vector<double>> eigvecs;
  engine.run(eigvals, eigvecs);
  vector<double> correct(num_eigs);
  std::iota(correct.begin(), correct.end(), -num_eigs / 2);
  for (auto& x : correct) x = -2.0 * std::cos(2.0 * M_PI * x / n);
  std::sort(correct.begin(), correct.end());
  EXPECT_EQ(correct.size(), eigvals.size());
  for (size_t i = 0; i < correct.size() && i < eigvals.size(); ++i) EXPECT_NEAR(correct[i], eigvals[i], 4 * engine.eps);
  std::printf("  Lanczos runs: %zu\n", engine.getIterationCounts().size());
}

// Diagnostic sweep of the degenerate ring: operator kind x start-vector policy (prints, never fails the suite)
void DEGENERATE_SWEEP() {
  const size_t n = 50;
  const int num_eigs = 26;
  vector<double> correct(num_eigs);
  std::iota(correct.begin(), correct.end(), -num_eigs / 2);
  for (auto& x : correct) x = -2.0 * std::cos(2.0 * M_PI * x / n);
  std::sort(correct.begin(), correct.end());
  vector<size_t> rows, cols;
  vector<double> vals;
  for (size_t i = 0; i < n; ++i) {
    rows.push_back(i); cols.push_back((i + 1) % n); vals.push_back(-1.0);
    rows.push_back((i + 1) % n); cols.push_back(i); vals.push_back(-1.0);
  }
  for (int opkind = 0; opkind < 2; ++opkind)
    for (int seeded = 0; seeded < 2; ++seeded)
      for (int depth = 0; depth < 2; ++depth) {
        auto op = opkind == 0 ? chain_operator<double, double>(*g_ctx, n, -1.0, true) : DeviceOperator<double>::coo(*g_ctx, n, rows, cols, vals);
        LambdaLanczos<double> engine(op, n, false, 1);
        engine.num_eigs = num_eigs;
        engine.eps = 1e-14;
        engine.pipeline_depth = depth;
        if (seeded) engine.init_vector = vector_initializer<double>;
        vector<double> eigvals;
        vector<vector<double>> eigvecs;
        engine.run(eigvals, eigvecs);
        double err = 0;
        for (size_t i = 0; i < correct.size() && i < eigvals.size(); ++i) err = std::max(err, std::abs(correct[i] - eigvals[i]));
        std::printf("  sweep op=%s start=%s depth=%d: max err %.3g, runs", opkind == 0 ? "callback" : "coo", seeded ? "seeded" : "random", depth, err);
        for (auto c : engine.getIterationCounts()) std::printf(" %zu", c);
        std::printf(", refinements %zu\n", engine.statistics().refinements);
      }
}

// Per-run diagnostics of the degenerate ring with a fresh random start per run
void DEGENERATE_TRACE() {
  const size_t n = 50;
  vector<size_t> rows, cols;
  vector<double> vals;
  for (size_t i = 0; i < n; ++i) {
    rows.push_back(i); cols.push_back((i + 1) % n); vals.push_back(-1.0);
    rows.push_back((i + 1) % n); cols.push_back(i); vals.push_back(-1.0);
  }
  auto op = DeviceOperator<double>::coo(*g_ctx, n, rows, cols, vals);
  LambdaLanczos<double> engine(op, n, false, 26);
  engine.eps = 1e-14;
  engine.pipeline_depth = 0;
  std::mt19937 gen(12345);
  engine.init_vector = [&gen](vector<double>& v) {
    std::uniform_real_distribution<double> d(-1, 1);
    for (auto& x : v) x = d(gen);
  };
  vector<ll::DeviceVector<double>> locked;
  vector<vector<double>> locked_host;
  for (int run = 0; run < 4; ++run) {
    vector<double> ev;
    vector<ll::DeviceVector<double>> vec;
    const size_t it = engine.run_iteration(ev, vec, 5, locked);
    std::printf("  run %d: itern %zu, last beta %.3g\n", run, it, engine.last_beta().back());
    for (size_t r = 0; r < ev.size(); ++r) {
      vector<double> x = vec[r].to_host();
      double res = 0, nrm = 0, maxov = 0;
      for (size_t i = 0; i < n; ++i) {
        const double ax = -x[(i + 1) % n] - x[(i + n - 1) % n];
        res += (ax - ev[r] * x[i]) * (ax - ev[r] * x[i]);
        nrm += x[i] * x[i];
      }
      for (auto& q : locked_host) {
        double d = 0;
        for (size_t i = 0; i < n; ++i) d += q[i] * x[i];
        maxov = std::max(maxov, std::abs(d));
      }
      std::printf("    root %zu: lambda %.16g  k=%g  residual %.3g  norm-1 %.3g  max overlap with locked %.3g\n", r, ev[r],
                  std::acos(-ev[r] / 2) * n / (2 * M_PI), std::sqrt(res), std::sqrt(nrm) - 1, maxov);
    }
    for (size_t r = 0; r < ev.size(); ++r) {
      locked.push_back(vec[r]);
      locked_host.push_back(vec[r].to_host());
    }
  }
}

template <typename T> double overlap_with(const vector<T>& a, const vector<T>& b) {
  T ip = T();
  double na = 0, nb = 0;
  for (size_t i = 0; i < a.size(); ++i) {
    ip += ll::util::typed_conj(a[i]) * b[i];
    na += std::norm(a[i]);
    nb += std::norm(b[i]);
  }
  return std::abs(ip) / std::sqrt(na) / std::sqrt(nb);
}

void EXPONENTIATE_REAL() {  // test/exponentiator_test.cpp:31-81
  const size_t n = 3;
  double matrix[n][n] = {{2.0, 1.0, 1.0}, {1.0, 2.0, 1.0}, {1.0, 1.0, 2.0}};
  const double a = 3;
  Exponentiator<double> exponentiator(dense_operator<double>(*g_ctx, matrix, n), n);
  vector<double> input = {1, 0, 0}, output(n);
  const size_t itern = exponentiator.run(a, input, output);
  // exact: eigenpairs (4; (1,1,1)/sqrt3), (1; ...) => exp(aA) e0 = e^{a}(e0 - s/3) + e^{4a} s/3 with s = (1,1,1)
  vector<double> exact(n);
  for (size_t i = 0; i < n; ++i) exact[i] = std::exp(a) * ((i == 0 ? 1.0 : 0.0) - 1.0 / 3) + std::exp(4 * a) / 3;
  EXPECT_EQ(itern, (size_t)3);
  EXPECT_NEAR(1.0, overlap_with(exact, output), exponentiator.eps);
  exponentiator.taylor_run(a, input, output);
  EXPECT_NEAR(1.0, overlap_with(exact, output), exponentiator.eps);
}

void EXPONENTIATE_LARGE_MATRIX_AND_ZERO_DELTA() {  // :106-222 — periodic hopping chain, user kernel, complex
  using cd = complex<double>;
  const size_t n = 100;
  const double t = -1.0;
  Exponentiator<cd> exponentiator(chain_operator<cd, double2>(*g_ctx, n, t, true), n);
  vector<cd> input(n);
  input[0] = cd(1, 2);
  input[n - 1] = cd(1, 2);
  input[n / 2] = cd(8, 2);
  double nrm = 0;
  for (auto& x : input) nrm += std::norm(x);
  for (auto& x : input) x /= std::sqrt(nrm);
  vector<cd> output;  // left unsized, as the reference test does
  const cd a(0.0, 3.0);
  const size_t itern = exponentiator.run(a, input, output);
  // exact through plane waves: out_i = sum_k e^{a 2t cos k} e^{ik i}/n sum_j e^{-ik j} in_j
  vector<cd> exact(n, cd(0));
  for (size_t q = 0; q < n; ++q) {
    const double k = 2 * M_PI / n * q;
    cd proj(0);
    for (size_t j = 0; j < n; ++j) proj += std::exp(cd(0, -k * (double)j)) * input[j];
    const cd w = std::exp(a * (2 * t * std::cos(k))) * proj / (double)n;
    for (size_t i = 0; i < n; ++i) exact[i] += w * std::exp(cd(0, k * (double)i));
  }
  EXPECT_EQ(itern, (size_t)19);  // the reference stops after 19 iterations
  EXPECT_NEAR(1.0, overlap_with(exact, output), exponentiator.eps * 2);
  const size_t itern_t = exponentiator.taylor_run(a, input, output);
  EXPECT_EQ(itern_t, (size_t)37);
  EXPECT_NEAR(1.0, overlap_with(exact, output), exponentiator.eps * 2);
  exponentiator.full_orthogonalize = true;  // ZERO_DELTA: a = 0 => identity after 2 iterations
  const size_t it0 = exponentiator.run(cd(0, 0), input, output);
  EXPECT_EQ(it0, (size_t)2);
  EXPECT_NEAR(1.0, overlap_with(input, output), exponentiator.eps);
  EXPECT_EQ(exponentiator.taylor_run(cd(0, 0), input, output), (size_t)1);
}

// ---- the cases below use the reference's constructor VERBATIM: mv_mul is a host lambda over std::vectors
//      (DeviceOperator<T>::host_function stages the vectors around it; the Krylov loop stays on the GPU) ----
void SIMPLE_MATRIX_USE_COMPLEX_TYPE(bool fix_seed) {  // test/lambda_lanczos_test.cpp:310-341 and :343-373
  using cd = complex<double>;
  const size_t n = 3;
  cd matrix[n][n] = {{2.0, 1.0, 1.0}, {1.0, 2.0, 1.0}, {1.0, 1.0, 2.0}};
  auto matmul = [&](const vector<cd>& in, vector<cd>& out) {
    for (size_t i = 0; i < n; ++i)
      for (size_t j = 0; j < n; ++j) out[i] += matrix[i][j] * in[j];
  };
  LambdaLanczos<cd> engine(matmul, n, true, 1);
  if (fix_seed) engine.init_vector = vector_initializer<cd>;
  double eigvalue;
  vector<cd> eigvec(n);
  engine.run(eigvalue, eigvec);
  const cd phase = std::exp(cd(0.0, 1.0) * std::arg(eigvec[0]));
  EXPECT_NEAR(4.0, eigvalue, std::abs(4.0 * engine.eps));
  for (size_t i = 0; i < n; ++i) {
    EXPECT_NEAR((phase / std::sqrt((double)n)).real(), eigvec[i].real(), std::abs(4.0 * engine.eps * 10));
    EXPECT_NEAR((phase / std::sqrt((double)n)).imag(), eigvec[i].imag(), std::abs(4.0 * engine.eps * 10));
  }
}
void SIMPLE_MATRIX_USE_COMPLEX_TYPE_FIXED() { SIMPLE_MATRIX_USE_COMPLEX_TYPE(true); }
void SIMPLE_MATRIX_USE_COMPLEX_TYPE_NOT_FIX_RANDOM_SEED() { SIMPLE_MATRIX_USE_COMPLEX_TYPE(false); }

// A matrix with an exactly known largest eigenpair: a random diagonal rotated by `rounds` random plane rotations
// (Jacobi rotations for the symmetric case, 2x2 unitaries for the Hermitian one), the eigenvector rotated along.  The
// random draws are made in the order of the reference's generators (:538-594, :639-712) so that std::mt19937(1) yields
// the reference's very matrices.
template <typename T, typename U2>
void rotate_planes(vector<vector<T>>& a, vector<T>& v, size_t k, size_t l, const U2& u) {  // a <- U a U^H, v <- U v
  const size_t n = a.size();
  const T akk = a[k][k], akl = a[k][l], alk = a[l][k], all = a[l][l];
  for (size_t i = 0; i < n; ++i) {
    const T rk = u.kk * a[k][i] + u.kl * a[l][i], rl = u.lk * a[k][i] + u.ll * a[l][i];
    a[k][i] = rk;
    a[l][i] = rl;
  }
  for (size_t i = 0; i < n; ++i) {
    a[i][k] = ll::util::typed_conj(a[k][i]);
    a[i][l] = ll::util::typed_conj(a[l][i]);
  }
  const T ckk = ll::util::typed_conj(u.kk), ckl = ll::util::typed_conj(u.kl), clk = ll::util::typed_conj(u.lk), cll = ll::util::typed_conj(u.ll);
  a[k][k] = u.kk * (akk * ckk + akl * ckl) + u.kl * (alk * ckk + all * ckl);
  a[k][l] = u.kk * (akk * clk + akl * cll) + u.kl * (alk * clk + all * cll);
  a[l][k] = ll::util::typed_conj(a[k][l]);
  a[l][l] = u.lk * (akk * clk + akl * cll) + u.ll * (alk * clk + all * cll);
  const T vk = u.kk * v[k] + u.kl * v[l];
  v[l] = u.lk * v[k] + u.ll * v[l];
  v[k] = vk;
}
template <typename T> struct Plane { T kk, kl, lk, ll; };

template <typename T, bool HERMITIAN>
void known_extreme_matrix(vector<vector<T>>& a, vector<T>& eigvec, double& eigvalue, size_t n, size_t rounds) {
  std::mt19937 eng(1);
  std::uniform_int_distribution<size_t> dist_index(0, n - 1);
  std::uniform_real_distribution<double> dist_angle(0.0, 2 * M_PI);
  std::uniform_real_distribution<double> dist_element(1.0, n * 10);
  a.assign(n, vector<T>(n, T()));
  eigvec.assign(n, T());
  eigvalue = 1.0;
  size_t at = 0;
  for (size_t i = 0; i < n; ++i) {
    const double d = dist_element(eng);
    a[i][i] = d;
    if (d > eigvalue) {
      eigvalue = d;
      at = i;
    }
  }
  eigvec[at] = 1.0;
  for (size_t r = 0; r < rounds; ++r) {
    const size_t k = dist_index(eng);
    size_t l = dist_index(eng);
    while (k == l) l = dist_index(eng);
    const double theta = dist_angle(eng);
    Plane<T> u;
    if constexpr (HERMITIAN) {
      const double phi1 = dist_angle(eng), phi2 = dist_angle(eng);
      const T I_(0, 1);
      u.kk = std::exp(I_ * phi1) * std::cos(theta);
      u.kl = -std::exp(I_ * phi2) * std::sin(theta);
      u.lk = std::exp(-I_ * phi2) * std::sin(theta);
      u.ll = std::exp(-I_ * phi1) * std::cos(theta);
    } else {
      u.kk = std::cos(theta);
      u.kl = -std::sin(theta);
      u.lk = std::sin(theta);
      u.ll = std::cos(theta);
    }
    rotate_planes(a, eigvec, k, l, u);
  }
}

void RANDOM_SYMMETRIC_MATRIX() {  // test/lambda_lanczos_test.cpp:596-637
  const size_t n = 50;
  vector<vector<double>> matrix;
  vector<double> correct_eigvec;
  double correct_eigvalue = 0.0;
  known_extreme_matrix<double, false>(matrix, correct_eigvec, correct_eigvalue, n, n * 10);
  auto matmul = [&](const vector<double>& in, vector<double>& out) {
    for (size_t i = 0; i < n; ++i)
      for (size_t j = 0; j < n; ++j) out[i] += matrix[i][j] * in[j];
  };
  LambdaLanczos<double> engine(matmul, n, true, 1);
  engine.init_vector = vector_initializer<double>;
  double eigvalue;
  vector<double> eigvec(n);
  engine.run(eigvalue, eigvec);
  EXPECT_NEAR(correct_eigvalue, eigvalue, std::abs(correct_eigvalue * engine.eps));
  const int sign = (eigvec[0] * correct_eigvec[0] > 0) ? 1 : -1;
  for (size_t i = 0; i < n; ++i) EXPECT_NEAR(correct_eigvec[i] * sign, eigvec[i], std::abs(correct_eigvalue * engine.eps * n * n));
}

void RANDOM_HERMITIAN_MATRIX() {  // :714-755
  using cd = complex<double>;
  const size_t n = 10;
  vector<vector<cd>> matrix;
  vector<cd> correct_eigvec;
  double correct_eigvalue = 0.0;
  known_extreme_matrix<cd, true>(matrix, correct_eigvec, correct_eigvalue, n, n * 10);
  auto matmul = [&](const vector<cd>& in, vector<cd>& out) {
    for (size_t i = 0; i < n; ++i)
      for (size_t j = 0; j < n; ++j) out[i] += matrix[i][j] * in[j];
  };
  LambdaLanczos<cd> engine(matmul, n, true, 1);
  engine.init_vector = vector_initializer<cd>;
  engine.eps = 1e-14;
  double eigvalue;
  vector<cd> eigvec(n);
  engine.run(eigvalue, eigvec);
  EXPECT_NEAR(correct_eigvalue, eigvalue, std::abs(correct_eigvalue * engine.eps));
  const cd phase = std::exp(cd(0, 1) * (std::arg(eigvec[0]) - std::arg(correct_eigvec[0])));
  for (size_t i = 0; i < n; ++i) {
    EXPECT_NEAR((correct_eigvec[i] * phase).real(), eigvec[i].real(), std::abs(correct_eigvalue * engine.eps * 10));
    EXPECT_NEAR((correct_eigvec[i] * phase).imag(), eigvec[i].imag(), std::abs(correct_eigvalue * engine.eps * 10));
  }
}

void MANHATTAN_NORM() {  // :93-100, on a device vector
  using cd = complex<double>;
  ll::DeviceVector<cd> v(*g_ctx, 2);
  v.upload(vector<cd>{cd(1.0, 3.0), cd(-1.0, -1.0)});
  EXPECT_NEAR(1.0 + 3.0 + 1.0 + 1.0, ll::util::m_norm(v), 0.0);
  ll::DeviceVector<double> w(*g_ctx, 1001);
  vector<double> h(1001);
  double expect = 0;
  for (size_t i = 0; i < h.size(); ++i) {
    h[i] = (i % 2 ? -1.0 : 1.0) * (double)i * 0.25;
    expect += std::abs(h[i]);
  }
  w.upload(h);
  EXPECT_NEAR(expect, ll::util::m_norm(w), 1e-9);
}

// The reference's sample1_simple.cpp as it stands (host lambda, default random start vector, eigenvalue_offset absent).
void SAMPLE1_UNCHANGED() {
  const int n = 3;
  double matrix[n][n] = {{2.0, 1.0, 1.0}, {1.0, 2.0, 1.0}, {1.0, 1.0, 2.0}};
  auto mv_mul = [&](const vector<double>& in, vector<double>& out) {
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) out[i] += matrix[i][j] * in[j];
  };
  LambdaLanczos<double> engine(mv_mul, n, true, 1);
  vector<double> eigenvalues;
  vector<vector<double>> eigenvectors;
  engine.run(eigenvalues, eigenvectors);
  EXPECT_NEAR(4.0, eigenvalues[0], 4.0 * engine.eps);
  Exponentiator<double> expo(mv_mul, n);
  vector<double> input = {1, 0, 0}, output(n);
  EXPECT_EQ(expo.run(3.0, input, output), (size_t)3);
}

int main() {
  try {
    ll::Context ctx(0);
    g_ctx = &ctx;
    RUN(SIMPLE_MATRIX);
    RUN(SIMPLE_MATRIX_FLOAT);
    RUN(MULTIPLE_VALUE_RETURN_FEATURE);
    RUN(DYNAMIC_MATRIX);
    RUN(DYNAMIC_MATRIX_FUNCTOR);
    RUN(HERMITIAN_MATRIX);
    RUN(SINGLE_ELEMENT_MATRIX);
    RUN(MULTIPLE_EIGENPAIRS);
    RUN(MULTIPLE_DEGENERATE_EIGENPAIRS);
    RUN(DEGENERATE_SWEEP);
    RUN(DEGENERATE_TRACE);
    RUN(EXPONENTIATE_REAL);
    RUN(EXPONENTIATE_LARGE_MATRIX_AND_ZERO_DELTA);
    RUN(SIMPLE_MATRIX_USE_COMPLEX_TYPE_FIXED);
    RUN(SIMPLE_MATRIX_USE_COMPLEX_TYPE_NOT_FIX_RANDOM_SEED);
    RUN(RANDOM_SYMMETRIC_MATRIX);
    RUN(RANDOM_HERMITIAN_MATRIX);
    RUN(MANHATTAN_NORM);
    RUN(SAMPLE1_UNCHANGED);
  } catch (const std::exception& e) {
    std::printf("EXCEPTION: %s\n", e.what());
    return 2;
  }
  std::printf("%d checks, %d failed\n", g_checks, g_failed);
  return g_failed == 0 ? 0 : 1;
}
