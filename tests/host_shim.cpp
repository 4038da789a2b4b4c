// tests/host_shim.cpp — exposes the header-only host solvers (lambda_lanczos_b200/tridiagonal.hpp) through a tiny C ABI
// so that the CPU test-suite can exercise them without a GPU.  Test infrastructure only.
#include <cstdint>
#include <vector>

#include "lambda_lanczos_b200/tridiagonal.hpp"

using namespace lambda_lanczos_b200::tridiagonal;

extern "C" {

void ht_extreme(const double* a, const double* b, int64_t m, int64_t nroot, int find_max, double* out) {
  std::vector<double> r;
  extreme_eigenvalues(a, b, (size_t)m, (size_t)nroot, find_max != 0, r);
  for (size_t i = 0; i < r.size(); ++i) out[i] = r[i];
}

// Replays a growing tridiagonal matrix (m = 1..mmax) with warm-started brackets, as the engine does per iteration.
void ht_extreme_sequence(const double* a, const double* b, int64_t mmax, int64_t nroot, int find_max, double* out /* mmax x nroot */) {
  ExtremeState<double> st;
  std::vector<double> r;
  for (int64_t m = 1; m <= mmax; ++m) {
    extreme_eigenvalues(a, b, (size_t)m, (size_t)nroot, find_max != 0, r, &st);
    for (size_t i = 0; i < r.size(); ++i) out[(m - 1) * nroot + i] = r[i];
  }
}

int64_t ht_ql(const double* a, const double* b, int64_t m, double* values, double* vectors) {
  std::vector<double> v, z;
  size_t f = implicit_ql(a, b, (size_t)m, v, vectors ? &z : nullptr);
  for (int64_t i = 0; i < m; ++i) values[i] = v[i];
  if (vectors)
    for (int64_t i = 0; i < m * m; ++i) vectors[i] = z[i];
  return (int64_t)f;
}

void ht_eigvecs(const double* a, const double* b, int64_t m, const double* lambdas, int64_t nv, double* vectors) {
  std::vector<double> l(lambdas, lambdas + nv), z;
  eigenvectors_for(a, b, (size_t)m, l, z);
  for (int64_t i = 0; i < nv * m; ++i) vectors[i] = z[i];
}

int64_t ht_sturm(const double* a, const double* b, int64_t m, double x) {
  return (int64_t)sturm_count(a, b, (size_t)m, x, 1e-300);
}
}
