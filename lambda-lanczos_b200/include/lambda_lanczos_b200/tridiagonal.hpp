// tridiagonal.hpp — host-side eigen-solvers for the small real symmetric tridiagonal matrix T_k of the Lanczos
// recurrence.  They stay on the host (BASELINE.json north_star); what changes against the reference is their cost:
//
//   * the reference runs a full implicit-shift QR on T_k in EVERY iteration (lambda_lanczos.hpp:268 ->
//     lambda_lanczos_tridiagonal_impl.hpp:291-361, O(k^2) per step, O(k^3) with vectors), which is ~2.5% of its CPU
//     time but would be >10x the whole GPU iteration;
//   * here the per-iteration test only needs the `nroot` extreme Ritz values: extreme_eigenvalues() finds them by
//     Sturm-sequence bisection (O(k) per evaluation, all roots advanced together, brackets warm-started from the
//     previous iteration through Cauchy interlacing);
//   * eigenvectors of T_m for the final assembly come from a twisted factorisation (O(m) each) with inverse-iteration
//     clean-up for clustered values instead of an O(m^3) accumulation;
//   * implicit_ql() is a complete eigen-decomposition (values + vectors) used by the Exponentiator (k is a few dozen)
//     and selectable as the per-iteration Ritz solver for A/B comparisons.
//
// Independent implementation (textbook algorithms: Sturm count, implicit QL with Wilkinson shift, dqds-style twisted
// factorisation); API names follow the reference's tridiagonal:: namespace where they overlap.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <limits>
#include <numeric>
#include <vector>

namespace lambda_lanczos_b200 {
namespace tridiagonal {

// Number of eigenvalues of T (diagonal a[0..m), off-diagonal b[0..m-1)) strictly below x.
template <typename R>
inline size_t sturm_count(const R* a, const R* b, size_t m, R x, R pivmin) {
  size_t count = 0;
  R q = a[0] - x;
  if (std::abs(q) < pivmin) q = -pivmin;
  if (q < 0) ++count;
  for (size_t i = 1; i < m; ++i) {
    q = a[i] - x - b[i - 1] * b[i - 1] / q;
    if (std::abs(q) < pivmin) q = -pivmin;
    if (q < 0) ++count;
  }
  return count;
}

template <typename R>
struct Bounds {
  R lo, hi, norm, pivmin;
};

// Gerschgorin interval of T and the pivot floor used by the Sturm recurrences.
template <typename R>
inline Bounds<R> gerschgorin(const R* a, const R* b, size_t m) {
  R lo = a[0], hi = a[0], bmax = 0;
  for (size_t i = 0; i < m; ++i) {
    R r = (i > 0 ? std::abs(b[i - 1]) : R(0)) + (i + 1 < m ? std::abs(b[i]) : R(0));
    lo = std::min(lo, a[i] - r);
    hi = std::max(hi, a[i] + r);
    if (i + 1 < m) bmax = std::max(bmax, std::abs(b[i]));
  }
  const R eps = std::numeric_limits<R>::epsilon();
  const R norm = std::max(std::abs(lo), std::abs(hi));
  Bounds<R> g;
  g.norm = norm;
  g.lo = lo - 2 * eps * norm * R(m) - std::numeric_limits<R>::min();
  g.hi = hi + 2 * eps * norm * R(m) + std::numeric_limits<R>::min();
  g.pivmin = std::max(std::numeric_limits<R>::min() * std::max(R(1), bmax * bmax), std::numeric_limits<R>::min());
  return g;
}

// State carried from one Lanczos iteration to the next so that brackets can be warm-started.
template <typename R>
struct ExtremeState {
  std::vector<R> prev;   // extreme eigenvalues of T_{m-1}, in the order they were returned
  std::vector<R> delta;  // how far each of them moved in the last iteration (predicts the next move)
  size_t prev_m = 0;
  bool prev_max = false;
  void clear() {
    prev.clear();
    delta.clear();
    prev_m = 0;
  }
};

// Sturm counts at up to 16 shifts in ONE sweep over T: the recurrences of different shifts are independent, so their
// divisions pipeline (a single recurrence is bound by the divider's latency).  bb[i] = b[i]^2.
template <typename R>
inline void sturm_counts_batched(const R* a, const R* bb, size_t m, const R* x, size_t nx, R pivmin, size_t* cnt) {
  constexpr size_t kW = 16;
  R q[kW];
  for (size_t r = 0; r < nx; ++r) {
    q[r] = a[0] - x[r];
    if (std::abs(q[r]) < pivmin) q[r] = -pivmin;
    cnt[r] = q[r] < 0 ? 1 : 0;
  }
  for (size_t i = 1; i < m; ++i) {
    const R b2 = bb[i - 1];
    const R ai = a[i];
    for (size_t r = 0; r < nx; ++r) {
      R t = ai - x[r] - b2 / q[r];
      t = (std::abs(t) < pivmin) ? -pivmin : t;
      q[r] = t;
      cnt[r] += t < 0 ? 1 : 0;
    }
  }
}

// The `nroot` smallest (find_max = false, ascending) or largest (find_max = true, descending) eigenvalues of T_m —
// what lambda_lanczos.hpp:264-277 extracts from the full spectrum.  Sturm bisection with all roots advanced together;
// from the second call on the brackets are predicted from the previous iteration (Cauchy interlacing bounds the move
// of every extreme value, and its last move predicts the size of the next one), checked by one batched Sturm sweep
// and widened when the prediction fails — so a converging value costs a handful of sweeps, not ~50.
template <typename R>
inline void extreme_eigenvalues(const R* a, const R* b, size_t m, size_t nroot, bool find_max, std::vector<R>& out,
                                ExtremeState<R>* state = nullptr) {
  nroot = std::min(nroot, m);
  out.assign(nroot, R(0));
  if (m == 0 || nroot == 0) return;
  if (m == 1) {
    out[0] = a[0];
    if (state) {
      state->prev = out;
      state->delta.assign(1, std::numeric_limits<R>::infinity());
      state->prev_m = 1;
      state->prev_max = find_max;
    }
    return;
  }
  const Bounds<R> g = gerschgorin(a, b, m);
  const R eps = std::numeric_limits<R>::epsilon();
  constexpr size_t kMaxBatch = 8;
  std::vector<R> bb(m > 1 ? m - 1 : 1);
  for (size_t i = 0; i + 1 < m; ++i) bb[i] = b[i] * b[i];
  std::vector<R> lo(nroot), hi(nroot);
  std::vector<size_t> want(nroot);  // ascending index of each requested root
  for (size_t r = 0; r < nroot; ++r) {
    want[r] = find_max ? m - 1 - r : r;
    lo[r] = g.lo;
    hi[r] = g.hi;
  }
  const R slack = 4 * eps * g.norm;
  if (state && state->prev_m + 1 == m && state->prev_max == find_max && !state->prev.empty()) {
    const std::vector<R>& pv = state->prev;
    for (size_t r0 = 0; r0 < nroot; r0 += kMaxBatch) {
      const size_t nb = std::min(kMaxBatch, nroot - r0);
      // interlacing bracket [il, ih] and the predicted inner end `pred` (the side the value moves towards)
      R il[kMaxBatch], ih[kMaxBatch], pred[kMaxBatch], xs[2 * kMaxBatch];
      size_t cnt[2 * kMaxBatch];
      bool have[kMaxBatch];
      for (size_t k = 0; k < nb; ++k) {
        const size_t r = r0 + k;
        il[k] = g.lo;
        ih[k] = g.hi;
        have[k] = r < pv.size();
        if (!find_max) {  // the r-th smallest moves DOWN, and stays above the (r-1)-th smallest of T_{m-1}
          if (have[k]) ih[k] = std::min(g.hi, pv[r] + slack);
          if (r >= 1 && r - 1 < pv.size()) il[k] = std::max(g.lo, pv[r - 1] - slack);
        } else {  // mirrored
          if (have[k]) il[k] = std::max(g.lo, pv[r] - slack);
          if (r >= 1 && r - 1 < pv.size()) ih[k] = std::min(g.hi, pv[r - 1] + slack);
        }
        const R d = (have[k] && r < state->delta.size()) ? state->delta[r] : std::numeric_limits<R>::infinity();
        const R step = std::max(4 * d, 64 * slack);
        if (!find_max)
          pred[k] = (have[k] && std::isfinite(d)) ? std::max(il[k], pv[r] - step) : il[k];
        else
          pred[k] = (have[k] && std::isfinite(d)) ? std::min(ih[k], pv[r] + step) : ih[k];
        xs[2 * k] = find_max ? il[k] : pred[k];      // candidate lower end
        xs[2 * k + 1] = find_max ? pred[k] : ih[k];  // candidate upper end
      }
      sturm_counts_batched(a, bb.data(), m, xs, 2 * nb, g.pivmin, cnt);
      // ends that failed the check are replaced by the interlacing end (one more sweep), then by Gerschgorin
      R ys[2 * kMaxBatch];
      size_t cnt2[2 * kMaxBatch];
      bool retry = false;
      for (size_t k = 0; k < nb; ++k) {
        const size_t w = want[r0 + k];
        const bool lo_ok = cnt[2 * k] <= w, hi_ok = cnt[2 * k + 1] >= w + 1;
        ys[2 * k] = lo_ok ? xs[2 * k] : il[k];
        ys[2 * k + 1] = hi_ok ? xs[2 * k + 1] : ih[k];
        if (!lo_ok || !hi_ok) retry = true;
      }
      if (retry) sturm_counts_batched(a, bb.data(), m, ys, 2 * nb, g.pivmin, cnt2);
      for (size_t k = 0; k < nb; ++k) {
        const size_t w = want[r0 + k];
        const size_t cl = retry ? cnt2[2 * k] : cnt[2 * k], ch = retry ? cnt2[2 * k + 1] : cnt[2 * k + 1];
        lo[r0 + k] = (cl <= w) ? ys[2 * k] : g.lo;
        hi[r0 + k] = (ch >= w + 1) ? ys[2 * k + 1] : g.hi;
        if (!(lo[r0 + k] < hi[r0 + k])) {
          lo[r0 + k] = g.lo;
          hi[r0 + k] = g.hi;
        }
      }
    }
  }

  for (size_t r0 = 0; r0 < nroot; r0 += kMaxBatch) {
    const size_t nb = std::min(kMaxBatch, nroot - r0);
    R l[kMaxBatch], h[kMaxBatch], x[kMaxBatch];
    size_t cnt[kMaxBatch];
    bool active[kMaxBatch];
    for (size_t r = 0; r < nb; ++r) {
      l[r] = lo[r0 + r];
      h[r] = hi[r0 + r];
      active[r] = true;
    }
    for (int iter = 0; iter < 2 * std::numeric_limits<R>::digits + 16; ++iter) {
      bool any = false;
      for (size_t r = 0; r < nb; ++r) {
        const R tol = 2 * eps * std::max(std::abs(l[r]), std::abs(h[r])) + 2 * g.pivmin;
        x[r] = l[r] + (h[r] - l[r]) * R(0.5);
        active[r] = (h[r] - l[r] > tol) && x[r] > l[r] && x[r] < h[r];
        any = any || active[r];
      }
      if (!any) break;
      sturm_counts_batched(a, bb.data(), m, x, nb, g.pivmin, cnt);
      for (size_t r = 0; r < nb; ++r) {
        if (!active[r]) continue;
        if (cnt[r] >= want[r0 + r] + 1)
          h[r] = x[r];
        else
          l[r] = x[r];
      }
    }
    for (size_t r = 0; r < nb; ++r) out[r0 + r] = l[r] + (h[r] - l[r]) * R(0.5);
  }
  if (state) {
    std::vector<R> d(nroot, std::numeric_limits<R>::infinity());
    if (state->prev_m + 1 == m && state->prev_max == find_max)
      for (size_t r = 0; r < nroot && r < state->prev.size(); ++r) d[r] = std::abs(out[r] - state->prev[r]);
    state->delta = d;
    state->prev = out;
    state->prev_m = m;
    state->prev_max = find_max;
  }
}

// Complete eigen-decomposition by the implicit QL algorithm with Wilkinson shifts.  d[0..m) diagonal, e[0..m-1)
// off-diagonal.  On return `values` ascending; if `vectors` != nullptr, vectors[j*m + i] is component i of the
// eigenvector of values[j] (row j = eigenvector j, the storage convention of the reference's tridiagonal_eigenpairs,
// lambda_lanczos_tridiagonal_impl.hpp:220-222).  Returns the number of eigenvalues that hit the iteration limit.
template <typename R>
inline size_t implicit_ql(const R* d_in, const R* e_in, size_t m, std::vector<R>& values, std::vector<R>* vectors) {
  values.assign(d_in, d_in + m);
  if (m == 0) return 0;
  std::vector<R> e(m, R(0));
  for (size_t i = 0; i + 1 < m; ++i) e[i] = e_in[i];
  std::vector<R> z;  // column-major m x m: z[i + j*m] = component i of vector j (columns are rotated)
  if (vectors) {
    z.assign(m * m, R(0));
    for (size_t i = 0; i < m; ++i) z[i + i * m] = 1;
  }
  R* d = values.data();
  const R eps = std::numeric_limits<R>::epsilon();
  size_t failures = 0;
  for (size_t l = 0; l < m; ++l) {
    int iter = 0;
    for (;;) {
      size_t k = l;
      for (; k + 1 < m; ++k) {
        const R dd = std::abs(d[k]) + std::abs(d[k + 1]);
        if (std::abs(e[k]) <= eps * dd) break;
      }
      if (k == l) break;
      if (++iter > 60) {
        ++failures;
        break;
      }
      // Wilkinson shift from the leading 2x2 of the active block
      R gq = (d[l + 1] - d[l]) / (2 * e[l]);
      R r = std::hypot(gq, R(1));
      gq = d[k] - d[l] + e[l] / (gq + (gq >= 0 ? std::abs(r) : -std::abs(r)));
      R s = 1, c = 1, p = 0;
      bool underflow = false;
      size_t i = k;
      while (i-- > l) {
        R f = s * e[i];
        const R bq = c * e[i];
        r = std::hypot(f, gq);
        e[i + 1] = r;
        if (r == 0) {
          d[i + 1] -= p;
          e[k] = 0;
          underflow = true;
          break;
        }
        s = f / r;
        c = gq / r;
        gq = d[i + 1] - p;
        r = (d[i] - gq) * s + 2 * c * bq;
        p = s * r;
        d[i + 1] = gq + p;
        gq = c * r - bq;
        if (vectors) {
          R* zi = z.data() + i * m;
          R* zi1 = z.data() + (i + 1) * m;
          for (size_t t = 0; t < m; ++t) {
            f = zi1[t];
            zi1[t] = s * zi[t] + c * f;
            zi[t] = c * zi[t] - s * f;
          }
        }
      }
      if (underflow) continue;
      d[l] -= p;
      e[l] = gq;
      e[k] = 0;
    }
  }
  // ascending order
  std::vector<size_t> order(m);
  std::iota(order.begin(), order.end(), size_t(0));
  std::stable_sort(order.begin(), order.end(), [&](size_t x, size_t y) { return d[x] < d[y]; });
  std::vector<R> sorted(m);
  for (size_t j = 0; j < m; ++j) sorted[j] = d[order[j]];
  if (vectors) {
    vectors->assign(m * m, R(0));
    for (size_t j = 0; j < m; ++j) std::copy(z.begin() + order[j] * m, z.begin() + (order[j] + 1) * m, vectors->begin() + j * m);
  }
  values.swap(sorted);
  return failures;
}

// API-compatible wrappers (tridiagonal::tridiagonal_eigenvalues / tridiagonal_eigenpairs of the reference,
// lambda_lanczos_tridiagonal_impl.hpp:291-361): beta may carry a trailing unused entry, as the reference's does.
template <typename R>
inline size_t tridiagonal_eigenvalues(const std::vector<R>& alpha, const std::vector<R>& beta, std::vector<R>& eigenvalues) {
  return implicit_ql(alpha.data(), beta.data(), alpha.size(), eigenvalues, static_cast<std::vector<R>*>(nullptr));
}
template <typename R>
inline size_t tridiagonal_eigenpairs(const std::vector<R>& alpha, const std::vector<R>& beta, std::vector<R>& eigenvalues,
                                     std::vector<std::vector<R>>& eigenvectors) {
  std::vector<R> flat;
  const size_t m = alpha.size();
  size_t f = implicit_ql(alpha.data(), beta.data(), m, eigenvalues, &flat);
  eigenvectors.assign(m, std::vector<R>(m));
  for (size_t j = 0; j < m; ++j) std::copy(flat.begin() + j * m, flat.begin() + (j + 1) * m, eigenvectors[j].begin());
  return f;
}

namespace detail {

// Solve (T - lambda I) y = rhs in place by Gaussian elimination with partial pivoting (tridiagonal => at most one
// extra super-diagonal).  Near-singular pivots are replaced by `tiny`: this is inverse iteration, growth is the point.
template <typename R>
inline void shifted_solve(const R* a, const R* b, size_t m, R lambda, R tiny, std::vector<R>& y) {
  std::vector<R> dl(m, R(0)), dd(m), du(m, R(0)), du2(m, R(0));
  for (size_t i = 0; i < m; ++i) {
    dd[i] = a[i] - lambda;
    if (i + 1 < m) {
      dl[i] = b[i];
      du[i] = b[i];
    }
  }
  for (size_t i = 0; i + 1 < m; ++i) {
    if (std::abs(dd[i]) >= std::abs(dl[i])) {
      if (std::abs(dd[i]) < tiny) dd[i] = dd[i] < 0 ? -tiny : tiny;
      const R f = dl[i] / dd[i];
      dd[i + 1] -= f * du[i];
      y[i + 1] -= f * y[i];
    } else {  // swap rows i and i+1
      const R f = dd[i] / dl[i];
      dd[i] = dl[i];
      const R t = dd[i + 1];
      dd[i + 1] = du[i] - f * t;
      if (i + 2 < m) {
        du2[i] = du[i + 1];
        du[i + 1] = -f * du2[i];
      }
      du[i] = t;
      const R ty = y[i];
      y[i] = y[i + 1];
      y[i + 1] = ty - f * y[i + 1];
    }
  }
  if (std::abs(dd[m - 1]) < tiny) dd[m - 1] = dd[m - 1] < 0 ? -tiny : tiny;
  y[m - 1] /= dd[m - 1];
  if (m >= 2) y[m - 2] = (y[m - 2] - du[m - 2] * y[m - 1]) / dd[m - 2];
  for (size_t i = m - 2; i-- > 0;) y[i] = (y[i] - du[i] * y[i + 1] - du2[i] * y[i + 2]) / dd[i];
}

template <typename R>
inline R norm2(const std::vector<R>& v) {
  R scale = 0;
  for (R x : v) scale = std::max(scale, std::abs(x));
  if (scale == 0) return 0;
  R s = 0;
  for (R x : v) s += (x / scale) * (x / scale);
  return scale * std::sqrt(s);
}

}  // namespace detail

// Unit eigenvectors of T_m for the given eigenvalues (each already accurate to working precision, e.g. from
// extreme_eigenvalues), stored row-wise: vectors[r*m + i].  Isolated values: one twisted-factorisation solve, O(m).
// Values closer than 1e-3*||T|| to their predecessor are treated as a cluster and additionally cleaned by inverse
// iteration with Gram-Schmidt against the other members (cf. LAPACK xSTEIN's strategy).
template <typename R>
inline void eigenvectors_for(const R* a, const R* b, size_t m, const std::vector<R>& lambdas, std::vector<R>& vectors) {
  const size_t nv = lambdas.size();
  vectors.assign(nv * m, R(0));
  if (m == 0) return;
  if (m == 1) {
    for (size_t r = 0; r < nv; ++r) vectors[r] = 1;
    return;
  }
  const Bounds<R> g = gerschgorin(a, b, m);
  const R eps = std::numeric_limits<R>::epsilon();
  const R tiny = std::max(eps * g.norm, std::numeric_limits<R>::min() * R(1e8));
  std::vector<R> dplus(m), dminus(m), z(m);
  std::vector<size_t> cluster_start(nv, 0);
  for (size_t r = 0; r < nv; ++r) {
    const R lam = lambdas[r];
    // stationary (top-down) and progressive (bottom-up) pivots of T - lam I
    dplus[0] = a[0] - lam;
    for (size_t i = 0; i + 1 < m; ++i) {
      if (std::abs(dplus[i]) < tiny) dplus[i] = dplus[i] < 0 ? -tiny : tiny;
      dplus[i + 1] = (a[i + 1] - lam) - b[i] * b[i] / dplus[i];
    }
    dminus[m - 1] = a[m - 1] - lam;
    for (size_t i = m - 1; i-- > 0;) {
      if (std::abs(dminus[i + 1]) < tiny) dminus[i + 1] = dminus[i + 1] < 0 ? -tiny : tiny;
      dminus[i] = (a[i] - lam) - b[i] * b[i] / dminus[i + 1];
    }
    size_t twist = 0;
    R best = std::numeric_limits<R>::max();
    for (size_t i = 0; i < m; ++i) {
      const R gamma = std::abs(dplus[i] + dminus[i] - (a[i] - lam));
      if (gamma < best) {
        best = gamma;
        twist = i;
      }
    }
    z[twist] = 1;
    for (size_t i = twist; i-- > 0;) {
      R dp = dplus[i];
      if (std::abs(dp) < tiny) dp = dp < 0 ? -tiny : tiny;
      z[i] = -(b[i] / dp) * z[i + 1];
    }
    for (size_t i = twist; i + 1 < m; ++i) {
      R dm = dminus[i + 1];
      if (std::abs(dm) < tiny) dm = dm < 0 ? -tiny : tiny;
      z[i + 1] = -(b[i] / dm) * z[i];
    }
    R nz = detail::norm2(z);
    if (!(nz > 0) || !std::isfinite(nz)) {  // pathological: fall back to a unit vector and let inverse iteration work
      std::fill(z.begin(), z.end(), R(0));
      z[twist] = 1;
      nz = 1;
    }
    for (size_t i = 0; i < m; ++i) z[i] /= nz;

    // cluster detection against the previous requested value
    const bool clustered = r > 0 && std::abs(lambdas[r] - lambdas[r - 1]) < R(1e-3) * g.norm;
    cluster_start[r] = clustered ? cluster_start[r - 1] : r;
    if (clustered) {
      // nudge the shift apart inside the cluster so that the solves do not return the same direction
      const R shift = lam + R(r - cluster_start[r]) * 10 * eps * g.norm * (lambdas[r] >= lambdas[r - 1] ? 1 : -1);
      for (int sweep = 0; sweep < 4; ++sweep) {
        for (size_t p = cluster_start[r]; p < r; ++p) {
          R dot = 0;
          for (size_t i = 0; i < m; ++i) dot += vectors[p * m + i] * z[i];
          for (size_t i = 0; i < m; ++i) z[i] -= dot * vectors[p * m + i];
        }
        R nn = detail::norm2(z);
        if (!(nn > 0)) {
          for (size_t i = 0; i < m; ++i) z[i] = R(1) / R(i + 1 + r);
          nn = detail::norm2(z);
        }
        for (size_t i = 0; i < m; ++i) z[i] /= nn;
        if (sweep == 3) break;
        detail::shifted_solve(a, b, m, shift, tiny, z);
        nn = detail::norm2(z);
        for (size_t i = 0; i < m; ++i) z[i] /= nn;
      }
    }
    std::copy(z.begin(), z.end(), vectors.begin() + r * m);
  }
}

}  // namespace tridiagonal
}  // namespace lambda_lanczos_b200
