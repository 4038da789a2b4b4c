// tests/cpp/mock_llz.cpp — a TEST DOUBLE of the C ABI (include/llz.h) for CPU tests of the header-only host engine
// (lambda_lanczos_b200/lambda_lanczos.hpp, exponentiator.hpp): the control flow around the asynchronous
// llz_krylov_step / llz_krylov_fetch / llz_krylov_refine calls — pipelining depth, the helper thread that runs the Ritz
// solves, the hand-over of the DGKS refinement, the lock-step rule of row-sharded runs, error propagation — and the
// coefficient scaling of the lazily normalised Exponentiator.
//
// NOT a backend and never linked into the product: "device" memory is host memory, one worker thread plays the CUDA
// stream (steps are queued and executed asynchronously, with a seeded random delay so that the host threads see
// different timings on every run), double and complex<double> only, dense loops.  Only what the headers call is there.
#include <atomic>
#include <chrono>
#include <cmath>
#include <complex>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <random>
#include <thread>
#include <vector>

#include "llz.h"

namespace {
thread_local char g_error[256] = "";
int fail(int status, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
  return status;
}
size_t esize(int dtype) { return dtype == LLZ_F64 ? 8 : dtype == LLZ_C128 ? 16 : 0; }
using cd = std::complex<double>;
}  // namespace

struct llz_ctx_s {
  int rank = 0, nranks = 1;
  std::atomic<uint64_t> launches{0};
};
struct llz_vec_s {
  llz_ctx_t ctx;
  int dtype;
  int64_t n;
  std::vector<char> buf;
};
struct llz_op_s {
  llz_ctx_t ctx;
  int dtype;
  int64_t n;
  llz_apply_fn fn;
  void* user;
  int overwrites;
};
struct Job {
  int64_t k;
  llz_op_t op;
  double sigma;
  int orth;
};
struct llz_krylov_s {
  llz_ctx_t ctx;
  int dtype;
  int64_t n, cap;
  std::vector<std::vector<char>> cols;
  std::vector<double> alpha, beta, wnorm;
  std::vector<llz_vec_t> locked;
  std::thread worker;
  std::mutex mu;
  std::condition_variable cv;
  std::deque<Job> q;
  std::atomic<int64_t> k_enq{0};
  std::atomic<int64_t> k_done{0};
  bool quit = false, busy = false;
  std::mt19937 rng{12345};
  int delay_us = 0;
};

// what the tests read back: steps enqueued in every run since the last reset (one entry per llz_krylov_begin)
static std::mutex g_stat_mu;
static std::vector<int64_t> g_steps_per_run;

namespace {

template <class T> double re_dot(const T* a, const T* b, int64_t n);
template <> double re_dot<double>(const double* a, const double* b, int64_t n) {
  double s = 0;
  for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}
template <> double re_dot<cd>(const cd* a, const cd* b, int64_t n) {
  double s = 0;
  for (int64_t i = 0; i < n; ++i) s += (std::conj(a[i]) * b[i]).real();
  return s;
}
template <class T> T dot(const T* a, const T* b, int64_t n);
template <> double dot<double>(const double* a, const double* b, int64_t n) { return re_dot(a, b, n); }
template <> cd dot<cd>(const cd* a, const cd* b, int64_t n) {
  cd s = 0;
  for (int64_t i = 0; i < n; ++i) s += std::conj(a[i]) * b[i];
  return s;
}

template <class T> void gram_schmidt(llz_krylov_t kry, T* w, int64_t ncols) {  // classical: all coefficients first
  const int64_t n = kry->n;
  std::vector<const T*> basis;
  for (llz_vec_t q : kry->locked) basis.push_back(reinterpret_cast<const T*>(q->buf.data()));
  for (int64_t j = 0; j < ncols; ++j) basis.push_back(reinterpret_cast<const T*>(kry->cols[(size_t)j].data()));
  std::vector<T> h(basis.size());
  for (size_t j = 0; j < basis.size(); ++j) h[j] = dot(basis[j], w, n);
  for (size_t j = 0; j < basis.size(); ++j)
    for (int64_t i = 0; i < n; ++i) w[i] -= h[j] * basis[j][i];
}

template <class T> void do_step(llz_krylov_t kry, const Job& job) {
  const int64_t n = kry->n, k = job.k;
  const T* x = reinterpret_cast<const T*>(kry->cols[(size_t)k - 1].data());
  kry->cols[(size_t)k].assign((size_t)n * sizeof(T), 0);
  T* y = reinterpret_cast<T*>(kry->cols[(size_t)k].data());
  job.op->fn(job.op->user, x, y, n, nullptr);  // y arrives zero-filled: serves both callback contracts
  for (int64_t i = 0; i < n; ++i) y[i] += job.sigma * x[i];
  const T* u2 = k >= 2 ? reinterpret_cast<const T*>(kry->cols[(size_t)k - 2].data()) : nullptr;
  double alpha, beta, wn;
  if (job.orth == LLZ_ORTH_RECURRENCE_LAZY) {
    const double s1 = k >= 2 ? kry->beta[(size_t)k - 2] : 1.0, s2 = k >= 3 ? kry->beta[(size_t)k - 3] : 1.0;
    alpha = re_dot(x, y, n) / (s1 * s1);
    for (int64_t i = 0; i < n; ++i) y[i] = y[i] / s1 - (alpha / s1) * x[i] - (u2 ? (s1 / s2) * u2[i] : T(0));
    beta = wn = std::sqrt(re_dot(y, y, n));  // column k stays un-normalised
  } else {
    alpha = re_dot(x, y, n);
    const double bprev = k >= 2 ? kry->beta[(size_t)k - 2] : 0.0;
    for (int64_t i = 0; i < n; ++i) y[i] -= alpha * x[i] + (u2 ? bprev * u2[i] : T(0));
    wn = std::sqrt(re_dot(y, y, n));
    if (job.orth == LLZ_ORTH_FULL || job.orth == LLZ_ORTH_FULL_TWICE) gram_schmidt(kry, y, k);
    if (job.orth == LLZ_ORTH_FULL_TWICE) gram_schmidt(kry, y, k);
    beta = std::sqrt(re_dot(y, y, n));
    if (beta > 0)
      for (int64_t i = 0; i < n; ++i) y[i] /= beta;
  }
  kry->alpha[(size_t)k - 1] = alpha;
  kry->beta[(size_t)k - 1] = beta;
  kry->wnorm[(size_t)k - 1] = wn;
}

void worker_loop(llz_krylov_t kry) {
  for (;;) {
    Job job;
    {
      std::unique_lock<std::mutex> lk(kry->mu);
      kry->cv.wait(lk, [&] { return kry->quit || !kry->q.empty(); });
      if (kry->quit && kry->q.empty()) return;
      job = kry->q.front();
      kry->q.pop_front();
      kry->busy = true;
    }
    if (kry->delay_us > 0) std::this_thread::sleep_for(std::chrono::microseconds(kry->rng() % (unsigned)kry->delay_us));
    if (kry->dtype == LLZ_F64)
      do_step<double>(kry, job);
    else
      do_step<cd>(kry, job);
    {
      std::lock_guard<std::mutex> lk(kry->mu);
      kry->busy = false;
      kry->k_done.store(job.k, std::memory_order_release);
    }
    kry->cv.notify_all();
  }
}

void drain(llz_krylov_t kry) {  // cudaStreamSynchronize
  std::unique_lock<std::mutex> lk(kry->mu);
  kry->cv.wait(lk, [&] { return kry->q.empty() && !kry->busy; });
}

}  // namespace

extern "C" {

int llz_version(void) { return 100; }
const char* llz_status_string(int s) { return s == LLZ_OK ? "ok" : s == LLZ_ERR_OOM ? "out of device memory" : "error"; }
const char* llz_last_error(void) { return g_error; }

int llz_ctx_create(int, llz_ctx_t* ctx) {
  *ctx = new llz_ctx_s();
  if (const char* e = getenv("MOCK_NRANKS")) (*ctx)->nranks = atoi(e);
  return LLZ_OK;
}
int llz_ctx_destroy(llz_ctx_t ctx) {
  delete ctx;
  return LLZ_OK;
}
int llz_ctx_synchronize(llz_ctx_t) { return LLZ_OK; }
int llz_ctx_memcpy(llz_ctx_t, void* dst, const void* src, size_t bytes, int) {
  memcpy(dst, src, bytes);
  return LLZ_OK;
}
int llz_ctx_launch_count(llz_ctx_t ctx, uint64_t* c) {
  *c = ctx->launches.load();
  return LLZ_OK;
}
int llz_comm_unique_id(void* id) {
  memset(id, 0, 128);
  return LLZ_OK;
}
int llz_ctx_join(llz_ctx_t ctx, int rank, int nranks, const void*) {
  ctx->rank = rank;
  ctx->nranks = nranks;
  return LLZ_OK;
}
int llz_ctx_rank(llz_ctx_t ctx, int* rank, int* nranks) {
  *rank = ctx->rank;
  *nranks = ctx->nranks;
  return LLZ_OK;
}

int llz_op_create_callback(llz_ctx_t ctx, int dtype, int64_t n, llz_apply_fn fn, void* user, int overwrites, llz_op_t* op) {
  if (!esize(dtype)) return fail(LLZ_ERR_UNSUPPORTED, "mock: double / complex<double> only");
  *op = new llz_op_s{ctx, dtype, n, fn, user, overwrites};
  return LLZ_OK;
}
int llz_op_create_csr(llz_ctx_t, int, int64_t, int64_t, int64_t, const int64_t*, const int32_t*, const void*, int, llz_op_t*) {
  return fail(LLZ_ERR_UNSUPPORTED, "mock: callback operators only");
}
int llz_op_create_sell(llz_ctx_t, int, int64_t, int64_t, int64_t, const int64_t*, const int32_t*, const void*, int, int, llz_op_t*) {
  return fail(LLZ_ERR_UNSUPPORTED, "mock: callback operators only");
}
int llz_op_create_xxz(llz_ctx_t, int, int, int, double, double, int, llz_op_t*) { return fail(LLZ_ERR_UNSUPPORTED, "mock"); }
int llz_op_destroy(llz_op_t op) {
  delete op;
  return LLZ_OK;
}
int llz_op_shape(llz_op_t op, int64_t* nl, int64_t* ng, int64_t* r0) {
  if (nl) *nl = op->n;
  if (ng) *ng = op->n;
  if (r0) *r0 = 0;
  return LLZ_OK;
}
int llz_op_bytes(llz_op_t, int64_t* b) {
  *b = 0;
  return LLZ_OK;
}
int llz_op_gerschgorin_radius(llz_op_t, double*) { return fail(LLZ_ERR_UNSUPPORTED, "mock"); }
int llz_op_apply(llz_op_t op, llz_vec_t x, llz_vec_t y) {
  std::fill(y->buf.begin(), y->buf.end(), 0);
  return op->fn(op->user, x->buf.data(), y->buf.data(), op->n, nullptr) == 0 ? LLZ_OK : fail(LLZ_ERR_USER, "callback failed");
}

int llz_vec_create(llz_ctx_t ctx, int dtype, int64_t n, llz_vec_t* v) {
  if (!esize(dtype)) return fail(LLZ_ERR_UNSUPPORTED, "mock: double / complex<double> only");
  *v = new llz_vec_s{ctx, dtype, n, std::vector<char>((size_t)n * esize(dtype), 0)};
  return LLZ_OK;
}
int llz_vec_destroy(llz_vec_t v) {
  delete v;
  return LLZ_OK;
}
int llz_vec_upload(llz_vec_t v, const void* host) {
  memcpy(v->buf.data(), host, v->buf.size());
  return LLZ_OK;
}
int llz_vec_download(llz_vec_t v, void* host) {
  memcpy(host, v->buf.data(), v->buf.size());
  return LLZ_OK;
}
int llz_vec_device_ptr(llz_vec_t v, void** dev) {
  *dev = v->buf.data();
  return LLZ_OK;
}
int llz_vec_copy(llz_vec_t d, llz_vec_t s) {
  d->buf = s->buf;
  return LLZ_OK;
}
int llz_vec_fill_zero(llz_vec_t v) {
  std::fill(v->buf.begin(), v->buf.end(), 0);
  return LLZ_OK;
}
int llz_vec_dot(llz_vec_t a, llz_vec_t b, double out[2]) {
  if (a->dtype == LLZ_F64) {
    out[0] = dot(reinterpret_cast<const double*>(a->buf.data()), reinterpret_cast<const double*>(b->buf.data()), a->n);
    out[1] = 0;
  } else {
    const cd z = dot(reinterpret_cast<const cd*>(a->buf.data()), reinterpret_cast<const cd*>(b->buf.data()), a->n);
    out[0] = z.real();
    out[1] = z.imag();
  }
  return LLZ_OK;
}
int llz_vec_norm(llz_vec_t v, double* out) {
  double d[2];
  llz_vec_dot(v, v, d);
  *out = std::sqrt(d[0]);
  return LLZ_OK;
}
int llz_vec_m_norm(llz_vec_t v, double* out) {
  const double* p = reinterpret_cast<const double*>(v->buf.data());
  double s = 0;
  for (size_t i = 0; i < v->buf.size() / 8; ++i) s += std::fabs(p[i]);
  *out = s;
  return LLZ_OK;
}
int llz_vec_scale(llz_vec_t v, const double a[2]) {
  if (v->dtype == LLZ_F64) {
    double* p = reinterpret_cast<double*>(v->buf.data());
    for (int64_t i = 0; i < v->n; ++i) p[i] *= a[0];
  } else {
    cd* p = reinterpret_cast<cd*>(v->buf.data());
    for (int64_t i = 0; i < v->n; ++i) p[i] *= cd(a[0], a[1]);
  }
  return LLZ_OK;
}
int llz_vec_normalize(llz_vec_t v, double* norm_out) {
  double nrm;
  llz_vec_norm(v, &nrm);
  if (norm_out) *norm_out = nrm;
  const double inv[2] = {1.0 / nrm, 0.0};
  return llz_vec_scale(v, inv);
}
int llz_vec_axpy(llz_vec_t y, const double a[2], llz_vec_t x) {
  if (y->dtype == LLZ_F64) {
    double* p = reinterpret_cast<double*>(y->buf.data());
    const double* q = reinterpret_cast<const double*>(x->buf.data());
    for (int64_t i = 0; i < y->n; ++i) p[i] += a[0] * q[i];
  } else {
    cd* p = reinterpret_cast<cd*>(y->buf.data());
    const cd* q = reinterpret_cast<const cd*>(x->buf.data());
    for (int64_t i = 0; i < y->n; ++i) p[i] += cd(a[0], a[1]) * q[i];
  }
  return LLZ_OK;
}
int llz_vec_schmidt_orth(llz_vec_t, const llz_vec_t*, int64_t, int) { return fail(LLZ_ERR_UNSUPPORTED, "mock"); }

int llz_krylov_create(llz_ctx_t ctx, int dtype, int64_t n, int64_t max_cols, llz_krylov_t* out) {
  if (!esize(dtype) || max_cols < 2) return fail(LLZ_ERR_INVALID, "mock krylov_create: bad argument");
  llz_krylov_t k = new llz_krylov_s();
  k->ctx = ctx;
  k->dtype = dtype;
  k->n = n;
  k->cap = max_cols;
  if (const char* e = getenv("MOCK_BASIS_CAPACITY")) k->cap = std::min<int64_t>(k->cap, atoll(e));
  if (const char* e = getenv("MOCK_DELAY_US")) k->delay_us = atoi(e);
  k->cols.resize((size_t)k->cap);
  k->alpha.assign((size_t)k->cap + 2, 0);
  k->beta.assign((size_t)k->cap + 2, 0);
  k->wnorm.assign((size_t)k->cap + 2, 0);
  k->worker = std::thread(worker_loop, k);
  *out = k;
  return LLZ_OK;
}
int llz_krylov_destroy(llz_krylov_t k) {
  if (!k) return LLZ_OK;
  {
    std::lock_guard<std::mutex> lk(k->mu);
    k->quit = true;
  }
  k->cv.notify_all();
  k->worker.join();
  delete k;
  return LLZ_OK;
}
int llz_krylov_capacity(llz_krylov_t k, int64_t* c) {
  *c = k->cap;
  return LLZ_OK;
}
int llz_krylov_set_locked(llz_krylov_t k, const llz_vec_t* locked, int64_t count) {
  drain(k);
  k->locked.assign(locked, locked + count);
  return LLZ_OK;
}
int llz_krylov_begin(llz_krylov_t k, const void* start, int, double* norm_out) {
  drain(k);
  k->k_enq.store(0);
  k->k_done.store(0);
  const size_t bytes = (size_t)k->n * esize(k->dtype);
  k->cols[0].assign(reinterpret_cast<const char*>(start), reinterpret_cast<const char*>(start) + bytes);
  double nrm = 0;
  if (k->dtype == LLZ_F64) {
    double* u = reinterpret_cast<double*>(k->cols[0].data());
    gram_schmidt(k, u, 0);
    gram_schmidt(k, u, 0);
    nrm = std::sqrt(re_dot(u, u, k->n));
    for (int64_t i = 0; i < k->n; ++i) u[i] /= nrm;
  } else {
    cd* u = reinterpret_cast<cd*>(k->cols[0].data());
    gram_schmidt(k, u, 0);
    gram_schmidt(k, u, 0);
    nrm = std::sqrt(re_dot(u, u, k->n));
    for (int64_t i = 0; i < k->n; ++i) u[i] /= nrm;
  }
  if (norm_out) *norm_out = nrm;
  std::lock_guard<std::mutex> lk(g_stat_mu);
  g_steps_per_run.push_back(0);
  return LLZ_OK;
}
int llz_krylov_step(llz_krylov_t k, llz_op_t op, double sigma, int orth) {
  const int64_t kk = k->k_enq.load() + 1;
  if (kk + 1 > k->cap) return fail(LLZ_ERR_OOM, "mock: Krylov basis full (%lld columns)", (long long)k->cap);
  {
    std::lock_guard<std::mutex> lk(k->mu);
    k->q.push_back(Job{kk, op, sigma, orth});
    k->k_enq.store(kk);
  }
  k->cv.notify_all();
  k->ctx->launches++;
  std::lock_guard<std::mutex> lk(g_stat_mu);
  if (!g_steps_per_run.empty()) g_steps_per_run.back()++;
  return LLZ_OK;
}
int llz_krylov_fetch(llz_krylov_t k, int64_t kk, double* alpha, double* beta, double* wnorm) {
  if (kk < 1 || kk > k->k_enq) return fail(LLZ_ERR_INVALID, "mock krylov_fetch: iteration %lld not enqueued", (long long)kk);
  while (k->k_done.load(std::memory_order_acquire) < kk) std::this_thread::yield();
  if (alpha) *alpha = k->alpha[(size_t)kk - 1];
  if (beta) *beta = k->beta[(size_t)kk - 1];
  if (wnorm) *wnorm = k->wnorm[(size_t)kk - 1];
  return LLZ_OK;
}
int llz_krylov_refine(llz_krylov_t k, int64_t kk, double* shrink) {
  if (kk < 1 || kk > k->k_enq) return fail(LLZ_ERR_INVALID, "mock krylov_refine: iteration %lld not enqueued", (long long)kk);
  drain(k);
  double nu = 1.0;
  if (k->dtype == LLZ_F64) {
    double* u = reinterpret_cast<double*>(k->cols[(size_t)kk].data());
    gram_schmidt(k, u, kk);
    nu = std::sqrt(re_dot(u, u, k->n));
    for (int64_t i = 0; i < k->n; ++i) u[i] /= nu;
  } else {
    cd* u = reinterpret_cast<cd*>(k->cols[(size_t)kk].data());
    gram_schmidt(k, u, kk);
    nu = std::sqrt(re_dot(u, u, k->n));
    for (int64_t i = 0; i < k->n; ++i) u[i] /= nu;
  }
  k->beta[(size_t)kk - 1] *= nu;
  k->k_enq.store(kk);  // iterations enqueued beyond kk used the un-refined vector: dropped
  k->k_done.store(kk);
  if (shrink) *shrink = nu;
  return LLZ_OK;
}
int llz_krylov_steps(llz_krylov_t k, int64_t* out) {
  *out = k->k_enq.load();
  return LLZ_OK;
}
int llz_krylov_combine(llz_krylov_t k, int64_t m, int64_t nvec, const void* coeff, int normalize, const llz_vec_t* out) {
  drain(k);
  if (m < 1 || m > k->k_enq + 1) return fail(LLZ_ERR_INVALID, "mock krylov_combine: m = %lld", (long long)m);
  for (int64_t r = 0; r < nvec; ++r) {
    std::fill(out[r]->buf.begin(), out[r]->buf.end(), 0);
    if (k->dtype == LLZ_F64) {
      double* o = reinterpret_cast<double*>(out[r]->buf.data());
      const double* c = reinterpret_cast<const double*>(coeff) + r * m;
      for (int64_t j = 0; j < m; ++j) {
        const double* u = reinterpret_cast<const double*>(k->cols[(size_t)j].data());
        for (int64_t i = 0; i < k->n; ++i) o[i] += c[j] * u[i];
      }
    } else {
      cd* o = reinterpret_cast<cd*>(out[r]->buf.data());
      const cd* c = reinterpret_cast<const cd*>(coeff) + r * m;
      for (int64_t j = 0; j < m; ++j) {
        const cd* u = reinterpret_cast<const cd*>(k->cols[(size_t)j].data());
        for (int64_t i = 0; i < k->n; ++i) o[i] += c[j] * u[i];
      }
    }
    if (normalize) llz_vec_normalize(out[r], nullptr);
  }
  return LLZ_OK;
}

// ---- read-back for the tests ----
int64_t mock_runs(void) {
  std::lock_guard<std::mutex> lk(g_stat_mu);
  return (int64_t)g_steps_per_run.size();
}
int64_t mock_steps_of_run(int64_t r) {
  std::lock_guard<std::mutex> lk(g_stat_mu);
  return g_steps_per_run[(size_t)r];
}
void mock_reset_stats(void) {
  std::lock_guard<std::mutex> lk(g_stat_mu);
  g_steps_per_run.clear();
}

}  // extern "C"
