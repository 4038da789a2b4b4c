// llz_ctx.cu — contexts, error reporting, device vectors and the stand-alone util:: vector operations.
#include <cstdarg>
#include <cstdio>
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "llz_launch.hpp"

namespace llz {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int fail(int status, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
  return status;
}

cudaError_t dev_malloc(llz_ctx_t ctx, void** p, size_t bytes) {
  bytes = (std::max<size_t>(bytes, 1) + 255) / 256 * 256;
  auto it = ctx->vec_pool.find(bytes);
  if (it != ctx->vec_pool.end()) {
    *p = it->second;
    ctx->vec_pool.erase(it);
    ctx->vec_pool_bytes -= bytes;
    ctx->dev_sizes[*p] = bytes;
    return cudaSuccess;
  }
  cudaError_t e = cudaMalloc(p, bytes);
  if (e == cudaErrorMemoryAllocation && (ctx->cached_krylov || !ctx->vec_pool.empty())) {
    cudaGetLastError();
    cudaStreamSynchronize(ctx->stream);
    ctx_trim(ctx);
    e = cudaMalloc(p, bytes);
  }
  if (e == cudaSuccess) ctx->dev_sizes[*p] = bytes;
  return e;
}

void dev_free(llz_ctx_t ctx, void* p) {
  if (!p) return;
  auto it = ctx->dev_sizes.find(p);
  if (it == ctx->dev_sizes.end()) {  // not ours: plain free
    cudaStreamSynchronize(ctx->stream);
    cudaFree(p);
    return;
  }
  const size_t bytes = it->second;
  ctx->dev_sizes.erase(it);
  ctx_free(ctx, p, bytes);
}

int ctx_alloc(llz_ctx_t ctx, size_t bytes, void** out) {
  auto it = ctx->vec_pool.find(bytes);
  if (it != ctx->vec_pool.end()) {
    *out = it->second;
    ctx->vec_pool.erase(it);
    ctx->vec_pool_bytes -= bytes;
    return LLZ_OK;
  }
  cudaError_t e = cudaMalloc(out, bytes);
  if (e == cudaErrorMemoryAllocation) {  // give the cached buffers back and try once more
    cudaGetLastError();
    cudaStreamSynchronize(ctx->stream);
    ctx_trim(ctx);
    e = cudaMalloc(out, bytes);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(e == cudaErrorMemoryAllocation ? LLZ_ERR_OOM : LLZ_ERR_CUDA, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
  }
  return LLZ_OK;
}

void ctx_free(llz_ctx_t ctx, void* p, size_t bytes) {
  if (!p) return;
  if (ctx->vec_pool_bytes + bytes <= ctx->vec_pool_limit) {
    ctx->vec_pool.emplace(bytes, p);
    ctx->vec_pool_bytes += bytes;
    return;
  }
  cudaStreamSynchronize(ctx->stream);
  cudaFree(p);
}

void ctx_trim(llz_ctx_t ctx) {
  for (auto& kv : ctx->vec_pool) cudaFree(kv.second);
  ctx->vec_pool.clear();
  ctx->vec_pool_bytes = 0;
  if (ctx->cached_krylov) {
    llz_krylov_t k = ctx->cached_krylov;
    ctx->cached_krylov = nullptr;
    krylov_destroy_now(k);
  }
}

}  // namespace llz

using namespace llz;

extern "C" {

int llz_version(void) { return LLZ_VERSION; }

const char* llz_status_string(int status) {
  switch (status) {
    case LLZ_OK: return "ok";
    case LLZ_ERR_INVALID: return "invalid argument";
    case LLZ_ERR_CUDA: return "CUDA error";
    case LLZ_ERR_OOM: return "out of device memory";
    case LLZ_ERR_COMM: return "communication error";
    case LLZ_ERR_UNSUPPORTED: return "unsupported";
    case LLZ_ERR_NO_DEVICE: return "no CUDA device (this engine has no CPU path)";
    case LLZ_ERR_USER: return "user callback failed";
  }
  return "unknown status";
}

const char* llz_last_error(void) { return g_error; }

int llz_ctx_create_on_stream(int device, void* cuda_stream, llz_ctx_t* out) {
  if (!out) return fail(LLZ_ERR_INVALID, "ctx_create: null output");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(LLZ_ERR_NO_DEVICE, "no usable CUDA device (%s); the engine has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  if (device < 0 || device >= count) return fail(LLZ_ERR_INVALID, "device %d out of range [0,%d)", device, count);
  LLZ_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  LLZ_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail(LLZ_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                prop.minor);
  llz_ctx_t ctx = new llz_ctx_s();
  ctx->device = device;
  ctx->num_sms = prop.multiProcessorCount;
  ctx->l2_bytes = (size_t)prop.l2CacheSize;
  ctx->vec_pool_limit = (size_t)prop.totalGlobalMem / 4;
  {
    const char* env = getenv("LLZ_FUSED_ORTH");
    ctx->fuse_orth = !(env && env[0] == '0');
  }
  if (cuda_stream) {
    ctx->stream = (cudaStream_t)cuda_stream;
    ctx->own_stream = false;
  } else {
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
      delete ctx;
      return fail(LLZ_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
    }
    ctx->own_stream = true;
  }
  e = cudaMalloc(&ctx->d_partials, sizeof(double) * kMaxGrid * 2);
  if (e == cudaSuccess) e = cudaMalloc(&ctx->d_result, sizeof(double) * 8);
  if (e == cudaSuccess) e = cudaHostAlloc(&ctx->h_result, sizeof(double) * 8, cudaHostAllocDefault);
  if (e != cudaSuccess) {  // give back what was created so far
    if (ctx->d_partials) cudaFree(ctx->d_partials);
    if (ctx->d_result) cudaFree(ctx->d_result);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return fail(e == cudaErrorMemoryAllocation ? LLZ_ERR_OOM : LLZ_ERR_CUDA, "ctx_create: %s", cudaGetErrorString(e));
  }
  *out = ctx;
  return LLZ_OK;
}

int llz_ctx_create(int device, llz_ctx_t* out) { return llz_ctx_create_on_stream(device, nullptr, out); }

int llz_ctx_synchronize(llz_ctx_t ctx) {
  if (!ctx) return fail(LLZ_ERR_INVALID, "null ctx");
  LLZ_CUDA(cudaStreamSynchronize(ctx->stream));
  return LLZ_OK;
}

int llz_ctx_memcpy(llz_ctx_t ctx, void* dst, const void* src, size_t bytes, int to_device) {
  if (!ctx || (bytes > 0 && (!dst || !src))) return fail(LLZ_ERR_INVALID, "ctx_memcpy: null");
  LLZ_CUDA(cudaSetDevice(ctx->device));
  LLZ_CUDA(cudaMemcpyAsync(dst, src, bytes, to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, ctx->stream));
  LLZ_CUDA(cudaStreamSynchronize(ctx->stream));
  return LLZ_OK;
}

int llz_ctx_stream(llz_ctx_t ctx, void** s) {
  if (!ctx || !s) return fail(LLZ_ERR_INVALID, "null argument");
  *s = (void*)ctx->stream;
  return LLZ_OK;
}

int llz_ctx_release_cache(llz_ctx_t ctx) {
  if (!ctx) return fail(LLZ_ERR_INVALID, "null ctx");
  LLZ_CUDA(cudaSetDevice(ctx->device));
  LLZ_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx_trim(ctx);
  return LLZ_OK;
}

int llz_ctx_launch_count(llz_ctx_t ctx, uint64_t* count) {
  if (!ctx || !count) return fail(LLZ_ERR_INVALID, "null argument");
  *count = ctx->launches;
  return LLZ_OK;
}

int llz_ctx_profile(llz_ctx_t ctx, int enable) {
  if (!ctx) return fail(LLZ_ERR_INVALID, "null ctx");
  LLZ_CUDA(cudaStreamSynchronize(ctx->stream));
  for (auto& kv : ctx->prof) {
    for (auto& p : kv.second.pending) {
      ctx->event_pool.push_back(p.first);
      ctx->event_pool.push_back(p.second);
    }
    kv.second = ProfEntry();
  }
  ctx->profile = enable != 0;
  return LLZ_OK;
}

int llz_ctx_profile_read(llz_ctx_t ctx, const char* name, double* ms, int64_t* launches, double* bytes) {
  if (!ctx || !name) return fail(LLZ_ERR_INVALID, "null");
  LLZ_CUDA(cudaStreamSynchronize(ctx->stream));
  double total = 0.0, by = 0.0;
  int64_t count = 0;
  auto it = ctx->prof.find(name);
  if (it != ctx->prof.end()) {
    for (auto& p : it->second.pending) {
      float t = 0.f;
      if (cudaEventElapsedTime(&t, p.first, p.second) == cudaSuccess) it->second.ms += t;
      ctx->event_pool.push_back(p.first);
      ctx->event_pool.push_back(p.second);
    }
    it->second.pending.clear();
    total = it->second.ms;
    count = it->second.launches;
    by = it->second.bytes;
  }
  if (ms) *ms = total;
  if (launches) *launches = count;
  if (bytes) *bytes = by;
  return LLZ_OK;
}

int llz_ctx_rank(llz_ctx_t ctx, int* rank, int* nranks) {
  if (!ctx) return fail(LLZ_ERR_INVALID, "null ctx");
  if (rank) *rank = ctx->rank;
  if (nranks) *nranks = ctx->nranks;
  return LLZ_OK;
}

// ---------------------------------------------------------------------------------------------------------------
int llz_vec_create(llz_ctx_t ctx, int dtype, int64_t n, llz_vec_t* out) {
  if (!ctx || !out || n < 0 || dtype_size(dtype) == 0) return fail(LLZ_ERR_INVALID, "vec_create: bad argument");
  LLZ_CUDA(cudaSetDevice(ctx->device));
  llz_vec_t v = new llz_vec_s();
  v->ctx = ctx;
  v->dtype = dtype;
  v->n = n;
  const size_t bytes = ((size_t)n * dtype_size(dtype) + 255) / 256 * 256 + 256;
  const int st = ctx_alloc(ctx, bytes, &v->d);
  if (st != LLZ_OK) {
    delete v;
    return st;
  }
  *out = v;
  return LLZ_OK;
}

int llz_vec_destroy(llz_vec_t v) {
  if (!v) return LLZ_OK;
  if (v->owned && v->d) ctx_free(v->ctx, v->d, ((size_t)v->n * dtype_size(v->dtype) + 255) / 256 * 256 + 256);
  delete v;
  return LLZ_OK;
}

int llz_vec_upload(llz_vec_t v, const void* host) {
  if (!v || !host) return fail(LLZ_ERR_INVALID, "vec_upload: null");
  LLZ_CUDA(cudaMemcpyAsync(v->d, host, (size_t)v->n * dtype_size(v->dtype), cudaMemcpyHostToDevice, v->ctx->stream));
  LLZ_CUDA(cudaStreamSynchronize(v->ctx->stream));
  return LLZ_OK;
}

int llz_vec_download(llz_vec_t v, void* host) {
  if (!v || !host) return fail(LLZ_ERR_INVALID, "vec_download: null");
  LLZ_CUDA(cudaMemcpyAsync(host, v->d, (size_t)v->n * dtype_size(v->dtype), cudaMemcpyDeviceToHost, v->ctx->stream));
  LLZ_CUDA(cudaStreamSynchronize(v->ctx->stream));
  return LLZ_OK;
}

int llz_vec_device_ptr(llz_vec_t v, void** dev) {
  if (!v || !dev) return fail(LLZ_ERR_INVALID, "null");
  *dev = v->d;
  return LLZ_OK;
}

int llz_vec_copy(llz_vec_t dst, llz_vec_t src) {
  if (!dst || !src || dst->n != src->n || dst->dtype != src->dtype) return fail(LLZ_ERR_INVALID, "vec_copy: mismatch");
  LLZ_CUDA(cudaMemcpyAsync(dst->d, src->d, (size_t)src->n * dtype_size(src->dtype), cudaMemcpyDeviceToDevice,
                           dst->ctx->stream));
  return LLZ_OK;
}

int llz_vec_fill_zero(llz_vec_t v) {
  if (!v) return fail(LLZ_ERR_INVALID, "null");
  LLZ_CUDA(cudaMemsetAsync(v->d, 0, (size_t)v->n * dtype_size(v->dtype), v->ctx->stream));
  return LLZ_OK;
}

static int dot_to_host(llz_ctx_t ctx, int dtype, const void* a, const void* b, int64_t n, double out[2]) {
  int grid = 0;
  const int nc = dtype_nc(dtype);
  LLZ_TRY(launch_dot(ctx, dtype, a, b, n, ctx->d_partials, &grid));
  LLZ_TRY(launch_sum_partials(ctx, ctx->d_partials, grid, nc, ctx->d_result, nullptr));
  LLZ_TRY(comm_allreduce_sum(ctx, ctx->d_result, nc));
  LLZ_CUDA(cudaMemcpyAsync(ctx->h_result, ctx->d_result, sizeof(double) * 2, cudaMemcpyDeviceToHost, ctx->stream));
  LLZ_CUDA(cudaStreamSynchronize(ctx->stream));
  out[0] = ctx->h_result[0];
  out[1] = nc == 2 ? ctx->h_result[1] : 0.0;
  return LLZ_OK;
}

int llz_vec_dot(llz_vec_t a, llz_vec_t b, double out[2]) {
  if (!a || !b || !out || a->n != b->n || a->dtype != b->dtype) return fail(LLZ_ERR_INVALID, "vec_dot: mismatch");
  return dot_to_host(a->ctx, a->dtype, a->d, b->d, a->n, out);
}

int llz_vec_norm(llz_vec_t v, double* out) {
  if (!v || !out) return fail(LLZ_ERR_INVALID, "null");
  double d[2] = {0.0, 0.0};
  LLZ_TRY(dot_to_host(v->ctx, v->dtype, v->d, v->d, v->n, d));
  *out = sqrt(d[0]);
  return LLZ_OK;
}

int llz_vec_m_norm(llz_vec_t v, double* out) {
  if (!v || !out) return fail(LLZ_ERR_INVALID, "null");
  llz_ctx_t ctx = v->ctx;
  int grid = 0;
  LLZ_TRY(launch_asum(ctx, v->dtype, v->d, v->n, ctx->d_partials, &grid));
  LLZ_TRY(launch_sum_partials(ctx, ctx->d_partials, grid, 1, ctx->d_result, nullptr));
  LLZ_TRY(comm_allreduce_sum(ctx, ctx->d_result, 1));
  LLZ_CUDA(cudaMemcpyAsync(ctx->h_result, ctx->d_result, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  LLZ_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = ctx->h_result[0];
  return LLZ_OK;
}

int llz_vec_scale(llz_vec_t v, const double a[2]) {
  if (!v || !a) return fail(LLZ_ERR_INVALID, "null");
  return launch_scale(v->ctx, v->dtype, v->d, v->n, a);
}

int llz_vec_normalize(llz_vec_t v, double* norm_out) {
  if (!v) return fail(LLZ_ERR_INVALID, "null");
  double nrm = 0.0;
  LLZ_TRY(llz_vec_norm(v, &nrm));
  if (norm_out) *norm_out = nrm;
  // T(1)/norm in the working precision, as util::normalize does (linear_algebra.hpp:78-80)
  const bool single = (v->dtype == LLZ_F32 || v->dtype == LLZ_C64);
  const double inv[2] = {single ? (double)(1.0f / (float)nrm) : 1.0 / nrm, 0.0};
  return launch_scale(v->ctx, v->dtype, v->d, v->n, inv);
}

int llz_vec_axpy(llz_vec_t y, const double a[2], llz_vec_t x) {
  if (!y || !x || !a || y->n != x->n || y->dtype != x->dtype) return fail(LLZ_ERR_INVALID, "vec_axpy: mismatch");
  return launch_axpy(y->ctx, y->dtype, y->d, a, x->d, y->n);
}

int llz_vec_schmidt_orth(llz_vec_t w, const llz_vec_t* basis, int64_t count, int passes) {
  if (!w || count < 0 || (count > 0 && !basis)) return fail(LLZ_ERR_INVALID, "schmidt_orth: bad argument");
  if (count == 0) return LLZ_OK;
  llz_ctx_t ctx = w->ctx;
  const int nc = dtype_nc(w->dtype);
  if (passes < 1) passes = 1;
  // pointer table + coefficient / partial buffers, grown on demand
  if (count > ctx->ptr_capacity) {
    if (ctx->d_ptrs) cudaFree(ctx->d_ptrs);
    if (ctx->d_coef) cudaFree(ctx->d_coef);
    ctx->d_ptrs = nullptr;
    ctx->d_coef = nullptr;
    int64_t cap = count < 64 ? 64 : count * 2;
    LLZ_CUDA(cudaMalloc(&ctx->d_ptrs, sizeof(void*) * cap));
    LLZ_CUDA(cudaMalloc(&ctx->d_coef, sizeof(double) * 2 * cap));
    ctx->ptr_capacity = cap;
  }
  std::vector<const void*> ptrs((size_t)count);
  for (int64_t j = 0; j < count; ++j) {
    if (!basis[j] || basis[j]->n != w->n || basis[j]->dtype != w->dtype)
      return fail(LLZ_ERR_INVALID, "schmidt_orth: basis vector %lld mismatched", (long long)j);
    ptrs[(size_t)j] = basis[j]->d;
  }
  LLZ_CUDA(cudaMemcpyAsync(ctx->d_ptrs, ptrs.data(), sizeof(void*) * count, cudaMemcpyHostToDevice, ctx->stream));
  LLZ_CUDA(cudaStreamSynchronize(ctx->stream));  // ptrs is a stack-lifetime staging buffer
  const int chunk = max_project_cols(w->dtype);
  const size_t need = (size_t)kMaxGrid * ((size_t)(count < chunk ? count : chunk) * nc + 1);
  if (need > ctx->ph_capacity) {
    if (ctx->d_ph) cudaFree(ctx->d_ph);
    ctx->d_ph = nullptr;
    LLZ_CUDA(cudaMalloc(&ctx->d_ph, sizeof(double) * need));
    ctx->ph_capacity = need;
  }
  ColumnSet cs;
  cs.Q = (const void* const*)ctx->d_ptrs;
  cs.nq = (int)count;
  Fold nofold;
  for (int p = 0; p < passes; ++p) {
    for (int c0 = 0; c0 < count; c0 += chunk) {
      const int nc_cols = (int)((count - c0) < chunk ? (count - c0) : chunk);
      int grid = 0;
      LLZ_TRY(launch_project(ctx, w->dtype, cs, c0, nc_cols, w->d, w->n, nofold, ctx->d_ph, &grid));
      LLZ_TRY(launch_reduce(ctx, w->dtype, ctx->d_ph, grid, c0, nc_cols, ctx->d_coef, nullptr));
    }
    LLZ_TRY(comm_allreduce_sum(ctx, ctx->d_coef, (int)count * nc));
    const int uchunk = max_update_cols(w->dtype);
    for (int c0 = 0; c0 < count; c0 += uchunk) {
      const int nc_cols = (int)((count - c0) < uchunk ? (count - c0) : uchunk);
      int grid = 0;
      LLZ_TRY(launch_update(ctx, w->dtype, cs, c0, nc_cols, w->d, w->d, w->n, ctx->d_coef, nofold, nullptr, &grid));
    }
  }
  return LLZ_OK;
}

int llz_ctx_destroy(llz_ctx_t ctx) {
  if (!ctx) return LLZ_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  ctx_trim(ctx);
  comm_destroy(ctx);
  ctx_trim(ctx);  // the exchange buffers comm_destroy handed back to the pool
  if (ctx->d_bar) cudaFree(ctx->d_bar);
  if (ctx->d_partials) cudaFree(ctx->d_partials);
  if (ctx->d_result) cudaFree(ctx->d_result);
  if (ctx->h_result) cudaFreeHost(ctx->h_result);
  if (ctx->d_ptrs) cudaFree(ctx->d_ptrs);
  if (ctx->d_coef) cudaFree(ctx->d_coef);
  if (ctx->d_ph) cudaFree(ctx->d_ph);
  for (auto& kv : ctx->prof)
    for (auto& p : kv.second.pending) {
      cudaEventDestroy(p.first);
      cudaEventDestroy(p.second);
    }
  for (auto e : ctx->event_pool) cudaEventDestroy(e);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return LLZ_OK;
}

}  // extern "C"
