// llz_ops.cu — device operators: the replacement of the reference's `mv_mul` std::function
// (lambda_lanczos.hpp:120-126, exponentiator.hpp:35-41).  Built-in CSR SpMV with the alpha = Re<x, A x> dot fused
// into its epilogue, and the user callback adapter.  (The matrix-free XXZ operator lives in llz_xxz.cu.)
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "llz_device.cuh"
#include "llz_launch.hpp"

namespace llz {

// ------------------------------------------------------------------------------------------------------------------
// CSR SpMV: LPR lanes cooperate on one row (LPR in {1,2,4,8,16,32} chosen from the mean row length), persistent
// grid-stride over row blocks.  y_i = sum_p a_p x[col_p] + sigma x_i; the CTA's partial of Re(conj(x_i) y_i) goes to
// pa[blockIdx.x] (fixed order => reproducible alpha).
// ------------------------------------------------------------------------------------------------------------------
template <class T> __device__ __forceinline__ T shfl_down_t(T v, int o, int width);
template <> __device__ __forceinline__ float shfl_down_t(float v, int o, int w) { return __shfl_down_sync(0xffffffffu, v, o, w); }
template <> __device__ __forceinline__ double shfl_down_t(double v, int o, int w) { return __shfl_down_sync(0xffffffffu, v, o, w); }
template <> __device__ __forceinline__ float2 shfl_down_t(float2 v, int o, int w) {
  return make_float2(__shfl_down_sync(0xffffffffu, v.x, o, w), __shfl_down_sync(0xffffffffu, v.y, o, w));
}
template <> __device__ __forceinline__ double2 shfl_down_t(double2 v, int o, int w) {
  return make_double2(__shfl_down_sync(0xffffffffu, v.x, o, w), __shfl_down_sync(0xffffffffu, v.y, o, w));
}

template <class T, class IDX, int LPR>
__global__ void __launch_bounds__(kThreads, 4)
    k_csr_spmv_dot(const IDX* __restrict__ rowptr, const int32_t* __restrict__ colidx, const T* __restrict__ vals,
                   const T* __restrict__ x, T* __restrict__ y, int64_t n, typename Num<T>::R sigma, double* pa) {
  __shared__ double scratch[kWarps];
  constexpr int ROWS = kThreads / LPR;
  const int tid = threadIdx.x;
  const int sub = tid % LPR;
  double dot = 0.0;
  for (int64_t row0 = (int64_t)blockIdx.x * ROWS; row0 < n; row0 += (int64_t)gridDim.x * ROWS) {
    const int64_t row = row0 + tid / LPR;
    T sum = zero_of(T());
    if (row < n) {
      const IDX p1 = rowptr[row + 1];
      for (IDX p = rowptr[row] + sub; p < p1; p += LPR) fmadd(sum, vals[p], __ldg(x + colidx[p]));
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) sum = add_t(sum, shfl_down_t(sum, o, LPR));
    if (sub == 0 && row < n) {
      const T xi = x[row];
      T yi = sum;
      yi = add_t(yi, scale_real(xi, sigma));
      y[row] = yi;
      dot += re_conj_mul(xi, yi);
    }
  }
  const double t = block_sum(dot, scratch);
  if (tid == 0) pa[blockIdx.x] = t;
}

// ------------------------------------------------------------------------------------------------------------------
// CSR "stream" SpMV for short rows (the common case: stencils, lattice Hamiltonians, ~5-30 non-zeros per row).
// A CTA owns R consecutive rows whose non-zeros (<= cap, guaranteed by the host when it picks R) are contiguous in
// memory: (1) all threads stream vals/colidx of that range fully coalesced, gather x and park the products in shared
// memory; (2) one thread per row adds its products in column order (the same order as a sequential CSR loop, so
// y is bit-identical to the reference's sample-style mv_mul); (3) y is written coalesced and the alpha partial
// accumulated.  Matrix traffic is perfectly coalesced whatever the row lengths are.
// ------------------------------------------------------------------------------------------------------------------
template <class T, class IDX>
__global__ void __launch_bounds__(kThreads, 4)
    k_csr_stream_dot(const IDX* __restrict__ rowptr, const int32_t* __restrict__ colidx, const T* __restrict__ vals,
                     const T* __restrict__ x, T* __restrict__ y, int64_t n, typename Num<T>::R sigma, double* pa, int R,
                     int cap) {
  extern __shared__ __align__(16) unsigned char smem_s[];
  T* prod = reinterpret_cast<T*>(smem_s);
  IDX* rp = reinterpret_cast<IDX*>(prod + cap);
  __shared__ double scratch[kWarps];
  const int tid = threadIdx.x;
  double dot = 0.0;
  const int64_t nblocks = (n + R - 1) / R;
  for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
    const int64_t row0 = blk * R;
    const int nrows = (int)((n - row0 < R) ? (n - row0) : R);
    for (int i = tid; i <= nrows; i += kThreads) rp[i] = rowptr[row0 + i];
    __syncthreads();
    const IDX base = rp[0];
    const int cnt = (int)(rp[nrows] - base);
    for (int e = tid; e < cnt; e += kThreads) prod[e] = mul(vals[base + e], __ldg(x + colidx[base + e]));
    __syncthreads();
    for (int r = tid; r < nrows; r += kThreads) {
      const int j1 = (int)(rp[r + 1] - base);
      T s = zero_of(T());
      for (int j = (int)(rp[r] - base); j < j1; ++j) s = add_t(s, prod[j]);
      const T xi = x[row0 + r];
      const T yi = add_t(s, scale_real(xi, sigma));
      y[row0 + r] = yi;
      dot += re_conj_mul(xi, yi);
    }
    __syncthreads();
  }
  const double t = block_sum(dot, scratch);
  if (tid == 0) pa[blockIdx.x] = t;
}

template <class T> struct CsrOp : OpBase {
  int64_t n_cols = 0;
  int64_t nnz = 0;
  bool idx32 = true;
  void* d_rowptr = nullptr;  // int32 or int64
  int32_t* d_colidx = nullptr;
  T* d_vals = nullptr;
  int lpr = 8;
  int stream_rows = 0;  // rows per CTA block of the stream kernel (0 = use the lanes-per-row kernel)
  int stream_cap = 0;   // products parked in shared memory per block

  ~CsrOp() override {
    if (d_rowptr) cudaFree(d_rowptr);
    if (d_colidx) cudaFree(d_colidx);
    if (d_vals) cudaFree(d_vals);
  }

  template <class IDX, int LPR> int launch(const void* x, void* y, double sigma, double* pa, int* npa) {
    constexpr int ROWS = kThreads / LPR;
    int64_t blocks = (n_local + ROWS - 1) / ROWS;
    int64_t g = std::min<int64_t>(blocks, std::min<int64_t>(kMaxGrid, (int64_t)ctx->num_sms * 8));
    if (g < 1) g = 1;
    k_csr_spmv_dot<T, IDX, LPR><<<(int)g, kThreads, 0, ctx->stream>>>(
        (const IDX*)d_rowptr, d_colidx, d_vals, (const T*)x, (T*)y, n_local, (typename Num<T>::R)sigma, pa);
    *npa = (int)g;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(LLZ_ERR_CUDA, "launch k_csr_spmv_dot: %s", cudaGetErrorString(e));
    ctx->launches++;
    return LLZ_OK;
  }

  template <class IDX> int launch_lpr(const void* x, void* y, double sigma, double* pa, int* npa) {
    switch (lpr) {
      case 1: return launch<IDX, 1>(x, y, sigma, pa, npa);
      case 2: return launch<IDX, 2>(x, y, sigma, pa, npa);
      case 4: return launch<IDX, 4>(x, y, sigma, pa, npa);
      case 8: return launch<IDX, 8>(x, y, sigma, pa, npa);
      case 16: return launch<IDX, 16>(x, y, sigma, pa, npa);
      default: return launch<IDX, 32>(x, y, sigma, pa, npa);
    }
  }

  template <class IDX> int launch_stream(const void* x, void* y, double sigma, double* pa, int* npa) {
    const size_t smem = (size_t)stream_cap * sizeof(T) + (size_t)(stream_rows + 1) * sizeof(IDX);
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(k_csr_stream_dot<T, IDX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return fail(LLZ_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    const int64_t blocks = (n_local + stream_rows - 1) / stream_rows;
    int64_t g = std::min<int64_t>(blocks, std::min<int64_t>(kMaxGrid, (int64_t)ctx->num_sms * 8));
    if (g < 1) g = 1;
    k_csr_stream_dot<T, IDX><<<(int)g, kThreads, smem, ctx->stream>>>((const IDX*)d_rowptr, d_colidx, d_vals, (const T*)x, (T*)y,
                                                                     n_local, (typename Num<T>::R)sigma, pa, stream_rows, stream_cap);
    *npa = (int)g;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(LLZ_ERR_CUDA, "launch k_csr_stream_dot: %s", cudaGetErrorString(e));
    ctx->launches++;
    return LLZ_OK;
  }

  int apply_fused(const void* x, void* y, double sigma, double* pa, int* npa) override {
    if (stream_rows > 0) return idx32 ? launch_stream<int32_t>(x, y, sigma, pa, npa) : launch_stream<int64_t>(x, y, sigma, pa, npa);
    return idx32 ? launch_lpr<int32_t>(x, y, sigma, pa, npa) : launch_lpr<int64_t>(x, y, sigma, pa, npa);
  }
};

template <class T>
static int create_csr(llz_ctx_t ctx, int dtype, int64_t n_rows, int64_t n_cols, const int64_t* rowptr,
                      const int32_t* colidx, const void* vals, int host_arrays, llz_op_t* out) {
  auto* op = new CsrOp<T>();
  op->ctx = ctx;
  op->dtype = dtype;
  op->n_local = n_rows;
  op->n_cols = n_cols;
  const cudaMemcpyKind kind = host_arrays ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
  // rowptr[n_rows] = nnz: fetch it to size the arrays
  int64_t first = 0, last = 0;
  if (host_arrays) {
    first = rowptr[0];
    last = rowptr[n_rows];
  } else {
    LLZ_CUDA(cudaMemcpy(&first, rowptr, sizeof(int64_t), cudaMemcpyDeviceToHost));
    LLZ_CUDA(cudaMemcpy(&last, rowptr + n_rows, sizeof(int64_t), cudaMemcpyDeviceToHost));
  }
  if (first != 0 || last < 0) {
    delete op;
    return fail(LLZ_ERR_INVALID, "csr: rowptr must start at 0 (got %lld) and be non-decreasing", (long long)first);
  }
  op->nnz = last;
  op->idx32 = last < (int64_t)0x7fffffff;
  int s = LLZ_OK;
  auto guard = [&](cudaError_t e, const char* what) {
    if (e != cudaSuccess && s == LLZ_OK)
      s = fail(e == cudaErrorMemoryAllocation ? LLZ_ERR_OOM : LLZ_ERR_CUDA, "csr %s: %s", what, cudaGetErrorString(e));
  };
  guard(cudaMalloc(&op->d_colidx, std::max<size_t>(16, sizeof(int32_t) * (size_t)last)), "colidx alloc");
  guard(cudaMalloc(&op->d_vals, std::max<size_t>(16, sizeof(T) * (size_t)last)), "vals alloc");
  if (s == LLZ_OK) {
    guard(cudaMemcpyAsync(op->d_colidx, colidx, sizeof(int32_t) * (size_t)last, kind, ctx->stream), "colidx copy");
    guard(cudaMemcpyAsync(op->d_vals, vals, sizeof(T) * (size_t)last, kind, ctx->stream), "vals copy");
  }
  if (s == LLZ_OK) {
    if (op->idx32) {
      // narrow the row pointers to 32 bits (halves their traffic: A_bytes = nnz*(s+4) + 4(n+1))
      std::vector<int64_t> tmp;
      const int64_t* src = rowptr;
      if (!host_arrays) {
        tmp.resize((size_t)n_rows + 1);
        guard(cudaMemcpy(tmp.data(), rowptr, sizeof(int64_t) * (size_t)(n_rows + 1), cudaMemcpyDeviceToHost), "rowptr fetch");
        src = tmp.data();
      }
      std::vector<int32_t> r32((size_t)n_rows + 1);
      for (int64_t i = 0; i <= n_rows; ++i) r32[(size_t)i] = (int32_t)src[i];
      guard(cudaMalloc(&op->d_rowptr, sizeof(int32_t) * (size_t)(n_rows + 1)), "rowptr alloc");
      if (s == LLZ_OK)
        guard(cudaMemcpy(op->d_rowptr, r32.data(), sizeof(int32_t) * (size_t)(n_rows + 1), cudaMemcpyHostToDevice), "rowptr copy");
    } else {
      guard(cudaMalloc(&op->d_rowptr, sizeof(int64_t) * (size_t)(n_rows + 1)), "rowptr alloc");
      if (s == LLZ_OK) guard(cudaMemcpyAsync(op->d_rowptr, rowptr, sizeof(int64_t) * (size_t)(n_rows + 1), kind, ctx->stream), "rowptr copy");
    }
  }
  if (s == LLZ_OK) guard(cudaStreamSynchronize(ctx->stream), "sync");
  if (s != LLZ_OK) {
    delete op;
    return s;
  }
  // Stream kernel: largest power-of-two R <= 256 such that every block of R consecutive rows holds at most `cap`
  // non-zeros (cap bounded by ~40 KB of shared memory per CTA so several CTAs stay resident).
  {
    std::vector<int64_t> tmp;
    const int64_t* rp = rowptr;
    if (!host_arrays) {
      tmp.resize((size_t)n_rows + 1);
      if (cudaMemcpy(tmp.data(), rowptr, sizeof(int64_t) * (size_t)(n_rows + 1), cudaMemcpyDeviceToHost) != cudaSuccess) rp = nullptr;
      else rp = tmp.data();
    }
    const char* env = getenv("LLZ_SPMV");
    const bool want_stream = !(env && env[0] == 'v');
    const int cap_max = (int)((40 * 1024) / sizeof(T));
    if (rp && want_stream) {
      for (int R = 256; R >= 32; R /= 2) {
        int64_t worst = 0;
        for (int64_t r0 = 0; r0 < n_rows; r0 += R) {
          const int64_t r1 = std::min<int64_t>(n_rows, r0 + R);
          worst = std::max(worst, rp[r1] - rp[r0]);
        }
        if (worst <= cap_max) {
          op->stream_rows = R;
          op->stream_cap = (int)std::max<int64_t>(worst, 1);
          break;
        }
      }
    }
  }
  const double mean = n_rows > 0 ? (double)last / (double)n_rows : 1.0;
  int lpr = 1;
  while (lpr < 32 && lpr < mean) lpr *= 2;
  op->lpr = lpr;
  op->bytes = last * (int64_t)(sizeof(T) + 4) + (n_rows + 1) * (op->idx32 ? 4 : 8);
  llz_op_t h = new llz_op_s();
  h->impl = op;
  *out = h;
  return LLZ_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// User callback adapter
// ------------------------------------------------------------------------------------------------------------------
struct CallbackOp : OpBase {
  llz_apply_fn fn = nullptr;
  void* user = nullptr;
  bool overwrites = false;
  int apply_fused(const void* x, void* y, double sigma, double* pa, int* npa) override {
    (void)pa;
    *npa = 0;
    if (!overwrites) {
      cudaError_t e = cudaMemsetAsync(y, 0, (size_t)n_local * dtype_size(dtype), ctx->stream);  // reference contract: out pre-zeroed
      if (e != cudaSuccess) return fail(LLZ_ERR_CUDA, "memset: %s", cudaGetErrorString(e));
    }
    const int rc = fn(user, x, y, n_local, (void*)ctx->stream);
    if (rc != 0) return fail(LLZ_ERR_USER, "user operator callback returned %d", rc);
    if (sigma != 0.0) {
      const double a[2] = {sigma, 0.0};
      LLZ_TRY(launch_axpy(ctx, dtype, y, a, x, n_local));
    }
    return LLZ_OK;
  }
};

}  // namespace llz

using namespace llz;

extern "C" {

int llz_op_create_csr(llz_ctx_t ctx, int dtype, int64_t n_rows, int64_t n_cols, int64_t row0, const int64_t* rowptr,
                      const int32_t* colidx, const void* vals, int host_arrays, llz_op_t* op) {
  if (!ctx || !rowptr || !colidx || !vals || !op || n_rows < 1 || n_cols < 1)
    return fail(LLZ_ERR_INVALID, "op_create_csr: bad argument");
  if (ctx->nranks == 1 && (row0 != 0 || n_rows != n_cols))
    return fail(LLZ_ERR_INVALID, "op_create_csr: a single-rank operator must be square with row0 = 0");
  if (ctx->nranks > 1) return fail(LLZ_ERR_UNSUPPORTED, "row-sharded CSR is not built yet");
  LLZ_CUDA(cudaSetDevice(ctx->device));
  switch (dtype) {
    case LLZ_F32: return create_csr<float>(ctx, dtype, n_rows, n_cols, rowptr, colidx, vals, host_arrays, op);
    case LLZ_F64: return create_csr<double>(ctx, dtype, n_rows, n_cols, rowptr, colidx, vals, host_arrays, op);
    case LLZ_C64: return create_csr<float2>(ctx, dtype, n_rows, n_cols, rowptr, colidx, vals, host_arrays, op);
    case LLZ_C128: return create_csr<double2>(ctx, dtype, n_rows, n_cols, rowptr, colidx, vals, host_arrays, op);
  }
  return fail(LLZ_ERR_INVALID, "unknown dtype %d", dtype);
}

int llz_op_create_callback(llz_ctx_t ctx, int dtype, int64_t n_local, llz_apply_fn apply, void* user, int overwrites_y,
                           llz_op_t* op) {
  if (!ctx || !apply || !op || n_local < 1 || dtype_size(dtype) == 0) return fail(LLZ_ERR_INVALID, "op_create_callback: bad argument");
  auto* c = new CallbackOp();
  c->ctx = ctx;
  c->dtype = dtype;
  c->n_local = n_local;
  c->fn = apply;
  c->user = user;
  c->overwrites = overwrites_y != 0;
  c->bytes = 0;
  llz_op_t h = new llz_op_s();
  h->impl = c;
  *op = h;
  return LLZ_OK;
}

int llz_op_destroy(llz_op_t op) {
  if (!op) return LLZ_OK;
  if (op->impl) {
    cudaStreamSynchronize(op->impl->ctx->stream);
    delete op->impl;
  }
  delete op;
  return LLZ_OK;
}

int llz_op_rows(llz_op_t op, int64_t* n) {
  if (!op || !n) return fail(LLZ_ERR_INVALID, "null");
  *n = op->impl->n_local;
  return LLZ_OK;
}

int llz_op_bytes(llz_op_t op, int64_t* bytes) {
  if (!op || !bytes) return fail(LLZ_ERR_INVALID, "null");
  *bytes = op->impl->bytes;
  return LLZ_OK;
}

int llz_op_apply(llz_op_t op, llz_vec_t x, llz_vec_t y) {
  if (!op || !x || !y) return fail(LLZ_ERR_INVALID, "null");
  OpBase* o = op->impl;
  if (x->n != o->n_local || y->n != o->n_local || x->dtype != o->dtype || y->dtype != o->dtype)
    return fail(LLZ_ERR_INVALID, "op_apply: shape/dtype mismatch");
  int npa = 0;
  return o->apply_fused(x->d, y->d, 0.0, o->ctx->d_partials, &npa);
}

}  // extern "C"
