// llz_ops.cu — device operators: the replacement of the reference's `mv_mul` std::function
// (lambda_lanczos.hpp:120-126, exponentiator.hpp:35-41).  Built-in CSR (lanes-per-row and stream kernels) and
// SELL-32-sigma (register-staged and TMA-streamed kernels) SpMV with the alpha = Re<x, A x> dot fused into the
// epilogue, the halo exchange of their row-sharded form, the Gerschgorin row-sum kernels and the user callback
// adapter.  (The matrix-free XXZ operator lives in llz_xxz.cu.)
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

// Column indices of a row-sharded operator use the local "extended" numbering (llz_halo.cpp): [0, nloc) addresses the
// rank's own block of x, nloc + h the h-th entry of the halo buffer filled by the exchange before the launch.
// The halo is written by the PEERS' k_halo_push kernels while this kernel may already be resident (it waits for their
// announcement in its prologue), so it is not read-only for the kernel's lifetime: it must not go through the
// non-coherent path (ld.global.nc) — ld.global.cg reads it from L2, where the NVLink stores land.
template <class T> __device__ __forceinline__ T gather_x(const T* __restrict__ x, const T* halo, int32_t c, int32_t nloc) {
  return (c < nloc) ? __ldg(x + c) : __ldcg(halo + (c - nloc));
}

// *bad = 1 if any column index lies outside [0, n_cols)
__global__ void __launch_bounds__(kThreads) k_col_range(const int32_t* __restrict__ col, int64_t count, int32_t n_cols, int* bad) {
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < count; i += (int64_t)gridDim.x * kThreads)
    if (col[i] < 0 || col[i] >= n_cols) *bad = 1;
}

// sendbuf[i] = x[idx[i]]: the entries of the local block the peers asked for, grouped by peer.
template <class T> __global__ void __launch_bounds__(kThreads) k_pack(const T* __restrict__ x, const int32_t* __restrict__ idx, T* __restrict__ out, int64_t count) {
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < count; i += (int64_t)gridDim.x * kThreads) out[i] = __ldg(x + idx[i]);
}

// Peer-memory halo exchange: instead of pack -> ncclSend/ncclRecv, every rank stores the entries its peers need straight
// into their halo buffers (sub-allocations of the IPC-mapped window, double-buffered by message parity) and the last
// CTA announces the message; the SpMV kernels wait for the G announcements in their prologue.
template <class T>
__global__ void __launch_bounds__(kThreads) k_halo_push(const T* __restrict__ x, const int32_t* __restrict__ idx, long long n_send, HaloPush hp, PeerMsg msg) {
  __shared__ int last_cta;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n_send; i += (long long)gridDim.x * kThreads) {
    int q = 0;
    while (i >= hp.start[q + 1]) ++q;
    reinterpret_cast<T*>(hp.dst[q])[i - hp.start[q]] = __ldg(x + idx[i]);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();  // cumulative: covers the stores of the whole CTA ordered before it by the barrier
    const unsigned int t = atomicAdd(msg.ticket, 1u);
    last_cta = (t == gridDim.x - 1);
    if (last_cta) *msg.ticket = 0;
  }
  __syncthreads();
  if (last_cta && (int)threadIdx.x < msg.ch.G) {
    __threadfence_system();
    peer_announce(msg.ch, threadIdx.x, msg.seq);
  }
}

template <class T, class IDX, int LPR>
__global__ void __launch_bounds__(kThreads, 4)
    k_csr_spmv_dot(const IDX* __restrict__ rowptr, const int32_t* __restrict__ colidx, const T* __restrict__ vals,
                   const T* __restrict__ x, const T* halo, int32_t nloc, T* __restrict__ y, int64_t n,
                   typename Num<T>::R sigma, double* pa, PeerMsg msg, PeerMsg halo_msg) {
  __shared__ double scratch[kWarps];
  if (halo_msg.ch.G > 0) peer_wait(halo_msg.ch, halo_msg.seq);  // the peers' entries of x have landed in `halo`
  constexpr int ROWS = kThreads / LPR;
  const int tid = threadIdx.x;
  const int sub = tid % LPR;
  double dot = 0.0;
  for (int64_t row0 = (int64_t)blockIdx.x * ROWS; row0 < n; row0 += (int64_t)gridDim.x * ROWS) {
    const int64_t row = row0 + tid / LPR;
    T sum = zero_of(T());
    if (row < n) {
      const IDX p1 = rowptr[row + 1];
      for (IDX p = rowptr[row] + sub; p < p1; p += LPR) fmadd(sum, vals[p], gather_x(x, halo, colidx[p], nloc));
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
  finish_scalar(t, pa, msg, scratch);
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
                     const T* __restrict__ x, const T* halo, int32_t nloc, T* __restrict__ y, int64_t n,
                     typename Num<T>::R sigma, double* pa, int R, int cap, PeerMsg msg, PeerMsg halo_msg) {
  extern __shared__ __align__(16) unsigned char smem_s[];
  if (halo_msg.ch.G > 0) peer_wait(halo_msg.ch, halo_msg.seq);
  T* prod = reinterpret_cast<T*>(smem_s);
  IDX* rp = reinterpret_cast<IDX*>(smem_s + ((size_t)cap * sizeof(T) + 15) / 16 * 16);  // (8-byte row pointers after an odd count of floats)
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
    for (int e = tid; e < cnt; e += kThreads) prod[e] = mul_rn(vals[base + e], gather_x(x, halo, colidx[base + e], nloc));
    __syncthreads();
    for (int r = tid; r < nrows; r += kThreads) {
      const int j1 = (int)(rp[r + 1] - base);
      T s = zero_of(T());
      for (int j = (int)(rp[r] - base); j < j1; ++j) s = add_rn(s, prod[j]);
      const T xi = x[row0 + r];
      const T yi = add_t(s, scale_real(xi, sigma));
      y[row0 + r] = yi;
      dot += re_conj_mul(xi, yi);
    }
    __syncthreads();
  }
  const double t = block_sum(dot, scratch);
  finish_scalar(t, pa, msg, scratch);
}

// ------------------------------------------------------------------------------------------------------------------
// SELL-C-sigma (C = 32 = one warp per slice): rows are grouped in slices of 32, each slice stored column-major and
// padded to its longest row, so lane l of a warp walks row l of the slice with perfectly coalesced 128/256-byte
// accesses and no shared-memory staging or block barriers.  Rows may be sorted by length inside windows of sigma rows
// (perm) to cut the padding of irregular matrices.  Padding entries carry column -1 and are skipped (never 0 * x, so
// Inf/NaN in x cannot leak into rows that do not reference them).  Within a row the products are added in column
// order with separate multiply and add — the order and rounding of a sequential CSR loop (the reference's sample-style
// mv_mul), so y is bit-identical to the CSR kernels' and to the CPU's.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kSellC = 32;

template <class IDX> __global__ void __launch_bounds__(kThreads) k_sell_rowlen(const IDX* __restrict__ rowptr, int64_t n, int32_t* __restrict__ len) {
  for (int64_t r = (int64_t)blockIdx.x * kThreads + threadIdx.x; r < n; r += (int64_t)gridDim.x * kThreads)
    len[r] = (int32_t)(rowptr[r + 1] - rowptr[r]);
}

// One CTA (sigma threads) per window: stable sort of the window's rows by decreasing length, by ranking.
__global__ void k_sell_sort(const int32_t* __restrict__ len, int64_t n, int sigma, int32_t* __restrict__ perm) {
  extern __shared__ int32_t lens[];
  const int i = threadIdx.x;
  const int64_t row = (int64_t)blockIdx.x * sigma + i;
  const int32_t mine = row < n ? len[row] : -1;
  lens[i] = mine;
  __syncthreads();
  int rank = 0;
  for (int j = 0; j < sigma; ++j) {
    const int32_t o = lens[j];
    rank += (o > mine) || (o == mine && j < i);
  }
  if (row < n) perm[(int64_t)blockIdx.x * sigma + rank] = (int32_t)row;
}

// width[s] = longest row of slice s (one warp per slice)
__global__ void __launch_bounds__(kThreads) k_sell_width(const int32_t* __restrict__ len, const int32_t* __restrict__ perm, int64_t n,
                                                         int64_t n_slices, int32_t* __restrict__ width) {
  const int lane = threadIdx.x & 31;
  for (int64_t s = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5); s < n_slices; s += (int64_t)gridDim.x * kWarps) {
    const int64_t r = s * kSellC + lane;
    int32_t w = 0;
    if (r < n) w = len[perm ? perm[r] : r];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w = max(w, __shfl_xor_sync(0xffffffffu, w, o));
    if (lane == 0) width[s] = w;
  }
}

template <class T, class IDX>
__global__ void __launch_bounds__(kThreads) k_sell_fill(const IDX* __restrict__ rowptr, const int32_t* __restrict__ colidx, const T* __restrict__ vals,
                                                        const int32_t* __restrict__ perm, const int64_t* __restrict__ slice_ptr, int64_t n,
                                                        int64_t n_slices, int32_t* __restrict__ scol, T* __restrict__ sval) {
  const int lane = threadIdx.x & 31;
  for (int64_t s = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5); s < n_slices; s += (int64_t)gridDim.x * kWarps) {
    const int64_t p0 = slice_ptr[s];
    const int w = (int)((slice_ptr[s + 1] - p0) / kSellC);
    const int64_t r = s * kSellC + lane;
    IDX q0 = 0;
    int len = 0;
    if (r < n) {
      const int64_t src = perm ? perm[r] : r;
      q0 = rowptr[src];
      len = (int)(rowptr[src + 1] - q0);
    }
    for (int j = 0; j < w; ++j) {
      const int64_t dst = p0 + (int64_t)j * kSellC + lane;
      scol[dst] = j < len ? colidx[q0 + j] : -1;
      sval[dst] = j < len ? vals[q0 + j] : zero_of(T());
    }
  }
}

// Each warp walks TWO adjacent slices per step (64 rows: 2 x 5 x 384 B of matrix data in flight per warp on a 5-point
// stencil) and fetches the slice pointers of its next step before it starts on the current one, so the three dependent
// round trips (slice pointer -> column/value -> x gather) of consecutive steps overlap.
template <class T, int MINB, int Q = 4, int U = 2>
__global__ void __launch_bounds__(kThreads, MINB)
    k_sell_spmv_dot(const int64_t* __restrict__ slice_ptr, const int32_t* __restrict__ scol, const T* __restrict__ sval,
                    const int32_t* __restrict__ perm, const T* __restrict__ x, const T* halo, int32_t nloc,
                    T* __restrict__ y, int64_t n, int64_t n_slices, typename Num<T>::R sigma, double* pa, PeerMsg msg,
                    PeerMsg halo_msg) {
  __shared__ double scratch[kWarps];
  if (halo_msg.ch.G > 0) peer_wait(halo_msg.ch, halo_msg.seq);
  const int lane = threadIdx.x & 31;
  double dot = 0.0;
  const int64_t stride = (int64_t)gridDim.x * kWarps * U;
  int64_t s0 = ((int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5)) * U;
  // lanes 0..U hold slice_ptr[s0 + lane] (clamped), refreshed one step ahead
  auto fetch_ptr = [&](int64_t first) -> int64_t {
    const int64_t i = first + lane;
    return (lane <= U && first < n_slices) ? __ldg(slice_ptr + (i <= n_slices ? i : n_slices)) : 0;
  };
  int64_t sp = fetch_ptr(s0);
  for (; s0 < n_slices; s0 += stride) {
    int64_t p0[U];
    int w[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t a = __shfl_sync(0xffffffffu, sp, u), b = __shfl_sync(0xffffffffu, sp, u + 1);
      p0[u] = a;
      w[u] = (s0 + u < n_slices) ? (int)((b - a) / kSellC) : 0;
    }
    sp = fetch_ptr(s0 + stride);
    T sum[U];
#pragma unroll
    for (int u = 0; u < U; ++u) sum[u] = zero_of(T());
    const int wmax = U == 2 ? max(w[0], w[U - 1]) : w[0];
    for (int j = 0; j < wmax; j += Q) {  // Q entries per row in flight (8 when the whole matrix is one step of the grid)
      int32_t c[U][Q];
      T v[U][Q], xv[U][Q];
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          const bool in = j + q < w[u];
          const int64_t at = p0[u] + (int64_t)(j + q) * kSellC + lane;
          c[u][q] = in ? __ldg(scol + at) : -1;
          v[u][q] = in ? __ldg(sval + at) : zero_of(T());
        }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int q = 0; q < Q; ++q) xv[u][q] = c[u][q] >= 0 ? gather_x(x, halo, c[u][q], nloc) : zero_of(T());
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int q = 0; q < Q; ++q)
          if (c[u][q] >= 0) sum[u] = add_rn(sum[u], mul_rn(v[u][q], xv[u][q]));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = (s0 + u) * kSellC + lane;
      if (r < n) {
        const int64_t out = perm ? perm[r] : r;
        const T xi = x[out];
        const T yi = add_t(sum[u], scale_real(xi, sigma));
        y[out] = yi;
        dot += re_conj_mul(xi, yi);
      }
    }
  }
  const double t = block_sum(dot, scratch);
  finish_scalar(t, pa, msg, scratch);
}

// ------------------------------------------------------------------------------------------------------------------
// SELL SpMV with the matrix streamed by TMA: the column / value arrays of consecutive slices are contiguous, so a CTA
// fetches a whole chunk of 16 slices (512 rows) with two `cp.async.bulk` copies into shared memory, completion
// signalled on an mbarrier, two to three chunks ahead of the warps that consume them.  The matrix stream (94 % of the
// DRAM bytes of a stencil SpMV) is thereby decoupled from the dependent x gathers: no register staging, no load ->
// gather round-trip chain per slice, ~100 KB of matrix data in flight per SM.  Same per-row arithmetic as
// k_sell_spmv_dot (bit-identical y).
// ------------------------------------------------------------------------------------------------------------------
constexpr int kTmaChunk = 16;  // slices per chunk: 8 warps x 2 slices

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(b))
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(b)), "r"(parity)
      : "memory");
  return ok != 0;
}

template <class T, int STAGES>
__global__ void __launch_bounds__(kThreads, 2)
    k_sell_tma_spmv_dot(const int64_t* __restrict__ slice_ptr, const int32_t* __restrict__ scol, const T* __restrict__ sval,
                        const int32_t* __restrict__ perm, const T* __restrict__ x, const T* halo, int32_t nloc,
                        T* __restrict__ y, int64_t n, int64_t n_slices, typename Num<T>::R sigma, double* pa, PeerMsg msg,
                        PeerMsg halo_msg, int cap /* entries per stage */) {
  extern __shared__ __align__(128) unsigned char smem_t[];
  __shared__ double scratch[kWarps];
  __shared__ __align__(8) uint64_t full[STAGES];
  __shared__ long long sp[STAGES][kTmaChunk + 1];
  if (halo_msg.ch.G > 0) peer_wait(halo_msg.ch, halo_msg.seq);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t stage_bytes = (size_t)cap * (sizeof(T) + sizeof(int32_t));
  auto stage_vals = [&](int s) { return reinterpret_cast<T*>(smem_t + (size_t)s * stage_bytes); };
  auto stage_cols = [&](int s) { return reinterpret_cast<int32_t*>(smem_t + (size_t)s * stage_bytes + (size_t)cap * sizeof(T)); };
  const int64_t nchunks = (n_slices + kTmaChunk - 1) / kTmaChunk;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // warp 0 fetches chunk number i of this CTA into stage i % STAGES (slice pointers first, then the two bulk copies)
  auto issue = [&](int64_t i) {
    const int64_t c = (int64_t)blockIdx.x + i * (int64_t)gridDim.x;
    if (c >= nchunks) return;
    const int s = (int)(i % STAGES);
    const int64_t s0 = c * kTmaChunk;
    if (lane <= kTmaChunk) {
      const int64_t at = s0 + lane;
      sp[s][lane] = __ldg(slice_ptr + (at <= n_slices ? at : n_slices));
    }
    __syncwarp();
    if (lane == 0) {
      const long long p0 = sp[s][0];
      const uint32_t cnt = (uint32_t)(sp[s][kTmaChunk] - p0);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic reads of this stage are done (barrier)
      mbar_arrive_expect_tx(&full[s], cnt * (uint32_t)(sizeof(T) + sizeof(int32_t)));
      if (cnt > 0) {
        bulk_copy_g2s(stage_vals(s), sval + p0, cnt * (uint32_t)sizeof(T), &full[s]);
        bulk_copy_g2s(stage_cols(s), scol + p0, cnt * (uint32_t)sizeof(int32_t), &full[s]);
      }
    }
  };
  if (warp == 0)
    for (int i = 0; i < STAGES - 1; ++i) issue(i);
  double dot = 0.0;
  for (int64_t i = 0;; ++i) {
    const int64_t c = (int64_t)blockIdx.x + i * (int64_t)gridDim.x;
    if (c >= nchunks) break;
    const int s = (int)(i % STAGES);
    const uint32_t parity = (uint32_t)((i / STAGES) & 1);
    if (warp == 0) issue(i + STAGES - 1);  // refills the stage every warp finished with at the end of the last turn
    while (!mbar_try_wait(&full[s], parity)) {
    }
    const T* __restrict__ vs = stage_vals(s);
    const int32_t* __restrict__ cs = stage_cols(s);
    const long long base = sp[s][0];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int q = warp * 2 + u;
      const int64_t slice = c * kTmaChunk + q;
      if (slice >= n_slices) continue;
      const int off = (int)(sp[s][q] - base) + lane;
      const int w = (int)((sp[s][q + 1] - sp[s][q]) / kSellC);
      T sum = zero_of(T());
      int j = 0;
      for (; j + 4 <= w; j += 4) {
        int32_t cc[4];
        T vv[4], xv[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          cc[t] = cs[off + (j + t) * kSellC];
          vv[t] = vs[off + (j + t) * kSellC];
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) xv[t] = cc[t] >= 0 ? gather_x(x, halo, cc[t], nloc) : zero_of(T());
#pragma unroll
        for (int t = 0; t < 4; ++t)
          if (cc[t] >= 0) sum = add_rn(sum, mul_rn(vv[t], xv[t]));
      }
      for (; j < w; ++j) {
        const int32_t cc = cs[off + j * kSellC];
        const T vv = vs[off + j * kSellC];
        if (cc >= 0) sum = add_rn(sum, mul_rn(vv, gather_x(x, halo, cc, nloc)));
      }
      const int64_t r = slice * kSellC + lane;
      if (r < n) {
        const int64_t out = perm ? perm[r] : r;
        const T xi = x[out];
        const T yi = add_t(sum, scale_real(xi, sigma));
        y[out] = yi;
        dot += re_conj_mul(xi, yi);
      }
    }
    __syncthreads();  // every warp is done with stage s (and its slice pointers) before warp 0 refills it
  }
  const double t = block_sum(dot, scratch);
  finish_scalar(t, pa, msg, scratch);
}

// ------------------------------------------------------------------------------------------------------------------
// DIA storage for stencil / lattice operators (2-D Laplacian, tight-binding models ...): when every non-zero sits on one
// of a few diagonals (column - row constant), the matrix is re-stored diagonal by diagonal — vals[d][row], no column
// indices — so a row reads its products' operands at x[row + offset_d]: contiguous, perfectly coalesced streams instead
// of dependent gathers, and A_bytes drops from nnz*(s+4) to ndiag*n*s + 2n (a 16-bit presence mask per row: absent
// entries — grid boundaries — are skipped, never multiplied as zeros).  Products are added in ascending column order
// with separate multiply and add: y is bit-identical to the CSR / SELL kernels'.  Row-sharded, offsets that leave the
// block read the halo buffer, which for such operators is the contiguous range of rows just below and just above.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kDiaMax = 16;
struct DiaOffsets {
  int nd = 0;
  int off[kDiaMax] = {};  // ascending
};

// One thread per row: scatter the CSR entries of the row into their diagonals.  *bad is raised if an entry is not on a
// listed diagonal or the row is not sorted by column (the DIA kernel adds in ascending-column order).
template <class T, class IDX>
__global__ void __launch_bounds__(kThreads) k_dia_fill(const IDX* __restrict__ rowptr, const int32_t* __restrict__ colidx, const T* __restrict__ vals,
                                                       int64_t n, int32_t n_lo, DiaOffsets offs, int64_t ldv, T* __restrict__ dvals,
                                                       uint16_t* __restrict__ mask, int* bad) {
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    uint32_t m = 0;
    int last = -1;
    for (IDX p = rowptr[i]; p < rowptr[i + 1]; ++p) {
      const int64_t c = colidx[p];
      // extended index: own rows [0, n), the halo rows below the block at [-n_lo, 0), those above at [n, ...)
      const int64_t e = c < n ? c : (c - n < n_lo ? c - n - n_lo : n + (c - n - n_lo));
      const int64_t off = e - i;
      int d = -1;
      for (int q = 0; q < offs.nd; ++q)
        if (offs.off[q] == off) d = q;
      if (d < 0 || d <= last) {
        *bad = 1;
        continue;
      }
      last = d;
      dvals[(int64_t)d * ldv + i] = vals[p];
      m |= 1u << d;
    }
    mask[i] = (uint16_t)m;
  }
}

template <class T, int ND, bool SHARDED>
__global__ void __launch_bounds__(kThreads, ND <= 3 ? 4 : (ND <= 6 ? 3 : 2))
    k_dia_spmv_dot(const T* __restrict__ dvals, int64_t ldv, const uint16_t* __restrict__ mask, DiaOffsets offs, const T* __restrict__ x,
                   const T* halo, int32_t n_lo, T* __restrict__ y, int64_t n, typename Num<T>::R sigma, double* pa, PeerMsg msg,
                   PeerMsg halo_msg) {
  __shared__ double scratch[kWarps];
  if (SHARDED && halo_msg.ch.G > 0) peer_wait(halo_msg.ch, halo_msg.seq);
  constexpr int U = (sizeof(T) <= 8 && ND <= 8) ? 2 : 1;  // rows per thread and step (loads in flight vs registers)
  double dot = 0.0;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t i0 = (int64_t)blockIdx.x * kThreads + threadIdx.x; i0 < n; i0 += stride * U) {
    uint32_t m[U];
    T a[U][ND], xv[U][ND], xi[U];
    // every load of the step is issued before the first is used; the presence mask only gates the arithmetic (absent
    // entries are stored as zeros, and the x index of an absent entry is clamped into the vector), so nothing waits
    // for the mask
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      const bool row = i < n;
      const int64_t ic = row ? i : n - 1;
      m[u] = row ? mask[ic] : 0u;
      xi[u] = x[ic];
#pragma unroll
      for (int d = 0; d < ND; ++d) {
        if (d < offs.nd) {
          a[u][d] = __ldg(dvals + (int64_t)d * ldv + ic);
          const int64_t e = ic + offs.off[d];
          if (e >= 0 && e < n) {
            xv[u][d] = __ldg(x + e);
          } else if (SHARDED) {
            const int64_t h = e < 0 ? n_lo + e : n_lo + (e - n);
            xv[u][d] = ((m[u] >> d) & 1u) ? __ldcg(halo + h) : zero_of(T());  // written by the peers: coherent load
          } else {
            xv[u][d] = zero_of(T());
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < n) {
        T sum = zero_of(T());
#pragma unroll
        for (int d = 0; d < ND; ++d)
          if (d < offs.nd && ((m[u] >> d) & 1u)) sum = add_rn(sum, mul_rn(a[u][d], xv[u][d]));
        const T yi = add_t(sum, scale_real(xi[u], sigma));
        y[i] = yi;
        dot += re_conj_mul(xi[u], yi);
      }
    }
  }
  const double t = block_sum(dot, scratch);
  finish_scalar(t, pa, msg, scratch);
}

// Gerschgorin radius: per-CTA maximum of the absolute row sums (CSR: one thread per row; SELL: one lane per row).
template <class T, class IDX>
__global__ void __launch_bounds__(kThreads) k_csr_rowsum_max(const IDX* __restrict__ rowptr, const T* __restrict__ vals, int64_t n, double* out) {
  __shared__ double red[kThreads];
  double m = 0.0;
  for (int64_t r = (int64_t)blockIdx.x * kThreads + threadIdx.x; r < n; r += (int64_t)gridDim.x * kThreads) {
    double s = 0.0;
    for (IDX p = rowptr[r]; p < rowptr[r + 1]; ++p) s += abs1(vals[p]);
    m = fmax(m, s);
  }
  red[threadIdx.x] = m;
  __syncthreads();
  for (int o = kThreads / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + o]);
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = red[0];
}
template <class T>
__global__ void __launch_bounds__(kThreads) k_sell_rowsum_max(const int64_t* __restrict__ slice_ptr, const T* __restrict__ sval, int64_t n_slices, double* out) {
  __shared__ double red[kThreads];
  const int lane = threadIdx.x & 31;
  double m = 0.0;
  for (int64_t s = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5); s < n_slices; s += (int64_t)gridDim.x * kWarps) {
    const int64_t p0 = slice_ptr[s];
    const int w = (int)((slice_ptr[s + 1] - p0) / kSellC);
    double t = 0.0;
    for (int j = 0; j < w; ++j) t += abs1(sval[p0 + (int64_t)j * kSellC + lane]);  // padding holds zeros
    m = fmax(m, t);
  }
  red[threadIdx.x] = m;
  __syncthreads();
  for (int o = kThreads / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + o]);
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = red[0];
}

template <class T>
__global__ void __launch_bounds__(kThreads) k_dia_rowsum_max(const T* __restrict__ dvals, int64_t ldv, int nd, int64_t n, double* out) {
  __shared__ double red[kThreads];
  double m = 0.0;
  for (int64_t r = (int64_t)blockIdx.x * kThreads + threadIdx.x; r < n; r += (int64_t)gridDim.x * kThreads) {
    double t = 0.0;
    for (int d = 0; d < nd; ++d) t += abs1(dvals[(int64_t)d * ldv + r]);  // absent entries hold zeros
    m = fmax(m, t);
  }
  red[threadIdx.x] = m;
  __syncthreads();
  for (int o = kThreads / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + o]);
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = red[0];
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
  // row-sharded runs: halo exchange plan (empty for a single rank)
  int64_t n_halo = 0, n_send = 0;
  T* d_halo = nullptr;           // [n_halo] remote entries of x, grouped by owner
  T* d_sendbuf = nullptr;        // [n_send] packed entries of the local block, grouped by requesting peer
  int32_t* d_send_idx = nullptr; // [n_send] local indices to pack
  std::vector<size_t> send_off, send_bytes, recv_off, recv_bytes;
  // peer-memory variant: halo buffers live in this rank's IPC window (two copies, by message parity)
  int64_t win_off = -1;
  HaloPush push;                    // destinations for parity 0; parity 1 is + push_stride[q] bytes
  size_t push_stride[kMaxRanks] = {};
  const T* cur_halo = nullptr;      // what the next apply reads
  PeerMsg cur_halo_msg;             // ... after waiting for this message
  // SELL-C-sigma storage (convert_to_sell releases the CSR arrays)
  bool sell = false;
  int sell_sigma = 1;
  int64_t n_slices = 0, padded_nnz = 0;
  int64_t* d_slice_ptr = nullptr;
  int32_t* d_scol = nullptr;
  T* d_sval = nullptr;
  int32_t* d_perm = nullptr;  // null when rows are not sorted (sigma = 1)
  int tma_cap = 0;            // > 0: entries per shared-memory stage of the TMA-streamed kernel; 0: register-staged kernel
  int tma_stages = 0;
  // DIA storage (convert_to_dia releases the CSR arrays)
  bool dia = false;
  DiaOffsets dia_offs;
  int64_t dia_ld = 0;
  T* d_dvals = nullptr;
  uint16_t* d_dmask = nullptr;
  int32_t halo_lo = 0;        // row-sharded: halo entries that lie below the block ...
  bool halo_contiguous = true;  // ... and whether the halo is exactly the rows just below and just above the block

  ~CsrOp() override {
    if (d_slice_ptr) dev_free(ctx, d_slice_ptr);
    if (d_scol) dev_free(ctx, d_scol);
    if (d_sval) dev_free(ctx, d_sval);
    if (d_perm) dev_free(ctx, d_perm);
    if (d_dvals) dev_free(ctx, d_dvals);
    if (d_dmask) dev_free(ctx, d_dmask);
    if (d_rowptr) dev_free(ctx, d_rowptr);
    if (d_colidx) dev_free(ctx, d_colidx);
    if (d_vals) dev_free(ctx, d_vals);
    if (d_halo) dev_free(ctx, d_halo);
    if (win_off >= 0) comm_window_free(ctx, win_off, 2 * (size_t)std::max<int64_t>(n_halo, 1) * sizeof(T));
    if (d_sendbuf) dev_free(ctx, d_sendbuf);
    if (d_send_idx) dev_free(ctx, d_send_idx);
  }
  int32_t nloc32() const { return (int32_t)std::min<int64_t>(n_local, 0x7fffffff); }
  const char* storage() const override {
    if (dia) return "DIA";
    if (sell) return d_perm ? "SELL-32-sigma (sorted windows)" : "SELL-32";
    return stream_rows > 0 ? "CSR (stream kernel)" : "CSR (lanes-per-row kernel)";
  }

  // Bring the remote entries of x this block references into d_halo (no-op for a single rank).
  int prepare(const void* x) override {
    cur_halo = d_halo;
    cur_halo_msg = PeerMsg();
    if (ctx->nranks == 1) return LLZ_OK;
    if (win_off >= 0) {  // peer-memory exchange (every rank of the group takes this branch together)
      ProfScope ps(ctx, "halo", (double)(n_halo + n_send) * sizeof(T));
      PeerMsg m = comm_next_message(ctx, kChanHalo);
      HaloPush hp = push;
      if (m.seq & 1ull)
        for (int q = 0; q < hp.G; ++q) hp.dst[q] = static_cast<char*>(hp.dst[q]) + push_stride[q];
      const int64_t g = std::max<int64_t>(1, std::min<int64_t>((n_send + kThreads - 1) / kThreads, (int64_t)ctx->num_sms * 2));
      cudaError_t e = launch_chain(ctx, k_halo_push<T>, (int)g, kThreads, 0, (const T*)x, d_send_idx, (long long)n_send, hp, m);
      if (e != cudaSuccess) return fail(LLZ_ERR_CUDA, "launch k_halo_push: %s", cudaGetErrorString(e));
      ctx->launches++;
      cur_halo = reinterpret_cast<const T*>(static_cast<char*>(comm_window_ptr(ctx, ctx->rank, win_off)) +
                                            ((m.seq & 1ull) ? (size_t)std::max<int64_t>(n_halo, 1) * sizeof(T) : 0));
      cur_halo_msg = m;
      return LLZ_OK;
    }
    if (n_halo == 0 && n_send == 0) return LLZ_OK;
    ProfScope ps(ctx, "halo", (double)(n_halo + n_send) * sizeof(T));
    if (n_send > 0) {
      const int64_t g = std::min<int64_t>((n_send + kThreads - 1) / kThreads, (int64_t)ctx->num_sms * 4);
      k_pack<T><<<(int)g, kThreads, 0, ctx->stream>>>((const T*)x, d_send_idx, d_sendbuf, n_send);
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) return fail(LLZ_ERR_CUDA, "launch k_pack: %s", cudaGetErrorString(e));
      ctx->launches++;
    }
    return comm_exchange(ctx, (const char*)d_sendbuf, send_off.data(), send_bytes.data(), (char*)d_halo, recv_off.data(),
                         recv_bytes.data());
  }

  bool plan_halo_push(HaloPushPlan* plan) override {
    if (ctx->nranks == 1 || win_off < 0) return false;
    plan->msg = comm_next_message(ctx, kChanHalo);
    plan->hp = push;
    if (plan->msg.seq & 1ull)
      for (int q = 0; q < plan->hp.G; ++q) plan->hp.dst[q] = static_cast<char*>(plan->hp.dst[q]) + push_stride[q];
    plan->idx = d_send_idx;
    plan->n_send = n_send;
    return true;
  }
  void use_pushed_halo(const HaloPushPlan& plan) override {
    cur_halo = reinterpret_cast<const T*>(static_cast<char*>(comm_window_ptr(ctx, ctx->rank, win_off)) +
                                          ((plan.msg.seq & 1ull) ? (size_t)std::max<int64_t>(n_halo, 1) * sizeof(T) : 0));
    cur_halo_msg = plan.msg;
  }

  template <class IDX, int LPR> int launch(const void* x, void* y, double sigma, double* pa, int* npa, const PeerMsg& msg) {
    constexpr int ROWS = kThreads / LPR;
    int64_t blocks = (n_local + ROWS - 1) / ROWS;
    int64_t g = std::min<int64_t>(blocks, std::min<int64_t>(kMaxGrid, (int64_t)ctx->num_sms * 8));
    if (g < 1) g = 1;
    cudaError_t e = launch_chain(ctx, k_csr_spmv_dot<T, IDX, LPR>, (int)g, kThreads, 0, (const IDX*)d_rowptr, d_colidx, d_vals, (const T*)x,
                                 cur_halo, nloc32(), (T*)y, n_local, (typename Num<T>::R)sigma, pa, msg, cur_halo_msg);
    *npa = (int)g;
    if (e != cudaSuccess) return fail(LLZ_ERR_CUDA, "launch k_csr_spmv_dot: %s", cudaGetErrorString(e));
    ctx->launches++;
    return LLZ_OK;
  }

  template <class IDX> int launch_lpr(const void* x, void* y, double sigma, double* pa, int* npa, const PeerMsg& msg) {
    switch (lpr) {
      case 1: return launch<IDX, 1>(x, y, sigma, pa, npa, msg);
      case 2: return launch<IDX, 2>(x, y, sigma, pa, npa, msg);
      case 4: return launch<IDX, 4>(x, y, sigma, pa, npa, msg);
      case 8: return launch<IDX, 8>(x, y, sigma, pa, npa, msg);
      case 16: return launch<IDX, 16>(x, y, sigma, pa, npa, msg);
      default: return launch<IDX, 32>(x, y, sigma, pa, npa, msg);
    }
  }

  template <class IDX> int launch_stream(const void* x, void* y, double sigma, double* pa, int* npa, const PeerMsg& msg) {
    const size_t smem = ((size_t)stream_cap * sizeof(T) + 15) / 16 * 16 + (size_t)(stream_rows + 1) * sizeof(IDX);
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(k_csr_stream_dot<T, IDX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return fail(LLZ_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    const int64_t blocks = (n_local + stream_rows - 1) / stream_rows;
    int64_t g = std::min<int64_t>(blocks, std::min<int64_t>(kMaxGrid, (int64_t)ctx->num_sms * 8));
    if (g < 1) g = 1;
    cudaError_t e = launch_chain(ctx, k_csr_stream_dot<T, IDX>, (int)g, kThreads, smem, (const IDX*)d_rowptr, d_colidx, d_vals, (const T*)x,
                                 cur_halo, nloc32(), (T*)y, n_local, (typename Num<T>::R)sigma, pa, stream_rows, stream_cap, msg,
                                 cur_halo_msg);
    *npa = (int)g;
    if (e != cudaSuccess) return fail(LLZ_ERR_CUDA, "launch k_csr_stream_dot: %s", cudaGetErrorString(e));
    ctx->launches++;
    return LLZ_OK;
  }

  template <int STAGES> int launch_sell_tma(const void* x, void* y, double sigma, double* pa, int* npa, const PeerMsg& msg) {
    const size_t smem = (size_t)STAGES * tma_cap * (sizeof(T) + sizeof(int32_t));
    cudaError_t e = cudaFuncSetAttribute(k_sell_tma_spmv_dot<T, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(LLZ_ERR_CUDA, "cudaFuncSetAttribute(k_sell_tma_spmv_dot, %zu): %s", smem, cudaGetErrorString(e));
    const int64_t nchunks = (n_slices + kTmaChunk - 1) / kTmaChunk;
    const int per_sm = smem * 2 <= 200 * 1024 ? 2 : 1;
    int64_t g = std::max<int64_t>(1, std::min<int64_t>(nchunks, (int64_t)ctx->num_sms * per_sm));
    e = launch_chain(ctx, k_sell_tma_spmv_dot<T, STAGES>, (int)g, kThreads, smem, d_slice_ptr, d_scol, d_sval, d_perm, (const T*)x, cur_halo,
                     nloc32(), (T*)y, n_local, n_slices, (typename Num<T>::R)sigma, pa, msg, cur_halo_msg, tma_cap);
    *npa = (int)g;
    if (e != cudaSuccess) return fail(LLZ_ERR_CUDA, "launch k_sell_tma_spmv_dot: %s", cudaGetErrorString(e));
    ctx->launches++;
    return LLZ_OK;
  }

  int launch_sell(const void* x, void* y, double sigma, double* pa, int* npa, const PeerMsg& msg) {
    if (tma_cap > 0) return tma_stages >= 3 ? launch_sell_tma<3>(x, y, sigma, pa, npa, msg) : launch_sell_tma<2>(x, y, sigma, pa, npa, msg);
    const int64_t per_cta = kWarps * 2;  // slices one CTA covers per step
    int64_t g = std::min<int64_t>((n_slices + per_cta - 1) / per_cta, std::min<int64_t>(kMaxGrid, (int64_t)ctx->num_sms * 8));
    if (g < 1) g = 1;
    // real element types fit 64 registers (4 resident CTAs, more loads in flight); complex ones need 80 (3 CTAs)
    static const int env_blocks = getenv("LLZ_SELL_BLOCKS") ? atoi(getenv("LLZ_SELL_BLOCKS")) : 0;
    const int blocks = env_blocks ? env_blocks : (Num<T>::NC == 1 ? 4 : 3);
    cudaError_t e;
    if (g * per_cta >= n_slices && (int64_t)g <= (int64_t)ctx->num_sms * 2) {
      // small matrix: every warp has at most one step and the grid does not fill the GPU — latency-bound, so one slice
      // per warp (twice the warps) with twice the entries per row in flight (the registers that costs do not limit
      // occupancy here)
      g = std::max<int64_t>(1, (n_slices + kWarps - 1) / kWarps);
      e = launch_chain(ctx, k_sell_spmv_dot<T, 2, 8, 1>, (int)g, kThreads, 0, d_slice_ptr, d_scol, d_sval, d_perm, (const T*)x, cur_halo, nloc32(),
                       (T*)y, n_local, n_slices, (typename Num<T>::R)sigma, pa, msg, cur_halo_msg);
    } else if (blocks >= 4)
      e = launch_chain(ctx, k_sell_spmv_dot<T, 4>, (int)g, kThreads, 0, d_slice_ptr, d_scol, d_sval, d_perm, (const T*)x, cur_halo, nloc32(),
                       (T*)y, n_local, n_slices, (typename Num<T>::R)sigma, pa, msg, cur_halo_msg);
    else
      e = launch_chain(ctx, k_sell_spmv_dot<T, 3>, (int)g, kThreads, 0, d_slice_ptr, d_scol, d_sval, d_perm, (const T*)x, cur_halo, nloc32(),
                       (T*)y, n_local, n_slices, (typename Num<T>::R)sigma, pa, msg, cur_halo_msg);
    *npa = (int)g;
    if (e != cudaSuccess) return fail(LLZ_ERR_CUDA, "launch k_sell_spmv_dot: %s", cudaGetErrorString(e));
    ctx->launches++;
    return LLZ_OK;
  }

  template <int ND> int launch_dia_nd(const void* x, void* y, double sigma, double* pa, int* npa, const PeerMsg& msg) {
    int64_t g = std::min<int64_t>((n_local + kThreads - 1) / kThreads, std::min<int64_t>(kMaxGrid, (int64_t)ctx->num_sms * 8));
    if (g < 1) g = 1;
    cudaError_t e;
    if (ctx->nranks > 1)
      e = launch_chain(ctx, k_dia_spmv_dot<T, ND, true>, (int)g, kThreads, 0, (const T*)d_dvals, dia_ld, (const uint16_t*)d_dmask, dia_offs,
                       (const T*)x, cur_halo, halo_lo, (T*)y, n_local, (typename Num<T>::R)sigma, pa, msg, cur_halo_msg);
    else
      e = launch_chain(ctx, k_dia_spmv_dot<T, ND, false>, (int)g, kThreads, 0, (const T*)d_dvals, dia_ld, (const uint16_t*)d_dmask, dia_offs,
                       (const T*)x, (const T*)nullptr, 0, (T*)y, n_local, (typename Num<T>::R)sigma, pa, msg, PeerMsg());
    *npa = (int)g;
    if (e != cudaSuccess) return fail(LLZ_ERR_CUDA, "launch k_dia_spmv_dot: %s", cudaGetErrorString(e));
    ctx->launches++;
    return LLZ_OK;
  }
  int launch_dia(const void* x, void* y, double sigma, double* pa, int* npa, const PeerMsg& msg) {
    if (dia_offs.nd <= 3) return launch_dia_nd<3>(x, y, sigma, pa, npa, msg);   // 1-D chains
    if (dia_offs.nd <= 6) return launch_dia_nd<6>(x, y, sigma, pa, npa, msg);   // 2-D 5-point stencils, square-lattice hopping
    if (dia_offs.nd <= 8) return launch_dia_nd<8>(x, y, sigma, pa, npa, msg);   // 3-D 7-point stencils
    return launch_dia_nd<16>(x, y, sigma, pa, npa, msg);
  }

  // Re-store the uploaded CSR arrays diagonal by diagonal when (almost) every non-zero sits on one of at most 16
  // diagonals.  The candidate offsets come from a sample of rows (host); the device fill kernel verifies that EVERY
  // entry is on one of them and that rows are sorted by column, else the CSR arrays are kept (*done = false).
  // `rowptr` / `col_local`: host copies (columns in the local extended numbering when row-sharded).
  template <class IDX> int convert_to_dia(const int64_t* rowptr, const int32_t* col_local, bool* done) {
    *done = false;
    const int64_t n = n_local;
    if (n < 1 || nnz < 1 || (ctx->nranks > 1 && !halo_contiguous) || n + (int64_t)halo_lo + (n_halo - halo_lo) >= 0x7fffffff) return LLZ_OK;
    const int64_t n_hi = n_halo - halo_lo;
    std::vector<int64_t> offsets;
    auto scan_row = [&](int64_t i) -> bool {
      for (int64_t p = rowptr[i]; p < rowptr[i + 1]; ++p) {
        const int64_t c = col_local[p];
        const int64_t e = c < n ? c : (c - n < halo_lo ? c - n - halo_lo : n + (c - n - halo_lo));
        const int64_t off = e - i;
        if (std::find(offsets.begin(), offsets.end(), off) == offsets.end()) {
          if ((int)offsets.size() == kDiaMax) return false;
          offsets.push_back(off);
        }
      }
      return true;
    };
    const int64_t stride = std::max<int64_t>(1, n / 4096);
    for (int64_t i = 0; i < n; i += stride)
      if (!scan_row(i)) return LLZ_OK;
    for (int64_t i = 0; i < std::min<int64_t>(n, 64); ++i)
      if (!scan_row(i) || !scan_row(n - 1 - i)) return LLZ_OK;
    if (offsets.empty()) return LLZ_OK;
    const int nd = (int)offsets.size();
    if ((double)nd * (double)n > 1.3 * (double)nnz) return LLZ_OK;  // too many absent entries: SELL stores less
    std::sort(offsets.begin(), offsets.end());
    if (offsets.front() < -(n + (int64_t)halo_lo) || offsets.back() > n + n_hi) return LLZ_OK;
    DiaOffsets offs;
    offs.nd = nd;
    for (int d = 0; d < nd; ++d) offs.off[d] = (int)offsets[(size_t)d];
    const int64_t ldv = (n + 31) / 32 * 32;
    T* dv = nullptr;
    uint16_t* dm = nullptr;
    cudaError_t e = dev_malloc(ctx, &dv, sizeof(T) * (size_t)ldv * nd);
    if (e == cudaSuccess) e = dev_malloc(ctx, &dm, sizeof(uint16_t) * (size_t)n);
    if (e != cudaSuccess) {  // not enough room for both copies: keep what we have
      if (dv) dev_free(ctx, dv);
      (void)cudaGetLastError();
      return LLZ_OK;
    }
    int* d_bad = reinterpret_cast<int*>(ctx->d_result);
    e = cudaMemsetAsync(dv, 0, sizeof(T) * (size_t)ldv * nd, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream);
    if (e == cudaSuccess) {
      const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + kThreads - 1) / kThreads, (int64_t)ctx->num_sms * 8));
      k_dia_fill<T, IDX><<<grid, kThreads, 0, ctx->stream>>>((const IDX*)d_rowptr, d_colidx, d_vals, n, halo_lo, offs, ldv, dv, dm, d_bad);
      e = cudaGetLastError();
      ctx->launches++;
    }
    int bad = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess || bad) {
      dev_free(ctx, dv);
      dev_free(ctx, dm);
      if (e != cudaSuccess) return fail(LLZ_ERR_CUDA, "DIA conversion: %s", cudaGetErrorString(e));
      return LLZ_OK;
    }
    d_dvals = dv;
    d_dmask = dm;
    dia_offs = offs;
    dia_ld = ldv;
    dev_free(ctx, d_rowptr);
    dev_free(ctx, d_colidx);
    dev_free(ctx, d_vals);
    d_rowptr = nullptr;
    d_colidx = nullptr;
    d_vals = nullptr;
    dia = true;
    bytes = (int64_t)nd * n * (int64_t)sizeof(T) + 2 * n;
    *done = true;
    return LLZ_OK;
  }

  // Re-store the uploaded CSR arrays as SELL-32-sigma, entirely on the device (only the slice widths visit the host
  // for the prefix sum).  sigma: 1 = keep the row order, 32..1024 (multiple of 32) = sort windows of sigma rows by
  // length, 0 = pick (sort only if it saves more than 5 % of the padded storage).
  template <class IDX> int convert_to_sell(int sigma) {
    const int64_t n = n_local;
    n_slices = (n + kSellC - 1) / kSellC;
    int32_t* d_len = nullptr;
    int32_t* d_width = nullptr;
    LLZ_CUDA(dev_malloc(ctx, &d_len, sizeof(int32_t) * (size_t)n));
    LLZ_CUDA(dev_malloc(ctx, &d_width, sizeof(int32_t) * (size_t)n_slices));
    const int grid = (int)std::min<int64_t>((n + kThreads - 1) / kThreads, (int64_t)ctx->num_sms * 8);
    const int wgrid = (int)std::min<int64_t>((n_slices + kWarps - 1) / kWarps, (int64_t)ctx->num_sms * 8);
    k_sell_rowlen<IDX><<<grid, kThreads, 0, ctx->stream>>>((const IDX*)d_rowptr, n, d_len);
    std::vector<int32_t> width((size_t)n_slices);
    auto widths_for = [&](const int32_t* perm, int64_t* total) -> int {
      k_sell_width<<<wgrid, kThreads, 0, ctx->stream>>>(d_len, perm, n, n_slices, d_width);
      LLZ_CUDA(cudaMemcpyAsync(width.data(), d_width, sizeof(int32_t) * (size_t)n_slices, cudaMemcpyDeviceToHost, ctx->stream));
      LLZ_CUDA(cudaStreamSynchronize(ctx->stream));
      int64_t t = 0;
      for (int32_t w : width) t += (int64_t)w * kSellC;
      *total = t;
      return LLZ_OK;
    };
    auto sort_windows = [&](int sg) -> int {
      if (!d_perm) LLZ_CUDA(dev_malloc(ctx, &d_perm, sizeof(int32_t) * (size_t)n));
      const int64_t windows = (n + sg - 1) / sg;
      k_sell_sort<<<(unsigned)windows, sg, sg * sizeof(int32_t), ctx->stream>>>(d_len, n, sg, d_perm);
      return LLZ_OK;
    };
    int64_t total = 0;
    int s = LLZ_OK;
    if (sigma == 0) {
      s = widths_for(nullptr, &total);
      if (s == LLZ_OK && total > nnz + nnz / 20) {  // > 5 % padding: try sorting
        int64_t sorted_total = 0;
        s = sort_windows(256);
        if (s == LLZ_OK) s = widths_for(d_perm, &sorted_total);
        if (s == LLZ_OK && sorted_total + total / 20 < total) {
          total = sorted_total;
          sigma = 256;
        } else if (s == LLZ_OK) {
          dev_free(ctx, d_perm);
          d_perm = nullptr;
          s = widths_for(nullptr, &total);
          sigma = 1;
        }
      } else {
        sigma = 1;
      }
    } else if (sigma == 1) {
      s = widths_for(nullptr, &total);
    } else {
      s = sort_windows(sigma);
      if (s == LLZ_OK) s = widths_for(d_perm, &total);
    }
    if (s != LLZ_OK) {
      dev_free(ctx, d_len);
      dev_free(ctx, d_width);
      return s;
    }
    sell_sigma = sigma;
    padded_nnz = total;
    std::vector<int64_t> sp((size_t)n_slices + 1, 0);
    for (int64_t i = 0; i < n_slices; ++i) sp[(size_t)i + 1] = sp[(size_t)i] + (int64_t)width[(size_t)i] * kSellC;
    {  // TMA-streamed kernel: the largest chunk of 16 slices must fit a shared-memory stage, >= 2 stages in <= 200 KB
      int64_t worst = 0;
      for (int64_t c0 = 0; c0 < n_slices; c0 += kTmaChunk)
        worst = std::max(worst, sp[(size_t)std::min<int64_t>(n_slices, c0 + kTmaChunk)] - sp[(size_t)c0]);
      const size_t stage = (size_t)worst * (sizeof(T) + sizeof(int32_t));
      const char* env = getenv("LLZ_SELL_TMA");
      tma_cap = 0;
      if ((env && env[0] == '1') && worst > 0 && stage * 2 <= 200 * 1024) {  // opt-in until validated on hardware
        tma_cap = (int)worst;
        tma_stages = stage * 3 <= 100 * 1024 ? 3 : (stage * 3 <= 200 * 1024 && stage * 2 > 100 * 1024 ? 3 : 2);
      }
    }
    cudaError_t e = dev_malloc(ctx, &d_slice_ptr, sizeof(int64_t) * sp.size());
    if (e == cudaSuccess) e = dev_malloc(ctx, &d_scol, std::max<size_t>(16, sizeof(int32_t) * (size_t)total));
    if (e == cudaSuccess) e = dev_malloc(ctx, &d_sval, std::max<size_t>(16, sizeof(T) * (size_t)total));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_slice_ptr, sp.data(), sizeof(int64_t) * sp.size(), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
      k_sell_fill<T, IDX><<<wgrid, kThreads, 0, ctx->stream>>>((const IDX*)d_rowptr, d_colidx, d_vals, d_perm, d_slice_ptr, n, n_slices, d_scol, d_sval);
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    dev_free(ctx, d_len);
    dev_free(ctx, d_width);
    if (e != cudaSuccess) return fail(e == cudaErrorMemoryAllocation ? LLZ_ERR_OOM : LLZ_ERR_CUDA, "SELL conversion: %s", cudaGetErrorString(e));
    dev_free(ctx, d_rowptr);
    dev_free(ctx, d_colidx);
    dev_free(ctx, d_vals);
    d_rowptr = nullptr;
    d_colidx = nullptr;
    d_vals = nullptr;
    sell = true;
    ctx->launches += 3;
    bytes = padded_nnz * (int64_t)(sizeof(T) + 4) + (n_slices + 1) * 8 + (d_perm ? n * 4 : 0);
    return LLZ_OK;
  }

  int abs_row_sum_max(double* out) override {
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n_local + kThreads - 1) / kThreads, (int64_t)ctx->num_sms * 4));
    double* d = ctx->d_partials;  // kMaxGrid * 2 doubles
    if (dia)
      k_dia_rowsum_max<T><<<grid, kThreads, 0, ctx->stream>>>(d_dvals, dia_ld, dia_offs.nd, n_local, d);
    else if (sell)
      k_sell_rowsum_max<T><<<grid, kThreads, 0, ctx->stream>>>(d_slice_ptr, d_sval, n_slices, d);
    else if (idx32)
      k_csr_rowsum_max<T, int32_t><<<grid, kThreads, 0, ctx->stream>>>((const int32_t*)d_rowptr, d_vals, n_local, d);
    else
      k_csr_rowsum_max<T, int64_t><<<grid, kThreads, 0, ctx->stream>>>((const int64_t*)d_rowptr, d_vals, n_local, d);
    std::vector<double> h((size_t)grid);
    LLZ_CUDA(cudaMemcpyAsync(h.data(), d, sizeof(double) * (size_t)grid, cudaMemcpyDeviceToHost, ctx->stream));
    LLZ_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->launches++;
    *out = *std::max_element(h.begin(), h.end());
    return LLZ_OK;
  }

  int apply_fused(const void* x, void* y, double sigma, double* pa, int* npa, const PeerMsg* alpha_msg) override {
    const PeerMsg msg = alpha_msg ? *alpha_msg : PeerMsg();
    if (dia) return launch_dia(x, y, sigma, pa, npa, msg);
    if (sell) return launch_sell(x, y, sigma, pa, npa, msg);
    if (stream_rows > 0) return idx32 ? launch_stream<int32_t>(x, y, sigma, pa, npa, msg) : launch_stream<int64_t>(x, y, sigma, pa, npa, msg);
    return idx32 ? launch_lpr<int32_t>(x, y, sigma, pa, npa, msg) : launch_lpr<int64_t>(x, y, sigma, pa, npa, msg);
  }
};

// Row-sharded set-up: every rank learns the row ranges of the group, plans its halo (llz_halo_plan, host), tells each
// owner which of its entries it needs, and receives the peers' requests.  Returns the column indices in the local
// extended numbering.
template <class T>
static int plan_sharded_csr(CsrOp<T>* op, int64_t row0, const int64_t* rowptr, const int32_t* colidx, std::vector<int32_t>& col_local) {
  llz_ctx_t ctx = op->ctx;
  const int G = ctx->nranks, me = ctx->rank;
  const int64_t n_rows = op->n_local;
  std::vector<int64_t> ranges((size_t)G * 2);
  const int64_t mine[2] = {row0, n_rows};
  LLZ_TRY(comm_allgather_host(ctx, mine, ranges.data(), sizeof(mine)));
  std::vector<int64_t> bounds((size_t)G + 1, 0);
  for (int r = 0; r < G; ++r) {
    if (ranges[2 * r] != bounds[r])
      return fail(LLZ_ERR_INVALID, "csr: the row blocks of the ranks must be contiguous and ordered by rank (rank %d starts at %lld, expected %lld)",
                  r, (long long)ranges[2 * r], (long long)bounds[r]);
    bounds[r + 1] = bounds[r] + ranges[2 * r + 1];
  }
  if (bounds[G] != op->n_cols) return fail(LLZ_ERR_INVALID, "csr: the row blocks cover %lld rows, the operator has %lld columns", (long long)bounds[G], (long long)op->n_cols);
  const int64_t nnz = rowptr[n_rows];
  col_local.resize((size_t)std::max<int64_t>(nnz, 1));
  std::vector<int64_t> need((size_t)G, 0);
  std::vector<int32_t> halo_cols;
  LLZ_TRY(halo_plan_vectors(n_rows, row0, rowptr, colidx, G, bounds.data(), col_local.data(), halo_cols, need.data()));
  const int64_t n_halo = (int64_t)halo_cols.size();
  {  // is the halo exactly the rows just below and just above the block?  (what the DIA storage can address)
    int64_t lo = 0;
    while (lo < n_halo && halo_cols[(size_t)lo] < row0) ++lo;
    bool contig = true;
    for (int64_t t = 0; t < lo && contig; ++t) contig = halo_cols[(size_t)t] == row0 - lo + t;
    for (int64_t t = lo; t < n_halo && contig; ++t) contig = halo_cols[(size_t)t] == row0 + n_rows + (t - lo);
    op->halo_lo = (int32_t)lo;
    op->halo_contiguous = contig;
  }
  // need_all[q*G + p] = number of entries rank q needs from rank p
  std::vector<int64_t> need_all((size_t)G * G);
  LLZ_TRY(comm_allgather_host(ctx, need.data(), need_all.data(), sizeof(int64_t) * G));
  op->n_halo = n_halo;
  op->send_off.assign(G, 0);
  op->send_bytes.assign(G, 0);
  op->recv_off.assign(G, 0);
  op->recv_bytes.assign(G, 0);
  std::vector<size_t> req_off(G, 0), req_bytes(G, 0), got_off(G, 0), got_bytes(G, 0);
  int64_t n_send = 0, h = 0;
  for (int p = 0; p < G; ++p) {
    const int64_t cnt_recv = need[p], cnt_send = need_all[(size_t)p * G + me];
    op->recv_off[p] = (size_t)h * sizeof(T);
    op->recv_bytes[p] = (size_t)cnt_recv * sizeof(T);
    req_off[p] = (size_t)h * sizeof(int32_t);
    req_bytes[p] = (size_t)cnt_recv * sizeof(int32_t);
    h += cnt_recv;
    op->send_off[p] = (size_t)n_send * sizeof(T);
    op->send_bytes[p] = (size_t)cnt_send * sizeof(T);
    got_off[p] = (size_t)n_send * sizeof(int32_t);
    got_bytes[p] = (size_t)cnt_send * sizeof(int32_t);
    n_send += cnt_send;
  }
  op->n_send = n_send;
  // requests: for each halo entry the index inside its owner's block
  std::vector<int32_t> req((size_t)std::max<int64_t>(n_halo, 1));
  {
    int owner = 0;
    for (int64_t i = 0; i < n_halo; ++i) {
      while (halo_cols[i] >= bounds[owner + 1]) ++owner;
      req[(size_t)i] = (int32_t)(halo_cols[i] - bounds[owner]);
    }
  }
  // Peer-memory exchange: every rank reserves 2 x n_halo entries in its IPC window; it is used only if ALL ranks got
  // their space (the offsets are all-gathered, so the decision is the same everywhere).
  {
    const size_t mine_bytes = 2 * (size_t)std::max<int64_t>(n_halo, 1) * sizeof(T);
    int64_t my_off = comm_window_alloc(ctx, mine_bytes);
    std::vector<int64_t> offs((size_t)G, -1);
    LLZ_TRY(comm_allgather_host(ctx, &my_off, offs.data(), sizeof(int64_t)));
    bool all = true;
    for (int r = 0; r < G; ++r) all = all && offs[(size_t)r] >= 0;
    if (!all) {
      if (my_off >= 0) comm_window_free(ctx, my_off, mine_bytes);
    } else {
      op->win_off = my_off;
      op->push.G = G;
      int64_t st = 0;
      for (int q = 0; q < G; ++q) {
        op->push.start[q] = st;
        st += need_all[(size_t)q * G + me];
        // my segment inside rank q's halo buffer starts after what q receives from the ranks below me
        int64_t before = 0, q_halo = 0;
        for (int pp = 0; pp < G; ++pp) {
          if (pp < me) before += need_all[(size_t)q * G + pp];
          q_halo += need_all[(size_t)q * G + pp];
        }
        op->push.dst[q] = static_cast<char*>(comm_window_ptr(ctx, q, offs[(size_t)q])) + (size_t)before * sizeof(T);
        op->push_stride[q] = (size_t)std::max<int64_t>(q_halo, 1) * sizeof(T);
      }
      op->push.start[G] = st;
    }
  }
  int32_t* d_req = nullptr;
  LLZ_CUDA(dev_malloc(ctx, &d_req, sizeof(int32_t) * req.size()));
  LLZ_CUDA(dev_malloc(ctx, &op->d_send_idx, sizeof(int32_t) * (size_t)std::max<int64_t>(n_send, 1)));
  LLZ_CUDA(dev_malloc(ctx, &op->d_sendbuf, sizeof(T) * (size_t)std::max<int64_t>(n_send, 1)));
  LLZ_CUDA(dev_malloc(ctx, &op->d_halo, sizeof(T) * (size_t)std::max<int64_t>(n_halo, 1)));
  LLZ_CUDA(cudaMemcpyAsync(d_req, req.data(), sizeof(int32_t) * req.size(), cudaMemcpyHostToDevice, ctx->stream));
  int s = comm_exchange(ctx, (const char*)d_req, req_off.data(), req_bytes.data(), (char*)op->d_send_idx, got_off.data(), got_bytes.data());
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  dev_free(ctx, d_req);
  if (s != LLZ_OK) return s;
  if (e != cudaSuccess) return fail(LLZ_ERR_CUDA, "csr halo set-up: %s", cudaGetErrorString(e));
  return LLZ_OK;
}

template <class T>
static int create_csr(llz_ctx_t ctx, int dtype, int64_t n_rows, int64_t n_cols, int64_t row0, const int64_t* rowptr_in,
                      const int32_t* colidx_in, const void* vals, int host_arrays, int sell_sigma, llz_op_t* out) {
  auto* op = new CsrOp<T>();
  op->ctx = ctx;
  op->dtype = dtype;
  op->n_local = n_rows;
  op->n_global = n_cols;
  op->row0 = row0;
  op->n_cols = n_cols;
  int s = LLZ_OK;
  auto guard = [&](cudaError_t e, const char* what) {
    if (e != cudaSuccess && s == LLZ_OK)
      s = fail(e == cudaErrorMemoryAllocation ? LLZ_ERR_OOM : LLZ_ERR_CUDA, "csr %s: %s", what, cudaGetErrorString(e));
  };
  // the row pointers are always inspected on the host (kernel selection, 32-bit narrowing)
  std::vector<int64_t> rp_host;
  const int64_t* rowptr = rowptr_in;
  if (!host_arrays) {
    rp_host.resize((size_t)n_rows + 1);
    guard(cudaMemcpy(rp_host.data(), rowptr_in, sizeof(int64_t) * (size_t)(n_rows + 1), cudaMemcpyDeviceToHost), "rowptr fetch");
    rowptr = rp_host.data();
  }
  if (s != LLZ_OK) {
    delete op;
    return s;
  }
  const int64_t first = rowptr[0], last = rowptr[n_rows];
  bool monotone = first == 0 && last >= 0;
  for (int64_t i = 0; monotone && i < n_rows; ++i) monotone = rowptr[i + 1] >= rowptr[i];
  if (!monotone) {
    delete op;
    return fail(LLZ_ERR_INVALID, "csr: rowptr must start at 0 (got %lld) and be non-decreasing", (long long)first);
  }
  if (host_arrays && ctx->nranks == 1) {  // (row-sharded: llz_halo_plan range-checks the columns; device arrays: k_col_range below)
    for (int64_t p = 0; p < last; ++p)
      if (colidx_in[p] < 0 || colidx_in[p] >= n_cols) {
        delete op;
        return fail(LLZ_ERR_INVALID, "csr: column index %d at position %lld outside [0, %lld)", (int)colidx_in[p], (long long)p, (long long)n_cols);
      }
  }
  op->nnz = last;
  op->idx32 = last < (int64_t)0x7fffffff;

  // row-sharded: rewrite the column indices to the local extended numbering and set up the halo exchange
  const int32_t* colidx = colidx_in;
  cudaMemcpyKind col_kind = host_arrays ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
  std::vector<int32_t> col_host, col_local;
  if (ctx->nranks > 1) {
    if (!host_arrays) {
      col_host.resize((size_t)std::max<int64_t>(last, 1));
      guard(cudaMemcpy(col_host.data(), colidx_in, sizeof(int32_t) * (size_t)last, cudaMemcpyDeviceToHost), "colidx fetch");
      colidx = col_host.data();
    }
    if (s == LLZ_OK) s = plan_sharded_csr<T>(op, row0, rowptr, colidx, col_local);
    if (s != LLZ_OK) {
      delete op;
      return s;
    }
    colidx = col_local.data();
    col_kind = cudaMemcpyHostToDevice;
  }

  const cudaMemcpyKind kind = host_arrays ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
  guard(dev_malloc(ctx, &op->d_colidx, std::max<size_t>(16, sizeof(int32_t) * (size_t)last)), "colidx alloc");
  guard(dev_malloc(ctx, &op->d_vals, std::max<size_t>(16, sizeof(T) * (size_t)last)), "vals alloc");
  if (s == LLZ_OK) {
    guard(cudaMemcpyAsync(op->d_colidx, colidx, sizeof(int32_t) * (size_t)last, col_kind, ctx->stream), "colidx copy");
    guard(cudaMemcpyAsync(op->d_vals, vals, sizeof(T) * (size_t)last, kind, ctx->stream), "vals copy");
  }
  if (s == LLZ_OK) {
    if (op->idx32) {
      // narrow the row pointers to 32 bits (halves their traffic: A_bytes = nnz*(s+4) + 4(n+1))
      std::vector<int32_t> r32((size_t)n_rows + 1);
      for (int64_t i = 0; i <= n_rows; ++i) r32[(size_t)i] = (int32_t)rowptr[i];
      guard(dev_malloc(ctx, &op->d_rowptr, sizeof(int32_t) * (size_t)(n_rows + 1)), "rowptr alloc");
      if (s == LLZ_OK)
        guard(cudaMemcpyAsync(op->d_rowptr, r32.data(), sizeof(int32_t) * (size_t)(n_rows + 1), cudaMemcpyHostToDevice, ctx->stream), "rowptr copy");
    } else {
      guard(dev_malloc(ctx, &op->d_rowptr, sizeof(int64_t) * (size_t)(n_rows + 1)), "rowptr alloc");
      if (s == LLZ_OK) guard(cudaMemcpyAsync(op->d_rowptr, rowptr, sizeof(int64_t) * (size_t)(n_rows + 1), cudaMemcpyHostToDevice, ctx->stream), "rowptr copy");
    }
  }
  if (s == LLZ_OK && !host_arrays && ctx->nranks == 1 && last > 0) {  // device arrays: range-check the columns on the device
    int* d_bad = reinterpret_cast<int*>(ctx->d_result);
    guard(cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream), "memset");
    const int g = (int)std::max<int64_t>(1, std::min<int64_t>((last + kThreads - 1) / kThreads, (int64_t)ctx->num_sms * 8));
    k_col_range<<<g, kThreads, 0, ctx->stream>>>(op->d_colidx, last, (int32_t)std::min<int64_t>(n_cols, 0x7fffffff), d_bad);
    int bad = 0;
    guard(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream), "copy");
    guard(cudaStreamSynchronize(ctx->stream), "sync");
    if (s == LLZ_OK && bad) s = fail(LLZ_ERR_INVALID, "csr: a column index lies outside [0, %lld)", (long long)n_cols);
  }
  if (s == LLZ_OK) guard(cudaStreamSynchronize(ctx->stream), "sync");
  if (s != LLZ_OK) {
    delete op;
    return s;
  }
  // Stream kernel: largest power-of-two R <= 256 such that every block of R consecutive rows holds at most `cap`
  // non-zeros (cap bounded by ~40 KB of shared memory per CTA so several CTAs stay resident).
  {
    const char* env = getenv("LLZ_SPMV");
    const bool want_stream = !(env && env[0] == 'v');
    const int cap_max = (int)((40 * 1024) / sizeof(T));
    if (want_stream) {
      for (int R = 256; R >= 32; R /= 2) {
        int64_t worst = 0;
        for (int64_t r0 = 0; r0 < n_rows; r0 += R) {
          const int64_t r1 = std::min<int64_t>(n_rows, r0 + R);
          worst = std::max(worst, rowptr[r1] - rowptr[r0]);
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
  bool as_dia = false;
  {
    const char* env = getenv("LLZ_DIA");
    const int32_t* host_cols = ctx->nranks > 1 ? col_local.data() : (host_arrays ? colidx_in : nullptr);
    if (sell_sigma == 0 && host_cols && !(env && env[0] == '0')) {
      s = op->idx32 ? op->template convert_to_dia<int32_t>(rowptr, host_cols, &as_dia) : op->template convert_to_dia<int64_t>(rowptr, host_cols, &as_dia);
      if (s != LLZ_OK) {
        delete op;
        return s;
      }
    }
  }
  if (sell_sigma >= 0 && !as_dia) {
    s = op->idx32 ? op->template convert_to_sell<int32_t>(sell_sigma) : op->template convert_to_sell<int64_t>(sell_sigma);
    if (s != LLZ_OK) {
      delete op;
      return s;
    }
  }
  llz_op_t h = new llz_op_s();
  h->impl = op;
  *out = h;
  return LLZ_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// User callback adapter
// ------------------------------------------------------------------------------------------------------------------
struct CallbackOp : OpBase {
  const char* storage() const override { return "user callback"; }
  llz_apply_fn fn = nullptr;
  void* user = nullptr;
  bool overwrites = false;
  int apply_fused(const void* x, void* y, double sigma, double* pa, int* npa, const PeerMsg*) override {
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

static int create_sparse(llz_ctx_t ctx, int dtype, int64_t n_rows, int64_t n_cols, int64_t row0, const int64_t* rowptr,
                         const int32_t* colidx, const void* vals, int host_arrays, int sell_sigma, llz_op_t* op) {
  if (!ctx || !rowptr || !colidx || !vals || !op || n_rows < 1 || n_cols < 1)
    return fail(LLZ_ERR_INVALID, "op_create_csr/sell: bad argument");
  if (ctx->nranks == 1 && (row0 != 0 || n_rows != n_cols))
    return fail(LLZ_ERR_INVALID, "op_create_csr/sell: a single-rank operator must be square with row0 = 0");
  if (ctx->nranks > 1 && (row0 < 0 || row0 + n_rows > n_cols))
    return fail(LLZ_ERR_INVALID, "op_create_csr/sell: row block [%lld, %lld) outside the %lld rows of the operator", (long long)row0,
                (long long)(row0 + n_rows), (long long)n_cols);
  LLZ_CUDA(cudaSetDevice(ctx->device));
  switch (dtype) {
    case LLZ_F32: return create_csr<float>(ctx, dtype, n_rows, n_cols, row0, rowptr, colidx, vals, host_arrays, sell_sigma, op);
    case LLZ_F64: return create_csr<double>(ctx, dtype, n_rows, n_cols, row0, rowptr, colidx, vals, host_arrays, sell_sigma, op);
    case LLZ_C64: return create_csr<float2>(ctx, dtype, n_rows, n_cols, row0, rowptr, colidx, vals, host_arrays, sell_sigma, op);
    case LLZ_C128: return create_csr<double2>(ctx, dtype, n_rows, n_cols, row0, rowptr, colidx, vals, host_arrays, sell_sigma, op);
  }
  return fail(LLZ_ERR_INVALID, "unknown dtype %d", dtype);
}

int llz_op_create_csr(llz_ctx_t ctx, int dtype, int64_t n_rows, int64_t n_cols, int64_t row0, const int64_t* rowptr,
                      const int32_t* colidx, const void* vals, int host_arrays, llz_op_t* op) {
  return create_sparse(ctx, dtype, n_rows, n_cols, row0, rowptr, colidx, vals, host_arrays, -1, op);
}

int llz_op_create_sell(llz_ctx_t ctx, int dtype, int64_t n_rows, int64_t n_cols, int64_t row0, const int64_t* rowptr,
                       const int32_t* colidx, const void* vals, int host_arrays, int sigma, llz_op_t* op) {
  if (sigma < 0 || sigma > 1024 || (sigma > 1 && sigma % 32 != 0))
    return fail(LLZ_ERR_INVALID, "op_create_sell: sigma must be 0 (auto), 1 (no sorting) or a multiple of 32 up to 1024 (got %d)", sigma);
  return create_sparse(ctx, dtype, n_rows, n_cols, row0, rowptr, colidx, vals, host_arrays, sigma, op);
}

int llz_op_create_callback(llz_ctx_t ctx, int dtype, int64_t n_local, llz_apply_fn apply, void* user, int overwrites_y,
                           llz_op_t* op) {
  if (!ctx || !apply || !op || n_local < 1 || dtype_size(dtype) == 0) return fail(LLZ_ERR_INVALID, "op_create_callback: bad argument");
  auto* c = new CallbackOp();
  c->ctx = ctx;
  c->dtype = dtype;
  c->n_local = n_local;
  c->n_global = n_local;
  if (ctx->nranks > 1) {  // joined context: the callback owns a row block; the blocks are stacked in rank order
    std::vector<int64_t> all((size_t)ctx->nranks);
    int s = comm_allgather_host(ctx, &n_local, all.data(), sizeof(int64_t));
    if (s != LLZ_OK) {
      delete c;
      return s;
    }
    c->n_global = 0;
    for (int r = 0; r < ctx->nranks; ++r) {
      if (r == ctx->rank) c->row0 = c->n_global;
      c->n_global += all[(size_t)r];
    }
  }
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

int llz_op_shape(llz_op_t op, int64_t* n_local, int64_t* n_global, int64_t* row0) {
  if (!op || !op->impl) return fail(LLZ_ERR_INVALID, "null");
  if (n_local) *n_local = op->impl->n_local;
  if (n_global) *n_global = op->impl->n_global;
  if (row0) *row0 = op->impl->row0;
  return LLZ_OK;
}

int llz_op_gerschgorin_radius(llz_op_t op, double* radius) {
  if (!op || !op->impl || !radius) return fail(LLZ_ERR_INVALID, "null");
  OpBase* o = op->impl;
  double local = 0.0;
  LLZ_TRY(o->abs_row_sum_max(&local));
  std::vector<double> all((size_t)o->ctx->nranks, local);
  LLZ_TRY(comm_allgather_host(o->ctx, &local, all.data(), sizeof(double)));
  *radius = *std::max_element(all.begin(), all.end());
  return LLZ_OK;
}

const char* llz_op_storage(llz_op_t op) { return (op && op->impl) ? op->impl->storage() : ""; }

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
  LLZ_TRY(o->prepare(x->d));
  ProfScope ps(o->ctx, "spmv", (double)o->bytes + (double)o->n_local * (double)dtype_size(o->dtype) * 2);
  return o->apply_fused(x->d, y->d, 0.0, o->ctx->d_partials, &npa);
}

}  // extern "C"
