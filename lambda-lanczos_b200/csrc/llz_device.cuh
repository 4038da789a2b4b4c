// llz_device.cuh — device-side numeric traits, 128-bit streaming loads/stores and deterministic reductions shared by
// every kernel of the engine.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "llz_peer.cuh"

namespace llz {

constexpr int kThreads = 256;          // every streaming kernel uses 8 warps per CTA
constexpr int kWarps = kThreads / 32;

// ---------------------------------------------------------------------------------------------------------------
// Scalar traits.  T is the storage type of vector elements: float, double, float2 (complex<float>),
// double2 (complex<double>).  R is real_t<T> (util/common.hpp:80-102 of the reference).
// ---------------------------------------------------------------------------------------------------------------
template <class T> struct Num;
template <> struct Num<float> {
  using R = float;
  static constexpr int NC = 1;   // real components per element
  static constexpr int VEC = 4;  // elements per 128-bit access
};
template <> struct Num<double> {
  using R = double;
  static constexpr int NC = 1;
  static constexpr int VEC = 2;
};
template <> struct Num<float2> {
  using R = float;
  static constexpr int NC = 2;
  static constexpr int VEC = 2;
};
template <> struct Num<double2> {
  using R = double;
  static constexpr int NC = 2;
  static constexpr int VEC = 1;
};

__device__ __forceinline__ float zero_of(float) { return 0.f; }
__device__ __forceinline__ double zero_of(double) { return 0.0; }
__device__ __forceinline__ float2 zero_of(float2) { return make_float2(0.f, 0.f); }
__device__ __forceinline__ double2 zero_of(double2) { return make_double2(0.0, 0.0); }

__device__ __forceinline__ float add_t(float a, float b) { return a + b; }
__device__ __forceinline__ double add_t(double a, double b) { return a + b; }
__device__ __forceinline__ float2 add_t(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 add_t(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }

// acc += conj(a) * b   (the reference's inner product conjugates its FIRST argument, linear_algebra.hpp:41,49)
__device__ __forceinline__ void fma_conj(float& acc, float a, float b) { acc = fmaf(a, b, acc); }
__device__ __forceinline__ void fma_conj(double& acc, double a, double b) { acc = fma(a, b, acc); }
__device__ __forceinline__ void fma_conj(float2& acc, float2 a, float2 b) {
  acc.x = fmaf(a.x, b.x, fmaf(a.y, b.y, acc.x));
  acc.y = fmaf(a.x, b.y, fmaf(-a.y, b.x, acc.y));
}
__device__ __forceinline__ void fma_conj(double2& acc, double2 a, double2 b) {
  acc.x = fma(a.x, b.x, fma(a.y, b.y, acc.x));
  acc.y = fma(a.x, b.y, fma(-a.y, b.x, acc.y));
}

// acc -= c * v
__device__ __forceinline__ void fnma(float& acc, float c, float v) { acc = fmaf(-c, v, acc); }
__device__ __forceinline__ void fnma(double& acc, double c, double v) { acc = fma(-c, v, acc); }
__device__ __forceinline__ void fnma(float2& acc, float2 c, float2 v) {
  acc.x = fmaf(-c.x, v.x, fmaf(c.y, v.y, acc.x));
  acc.y = fmaf(-c.x, v.y, fmaf(-c.y, v.x, acc.y));
}
__device__ __forceinline__ void fnma(double2& acc, double2 c, double2 v) {
  acc.x = fma(-c.x, v.x, fma(c.y, v.y, acc.x));
  acc.y = fma(-c.x, v.y, fma(-c.y, v.x, acc.y));
}

// acc += c * v
__device__ __forceinline__ void fmadd(float& acc, float c, float v) { acc = fmaf(c, v, acc); }
__device__ __forceinline__ void fmadd(double& acc, double c, double v) { acc = fma(c, v, acc); }
__device__ __forceinline__ void fmadd(float2& acc, float2 c, float2 v) {
  acc.x = fmaf(c.x, v.x, fmaf(-c.y, v.y, acc.x));
  acc.y = fmaf(c.x, v.y, fmaf(c.y, v.x, acc.y));
}
__device__ __forceinline__ void fmadd(double2& acc, double2 c, double2 v) {
  acc.x = fma(c.x, v.x, fma(-c.y, v.y, acc.x));
  acc.y = fma(c.x, v.y, fma(c.y, v.x, acc.y));
}

// acc -= r * v with a REAL factor (three-term recurrence, lambda_lanczos.hpp:251-257)
__device__ __forceinline__ void fnma_real(float& acc, float r, float v) { acc = fmaf(-r, v, acc); }
__device__ __forceinline__ void fnma_real(double& acc, double r, double v) { acc = fma(-r, v, acc); }
__device__ __forceinline__ void fnma_real(float2& acc, float r, float2 v) {
  acc.x = fmaf(-r, v.x, acc.x);
  acc.y = fmaf(-r, v.y, acc.y);
}
__device__ __forceinline__ void fnma_real(double2& acc, double r, double2 v) {
  acc.x = fma(-r, v.x, acc.x);
  acc.y = fma(-r, v.y, acc.y);
}

__device__ __forceinline__ float scale_real(float v, float r) { return v * r; }
__device__ __forceinline__ double scale_real(double v, double r) { return v * r; }
__device__ __forceinline__ float2 scale_real(float2 v, float r) { return make_float2(v.x * r, v.y * r); }
__device__ __forceinline__ double2 scale_real(double2 v, double r) { return make_double2(v.x * r, v.y * r); }

// v * c with a full scalar of type T
__device__ __forceinline__ float mul(float c, float v) { return c * v; }
__device__ __forceinline__ double mul(double c, double v) { return c * v; }
__device__ __forceinline__ float2 mul(float2 c, float2 v) { return make_float2(c.x * v.x - c.y * v.y, c.x * v.y + c.y * v.x); }
__device__ __forceinline__ double2 mul(double2 c, double2 v) {
  return make_double2(c.x * v.x - c.y * v.y, c.x * v.y + c.y * v.x);
}

// Product and sum with one rounding each and NO fused contraction, whatever -fmad says: the arithmetic of a plain
// `out[i] += a * x[j]` loop compiled for a CPU without FMA contraction (the reference's sample-style mv_mul), so the
// SpMV kernels that use them reproduce the CPU's y bit for bit.
__device__ __forceinline__ float mul_rn(float c, float v) { return __fmul_rn(c, v); }
__device__ __forceinline__ double mul_rn(double c, double v) { return __dmul_rn(c, v); }
__device__ __forceinline__ float2 mul_rn(float2 c, float2 v) {
  return make_float2(__fsub_rn(__fmul_rn(c.x, v.x), __fmul_rn(c.y, v.y)), __fadd_rn(__fmul_rn(c.x, v.y), __fmul_rn(c.y, v.x)));
}
__device__ __forceinline__ double2 mul_rn(double2 c, double2 v) {
  return make_double2(__dsub_rn(__dmul_rn(c.x, v.x), __dmul_rn(c.y, v.y)), __dadd_rn(__dmul_rn(c.x, v.y), __dmul_rn(c.y, v.x)));
}
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float2 add_rn(float2 a, float2 b) { return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)); }
__device__ __forceinline__ double2 add_rn(double2 a, double2 b) { return make_double2(__dadd_rn(a.x, b.x), __dadd_rn(a.y, b.y)); }

__device__ __forceinline__ double abs2(float v) { return (double)v * (double)v; }
__device__ __forceinline__ double abs2(double v) { return v * v; }
__device__ __forceinline__ double abs2(float2 v) { return (double)v.x * v.x + (double)v.y * v.y; }
__device__ __forceinline__ double abs2(double2 v) { return v.x * v.x + v.y * v.y; }

// Re(conj(a) * b) in double (alpha = Re<u,Au>, lambda_lanczos.hpp:248)
__device__ __forceinline__ double re_conj_mul(float a, float b) { return (double)a * (double)b; }
__device__ __forceinline__ double re_conj_mul(double a, double b) { return a * b; }
__device__ __forceinline__ double re_conj_mul(float2 a, float2 b) { return (double)a.x * b.x + (double)a.y * b.y; }
__device__ __forceinline__ double re_conj_mul(double2 a, double2 b) { return a.x * b.x + a.y * b.y; }

// |v| as the reference's row-sum tool takes it (std::abs, also for complex entries)
__device__ __forceinline__ double abs1(float v) { return fabs((double)v); }
__device__ __forceinline__ double abs1(double v) { return fabs(v); }
__device__ __forceinline__ double abs1(float2 v) { return hypot((double)v.x, (double)v.y); }
__device__ __forceinline__ double abs1(double2 v) { return hypot(v.x, v.y); }

// component access (for reductions that treat an accumulator as NC reals)
__device__ __forceinline__ float comp(float v, int) { return v; }
__device__ __forceinline__ double comp(double v, int) { return v; }
__device__ __forceinline__ float comp(float2 v, int c) { return c ? v.y : v.x; }
__device__ __forceinline__ double comp(double2 v, int c) { return c ? v.y : v.x; }

// build T from doubles (coefficients travel as double across kernels)
template <class T> __device__ __forceinline__ T from_double(double re, double im);
template <> __device__ __forceinline__ float from_double<float>(double re, double) { return (float)re; }
template <> __device__ __forceinline__ double from_double<double>(double re, double) { return re; }
template <> __device__ __forceinline__ float2 from_double<float2>(double re, double im) { return make_float2((float)re, (float)im); }
template <> __device__ __forceinline__ double2 from_double<double2>(double re, double im) { return make_double2(re, im); }

// ---------------------------------------------------------------------------------------------------------------
// 128-bit packets.  The Krylov basis is streamed exactly once per kernel, so loads bypass L1 allocation.
// ---------------------------------------------------------------------------------------------------------------
template <class T> struct alignas(16) Pack {
  T e[Num<T>::VEC];
};

template <class T> __device__ __forceinline__ Pack<T> ld_stream(const T* p) {
  Pack<T> r;
  uint32_t* u = reinterpret_cast<uint32_t*>(&r);
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3])
               : "l"(p));
  return r;
}
// coherent 128-bit load (data written earlier in the SAME kernel launch sequence but possibly aliased with a store
// target of this kernel, e.g. the in-place update of column k)
template <class T> __device__ __forceinline__ Pack<T> ld_plain(const T* p) {
  Pack<T> r;
  uint32_t* u = reinterpret_cast<uint32_t*>(&r);
  asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]) : "l"(p));
  return r;
}
template <class T> __device__ __forceinline__ void st_pack(T* p, const Pack<T>& v) {
  const uint32_t* u = reinterpret_cast<const uint32_t*>(&v);
  asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]) : "memory");
}

// Guarded variants for the ragged tail: elements at index >= n read as zero / are not written.
template <class T> __device__ __forceinline__ Pack<T> ld_guard(const T* base, int64_t i, int64_t n) {
  Pack<T> r;
#pragma unroll
  for (int e = 0; e < Num<T>::VEC; ++e) r.e[e] = (i + e < n) ? base[i + e] : zero_of(T());
  return r;
}
template <class T> __device__ __forceinline__ void st_guard(T* base, int64_t i, int64_t n, const Pack<T>& v) {
#pragma unroll
  for (int e = 0; e < Num<T>::VEC; ++e)
    if (i + e < n) base[i + e] = v.e[e];
}

// ---------------------------------------------------------------------------------------------------------------
// Reductions.  All are fixed-order (no floating-point atomics), so every run and every CTA that repeats the same
// reduction obtains bit-identical scalars.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum over the CTA (NW warps; the streaming kernels all run kWarps); result valid in every thread.  `scratch` needs NW
// doubles.
template <int NW = kWarps> __device__ __forceinline__ double block_sum(double v, double* scratch) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();  // protect scratch from a previous use
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < NW; ++w) t += scratch[w];
  return t;
}

// Every CTA re-derives a scalar from per-CTA partials written by the previous kernel (fixed order => identical
// everywhere).  `count` partials, each `stride` doubles apart.
__device__ __forceinline__ double block_sum_partials(const double* __restrict__ p, int count, double* scratch) {
  double v = 0.0;
  for (int i = threadIdx.x; i < count; i += kThreads) v += p[i];
  return block_sum(v, scratch);
}

// Epilogue of every kernel that produces a scalar as per-CTA partials (alpha in the operators, ||u||^2 in the update):
// stores the CTA's partial; row-sharded with peer channels, the LAST CTA to finish (ticket) also sums the partials and
// delivers the rank's value to every GPU's inbox — thread p stores to rank p, fences and announces — so no separate
// reduction kernel or collective sits between the producer and the consumer.  Call with all threads of the CTA.
// `push` (optional): the kernel also stored its output vector into the peers' exchange buffers (fused all-gather);
// the last CTA announces that message too.  (The callers' block_sum barrier orders every thread's stores before thread
// 0's fence, which is cumulative.)
template <int NW = kWarps>
__device__ __forceinline__ void finish_scalar(double cta_value, double* partials, const PeerMsg& msg, double* scratch,
                                              const GatherPush* push = nullptr) {
  if (threadIdx.x == 0) partials[blockIdx.x] = cta_value;
  if (msg.ch.G == 0) return;
  __shared__ int last_cta;
  const bool pushing = push != nullptr && push->G > 0;
  if (threadIdx.x == 0) {
    if (pushing)
      __threadfence_system();
    else
      __threadfence();
    const unsigned int t = atomicAdd(msg.ticket, 1u);
    last_cta = (t == gridDim.x - 1);
    if (last_cta) *msg.ticket = 0;
  }
  __syncthreads();
  if (!last_cta) return;
  __threadfence();
  double v = 0.0;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += NW * 32) v += __ldcg(partials + i);
  v = block_sum<NW>(v, scratch);
  if ((int)threadIdx.x < msg.ch.G) {
    peer_slot(msg.ch, threadIdx.x, msg.seq, msg.ch.rank)[0] = v;
    __threadfence_system();
    peer_announce(msg.ch, threadIdx.x, msg.seq);
  } else if (pushing && (int)threadIdx.x >= 32 && (int)threadIdx.x - 32 < push->msg.ch.G) {  // a second warp, concurrently
    __threadfence_system();
    peer_announce(push->msg.ch, threadIdx.x - 32, push->msg.seq);
  }
}

// Store a 128-bit packet of the vector a kernel writes into every peer's exchange buffer as well (fused all-gather).
template <class T> __device__ __forceinline__ void push_pack(const GatherPush& push, int64_t idx, const Pack<T>& v) {
  for (int p = 0; p < push.G; ++p)
    if (p != push.rank && idx >= push.lo[p] && idx < push.hi[p]) st_pack(reinterpret_cast<T*>(push.dst[p]) + idx, v);
}
template <class T> __device__ __forceinline__ void push_guard(const GatherPush& push, int64_t idx, int64_t n, const Pack<T>& v) {
  for (int p = 0; p < push.G; ++p)
    if (p != push.rank && idx >= push.lo[p] && idx < push.hi[p]) st_guard(reinterpret_cast<T*>(push.dst[p]), idx, n, v);
}

// Transposed warp reduction: each lane holds M partial sums (M a power of two <= 32); afterwards the total of value
// index v sits in acc[0] of the lanes whose top log2(M) lane bits spell v, replicated over the low lane bits.
// Costs M-1 + log2(32/M) shuffles instead of 5*M.
template <int M, class R> __device__ __forceinline__ void warp_transpose_sum(R (&acc)[M], int lane) {
  int bit = 16;
#pragma unroll
  for (int w = M; w > 1; w >>= 1) {
    const bool upper = (lane & bit) != 0;
#pragma unroll
    for (int c = 0; c < w / 2; ++c) {
      R send = upper ? acc[c] : acc[c + w / 2];
      R keep = upper ? acc[c + w / 2] : acc[c];
      acc[c] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
    }
    bit >>= 1;
  }
#pragma unroll
  for (; bit > 0; bit >>= 1) acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], bit);
}
// value index owned by a lane after warp_transpose_sum<M>
template <int M> __device__ __forceinline__ int transpose_owner_index(int lane) {
  int v = 0, bit = 16;
#pragma unroll
  for (int w = M; w > 1; w >>= 1) {
    v = (v << 1) | ((lane & bit) ? 1 : 0);
    bit >>= 1;
  }
  return v;
}
// lanes with all remaining low bits zero write the result
template <int M> __device__ __forceinline__ bool transpose_is_writer(int lane) {
  int low = 32 / M - 1;  // mask of the low bits that were reduced by plain butterflies
  return (lane & low) == 0;
}

}  // namespace llz
