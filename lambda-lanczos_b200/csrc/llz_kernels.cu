// llz_kernels.cu — the bandwidth-bound streaming kernels of the Lanczos iteration (sm_100a).
//
//   k_project      h = [Q,V]^H w'        tall-skinny GEMV-T, w' = w - alpha u_{k-1} - beta u_{k-2} folded in registers
//   k_reduce       coef = sum over CTAs of the h partials (+ recurrence coefficients)
//   k_update       u_k = w - [Q,V] coef  tall-skinny GEMV-N, ||u_k||^2 partials in the epilogue
//   k_scale_norm   u_k *= 1/beta, publishes alpha/beta to the device bank and the pinned host mirror
//   k_recurrence   u_k = w - alpha u_{k-1} - beta u_{k-2} (+ norm partials)  [no reorthogonalisation]
//   k_combine      out_r = sum_j y_r[j] u_j for up to 5 vectors in one pass over the basis
//   k_dot, k_scale, k_axpy, k_sum_partials   the util:: vector helpers
//
// Together k_project + k_reduce + k_update replace the reference's modified Gram-Schmidt sweep
// (util::schmidt_orth, util/linear_algebra.hpp:133-144, called at lambda_lanczos.hpp:259-260): two streaming passes
// over the basis instead of 2k dependent vector passes.  Every kernel is persistent (grid <= SMs x resident CTAs),
// walks contiguous row slabs with 128-bit loads, keeps the w slab in registers across all columns, and reduces in a
// fixed order (no floating-point atomics), so results are reproducible bit for bit.
#include <algorithm>
#include <map>

#include "llz_device.cuh"
#include "llz_launch.hpp"

namespace llz {

constexpr int kCT = 8;  // columns per register tile (16 in the small-n instantiations: more loads in flight per thread)

// ------------------------------------------------------------------------------------------------------------------
struct ProjectArgs {
  const void* V;
  int64_t ld;
  const void* const* Q;
  int nq;
  int col0, ncols;
  const void* w;
  int64_t n;
  int fold;
  const double* pa;
  int npa;
  const double* beta_prev;
  double* alpha_out;
  const void* u1;  // V[:, nv-1]
  const void* u2;  // V[:, nv-2]
  double* ph;
  PeerMsg alpha_msg;
};

template <class T> __device__ __forceinline__ const T* column_ptr(const void* V, int64_t ld, const void* const* Q, int nq, int j) {
  return (j < nq) ? reinterpret_cast<const T*>(Q[j]) : reinterpret_cast<const T*>(V) + (int64_t)(j - nq) * ld;
}

template <class T, int VPT, bool FULL, int CT>
__device__ __forceinline__ void project_slab(const ProjectArgs& a, int64_t base, typename Num<T>::R alpha,
                                             typename Num<T>::R beta, double* hs_warp, int lane, double& wnorm2) {
  using R = typename Num<T>::R;
  constexpr int NC = Num<T>::NC, VEC = Num<T>::VEC, M = CT * NC;
  const int tid = threadIdx.x;
  const T* w = reinterpret_cast<const T*>(a.w);

  Pack<T> wp[VPT];
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int64_t idx = base + (int64_t)(i * kThreads + tid) * VEC;
    wp[i] = FULL ? ld_stream(w + idx) : ld_guard(w, idx, a.n);
  }
  if (a.fold >= 1) {
    const T* u1 = reinterpret_cast<const T*>(a.u1);
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int64_t idx = base + (int64_t)(i * kThreads + tid) * VEC;
      Pack<T> v = FULL ? ld_stream(u1 + idx) : ld_guard(u1, idx, a.n);
#pragma unroll
      for (int e = 0; e < VEC; ++e) fnma_real(wp[i].e[e], alpha, v.e[e]);
    }
    if (a.fold >= 2) {
      const T* u2 = reinterpret_cast<const T*>(a.u2);
#pragma unroll
      for (int i = 0; i < VPT; ++i) {
        const int64_t idx = base + (int64_t)(i * kThreads + tid) * VEC;
        Pack<T> v = FULL ? ld_stream(u2 + idx) : ld_guard(u2, idx, a.n);
#pragma unroll
        for (int e = 0; e < VEC; ++e) fnma_real(wp[i].e[e], beta, v.e[e]);
      }
    }
  }

#pragma unroll
  for (int i = 0; i < VPT; ++i)
#pragma unroll
    for (int e = 0; e < VEC; ++e) wnorm2 += abs2(wp[i].e[e]);  // ||w'||^2: the DGKS cancellation test needs it

  for (int j0 = 0; j0 < a.ncols; j0 += CT) {
    T acc[CT];
#pragma unroll
    for (int c = 0; c < CT; ++c) acc[c] = zero_of(T());
    if (j0 + CT <= a.ncols) {
#pragma unroll
      for (int c = 0; c < CT; ++c) {
        const T* col = column_ptr<T>(a.V, a.ld, a.Q, a.nq, a.col0 + j0 + c);
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
          const int64_t idx = base + (int64_t)(i * kThreads + tid) * VEC;
          Pack<T> v = FULL ? ld_stream(col + idx) : ld_guard(col, idx, a.n);
#pragma unroll
          for (int e = 0; e < VEC; ++e) fma_conj(acc[c], v.e[e], wp[i].e[e]);
        }
      }
    } else {
#pragma unroll
      for (int c = 0; c < CT; ++c) {
        if (j0 + c < a.ncols) {
          const T* col = column_ptr<T>(a.V, a.ld, a.Q, a.nq, a.col0 + j0 + c);
#pragma unroll
          for (int i = 0; i < VPT; ++i) {
            const int64_t idx = base + (int64_t)(i * kThreads + tid) * VEC;
            Pack<T> v = FULL ? ld_stream(col + idx) : ld_guard(col, idx, a.n);
#pragma unroll
            for (int e = 0; e < VEC; ++e) fma_conj(acc[c], v.e[e], wp[i].e[e]);
          }
        }
      }
    }
    R red[M];
#pragma unroll
    for (int c = 0; c < CT; ++c)
#pragma unroll
      for (int k = 0; k < NC; ++k) red[k * CT + c] = comp(acc[c], k);
    warp_transpose_sum<M>(red, lane);
    if (transpose_is_writer<M>(lane)) {
      const int v = transpose_owner_index<M>(lane);
      const int c = v % CT, k = v / CT;
      if (j0 + c < a.ncols) hs_warp[(j0 + c) * NC + k] += (double)red[0];
    }
  }
}

template <class T, int VPT, int CT = kCT>
__device__ __forceinline__ void project_body(const ProjectArgs& a, double* smem_d) {
  using R = typename Num<T>::R;
  constexpr int NC = Num<T>::NC, VEC = Num<T>::VEC;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int width = a.ncols * NC;
  double* hs = smem_d;                     // [kWarps][width]
  double* scratch = smem_d + kWarps * width;  // [kWarps]
  for (int i = tid; i < kWarps * width; i += kThreads) hs[i] = 0.0;

  R alpha = 0, beta = 0;
  if (a.fold >= 1) {
    double al;
    if (a.alpha_msg.ch.G > 0) {  // row-sharded: every rank delivered its part of <u, Au> to this GPU's inbox
      peer_wait(a.alpha_msg.ch, a.alpha_msg.seq);
      al = peer_sum(a.alpha_msg.ch, a.alpha_msg.seq, 0);
    } else {
      al = block_sum_partials(a.pa, a.npa, scratch);
    }
    alpha = (R)al;
    if (blockIdx.x == 0 && tid == 0 && a.alpha_out) *a.alpha_out = al;
    if (a.fold >= 2) beta = (R)(*a.beta_prev);
  }
  __syncthreads();

  constexpr int64_t SLAB = (int64_t)kThreads * VPT * VEC;
  const int64_t nslabs = (a.n + SLAB - 1) / SLAB;
  double* hs_warp = hs + warp * width;
  double wnorm2 = 0.0;
  for (int64_t s = blockIdx.x; s < nslabs; s += gridDim.x) {
    const int64_t base = s * SLAB;
    if (base + SLAB <= a.n)
      project_slab<T, VPT, true, CT>(a, base, alpha, beta, hs_warp, lane, wnorm2);
    else
      project_slab<T, VPT, false, CT>(a, base, alpha, beta, hs_warp, lane, wnorm2);
  }
  const double wn = block_sum(wnorm2, scratch);  // also orders the hs writes before the cross-warp sum below
  double* out = a.ph + (size_t)blockIdx.x * (width + 1);
  for (int i = tid; i < width; i += kThreads) {
    double s = 0.0;
#pragma unroll
    for (int wi = 0; wi < kWarps; ++wi) s += hs[wi * width + i];
    out[i] = s;
  }
  if (tid == 0) out[width] = wn;
}

template <class T, int VPT>
__global__ void __launch_bounds__(kThreads, 2) k_project(ProjectArgs a) {
  extern __shared__ __align__(16) double smem_d[];
  project_body<T, VPT>(a, smem_d);
}

// ------------------------------------------------------------------------------------------------------------------
struct ReduceArgs {
  const double* ph;
  int grid;
  double* wnorm2;  // receives sum of the extra ||w'||^2 slot (may be null)
  int width;  // ncols*NC of the chunk; each CTA row of ph holds width + 1 doubles
  double* coef;  // already offset to the chunk
  // row-sharded with peer channels: deliver the sums to every rank's inbox instead
  PeerMsg msg;
  int slot_base;    // payload index of the chunk's first value
  int wnorm_index;  // payload index of ||w'||^2, or -1
  int publish;      // announce the message (last chunk of the pass)
  unsigned int* ticket;
};

// 32 values per CTA; the 8 warps each add every 8th per-CTA partial (coalesced 256-byte rows), then the 8 partial sums
// are added in a fixed order.  Row-sharded: the result goes straight into every peer's inbox over NVLink and the last
// CTA to finish announces the message — no collective call between the projection and the update.
// (`block` of `nblocks`: the kernel's CTAs, or the first CTAs of the fused kernel's grid; the per-CTA partials may have
//  been written by other CTAs of the SAME launch there, hence the L2 loads)
__device__ __forceinline__ void reduce_body(const ReduceArgs& a, int block, int nblocks) {
  __shared__ double part[kWarps][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int i = block * 32 + tx;
  double s = 0.0;
  if (i <= a.width)
    for (int c = ty; c < a.grid; c += kWarps) s += __ldcg(a.ph + (size_t)c * (a.width + 1) + i);
  part[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && i <= a.width) {
    double t = 0.0;
#pragma unroll
    for (int y = 0; y < kWarps; ++y) t += part[y][tx];
    if (a.msg.ch.G > 0) {
      const int idx = (i == a.width) ? a.wnorm_index : a.slot_base + i;
      if (idx >= 0)
        for (int p = 0; p < a.msg.ch.G; ++p) peer_slot(a.msg.ch, p, a.msg.seq, a.msg.ch.rank)[idx] = t;
    } else if (i == a.width) {
      if (a.wnorm2) *a.wnorm2 = t;
    } else {
      a.coef[i] = t;
    }
  }
  if (a.msg.ch.G > 0 && a.publish) {
    __shared__ int is_last;
    __syncthreads();  // the CTA's NVLink stores happen-before thread 0's fence (bar.sync), which makes them — being
                      // cumulative — visible system-wide before the ticket is taken
    if (threadIdx.x == 0) {
      __threadfence_system();
      const unsigned int t = atomicAdd(a.ticket, 1u);
      is_last = (t == (unsigned int)nblocks - 1);
      if (is_last) *a.ticket = 0;
    }
    __syncthreads();
    // every CTA's stores are visible system-wide: thread p announces to rank p (G release-stores in parallel — one
    // thread doing all of them would pay G NVLink round trips back to back)
    if (is_last && (int)threadIdx.x < a.msg.ch.G) {
      __threadfence_system();
      peer_announce(a.msg.ch, threadIdx.x, a.msg.seq);
    }
  }
}

__global__ void __launch_bounds__(kThreads) k_reduce(ReduceArgs a) { reduce_body(a, blockIdx.x, gridDim.x); }

// ------------------------------------------------------------------------------------------------------------------
struct UpdateArgs {
  const void* V;
  int64_t ld;
  const void* const* Q;
  int nq;
  int col0, ncols;
  const void* w;
  void* out;
  int64_t n;
  const double* coef;  // full coefficient array (indexed by absolute column)
  double* pb;          // norm partials or null
  // Three-term recurrence, applied FIRST and with the very operation sequence k_project used, so that the w' this
  // kernel starts from is bit-identical to the w' the coefficients were computed from.  (Folding alpha/beta into the
  // coefficients instead would re-derive w' with different rounding, ~eps*||w|| in arbitrary directions, which is
  // fatal near breakdown where ||w'|| << ||w||.)
  int fold;                 // 0: none, 1: alpha*u1, 2: alpha*u1 + beta*u2
  const double* alpha;      // device scalars
  const double* beta_prev;
  const void* u1;           // basis columns k-1 and k-2; their projection coefficients sit at coef[cu1], coef[cu2]
  const void* u2;
  int cu1, cu2;
  double* wnorm_out;  // row-sharded: receives element wnorm_index of the coefficient message (||w'||^2), if non-null
  int wnorm_index;
  PeerMsg coef_msg;  // row-sharded: coefficients = sum over ranks of this message (element col*NC + c)
  PeerMsg norm_msg;  // row-sharded: deliver sum(pb) group-wide from the last CTA
  GatherPush push;   // row-sharded: also store `out` into the peers' exchange buffers (fused all-gather)
};

template <class T, int VPT, bool FULL, int CT>
__device__ __forceinline__ double update_slab(const UpdateArgs& a, int64_t base, const T* cs, typename Num<T>::R alpha,
                                              typename Num<T>::R beta, T h1, T h2) {
  constexpr int VEC = Num<T>::VEC;
  const int tid = threadIdx.x;
  const T* w = reinterpret_cast<const T*>(a.w);
  T* out = reinterpret_cast<T*>(a.out);
  Pack<T> acc[VPT];
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int64_t idx = base + (int64_t)(i * kThreads + tid) * VEC;
    acc[i] = FULL ? ld_plain(w + idx) : ld_guard(w, idx, a.n);
  }
  if (a.fold >= 1) {
    const T* u1 = reinterpret_cast<const T*>(a.u1);
    Pack<T> v1[VPT];
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int64_t idx = base + (int64_t)(i * kThreads + tid) * VEC;
      v1[i] = FULL ? ld_stream(u1 + idx) : ld_guard(u1, idx, a.n);
#pragma unroll
      for (int e = 0; e < VEC; ++e) fnma_real(acc[i].e[e], alpha, v1[i].e[e]);
    }
    if (a.fold >= 2) {
      const T* u2 = reinterpret_cast<const T*>(a.u2);
#pragma unroll
      for (int i = 0; i < VPT; ++i) {
        const int64_t idx = base + (int64_t)(i * kThreads + tid) * VEC;
        Pack<T> v2 = FULL ? ld_stream(u2 + idx) : ld_guard(u2, idx, a.n);
#pragma unroll
        for (int e = 0; e < VEC; ++e) fnma_real(acc[i].e[e], beta, v2.e[e]);  // acc == w' of k_project, bit for bit
#pragma unroll
        for (int e = 0; e < VEC; ++e) fnma(acc[i].e[e], h2, v2.e[e]);
      }
    }
#pragma unroll
    for (int i = 0; i < VPT; ++i)
#pragma unroll
      for (int e = 0; e < VEC; ++e) fnma(acc[i].e[e], h1, v1[i].e[e]);
  }
  int j0 = 0;
  for (; j0 + CT <= a.ncols; j0 += CT) {
    Pack<T> v[CT][VPT];
#pragma unroll
    for (int c = 0; c < CT; ++c) {
      const T* col = column_ptr<T>(a.V, a.ld, a.Q, a.nq, a.col0 + j0 + c);
#pragma unroll
      for (int i = 0; i < VPT; ++i) {
        const int64_t idx = base + (int64_t)(i * kThreads + tid) * VEC;
        v[c][i] = FULL ? ld_stream(col + idx) : ld_guard(col, idx, a.n);
      }
    }
#pragma unroll
    for (int c = 0; c < CT; ++c) {
      const T cj = cs[j0 + c];
#pragma unroll
      for (int i = 0; i < VPT; ++i)
#pragma unroll
        for (int e = 0; e < VEC; ++e) fnma(acc[i].e[e], cj, v[c][i].e[e]);
    }
  }
  for (; j0 < a.ncols; ++j0) {
    const T* col = column_ptr<T>(a.V, a.ld, a.Q, a.nq, a.col0 + j0);
    const T cj = cs[j0];
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int64_t idx = base + (int64_t)(i * kThreads + tid) * VEC;
      Pack<T> v = FULL ? ld_stream(col + idx) : ld_guard(col, idx, a.n);
#pragma unroll
      for (int e = 0; e < VEC; ++e) fnma(acc[i].e[e], cj, v.e[e]);
    }
  }
  double nrm = 0.0;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int64_t idx = base + (int64_t)(i * kThreads + tid) * VEC;
    if (FULL)
      st_pack(out + idx, acc[i]);
    else
      st_guard(out, idx, a.n, acc[i]);
    if (a.push.G > 0) {
      if (FULL)
        push_pack<T>(a.push, idx, acc[i]);
      else
        push_guard<T>(a.push, idx, a.n, acc[i]);
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) nrm += abs2(acc[i].e[e]);  // guarded lanes hold zeros
  }
  return nrm;
}

template <class T, int VPT, int CT = kCT>
__device__ __forceinline__ void update_body(const UpdateArgs& a, unsigned char* smem_u) {
  constexpr int NC = Num<T>::NC, VEC = Num<T>::VEC;
  T* cs = reinterpret_cast<T*>(smem_u);
  __shared__ double scratch[kWarps];
  const int tid = threadIdx.x;
  const bool peer = a.coef_msg.ch.G > 0;
  if (peer) peer_wait(a.coef_msg.ch, a.coef_msg.seq);
  if (peer && a.wnorm_out && a.wnorm_index >= 0 && blockIdx.x == 0 && tid == 0)
    *a.wnorm_out = peer_sum(a.coef_msg.ch, a.coef_msg.seq, a.wnorm_index);
  auto coef_at = [&](int col) -> T {
    if (peer) {
      const double re = peer_sum(a.coef_msg.ch, a.coef_msg.seq, col * NC);
      const double im = NC == 2 ? peer_sum(a.coef_msg.ch, a.coef_msg.seq, col * NC + 1) : 0.0;
      return from_double<T>(re, im);
    }
    const double* c = a.coef + (size_t)col * NC;  // (possibly written by other CTAs of this very launch: L2 loads)
    return from_double<T>(__ldcg(c), NC == 2 ? __ldcg(c + NC - 1) : 0.0);
  };
  for (int i = tid; i < a.ncols; i += kThreads) cs[i] = coef_at(a.col0 + i);
  __syncthreads();
  using R = typename Num<T>::R;
  R alpha = 0, beta = 0;
  T h1 = zero_of(T()), h2 = zero_of(T());
  if (a.fold >= 1) {
    alpha = (R)__ldcg(a.alpha);
    h1 = coef_at(a.cu1);
    if (a.fold >= 2) {
      beta = (R)(*a.beta_prev);
      h2 = coef_at(a.cu2);
    }
  }
  constexpr int64_t SLAB = (int64_t)kThreads * VPT * VEC;
  const int64_t nslabs = (a.n + SLAB - 1) / SLAB;
  double nrm = 0.0;
  for (int64_t s = blockIdx.x; s < nslabs; s += gridDim.x) {
    const int64_t base = s * SLAB;
    nrm += (base + SLAB <= a.n) ? update_slab<T, VPT, true, CT>(a, base, cs, alpha, beta, h1, h2)
                                : update_slab<T, VPT, false, CT>(a, base, cs, alpha, beta, h1, h2);
  }
  if (a.pb) {
    const double t = block_sum(nrm, scratch);
    finish_scalar(t, a.pb, a.norm_msg, scratch, &a.push);
  }
}

template <class T, int VPT>
__global__ void __launch_bounds__(kThreads, 2) k_update(UpdateArgs a) {
  extern __shared__ __align__(16) unsigned char smem_u[];
  update_body<T, VPT>(a, smem_u);
}

// ------------------------------------------------------------------------------------------------------------------
struct ScaleNormArgs {
  void* x;
  int64_t n;
  const double* pb;
  int npb;
  ScalarSink sink;
};

// beta = sqrt(||u||^2) from the per-CTA partials (or the peers' message), identically in every CTA; CTA 0 publishes the
// iteration's scalars to the device bank and the pinned host mirror.  Call with all threads of the CTA.
__device__ __forceinline__ double norm_and_publish(const ScaleNormArgs& a, double* scratch) {
  double beta2;
  if (a.sink.beta_msg.ch.G > 0) {
    peer_wait(a.sink.beta_msg.ch, a.sink.beta_msg.seq);
    beta2 = peer_sum(a.sink.beta_msg.ch, a.sink.beta_msg.seq, 0);
  } else {
    double v = 0.0;
    for (int i = threadIdx.x; i < a.npb; i += kThreads) v += __ldcg(a.pb + i);
    beta2 = block_sum(v, scratch);
  }
  const double beta = sqrt(beta2);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (a.sink.beta_out) *a.sink.beta_out = beta;
    if (a.sink.h_beta) *a.sink.h_beta = beta;
    if (a.sink.h_alpha && a.sink.alpha_in) *a.sink.h_alpha = __ldcg(a.sink.alpha_in);
    if (a.sink.h_wnorm) {
      double wn = beta;
      if (a.sink.wnorm2_in) wn = sqrt(__ldcg(a.sink.wnorm2_in));
      *a.sink.h_wnorm = wn;
    }
    if (a.sink.h_flag) {
      __threadfence_system();
      *reinterpret_cast<volatile long long*>(a.sink.h_flag) = a.sink.flag_value;
    }
  }
  return beta;
}

template <class T> __global__ void __launch_bounds__(kThreads, 4) k_scale_norm(ScaleNormArgs a) {
  using R = typename Num<T>::R;
  constexpr int VEC = Num<T>::VEC;
  __shared__ double scratch[kWarps];
  const int tid = threadIdx.x;
  const double beta = norm_and_publish(a, scratch);
  if (!(beta > 0.0) || !isfinite(beta)) return;  // breakdown: the reference leaves u_k un-normalised (:279-283)
  const R inv = (R)1 / (R)beta;                   // normalize() multiplies by T(1)/norm (linear_algebra.hpp:78-80)
  T* x = reinterpret_cast<T*>(a.x);
  const int64_t npacks = a.n / VEC;
  for (int64_t p = (int64_t)blockIdx.x * kThreads + tid; p < npacks; p += (int64_t)gridDim.x * kThreads) {
    Pack<T> v = ld_plain(x + p * VEC);
#pragma unroll
    for (int e = 0; e < VEC; ++e) v.e[e] = scale_real(v.e[e], inv);
    st_pack(x + p * VEC, v);
  }
  if (blockIdx.x == 0) {
    const int64_t i = npacks * VEC + tid;
    if (i < a.n) x[i] = scale_real(x[i], inv);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// The whole orthogonalisation step of a Lanczos iteration in ONE launch: projection, reduction of the per-CTA partials,
// update, norm and normalisation (k_project + k_reduce + k_update + k_scale_norm) separated by grid-wide barriers
// instead of kernel boundaries.  Launched cooperatively (every CTA resident), one or two CTAs per SM as before; a
// barrier costs one L2 atomic and a poll (~2 us) where a launch boundary costs the drain of the grid, the launch
// latency and the host's launch work — at n = 100k that is half of the iteration, and row-sharded it is what the
// three latency-bound launches per iteration were.  Same arithmetic, same order: bit-identical to the separate kernels.
// ------------------------------------------------------------------------------------------------------------------
struct OrthArgs {
  ProjectArgs p;
  ReduceArgs r;
  int nred;  // CTAs that take part in the reduction (32 values each)
  UpdateArgs u;
  ScaleNormArgs s;
  HaloPushPlan halo;            // row-sharded sparse operators: push the normalised vector's halo entries to the peers
  unsigned long long* bar;      // device counter, monotonic across launches
  unsigned long long bar_base;  // its value when this launch starts
};

__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// Grid-wide barrier number `index` (1, 2, ...) of this launch.  Every CTA is resident (cooperative launch).
__device__ __forceinline__ void grid_barrier(const OrthArgs& a, int index) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();  // this CTA's global writes (ordered before by the barrier above) are visible before it arrives
    atomicAdd(a.bar, 1ull);
    const unsigned long long target = a.bar_base + (unsigned long long)index * gridDim.x;
    while (ld_acquire_gpu(a.bar) < target) {
    }
  }
  __syncthreads();
}

template <class T, int VPT, int CT>
__global__ void __launch_bounds__(kThreads, 2) k_orth(OrthArgs a) {
  using R = typename Num<T>::R;
  constexpr int VEC = Num<T>::VEC;
  extern __shared__ __align__(16) unsigned char smem_o[];
  __shared__ double scratch[kWarps];
  project_body<T, VPT, CT>(a.p, reinterpret_cast<double*>(smem_o));
  grid_barrier(a, 1);
  if ((int)blockIdx.x < a.nred) reduce_body(a.r, blockIdx.x, a.nred);
  grid_barrier(a, 2);
  update_body<T, VPT, CT>(a.u, smem_o);  // (row-sharded: waits for the peers' coefficient messages in its prologue)
  grid_barrier(a, 3);
  const double beta = norm_and_publish(a.s, scratch);
  const bool scale = beta > 0.0 && isfinite(beta);  // breakdown: the reference leaves u_k un-normalised (:279-283)
  const R inv = scale ? (R)1 / (R)beta : (R)1;
  // every thread re-scales exactly the packets it wrote in the update phase (same slab walk)
  T* x = reinterpret_cast<T*>(a.u.out);
  constexpr int64_t SLAB = (int64_t)kThreads * VPT * VEC;
  const int64_t nslabs = (a.u.n + SLAB - 1) / SLAB;
  for (int64_t sl = blockIdx.x; scale && sl < nslabs; sl += gridDim.x) {
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int64_t idx = sl * SLAB + (int64_t)(i * kThreads + threadIdx.x) * VEC;
      if (idx + VEC <= a.u.n) {
        Pack<T> v = ld_plain(x + idx);
#pragma unroll
        for (int e = 0; e < VEC; ++e) v.e[e] = scale_real(v.e[e], inv);
        st_pack(x + idx, v);
      } else {
#pragma unroll
        for (int e = 0; e < VEC; ++e)
          if (idx + e < a.u.n) x[idx + e] = scale_real(x[idx + e], inv);
      }
    }
  }
  if (a.halo.hp.G == 0) return;
  // ---- halo of the next operator application: the normalised entries the peers reference, straight into their halo
  //      segments over NVLink (what k_halo_push would do in a launch of its own before the apply) ----
  grid_barrier(a, 4);  // every packet of the vector is final (and visible: the loads below go to L2)
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < a.halo.n_send; i += (long long)gridDim.x * kThreads) {
    int q = 0;
    while (i >= a.halo.hp.start[q + 1]) ++q;
    reinterpret_cast<T*>(a.halo.hp.dst[q])[i - a.halo.hp.start[q]] = __ldcg(x + a.halo.idx[i]);
  }
  __shared__ int last_cta;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();  // cumulative: covers the stores of the whole CTA ordered before it by the barrier
    const unsigned int t = atomicAdd(a.halo.msg.ticket, 1u);
    last_cta = (t == gridDim.x - 1);
    if (last_cta) *a.halo.msg.ticket = 0;
  }
  __syncthreads();
  if (last_cta && (int)threadIdx.x < a.halo.msg.ch.G) {
    __threadfence_system();
    peer_announce(a.halo.msg.ch, threadIdx.x, a.halo.msg.seq);
  }
}

// ------------------------------------------------------------------------------------------------------------------
struct RecurrenceArgs {
  const void* w;
  const void* u1;
  const void* u2;
  void* out;
  int64_t n;
  int fold;
  const double* pa;
  int npa;
  const double* beta_prev;
  double* alpha_out;
  double* pb;
  PeerMsg alpha_msg;
  PeerMsg norm_msg;
  GatherPush push;
  // lazy normalisation (LLZ_ORTH_RECURRENCE_LAZY): u1 / u2 are stored scaled by *scale1 / *scale2 (null: 1), w = A u1
  // carries scale1 too; the output stays un-normalised and the LAST CTA finishes ||out|| and publishes per `sink`
  int lazy;
  const double* scale1;
  const double* scale2;
  ScalarSink sink;
  unsigned int* ticket;
};

template <class T> __global__ void __launch_bounds__(kThreads, 4) k_recurrence(RecurrenceArgs a) {
  using R = typename Num<T>::R;
  constexpr int VEC = Num<T>::VEC;
  __shared__ double scratch[kWarps];
  const int tid = threadIdx.x;
  R alpha = 0, beta = 0;
  if (a.fold >= 1) {
    double al;
    if (a.alpha_msg.ch.G > 0) {
      peer_wait(a.alpha_msg.ch, a.alpha_msg.seq);
      al = peer_sum(a.alpha_msg.ch, a.alpha_msg.seq, 0);
    } else {
      al = block_sum_partials(a.pa, a.npa, scratch);
    }
    R cw = (R)1;  // factor of w (lazy mode)
    double alpha_true = al;
    if (a.lazy) {
      // stored: x~ = s1 u1, u2~ = s2 u2, w~ = A x~.  With beta_prev = s1:
      //   r = w~/s1 - (alpha~/s1^3) x~ - (s1/s2) u2~,   alpha = alpha~/s1^2
      const double s1 = a.scale1 ? *a.scale1 : 1.0, s2 = a.scale2 ? *a.scale2 : 1.0;
      alpha_true = al / (s1 * s1);
      cw = (R)(1.0 / s1);
      al = alpha_true / s1;
      if (a.fold >= 2) beta = (R)(s1 / s2);
    } else if (a.fold >= 2) {
      beta = (R)(*a.beta_prev);
    }
    alpha = (R)al;
    if (blockIdx.x == 0 && tid == 0 && a.alpha_out) *a.alpha_out = alpha_true;
    if (a.lazy) {
      const T* w = reinterpret_cast<const T*>(a.w);
      const T* u1 = reinterpret_cast<const T*>(a.u1);
      const T* u2 = reinterpret_cast<const T*>(a.u2);
      T* out = reinterpret_cast<T*>(a.out);
      const int64_t npacks = (a.n + VEC - 1) / VEC;
      double nrm = 0.0;
      for (int64_t p = (int64_t)blockIdx.x * kThreads + tid; p < npacks; p += (int64_t)gridDim.x * kThreads) {
        const int64_t idx = p * VEC;
        const bool full = idx + VEC <= a.n;
        Pack<T> acc = full ? ld_plain(w + idx) : ld_guard(w, idx, a.n);
        Pack<T> v = full ? ld_stream(u1 + idx) : ld_guard(u1, idx, a.n);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          acc.e[e] = scale_real(acc.e[e], cw);
          fnma_real(acc.e[e], alpha, v.e[e]);
        }
        if (a.fold >= 2) {
          Pack<T> v2 = full ? ld_stream(u2 + idx) : ld_guard(u2, idx, a.n);
#pragma unroll
          for (int e = 0; e < VEC; ++e) fnma_real(acc.e[e], beta, v2.e[e]);
        }
        if (full)
          st_pack(out + idx, acc);
        else
          st_guard(out, idx, a.n, acc);
#pragma unroll
        for (int e = 0; e < VEC; ++e) nrm += abs2(acc.e[e]);
      }
      // per-CTA partial; the last CTA to arrive sums them in index order (every run: the same bits) and publishes
      const double t = block_sum(nrm, scratch);
      __shared__ int last_cta;
      if (tid == 0) {
        a.pb[blockIdx.x] = t;
        __threadfence();
        const unsigned int tk = atomicAdd(a.ticket, 1u);
        last_cta = (tk == gridDim.x - 1);
        if (last_cta) *a.ticket = 0;
      }
      __syncthreads();
      if (!last_cta) return;
      __threadfence();
      double v = 0.0;
      for (int i = tid; i < (int)gridDim.x; i += kThreads) v += __ldcg(a.pb + i);
      const double beta_new = sqrt(block_sum(v, scratch));
      if (tid == 0) {
        if (a.sink.beta_out) *a.sink.beta_out = beta_new;
        if (a.sink.h_beta) *a.sink.h_beta = beta_new;
        if (a.sink.h_alpha) *a.sink.h_alpha = alpha_true;
        if (a.sink.h_wnorm) *a.sink.h_wnorm = beta_new;
        if (a.sink.h_flag) {
          __threadfence_system();
          *reinterpret_cast<volatile long long*>(a.sink.h_flag) = a.sink.flag_value;
        }
      }
      return;
    }
  }
  const T* w = reinterpret_cast<const T*>(a.w);
  const T* u1 = reinterpret_cast<const T*>(a.u1);
  const T* u2 = reinterpret_cast<const T*>(a.u2);
  T* out = reinterpret_cast<T*>(a.out);
  const int64_t npacks = (a.n + VEC - 1) / VEC;
  double nrm = 0.0;
  for (int64_t p = (int64_t)blockIdx.x * kThreads + tid; p < npacks; p += (int64_t)gridDim.x * kThreads) {
    const int64_t idx = p * VEC;
    const bool full = idx + VEC <= a.n;
    Pack<T> acc = full ? ld_plain(w + idx) : ld_guard(w, idx, a.n);
    if (a.fold >= 1) {
      Pack<T> v = full ? ld_stream(u1 + idx) : ld_guard(u1, idx, a.n);
#pragma unroll
      for (int e = 0; e < VEC; ++e) fnma_real(acc.e[e], alpha, v.e[e]);
    }
    if (a.fold >= 2) {
      Pack<T> v = full ? ld_stream(u2 + idx) : ld_guard(u2, idx, a.n);
#pragma unroll
      for (int e = 0; e < VEC; ++e) fnma_real(acc.e[e], beta, v.e[e]);
    }
    if (full)
      st_pack(out + idx, acc);
    else
      st_guard(out, idx, a.n, acc);
    if (a.push.G > 0) {
      if (full)
        push_pack<T>(a.push, idx, acc);
      else
        push_guard<T>(a.push, idx, a.n, acc);
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) nrm += abs2(acc.e[e]);
  }
  if (a.pb) {
    const double t = block_sum(nrm, scratch);
    finish_scalar(t, a.pb, a.norm_msg, scratch, &a.push);
  }
}

// ------------------------------------------------------------------------------------------------------------------
constexpr int kMaxCombine = 5;  // num_eigs_per_iteration default (lambda_lanczos.hpp:173)

struct CombineArgs {
  const void* V;
  int64_t ld;
  int col0, ncols;
  const void* coef;  // device, T[nvec][ldc]
  int64_t ldc;
  int nvec;
  void* out[kMaxCombine];
  int64_t n;
  int accumulate;
  double* pb;  // [nvec][kMaxGrid] or null
};

template <class T, int NV, bool FULL>
__device__ __forceinline__ void combine_slab(const CombineArgs& a, int64_t base, const T* cs, double (&nrm)[NV]) {
  constexpr int VEC = Num<T>::VEC;
  const int tid = threadIdx.x;
  const int64_t idx = base + (int64_t)tid * VEC;
  Pack<T> acc[NV];
#pragma unroll
  for (int r = 0; r < NV; ++r) {
    if (a.accumulate && r < a.nvec) {
      const T* o = reinterpret_cast<const T*>(a.out[r]);
      acc[r] = FULL ? ld_plain(o + idx) : ld_guard(o, idx, a.n);
    } else {
#pragma unroll
      for (int e = 0; e < VEC; ++e) acc[r].e[e] = zero_of(T());
    }
  }
  const T* V = reinterpret_cast<const T*>(a.V) + (int64_t)a.col0 * a.ld;
  int j0 = 0;
  for (; j0 + 4 <= a.ncols; j0 += 4) {
    Pack<T> v[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const T* col = V + (int64_t)(j0 + c) * a.ld;
      v[c] = FULL ? ld_stream(col + idx) : ld_guard(col, idx, a.n);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int r = 0; r < NV; ++r) {
        const T y = cs[r * a.ncols + j0 + c];
#pragma unroll
        for (int e = 0; e < VEC; ++e) fmadd(acc[r].e[e], y, v[c].e[e]);
      }
  }
  for (; j0 < a.ncols; ++j0) {
    const T* col = V + (int64_t)j0 * a.ld;
    Pack<T> v = FULL ? ld_stream(col + idx) : ld_guard(col, idx, a.n);
#pragma unroll
    for (int r = 0; r < NV; ++r) {
      const T y = cs[r * a.ncols + j0];
#pragma unroll
      for (int e = 0; e < VEC; ++e) fmadd(acc[r].e[e], y, v.e[e]);
    }
  }
#pragma unroll
  for (int r = 0; r < NV; ++r) {
    if (r < a.nvec) {
      T* o = reinterpret_cast<T*>(a.out[r]);
      if (FULL)
        st_pack(o + idx, acc[r]);
      else
        st_guard(o, idx, a.n, acc[r]);
#pragma unroll
      for (int e = 0; e < VEC; ++e) nrm[r] += abs2(acc[r].e[e]);
    }
  }
}

template <class T, int NV> __global__ void __launch_bounds__(kThreads, 2) k_combine(CombineArgs a) {
  constexpr int VEC = Num<T>::VEC;
  extern __shared__ __align__(16) unsigned char smem_c[];
  T* cs = reinterpret_cast<T*>(smem_c);  // [NV][ncols], zero rows beyond nvec
  __shared__ double scratch[kWarps];
  const int tid = threadIdx.x;
  const T* coef = reinterpret_cast<const T*>(a.coef);
  for (int i = tid; i < NV * a.ncols; i += kThreads) {
    const int r = i / a.ncols, j = i % a.ncols;
    cs[i] = (r < a.nvec) ? coef[(int64_t)r * a.ldc + a.col0 + j] : zero_of(T());
  }
  __syncthreads();
  constexpr int64_t SLAB = (int64_t)kThreads * VEC;
  const int64_t nslabs = (a.n + SLAB - 1) / SLAB;
  double nrm[NV];
#pragma unroll
  for (int r = 0; r < NV; ++r) nrm[r] = 0.0;
  for (int64_t s = blockIdx.x; s < nslabs; s += gridDim.x) {
    const int64_t base = s * SLAB;
    if (base + SLAB <= a.n)
      combine_slab<T, NV, true>(a, base, cs, nrm);
    else
      combine_slab<T, NV, false>(a, base, cs, nrm);
  }
  if (a.pb) {
#pragma unroll
    for (int r = 0; r < NV; ++r) {
      if (r < a.nvec) {
        const double t = block_sum(nrm[r], scratch);
        if (tid == 0) a.pb[(size_t)r * kMaxGrid + blockIdx.x] = t;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
template <class T> __global__ void __launch_bounds__(kThreads, 4) k_dot(const T* __restrict__ x, const T* __restrict__ y, int64_t n, double* partials) {
  constexpr int NC = Num<T>::NC, VEC = Num<T>::VEC;
  __shared__ double scratch[kWarps];
  const int tid = threadIdx.x;
  T acc = zero_of(T());
  double re = 0.0, im = 0.0;
  const int64_t npacks = (n + VEC - 1) / VEC;
  for (int64_t p = (int64_t)blockIdx.x * kThreads + tid; p < npacks; p += (int64_t)gridDim.x * kThreads) {
    const int64_t idx = p * VEC;
    const bool full = idx + VEC <= n;
    Pack<T> a = full ? ld_plain(x + idx) : ld_guard(x, idx, n);
    Pack<T> b = full ? ld_plain(y + idx) : ld_guard(y, idx, n);
    acc = zero_of(T());
#pragma unroll
    for (int e = 0; e < VEC; ++e) fma_conj(acc, a.e[e], b.e[e]);
    re += (double)comp(acc, 0);
    if (NC == 2) im += (double)comp(acc, 1);
  }
  const double tre = block_sum(re, scratch);
  if (tid == 0) partials[(size_t)blockIdx.x * NC] = tre;
  if (NC == 2) {
    const double tim = block_sum(im, scratch);
    if (tid == 0) partials[(size_t)blockIdx.x * NC + 1] = tim;
  }
}

template <class T> __global__ void __launch_bounds__(kThreads, 4) k_redot(const T* __restrict__ x, const T* __restrict__ y, int64_t n, double* partials, PeerMsg msg) {
  constexpr int VEC = Num<T>::VEC;
  __shared__ double scratch[kWarps];
  double re = 0.0;
  const int64_t npacks = (n + VEC - 1) / VEC;
  for (int64_t p = (int64_t)blockIdx.x * kThreads + threadIdx.x; p < npacks; p += (int64_t)gridDim.x * kThreads) {
    const int64_t idx = p * VEC;
    const bool full = idx + VEC <= n;
    Pack<T> a = full ? ld_plain(x + idx) : ld_guard(x, idx, n);
    Pack<T> b = full ? ld_plain(y + idx) : ld_guard(y, idx, n);
#pragma unroll
    for (int e = 0; e < VEC; ++e) re += re_conj_mul(a.e[e], b.e[e]);
  }
  const double t = block_sum(re, scratch);
  finish_scalar(t, partials, msg, scratch);
}

// per-CTA partials of sum_i |Re v_i| + |Im v_i|  (util::m_norm, util/linear_algebra.hpp:83-125: the _ASUM definition)
template <class T> __global__ void __launch_bounds__(kThreads, 4) k_asum(const T* __restrict__ x, int64_t n, double* partials) {
  constexpr int NC = Num<T>::NC, VEC = Num<T>::VEC;
  __shared__ double scratch[kWarps];
  double s = 0.0;
  const int64_t npacks = (n + VEC - 1) / VEC;
  for (int64_t p = (int64_t)blockIdx.x * kThreads + threadIdx.x; p < npacks; p += (int64_t)gridDim.x * kThreads) {
    const int64_t idx = p * VEC;
    Pack<T> a = (idx + VEC <= n) ? ld_plain(x + idx) : ld_guard(x, idx, n);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      s += fabs((double)comp(a.e[e], 0));
      if (NC == 2) s += fabs((double)comp(a.e[e], 1));
    }
  }
  const double t = block_sum(s, scratch);
  if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

// result[c] = sum_i partials[i*nc + c]; also mirrored to pinned host memory when h_result != null
__global__ void __launch_bounds__(kThreads) k_sum_partials(const double* partials, int count, int nc, double* result, double* h_result) {
  __shared__ double scratch[kWarps];
  for (int c = 0; c < nc; ++c) {
    double v = 0.0;
    for (int i = threadIdx.x; i < count; i += kThreads) v += partials[(size_t)i * nc + c];
    const double t = block_sum(v, scratch);
    if (threadIdx.x == 0) {
      if (result) result[c] = t;
      if (h_result) h_result[c] = t;
    }
  }
}

template <class T> __global__ void __launch_bounds__(kThreads, 4) k_scale(T* x, int64_t n, T a) {
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) x[i] = mul(a, x[i]);
}
template <class T> __global__ void __launch_bounds__(kThreads, 4) k_axpy(T* y, T a, const T* __restrict__ x, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    T acc = y[i];
    fmadd(acc, a, x[i]);
    y[i] = acc;
  }
}

// ==================================================================================================================
// Host launchers
// ==================================================================================================================
namespace {

constexpr size_t kSmemBudget = 100 * 1024;  // per CTA; two CTAs per SM stay resident

inline int persistent_grid(llz_ctx_t ctx, int64_t work_items, int ctas_per_sm) {
  int64_t g = (int64_t)ctx->num_sms * ctas_per_sm;
  if (g > kMaxGrid) g = kMaxGrid;
  if (g > work_items) g = work_items;
  if (g < 1) g = 1;
  return (int)g;
}

// Rows one CTA covers per slab with VPT packets per thread.
template <class T> constexpr int64_t slab_rows(int vpt) { return (int64_t)kThreads * vpt * Num<T>::VEC; }

// VPT = 2 once there is enough work to fill the machine twice over, else 1 (keeps small problems spread over SMs).
template <class T> inline int pick_vpt(llz_ctx_t ctx, int64_t n) {
  return (n >= slab_rows<T>(2) * ctx->num_sms * 4) ? 2 : 1;
}

template <class F> inline int set_smem(F* kernel, size_t bytes) {
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return fail(LLZ_ERR_CUDA, "cudaFuncSetAttribute(smem=%zu): %s", bytes, cudaGetErrorString(e));
  }
  return LLZ_OK;
}

inline int check_launch(llz_ctx_t ctx, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(LLZ_ERR_CUDA, "launch %s: %s", what, cudaGetErrorString(e));
  ctx->launches++;
  return LLZ_OK;
}

}  // namespace

template <class T> inline T from_double_host(const double a[2]);
template <> inline float from_double_host<float>(const double a[2]) { return (float)a[0]; }
template <> inline double from_double_host<double>(const double a[2]) { return a[0]; }
template <> inline float2 from_double_host<float2>(const double a[2]) { return make_float2((float)a[0], (float)a[1]); }
template <> inline double2 from_double_host<double2>(const double a[2]) { return make_double2(a[0], a[1]); }

#define LLZ_DISPATCH(dtype, ...)                                               \
  switch (dtype) {                                                             \
    case LLZ_F32: { using T = float; __VA_ARGS__; } break;                     \
    case LLZ_F64: { using T = double; __VA_ARGS__; } break;                    \
    case LLZ_C64: { using T = float2; __VA_ARGS__; } break;                    \
    case LLZ_C128: { using T = double2; __VA_ARGS__; } break;                  \
    default: return fail(LLZ_ERR_INVALID, "unknown dtype %d", dtype);          \
  }

int max_project_cols(int dtype) {
  const size_t per_col = (size_t)kWarps * dtype_nc(dtype) * sizeof(double);
  return (int)((kSmemBudget - kWarps * sizeof(double)) / per_col);
}
int max_update_cols(int dtype) { return (int)(kSmemBudget / dtype_size(dtype)); }
int max_combine_cols(int dtype, int nvec) {
  (void)nvec;
  return (int)(kSmemBudget / (dtype_size(dtype) * kMaxCombine));
}

template <class T, int VPT>
static int project_impl(llz_ctx_t ctx, const ProjectArgs& a, int* grid_out) {
  const size_t smem = ((size_t)kWarps * a.ncols * Num<T>::NC + kWarps) * sizeof(double);
  LLZ_TRY(set_smem(k_project<T, VPT>, smem));
  const int64_t nslabs = (a.n + slab_rows<T>(VPT) - 1) / slab_rows<T>(VPT);
  const int grid = persistent_grid(ctx, nslabs, 2);
  cudaError_t le = launch_chain(ctx, k_project<T, VPT>, grid, kThreads, smem, a);
  *grid_out = grid;
  if (le != cudaSuccess) return fail(LLZ_ERR_CUDA, "launch k_project: %s", cudaGetErrorString(le));
  return check_launch(ctx, "k_project");
}

static int make_project_args(int dtype, const ColumnSet& cs, int col0, int ncols, const void* w, int64_t n, const Fold& fold,
                             double* ph, ProjectArgs& a) {
  if (ncols < 0 || ncols > max_project_cols(dtype)) return fail(LLZ_ERR_INVALID, "project: %d columns per launch", ncols);
  a.V = cs.V;
  a.ld = cs.ld;
  a.Q = cs.Q;
  a.nq = cs.nq;
  a.col0 = col0;
  a.ncols = ncols;
  a.w = w;
  a.n = n;
  a.fold = fold.mode;
  a.pa = fold.alpha_partials;
  a.npa = fold.n_partials;
  a.beta_prev = fold.beta_prev;
  a.alpha_out = fold.alpha_out;
  a.alpha_msg = fold.alpha_msg;
  a.ph = ph;
  const size_t es = dtype_size(dtype);
  a.u1 = cs.nv >= 1 ? (const char*)cs.V + (size_t)(cs.nv - 1) * cs.ld * es : nullptr;
  a.u2 = cs.nv >= 2 ? (const char*)cs.V + (size_t)(cs.nv - 2) * cs.ld * es : nullptr;
  if (a.fold >= 1 && !a.u1) return fail(LLZ_ERR_INVALID, "project: fold needs a previous basis column");
  if (a.fold >= 2 && !a.u2) return fail(LLZ_ERR_INVALID, "project: fold=2 needs two previous basis columns");
  return LLZ_OK;
}

int launch_project(llz_ctx_t ctx, int dtype, const ColumnSet& cs, int col0, int ncols, const void* w, int64_t n,
                   const Fold& fold, double* ph, int* grid_out) {
  ProjectArgs a;
  LLZ_TRY(make_project_args(dtype, cs, col0, ncols, w, n, fold, ph, a));
  LLZ_DISPATCH(dtype, {
    if (pick_vpt<T>(ctx, n) == 2) return project_impl<T, 2>(ctx, a, grid_out);
    return project_impl<T, 1>(ctx, a, grid_out);
  });
  return LLZ_OK;
}

static void make_reduce_args(int dtype, const double* ph, int grid, int col0, int ncols, double* coef, double* wnorm2,
                             const PeerMsg& msg, int wnorm_index, int publish, ReduceArgs& a) {
  const int nc = dtype_nc(dtype);
  a.ph = ph;
  a.grid = grid;
  a.wnorm2 = wnorm2;
  a.width = ncols * nc;
  a.coef = coef + (size_t)col0 * nc;
  a.msg = msg;
  a.slot_base = col0 * nc;
  a.wnorm_index = wnorm_index;
  a.publish = publish;
  a.ticket = msg.ticket;
}

int launch_reduce(llz_ctx_t ctx, int dtype, const double* ph, int grid, int col0, int ncols, double* coef, double* wnorm2,
                  const PeerMsg& msg, int wnorm_index, int publish) {
  ReduceArgs a;
  make_reduce_args(dtype, ph, grid, col0, ncols, coef, wnorm2, msg, wnorm_index, publish, a);
  cudaError_t le = launch_chain(ctx, k_reduce, (a.width + 1 + 31) / 32, kThreads, 0, a);
  if (le != cudaSuccess) return fail(LLZ_ERR_CUDA, "launch k_reduce: %s", cudaGetErrorString(le));
  return check_launch(ctx, "k_reduce");
}

template <class T, int VPT>
static int update_impl(llz_ctx_t ctx, const UpdateArgs& a, int* grid_out) {
  const size_t smem = std::max<size_t>(16, (size_t)a.ncols * sizeof(T));
  LLZ_TRY(set_smem(k_update<T, VPT>, smem));
  const int64_t nslabs = (a.n + slab_rows<T>(VPT) - 1) / slab_rows<T>(VPT);
  const int grid = persistent_grid(ctx, nslabs, 2);
  cudaError_t le = launch_chain(ctx, k_update<T, VPT>, grid, kThreads, smem, a);
  *grid_out = grid;
  if (le != cudaSuccess) return fail(LLZ_ERR_CUDA, "launch k_update: %s", cudaGetErrorString(le));
  return check_launch(ctx, "k_update");
}

static int make_update_args(int dtype, const ColumnSet& cs, int col0, int ncols, const void* w, void* out, int64_t n,
                            const double* coef, const Fold& fold, double* norm_partials, const PeerMsg& coef_msg,
                            const PeerMsg& norm_msg, UpdateArgs& a) {
  if (ncols < 0 || ncols > max_update_cols(dtype)) return fail(LLZ_ERR_INVALID, "update: %d columns per launch", ncols);
  a.V = cs.V;
  a.ld = cs.ld;
  a.Q = cs.Q;
  a.nq = cs.nq;
  a.col0 = col0;
  a.ncols = ncols;
  a.w = w;
  a.out = out;
  a.n = n;
  a.coef = coef;
  a.coef_msg = coef_msg;
  a.wnorm_out = fold.wnorm_out;
  a.wnorm_index = fold.wnorm_index;
  a.norm_msg = norm_partials ? norm_msg : PeerMsg();
  a.push = (norm_partials && a.norm_msg.ch.G > 0) ? fold.push : GatherPush();
  a.pb = norm_partials;
  a.fold = fold.mode;
  a.alpha = fold.alpha_out;
  a.beta_prev = fold.beta_prev;
  const size_t es = dtype_size(dtype);
  a.u1 = cs.nv >= 1 ? (const char*)cs.V + (size_t)(cs.nv - 1) * cs.ld * es : nullptr;
  a.u2 = cs.nv >= 2 ? (const char*)cs.V + (size_t)(cs.nv - 2) * cs.ld * es : nullptr;
  a.cu1 = cs.nq + cs.nv - 1;
  a.cu2 = cs.nq + cs.nv - 2;
  if (a.fold >= 1 && (!a.u1 || !a.alpha)) return fail(LLZ_ERR_INVALID, "update: fold needs a previous basis column and alpha");
  if (a.fold >= 2 && (!a.u2 || !a.beta_prev)) return fail(LLZ_ERR_INVALID, "update: fold=2 needs two previous basis columns and beta");
  return LLZ_OK;
}

int launch_update(llz_ctx_t ctx, int dtype, const ColumnSet& cs, int col0, int ncols, const void* w, void* out,
                  int64_t n, const double* coef, const Fold& fold, double* norm_partials, int* grid_out,
                  const PeerMsg& coef_msg, const PeerMsg& norm_msg) {
  UpdateArgs a;
  LLZ_TRY(make_update_args(dtype, cs, col0, ncols, w, out, n, coef, fold, norm_partials, coef_msg, norm_msg, a));
  LLZ_DISPATCH(dtype, {
    if (pick_vpt<T>(ctx, n) == 2) return update_impl<T, 2>(ctx, a, grid_out);
    return update_impl<T, 1>(ctx, a, grid_out);
  });
  return LLZ_OK;
}

// The fused orthogonalisation step (k_orth): one cooperative launch for project + reduce + update + norm + normalise of
// a full-reorthogonalisation iteration whose columns fit one projection chunk.  *fused = 0 (nothing launched) when the
// shape does not allow it (too many columns for the shared-memory accumulators, or the grid would not be co-resident).
// Resident CTAs per SM of k_orth for a given amount of dynamic shared memory (cached: the query costs microseconds).
template <class T, int VPT, int CT> static int orth_ctas_per_sm(size_t smem) {
  static std::map<size_t, int> cache;
  auto it = cache.find(smem);
  if (it != cache.end()) return it->second;
  int per_sm = 0;
  if (set_smem(k_orth<T, VPT, CT>, smem) != LLZ_OK) return 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_orth<T, VPT, CT>, kThreads, smem) != cudaSuccess) per_sm = 0;
  cache[smem] = per_sm;
  return per_sm;
}
template <class T> static size_t orth_smem(int total_cols) {
  // 4 KB granules: the shared-memory attribute and the occupancy query are then made once per granule, not per iteration
  const size_t need = std::max(((size_t)kWarps * total_cols * Num<T>::NC + kWarps) * sizeof(double), std::max<size_t>(16, (size_t)total_cols * sizeof(T)));
  return (need + 4095) / 4096 * 4096;
}
template <class T, int VPT, int CT> static int orth_grid(llz_ctx_t ctx, int total_cols, int64_t n) {
  const int per_sm = orth_ctas_per_sm<T, VPT, CT>(orth_smem<T>(total_cols));
  if (per_sm < 1) return 0;
  const int64_t nslabs = (n + slab_rows<T>(VPT) - 1) / slab_rows<T>(VPT);
  const int grid = persistent_grid(ctx, nslabs, std::min(per_sm, 2));
  return (total_cols * Num<T>::NC + 1 + 31) / 32 <= grid ? grid : 0;  // the reduction needs one CTA per 32 values
}

template <class T, int VPT, int CT>
static int orth_impl(llz_ctx_t ctx, OrthArgs& a, int* fused, int* grid_out) {
  const int grid = orth_grid<T, VPT, CT>(ctx, a.p.ncols, a.p.n);
  if (grid < 1) {
    *fused = 0;
    return LLZ_OK;
  }
  const size_t smem = orth_smem<T>(a.p.ncols);
  a.nred = (a.r.width + 1 + 31) / 32;
  if (!ctx->d_bar) {
    LLZ_CUDA(cudaMalloc(&ctx->d_bar, sizeof(unsigned long long)));
    LLZ_CUDA(cudaMemsetAsync(ctx->d_bar, 0, sizeof(unsigned long long), ctx->stream));
    ctx->bar_count = 0;
  }
  a.r.grid = grid;
  a.s.npb = grid;
  a.bar = ctx->d_bar;
  a.bar_base = ctx->bar_count;
  ctx->bar_count += (a.halo.hp.G > 0 ? 4ull : 3ull) * (unsigned long long)grid;
  void* params[] = {&a};
  cudaError_t le = cudaLaunchCooperativeKernel((const void*)k_orth<T, VPT, CT>, dim3((unsigned)grid), dim3(kThreads), params, smem, ctx->stream);
  if (le != cudaSuccess) return fail(LLZ_ERR_CUDA, "launch k_orth: %s", cudaGetErrorString(le));
  *fused = 1;
  *grid_out = grid;
  return check_launch(ctx, "k_orth");
}

// Small problems (one packet per thread, fewer threads than the GPU holds) keep 16 columns of loads in flight per
// thread instead of 8: there the streaming passes are bound by latency x bytes in flight, not by bandwidth.
template <class T> constexpr int small_ct() { return sizeof(T) <= 8 ? 16 : kCT; }

bool orth_fusable(llz_ctx_t ctx, int dtype, int total_cols, int64_t n) {
  if (total_cols < 1 || total_cols > max_project_cols(dtype)) return false;
  switch (dtype) {
    case LLZ_F32: return (pick_vpt<float>(ctx, n) == 2 ? orth_grid<float, 2, kCT>(ctx, total_cols, n) : orth_grid<float, 1, small_ct<float>()>(ctx, total_cols, n)) > 0;
    case LLZ_F64: return (pick_vpt<double>(ctx, n) == 2 ? orth_grid<double, 2, kCT>(ctx, total_cols, n) : orth_grid<double, 1, small_ct<double>()>(ctx, total_cols, n)) > 0;
    case LLZ_C64: return (pick_vpt<float2>(ctx, n) == 2 ? orth_grid<float2, 2, kCT>(ctx, total_cols, n) : orth_grid<float2, 1, small_ct<float2>()>(ctx, total_cols, n)) > 0;
    case LLZ_C128: return (pick_vpt<double2>(ctx, n) == 2 ? orth_grid<double2, 2, kCT>(ctx, total_cols, n) : orth_grid<double2, 1, small_ct<double2>()>(ctx, total_cols, n)) > 0;
  }
  return false;
}

int launch_orth(llz_ctx_t ctx, int dtype, const ColumnSet& cs, void* w, int64_t n, const Fold& fold, double* ph, double* coef,
                double* wnorm2, const PeerMsg& coef_msg, int wnorm_index, double* norm_partials, const ScalarSink& sink,
                const HaloPushPlan& halo, int* fused, int* grid_out) {
  *fused = 0;
  const int total = cs.ncols();
  if (total < 1 || total > max_project_cols(dtype)) return LLZ_OK;
  OrthArgs a;
  LLZ_TRY(make_project_args(dtype, cs, 0, total, w, n, fold, ph, a.p));
  make_reduce_args(dtype, ph, 0, 0, total, coef, wnorm2, coef_msg, wnorm_index, 1, a.r);
  LLZ_TRY(make_update_args(dtype, cs, 0, total - fold.mode, w, w, n, coef, fold, norm_partials, coef_msg, fold.norm_msg, a.u));
  a.s.x = w;
  a.s.n = n;
  a.s.pb = norm_partials;
  a.s.npb = 0;
  a.s.sink = sink;
  a.halo = halo;
  LLZ_DISPATCH(dtype, {
    if (pick_vpt<T>(ctx, n) == 2) return orth_impl<T, 2, kCT>(ctx, a, fused, grid_out);
    return orth_impl<T, 1, small_ct<T>()>(ctx, a, fused, grid_out);
  });
  return LLZ_OK;
}

int launch_scale_by_norm(llz_ctx_t ctx, int dtype, void* x, int64_t n, const double* norm_partials, int n_partials,
                         const ScalarSink& sink) {
  ScaleNormArgs a;
  a.x = x;
  a.n = n;
  a.pb = norm_partials;
  a.npb = n_partials;
  a.sink = sink;
  LLZ_DISPATCH(dtype, {
    const int64_t npacks = (n + Num<T>::VEC - 1) / Num<T>::VEC;
    const int grid = persistent_grid(ctx, (npacks + kThreads - 1) / kThreads, 4);
    cudaError_t le = launch_chain(ctx, k_scale_norm<T>, grid, kThreads, 0, a);
    if (le != cudaSuccess) return fail(LLZ_ERR_CUDA, "launch k_scale_norm: %s", cudaGetErrorString(le));
  });
  return check_launch(ctx, "k_scale_norm");
}

int launch_recurrence(llz_ctx_t ctx, int dtype, const void* w, const void* u1, const void* u2, void* out, int64_t n,
                      const Fold& fold, double* norm_partials, int* grid_out, const LazyRecurrence* lazy) {
  RecurrenceArgs a;
  a.lazy = lazy ? 1 : 0;
  a.scale1 = lazy ? lazy->scale1 : nullptr;
  a.scale2 = lazy ? lazy->scale2 : nullptr;
  if (lazy) a.sink = lazy->sink;
  a.ticket = lazy ? lazy->ticket : nullptr;
  a.w = w;
  a.u1 = u1;
  a.u2 = u2;
  a.out = out;
  a.n = n;
  a.fold = fold.mode;
  a.pa = fold.alpha_partials;
  a.npa = fold.n_partials;
  a.beta_prev = fold.beta_prev;
  a.alpha_out = fold.alpha_out;
  a.alpha_msg = fold.alpha_msg;
  a.norm_msg = norm_partials ? fold.norm_msg : PeerMsg();
  a.push = (norm_partials && a.norm_msg.ch.G > 0) ? fold.push : GatherPush();
  a.pb = norm_partials;
  LLZ_DISPATCH(dtype, {
    const int64_t npacks = (n + Num<T>::VEC - 1) / Num<T>::VEC;
    const int grid = persistent_grid(ctx, (npacks + kThreads - 1) / kThreads, 4);
    cudaError_t le = launch_chain(ctx, k_recurrence<T>, grid, kThreads, 0, a);
    if (le != cudaSuccess) return fail(LLZ_ERR_CUDA, "launch k_recurrence: %s", cudaGetErrorString(le));
    *grid_out = grid;
  });
  return check_launch(ctx, "k_recurrence");
}

int launch_combine(llz_ctx_t ctx, int dtype, const void* V, int64_t ld, int col0, int ncols, const void* coef,
                   int64_t ldc, int nvec, void* const* out, int64_t n, int accumulate, double* norm_partials,
                   int* grid_out) {
  if (nvec < 1 || nvec > kMaxCombine) return fail(LLZ_ERR_INVALID, "combine: nvec=%d", nvec);
  if (ncols > max_combine_cols(dtype, nvec)) return fail(LLZ_ERR_INVALID, "combine: %d columns per launch", ncols);
  CombineArgs a;
  a.V = V;
  a.ld = ld;
  a.col0 = col0;
  a.ncols = ncols;
  a.coef = coef;
  a.ldc = ldc;
  a.nvec = nvec;
  for (int r = 0; r < kMaxCombine; ++r) a.out[r] = r < nvec ? out[r] : nullptr;
  a.n = n;
  a.accumulate = accumulate;
  a.pb = norm_partials;
  LLZ_DISPATCH(dtype, {
    const int64_t nslabs = (n + slab_rows<T>(1) - 1) / slab_rows<T>(1);
    const int grid = persistent_grid(ctx, nslabs, 2);
    if (nvec == 1) {
      const size_t smem = std::max<size_t>(16, (size_t)ncols * sizeof(T));
      LLZ_TRY(set_smem(k_combine<T, 1>, smem));
      k_combine<T, 1><<<grid, kThreads, smem, ctx->stream>>>(a);
    } else {
      const size_t smem = std::max<size_t>(16, (size_t)ncols * sizeof(T) * kMaxCombine);
      LLZ_TRY(set_smem(k_combine<T, kMaxCombine>, smem));
      k_combine<T, kMaxCombine><<<grid, kThreads, smem, ctx->stream>>>(a);
    }
    *grid_out = grid;
  });
  return check_launch(ctx, "k_combine");
}

int launch_dot(llz_ctx_t ctx, int dtype, const void* a, const void* b, int64_t n, double* partials, int* grid_out) {
  LLZ_DISPATCH(dtype, {
    const int64_t npacks = (n + Num<T>::VEC - 1) / Num<T>::VEC;
    const int grid = persistent_grid(ctx, (npacks + kThreads - 1) / kThreads, 4);
    k_dot<T><<<grid, kThreads, 0, ctx->stream>>>((const T*)a, (const T*)b, n, partials);
    *grid_out = grid;
  });
  return check_launch(ctx, "k_dot");
}

int launch_asum(llz_ctx_t ctx, int dtype, const void* a, int64_t n, double* partials, int* grid_out) {
  LLZ_DISPATCH(dtype, {
    const int64_t npacks = (n + Num<T>::VEC - 1) / Num<T>::VEC;
    const int grid = persistent_grid(ctx, (npacks + kThreads - 1) / kThreads, 4);
    k_asum<T><<<grid, kThreads, 0, ctx->stream>>>((const T*)a, n, partials);
    *grid_out = grid;
  });
  return check_launch(ctx, "k_asum");
}

int launch_redot(llz_ctx_t ctx, int dtype, const void* a, const void* b, int64_t n, double* partials, int* grid_out,
                 const PeerMsg& msg) {
  LLZ_DISPATCH(dtype, {
    const int64_t npacks = (n + Num<T>::VEC - 1) / Num<T>::VEC;
    const int grid = persistent_grid(ctx, (npacks + kThreads - 1) / kThreads, 4);
    cudaError_t le = launch_chain(ctx, k_redot<T>, grid, kThreads, 0, (const T*)a, (const T*)b, n, partials, msg);
    if (le != cudaSuccess) return fail(LLZ_ERR_CUDA, "launch k_redot: %s", cudaGetErrorString(le));
    *grid_out = grid;
  });
  return check_launch(ctx, "k_redot");
}

int launch_sum_partials(llz_ctx_t ctx, const double* partials, int count, int nc, double* result, double* h_result) {
  k_sum_partials<<<1, kThreads, 0, ctx->stream>>>(partials, count, nc, result, h_result);
  return check_launch(ctx, "k_sum_partials");
}

int launch_scale(llz_ctx_t ctx, int dtype, void* x, int64_t n, const double a[2]) {
  LLZ_DISPATCH(dtype, {
    const int grid = persistent_grid(ctx, (n + kThreads - 1) / kThreads, 8);
    k_scale<T><<<grid, kThreads, 0, ctx->stream>>>((T*)x, n, from_double_host<T>(a));
  });
  return check_launch(ctx, "k_scale");
}

int launch_axpy(llz_ctx_t ctx, int dtype, void* y, const double a[2], const void* x, int64_t n) {
  LLZ_DISPATCH(dtype, {
    const int grid = persistent_grid(ctx, (n + kThreads - 1) / kThreads, 8);
    k_axpy<T><<<grid, kThreads, 0, ctx->stream>>>((T*)y, from_double_host<T>(a), (const T*)x, n);
  });
  return check_launch(ctx, "k_axpy");
}

}  // namespace llz
