// llz_xxz.cu — matrix-free spin-1/2 XXZ chain operator in a fixed-Sz sector (BASELINE.json configs 4 and 5):
//     H = sum_<i,i+1> [ Jxy/2 (S+_i S-_{i+1} + h.c.) + Jz Sz_i Sz_{i+1} ]
// on L sites, basis = all L-bit integers with n_up set bits in increasing order.  Nothing of the matrix is stored:
// y_r = diag(s_r) x_r + Jxy/2 * sum_{anti-parallel bonds b of s_r} x[rank(s_r ^ flip_b)]   (H is symmetric, so the
// gather form needs no atomics).  The basis states come from a 4-byte-per-row table built on the device (one
// unranking per 32 states, Gosper steps between); the rank of a flipped state is the row index plus or minus a
// binomial coefficient (k_xxz_apply), the two Lin tables (low / high half of the bit string, 2^(L/2) entries each)
// only serve the periodic wrap bond.  Row-sharded, the block reads index ranges of x owned by other ranks: they arrive
// either pushed by the kernel that produced x (fused all-gather, llz_peer.cuh) or by grouped send/recv.
// The alpha = Re<x, Hx> dot is fused into the epilogue like the CSR kernel's.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "llz_device.cuh"
#include "llz_launch.hpp"

namespace llz {

__constant__ unsigned long long c_binom[34][34];

struct XxzParams {
  int L, n_up, half;   // half = number of low bits covered by the low table
  int periodic;
  double jz4, jxy2;    // Jz/4, Jxy/2
  const uint32_t* rank_lo;  // [2^half]
  const uint32_t* rank_hi;  // [2^(L-half)]
  int64_t n;
  int64_t row0;        // first local row (row-sharded runs)
  const void* x_all;   // row-sharded runs: the whole input vector, gathered before the launch (entries of the local
                       // block are read from x itself); null for a single rank
  const double* x_scale;  // fused all-gather: the remote entries in x_all are still un-normalised, multiply them by
                          // 1 / *x_scale (the norm the producer's successor kernel divided the local block by)
};

__device__ __forceinline__ uint32_t xxz_unrank(int64_t idx, int L, int n_up) {
  uint32_t s = 0;
  int p = L - 1;
  unsigned long long rem = (unsigned long long)idx;
  for (int k = n_up; k >= 1; --k) {
    while (c_binom[p][k] > rem) --p;
    s |= (1u << p);
    rem -= c_binom[p][k];
    --p;
  }
  return s;
}

__device__ __forceinline__ uint32_t gosper_next(uint32_t s) {
  const uint32_t t = s | (s - 1);
  return (t + 1) | (((~t & -~t) - 1) >> __ffs(s));
}

// states[r] = the r-th basis state of the local block (one unranking per 32 consecutive states, Gosper steps between).
__global__ void __launch_bounds__(kThreads) k_xxz_states(uint32_t* __restrict__ states, int64_t n, int64_t row0, int L, int n_up) {
  constexpr int CH = 32;
  const int64_t nchunks = (n + CH - 1) / CH;
  for (int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x; c < nchunks; c += (int64_t)gridDim.x * kThreads) {
    const int64_t first = c * CH;
    uint32_t s = xxz_unrank(row0 + first, L, n_up);
    const int cnt = (int)((n - first < CH) ? (n - first) : CH);
    for (int i = 0; i < cnt; ++i) {
      states[first + i] = s;
      s = gosper_next(s);
    }
  }
}

// One thread per basis state, consecutive threads on consecutive states, and a loop over the BONDS that is uniform
// across the warp.  The rank of the flipped state needs no table: in the combinatorial number system
// rank(s) = sum_j C(p_j, j) (p_j = position of the j-th set bit), exchanging the adjacent bits (b, b+1) moves one set
// bit by one place and changes the rank by exactly
//        +C(b, c)  if bit b was the set one,   -C(b, c)  if bit b+1 was,      c = popcount(s & ((1 << b) - 1)),
// a shared-memory lookup in Pascal's triangle.  For a given bond the 32 lanes hold 32 consecutive states, which share
// all but their lowest bits: on high bonds c — hence the offset — is the same in every lane and the x gather is one
// contiguous 256-byte read; on low bonds the offsets are a few hundred elements at most and hit L1.  Only the periodic
// wrap bond (bits L-1 and 0) re-ranks through the two Lin tables.  The state itself comes from a 4-byte-per-row table
// (A_bytes = 4 n).
template <class T, bool SHARDED>
__global__ void __launch_bounds__(kThreads, 6)
    k_xxz_apply(const T* __restrict__ x, T* __restrict__ y, const uint32_t* __restrict__ states, XxzParams p,
                typename Num<T>::R sigma, double* pa, PeerMsg msg, PeerMsg gather_msg) {
  using R = typename Num<T>::R;
  __shared__ double scratch[kWarps];
  __shared__ uint32_t pascal[32 * 33];  // pascal[b*33 + c] = C(b, c); the odd stride spreads rows over the banks
  for (int i = threadIdx.x; i < 32 * 32; i += kThreads) pascal[(i >> 5) * 33 + (i & 31)] = (uint32_t)c_binom[i >> 5][i & 31];
  pdl_prologue();  // (the table above depends on nothing an earlier kernel wrote, so it is built while that one drains)
  if (SHARDED && gather_msg.ch.G > 0) peer_wait(gather_msg.ch, gather_msg.seq);  // the peers' blocks have landed in x_all
  R inv = (R)1;
  const bool rescale = SHARDED && p.x_scale != nullptr;
  if (rescale) inv = (R)1 / (R)(*p.x_scale);  // exactly the factor k_scale_norm applied to the owner's copy
  __syncthreads();
  const uint32_t lo_mask = (1u << p.half) - 1u;
  const int nbonds = p.periodic ? p.L : p.L - 1;
  const int inner = p.L - 1;  // bonds (b, b+1) that stay inside the bit string
  const uint32_t inner_mask = (1u << (p.L - 1)) - 1u;
  const uint32_t top = 1u << (p.L - 1);
  const T* __restrict__ xg = reinterpret_cast<const T*>(p.x_all);
  // every index fits 32 bits: the largest sector (L = 32, 16 up spins) has 601 080 390 states
  const int32_t n = (int32_t)p.n, row0 = (int32_t)p.row0;
  const int32_t stride = (int32_t)(gridDim.x * kThreads);
  double dot = 0.0;
  for (int32_t r = (int32_t)(blockIdx.x * kThreads + threadIdx.x); r < n; r += stride) {
    const uint32_t s = __ldg(states + r);
    // anti-parallel bonds: bit b set <=> sites b and b+1 differ
    const uint32_t d = (s ^ (s >> 1)) & inner_mask;
    T acc = zero_of(T());
    // walk the bonds with the state and the bond mask shifted down one place per step; `pi` tracks the address of
    // pascal[b][c], c = number of set bits below b (next row: +33, one more set bit below: +1)
    uint32_t sc = s, dc = d;
    const uint32_t* pi = pascal;
#pragma unroll 4
    for (int b = 0; b < inner; ++b) {
      const uint32_t up = sc & 1u;
      if (dc & 1u) {
        const int32_t delta = (int32_t)*pi;
        const int32_t j = up ? r + delta : r - delta;  // local index of the flipped state (may leave the block)
        T xv;
        if (SHARDED) {
          if ((uint32_t)j < (uint32_t)n) {
            xv = __ldg(x + j);
          } else {
            xv = __ldcg(xg + (j + row0));  // written by peers: not through the non-coherent path
            if (rescale) xv = scale_real(xv, inv);
          }
        } else {
          xv = __ldg(x + j);
        }
        acc = add_t(acc, xv);
      }
      pi += 33 + up;
      sc >>= 1;
      dc >>= 1;
    }
    int anti = __popc(d);
    if (p.periodic && (((s >> (p.L - 1)) ^ s) & 1u)) {  // wrap bond (bits L-1 and 0): re-rank through the Lin tables
      ++anti;
      const uint32_t t = s ^ (top | 1u);
      const int32_t jg = (int32_t)(__ldg(p.rank_lo + (t & lo_mask)) + __ldg(p.rank_hi + (t >> p.half)));
      const int32_t j = jg - row0;
      T xv;
      if (SHARDED) {
        if ((uint32_t)j < (uint32_t)n) {
          xv = __ldg(x + j);
        } else {
          xv = __ldcg(xg + jg);
          if (rescale) xv = scale_real(xv, inv);
        }
      } else {
        xv = __ldg(x + j);
      }
      acc = add_t(acc, xv);
    }
    const R diag = (R)(p.jz4 * (double)(nbonds - 2 * anti));
    const T xi = x[r];
    T yi = scale_real(acc, (R)p.jxy2);
    yi = add_t(yi, scale_real(xi, diag + sigma));
    y[r] = yi;
    dot += re_conj_mul(xi, yi);
  }
  const double t = block_sum(dot, scratch);
  finish_scalar(t, pa, msg, scratch);
}

// Which entries of the input vector does this row block read from every other rank?  Per owner the smallest and the
// largest global index (same target arithmetic as k_xxz_apply), so that only those ranges travel: with 8 ranks a block
// reads 13-38 % of the vector, not the 87.5 % a plain all-gather moves.
struct XxzNeed {
  long long lo[kMaxRanks];
  long long hi[kMaxRanks];
};
__global__ void __launch_bounds__(kThreads) k_xxz_need_ranges(const uint32_t* __restrict__ states, XxzParams p, int G, XxzNeed bounds /* lo = block starts */,
                                                                unsigned long long* out_lo, unsigned long long* out_hi) {
  __shared__ unsigned long long s_lo[kMaxRanks], s_hi[kMaxRanks];
  __shared__ uint32_t pascal[32 * 33];
  for (int i = threadIdx.x; i < 32 * 32; i += kThreads) pascal[(i >> 5) * 33 + (i & 31)] = (uint32_t)c_binom[i >> 5][i & 31];
  if (threadIdx.x < kMaxRanks) {
    s_lo[threadIdx.x] = ~0ull;
    s_hi[threadIdx.x] = 0ull;
  }
  __syncthreads();
  const uint32_t lo_mask = (1u << p.half) - 1u;
  const int inner = p.L - 1;
  const uint32_t inner_mask = (1u << (p.L - 1)) - 1u;
  auto note = [&](long long jg) {
    if (jg >= p.row0 && jg < p.row0 + p.n) return;
    int o = 0;
    while (o + 1 < G && jg >= bounds.lo[o + 1]) ++o;
    atomicMin(&s_lo[o], (unsigned long long)jg);
    atomicMax(&s_hi[o], (unsigned long long)jg + 1ull);
  };
  for (long long r = (long long)blockIdx.x * kThreads + threadIdx.x; r < p.n; r += (long long)gridDim.x * kThreads) {
    const uint32_t s = states[r];
    uint32_t sc = s, dc = (s ^ (s >> 1)) & inner_mask;
    const uint32_t* pi = pascal;
    for (int b = 0; b < inner; ++b) {
      const uint32_t up = sc & 1u;
      if (dc & 1u) note(p.row0 + (up ? r + (long long)*pi : r - (long long)*pi));
      pi += 33 + up;
      sc >>= 1;
      dc >>= 1;
    }
    if (p.periodic && (((s >> (p.L - 1)) ^ s) & 1u)) {
      const uint32_t t = s ^ ((1u << (p.L - 1)) | 1u);
      note((long long)p.rank_lo[t & lo_mask] + (long long)p.rank_hi[t >> p.half]);
    }
  }
  __syncthreads();
  if ((int)threadIdx.x < G) {
    if (s_lo[threadIdx.x] != ~0ull) atomicMin(out_lo + threadIdx.x, s_lo[threadIdx.x]);
    if (s_hi[threadIdx.x] != 0ull) atomicMax(out_hi + threadIdx.x, s_hi[threadIdx.x]);
  }
}

struct XxzOpBase : OpBase {
  XxzParams prm;
  uint32_t* d_lo = nullptr;
  uint32_t* d_hi = nullptr;
  uint32_t* d_states = nullptr;  // basis states of the local block
  void* d_xall = nullptr;  // row-sharded runs: gathered input vector (n_global elements), NCCL path
  std::vector<size_t> send_off, send_bytes, recv_off, recv_bytes;
  // fused all-gather: a double-buffered (by message parity) whole-vector buffer on every rank, mapped into every peer
  ExchangeBuffer* xb = nullptr;
  long long push_lo[kMaxRanks] = {}, push_hi[kMaxRanks] = {};  // local element range each peer reads from this block
  PeerMsg cur_gather_msg;  // what the next apply waits for (none: x_all was filled in stream order)
  ~XxzOpBase() override {
    if (d_lo) dev_free(ctx, d_lo);
    if (d_hi) dev_free(ctx, d_hi);
    if (d_xall) dev_free(ctx, d_xall);
    if (d_states) dev_free(ctx, d_states);
    if (xb) comm_exchange_buffer_release(ctx, xb);
  }
  size_t vec_bytes() const { return ((size_t)n_global * dtype_size(dtype) + 255) / 256 * 256; }  // stride between the two copies
  bool plan_push(GatherPush* push) override {
    if (!xb) return false;
    push->msg = comm_next_message(ctx, kChanGather);
    push->G = ctx->nranks;
    push->rank = ctx->rank;
    const size_t off = ((push->msg.seq & 1ull) ? vec_bytes() : 0) + (size_t)row0 * dtype_size(dtype);
    for (int r = 0; r < ctx->nranks; ++r) {
      push->dst[r] = static_cast<char*>(xb->peer[r]) + off;
      push->lo[r] = push_lo[r];
      push->hi[r] = push_hi[r];
    }
    return true;
  }
  void use_pushed(const GatherPush& push, const double* scale) override {
    prm.x_all = static_cast<char*>(xb->local) + ((push.msg.seq & 1ull) ? vec_bytes() : 0);
    prm.x_scale = scale;
    cur_gather_msg = push.msg;
  }
  // every row has at most `nbonds` off-diagonal entries Jxy/2 and a diagonal of magnitude <= nbonds |Jz|/4
  int abs_row_sum_max(double* out) override {
    const int nbonds = prm.periodic ? prm.L : prm.L - 1;
    *out = nbonds * (fabs(prm.jz4) + fabs(prm.jxy2));
    return LLZ_OK;
  }
  // Row-sharded: every rank sends its block of x to every peer (bit flips on high sites land anywhere in the sector).
  int prepare(const void* x) override {
    if (ctx->nranks == 1) return LLZ_OK;
    ProfScope ps(ctx, "halo", (double)(n_global - n_local) * (double)dtype_size(dtype));
    char* dst = (char*)d_xall;
    if (xb) {  // keep the parity of the exchange buffers alternating with the fused messages
      const PeerMsg m = comm_next_message(ctx, kChanGather);
      dst = static_cast<char*>(xb->local) + ((m.seq & 1ull) ? vec_bytes() : 0);
    }
    prm.x_all = dst;
    prm.x_scale = nullptr;
    cur_gather_msg = PeerMsg();
    return comm_exchange(ctx, (const char*)x, send_off.data(), send_bytes.data(), dst, recv_off.data(), recv_bytes.data());
  }
};

template <class T> struct XxzOp : XxzOpBase {
  int apply_fused(const void* x, void* y, double sigma, double* pa, int* npa, const PeerMsg* alpha_msg) override {
    // persistent: exactly the 6 CTAs per SM that __launch_bounds__(kThreads, 6) keeps resident
    int64_t g = std::min<int64_t>((n_local + kThreads - 1) / kThreads, std::min<int64_t>(kMaxGrid, (int64_t)ctx->num_sms * 6));
    if (g < 1) g = 1;
    const PeerMsg msg = alpha_msg ? *alpha_msg : PeerMsg();
    cudaError_t e;
    if (prm.x_all)
      e = launch_chain(ctx, k_xxz_apply<T, true>, (int)g, kThreads, 0, (const T*)x, (T*)y, (const uint32_t*)d_states, prm,
                       (typename Num<T>::R)sigma, pa, msg, cur_gather_msg);
    else
      e = launch_chain(ctx, k_xxz_apply<T, false>, (int)g, kThreads, 0, (const T*)x, (T*)y, (const uint32_t*)d_states, prm,
                       (typename Num<T>::R)sigma, pa, msg, PeerMsg());
    *npa = (int)g;
    if (e != cudaSuccess) return fail(LLZ_ERR_CUDA, "launch k_xxz_apply: %s", cudaGetErrorString(e));
    ctx->launches++;
    return LLZ_OK;
  }
};

static unsigned long long host_binom[34][34];
static bool binom_ready = false;
static void init_binom() {
  if (binom_ready) return;
  for (int n = 0; n < 34; ++n) {
    for (int k = 0; k < 34; ++k) host_binom[n][k] = 0;
    host_binom[n][0] = 1;
    for (int k = 1; k <= n; ++k) host_binom[n][k] = host_binom[n - 1][k - 1] + (k <= n - 1 ? host_binom[n - 1][k] : 0);
  }
  binom_ready = true;
}

}  // namespace llz

using namespace llz;

extern "C" int llz_op_create_xxz(llz_ctx_t ctx, int dtype, int L, int n_up, double jz, double jxy, int periodic, llz_op_t* out) {
  if (!ctx || !out || L < 2 || L > 32 || n_up < 0 || n_up > L || dtype_size(dtype) == 0)
    return fail(LLZ_ERR_INVALID, "op_create_xxz: need 2 <= L <= 32, 0 <= n_up <= L");
  LLZ_CUDA(cudaSetDevice(ctx->device));
  init_binom();
  LLZ_CUDA(cudaMemcpyToSymbol(c_binom, host_binom, sizeof(host_binom)));
  XxzOpBase* op = nullptr;
  switch (dtype) {
    case LLZ_F32: op = new XxzOp<float>(); break;
    case LLZ_F64: op = new XxzOp<double>(); break;
    case LLZ_C64: op = new XxzOp<float2>(); break;
    case LLZ_C128: op = new XxzOp<double2>(); break;
  }
  op->ctx = ctx;
  op->dtype = dtype;
  op->n_global = (int64_t)host_binom[L][n_up];
  op->bytes = 0;
  {
    const int s0 = llz_partition(op->n_global, ctx->rank, ctx->nranks, &op->row0, &op->n_local);
    if (s0 != LLZ_OK || op->n_local < 1) {
      delete op;
      return s0 != LLZ_OK ? s0 : fail(LLZ_ERR_INVALID, "op_create_xxz: sector of %lld states is too small for %d ranks", (long long)host_binom[L][n_up], ctx->nranks);
    }
  }
  const int half = L / 2;
  // Lin tables: rank(s) = lo[s & lo_mask] + hi[s >> half].  The j-th set bit (1-based, counted from bit 0) at
  // position q contributes C(q, j); for the high half j continues from the number of set bits in the low half,
  // which the fixed magnetisation determines: popcount(low) = n_up - popcount(high).
  std::vector<uint32_t> lo((size_t)1 << half), hi((size_t)1 << (L - half));
  for (uint32_t b = 0; b < lo.size(); ++b) {
    unsigned long long r = 0;
    int j = 0;
    for (int q = 0; q < half; ++q)
      if (b >> q & 1u) r += host_binom[q][++j];
    lo[b] = (uint32_t)r;
  }
  for (uint32_t b = 0; b < hi.size(); ++b) {
    const int pc = __builtin_popcount(b);
    unsigned long long r = 0;
    if (pc <= n_up) {
      int j = n_up - pc;
      for (int q = 0; q < L - half; ++q)
        if (b >> q & 1u) r += host_binom[q + half][++j];
    }
    hi[b] = (uint32_t)r;
  }
  cudaError_t e = dev_malloc(ctx, &op->d_lo, lo.size() * sizeof(uint32_t));
  if (e == cudaSuccess) e = dev_malloc(ctx, &op->d_hi, hi.size() * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMemcpy(op->d_lo, lo.data(), lo.size() * sizeof(uint32_t), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(op->d_hi, hi.data(), hi.size() * sizeof(uint32_t), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    delete op;
    return fail(LLZ_ERR_CUDA, "op_create_xxz: %s", cudaGetErrorString(e));
  }
  op->prm.L = L;
  op->prm.n_up = n_up;
  op->prm.half = half;
  op->prm.periodic = periodic ? 1 : 0;
  op->prm.jz4 = jz * 0.25;
  op->prm.jxy2 = jxy * 0.5;
  op->prm.rank_lo = op->d_lo;
  op->prm.rank_hi = op->d_hi;
  op->prm.n = op->n_local;
  op->prm.row0 = op->row0;
  op->prm.x_all = nullptr;
  op->prm.x_scale = nullptr;
  op->bytes = op->n_local * 4;  // the state table is the only stored part of the operator
  e = dev_malloc(ctx, &op->d_states, sizeof(uint32_t) * (size_t)op->n_local);
  if (e != cudaSuccess) {
    delete op;
    return fail(LLZ_ERR_OOM, "op_create_xxz: state table (%lld entries): %s", (long long)op->n_local, cudaGetErrorString(e));
  }
  {
    const int64_t nchunks = (op->n_local + 31) / 32;
    const int grid = (int)std::min<int64_t>((nchunks + kThreads - 1) / kThreads, (int64_t)ctx->num_sms * 8);
    k_xxz_states<<<std::max(grid, 1), kThreads, 0, ctx->stream>>>(op->d_states, op->n_local, op->row0, L, n_up);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      delete op;
      return fail(LLZ_ERR_CUDA, "op_create_xxz: state table kernel: %s", cudaGetErrorString(e));
    }
    ctx->launches++;
  }
  if (ctx->nranks > 1) {
    const size_t es = dtype_size(dtype);
    // fused all-gather buffers (two whole vectors, IPC-mapped group-wide) when the group has peer channels; the NCCL
    // gather then lands in them too.  Otherwise a private gather buffer.
    const char* env = getenv("LLZ_FUSED_GATHER");
    if (!(env && env[0] == '0')) op->xb = comm_exchange_buffer_acquire(ctx, 2 * op->vec_bytes());
    if (!op->xb) {
      e = dev_malloc(ctx, &op->d_xall, (size_t)op->n_global * es);
      if (e != cudaSuccess) {
        delete op;
        return fail(LLZ_ERR_OOM, "op_create_xxz: gathered input vector (%lld elements): %s", (long long)op->n_global, cudaGetErrorString(e));
      }
    }
    op->prm.x_all = op->xb ? op->xb->local : op->d_xall;
    const int G = ctx->nranks;
    // what this block reads from every other rank: [need_lo, need_hi) per owner, widened to multiples of 4 elements
    XxzNeed bounds;
    std::vector<int64_t> starts((size_t)G + 1, op->n_global);
    for (int r = 0; r < G; ++r) {
      int64_t nl = 0;
      llz_partition(op->n_global, r, G, &starts[(size_t)r], &nl);
      bounds.lo[r] = starts[(size_t)r];
      bounds.hi[r] = starts[(size_t)r] + nl;
    }
    unsigned long long* d_rng = nullptr;
    std::vector<unsigned long long> h_rng((size_t)2 * kMaxRanks);
    for (int r = 0; r < kMaxRanks; ++r) {
      h_rng[(size_t)r] = ~0ull;
      h_rng[(size_t)kMaxRanks + r] = 0ull;
    }
    e = dev_malloc(ctx, &d_rng, sizeof(unsigned long long) * h_rng.size());
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_rng, h_rng.data(), sizeof(unsigned long long) * h_rng.size(), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
      const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((op->n_local + kThreads - 1) / kThreads, (int64_t)ctx->num_sms * 8));
      k_xxz_need_ranges<<<grid, kThreads, 0, ctx->stream>>>(op->d_states, op->prm, G, bounds, d_rng, d_rng + kMaxRanks);
      e = cudaGetLastError();
      ctx->launches++;
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_rng.data(), d_rng, sizeof(unsigned long long) * h_rng.size(), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (d_rng) dev_free(ctx, d_rng);
    if (e != cudaSuccess) {
      delete op;
      return fail(LLZ_ERR_CUDA, "op_create_xxz: halo range kernel: %s", cudaGetErrorString(e));
    }
    std::vector<int64_t> need((size_t)2 * G, 0), need_all((size_t)2 * G * G, 0);  // [lo, hi) per owner, global indices
    for (int r = 0; r < G; ++r) {
      if (r == ctx->rank || h_rng[(size_t)r] == ~0ull) continue;  // nothing read from r
      int64_t lo = (int64_t)h_rng[(size_t)r] & ~(int64_t)3, hi = ((int64_t)h_rng[(size_t)kMaxRanks + r] + 3) & ~(int64_t)3;
      lo = std::max<int64_t>(lo, (int64_t)bounds.lo[r]);  // block starts are multiples of 4
      hi = std::min<int64_t>(hi, (int64_t)bounds.hi[r]);
      need[(size_t)2 * r] = lo;
      need[(size_t)2 * r + 1] = hi;
    }
    int st = comm_allgather_host(ctx, need.data(), need_all.data(), sizeof(int64_t) * 2 * G);
    if (st != LLZ_OK) {
      delete op;
      return st;
    }
    op->send_off.assign(G, 0);
    op->send_bytes.assign(G, 0);
    op->recv_off.assign(G, 0);
    op->recv_bytes.assign(G, 0);
    for (int r = 0; r < G; ++r) {
      if (r == ctx->rank) continue;
      // what I read from r lands at its global position in my gathered vector
      op->recv_off[r] = (size_t)need[(size_t)2 * r] * es;
      op->recv_bytes[r] = (size_t)(need[(size_t)2 * r + 1] - need[(size_t)2 * r]) * es;
      // what r reads from me: a range of my block (local element indices for the pushing kernels)
      const int64_t lo = need_all[((size_t)r * G + ctx->rank) * 2], hi = need_all[((size_t)r * G + ctx->rank) * 2 + 1];
      op->push_lo[r] = lo - op->row0;
      op->push_hi[r] = hi - op->row0;
      op->send_off[r] = (size_t)(lo - op->row0) * es;
      op->send_bytes[r] = (size_t)(hi - lo) * es;
    }
  }
  llz_op_t h = new llz_op_s();
  h->impl = op;
  *out = h;
  return LLZ_OK;
}
