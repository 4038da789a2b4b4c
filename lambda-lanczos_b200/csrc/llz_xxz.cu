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
#include <type_traits>
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

// y_r = Jxy/2 * (sum of the flipped neighbours) + (diag + sigma) * x_r with a fixed rounding sequence (product, then one
// fused multiply-add), so that the per-state and the block kernel — and every instantiation of them — agree bit for bit.
__device__ __forceinline__ float xxz_row(float acc, float xi, float jxy2, float dg) { return fmaf(xi, dg, __fmul_rn(acc, jxy2)); }
__device__ __forceinline__ double xxz_row(double acc, double xi, double jxy2, double dg) { return fma(xi, dg, __dmul_rn(acc, jxy2)); }
__device__ __forceinline__ float2 xxz_row(float2 acc, float2 xi, float jxy2, float dg) {
  return make_float2(xxz_row(acc.x, xi.x, jxy2, dg), xxz_row(acc.y, xi.y, jxy2, dg));
}
__device__ __forceinline__ double2 xxz_row(double2 acc, double2 xi, double jxy2, double dg) {
  return make_double2(xxz_row(acc.x, xi.x, jxy2, dg), xxz_row(acc.y, xi.y, jxy2, dg));
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
    const T yi = xxz_row(acc, xi, (R)p.jxy2, diag + sigma);
    y[r] = yi;
    dot += re_conj_mul(xi, yi);
  }
  const double t = block_sum(dot, scratch);
  finish_scalar(t, pa, msg, scratch);
}


// ------------------------------------------------------------------------------------------------------------------
// Block kernel (the default).  Split a basis state into its m low bits and its high bits, s = (h << m) | l.  The states
// that share h are CONSECUTIVE rows — a block of C(m, p) rows, p = n_up - popcount(h), starting at row base(h) — and
// inside a block the row index is the rank of l among the m-bit strings of popcount p, whatever h is.  Hence
//   * a bond inside the low bits maps a block onto itself, and the row it leads to depends on (p, l) only: the CTA
//     stages the block's x segment in shared memory (one contiguous read) and follows per-(p, l) neighbour lists it
//     built once per popcount class — 13 of the ~28 bonds never leave shared memory;
//   * a bond inside the high bits maps the whole block onto another block with the same l's: x[base(h') + i], a
//     contiguous, perfectly coalesced stream whose offset is uniform over the CTA;
//   * only the bond across the split (bits m-1 | m) and the periodic wrap bond (bits L-1 | 0) gather individually.
// Nothing per state is stored: no state table, no rank table — A_bytes is a few hundred KB of popcount-class tables.
// CTAs take the blocks in the order of a host-built list (grouped by popcount class, so the neighbour lists are
// rebuilt ~m times per launch, and so that the blocks in flight are each other's high-bond neighbours in L2).
// Products are added in bond order 0, 1, ..., L-2, wrap — the order of the plain per-state loop (k_xxz_apply), so
// both kernels produce the same bits, and a row-sharded run the same bits as a single GPU.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kXxzMaxLow = 13;       // blocks of at most C(13,6) = 1716 rows: <= 7 rows per thread of a 256-thread CTA

struct XxzBlockParams {
  int L, n_up, m, periodic;
  double jz4, jxy2;
  const uint16_t* t_lo;   // [2^m]     row of a low configuration inside its block
  const uint16_t* t_ls;   // [2^m]     low configurations grouped by popcount: t_ls[ls_off[p] + i]
  const uint4* blocks;    // [nblocks] (h, base(h), base(h ^ top bit), 0) of the blocks that intersect the local rows
  int nblocks;
  int bsmax;              // rows of the largest block
  int ls_off[kXxzMaxLow + 2];
  int32_t n, row0;        // local rows [row0, row0 + n)
  const void* x_all;      // row-sharded: the whole input vector (see XxzParams)
  const double* x_scale;
};

struct XxzBlockDesc {
  int32_t base, bs, i0, i1;  // first row, rows, local row range [i0, i1) of the block
  int32_t p;                 // popcount of the low part
  int32_t nhigh;             // anti-parallel bonds inside the high bits
  int32_t h0;                // bit 0 of h, or -1 when there is no bond across the split
  int32_t top;               // bit L-1 when it lies in h (else -1: read it from l), for the wrap bond
  int32_t wrap_base;         // base(h ^ top bit)
  int32_t mode[32];          // per high bond: 0 = the neighbour block is local, 1 = remote (gathered vector), 2 = both
  const void* src[32];       // per high bond (ascending bond order): address of row 0 of the neighbour block's x segment
};

template <class T, bool SHARDED>
__device__ __forceinline__ T xxz_load(const T* __restrict__ x, const T* xg, int32_t g, int32_t row0, int32_t n, bool rescale,
                                      typename Num<T>::R inv) {
  if (!SHARDED) return __ldg(x + g);
  const int32_t j = g - row0;
  if ((uint32_t)j < (uint32_t)n) return __ldg(x + j);
  T v = __ldcg(xg + g);  // written by peers while this kernel may be resident: not through the non-coherent path
  return rescale ? scale_real(v, inv) : v;
}

// cp.async of one element (4, 8 or 16 bytes) into shared memory
template <int BYTES> __device__ __forceinline__ void cp_async_elem(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  if (BYTES == 16)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(gsrc), "n"(BYTES) : "memory");
}

// resident CTAs the register allocation must allow: 4 of 256 threads (64 registers) unless the loads a thread keeps in
// flight alone need more
template <class T, int NT, int R, int D> constexpr int xxz_min_blocks() {
  return (int)sizeof(T) * R * D <= 64 ? 1024 / NT : 512 / NT;
}

template <class T, bool SHARDED, int NT, int R, int D>
__global__ void __launch_bounds__(NT, xxz_min_blocks<T, NT, R, D>())
    k_xxz_block_apply(const T* __restrict__ x, T* __restrict__ y, XxzBlockParams p, typename Num<T>::R sigma, double* pa,
                      PeerMsg msg, PeerMsg gather_msg) {
  using RT = typename Num<T>::R;
  constexpr int NW = NT / 32;
  extern __shared__ __align__(16) unsigned char smem_x[];
  __shared__ double scratch[NW];
  __shared__ uint32_t pascal[32 * 33];
  __shared__ XxzBlockDesc desc[3];  // block it % 3: written one block ahead, while the slowest warp may still read it - 1
  const int tid = threadIdx.x;
  for (int i = tid; i < 32 * 32; i += NT) pascal[(i >> 5) * 33 + (i & 31)] = (uint32_t)c_binom[i >> 5][i & 31];
  // shared memory: two x tiles of bsmax + 1 elements (the last one is the zero the padding entries of the neighbour
  // lists point at) | neighbour offsets [bsmax][E] (E = m - 1 rounded up to 4: a row's list is E/4 8-byte loads) |
  // meta [bsmax] | low states [bsmax]
  const int bsmax = p.bsmax;
  const int m = p.m, L = p.L;
  const int E4 = (m - 1 + 3) / 4;  // 8-byte words of neighbour offsets per row
  const size_t tile_bytes = (((size_t)bsmax + 1) * sizeof(T) + 15) / 16 * 16;
  uint2* nb = reinterpret_cast<uint2*>(smem_x + 2 * tile_bytes);
  uint16_t* mt = reinterpret_cast<uint16_t*>(nb + (size_t)E4 * bsmax);
  uint16_t* ls = mt + bsmax;
  if (tid < 2) reinterpret_cast<T*>(smem_x + tid * tile_bytes)[bsmax] = zero_of(T());
  if (SHARDED && gather_msg.ch.G > 0) peer_wait(gather_msg.ch, gather_msg.seq);
  RT inv = (RT)1;
  const bool rescale = SHARDED && p.x_scale != nullptr;
  if (rescale) inv = (RT)1 / (RT)(*p.x_scale);
  const T* xg = reinterpret_cast<const T*>(p.x_all);
  const int nhb = L - m - 1;  // bonds inside the high bits
  const int nbonds = p.periodic ? L : L - 1;
  const bool wrap_in_low = L - 1 < m;
  const uint32_t wrap_lmask = 1u | (wrap_in_low ? (1u << (L - 1)) : 0u);
  __syncthreads();
  auto load = [&](int32_t g) -> T { return xxz_load<T, SHARDED>(x, xg, g, p.row0, p.n, rescale, inv); };

  // warp 0: the descriptor of a work item from its block-list entry (shared-memory look-ups only: the entry itself was
  // fetched an iteration earlier, so nothing here waits for global memory)
  auto describe = [&](const uint4 blk, XxzBlockDesc& d) {
    const int lane = tid;
    const uint32_t h = blk.x;
    const int pp = p.n_up - __popc(h);
    const int32_t base = (int32_t)blk.y, bs = (int32_t)pascal[m * 33 + pp];
    bool anti = false;
    int32_t dl = 0;
    if (lane < nhb) {
      anti = (((h >> lane) ^ (h >> (lane + 1))) & 1u) != 0;
      const int c = pp + __popc(h & ((1u << lane) - 1u));
      const int32_t dd = (int32_t)pascal[(m + lane) * 33 + c];
      dl = ((h >> lane) & 1u) ? dd : -dd;
    }
    const uint32_t mask = __ballot_sync(0xffffffffu, anti);
    if (anti) {
      const int k = __popc(mask & ((1u << lane) - 1u));
      const int32_t g0 = base + dl;  // the neighbour block: rows [g0, g0 + bs)
      int mode = 0;
      if (SHARDED) mode = (g0 >= p.row0 && g0 + bs <= p.row0 + p.n) ? 0 : ((g0 + bs <= p.row0 || g0 >= p.row0 + p.n) ? 1 : 2);
      d.mode[k] = mode;
      d.src[k] = mode == 1 ? (const void*)(xg + g0) : (const void*)(x + (g0 - p.row0));
    }
    if (lane == 0) {
      d.base = base;
      d.bs = bs;
      d.p = pp;
      d.i0 = max(0, p.row0 - base);
      d.i1 = min(bs, p.row0 + p.n - base);
      d.nhigh = __popc(mask);
      d.h0 = (L > m) ? (int32_t)(h & 1u) : -1;
      d.top = wrap_in_low ? -1 : (int32_t)((h >> (L - 1 - m)) & 1u);
      d.wrap_base = (int32_t)blk.z;
    }
  };
  // asynchronous copy of a block's x segment into tile `buf`.  Row-sharded: a block that straddles the boundary of the
  // local rows is staged through registers instead (its remote entries come from the gathered vector, re-scaled).
  auto stage = [&](const XxzBlockDesc& d, int buf) {
    T* xs = reinterpret_cast<T*>(smem_x + buf * tile_bytes);
    if (SHARDED && (d.i0 > 0 || d.i1 < d.bs)) {
      for (int i = tid; i < d.bs; i += NT) xs[i] = load(d.base + i);
    } else {
      const T* src = x + (d.base - p.row0);
      for (int i = tid; i < d.bs; i += NT) cp_async_elem<(int)sizeof(T)>(xs + i, src + i);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // element `i` of the x segment of the neighbour block of high bond k
  auto high = [&](const XxzBlockDesc& d, int k, uint32_t i) -> T {
    const T* src = reinterpret_cast<const T*>(d.src[k]);
    if (!SHARDED) return __ldg(src + i);
    const int mode = d.mode[k];
    if (mode == 0) return __ldg(src + i);
    if (mode == 1) {
      const T v = __ldcg(src + i);
      return rescale ? scale_real(v, inv) : v;
    }
    return load((int32_t)(src - x) + p.row0 + (int32_t)i);
  };

  int it = blockIdx.x;
  const int stride = (int)gridDim.x;
  uint4 entry = make_uint4(0u, 0u, 0u, 0u);  // warp 0: the block-list entry of the block after the next one
  if (tid < 32) {
    if (it < p.nblocks) describe(__ldg(p.blocks + it), desc[0]);
    if (it + stride < p.nblocks) entry = __ldg(p.blocks + it + stride);
  }
  __syncthreads();
  if (it < p.nblocks) stage(desc[0], 0);
  int cur_p = -1, buf = 0, ds = 0;
  double dot = 0.0;
  for (; it < p.nblocks; it += stride, buf ^= 1, ds = ds == 2 ? 0 : ds + 1) {
    const XxzBlockDesc& d = desc[ds];
    XxzBlockDesc& dnext = desc[ds == 2 ? 0 : ds + 1];
    const bool more = it + stride < p.nblocks;
    if (tid < 32 && more) {
      describe(entry, dnext);
      if (it + 2 * stride < p.nblocks) entry = __ldg(p.blocks + it + 2 * stride);
    }
    // ---- neighbour lists of this popcount class ----
    if (d.p != cur_p) {
      __syncthreads();  // the previous block's readers of the lists are done
      cur_p = d.p;
      const int32_t bs = d.bs;
      const uint16_t* lsp = p.t_ls + p.ls_off[cur_p];
      const uint32_t zero_slot = (uint32_t)bsmax * (uint32_t)sizeof(T);
      uint16_t* nb16 = reinterpret_cast<uint16_t*>(nb);
      for (int i = tid; i < bs; i += NT) {
        const uint32_t l = __ldg(lsp + i);
        int c = 0, anti = 0;
        uint16_t* row = nb16 + (size_t)i * (4 * E4);
        for (int b = 0; b < 4 * E4; ++b) {
          const uint32_t up = (l >> b) & 1u;
          uint32_t off = zero_slot;
          if (b + 1 < m) {
            if (((l >> (b + 1)) & 1u) != up) {
              const int32_t dd = (int32_t)pascal[b * 33 + c];
              off = (uint32_t)(up ? i + dd : i - dd) * (uint32_t)sizeof(T);
              ++anti;
            }
            c += (int)up;
          }
          row[b] = (uint16_t)off;
        }
        // bond across the split: rank distance C(m-1, c), direction from bit m-1
        const uint32_t dsplit = pascal[(m - 1) * 33 + c];
        const uint32_t upm = (l >> (m - 1)) & 1u;
        mt[i] = (uint16_t)(dsplit | (upm << 11) | ((uint32_t)anti << 12));
        ls[i] = (uint16_t)l;
      }
    }
    const int32_t base = d.base, i0 = d.i0, i1 = d.i1;
    const int nhigh = d.nhigh;
    // ---- R rows per thread and pass, bond-major: the loads of a bond are independent across the rows, only the
    //      additions of a row form a chain (in bond order); D bonds' worth of loads stay in flight, the first D of a
    //      pass are issued before the previous pass is finished (for the first pass: before the tile barrier).  A warp
    //      whose rows all lie beyond the block skips the pass. ----
    uint32_t ir[R];
    bool act[R];
    T hv[D][R];
    auto begin_pass = [&](int32_t ib) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int32_t i = ib + r * NT;
        act[r] = i < i1;
        ir[r] = (uint32_t)(act[r] ? i : i1 - 1);  // rows beyond the block repeat its last row (and store nothing)
      }
#pragma unroll
      for (int u = 0; u < D; ++u) {
        if (u < nhigh) {
#pragma unroll
          for (int r = 0; r < R; ++r) hv[u][r] = high(d, u, ir[r]);
        } else {
#pragma unroll
          for (int r = 0; r < R; ++r) hv[u][r] = zero_of(T());
        }
      }
    };
    int32_t ib = i0 + tid;
    if (ib - (tid & 31) < i1) begin_pass(ib);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();  // tile `buf` has landed; every thread is done with the previous block, i.e. with tile buf ^ 1
    if (more) stage(dnext, buf ^ 1);  // the next block's segment travels while this one is computed
    const T* xs = reinterpret_cast<const T*>(smem_x + buf * tile_bytes);
    const unsigned char* xb = reinterpret_cast<const unsigned char*>(xs);
    T* yb = y + (base - p.row0);
    while (ib - (tid & 31) < i1) {
      uint32_t meta[R];
      T vsplit[R], vwrap[R];
      int extra[R];  // anti-parallel bonds of the row beyond the low and the high ones
#pragma unroll
      for (int r = 0; r < R; ++r) {
        meta[r] = mt[ir[r]];
        const uint32_t upm = (meta[r] >> 11) & 1u;
        const bool split = d.h0 >= 0 && (int32_t)upm != d.h0;
        const int32_t dd = (int32_t)(meta[r] & 0x7ffu);
        vsplit[r] = split ? load(base + (int32_t)ir[r] + (upm ? dd : -dd)) : zero_of(T());
        extra[r] = split ? 1 : 0;
        vwrap[r] = zero_of(T());
      }
      if (p.periodic) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const uint32_t l = ls[ir[r]];
          const uint32_t topbit = wrap_in_low ? (l >> (L - 1)) & 1u : (uint32_t)d.top;
          const bool wrap = (l & 1u) != topbit;
          if (wrap) vwrap[r] = load(d.wrap_base + (int32_t)__ldg(p.t_lo + (l ^ wrap_lmask)));
          extra[r] += wrap ? 1 : 0;
        }
      }
      T acc[R];
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = zero_of(T());
      for (int e = 0; e < E4; ++e) {
        uint2 q[R];
#pragma unroll
        for (int r = 0; r < R; ++r) q[r] = nb[(size_t)ir[r] * E4 + e];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const T v0 = *reinterpret_cast<const T*>(xb + (q[r].x & 0xffffu)), v1 = *reinterpret_cast<const T*>(xb + (q[r].x >> 16));
          const T v2 = *reinterpret_cast<const T*>(xb + (q[r].y & 0xffffu)), v3 = *reinterpret_cast<const T*>(xb + (q[r].y >> 16));
          acc[r] = add_t(add_t(add_t(add_t(acc[r], v0), v1), v2), v3);
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = add_t(acc[r], vsplit[r]);
      for (int k0 = 0; k0 < nhigh; k0 += D) {
#pragma unroll
        for (int u = 0; u < D; ++u) {
#pragma unroll
          for (int r = 0; r < R; ++r) acc[r] = add_t(acc[r], hv[u][r]);
          if (k0 + u + D < nhigh) {  // refill the slot just consumed
#pragma unroll
            for (int r = 0; r < R; ++r) hv[u][r] = high(d, k0 + u + D, ir[r]);
          } else {
#pragma unroll
            for (int r = 0; r < R; ++r) hv[u][r] = zero_of(T());
          }
        }
      }
      // this pass's rows, then the next pass's first loads (while the results below are formed and stored)
      uint32_t irow[R];
      bool arow[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        irow[r] = ir[r];
        arow[r] = act[r];
      }
      ib += NT * R;
      if (ib - (tid & 31) < i1) begin_pass(ib);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (arow[r]) {
          const T a = add_t(acc[r], vwrap[r]);
          const int anti = (int)(meta[r] >> 12) + nhigh + extra[r];
          const RT diag = (RT)(p.jz4 * (double)(nbonds - 2 * anti));
          const T xi = xs[irow[r]];
          const T yi = xxz_row(a, xi, (RT)p.jxy2, diag + sigma);
          yb[irow[r]] = yi;
          dot += re_conj_mul(xi, yi);
        }
      }
    }
  }
  const double t = block_sum<NW>(dot, scratch);
  finish_scalar<NW>(t, pa, msg, scratch);
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
  uint32_t* d_states = nullptr;  // basis states of the local block (per-state kernel and set-up only)
  // block kernel: popcount-class tables and the block list
  bool per_state = false;  // LLZ_XXZ_KERNEL=state: the plain one-thread-per-state kernel (kept as the A/B reference)
  XxzBlockParams bp;
  uint16_t* d_tlo = nullptr;
  uint16_t* d_tls = nullptr;
  uint4* d_blocks = nullptr;
  size_t block_smem = 0;
  int block_grid = 1;
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
    if (d_tlo) dev_free(ctx, d_tlo);
    if (d_tls) dev_free(ctx, d_tls);
    if (d_blocks) dev_free(ctx, d_blocks);
    if (xb) comm_exchange_buffer_release(ctx, xb);
  }
  virtual int plan_block_kernel() = 0;
  const char* storage() const override { return per_state ? "XXZ matrix-free (state table)" : "XXZ matrix-free (block kernel)"; }
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
  // threads per CTA, rows per thread and pass, high bonds in flight: measured best of {1,2,4} x {2,4,8} rows x bonds and
  // 256/512 threads on L = 28 for double and complex<double> (tools/bench_xxz.py, profiles/r02_xxz_block_kernel.md)
  template <class F> int dispatch(F&& f) {
    using std::integral_constant;
    return f(integral_constant<int, 256>(), integral_constant<int, 2>(), integral_constant<int, 4>());
  }
  // shared memory of the block kernel and the persistent grid that keeps every SM full
  int plan_block_kernel() override {
    const size_t tile = (((size_t)bp.bsmax + 1) * sizeof(T) + 15) / 16 * 16;  // two of them: the next block's segment is prefetched
    block_smem = 2 * tile + (size_t)((bp.m - 1 + 3) / 4 * 4 + 2) * bp.bsmax * sizeof(uint16_t);
    if (block_smem > 200 * 1024) return fail(LLZ_ERR_INVALID, "op_create_xxz: %d low bits need %zu bytes of shared memory", bp.m, block_smem);
    int per_sm = 1;
    const size_t smem = block_smem;
    LLZ_TRY(dispatch([&](auto nt, auto r, auto dd) {
      constexpr int NT = decltype(nt)::value, RR = decltype(r)::value, DD = decltype(dd)::value;
      if (smem > 48 * 1024) {
        cudaFuncSetAttribute(k_xxz_block_apply<T, false, NT, RR, DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_xxz_block_apply<T, true, NT, RR, DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      }
      int a = 0, b = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_xxz_block_apply<T, false, NT, RR, DD>, NT, smem);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_xxz_block_apply<T, true, NT, RR, DD>, NT, smem);
      per_sm = std::max(1, std::min(a, b));
      return (int)LLZ_OK;
    }));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(LLZ_ERR_CUDA, "op_create_xxz: block kernel set-up: %s", cudaGetErrorString(e));
    block_grid = (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)bp.nblocks, (int64_t)kMaxGrid, (int64_t)ctx->num_sms * per_sm}));
    return LLZ_OK;
  }
  int launch_block(const void* x, void* y, double sigma, double* pa, const PeerMsg& msg) {
    bp.x_all = prm.x_all;
    bp.x_scale = prm.x_scale;
    return dispatch([&](auto nt, auto r, auto dd) {
      constexpr int NT = decltype(nt)::value, RR = decltype(r)::value, DD = decltype(dd)::value;
      cudaError_t e;
      if (prm.x_all)
        e = launch_chain(ctx, k_xxz_block_apply<T, true, NT, RR, DD>, block_grid, NT, block_smem, (const T*)x, (T*)y, bp,
                         (typename Num<T>::R)sigma, pa, msg, cur_gather_msg);
      else
        e = launch_chain(ctx, k_xxz_block_apply<T, false, NT, RR, DD>, block_grid, NT, block_smem, (const T*)x, (T*)y, bp,
                         (typename Num<T>::R)sigma, pa, msg, PeerMsg());
      if (e != cudaSuccess) return fail(LLZ_ERR_CUDA, "launch k_xxz_block_apply: %s", cudaGetErrorString(e));
      return (int)LLZ_OK;
    });
  }
  int apply_fused(const void* x, void* y, double sigma, double* pa, int* npa, const PeerMsg* alpha_msg) override {
    const PeerMsg msg = alpha_msg ? *alpha_msg : PeerMsg();
    cudaError_t e;
    if (!per_state) {
      LLZ_TRY(launch_block(x, y, sigma, pa, msg));
      *npa = block_grid;
      ctx->launches++;
      return LLZ_OK;
    }
    // persistent: exactly the 6 CTAs per SM that __launch_bounds__(kThreads, 6) keeps resident
    int64_t g = std::min<int64_t>((n_local + kThreads - 1) / kThreads, std::min<int64_t>(kMaxGrid, (int64_t)ctx->num_sms * 6));
    if (g < 1) g = 1;
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

// Low bits of the split used by the block kernel: 12 (blocks of up to 924 rows, ~40 KB of shared memory per CTA: four
// resident CTAs; measured best on L = 28, 11 and 13 are ~20 % slower) once the sector is large; fewer for short chains
// so that there are enough blocks to spread over the SMs.  LLZ_XXZ_M overrides (tests cover every split).
static int xxz_low_bits(int L) {
  int m = std::min(L, std::max(6, std::min(12, L - 10)));
  if (const char* env = getenv("LLZ_XXZ_M")) {
    const int v = atoi(env);
    if (v >= 1) m = std::min({v, L, kXxzMaxLow});
  }
  return m;
}

// Popcount-class tables and the block list of the block kernel (host, O(2^m + 2^(L-m))).
static int build_block_tables(XxzOpBase* op, int L, int n_up) {
  llz_ctx_t ctx = op->ctx;
  const int m = xxz_low_bits(L);
  const int hb = L - m;
  XxzBlockParams& bp = op->bp;
  bp.L = L;
  bp.n_up = n_up;
  bp.m = m;
  bp.periodic = op->prm.periodic;
  bp.jz4 = op->prm.jz4;
  bp.jxy2 = op->prm.jxy2;
  bp.n = (int32_t)op->n_local;
  bp.row0 = (int32_t)op->row0;
  bp.x_all = nullptr;
  bp.x_scale = nullptr;
  std::vector<uint16_t> tlo((size_t)1 << m), tls((size_t)1 << m);
  bp.ls_off[0] = 0;
  for (int pp = 0; pp <= m; ++pp) bp.ls_off[pp + 1] = bp.ls_off[pp] + (int)host_binom[m][pp];
  for (uint32_t l = 0; l < tlo.size(); ++l) {
    unsigned long long r = 0;
    int j = 0;
    for (int q = 0; q < m; ++q)
      if (l >> q & 1u) r += host_binom[q][++j];
    tlo[l] = (uint16_t)r;
    tls[(size_t)bp.ls_off[j] + r] = (uint16_t)l;
  }
  std::vector<uint32_t> thi((size_t)1 << hb, 0u);
  // blocks that intersect the local rows, by popcount class; classes in order of decreasing block size
  std::vector<std::vector<uint4>> by_class((size_t)m + 1);
  int bsmax = 1;
  for (uint32_t h = 0; h < thi.size(); ++h) {
    const int pp = n_up - __builtin_popcount(h);
    if (pp < 0 || pp > m) continue;
    unsigned long long r = 0;
    int j = pp;
    for (int q = 0; q < hb; ++q)
      if (h >> q & 1u) r += host_binom[q + m][++j];
    thi[h] = (uint32_t)r;
    const int64_t bs = (int64_t)host_binom[m][pp];
    if ((int64_t)r < op->row0 + op->n_local && (int64_t)r + bs > op->row0) {
      by_class[(size_t)pp].push_back(make_uint4(h, (uint32_t)r, 0u, 0u));
      bsmax = std::max(bsmax, (int)bs);
    }
  }
  std::vector<int> order;
  for (int pp = 0; pp <= m; ++pp) order.push_back(pp);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return host_binom[m][a] > host_binom[m][b]; });
  std::vector<uint4> blocks;
  for (int pp : order) blocks.insert(blocks.end(), by_class[(size_t)pp].begin(), by_class[(size_t)pp].end());
  if (L - 1 >= m)  // the wrap bond toggles bit L-1, which lies in the high part: first row of the block it leads to
    for (uint4& b : blocks) b.z = thi[b.x ^ (1u << (L - 1 - m))];
  if (blocks.empty()) return fail(LLZ_ERR_INVALID, "op_create_xxz: no basis state in the local row block");
  bp.nblocks = (int)blocks.size();
  bp.bsmax = (bsmax + 1) & ~1;
  cudaError_t e = dev_malloc(ctx, &op->d_tlo, tlo.size() * sizeof(uint16_t));
  if (e == cudaSuccess) e = dev_malloc(ctx, &op->d_tls, tls.size() * sizeof(uint16_t));
  if (e == cudaSuccess) e = dev_malloc(ctx, &op->d_blocks, blocks.size() * sizeof(uint4));
  if (e == cudaSuccess) e = cudaMemcpy(op->d_tlo, tlo.data(), tlo.size() * sizeof(uint16_t), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(op->d_tls, tls.data(), tls.size() * sizeof(uint16_t), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(op->d_blocks, blocks.data(), blocks.size() * sizeof(uint4), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return fail(e == cudaErrorMemoryAllocation ? LLZ_ERR_OOM : LLZ_ERR_CUDA, "op_create_xxz: block tables: %s", cudaGetErrorString(e));
  bp.t_lo = op->d_tlo;
  bp.t_ls = op->d_tls;
  bp.blocks = op->d_blocks;
  op->bytes = (int64_t)(tlo.size() + tls.size()) * 2 + (int64_t)blocks.size() * 16;
  return op->plan_block_kernel();
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
  {
    const char* env = getenv("LLZ_XXZ_KERNEL");
    op->per_state = env && env[0] == 's';
  }
  if (!op->per_state) {
    const int st = build_block_tables(op, L, n_up);
    if (st != LLZ_OK) {
      delete op;
      return st;
    }
  } else {
    op->bytes = op->n_local * 4;  // the state table is the only stored part of the operator
  }
  // the state table: what the per-state kernel reads, and what the row-sharded set-up derives its index ranges from
  if (op->per_state || ctx->nranks > 1) {
    e = dev_malloc(ctx, &op->d_states, sizeof(uint32_t) * (size_t)op->n_local);
    if (e != cudaSuccess) {
      delete op;
      return fail(LLZ_ERR_OOM, "op_create_xxz: state table (%lld entries): %s", (long long)op->n_local, cudaGetErrorString(e));
    }
  }
  if (op->d_states) {
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
  if (!op->per_state && op->d_states) {  // only the set-up needed it
    dev_free(ctx, op->d_states);
    op->d_states = nullptr;
  }
  llz_op_t h = new llz_op_s();
  h->impl = op;
  *out = h;
  return LLZ_OK;
}
