// llz_peer.cuh — scalar reductions across the GPUs of one box WITHOUT a collective call: every rank stores its
// contribution straight into every peer's "inbox" over NVLink/NVSwitch (CUDA IPC mapped peer memory) from inside the
// kernel that produced it, and the consuming kernel sums the G contributions in rank order in its prologue.
//
//   producer (any kernel)                      consumer (next kernel on every rank)
//   -------------------------------------      ------------------------------------------------
//   st.global   peer[p].data[parity][me][i]    thread r < G: spin on ld.acquire.sys local.flag[parity][r] >= seq
//   fence.sys                                  bar.sync
//   st.release.sys peer[p].flag[parity][me]    sum_r ld.cg local.data[parity][r][i]   (rank order => every GPU
//        = seq                                                                          obtains the SAME bits)
//
// This replaces "reduce kernel + ncclAllReduce + consumer" (three stream operations, ~10-20 us of launch + protocol
// latency for 8 bytes) by one NVLink store and one poll (~2-4 us), and it is what lets the Lanczos iteration keep its
// single-GPU kernel sequence when row-sharded.  Slots are double-buffered by the parity of the sequence number: a rank
// can only reach message seq+2 after it consumed seq+1 from every peer, which every peer produced after consuming seq.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace llz {

constexpr int kMaxRanks = 16;

// One message channel, passed to kernels by value.  `inbox[p]` is the address, valid on THIS GPU, of rank p's inbox
// region: data [2][G][payload] doubles followed by flags [2][G] 64-bit sequence numbers.
struct PeerChannel {
  int G = 0;        // 0 => channel unused (single rank or NCCL path)
  int rank = 0;
  int payload = 0;  // doubles per slot
  unsigned long long* status = nullptr;  // device word: set non-zero when a wait timed out (a peer died)
  double* inbox[kMaxRanks] = {};
};

// A group-wide scalar or coefficient vector delivered through a channel: message `seq` of `ch`.
// ch.G == 0 means "not used": the consumer reads this rank's own device memory instead.
struct PeerMsg {
  PeerChannel ch;
  unsigned long long seq = 0;
  unsigned int* ticket = nullptr;  // device counter for last-CTA detection in the producing kernel
};

// Fused all-gather: where the producer of a vector additionally stores its row block (one destination per peer,
// already offset to this rank's rows inside the peer's exchange buffer) and the message that announces it.
struct GatherPush {
  int G = 0;  // 0 => no push
  int rank = 0;
  void* dst[kMaxRanks] = {};
  // peer p only needs the local elements [lo[p], hi[p]) (multiples of 4, so whole 128-bit packets); empty: lo >= hi
  long long lo[kMaxRanks] = {};
  long long hi[kMaxRanks] = {};
  PeerMsg msg;
};

// Halo of a row-sharded CSR / SELL / DIA operator pushed by the owner: entries idx[start[q] .. start[q+1]) of the local
// vector go to dst[q], rank q's halo segment for this rank (mapped here).  Either a kernel of its own before the apply
// (k_halo_push) or the tail of the fused orthogonalisation kernel, which has just produced the vector.
struct HaloPush {
  int G = 0;
  long long start[kMaxRanks + 1] = {};  // send entries [start[q], start[q+1]) go to rank q
  void* dst[kMaxRanks] = {};            // where they go
};
struct HaloPushPlan {
  HaloPush hp;                  // hp.G == 0: nothing to push
  const int32_t* idx = nullptr; // [n_send] local indices, grouped by requesting rank
  long long n_send = 0;
  PeerMsg msg;                  // announced by the last CTA once every entry is stored
};

#ifdef __CUDACC__
__device__ __forceinline__ double* peer_slot(const PeerChannel& ch, int owner, unsigned long long seq, int src) {
  return ch.inbox[owner] + ((size_t)(seq & 1ull) * ch.G + src) * ch.payload;
}
__device__ __forceinline__ unsigned long long* peer_flag(const PeerChannel& ch, int owner, unsigned long long seq, int src) {
  return reinterpret_cast<unsigned long long*>(ch.inbox[owner] + (size_t)2 * ch.G * ch.payload) + (seq & 1ull) * ch.G + src;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Called by ONE thread after this rank's data stores to owner `p` are complete (and fenced by the caller if other
// threads made them): announces message `seq` to rank p.
__device__ __forceinline__ void peer_announce(const PeerChannel& ch, int p, unsigned long long seq) {
  st_release_sys(peer_flag(ch, p, seq, ch.rank), seq);
}

// Called by ALL threads of a CTA (blockDim >= G): returns once message `seq` of every rank has landed in the local
// inbox.  Gives up after ~2 minutes (a peer process died) and raises the status word so the host reports it instead
// of hanging the GPU for ever.
__device__ __forceinline__ void peer_wait(const PeerChannel& ch, unsigned long long seq) {
  if ((int)threadIdx.x < ch.G) {
    const unsigned long long* f = peer_flag(ch, ch.rank, seq, threadIdx.x);
    if (ld_acquire_sys(f) < seq) {
      const unsigned long long t0 = global_timer_ns();
      unsigned spins = 0;
      while (ld_acquire_sys(f) < seq) {
        if ((++spins & 0xfff) == 0) {
          if (ch.status && *reinterpret_cast<volatile unsigned long long*>(ch.status) != 0) break;  // already failed
          if (global_timer_ns() - t0 > 120000000000ull) {
            if (ch.status) atomicExch(ch.status, seq | (1ull << 63));
            break;
          }
        }
      }
    }
  }
  __syncthreads();
}

// Sum over the ranks, in rank order, of element i of message `seq` (after peer_wait).
__device__ __forceinline__ double peer_sum(const PeerChannel& ch, unsigned long long seq, int i) {
  double s = 0.0;
  for (int r = 0; r < ch.G; ++r) s += __ldcg(peer_slot(ch, ch.rank, seq, r) + i);
  return s;
}
#endif

}  // namespace llz
