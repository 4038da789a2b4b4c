// llz_internal.hpp — host-side object definitions behind the opaque handles of include/llz.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <utility>
#include <vector>

#include "../../include/llz.h"

namespace llz {

void set_error(const char* fmt, ...);
int fail(int status, const char* fmt, ...);

#define LLZ_CUDA(expr)                                                                                  \
  do {                                                                                                  \
    cudaError_t e__ = (expr);                                                                           \
    if (e__ != cudaSuccess)                                                                             \
      return ::llz::fail(e__ == cudaErrorMemoryAllocation ? LLZ_ERR_OOM : LLZ_ERR_CUDA, "%s: %s (%s:%d)", #expr, \
                         cudaGetErrorString(e__), __FILE__, __LINE__);                                  \
  } while (0)

#define LLZ_TRY(expr)             \
  do {                            \
    int s__ = (expr);             \
    if (s__ != LLZ_OK) return s__; \
  } while (0)

inline size_t dtype_size(int dtype) {
  switch (dtype) {
    case LLZ_F32: return 4;
    case LLZ_F64: return 8;
    case LLZ_C64: return 8;
    case LLZ_C128: return 16;
  }
  return 0;
}
inline int dtype_nc(int dtype) { return (dtype == LLZ_C64 || dtype == LLZ_C128) ? 2 : 1; }

struct Comm;         // inter-GPU plumbing (llz_comm.cu)
struct PeerChannel;  // peer-memory message channel (llz_peer.cuh)
struct PeerMsg;
struct GatherPush;
struct HaloPushPlan;

// Per-kernel device-time accounting (CUDA events on the context's stream), keyed by a short kernel-family name.
struct ProfEntry {
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
  double ms = 0.0;
  int64_t launches = 0;
  double bytes = 0.0;  // algorithmic bytes the timed launches had to move (SURVEY.md §8d accounting)
};

}  // namespace llz

struct llz_ctx_s {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int num_sms = 0;
  size_t l2_bytes = 0;
  uint64_t launches = 0;
  int rank = 0, nranks = 1;
  llz::Comm* comm = nullptr;
  // scratch shared by the stand-alone vector kernels: per-CTA partials + pinned result slot
  double* d_partials = nullptr;  // kMaxGrid * 2 doubles
  double* d_result = nullptr;    // 8 doubles
  double* h_result = nullptr;    // pinned, 8 doubles
  void** d_ptrs = nullptr;       // pointer table for llz_vec_schmidt_orth
  double* d_coef = nullptr;      // coefficients for llz_vec_schmidt_orth
  double* d_ph = nullptr;        // h partials for llz_vec_schmidt_orth
  size_t ph_capacity = 0;
  int64_t ptr_capacity = 0;
  // Device-memory cache of the context (SURVEY.md §8b "ownership": the basis slab and work vectors are owned by the
  // ctx and reused across runs).  Everything runs on ONE stream, so a buffer handed back can be reused by later
  // stream-ordered work without a synchronisation.
  std::multimap<size_t, void*> vec_pool;  // free device buffers by size
  std::map<void*, size_t> dev_sizes;      // live dev_malloc allocations
  size_t vec_pool_bytes = 0;
  size_t vec_pool_limit = 0;              // set at creation (a quarter of the device memory)
  llz_krylov_t cached_krylov = nullptr;   // last destroyed Krylov workspace, revived by a matching llz_krylov_create
  // grid-wide barrier of the fused orthogonalisation kernel: a device counter that only ever grows, and its value
  // after the launches enqueued so far (every launch adds 3 x its grid)
  unsigned long long* d_bar = nullptr;
  unsigned long long bar_count = 0;
  bool fuse_orth = true;  // LLZ_FUSED_ORTH=0 keeps the separate kernels (A/B measurements, tests)
  // profiling
  bool profile = false;
  std::map<std::string, llz::ProfEntry> prof;
  std::vector<cudaEvent_t> event_pool;
};

struct llz_vec_s {
  llz_ctx_t ctx = nullptr;
  int dtype = 0;
  int64_t n = 0;
  void* d = nullptr;
  bool owned = true;
};

namespace llz {

constexpr int kMaxGrid = 4096;  // upper bound on the grid of any persistent kernel (=> on per-CTA partial arrays)

// Operator interface (device side of the reference's mv_mul)
struct OpBase {
  llz_ctx_t ctx = nullptr;
  int dtype = 0;
  int64_t n_local = 0;
  int64_t n_global = 0;  // rows of the whole operator (= n_local for a single rank)
  int64_t row0 = 0;      // first global row of the local block
  int64_t bytes = 0;
  virtual ~OpBase() {}
  virtual const char* storage() const { return "operator"; }  // how the operator is held on the device (llz_op_storage)
  // Row-sharded operators: make the remote entries of x that the local rows reference available (halo exchange /
  // all-gather over NCCL, enqueued on ctx->stream).  Called once before every apply_fused.
  virtual int prepare(const void* x) {
    (void)x;
    return LLZ_OK;
  }
  // Fused all-gather (row-sharded operators that read the WHOLE input vector): instead of gathering x before the
  // apply, the kernel that produces the vector (update / recurrence of the previous iteration) also stores its block
  // into every peer's exchange buffer.  plan_push() hands the producer the destinations and the message to announce
  // (false: not supported, gather in prepare()); use_pushed() tells the operator that its next apply finds the input
  // in the pushed buffers, still un-normalised: remote entries are multiplied by 1 / *scale on the fly.
  virtual bool plan_push(GatherPush* push) {
    (void)push;
    return false;
  }
  virtual void use_pushed(const GatherPush& push, const double* scale) {
    (void)push;
    (void)scale;
  }
  // Halo of a row-sharded sparse operator: the kernel that produces (and normalises) the next input vector stores the
  // entries the peers reference into their halo segments itself.  plan_halo_push() hands it the plan (false: not
  // supported / nothing to send — prepare() does the exchange before the apply as usual); use_pushed_halo() tells the
  // operator that its next apply finds the halo announced by that plan's message.
  virtual bool plan_halo_push(HaloPushPlan* plan) {
    (void)plan;
    return false;
  }
  virtual void use_pushed_halo(const HaloPushPlan& plan) { (void)plan; }
  // max_i sum_j |a_ij| over the LOCAL rows (Gerschgorin radius); LLZ_ERR_UNSUPPORTED for operators without stored
  // or analytically known entries (user callbacks).
  virtual int abs_row_sum_max(double* out) {
    (void)out;
    return fail(LLZ_ERR_UNSUPPORTED, "this operator kind cannot report its row sums");
  }
  // y = A x + sigma x ; per-CTA partials of Re<x,y> into alpha_partials[0..*n_partials) (device), all on ctx->stream.
  // Returns LLZ_OK or an error.  Implementations that cannot fuse the dot leave *n_partials = 0 and the engine runs
  // a separate dot kernel.
  // Row-sharded with peer channels: `alpha_msg` (may be null) asks the kernel to also deliver the rank's alpha to every
  // GPU (finish_scalar, llz_device.cuh); implementations that leave *n_partials = 0 ignore it.
  virtual int apply_fused(const void* x, void* y, double sigma, double* alpha_partials, int* n_partials,
                          const PeerMsg* alpha_msg = nullptr) = 0;
};

}  // namespace llz

namespace llz {
// Group-wide sums over the row-sharded ranks (no-ops for a single rank).  `d` is device memory on ctx->stream.
int comm_allreduce_sum(llz_ctx_t ctx, double* d, int count);
// Per-CTA partials of a scalar: leaves them alone for one rank; otherwise folds them to one group-wide value
// (d[0], *count = 1).
int comm_allreduce_partials(llz_ctx_t ctx, double* d, int* count);
void comm_destroy(llz_ctx_t ctx);
// Peer-memory channels (llz_peer.cuh) set up by llz_ctx_join: 0 = alpha, 1 = beta^2, 2 = projection coefficients
bool comm_p2p(llz_ctx_t ctx);
int comm_coef_capacity(llz_ctx_t ctx);
int comm_check_peers(llz_ctx_t ctx);
// Halo planning (llz_halo.cpp): one pass producing the sorted unique remote columns, their count per owner and the
// column indices in the local extended numbering.
int halo_plan_vectors(int64_t n_rows, int64_t row0, const int64_t* rowptr, const int32_t* colidx, int nranks,
                      const int64_t* boundaries, int32_t* colidx_local, std::vector<int32_t>& remote, int64_t* per_owner);
// Whole-vector exchange buffers for the fused all-gather (llz_comm.cu)
struct ExchangeBuffer {
  size_t bytes = 0;
  void* local = nullptr;
  void* peer[16] = {};  // address of every rank's buffer, valid on this GPU (peer[rank] == local)
  bool in_use = false, usable = false;
};
ExchangeBuffer* comm_exchange_buffer_acquire(llz_ctx_t ctx, size_t bytes);
void comm_exchange_buffer_release(llz_ctx_t ctx, ExchangeBuffer* b);
int64_t comm_window_alloc(llz_ctx_t ctx, size_t bytes);
void comm_window_free(llz_ctx_t ctx, int64_t off, size_t bytes);
void* comm_window_ptr(llz_ctx_t ctx, int rank, int64_t off);
// Pooled device allocations of a context (llz_ctx.cu)
int ctx_alloc(llz_ctx_t ctx, size_t bytes, void** out);
void ctx_free(llz_ctx_t ctx, void* p, size_t bytes);
// Pooled device allocation for operator arrays: served from the context's free list when a buffer of the same
// (256-byte rounded) size is there, else cudaMalloc — which gives the context's cached memory back to the driver and
// retries once when the device is full.  dev_free returns the buffer to the free list (safe without a synchronisation
// because every kernel and copy of a context is ordered on its one stream).
cudaError_t dev_malloc(llz_ctx_t ctx, void** p, size_t bytes);
void dev_free(llz_ctx_t ctx, void* p);
template <class P> inline cudaError_t dev_malloc(llz_ctx_t ctx, P** p, size_t bytes) { return dev_malloc(ctx, (void**)p, bytes); }
void ctx_trim(llz_ctx_t ctx);            // release every cached buffer (vector pool + cached Krylov workspace)
void krylov_destroy_now(llz_krylov_t k); // really free a workspace (llz_krylov.cu)
int comm_allgather_bytes(llz_ctx_t ctx, const void* send, void* recv, size_t bytes_per_rank);
int comm_allgather_host(llz_ctx_t ctx, const void* send, void* recv, size_t bytes_per_rank);
// Grouped point-to-point exchange (device buffers): for every peer p send send_bytes[p] from send_base + send_off[p]
// and receive recv_bytes[p] into recv_base + recv_off[p].
int comm_exchange(llz_ctx_t ctx, const char* send_base, const size_t* send_off, const size_t* send_bytes, char* recv_base,
                  const size_t* recv_off, const size_t* recv_bytes);

// RAII: records a start/stop event pair around the launches issued in its scope when ctx->profile is on.
struct ProfScope {
  llz_ctx_t ctx;
  ProfEntry* entry = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  ProfScope(llz_ctx_t c, const char* name, double bytes) : ctx(c) {
    if (!c->profile) return;
    entry = &c->prof[name];
    entry->bytes += bytes;
    e0 = take();
    e1 = take();
    cudaEventRecord(e0, c->stream);
  }
  cudaEvent_t take() {
    if (!ctx->event_pool.empty()) {
      cudaEvent_t e = ctx->event_pool.back();
      ctx->event_pool.pop_back();
      return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
  }
  ~ProfScope() {
    if (!entry) return;
    cudaEventRecord(e1, ctx->stream);
    entry->pending.emplace_back(e0, e1);
    entry->launches++;
  }
};
}  // namespace llz

struct llz_op_s {
  llz::OpBase* impl = nullptr;
};
