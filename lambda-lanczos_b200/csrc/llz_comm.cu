// llz_comm.cu — inter-GPU plumbing for row-sharded runs: one process per GPU of one box.
//
//   * an NCCL communicator per joined context (bound at run time, see NcclApi) for set-up traffic, stand-alone vector
//     reductions and as the fall-back transport;
//   * CUDA-IPC mapped peer memory for everything inside the Lanczos iteration (llz_peer.cuh): the scalar channels
//     (alpha, beta^2, projection coefficients, halo / gather announcements), a halo window that row-sharded CSR / SELL
//     operators sub-allocate, and whole-vector exchange buffers for operators that read the entire input vector.
//     Producing kernels store into the peers' memory over NVLink and consuming kernels wait in their prologue, so a
//     sharded iteration issues no collective call at all.
// Single-rank contexts never reach any of this.
#include <dlfcn.h>
#include <nccl.h>  // types and enums only: the library itself is bound at run time (see NcclApi)

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <utility>
#include <vector>

#include "llz_launch.hpp"
#include "llz_peer.cuh"

namespace llz {

// NCCL is bound lazily, the first time a context joins a group: single-GPU processes never load it, and a process
// that already holds a copy (e.g. PyTorch's bundled libnccl.so.2) shares that copy instead of pulling a second,
// possibly older one under the same SONAME.
struct NcclApi {
  bool ok = false;
  const char* why = "";
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi& nccl() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // a copy this process already uses
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
  if (!h) {
    api.why = "libnccl.so.2 not found";
    return api;
  }
  auto get = [&](const char* name, void** fn) {
    *fn = dlsym(h, name);
    return *fn != nullptr;
  };
  api.ok = get("ncclGetUniqueId", (void**)&api.GetUniqueId) && get("ncclCommInitRank", (void**)&api.CommInitRank) &&
           get("ncclCommDestroy", (void**)&api.CommDestroy) && get("ncclAllReduce", (void**)&api.AllReduce) &&
           get("ncclAllGather", (void**)&api.AllGather) && get("ncclSend", (void**)&api.Send) && get("ncclRecv", (void**)&api.Recv) &&
           get("ncclGroupStart", (void**)&api.GroupStart) && get("ncclGroupEnd", (void**)&api.GroupEnd) &&
           get("ncclGetErrorString", (void**)&api.GetErrorString);
  if (!api.ok) api.why = "libnccl.so.2 lacks an expected symbol";
  return api;
}

struct Comm {
  ncclComm_t nccl = nullptr;
  // peer-memory channels for the per-iteration scalar reductions (llz_peer.cuh); G == 0 when unavailable
  void* p2p_local = nullptr;
  void* p2p_peer[kMaxRanks] = {};
  PeerChannel ch[5];
  unsigned long long seq[5] = {0, 0, 0, 0, 0};
  std::vector<ExchangeBuffer*> xbufs;  // IPC-mapped whole-vector buffers (fused all-gather), kept until the context dies
  unsigned int* ticket = nullptr;
  // halo window: part of the same IPC-mapped allocation, sub-allocated by row-sharded operators for the entries of x
  // their peers push to them (first-fit free list of [offset, bytes), offsets relative to the window)
  size_t win_off = 0, win_bytes = 0;
  std::vector<std::pair<size_t, size_t>> win_free;
};

#define LLZ_NCCL(expr)                                                                                       \
  do {                                                                                                       \
    ncclResult_t r__ = (expr);                                                                               \
    if (r__ != ncclSuccess) return ::llz::fail(LLZ_ERR_COMM, "%s: %s", #expr, nccl().GetErrorString(r__));   \
  } while (0)

int comm_allreduce_sum(llz_ctx_t ctx, double* d, int count) {
  if (ctx->nranks == 1 || count <= 0) return LLZ_OK;
  LLZ_NCCL(nccl().AllReduce(d, d, (size_t)count, ncclDouble, ncclSum, ctx->comm->nccl, ctx->stream));
  return LLZ_OK;
}

int comm_allreduce_partials(llz_ctx_t ctx, double* d, int* count) {
  if (ctx->nranks == 1) return LLZ_OK;
  // fold this rank's per-CTA partials into d[0] (fixed order), then sum over the group
  LLZ_TRY(launch_sum_partials(ctx, d, *count, 1, d, nullptr));
  LLZ_NCCL(nccl().AllReduce(d, d, 1, ncclDouble, ncclSum, ctx->comm->nccl, ctx->stream));
  *count = 1;
  return LLZ_OK;
}

int comm_allgather_bytes(llz_ctx_t ctx, const void* send, void* recv, size_t bytes_per_rank) {
  if (ctx->nranks == 1) {
    if (send != recv) {
      cudaError_t e = cudaMemcpyAsync(recv, send, bytes_per_rank, cudaMemcpyDeviceToDevice, ctx->stream);
      if (e != cudaSuccess) return fail(LLZ_ERR_CUDA, "allgather copy: %s", cudaGetErrorString(e));
    }
    return LLZ_OK;
  }
  LLZ_NCCL(nccl().AllGather(send, recv, bytes_per_rank, ncclChar, ctx->comm->nccl, ctx->stream));
  return LLZ_OK;
}

// Grouped point-to-point exchange: for every peer p, send send_bytes[p] from send_base + send_off[p] and receive
// recv_bytes[p] into recv_base + recv_off[p].
int comm_exchange(llz_ctx_t ctx, const char* send_base, const size_t* send_off, const size_t* send_bytes,
                  char* recv_base, const size_t* recv_off, const size_t* recv_bytes) {
  if (ctx->nranks == 1) return LLZ_OK;
  LLZ_NCCL(nccl().GroupStart());
  for (int p = 0; p < ctx->nranks; ++p) {
    if (p == ctx->rank) continue;
    if (send_bytes[p]) LLZ_NCCL(nccl().Send(send_base + send_off[p], send_bytes[p], ncclChar, p, ctx->comm->nccl, ctx->stream));
    if (recv_bytes[p]) LLZ_NCCL(nccl().Recv(recv_base + recv_off[p], recv_bytes[p], ncclChar, p, ctx->comm->nccl, ctx->stream));
  }
  LLZ_NCCL(nccl().GroupEnd());
  return LLZ_OK;
}

// Host-to-host all-gather of a small fixed-size record per rank (set-up time only: goes through device scratch and
// synchronises the stream).
int comm_allgather_host(llz_ctx_t ctx, const void* send, void* recv, size_t bytes_per_rank) {
  if (ctx->nranks == 1) {
    memcpy(recv, send, bytes_per_rank);
    return LLZ_OK;
  }
  void* d_send = nullptr;
  void* d_recv = nullptr;
  LLZ_CUDA(cudaMalloc(&d_send, std::max<size_t>(bytes_per_rank, 16)));
  LLZ_CUDA(cudaMalloc(&d_recv, std::max<size_t>(bytes_per_rank * ctx->nranks, 16)));
  LLZ_CUDA(cudaMemcpyAsync(d_send, send, bytes_per_rank, cudaMemcpyHostToDevice, ctx->stream));
  int s = comm_allgather_bytes(ctx, d_send, d_recv, bytes_per_rank);
  if (s == LLZ_OK) {
    cudaError_t e = cudaMemcpyAsync(recv, d_recv, bytes_per_rank * ctx->nranks, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) s = fail(LLZ_ERR_CUDA, "allgather_host: %s", cudaGetErrorString(e));
  }
  cudaFree(d_send);
  cudaFree(d_recv);
  return s;
}

// ---- peer-memory channels ------------------------------------------------------------------------------------
constexpr int kCoefPayload = 32768;  // doubles per slot of the coefficient channel (columns x NC + 1 must fit)
constexpr int kScalarPayload = 8;

static size_t channel_bytes(int G, int payload) { return ((size_t)2 * G * payload + (size_t)2 * G) * sizeof(double); }

// Collective over the group (called from llz_ctx_join): every rank allocates its inbox, exports it with CUDA IPC,
// all-gathers the handles over NCCL and maps every peer's inbox.  Any failure on any rank (IPC not permitted, no
// peer access) leaves the whole group on the NCCL path — decided together, so the ranks never disagree.
static int p2p_setup(llz_ctx_t ctx) {
  Comm* c = ctx->comm;
  const int G = ctx->nranks;
  const char* env = getenv("LLZ_P2P");
  int want = !(env && env[0] == '0');
  size_t win_bytes = (size_t)256 << 20;
  if (const char* w = getenv("LLZ_HALO_WINDOW_MB")) win_bytes = (size_t)std::max(0L, atol(w)) << 20;
  const size_t chan_bytes = (channel_bytes(G, kScalarPayload) * 4 + channel_bytes(G, kCoefPayload) + 256 + 255) / 256 * 256;
  const size_t bytes = chan_bytes + win_bytes;
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  int ok = want;
  if (ok && cudaMalloc(&c->p2p_local, bytes) != cudaSuccess) ok = 0;
  if (ok && cudaMemset(c->p2p_local, 0, chan_bytes) != cudaSuccess) ok = 0;
  if (ok && cudaIpcGetMemHandle(&mine, c->p2p_local) != cudaSuccess) ok = 0;
  cudaGetLastError();
  struct Rec {
    cudaIpcMemHandle_t h;
    int ok;
    int pad;
  };
  Rec rec;
  memset(&rec, 0, sizeof(rec));
  rec.h = mine;
  rec.ok = ok;
  std::vector<Rec> all((size_t)G);
  LLZ_TRY(comm_allgather_host(ctx, &rec, all.data(), sizeof(Rec)));
  for (int r = 0; r < G; ++r) ok = ok && all[(size_t)r].ok;
  if (ok) {
    for (int r = 0; r < G && ok; ++r) {
      if (r == ctx->rank) {
        c->p2p_peer[r] = c->p2p_local;
        continue;
      }
      if (cudaIpcOpenMemHandle(&c->p2p_peer[r], all[(size_t)r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        c->p2p_peer[r] = nullptr;
        ok = 0;
      }
    }
  }
  // second round: did every rank map every peer?
  int mapped = ok;
  std::vector<int> mapped_all((size_t)G);
  LLZ_TRY(comm_allgather_host(ctx, &mapped, mapped_all.data(), sizeof(int)));
  for (int r = 0; r < G; ++r) ok = ok && mapped_all[(size_t)r];
  if (!ok) {
    for (int r = 0; r < G; ++r)
      if (r != ctx->rank && c->p2p_peer[r]) cudaIpcCloseMemHandle(c->p2p_peer[r]);
    if (c->p2p_local) cudaFree(c->p2p_local);
    c->p2p_local = nullptr;
    memset(c->p2p_peer, 0, sizeof(c->p2p_peer));
    cudaGetLastError();
    return LLZ_OK;  // NCCL path
  }
  const int payload[5] = {kScalarPayload, kScalarPayload, kCoefPayload, kScalarPayload, kScalarPayload};
  size_t off = 256;  // first 256 bytes: status word (+0) and the last-CTA ticket (+64)
  c->win_off = chan_bytes;
  c->win_bytes = win_bytes;
  c->win_free.assign(1, std::make_pair((size_t)0, win_bytes));
  for (int k = 0; k < 5; ++k) {
    PeerChannel& ch = c->ch[k];
    ch.G = G;
    ch.rank = ctx->rank;
    ch.payload = payload[k];
    ch.status = reinterpret_cast<unsigned long long*>(c->p2p_local);
    for (int r = 0; r < G; ++r) ch.inbox[r] = reinterpret_cast<double*>(static_cast<char*>(c->p2p_peer[r]) + off);
    off += channel_bytes(G, payload[k]);
  }
  c->ticket = reinterpret_cast<unsigned int*>(static_cast<char*>(c->p2p_local) + 64);
  return LLZ_OK;
}

bool comm_p2p(llz_ctx_t ctx) { return ctx->nranks > 1 && ctx->comm && ctx->comm->ch[0].G > 0; }
int comm_coef_capacity(llz_ctx_t ctx) { return comm_p2p(ctx) ? kCoefPayload : 0; }

// Halo window of this rank: first-fit sub-allocation (256-byte granularity); returns the offset inside the window or
// -1 when there is no peer window or no room (the operator then uses the NCCL exchange).
int64_t comm_window_alloc(llz_ctx_t ctx, size_t bytes) {
  if (!comm_p2p(ctx) || bytes == 0) return -1;
  Comm* c = ctx->comm;
  bytes = (bytes + 255) / 256 * 256;
  for (size_t i = 0; i < c->win_free.size(); ++i) {
    if (c->win_free[i].second >= bytes) {
      const size_t off = c->win_free[i].first;
      c->win_free[i].first += bytes;
      c->win_free[i].second -= bytes;
      if (c->win_free[i].second == 0) c->win_free.erase(c->win_free.begin() + i);
      return (int64_t)off;
    }
  }
  return -1;
}
void comm_window_free(llz_ctx_t ctx, int64_t off, size_t bytes) {
  if (!ctx->comm || off < 0) return;
  Comm* c = ctx->comm;
  bytes = (bytes + 255) / 256 * 256;
  c->win_free.emplace_back((size_t)off, bytes);
  std::sort(c->win_free.begin(), c->win_free.end());
  for (size_t i = 0; i + 1 < c->win_free.size();) {  // coalesce neighbours
    if (c->win_free[i].first + c->win_free[i].second == c->win_free[i + 1].first) {
      c->win_free[i].second += c->win_free[i + 1].second;
      c->win_free.erase(c->win_free.begin() + i + 1);
    } else {
      ++i;
    }
  }
}
// Address, valid on THIS GPU, of byte `off` of rank r's halo window.
void* comm_window_ptr(llz_ctx_t ctx, int r, int64_t off) {
  Comm* c = ctx->comm;
  return static_cast<char*>(c->p2p_peer[r]) + c->win_off + off;
}

// A buffer of `bytes` on every rank, each mapped into every peer (CUDA IPC) — where producers store the vector they
// write so that the next operator apply finds the whole input vector without an all-gather.  Collective (all ranks
// call it in the same order); buffers are recycled by size and only freed with the context, so that releasing one is
// a local operation.  nullptr when the group has no peer channels or any rank failed to allocate / map.
ExchangeBuffer* comm_exchange_buffer_acquire(llz_ctx_t ctx, size_t bytes) {
  if (!comm_p2p(ctx) || bytes == 0) return nullptr;
  Comm* c = ctx->comm;
  for (ExchangeBuffer* b : c->xbufs)
    if (b->usable && !b->in_use && b->bytes == bytes) {
      b->in_use = true;
      return b;
    }
  const int G = ctx->nranks;
  ExchangeBuffer* b = new ExchangeBuffer();
  b->bytes = bytes;
  struct Rec {
    cudaIpcMemHandle_t h;
    int ok;
    int pad;
  } rec;
  memset(&rec, 0, sizeof(rec));
  rec.ok = dev_malloc(ctx, &b->local, bytes) == cudaSuccess;
  if (rec.ok && cudaIpcGetMemHandle(&rec.h, b->local) != cudaSuccess) rec.ok = 0;
  cudaGetLastError();
  std::vector<Rec> all((size_t)G);
  int ok = comm_allgather_host(ctx, &rec, all.data(), sizeof(Rec)) == LLZ_OK;
  for (int r = 0; r < G && ok; ++r) ok = all[(size_t)r].ok;
  for (int r = 0; r < G && ok; ++r) {
    if (r == ctx->rank) {
      b->peer[r] = b->local;
    } else if (cudaIpcOpenMemHandle(&b->peer[r], all[(size_t)r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      b->peer[r] = nullptr;
      ok = 0;
    }
  }
  int mapped = ok;
  std::vector<int> mapped_all((size_t)G, 0);
  if (comm_allgather_host(ctx, &mapped, mapped_all.data(), sizeof(int)) != LLZ_OK) ok = 0;
  for (int r = 0; r < G; ++r) ok = ok && mapped_all[(size_t)r];
  b->in_use = ok;
  b->usable = ok;
  c->xbufs.push_back(b);  // kept even when unusable: its memory and mappings are released with the context
  return ok ? b : nullptr;
}
void comm_exchange_buffer_release(llz_ctx_t ctx, ExchangeBuffer* b) {
  (void)ctx;
  if (b) b->in_use = false;
}

// The NEXT message of a channel (every rank calls this in the same order); an unused message (ch.G == 0) when the
// group has no peer channels.
PeerMsg comm_next_message(llz_ctx_t ctx, int which) {
  PeerMsg m;
  if (!comm_p2p(ctx)) return m;
  Comm* c = ctx->comm;
  m.ch = c->ch[which];
  m.seq = ++c->seq[which];
  m.ticket = c->ticket;
  return m;
}

// Non-zero when a kernel gave up waiting for a peer (the peer process died or diverged).
int comm_check_peers(llz_ctx_t ctx) {
  if (!comm_p2p(ctx)) return LLZ_OK;
  unsigned long long st = 0;
  if (cudaMemcpy(&st, ctx->comm->p2p_local, sizeof(st), cudaMemcpyDeviceToHost) != cudaSuccess) return fail(LLZ_ERR_CUDA, "peer status read failed");
  if (st) return fail(LLZ_ERR_COMM, "a kernel timed out waiting for message %llu of a peer GPU", st & ~(1ull << 63));
  return LLZ_OK;
}

void comm_destroy(llz_ctx_t ctx) {
  if (ctx->comm && ctx->comm->p2p_local) {
    Comm* c = ctx->comm;
    for (int r = 0; r < ctx->nranks; ++r)
      if (r != ctx->rank && c->p2p_peer[r]) cudaIpcCloseMemHandle(c->p2p_peer[r]);
    for (ExchangeBuffer* b : c->xbufs)
      for (int r = 0; r < ctx->nranks; ++r)
        if (r != ctx->rank && b->peer[r]) cudaIpcCloseMemHandle(b->peer[r]);
    // every rank must have unmapped this rank's inbox before it is freed: one blocking collective
    double* d = ctx->d_result;
    if (d && nccl().AllReduce(d, d, 1, ncclDouble, ncclSum, c->nccl, ctx->stream) == ncclSuccess) cudaStreamSynchronize(ctx->stream);
    cudaFree(c->p2p_local);
    c->p2p_local = nullptr;
    for (ExchangeBuffer* b : c->xbufs) {
      if (b->local) dev_free(ctx, b->local);
      delete b;
    }
    c->xbufs.clear();
    cudaGetLastError();
  }
  if (ctx->comm) {
    if (ctx->comm->nccl) nccl().CommDestroy(ctx->comm->nccl);
    delete ctx->comm;
    ctx->comm = nullptr;
  }
}

}  // namespace llz

using namespace llz;

extern "C" {

int llz_ctx_peer_channels(llz_ctx_t ctx, int* enabled) {
  if (!ctx || !enabled) return fail(LLZ_ERR_INVALID, "null argument");
  *enabled = comm_p2p(ctx) ? 1 : 0;
  return LLZ_OK;
}

int llz_comm_unique_id(void* id128) {
  if (!id128) return fail(LLZ_ERR_INVALID, "null id buffer");
  if (!nccl().ok) return fail(LLZ_ERR_COMM, "NCCL unavailable: %s", nccl().why);
  static_assert(sizeof(ncclUniqueId) <= 128, "ncclUniqueId larger than the 128-byte blob of the ABI");
  ncclUniqueId id;
  LLZ_NCCL(nccl().GetUniqueId(&id));
  memset(id128, 0, 128);
  memcpy(id128, &id, sizeof(id));
  return LLZ_OK;
}

int llz_ctx_join(llz_ctx_t ctx, int rank, int nranks, const void* id128) {
  if (!ctx || nranks < 1 || rank < 0 || rank >= nranks) return fail(LLZ_ERR_INVALID, "ctx_join: bad rank %d of %d", rank, nranks);
  if (ctx->comm) return fail(LLZ_ERR_INVALID, "ctx_join: context already joined");
  if (nranks == 1) return LLZ_OK;
  if (!id128) return fail(LLZ_ERR_INVALID, "ctx_join: null id");
  if (nranks > kMaxRanks) return fail(LLZ_ERR_UNSUPPORTED, "ctx_join: at most %d ranks (one box)", kMaxRanks);
  if (!nccl().ok) return fail(LLZ_ERR_COMM, "NCCL unavailable: %s", nccl().why);
  LLZ_CUDA(cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  Comm* c = new Comm();
  ncclResult_t r = nccl().CommInitRank(&c->nccl, nranks, id, rank);
  if (r != ncclSuccess) {
    delete c;
    return fail(LLZ_ERR_COMM, "ncclCommInitRank: %s", nccl().GetErrorString(r));
  }
  ctx->comm = c;
  ctx->rank = rank;
  ctx->nranks = nranks;
  return p2p_setup(ctx);
}

}  // extern "C"
