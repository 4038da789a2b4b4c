// llz_comm.cu — inter-GPU plumbing for row-sharded runs: one process per GPU, one NCCL communicator per context,
// all collectives enqueued on the context's stream (NVLink 5 / NVSwitch underneath).  Single-rank contexts never
// reach NCCL.
//
// Per Lanczos iteration a joined context issues: one halo exchange before the SpMV (only the remote entries the local
// rows reference, ncclSend/ncclRecv grouped), one all-reduce of the packed projection coefficients, and all-reduces of
// the alpha / beta^2 scalars.
#include <nccl.h>

#include <algorithm>
#include <cstring>

#include "llz_launch.hpp"

namespace llz {

struct Comm {
  ncclComm_t nccl = nullptr;
};

#define LLZ_NCCL(expr)                                                                                       \
  do {                                                                                                       \
    ncclResult_t r__ = (expr);                                                                               \
    if (r__ != ncclSuccess) return ::llz::fail(LLZ_ERR_COMM, "%s: %s", #expr, ncclGetErrorString(r__));      \
  } while (0)

int comm_allreduce_sum(llz_ctx_t ctx, double* d, int count) {
  if (ctx->nranks == 1 || count <= 0) return LLZ_OK;
  LLZ_NCCL(ncclAllReduce(d, d, (size_t)count, ncclDouble, ncclSum, ctx->comm->nccl, ctx->stream));
  return LLZ_OK;
}

int comm_allreduce_partials(llz_ctx_t ctx, double* d, int* count) {
  if (ctx->nranks == 1) return LLZ_OK;
  // fold this rank's per-CTA partials into d[0] (fixed order), then sum over the group
  LLZ_TRY(launch_sum_partials(ctx, d, *count, 1, d, nullptr));
  LLZ_NCCL(ncclAllReduce(d, d, 1, ncclDouble, ncclSum, ctx->comm->nccl, ctx->stream));
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
  LLZ_NCCL(ncclAllGather(send, recv, bytes_per_rank, ncclChar, ctx->comm->nccl, ctx->stream));
  return LLZ_OK;
}

// Grouped point-to-point exchange: for every peer p, send send_bytes[p] from send_base + send_off[p] and receive
// recv_bytes[p] into recv_base + recv_off[p].
int comm_exchange(llz_ctx_t ctx, const char* send_base, const size_t* send_off, const size_t* send_bytes,
                  char* recv_base, const size_t* recv_off, const size_t* recv_bytes) {
  if (ctx->nranks == 1) return LLZ_OK;
  LLZ_NCCL(ncclGroupStart());
  for (int p = 0; p < ctx->nranks; ++p) {
    if (p == ctx->rank) continue;
    if (send_bytes[p]) LLZ_NCCL(ncclSend(send_base + send_off[p], send_bytes[p], ncclChar, p, ctx->comm->nccl, ctx->stream));
    if (recv_bytes[p]) LLZ_NCCL(ncclRecv(recv_base + recv_off[p], recv_bytes[p], ncclChar, p, ctx->comm->nccl, ctx->stream));
  }
  LLZ_NCCL(ncclGroupEnd());
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

void comm_destroy(llz_ctx_t ctx) {
  if (ctx->comm) {
    if (ctx->comm->nccl) ncclCommDestroy(ctx->comm->nccl);
    delete ctx->comm;
    ctx->comm = nullptr;
  }
}

}  // namespace llz

using namespace llz;

extern "C" {

int llz_comm_unique_id(void* id128) {
  if (!id128) return fail(LLZ_ERR_INVALID, "null id buffer");
  static_assert(sizeof(ncclUniqueId) <= 128, "ncclUniqueId larger than the 128-byte blob of the ABI");
  ncclUniqueId id;
  LLZ_NCCL(ncclGetUniqueId(&id));
  memset(id128, 0, 128);
  memcpy(id128, &id, sizeof(id));
  return LLZ_OK;
}

int llz_ctx_join(llz_ctx_t ctx, int rank, int nranks, const void* id128) {
  if (!ctx || nranks < 1 || rank < 0 || rank >= nranks) return fail(LLZ_ERR_INVALID, "ctx_join: bad rank %d of %d", rank, nranks);
  if (ctx->comm) return fail(LLZ_ERR_INVALID, "ctx_join: context already joined");
  if (nranks == 1) return LLZ_OK;
  if (!id128) return fail(LLZ_ERR_INVALID, "ctx_join: null id");
  LLZ_CUDA(cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  Comm* c = new Comm();
  ncclResult_t r = ncclCommInitRank(&c->nccl, nranks, id, rank);
  if (r != ncclSuccess) {
    delete c;
    return fail(LLZ_ERR_COMM, "ncclCommInitRank: %s", ncclGetErrorString(r));
  }
  ctx->comm = c;
  ctx->rank = rank;
  ctx->nranks = nranks;
  return LLZ_OK;
}

}  // extern "C"
