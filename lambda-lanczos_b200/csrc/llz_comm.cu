// llz_comm.cu — inter-GPU plumbing for row-sharded runs (one process per GPU).  Single-rank contexts never touch it.
#include "llz_launch.hpp"

namespace llz {

struct Comm {
  int placeholder = 0;
};

int comm_allreduce_sum(llz_ctx_t ctx, double* d, int count) {
  (void)d;
  (void)count;
  if (ctx->nranks == 1) return LLZ_OK;
  return fail(LLZ_ERR_UNSUPPORTED, "multi-rank reductions are not built yet");
}

int comm_allreduce_partials(llz_ctx_t ctx, double* d, int* count) {
  (void)d;
  (void)count;
  if (ctx->nranks == 1) return LLZ_OK;
  return fail(LLZ_ERR_UNSUPPORTED, "multi-rank reductions are not built yet");
}

void comm_destroy(llz_ctx_t ctx) {
  delete ctx->comm;
  ctx->comm = nullptr;
}

}  // namespace llz

extern "C" {

int llz_comm_unique_id(void* id128) {
  (void)id128;
  return llz::fail(LLZ_ERR_UNSUPPORTED, "multi-rank support is not built yet");
}

int llz_ctx_join(llz_ctx_t ctx, int rank, int nranks, const void* id128) {
  (void)id128;
  if (!ctx) return llz::fail(LLZ_ERR_INVALID, "null ctx");
  if (nranks == 1 && rank == 0) return LLZ_OK;
  return llz::fail(LLZ_ERR_UNSUPPORTED, "multi-rank support is not built yet");
}

}  // extern "C"
