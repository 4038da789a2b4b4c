// llz_krylov.cu — the device-resident Krylov workspace: basis store (CUDA virtual-memory backed, grows without copies),
// the alpha/beta scalar bank with its pinned host mirror, and the per-iteration kernel sequence.
//
// Replaces, on the device, the storage and loop body of LambdaLanczos<T>::run_iteration
// (lambda_lanczos.hpp:221-223 for `u`, `alpha`, `beta`; :240-285 for the iteration) and of Exponentiator<T>::run
// (exponentiator.hpp:90-92, :106-160).  The host keeps the control flow (convergence tests, tridiagonal solves).
#include <cuda.h>

#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>

#include "llz_launch.hpp"

namespace llz {

// ---- CUDA virtual memory management through runtime-resolved driver entry points (no libcuda link dependency) ----
struct VmmApi {
  bool ok = false;
  CUresult (*AddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*AddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*Create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*Release)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*Map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*Unmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*SetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*GetGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
};

static VmmApi& vmm() {
  static VmmApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  const char* env = getenv("LLZ_BASIS_VMM");
  if (env && env[0] == '0') return api;
  auto get = [](const char* name, void** fn) {
    cudaDriverEntryPointQueryResult q;
    return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && *fn;
  };
  bool ok = get("cuMemAddressReserve", (void**)&api.AddressReserve) && get("cuMemAddressFree", (void**)&api.AddressFree) &&
            get("cuMemCreate", (void**)&api.Create) && get("cuMemRelease", (void**)&api.Release) &&
            get("cuMemMap", (void**)&api.Map) && get("cuMemUnmap", (void**)&api.Unmap) &&
            get("cuMemSetAccess", (void**)&api.SetAccess) &&
            get("cuMemGetAllocationGranularity", (void**)&api.GetGranularity);
  api.ok = ok;
  return api;
}

}  // namespace llz

using namespace llz;

struct llz_krylov_s {
  llz_ctx_t ctx = nullptr;
  int dtype = 0;
  int64_t n = 0;
  int64_t ld = 0;        // elements per column slot
  size_t col_bytes = 0;  // ld * sizeof(T)
  int64_t cap_cols = 0;  // columns the store may ever hold
  int64_t requested_cols = 0;  // max_cols the creator asked for (cache key)
  // basis store
  bool use_vmm = false;
  CUdeviceptr va = 0;
  size_t va_size = 0;
  size_t chunk_bytes = 0;
  size_t mapped_bytes = 0;
  std::vector<CUmemGenericAllocationHandle> handles;
  void* plain = nullptr;  // fallback: one cudaMalloc
  // scalar bank
  double* d_alpha = nullptr;
  double* d_beta = nullptr;
  double* d_pa = nullptr;  // alpha partials
  double* d_pb = nullptr;  // norm partials (kMaxCombine * kMaxGrid)
  double* d_ph = nullptr;  // projection partials
  size_t ph_cap = 0;
  double* d_coef = nullptr;
  size_t coef_cap = 0;
  void* d_ycoef = nullptr;  // combine coefficients (device, T)
  size_t ycoef_cap = 0;
  double* d_misc = nullptr;
  unsigned int* d_ticket = nullptr;  // last-CTA detection of the lazy recurrence kernel
  double* h_alpha = nullptr;  // pinned + mapped
  double* h_beta = nullptr;
  double* h_misc = nullptr;
  double* h_wnorm = nullptr;  // ||w'|| before the Gram-Schmidt pass of each iteration (DGKS cancellation test)
  long long* h_flag = nullptr;
  int64_t scalar_cap = 0;
  // locked vectors
  const void** d_qptrs = nullptr;
  int q_cap = 0;
  int nq = 0;
  // fused all-gather: the vector of column `pushed_col` was also stored into the peers' exchange buffers by the
  // kernel that produced it, for operator `pushed_op`
  bool pushed_valid = false;
  int64_t pushed_col = -1;
  OpBase* pushed_op = nullptr;
  GatherPush pushed;
  bool halo_pushed = false;  // ... or, for sparse operators, its halo entries into the peers' halo segments
  HaloPushPlan halo_plan;
  // state
  int64_t k = 0;

  char* col(int64_t j) const { return (use_vmm ? (char*)va : (char*)plain) + (size_t)j * col_bytes; }
};

namespace {

int ensure_cols(llz_krylov_t kry, int64_t cols) {
  if (cols > kry->cap_cols)
    return fail(LLZ_ERR_OOM, "Krylov basis full: %lld columns of %zu bytes is all that fits (device memory / max_cols)",
                (long long)kry->cap_cols, kry->col_bytes);
  if (!kry->use_vmm) return LLZ_OK;
  const size_t need = (size_t)cols * kry->col_bytes;
  VmmApi& api = vmm();
  while (kry->mapped_bytes < need) {
    size_t sz = kry->chunk_bytes;
    if (kry->mapped_bytes + sz > kry->va_size) sz = kry->va_size - kry->mapped_bytes;
    CUmemAllocationProp prop;
    memset(&prop, 0, sizeof(prop));
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = kry->ctx->device;
    CUmemGenericAllocationHandle h;
    CUresult r = api.Create(&h, sz, &prop, 0);
    if (r != CUDA_SUCCESS)
      return fail(r == CUDA_ERROR_OUT_OF_MEMORY ? LLZ_ERR_OOM : LLZ_ERR_CUDA,
                  "cuMemCreate(%zu) failed (%d) growing the Krylov basis past %zu bytes", sz, (int)r, kry->mapped_bytes);
    r = api.Map(kry->va + kry->mapped_bytes, sz, 0, h, 0);
    if (r != CUDA_SUCCESS) {
      api.Release(h);
      return fail(LLZ_ERR_CUDA, "cuMemMap failed (%d)", (int)r);
    }
    CUmemAccessDesc acc;
    memset(&acc, 0, sizeof(acc));
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = kry->ctx->device;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    r = api.SetAccess(kry->va + kry->mapped_bytes, sz, &acc, 1);
    if (r != CUDA_SUCCESS) {
      api.Unmap(kry->va + kry->mapped_bytes, sz);
      api.Release(h);
      return fail(LLZ_ERR_CUDA, "cuMemSetAccess failed (%d)", (int)r);
    }
    kry->handles.push_back(h);
    kry->mapped_bytes += sz;
  }
  return LLZ_OK;
}

int ensure_ph(llz_krylov_t kry, size_t doubles) {
  if (doubles <= kry->ph_cap) return LLZ_OK;
  if (kry->d_ph) {
    LLZ_CUDA(cudaStreamSynchronize(kry->ctx->stream));
    cudaFree(kry->d_ph);
    kry->d_ph = nullptr;
  }
  size_t cap = doubles * 2;
  LLZ_CUDA(cudaMalloc(&kry->d_ph, cap * sizeof(double)));
  kry->ph_cap = cap;
  return LLZ_OK;
}

// NCCL path only (no peer channels): fold the per-CTA partials of a group-wide scalar and all-reduce them.
int allreduce_scalar(llz_ctx_t ctx, double* partials, int* count) {
  if (ctx->nranks == 1 || comm_p2p(ctx)) return LLZ_OK;
  ProfScope ps(ctx, "exchange", 0.0);
  return comm_allreduce_partials(ctx, partials, count);
}

// One classical Gram-Schmidt pass of `w` against cs: project -> reduce -> update (in place).  With an empty column
// set only the update kernel runs (it then just produces the norm partials of w).  `wnorm` tells the caller where
// ||w'||^2 of this pass can be read (peer message element or kry->d_misc[1]).
int cgs_pass(llz_krylov_t kry, const ColumnSet& cs, void* w, const Fold& fold, bool want_norm, int* norm_grid) {
  // (row-sharded with peer channels: fold.norm_msg, if used, is delivered by the update kernel that writes the norm
  //  partials; the coefficients travel as one message of the coefficient channel)
  llz_ctx_t ctx = kry->ctx;
  const int nc = dtype_nc(kry->dtype);
  const int total = cs.ncols();
  const int max_grid = std::min(kMaxGrid, ctx->num_sms * 2);
  PeerMsg coef_msg;  // row-sharded with peer channels: the coefficients travel as one message
  bool peer = false;
  if (total > 0) {
    peer = comm_p2p(ctx) && total * nc + 1 <= comm_coef_capacity(ctx);
    if (peer) coef_msg = comm_next_message(ctx, kChanCoef);
    const int pchunk = max_project_cols(kry->dtype);
    for (int c0 = 0; c0 < total; c0 += pchunk) {
      const int cols = std::min(pchunk, total - c0);
      const bool last = c0 + cols >= total;
      LLZ_TRY(ensure_ph(kry, (size_t)max_grid * ((size_t)cols * nc + 1)));
      int grid = 0;
      {
        ProfScope ps(ctx, "project", (double)kry->n * (double)dtype_size(kry->dtype) * (cols + 1 + fold.mode));
        LLZ_TRY(launch_project(ctx, kry->dtype, cs, c0, cols, w, kry->n, fold, kry->d_ph, &grid));
      }
      {
        ProfScope ps(ctx, "reduce", 0.0);
        LLZ_TRY(launch_reduce(ctx, kry->dtype, kry->d_ph, grid, c0, cols, kry->d_coef, last ? kry->d_misc + 1 : nullptr, coef_msg,
                              last ? total * nc : -1, last ? 1 : 0));
      }
    }
    if (!peer && ctx->nranks > 1) {
      ProfScope ps(ctx, "exchange", 0.0);
      LLZ_TRY(comm_allreduce_sum(ctx, kry->d_coef, total * nc));
      LLZ_TRY(comm_allreduce_sum(ctx, kry->d_misc + 1, 1));
    }
  }
  // update: the fold.mode trailing basis columns are consumed by the kernel's recurrence prologue (first chunk)
  const int generic = total - fold.mode;
  const int uchunk = max_update_cols(kry->dtype);
  int c0 = 0;
  do {
    const int cols = std::min(uchunk, generic - c0);
    const bool last = c0 + cols >= generic;
    Fold f = (c0 == 0) ? fold : Fold();
    if (c0 == 0 && peer) {  // ||w'||^2 travelled with the coefficients: keep a local copy (kry->d_misc[1], as on one rank)
      f.wnorm_out = kry->d_misc + 1;
      f.wnorm_index = total * nc;
    }
    f.norm_msg = fold.norm_msg;
    f.push = fold.push;  // (launch_update only honours it in the chunk that writes the final vector and its norm)
    ProfScope ps(ctx, "update", (double)kry->n * (double)dtype_size(kry->dtype) * (cols + f.mode + 2));
    LLZ_TRY(launch_update(ctx, kry->dtype, cs, c0, cols, w, w, kry->n, kry->d_coef, f,
                          (last && want_norm) ? kry->d_pb : nullptr, norm_grid, coef_msg, fold.norm_msg));
    c0 += cols;
  } while (c0 < generic);
  return LLZ_OK;
}

}  // namespace

static int krylov_init(llz_krylov_t kry, int dtype, int64_t n, int64_t max_cols);

extern "C" {

int llz_krylov_create(llz_ctx_t ctx, int dtype, int64_t n, int64_t max_cols, llz_krylov_t* out) {
  if (!ctx || !out || n < 1 || max_cols < 2 || dtype_size(dtype) == 0)
    return fail(LLZ_ERR_INVALID, "krylov_create: bad argument (n=%lld, max_cols=%lld)", (long long)n, (long long)max_cols);
  LLZ_CUDA(cudaSetDevice(ctx->device));
  if (ctx->cached_krylov) {
    llz_krylov_t c = ctx->cached_krylov;
    ctx->cached_krylov = nullptr;
    if (c->dtype == dtype && c->n == n && c->requested_cols == max_cols) {  // revive: mapped basis memory is kept
      c->k = 0;
      c->nq = 0;
      c->pushed_valid = false;
      c->halo_pushed = false;
      c->h_flag[0] = 0;
      *out = c;
      return LLZ_OK;
    }
    krylov_destroy_now(c);
  }
  llz_krylov_t kry = new llz_krylov_s();
  kry->ctx = ctx;
  // every early return below (LLZ_CUDA / fail) must give back the object, its VA reservation and what was allocated
  const int st = krylov_init(kry, dtype, n, max_cols);
  if (st != LLZ_OK) {
    krylov_destroy_now(kry);
    return st;
  }
  *out = kry;
  return LLZ_OK;
}

}  // extern "C"

static int krylov_init(llz_krylov_t kry, int dtype, int64_t n, int64_t max_cols) {
  llz_ctx_t ctx = kry->ctx;
  const size_t es = dtype_size(dtype);
  kry->requested_cols = max_cols;
  kry->dtype = dtype;
  kry->n = n;
  // column slots: 256-byte aligned, plus a 1280-byte skew so that consecutive columns of a power-of-two sized
  // problem do not land on the same L2 slice / HBM channel for equal row offsets
  size_t bytes = ((size_t)n * es + 255) / 256 * 256;
  if (bytes >= (1u << 20)) bytes += 1280;
  kry->col_bytes = bytes;
  kry->ld = (int64_t)(bytes / es);

  size_t free_b = 0, total_b = 0;
  LLZ_CUDA(cudaMemGetInfo(&free_b, &total_b));
  int64_t fit = (int64_t)((double)free_b * 0.96 / (double)bytes);
  if (fit < 2) return fail(LLZ_ERR_OOM, "not even two Lanczos vectors of %zu bytes fit in %zu free bytes", bytes, free_b);
  kry->cap_cols = std::min<int64_t>(max_cols, fit);

  VmmApi& api = vmm();
  if (api.ok) {
    CUmemAllocationProp prop;
    memset(&prop, 0, sizeof(prop));
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = ctx->device;
    size_t gran = 0;
    if (api.GetGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS && gran > 0) {
      // physical chunk: >= 64 MiB and >= 4 columns, multiple of the granularity
      size_t chunk = std::max<size_t>((size_t)64 << 20, 4 * bytes);
      chunk = (chunk + gran - 1) / gran * gran;
      size_t va_size = ((size_t)kry->cap_cols * bytes + chunk - 1) / chunk * chunk;
      CUdeviceptr va = 0;
      if (api.AddressReserve(&va, va_size, gran, 0, 0) == CUDA_SUCCESS) {
        kry->use_vmm = true;
        kry->va = va;
        kry->va_size = va_size;
        kry->chunk_bytes = chunk;
      }
    }
  }
  if (!kry->use_vmm) {
    // fallback: a single allocation of the whole capacity (bounded so that small problems do not grab the GPU)
    int64_t cols = std::min<int64_t>(kry->cap_cols, std::max<int64_t>(2, (int64_t)(((size_t)8 << 30) / bytes)));
    cudaError_t e = cudaMalloc(&kry->plain, (size_t)cols * bytes);
    if (e != cudaSuccess)
      return fail(LLZ_ERR_OOM, "cudaMalloc of the Krylov basis (%lld columns) failed: %s", (long long)cols, cudaGetErrorString(e));
    kry->cap_cols = cols;
  }

  kry->scalar_cap = kry->cap_cols + 2;
  const size_t sc = (size_t)kry->scalar_cap;
  LLZ_CUDA(cudaMalloc(&kry->d_alpha, sc * sizeof(double)));
  LLZ_CUDA(cudaMalloc(&kry->d_beta, sc * sizeof(double)));
  LLZ_CUDA(cudaMemsetAsync(kry->d_alpha, 0, sc * sizeof(double), ctx->stream));
  LLZ_CUDA(cudaMemsetAsync(kry->d_beta, 0, sc * sizeof(double), ctx->stream));
  LLZ_CUDA(cudaMalloc(&kry->d_pa, kMaxGrid * sizeof(double)));
  LLZ_CUDA(cudaMalloc(&kry->d_pb, (size_t)kMaxGrid * 5 * sizeof(double)));
  LLZ_CUDA(cudaMalloc(&kry->d_misc, 8 * sizeof(double)));
  LLZ_CUDA(cudaMalloc(&kry->d_ticket, sizeof(unsigned int)));
  LLZ_CUDA(cudaMemsetAsync(kry->d_ticket, 0, sizeof(unsigned int), ctx->stream));
  kry->coef_cap = (sc + 64) * 2;
  LLZ_CUDA(cudaMalloc(&kry->d_coef, kry->coef_cap * sizeof(double)));
  LLZ_CUDA(cudaHostAlloc(&kry->h_alpha, sc * sizeof(double), cudaHostAllocMapped));
  LLZ_CUDA(cudaHostAlloc(&kry->h_beta, sc * sizeof(double), cudaHostAllocMapped));
  LLZ_CUDA(cudaHostAlloc(&kry->h_misc, 8 * sizeof(double), cudaHostAllocMapped));
  LLZ_CUDA(cudaHostAlloc(&kry->h_wnorm, sc * sizeof(double), cudaHostAllocMapped));
  LLZ_CUDA(cudaHostAlloc(&kry->h_flag, sizeof(long long) * 2, cudaHostAllocMapped));
  kry->h_flag[0] = 0;
  return ensure_cols(kry, 2);
}

extern "C" {

int llz_krylov_destroy(llz_krylov_t kry) {
  if (!kry) return LLZ_OK;
  llz_ctx_t ctx = kry->ctx;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->cached_krylov) krylov_destroy_now(ctx->cached_krylov);
  ctx->cached_krylov = kry;  // keep the mapped basis for the next run on this context (llz_ctx_release_cache frees it)
  return LLZ_OK;
}

}  // extern "C"

void llz::krylov_destroy_now(llz_krylov_t kry) {
  if (!kry) return;
  cudaSetDevice(kry->ctx->device);
  cudaStreamSynchronize(kry->ctx->stream);
  if (kry->use_vmm && kry->va) {
    VmmApi& api = vmm();
    size_t off = 0;
    for (auto h : kry->handles) {
      size_t sz = std::min(kry->chunk_bytes, kry->va_size - off);
      api.Unmap(kry->va + off, sz);
      api.Release(h);
      off += sz;
    }
    api.AddressFree(kry->va, kry->va_size);
  } else if (kry->plain) {
    cudaFree(kry->plain);
  }
  cudaFree(kry->d_alpha);
  cudaFree(kry->d_beta);
  cudaFree(kry->d_pa);
  cudaFree(kry->d_pb);
  cudaFree(kry->d_misc);
  if (kry->d_ticket) cudaFree(kry->d_ticket);
  cudaFree(kry->d_coef);
  if (kry->d_ph) cudaFree(kry->d_ph);
  if (kry->d_ycoef) cudaFree(kry->d_ycoef);
  if (kry->d_qptrs) cudaFree(kry->d_qptrs);
  for (void* h : {(void*)kry->h_alpha, (void*)kry->h_beta, (void*)kry->h_misc, (void*)kry->h_wnorm, (void*)kry->h_flag})
    if (h) cudaFreeHost(h);  // (a workspace whose creation failed half-way holds nulls)
  delete kry;
}

extern "C" {

int llz_krylov_capacity(llz_krylov_t kry, int64_t* max_cols) {
  if (!kry || !max_cols) return fail(LLZ_ERR_INVALID, "null");
  *max_cols = kry->cap_cols;
  return LLZ_OK;
}

int llz_krylov_set_locked(llz_krylov_t kry, const llz_vec_t* locked, int64_t count) {
  if (!kry || count < 0 || (count > 0 && !locked)) return fail(LLZ_ERR_INVALID, "set_locked: bad argument");
  if (count > kry->q_cap) {
    LLZ_CUDA(cudaStreamSynchronize(kry->ctx->stream));
    if (kry->d_qptrs) cudaFree(kry->d_qptrs);
    kry->d_qptrs = nullptr;
    int cap = (int)std::max<int64_t>(16, count * 2);
    LLZ_CUDA(cudaMalloc(&kry->d_qptrs, sizeof(void*) * cap));
    kry->q_cap = cap;
  }
  if ((size_t)(kry->scalar_cap + count) * 2 > kry->coef_cap) {
    LLZ_CUDA(cudaStreamSynchronize(kry->ctx->stream));
    cudaFree(kry->d_coef);
    kry->coef_cap = (size_t)(kry->scalar_cap + count + 64) * 2;
    LLZ_CUDA(cudaMalloc(&kry->d_coef, kry->coef_cap * sizeof(double)));
  }
  std::vector<const void*> ptrs((size_t)count);
  for (int64_t j = 0; j < count; ++j) {
    if (!locked[j] || locked[j]->n != kry->n || locked[j]->dtype != kry->dtype)
      return fail(LLZ_ERR_INVALID, "set_locked: vector %lld mismatched", (long long)j);
    ptrs[(size_t)j] = locked[j]->d;
  }
  if (count > 0) {
    LLZ_CUDA(cudaMemcpyAsync(kry->d_qptrs, ptrs.data(), sizeof(void*) * count, cudaMemcpyHostToDevice, kry->ctx->stream));
    LLZ_CUDA(cudaStreamSynchronize(kry->ctx->stream));
  }
  kry->nq = (int)count;
  return LLZ_OK;
}

int llz_krylov_begin(llz_krylov_t kry, const void* start, int host, double* norm_out) {
  if (!kry || !start) return fail(LLZ_ERR_INVALID, "krylov_begin: null");
  llz_ctx_t ctx = kry->ctx;
  LLZ_CUDA(cudaSetDevice(ctx->device));
  LLZ_CUDA(cudaStreamSynchronize(ctx->stream));  // nothing of a previous run may still publish scalars
  if (ctx->nranks > 1) {
    // Row-sharded: line the ranks up on the host before the first kernel that waits for a peer's message, so that a
    // rank that arrives late (still building its operator, reading input ...) makes the others wait here, in a
    // collective, and not inside a spinning kernel with a time-out.
    LLZ_CUDA(cudaMemsetAsync(kry->d_misc + 2, 0, sizeof(double), ctx->stream));
    LLZ_TRY(comm_allreduce_sum(ctx, kry->d_misc + 2, 1));
    LLZ_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  kry->k = 0;
  kry->pushed_valid = false;
  kry->halo_pushed = false;
  kry->h_flag[0] = 0;
  const size_t bytes = (size_t)kry->n * dtype_size(kry->dtype);
  void* u0 = kry->col(0);
  if (start != u0)
    LLZ_CUDA(cudaMemcpyAsync(u0, start, bytes, host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, ctx->stream));
  ColumnSet cs;
  cs.V = kry->col(0);
  cs.ld = kry->ld;
  cs.nv = 0;
  cs.Q = (const void* const*)kry->d_qptrs;
  cs.nq = kry->nq;
  Fold nofold;
  int grid = 0;
  // the start vector may lie almost inside span(locked): orthogonalise twice ("twice is enough")
  if (kry->nq > 0) LLZ_TRY(cgs_pass(kry, cs, u0, nofold, false, &grid));
  ScalarSink sink;
  nofold.norm_msg = sink.beta_msg = comm_next_message(ctx, kChanBeta);
  LLZ_TRY(cgs_pass(kry, cs, u0, nofold, true, &grid));
  LLZ_TRY(allreduce_scalar(ctx, kry->d_pb, &grid));
  sink.beta_out = kry->d_misc;
  sink.h_beta = kry->h_misc;
  {
    ProfScope ps(ctx, "scale", (double)kry->n * (double)dtype_size(kry->dtype) * 2);
    LLZ_TRY(launch_scale_by_norm(ctx, kry->dtype, u0, kry->n, kry->d_pb, grid, sink));
  }
  LLZ_CUDA(cudaStreamSynchronize(ctx->stream));
  LLZ_TRY(comm_check_peers(ctx));
  if (norm_out) *norm_out = kry->h_misc[0];
  return LLZ_OK;
}

int llz_krylov_step(llz_krylov_t kry, llz_op_t op, double sigma, int orth) {
  if (!kry || !op || !op->impl) return fail(LLZ_ERR_INVALID, "krylov_step: null");
  if (op->impl->n_local != kry->n || op->impl->dtype != kry->dtype)
    return fail(LLZ_ERR_INVALID, "krylov_step: operator (n=%lld, dtype %d) does not match workspace (n=%lld, dtype %d)",
                (long long)op->impl->n_local, op->impl->dtype, (long long)kry->n, kry->dtype);
  llz_ctx_t ctx = kry->ctx;
  const int64_t k = kry->k + 1;
  LLZ_TRY(ensure_cols(kry, k + 1));
  void* x = kry->col(k - 1);
  void* y = kry->col(k);

  int npa = 0;
  if (kry->pushed_valid && kry->pushed_op == op->impl && kry->pushed_col == k - 1) {
    // the previous iteration's update / recurrence kernel already delivered x to every GPU (un-normalised: the
    // consumer divides the remote entries by beta_{k-2}, as k_scale_norm did with the local block)
    op->impl->use_pushed(kry->pushed, kry->d_beta + (k - 2));
  } else if (kry->halo_pushed && kry->pushed_op == op->impl && kry->pushed_col == k - 1) {
    op->impl->use_pushed_halo(kry->halo_plan);  // the fused kernel of the previous iteration pushed the halo of x
  } else {
    LLZ_TRY(op->impl->prepare(x));
  }
  kry->pushed_valid = false;
  kry->halo_pushed = false;
  Fold fold;
  fold.alpha_msg = comm_next_message(ctx, kChanAlpha);  // delivered by the last CTA of the kernel that computes <x, Ax>
  {
    ProfScope ps(ctx, "spmv", (double)op->impl->bytes + (double)kry->n * (double)dtype_size(kry->dtype) * 2);
    LLZ_TRY(op->impl->apply_fused(x, y, sigma, kry->d_pa, &npa, &fold.alpha_msg));
  }
  if (npa == 0) {
    ProfScope ps(ctx, "dot", (double)kry->n * (double)dtype_size(kry->dtype) * 2);
    LLZ_TRY(launch_redot(ctx, kry->dtype, x, y, kry->n, kry->d_pa, &npa, fold.alpha_msg));
  }
  LLZ_TRY(allreduce_scalar(ctx, kry->d_pa, &npa));
  fold.mode = (k == 1) ? 1 : 2;
  fold.alpha_partials = kry->d_pa;
  fold.n_partials = npa;
  fold.beta_prev = (k >= 2) ? kry->d_beta + (k - 2) : nullptr;
  fold.alpha_out = kry->d_alpha + (k - 1);

  int grid = 0;
  ScalarSink sink;
  sink.beta_msg = comm_next_message(ctx, kChanBeta);  // delivered by the kernel that writes the norm partials
  if (sink.beta_msg.ch.G > 0 && op->impl->plan_push(&fold.push)) {  // ... which also pushes u_k to the peers
    kry->pushed_valid = true;
    kry->pushed_col = k;
    kry->pushed_op = op->impl;
    kry->pushed = fold.push;
  }
  if (orth == LLZ_ORTH_RECURRENCE_LAZY) {
    if (ctx->nranks > 1) return fail(LLZ_ERR_UNSUPPORTED, "krylov_step: LLZ_ORTH_RECURRENCE_LAZY is for a single rank");
    LazyRecurrence lazy;
    lazy.scale1 = (k >= 2) ? kry->d_beta + (k - 2) : nullptr;  // column j >= 1 is stored as beta_{j-1} u_j, column 0 is normalised
    lazy.scale2 = (k >= 3) ? kry->d_beta + (k - 3) : nullptr;
    lazy.ticket = kry->d_ticket;
    lazy.sink.beta_out = kry->d_beta + (k - 1);
    lazy.sink.h_alpha = kry->h_alpha + (k - 1);
    lazy.sink.h_beta = kry->h_beta + (k - 1);
    lazy.sink.h_wnorm = kry->h_wnorm + (k - 1);
    lazy.sink.h_flag = kry->h_flag;
    lazy.sink.flag_value = k;
    ProfScope ps(ctx, "recurrence", (double)kry->n * (double)dtype_size(kry->dtype) * (2 + fold.mode));
    LLZ_TRY(launch_recurrence(ctx, kry->dtype, y, kry->col(k - 1), k >= 2 ? kry->col(k - 2) : nullptr, y, kry->n, fold, kry->d_pb, &grid,
                              &lazy));
    kry->k = k;
    return LLZ_OK;
  }
  if (orth == LLZ_ORTH_RECURRENCE) {
    fold.norm_msg = sink.beta_msg;
    ProfScope ps(ctx, "recurrence", (double)kry->n * (double)dtype_size(kry->dtype) * (2 + fold.mode));
    LLZ_TRY(launch_recurrence(ctx, kry->dtype, y, kry->col(k - 1), k >= 2 ? kry->col(k - 2) : nullptr, y, kry->n, fold,
                              kry->d_pb, &grid));
  } else {
    ColumnSet cs;
    cs.V = kry->col(0);
    cs.ld = kry->ld;
    cs.nv = (int)k;
    cs.Q = (const void* const*)kry->d_qptrs;
    cs.nq = kry->nq;
    if (orth != LLZ_ORTH_FULL_TWICE) fold.norm_msg = sink.beta_msg;
    // One cooperative launch for project + reduce + update + norm + normalise when the columns fit one projection
    // chunk and the scalars travel through peer memory (or there is one rank); else the separate kernels below.
    const int total = cs.ncols(), nc = dtype_nc(kry->dtype);
    const bool peer = comm_p2p(ctx) && total * nc + 1 <= comm_coef_capacity(ctx);
    if (orth == LLZ_ORTH_FULL && ctx->fuse_orth && (ctx->nranks == 1 || peer) && orth_fusable(ctx, kry->dtype, total, kry->n)) {
      const int max_grid = std::min(kMaxGrid, ctx->num_sms * 2);
      LLZ_TRY(ensure_ph(kry, (size_t)max_grid * ((size_t)total * nc + 1)));
      PeerMsg coef_msg;
      if (peer) {
        coef_msg = comm_next_message(ctx, kChanCoef);
        fold.wnorm_out = kry->d_misc + 1;
        fold.wnorm_index = total * nc;
      }
      sink.beta_out = kry->d_beta + (k - 1);
      sink.alpha_in = kry->d_alpha + (k - 1);
      sink.h_alpha = kry->h_alpha + (k - 1);
      sink.h_beta = kry->h_beta + (k - 1);
      sink.h_wnorm = kry->h_wnorm + (k - 1);
      sink.wnorm2_in = kry->d_misc + 1;
      sink.h_flag = kry->h_flag;
      sink.flag_value = k;
      const double es = (double)dtype_size(kry->dtype);
      ProfScope ps(ctx, "orth", (double)kry->n * es * ((total + 1 + fold.mode) + (total + 2) + 2));
      HaloPushPlan halo;
      if (peer && !kry->pushed_valid && op->impl->plan_halo_push(&halo)) {  // ... and the halo of u_k for the next apply
        kry->halo_pushed = true;
        kry->pushed_col = k;
        kry->pushed_op = op->impl;
        kry->halo_plan = halo;
      }
      int fused = 0;
      LLZ_TRY(launch_orth(ctx, kry->dtype, cs, y, kry->n, fold, kry->d_ph, kry->d_coef, kry->d_misc + 1, coef_msg, peer ? total * nc : -1,
                          kry->d_pb, sink, halo, &fused, &grid));
      if (!fused) return fail(LLZ_ERR_CUDA, "krylov_step: the fused orthogonalisation kernel declined a shape it had accepted");
      kry->k = k;
      return LLZ_OK;
    }
    LLZ_TRY(cgs_pass(kry, cs, y, fold, orth != LLZ_ORTH_FULL_TWICE, &grid));
    if (orth == LLZ_ORTH_FULL_TWICE) {
      Fold nofold;
      nofold.norm_msg = sink.beta_msg;
      nofold.push = fold.push;
      LLZ_TRY(cgs_pass(kry, cs, y, nofold, true, &grid));
    }
  }
  LLZ_TRY(allreduce_scalar(ctx, kry->d_pb, &grid));
  sink.beta_out = kry->d_beta + (k - 1);
  sink.alpha_in = kry->d_alpha + (k - 1);
  sink.h_alpha = kry->h_alpha + (k - 1);
  sink.h_beta = kry->h_beta + (k - 1);
  sink.h_wnorm = kry->h_wnorm + (k - 1);
  sink.wnorm2_in = (orth == LLZ_ORTH_RECURRENCE) ? nullptr : kry->d_misc + 1;
  sink.h_flag = kry->h_flag;
  sink.flag_value = k;
  {
    ProfScope ps(ctx, "scale", (double)kry->n * (double)dtype_size(kry->dtype) * 2);
    LLZ_TRY(launch_scale_by_norm(ctx, kry->dtype, y, kry->n, kry->d_pb, grid, sink));
  }
  kry->k = k;
  return LLZ_OK;
}

int llz_krylov_fetch(llz_krylov_t kry, int64_t k, double* alpha, double* beta, double* wnorm) {
  if (!kry || k < 1 || k > kry->k) return fail(LLZ_ERR_INVALID, "krylov_fetch: iteration %lld not enqueued", (long long)k);
  volatile long long* flag = kry->h_flag;
  uint64_t spins = 0;
  while (*flag < k) {
    if ((++spins & 0x3ff) == 0) {
      cudaError_t e = cudaStreamQuery(kry->ctx->stream);
      if (e == cudaSuccess) {
        if (*flag >= k) break;
        return fail(LLZ_ERR_CUDA, "krylov_fetch: stream idle but iteration %lld never published its scalars", (long long)k);
      }
      if (e != cudaErrorNotReady) return fail(LLZ_ERR_CUDA, "krylov_fetch: %s", cudaGetErrorString(e));
    }
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  }
  __sync_synchronize();
  if (alpha) *alpha = ((volatile double*)kry->h_alpha)[k - 1];
  if (beta) *beta = ((volatile double*)kry->h_beta)[k - 1];
  if (wnorm) *wnorm = ((volatile double*)kry->h_wnorm)[k - 1];
  return LLZ_OK;
}

int llz_krylov_refine(llz_krylov_t kry, int64_t k, double* shrink) {
  if (!kry || k < 1 || k > kry->k) return fail(LLZ_ERR_INVALID, "krylov_refine: iteration %lld not enqueued", (long long)k);
  llz_ctx_t ctx = kry->ctx;
  void* y = kry->col(k);
  ColumnSet cs;
  cs.V = kry->col(0);
  cs.ld = kry->ld;
  cs.nv = (int)k;
  cs.Q = (const void* const*)kry->d_qptrs;
  cs.nq = kry->nq;
  Fold nofold;
  int grid = 0;
  ScalarSink sink;
  nofold.norm_msg = sink.beta_msg = comm_next_message(ctx, kChanBeta);
  LLZ_TRY(cgs_pass(kry, cs, y, nofold, true, &grid));
  LLZ_TRY(allreduce_scalar(ctx, kry->d_pb, &grid));
  sink.beta_out = kry->d_misc;
  sink.h_beta = kry->h_misc;
  {
    ProfScope ps(ctx, "scale", (double)kry->n * (double)dtype_size(kry->dtype) * 2);
    LLZ_TRY(launch_scale_by_norm(ctx, kry->dtype, y, kry->n, kry->d_pb, grid, sink));
  }
  LLZ_CUDA(cudaStreamSynchronize(ctx->stream));  // also drains the speculative iterations queued after k
  const double nu = kry->h_misc[0];              // column k had unit norm: nu is the shrink factor of this pass
  const double beta_new = kry->h_beta[k - 1] * nu;
  kry->h_beta[k - 1] = beta_new;
  LLZ_CUDA(cudaMemcpyAsync(kry->d_beta + (k - 1), kry->h_beta + (k - 1), sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  LLZ_CUDA(cudaStreamSynchronize(ctx->stream));
  kry->k = k;  // iterations enqueued beyond k used the un-refined vector: drop them
  kry->pushed_valid = false;  // ... and so did the copy of it the peers were sent
  kry->halo_pushed = false;
  kry->h_flag[0] = k;
  if (shrink) *shrink = nu;
  return LLZ_OK;
}

int llz_krylov_steps(llz_krylov_t kry, int64_t* k) {
  if (!kry || !k) return fail(LLZ_ERR_INVALID, "null");
  *k = kry->k;
  return LLZ_OK;
}

int llz_krylov_combine(llz_krylov_t kry, int64_t m, int64_t nvec, const void* coeff, int normalize,
                       const llz_vec_t* out) {
  if (!kry || !coeff || !out || m < 1 || nvec < 1 || m > kry->k + 1)
    return fail(LLZ_ERR_INVALID, "krylov_combine: bad argument (m=%lld, nvec=%lld, stored=%lld)", (long long)m,
                (long long)nvec, kry ? (long long)kry->k + 1 : 0LL);
  llz_ctx_t ctx = kry->ctx;
  const size_t es = dtype_size(kry->dtype);
  const size_t need = (size_t)m * (size_t)nvec * es;
  if (need > kry->ycoef_cap) {
    LLZ_CUDA(cudaStreamSynchronize(ctx->stream));
    if (kry->d_ycoef) cudaFree(kry->d_ycoef);
    kry->d_ycoef = nullptr;
    LLZ_CUDA(cudaMalloc(&kry->d_ycoef, need * 2));
    kry->ycoef_cap = need * 2;
  }
  LLZ_CUDA(cudaMemcpyAsync(kry->d_ycoef, coeff, need, cudaMemcpyHostToDevice, ctx->stream));
  LLZ_CUDA(cudaStreamSynchronize(ctx->stream));  // `coeff` is caller-owned pageable memory
  for (int64_t r0 = 0; r0 < nvec; r0 += 5) {
    const int nv = (int)std::min<int64_t>(5, nvec - r0);
    void* outs[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    for (int r = 0; r < nv; ++r) {
      llz_vec_t v = out[r0 + r];
      if (!v || v->n != kry->n || v->dtype != kry->dtype) return fail(LLZ_ERR_INVALID, "krylov_combine: output %lld mismatched", (long long)(r0 + r));
      outs[r] = v->d;
    }
    const int chunk = max_combine_cols(kry->dtype, nv);
    int grid = 0;
    for (int64_t c0 = 0; c0 < m; c0 += chunk) {
      const int cols = (int)std::min<int64_t>(chunk, m - c0);
      const bool last = c0 + cols >= m;
      ProfScope ps(ctx, "combine", (double)kry->n * (double)dtype_size(kry->dtype) * (cols + nv * (c0 > 0 ? 2 : 1)));
      LLZ_TRY(launch_combine(ctx, kry->dtype, kry->col(0), kry->ld, (int)c0, cols, (const char*)kry->d_ycoef + (size_t)r0 * m * es,
                             m, nv, outs, kry->n, c0 > 0, (last && normalize) ? kry->d_pb : nullptr, &grid));
    }
    if (normalize) {
      for (int r = 0; r < nv; ++r) {
        int g = grid;
        LLZ_TRY(comm_allreduce_partials(ctx, kry->d_pb + (size_t)r * kMaxGrid, &g));
        ScalarSink none;
        ProfScope ps(ctx, "scale", (double)kry->n * (double)dtype_size(kry->dtype) * 2);
        LLZ_TRY(launch_scale_by_norm(ctx, kry->dtype, outs[r], kry->n, kry->d_pb + (size_t)r * kMaxGrid, g, none));
      }
    }
  }
  if (comm_p2p(ctx)) {  // end of a run: report a peer that stopped answering instead of returning garbage
    LLZ_CUDA(cudaStreamSynchronize(ctx->stream));
    LLZ_TRY(comm_check_peers(ctx));
  }
  return LLZ_OK;
}

int llz_krylov_column_ptr(llz_krylov_t kry, int64_t j, void** dev) {
  if (!kry || !dev || j < 0 || j > kry->k) return fail(LLZ_ERR_INVALID, "column %lld not stored", (long long)j);
  *dev = kry->col(j);
  return LLZ_OK;
}

int llz_krylov_download_column(llz_krylov_t kry, int64_t j, void* host) {
  if (!kry || !host || j < 0 || j > kry->k) return fail(LLZ_ERR_INVALID, "column %lld not stored", (long long)j);
  LLZ_CUDA(cudaMemcpyAsync(host, kry->col(j), (size_t)kry->n * dtype_size(kry->dtype), cudaMemcpyDeviceToHost,
                           kry->ctx->stream));
  LLZ_CUDA(cudaStreamSynchronize(kry->ctx->stream));
  return LLZ_OK;
}

}  // extern "C"
