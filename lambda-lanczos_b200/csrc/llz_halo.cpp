// llz_halo.cpp — host-side planning for row-sharded operators (no CUDA in this file; tested on CPU).
//
//   llz_partition   contiguous, balanced row blocks: rank r owns [b(r), b(r+1)), b(r) = floor(n*r/G) rounded down to a
//                   multiple of 4 (b(0) = 0, b(G) = n)
//   llz_halo_plan   for the local row block of a CSR matrix with GLOBAL column indices: which remote entries of x the
//                   block references (sorted, hence grouped by owner), and the column indices rewritten to the local
//                   "extended" numbering  [0, n_rows) = own block,  n_rows + h = h-th halo entry.
//
// The reference has no counterpart (single address space, lambda_lanczos.hpp:243 calls mv_mul on whole vectors); this
// is the piece SURVEY.md §8e asks for: "halo exchange of only the remote columns actually referenced".
#include <algorithm>
#include <cstring>
#include <vector>

#include "llz_internal.hpp"

using namespace llz;

namespace llz {

// One planning pass shared by the C entry point and the sharded CSR constructor: `remote` receives the sorted unique
// remote columns (grouped by owner because sorted), per_owner their count per rank, colidx_local (may be null) the
// column indices in the local extended numbering.
int halo_plan_vectors(int64_t n_rows, int64_t row0, const int64_t* rowptr, const int32_t* colidx, int nranks,
                      const int64_t* boundaries, int32_t* colidx_local, std::vector<int32_t>& remote, int64_t* per_owner) {
  const int64_t nnz = rowptr[n_rows];
  const int64_t lo = row0, hi = row0 + n_rows;
  const int64_t n_global = boundaries[nranks];
  remote.clear();
  for (int64_t p = 0; p < nnz; ++p) {
    const int64_t c = colidx[p];
    if (c < 0 || c >= n_global) return fail(LLZ_ERR_INVALID, "halo_plan: column %lld outside [0, %lld)", (long long)c, (long long)n_global);
    if (c < lo || c >= hi) remote.push_back((int32_t)c);
  }
  std::sort(remote.begin(), remote.end());
  remote.erase(std::unique(remote.begin(), remote.end()), remote.end());
  if (n_rows + (int64_t)remote.size() >= (int64_t)0x7fffffff)
    return fail(LLZ_ERR_UNSUPPORTED, "halo_plan: local block + halo exceeds 32-bit column indices");
  if (per_owner) {
    for (int r = 0; r < nranks; ++r) {
      auto a = std::lower_bound(remote.begin(), remote.end(), (int32_t)std::min<int64_t>(boundaries[r], 0x7fffffff));
      auto b = std::lower_bound(remote.begin(), remote.end(), (int32_t)std::min<int64_t>(boundaries[r + 1], 0x7fffffff));
      per_owner[r] = (int64_t)(b - a);
    }
  }
  if (colidx_local) {
    for (int64_t p = 0; p < nnz; ++p) {
      const int64_t c = colidx[p];
      if (c >= lo && c < hi) {
        colidx_local[p] = (int32_t)(c - lo);
      } else {
        const auto it = std::lower_bound(remote.begin(), remote.end(), (int32_t)c);
        colidx_local[p] = (int32_t)(n_rows + (it - remote.begin()));
      }
    }
  }
  return LLZ_OK;
}

}  // namespace llz

extern "C" {

int llz_partition(int64_t n_global, int rank, int nranks, int64_t* row0, int64_t* n_local) {
  if (n_global < 0 || nranks < 1 || rank < 0 || rank >= nranks || !row0 || !n_local)
    return fail(LLZ_ERR_INVALID, "partition: bad argument (n=%lld, rank %d of %d)", (long long)n_global, rank, nranks);
  // 128-bit products are not needed: n < 2^40 and nranks <= 2^10 in any realistic run
  // interior boundaries are rounded down to a multiple of 4 elements, so that a row block starts 16-byte aligned inside
  // a gathered vector for every element type (the fused all-gather stores 128-bit packets into the peers' buffers)
  auto bound = [&](int r) -> int64_t {
    if (r <= 0) return 0;
    if (r >= nranks) return n_global;
    return (int64_t)((__int128)n_global * r / nranks) & ~(int64_t)3;
  };
  const int64_t a = bound(rank);
  const int64_t b = bound(rank + 1);
  *row0 = a;
  *n_local = b - a;
  return LLZ_OK;
}

int llz_halo_plan(int64_t n_rows, int64_t row0, const int64_t* rowptr, const int32_t* colidx, int nranks,
                  const int64_t* boundaries, int32_t* colidx_local, int64_t* halo_cols, int64_t halo_capacity,
                  int64_t* n_halo, int64_t* per_owner) {
  if (n_rows < 0 || !rowptr || !colidx || nranks < 1 || !boundaries || !n_halo)
    return fail(LLZ_ERR_INVALID, "halo_plan: bad argument");
  std::vector<int32_t> remote;
  std::vector<int64_t> owners((size_t)nranks, 0);
  LLZ_TRY(llz::halo_plan_vectors(n_rows, row0, rowptr, colidx, nranks, boundaries, colidx_local, remote, owners.data()));
  *n_halo = (int64_t)remote.size();
  if (per_owner)
    for (int r = 0; r < nranks; ++r) per_owner[r] = owners[(size_t)r];
  if (halo_cols) {
    if (halo_capacity < (int64_t)remote.size())
      return fail(LLZ_ERR_INVALID, "halo_plan: halo_cols holds %lld entries, %lld needed", (long long)halo_capacity, (long long)remote.size());
    for (size_t i = 0; i < remote.size(); ++i) halo_cols[i] = remote[i];
  }
  return LLZ_OK;
}

}  // extern "C"
