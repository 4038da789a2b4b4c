// llz_launch.hpp — typed-erased launchers of the streaming kernels (llz_kernels.cu).  All launch on ctx->stream.
#pragma once
#include <cstring>

#include "llz_internal.hpp"
#include "llz_peer.cuh"

namespace llz {

#ifdef __CUDACC__
// Launch a kernel of the per-iteration chain on the context's stream.
template <class... KArgs, class... Args>
inline cudaError_t launch_chain(llz_ctx_t ctx, void (*kernel)(KArgs...), int grid, int block, size_t smem, Args&&... args) {
  kernel<<<grid, block, smem, ctx->stream>>>(KArgs(args)...);
  return cudaGetLastError();
}
#endif

// The NEXT message of a peer channel (llz_comm.cu; every rank calls this in the same order); unused (ch.G == 0) when
// the context has no peer channels.
PeerMsg comm_next_message(llz_ctx_t ctx, int which);
enum { kChanAlpha = 0, kChanBeta = 1, kChanCoef = 2, kChanHalo = 3, kChanGather = 4 };


// The set of orthonormal columns a vector is projected on: `nq` separately allocated vectors (device pointer table)
// followed by `nv` contiguous columns of the Krylov basis.  Column index space: [0,nq) = Q, [nq,nq+nv) = V.
struct ColumnSet {
  const void* V = nullptr;
  int64_t ld = 0;  // elements between consecutive basis columns
  int nv = 0;
  const void* const* Q = nullptr;  // device array of nq device pointers
  int nq = 0;
  int ncols() const { return nq + nv; }
};

// Folding of the three-term recurrence into the projection / update (lambda_lanczos.hpp:251-257):
//   w' = w - alpha u_{k-1} - beta_{k-2} u_{k-2},  alpha = sum(alpha_partials)
struct Fold {
  int mode = 0;                         // 0: none, 1: alpha only (first iteration), 2: alpha and beta
  const double* alpha_partials = nullptr;
  int n_partials = 0;
  const double* beta_prev = nullptr;    // device address of beta_{k-2}
  double* alpha_out = nullptr;          // device address receiving alpha_{k-1}
  PeerMsg alpha_msg;                    // row-sharded: alpha = sum over ranks of this message instead of the partials
  PeerMsg norm_msg;                     // row-sharded: the kernel that produces the norm partials (update / recurrence)
                                        // also delivers their sum as this message (consumed by scale_by_norm)
  GatherPush push;                      // row-sharded: ... and stores the vector it writes into the peers' exchange
                                        // buffers (fused all-gather for operators that read the whole input vector)
  double* wnorm_out = nullptr;          // row-sharded with peer channels: the update kernel — which waits for the
  int wnorm_index = -1;                 // coefficient message — copies element wnorm_index of it (||w'||^2) here, so
                                        // that no later kernel has to read a message slot a fast peer may already reuse
};

// Where scale_by_norm publishes the iteration's scalars.
struct ScalarSink {
  double* beta_out = nullptr;     // device bank slot for beta_{k-1} (may be null)
  const double* alpha_in = nullptr;  // device bank slot of alpha_{k-1} (copied to the host mirror)
  double* h_alpha = nullptr;      // mapped pinned host slots (may be null)
  double* h_beta = nullptr;
  const double* wnorm2_in = nullptr;  // device: ||w'||^2 before the Gram-Schmidt pass (null => report beta itself)
  double* h_wnorm = nullptr;      // mapped pinned slot receiving ||w'||
  long long* h_flag = nullptr;    // mapped pinned: set to `flag_value` after the scalars are visible
  long long flag_value = 0;
  PeerMsg beta_msg;               // row-sharded: ||u||^2 = sum over ranks of this message instead of the partials
};

int max_project_cols(int dtype);  // columns one projection launch can accumulate in shared memory
int max_update_cols(int dtype);
int max_combine_cols(int dtype, int nvec);

// h partials: ph[cta][ncols*NC + 1] for the column chunk [col0, col0+ncols); returns the grid used.
int launch_project(llz_ctx_t ctx, int dtype, const ColumnSet& cs, int col0, int ncols, const void* w, int64_t n,
                   const Fold& fold, double* ph, int* grid_out);
// coef[(col0+j)*NC+c] = sum_cta ph[cta][j*NC+c]
// ph rows hold ncols*NC + 1 doubles: the last one is the CTA's partial of ||w'||^2, summed into *wnorm2 if non-null.
// Row-sharded with peer channels: `msg` (ch.G > 0) receives this rank's sums instead of coef/wnorm2 — element
// (col0+j)*NC+c for the coefficients, element wnorm_index for ||w'||^2 (if >= 0) — in EVERY rank's inbox, and the
// message is announced when `publish` is set (last chunk of a pass).
int launch_reduce(llz_ctx_t ctx, int dtype, const double* ph, int grid, int col0, int ncols, double* coef, double* wnorm2,
                  const PeerMsg& msg = PeerMsg(), int wnorm_index = -1, int publish = 0);
// out = w' - sum_j coef_j col_j over the chunk [col0, col0+ncols), where w' = w - alpha u_{k-1} - beta u_{k-2} per `fold`
// (then the chunk must EXCLUDE those fold.mode trailing basis columns: the kernel applies their coefficients itself).
// If norm_partials != null, per-CTA partials of ||out||^2.
int launch_update(llz_ctx_t ctx, int dtype, const ColumnSet& cs, int col0, int ncols, const void* w, void* out,
                  int64_t n, const double* coef, const Fold& fold, double* norm_partials, int* grid_out,
                  const PeerMsg& coef_msg = PeerMsg(), const PeerMsg& norm_msg = PeerMsg());
// One cooperative launch for the whole orthogonalisation step of a full-reorthogonalisation iteration: project on all
// columns of `cs` (recurrence folded per `fold`), reduce, update `w` in place, norm, normalise, publish per `sink`.
// *fused = 0 and nothing is launched when the shape does not allow it; the caller then issues the separate kernels.
bool orth_fusable(llz_ctx_t ctx, int dtype, int total_cols, int64_t n);  // ask BEFORE drawing peer messages for the step
int launch_orth(llz_ctx_t ctx, int dtype, const ColumnSet& cs, void* w, int64_t n, const Fold& fold, double* ph, double* coef,
                double* wnorm2, const PeerMsg& coef_msg, int wnorm_index, double* norm_partials, const ScalarSink& sink,
                const HaloPushPlan& halo, int* fused, int* grid_out);
// x *= 1/sqrt(sum partials); publishes beta (and alpha) per `sink`.  Leaves x untouched when the norm is not > 0.
int launch_scale_by_norm(llz_ctx_t ctx, int dtype, void* x, int64_t n, const double* norm_partials, int n_partials,
                         const ScalarSink& sink);
// out = w - alpha u1 - beta u2 (no reorthogonalisation; exponentiator.hpp:112-118) + norm partials
// `lazy` (LLZ_ORTH_RECURRENCE_LAZY, single rank): u1 / u2 / w are stored scaled by *scale1 / *scale2 / *scale1 (null: 1),
// `out` stays un-normalised, and the kernel's last CTA finishes ||out|| and publishes the iteration's scalars per `sink`
// — no normalisation pass follows.
struct LazyRecurrence {
  const double* scale1 = nullptr;
  const double* scale2 = nullptr;
  ScalarSink sink;
  unsigned int* ticket = nullptr;
};
int launch_recurrence(llz_ctx_t ctx, int dtype, const void* w, const void* u1, const void* u2, void* out, int64_t n,
                      const Fold& fold, double* norm_partials, int* grid_out, const LazyRecurrence* lazy = nullptr);
// out_r (+)= sum_j coef[r*ldc + j] col_j for r < nvec (<= 5); coef is a DEVICE array of T; norm partials per vector
// at norm_partials[r*kMaxGrid + cta] when non-null.
int launch_combine(llz_ctx_t ctx, int dtype, const void* V, int64_t ld, int col0, int ncols, const void* coef,
                   int64_t ldc, int nvec, void* const* out /* host array of device pointers */, int64_t n,
                   int accumulate, double* norm_partials, int* grid_out);
// partials of <a,b> (NC doubles per CTA, interleaved) and of Re<a,b> only
int launch_dot(llz_ctx_t ctx, int dtype, const void* a, const void* b, int64_t n, double* partials, int* grid_out);
// partials of sum |Re a_i| + |Im a_i| (util::m_norm), one double per CTA
int launch_asum(llz_ctx_t ctx, int dtype, const void* a, int64_t n, double* partials, int* grid_out);
// partials of Re<a,b> only, one double per CTA (alpha when the operator cannot fuse the dot)
int launch_redot(llz_ctx_t ctx, int dtype, const void* a, const void* b, int64_t n, double* partials, int* grid_out,
                 const PeerMsg& msg = PeerMsg());
// result[0..NC) = sum of partials (single CTA)
int launch_sum_partials(llz_ctx_t ctx, const double* partials, int count, int nc, double* result, double* h_result);
int launch_scale(llz_ctx_t ctx, int dtype, void* x, int64_t n, const double a[2]);
int launch_axpy(llz_ctx_t ctx, int dtype, void* y, const double a[2], const void* x, int64_t n);

}  // namespace llz
