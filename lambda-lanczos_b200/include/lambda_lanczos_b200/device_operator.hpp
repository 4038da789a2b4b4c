// device_operator.hpp — DeviceOperator<T>: what the reference's
//     std::function<void(const std::vector<T>& in, std::vector<T>& out)> mv_mul
// (lambda_lanczos.hpp:126, exponentiator.hpp:41) becomes when the vectors live in HBM.  Copyable and reassignable
// like the std::function it replaces; the underlying llz_op_t is shared.
#pragma once
#include <algorithm>
#include <functional>
#include <utility>

#include "common.hpp"

namespace lambda_lanczos_b200 {

template <typename T>
class DeviceOperator {
 public:
  DeviceOperator() {}

  // Built-in CSR SpMV (32-bit column indices); host arrays are copied to the device.
  static DeviceOperator csr(const Context& ctx, size_t n, const int64_t* rowptr, const int32_t* colidx, const T* vals) {
    llz_op_t op = nullptr;
    check(llz_op_create_csr(ctx.get(), util::dtype_of<T>::value, (int64_t)n, (int64_t)n, 0, rowptr, colidx, vals, 1, &op),
          "llz_op_create_csr");
    return DeviceOperator(ctx, op, n);
  }
  // SELL-32-sigma: same CSR input, re-stored on the device in sliced-ELL form (coalesced, barrier-free SpMV for
  // stencil / lattice matrices); sigma = 0 picks whether to sort rows by length, 1 keeps the row order.
  static DeviceOperator sell(const Context& ctx, size_t n, const int64_t* rowptr, const int32_t* colidx, const T* vals, int sigma = 0) {
    llz_op_t op = nullptr;
    check(llz_op_create_sell(ctx.get(), util::dtype_of<T>::value, (int64_t)n, (int64_t)n, 0, rowptr, colidx, vals, 1, sigma, &op),
          "llz_op_create_sell");
    return DeviceOperator(ctx, op, n);
  }
  // COO triplets as in the reference's sparse sample (src/samples/sample2_sparse.cpp:14-47): duplicates are summed.
  static DeviceOperator coo(const Context& ctx, size_t n, const std::vector<size_t>& rows, const std::vector<size_t>& cols,
                            const std::vector<T>& vals) {
    std::vector<int64_t> rowptr(n + 1, 0);
    for (size_t r : rows) rowptr[r + 1]++;
    for (size_t i = 0; i < n; ++i) rowptr[i + 1] += rowptr[i];
    std::vector<int32_t> ci(vals.size());
    std::vector<T> v(vals.size());
    std::vector<int64_t> fill(rowptr.begin(), rowptr.end() - 1);
    for (size_t k = 0; k < vals.size(); ++k) {
      const int64_t p = fill[rows[k]]++;
      ci[(size_t)p] = (int32_t)cols[k];
      v[(size_t)p] = vals[k];
    }
    return csr(ctx, n, rowptr.data(), ci.data(), v.data());
  }
  // Matrix-free spin-1/2 XXZ chain in a fixed-Sz sector (BASELINE.json configs 4/5).
  static DeviceOperator xxz(const Context& ctx, int L, int n_up, double jz, double jxy, bool periodic) {
    llz_op_t op = nullptr;
    check(llz_op_create_xxz(ctx.get(), util::dtype_of<T>::value, L, n_up, jz, jxy, periodic ? 1 : 0, &op), "llz_op_create_xxz");
    int64_t n = 0;
    check(llz_op_rows(op, &n), "llz_op_rows");
    return DeviceOperator(ctx, op, (size_t)n);
  }
  // User device code: `fn(x_dev, y_dev, n, stream)` must enqueue y = A x (overwrites = true) or y += A x
  // (overwrites = false: y arrives zero-filled, the reference's contract, lambda_lanczos.hpp:124) on `stream`.
  // This is also how a user __device__ functor is plugged in: wrap its kernel launch in the callable.
  using DeviceFn = std::function<void(const T* x_dev, T* y_dev, size_t n, void* cuda_stream)>;
  static DeviceOperator callback(const Context& ctx, size_t n, DeviceFn fn, bool overwrites) {
    auto holder = std::make_shared<DeviceFn>(std::move(fn));
    llz_op_t op = nullptr;
    check(llz_op_create_callback(ctx.get(), util::dtype_of<T>::value, (int64_t)n, &DeviceOperator::trampoline, holder.get(),
                                 overwrites ? 1 : 0, &op),
          "llz_op_create_callback");
    DeviceOperator d(ctx, op, n);
    d.fn_ = holder;
    return d;
  }

  // The reference's own operator type (lambda_lanczos.hpp:120-126, exponentiator.hpp:35-41): a HOST callable that
  // ADDS A*in to a zero-filled `out`.  The Krylov loop still runs on the GPU; every application copies the input vector
  // to the host, calls `fn`, and copies the result back — the price of an operator that only exists as CPU code.  This is
  // what lets source written against the reference (its samples, its tests) compile unchanged.
  using HostFn = std::function<void(const std::vector<T>& in, std::vector<T>& out)>;
  static DeviceOperator host_function(const Context& ctx, size_t n, HostFn fn) {
    auto in = std::make_shared<std::vector<T>>(n);
    auto out = std::make_shared<std::vector<T>>(n);
    llz_ctx_t c = ctx.get();
    return callback(
        ctx, n,
        [c, in, out, fn](const T* x_dev, T* y_dev, size_t len, void*) {
          check(llz_ctx_memcpy(c, in->data(), x_dev, len * sizeof(T), 0), "llz_ctx_memcpy");
          std::fill(out->begin(), out->end(), T());  // the reference hands mv_mul a zero-filled output (:242-243)
          fn(*in, *out);
          check(llz_ctx_memcpy(c, y_dev, out->data(), len * sizeof(T), 1), "llz_ctx_memcpy");
        },
        true);
  }

  // Non-owning view of an operator created through the C ABI.
  static DeviceOperator borrow(const Context& ctx, llz_op_t op) {
    DeviceOperator d;
    d.ctx_ = ctx;
    d.h_.reset(op, [](llz_op_t) {});
    d.read_shape();
    return d;
  }
  // Row block [row0, row0 + n_rows) of an n x n CSR matrix with GLOBAL column indices, for a context that joined a
  // row-sharded group (Context::join); rowptr is local (starts at 0).
  static DeviceOperator csr_rows(const Context& ctx, size_t n, size_t row0, size_t n_rows, const int64_t* rowptr,
                                 const int32_t* colidx, const T* vals) {
    llz_op_t op = nullptr;
    check(llz_op_create_csr(ctx.get(), util::dtype_of<T>::value, (int64_t)n_rows, (int64_t)n, (int64_t)row0, rowptr, colidx, vals, 1, &op),
          "llz_op_create_csr");
    return DeviceOperator(ctx, op, n_rows);
  }

  bool valid() const { return (bool)h_; }
  llz_op_t get() const { return h_.get(); }
  size_t rows() const { return n_; }                // rows of the local block (= vector length on this GPU)
  size_t global_rows() const { return n_global_; }  // the reference's matrix_size
  size_t row_offset() const { return row0_; }       // first global row of the local block
  const Context& context() const { return ctx_; }
  int64_t bytes() const {
    int64_t b = 0;
    check(llz_op_bytes(h_.get(), &b), "llz_op_bytes");
    return b;
  }
  // max_i sum_j |a_ij|: all eigenvalues lie in [-r, r] — what to choose eigenvalue_offset from (the reference ships a
  // stand-alone tool for this, src/determine_eigenvalue_offset/determine_eigenvalue_offset.cpp)
  double gerschgorin_radius() const {
    double r = 0.0;
    check(llz_op_gerschgorin_radius(h_.get(), &r), "llz_op_gerschgorin_radius");
    return r;
  }
  // y = A x on device vectors
  void operator()(const DeviceVector<T>& x, DeviceVector<T>& y) const { check(llz_op_apply(h_.get(), x.get(), y.get()), "llz_op_apply"); }

 private:
  DeviceOperator(const Context& ctx, llz_op_t op, size_t n) : ctx_(ctx), n_(n) {
    h_.reset(op, [](llz_op_t p) { llz_op_destroy(p); });
    read_shape();
  }
  void read_shape() {
    int64_t nl = 0, ng = 0, r0 = 0;
    check(llz_op_shape(h_.get(), &nl, &ng, &r0), "llz_op_shape");
    n_ = (size_t)nl;
    n_global_ = (size_t)ng;
    row0_ = (size_t)r0;
  }
  static int trampoline(void* user, const void* x, void* y, int64_t n, void* stream) {
    try {
      (*static_cast<DeviceFn*>(user))(static_cast<const T*>(x), static_cast<T*>(y), (size_t)n, stream);
      return 0;
    } catch (...) {
      return 1;
    }
  }
  Context ctx_ = Context::none();
  std::shared_ptr<llz_op_s> h_;
  std::shared_ptr<DeviceFn> fn_;
  size_t n_ = 0, n_global_ = 0, row0_ = 0;
};

}  // namespace lambda_lanczos_b200
