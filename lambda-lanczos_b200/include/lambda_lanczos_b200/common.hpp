// common.hpp — type utilities and RAII wrappers over the C ABI (include/llz.h) shared by the host-side engine headers.
// Header-only C++14; needs -I<repo>/include for llz.h and linking against libllz.so.
#pragma once
#include <complex>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "llz.h"

namespace lambda_lanczos_b200 {
namespace util {

// real_t<T>: T for real types, the component type for std::complex (reference: util/common.hpp:80-102)
template <typename T> struct realTypeMap { typedef T type; };
template <typename T> struct realTypeMap<std::complex<T>> { typedef T type; };
template <typename T> using real_t = typename realTypeMap<T>::type;

// conjugate that is the identity on real types (reference: util/common.hpp:112-134)
template <typename T> inline T typed_conj(const T& v) { return v; }
template <typename T> inline std::complex<T> typed_conj(const std::complex<T>& v) { return std::conj(v); }

template <typename T> struct dtype_of;
template <> struct dtype_of<float> { static constexpr int value = LLZ_F32; };
template <> struct dtype_of<double> { static constexpr int value = LLZ_F64; };
template <> struct dtype_of<std::complex<float>> { static constexpr int value = LLZ_C64; };
template <> struct dtype_of<std::complex<double>> { static constexpr int value = LLZ_C128; };

template <typename T> inline void to_pair(const T& v, double out[2]) {
  out[0] = static_cast<double>(v);
  out[1] = 0.0;
}
template <typename T> inline void to_pair(const std::complex<T>& v, double out[2]) {
  out[0] = static_cast<double>(v.real());
  out[1] = static_cast<double>(v.imag());
}

}  // namespace util

// The reference reports nothing but asserts; the C ABI returns status codes and this layer turns a
// non-zero status into an exception.
class Error : public std::runtime_error {
 public:
  Error(int status, const std::string& what) : std::runtime_error(what), status_(status) {}
  int status() const { return status_; }

 private:
  int status_;
};

inline void check(int status, const char* where) {
  if (status != LLZ_OK)
    throw Error(status, std::string(where) + ": " + llz_status_string(status) + " — " + llz_last_error());
}

// One GPU + one stream (+ one rank of a row-sharded group).
class Context {
 public:
  explicit Context(int device = 0) {
    llz_ctx_t c = nullptr;
    check(llz_ctx_create(device, &c), "llz_ctx_create");
    h_.reset(c, [](llz_ctx_t p) { llz_ctx_destroy(p); });
  }
  // Non-owning view of a context created through the C ABI.
  static Context borrow(llz_ctx_t c) {
    Context x(Borrow{});
    x.h_.reset(c, [](llz_ctx_t) {});
    return x;
  }
  // An empty handle (no device attached); only useful as a placeholder.
  static Context none() { return Context(Borrow{}); }
  llz_ctx_t get() const { return h_.get(); }
  void synchronize() const { check(llz_ctx_synchronize(h_.get()), "llz_ctx_synchronize"); }
  // Join a row-sharded group: one process per GPU, `id128` = the blob of unique_id() made on rank 0 and handed to
  // every rank by whatever transport the application has.  Afterwards vectors are local row blocks.
  static std::vector<unsigned char> unique_id() {
    std::vector<unsigned char> id(128);
    check(llz_comm_unique_id(id.data()), "llz_comm_unique_id");
    return id;
  }
  void join(int rank, int nranks, const void* id128) const { check(llz_ctx_join(h_.get(), rank, nranks, id128), "llz_ctx_join"); }
  int rank() const {
    int r = 0, n = 1;
    check(llz_ctx_rank(h_.get(), &r, &n), "llz_ctx_rank");
    return r;
  }
  int nranks() const {
    int r = 0, n = 1;
    check(llz_ctx_rank(h_.get(), &r, &n), "llz_ctx_rank");
    return n;
  }
  uint64_t launch_count() const {
    uint64_t c = 0;
    check(llz_ctx_launch_count(h_.get(), &c), "llz_ctx_launch_count");
    return c;
  }
  // Process-wide default context on device 0, created on first use.
  static Context& default_context() {
    static Context ctx(0);
    return ctx;
  }

 private:
  struct Borrow {};
  explicit Context(Borrow) {}
  std::shared_ptr<llz_ctx_s> h_;
};

namespace detail {
inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
  __builtin_ia32_pause();
#elif defined(__aarch64__)
  asm volatile("yield");
#endif
}
}  // namespace detail

// Iterations the GPU runs ahead of the host's convergence test when the engine's `pipeline_depth` is negative (the
// default): a Lanczos iteration at depth k streams (2k + 7) vectors, so with vectors of a few MB it is shorter than the
// host's Ritz solve plus the launch latency and the launches must be queued several iterations ahead to keep the GPU
// busy; with vectors of hundreds of MB one iteration ahead hides the host entirely, and every speculative iteration
// past convergence is milliseconds thrown away — for the Exponentiator, which stops after ~10 iterations, one in ten,
// so it runs in lock-step there.  `vector_bytes` must be the same on every rank of a row-sharded group (the ranks
// replicate the control flow): pass global_rows * sizeof(T) / nranks.
inline int auto_pipeline_depth(size_t vector_bytes, bool lanczos) {
  if (vector_bytes < ((size_t)4 << 20)) return 4;
  if (vector_bytes < ((size_t)32 << 20)) return 2;
  return lanczos ? 1 : 0;
}

// Device vector of n_local elements of T.  Holds a reference to its context: the vector's memory goes back to the
// context's pool when it dies, so the context must outlive it (e.g. eigenvectors returned by run_device()).
template <typename T>
class DeviceVector {
 public:
  DeviceVector() {}
  DeviceVector(const Context& ctx, size_t n) : n_(n), ctx_(ctx) {
    llz_vec_t v = nullptr;
    check(llz_vec_create(ctx.get(), util::dtype_of<T>::value, (int64_t)n, &v), "llz_vec_create");
    h_.reset(v, [](llz_vec_t p) { llz_vec_destroy(p); });
  }
  bool valid() const { return (bool)h_; }
  size_t size() const { return n_; }
  llz_vec_t get() const { return h_.get(); }
  T* device_ptr() const {
    void* p = nullptr;
    check(llz_vec_device_ptr(h_.get(), &p), "llz_vec_device_ptr");
    return static_cast<T*>(p);
  }
  void upload(const T* host) { check(llz_vec_upload(h_.get(), host), "llz_vec_upload"); }
  void upload(const std::vector<T>& host) { upload(host.data()); }
  void download(T* host) const { check(llz_vec_download(h_.get(), host), "llz_vec_download"); }
  std::vector<T> to_host() const {
    std::vector<T> out(n_);
    download(out.data());
    return out;
  }

 private:
  size_t n_ = 0;
  Context ctx_ = Context::none();  // declared before h_: destroyed after it
  std::shared_ptr<llz_vec_s> h_;
};

// The reference's vector helpers (util/linear_algebra.hpp:30-144) on device vectors, same names and semantics.
namespace util {
template <typename T> inline T inner_prod(const DeviceVector<T>& v1, const DeviceVector<T>& v2) {  // conjugates v1 (:30-51)
  double out[2] = {0, 0};
  check(llz_vec_dot(v1.get(), v2.get(), out), "llz_vec_dot");
  std::complex<double> z(out[0], out[1]);
  return static_cast<T>(*reinterpret_cast<const typename std::conditional<std::is_same<T, real_t<T>>::value, double, std::complex<double>>::type*>(&z));
}
template <typename T> inline real_t<T> norm(const DeviceVector<T>& v) {  // :57-60
  double out = 0;
  check(llz_vec_norm(v.get(), &out), "llz_vec_norm");
  return (real_t<T>)out;
}
template <typename T> inline real_t<T> m_norm(const DeviceVector<T>& v) {  // :83-125
  double out = 0;
  check(llz_vec_m_norm(v.get(), &out), "llz_vec_m_norm");
  return (real_t<T>)out;
}
template <typename T> inline void scalar_mul(T a, DeviceVector<T>& v) {  // :66-72
  double z[2];
  to_pair(a, z);
  check(llz_vec_scale(v.get(), z), "llz_vec_scale");
}
template <typename T> inline void normalize(DeviceVector<T>& v) { check(llz_vec_normalize(v.get(), nullptr), "llz_vec_normalize"); }  // :78-80
template <typename T> inline void schmidt_orth(DeviceVector<T>& uorth, const std::vector<DeviceVector<T>>& us) {  // :133-144
  std::vector<llz_vec_t> hs;
  for (const auto& u : us) hs.push_back(u.get());
  check(llz_vec_schmidt_orth(uorth.get(), hs.data(), (int64_t)hs.size(), 1), "llz_vec_schmidt_orth");
}
}  // namespace util

}  // namespace lambda_lanczos_b200
