// functor_operator.cuh — plug a user `__device__` functor in where the reference takes its mv_mul lambda.
//
// The reference's matrix-free sample (src/samples/sample3_dynamic.cpp:17-22) writes the operator as a host lambda over
// std::vector.  On the GPU the same thing is a functor evaluated per row:
//
//     struct Hop {                                          // y_i = (A x)_i
//       size_t n;
//       __device__ double operator()(size_t i, const double* x) const {
//         return -(i > 0 ? x[i - 1] : 0.0) - (i + 1 < n ? x[i + 1] : 0.0);
//       }
//     };
//     auto mv_mul = lambda_lanczos_b200::make_functor_operator<double>(ctx, n, Hop{n});
//     LambdaLanczos<double> engine(mv_mul, n, false, 1);
//
// Include this header from a .cu translation unit compiled by nvcc for sm_100a (it defines a kernel template); the
// rest of the engine headers are plain C++.  Complex element types are seen by the functor as float2 / double2
// (layout-compatible with std::complex).  The functor OVERWRITES y_i, so no zero-fill pass is spent.
#pragma once
#include <cuda_runtime.h>

#include <complex>

#include "device_operator.hpp"

namespace lambda_lanczos_b200 {

template <typename T> struct device_element { typedef T type; };
template <> struct device_element<std::complex<float>> { typedef float2 type; };
template <> struct device_element<std::complex<double>> { typedef double2 type; };

template <typename D, typename F>
__global__ void __launch_bounds__(256) k_functor_apply(F f, const D* __restrict__ x, D* __restrict__ y, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) y[i] = f(i, x);
}

// `f(i, x)` returns row i of A x; it is copied to the device by value at every launch (keep it small: captured
// pointers must be device pointers).
template <typename T, typename F>
DeviceOperator<T> make_functor_operator(const Context& ctx, size_t n, F f) {
  typedef typename device_element<T>::type D;
  int device = 0, sms = 148;
  cudaGetDevice(&device);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  return DeviceOperator<T>::callback(
      ctx, n,
      [f, sms](const T* x, T* y, size_t len, void* stream) {
        size_t blocks = (len + 255) / 256;
        if (blocks > (size_t)sms * 8) blocks = (size_t)sms * 8;
        if (blocks < 1) blocks = 1;
        k_functor_apply<D, F><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(f, reinterpret_cast<const D*>(x), reinterpret_cast<D*>(y), len);
      },
      /*overwrites=*/true);
}

}  // namespace lambda_lanczos_b200
