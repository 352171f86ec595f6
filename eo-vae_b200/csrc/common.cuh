// Shared device/host helpers for the eo-vae sm_100a kernels.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

// ---------------------------------------------------------------- error reporting (C-ABI: eovae_last_error)
namespace eovae {
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
extern unsigned long long g_launches;
// C[batch][i][j] (+)= sum_k A[batch*a_bs + i*a_rs + k*a_cs] * B[batch*b_bs + k*b_rs + j*b_cs]   (fp32 SIMT, hypernet.cu)
int sgemm_batched(const float* a, long long a_rs, long long a_cs, long long a_bs, const float* b, long long b_rs, long long b_cs,
                  long long b_bs, float* c, long long ldc, long long c_bs, int batches, int m, int n, int k, int accumulate,
                  cudaStream_t st);
}  // namespace eovae

#define EOVAE_CHECK(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      eovae::set_error(__VA_ARGS__);      \
      return -1;                          \
    }                                     \
  } while (0)

#define EOVAE_CUDA(call)                                  \
  do {                                                    \
    if (eovae::check_cuda((call), #call) != 0) return -2; \
  } while (0)

// every kernel launch goes through this macro: it also feeds eovae_launch_count() (bench.py "gpu_launches")
#define EOVAE_LAUNCH_CHECK()              \
  do {                                    \
    ++eovae::g_launches;                  \
    EOVAE_CUDA(cudaGetLastError());       \
  } while (0)

// dtype codes used across the C-ABI
enum : int { EOVAE_BF16 = 0, EOVAE_F16 = 1, EOVAE_F32 = 2 };  // == EOVAE_DT_* in include/eovae.h

// ---------------------------------------------------------------- 16-bit <-> float helpers
template <typename T>
struct T16;
template <>
struct T16<__nv_bfloat16> {
  using v2 = __nv_bfloat162;
  static __device__ __forceinline__ float2 to_f2(uint32_t u) {
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
  }
  static __device__ __forceinline__ uint32_t from_f2(float a, float b) {
    __nv_bfloat162 r = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&r);
  }
  static __device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};
template <>
struct T16<__half> {
  using v2 = __half2;
  static __device__ __forceinline__ float2 to_f2(uint32_t u) {
    return __half22float2(*reinterpret_cast<const __half2*>(&u));
  }
  static __device__ __forceinline__ uint32_t from_f2(float a, float b) {
    __half2 r = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&r);
  }
  static __device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
  static __device__ __forceinline__ __half from_f(float v) { return __float2half_rn(v); }
};

template <>
struct T16<float> {  // scalar identity: lets the element-wise templates serve the fp32 validation path
  static __device__ __forceinline__ float to_f(float v) { return v; }
  static __device__ __forceinline__ float from_f(float v) { return v; }
};

__device__ __forceinline__ float2 unpack16(uint32_t u, int dtype) {
  return dtype == EOVAE_BF16 ? T16<__nv_bfloat16>::to_f2(u) : T16<__half>::to_f2(u);
}
__device__ __forceinline__ uint32_t pack16(float a, float b, int dtype) {
  return dtype == EOVAE_BF16 ? T16<__nv_bfloat16>::from_f2(a, b) : T16<__half>::from_f2(a, b);
}

// x * sigmoid(x) with one ex2 and one fast reciprocal (2 ulp)
__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

int eovae_num_sms();
