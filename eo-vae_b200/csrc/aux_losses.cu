// Optional EOConsistencyLoss branches: spectral-angle loss (SAMLoss, consistency_loss.py:186-210) and
// gradient-difference loss (GradientDifferenceLoss, alpha = 1, consistency_loss.py:241-269) on NCHW fp32 tensors.
// HBM bound: forward reads pred + target once (8 bytes / element), backward reads both once and writes the gradient
// (12 bytes / element; the neighbour taps of the gradient-difference adjoint hit L1/L2).  Sums are fp64 atomics of
// per-block fp32 partials, as in the pixel losses.
#include "../../include/eovae.h"
#include "common.cuh"

namespace {

__device__ __forceinline__ void block_add2(float s1, float s2, double* ws) {
  __shared__ float r1[8], r2[8];
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if ((threadIdx.x & 31) == 0) {
    r1[threadIdx.x >> 5] = s1;
    r2[threadIdx.x >> 5] = s2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t1 = 0, t2 = 0;
    for (int i = 0; i < (blockDim.x >> 5); ++i) {
      t1 += r1[i];
      t2 += r2[i];
    }
    atomicAdd(&ws[0], t1);
    atomicAdd(&ws[1], t2);
  }
}

// one thread per pixel (b, p): channel loop with stride hw (coalesced across the warp)
__global__ void __launch_bounds__(256) sam_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, int c,
                                                      long long hw, long long pixels, float eps, double* __restrict__ ws) {
  float s = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < pixels;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long base = (i / hw) * c * hw + (i % hw);
    float dot = 0.f, na = 0.f, nb = 0.f;
    for (int ch = 0; ch < c; ++ch) {
      const float x = __ldg(&a[base + ch * hw]), y = __ldg(&b[base + ch * hw]);
      dot = fmaf(x, y, dot);
      na = fmaf(x, x, na);
      nb = fmaf(y, y, nb);
    }
    s += 1.f - dot / (sqrtf(na) * sqrtf(nb) + eps);
  }
  block_add2(s, 0.f, ws);
}

// d/da of mean(1 - dot / (|a| |b| + eps)):  -(b_c / D - dot |b| a_c / (|a| D^2)) / pixels   (the norm's subgradient at 0 is 0)
__global__ void __launch_bounds__(256) sam_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, int c,
                                                      long long hw, long long pixels, float eps,
                                                      const float* __restrict__ gscale, float* __restrict__ ga) {
  const float k = gscale[0] / static_cast<float>(pixels);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < pixels;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long base = (i / hw) * c * hw + (i % hw);
    float dot = 0.f, na = 0.f, nb = 0.f;
    for (int ch = 0; ch < c; ++ch) {
      const float x = __ldg(&a[base + ch * hw]), y = __ldg(&b[base + ch * hw]);
      dot = fmaf(x, y, dot);
      na = fmaf(x, x, na);
      nb = fmaf(y, y, nb);
    }
    const float ra = sqrtf(na), rb = sqrtf(nb);
    const float d = ra * rb + eps;
    const float c1 = 1.f / d;
    const float c2 = ra > 0.f ? dot * rb / (ra * d * d) : 0.f;
    for (int ch = 0; ch < c; ++ch) {
      const float x = __ldg(&a[base + ch * hw]), y = __ldg(&b[base + ch * hw]);
      ga[base + ch * hw] = -k * (y * c1 - x * c2);
    }
  }
}

__device__ __forceinline__ float sgn(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }
// derivative of | |u| - |v| | with respect to u
__device__ __forceinline__ float gdl_d(float u, float v) { return sgn(fabsf(u) - fabsf(v)) * sgn(u); }

// one thread per element: horizontal pair (x, x+1) and vertical pair (y, y+1)
__global__ void __launch_bounds__(256) gdl_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, int h, int w,
                                                      long long count, double* __restrict__ ws) {
  float sx = 0.f, sy = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < count;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % w), y = static_cast<int>((i / w) % h);
    const float pa = __ldg(&a[i]), pb = __ldg(&b[i]);
    if (x + 1 < w) sx += fabsf(fabsf(__ldg(&a[i + 1]) - pa) - fabsf(__ldg(&b[i + 1]) - pb));
    if (y + 1 < h) sy += fabsf(fabsf(__ldg(&a[i + w]) - pa) - fabsf(__ldg(&b[i + w]) - pb));
  }
  block_add2(sx, sy, ws);
}

__global__ void __launch_bounds__(256) gdl_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, int h, int w,
                                                      long long count, float inv_nx, float inv_ny,
                                                      const float* __restrict__ gscale, float* __restrict__ ga) {
  const float kx = gscale[0] * inv_nx, ky = gscale[0] * inv_ny;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < count;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % w), y = static_cast<int>((i / w) % h);
    const float pa = __ldg(&a[i]), pb = __ldg(&b[i]);
    float g = 0.f;
    if (x > 0) g += kx * gdl_d(pa - __ldg(&a[i - 1]), pb - __ldg(&b[i - 1]));       // right end of the pair (x-1, x)
    if (x + 1 < w) g -= kx * gdl_d(__ldg(&a[i + 1]) - pa, __ldg(&b[i + 1]) - pb);   // left end of the pair (x, x+1)
    if (y > 0) g += ky * gdl_d(pa - __ldg(&a[i - w]), pb - __ldg(&b[i - w]));
    if (y + 1 < h) g -= ky * gdl_d(__ldg(&a[i + w]) - pa, __ldg(&b[i + w]) - pb);
    ga[i] = g;
  }
}

__global__ void sam_finalize_kernel(const double* ws, double pixels, float* out) { out[0] = static_cast<float>(ws[0] / pixels); }
__global__ void gdl_finalize_kernel(const double* ws, double nx, double ny, float* out) {
  out[0] = static_cast<float>(ws[0] / nx + ws[1] / ny);
}

unsigned grid_for(long long n) {
  long long blocks = (n + 255) / 256;
  const long long cap = 8LL * eovae_num_sms();
  if (blocks > cap) blocks = cap;
  return static_cast<unsigned>(blocks < 1 ? 1 : blocks);
}

}  // namespace

extern "C" {

int eovae_sam_loss(const float* pred, const float* target, int b, int c, long long hw, float eps, float* out, void* workspace,
                   size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  EOVAE_CHECK(b >= 1 && c >= 1 && hw >= 1, "sam_loss: bad shape");
  EOVAE_CHECK(workspace_bytes >= 2 * sizeof(double), "sam_loss: workspace too small");
  double* ws = static_cast<double*>(workspace);
  EOVAE_CUDA(cudaMemsetAsync(ws, 0, 2 * sizeof(double), stream));
  const long long pixels = static_cast<long long>(b) * hw;
  sam_fwd_kernel<<<grid_for(pixels), 256, 0, stream>>>(pred, target, c, hw, pixels, eps, ws);
  EOVAE_LAUNCH_CHECK();
  sam_finalize_kernel<<<1, 1, 0, stream>>>(ws, static_cast<double>(pixels), out);
  EOVAE_LAUNCH_CHECK();
  return 0;
}

int eovae_sam_loss_backward(const float* pred, const float* target, int b, int c, long long hw, float eps,
                            const float* grad_scale, float* grad_pred, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  EOVAE_CHECK(b >= 1 && c >= 1 && hw >= 1, "sam_loss_backward: bad shape");
  const long long pixels = static_cast<long long>(b) * hw;
  sam_bwd_kernel<<<grid_for(pixels), 256, 0, stream>>>(pred, target, c, hw, pixels, eps, grad_scale, grad_pred);
  EOVAE_LAUNCH_CHECK();
  return 0;
}

int eovae_grad_diff_loss(const float* pred, const float* target, long long planes, int h, int w, float* out, void* workspace,
                         size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  EOVAE_CHECK(planes >= 1 && h >= 2 && w >= 2, "grad_diff_loss: needs H, W >= 2");
  EOVAE_CHECK(workspace_bytes >= 2 * sizeof(double), "grad_diff_loss: workspace too small");
  double* ws = static_cast<double*>(workspace);
  EOVAE_CUDA(cudaMemsetAsync(ws, 0, 2 * sizeof(double), stream));
  const long long count = planes * h * w;
  gdl_fwd_kernel<<<grid_for(count), 256, 0, stream>>>(pred, target, h, w, count, ws);
  EOVAE_LAUNCH_CHECK();
  gdl_finalize_kernel<<<1, 1, 0, stream>>>(ws, static_cast<double>(planes) * h * (w - 1), static_cast<double>(planes) * (h - 1) * w, out);
  EOVAE_LAUNCH_CHECK();
  return 0;
}

int eovae_grad_diff_loss_backward(const float* pred, const float* target, long long planes, int h, int w,
                                  const float* grad_scale, float* grad_pred, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  EOVAE_CHECK(planes >= 1 && h >= 2 && w >= 2, "grad_diff_loss_backward: needs H, W >= 2");
  const long long count = planes * h * w;
  gdl_bwd_kernel<<<grid_for(count), 256, 0, stream>>>(pred, target, h, w, count,
                                                     static_cast<float>(1.0 / (static_cast<double>(planes) * h * (w - 1))),
                                                     static_cast<float>(1.0 / (static_cast<double>(planes) * (h - 1) * w)),
                                                     grad_scale, grad_pred);
  EOVAE_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
