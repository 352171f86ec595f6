// EQ-VAE regularisation transforms of the training step (new_autoencoder.py:460-464, 519-531, 611-636): the sampled latent
// is bilinearly rescaled (F.interpolate, align_corners=False) and rotated by k * 90 degrees (torch.rot90, dims=[-1,-2]); the
// reconstruction target is the input area-averaged to the reconstruction size (F.interpolate mode='area') and rotated the
// same way.  Tiny tensors (latent: 2 MB at batch 16) - one gather kernel each, fp32 NCHW.
#include "../../include/eovae.h"
#include "common.cuh"

namespace {

// torch.rot90(x, k, dims=[-1, -2]) on an (h, w) plane: output (a, b) reads x[sy][sx]
//   k = 0: (a, b)          k = 1: (h - 1 - b, a)          k = 2: (h - 1 - a, w - 1 - b)          k = 3: (b, w - 1 - a)
__device__ __forceinline__ void rot_src(int k, int a, int b, int h, int w, int& sy, int& sx) {
  switch (k & 3) {
    case 0: sy = a; sx = b; break;
    case 1: sy = h - 1 - b; sx = a; break;
    case 2: sy = h - 1 - a; sx = w - 1 - b; break;
    default: sy = b; sx = w - 1 - a; break;
  }
}

// PyTorch's area_pixel_compute_source_index (align_corners = False, no user scale factor): clamp below at 0
__device__ __forceinline__ void bilinear_taps(int dst, float scale, int in_size, int& i0, int& i1, float& l1) {
  float src = (static_cast<float>(dst) + 0.5f) * scale - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = static_cast<int>(src);
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = src - static_cast<float>(i0);
}

// out [planes][oh][ow] = rot90_k(resize(in [planes][h][w] -> [nh][nw]))
template <bool BACKWARD>
__global__ void resize_rot_kernel(const float* __restrict__ src, float* __restrict__ dst, long long planes, int h, int w, int nh,
                                  int nw, int k, float sy_scale, float sx_scale) {
  const int oh = (k & 1) ? nw : nh, ow = (k & 1) ? nh : nw;
  const long long total = planes * oh * ow;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i % ow), a = static_cast<int>((i / ow) % oh);
    const long long plane = i / (static_cast<long long>(oh) * ow);
    int ry, rx;
    rot_src(k, a, b, nh, nw, ry, rx);
    int y0, y1, x0, x1;
    float ly, lx;
    bilinear_taps(ry, sy_scale, h, y0, y1, ly);
    bilinear_taps(rx, sx_scale, w, x0, x1, lx);
    const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
    const long long base = plane * h * w;
    if (!BACKWARD) {
      dst[i] = w00 * __ldg(&src[base + y0 * w + x0]) + w01 * __ldg(&src[base + y0 * w + x1]) +
               w10 * __ldg(&src[base + y1 * w + x0]) + w11 * __ldg(&src[base + y1 * w + x1]);
    } else {  // src = gradient of the output, dst = gradient of the input (zero-initialised)
      const float g = src[i];
      atomicAdd(&dst[base + y0 * w + x0], w00 * g);
      atomicAdd(&dst[base + y0 * w + x1], w01 * g);
      atomicAdd(&dst[base + y1 * w + x0], w10 * g);
      atomicAdd(&dst[base + y1 * w + x1], w11 * g);
    }
  }
}

// F.interpolate(mode='area') == adaptive average pooling: window [floor(i h / nh), ceil((i + 1) h / nh))
__global__ void area_rot_kernel(const float* __restrict__ in, float* __restrict__ out, long long planes, int h, int w, int nh,
                                int nw, int k) {
  const int oh = (k & 1) ? nw : nh, ow = (k & 1) ? nh : nw;
  const long long total = planes * oh * ow;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i % ow), a = static_cast<int>((i / ow) % oh);
    const long long plane = i / (static_cast<long long>(oh) * ow);
    int ry, rx;
    rot_src(k, a, b, nh, nw, ry, rx);
    const int ys = static_cast<int>((static_cast<long long>(ry) * h) / nh);
    const int ye = static_cast<int>((static_cast<long long>(ry + 1) * h + nh - 1) / nh);
    const int xs = static_cast<int>((static_cast<long long>(rx) * w) / nw);
    const int xe = static_cast<int>((static_cast<long long>(rx + 1) * w + nw - 1) / nw);
    float s = 0.f;
    for (int y = ys; y < ye; ++y)
      for (int x = xs; x < xe; ++x) s += __ldg(&in[plane * h * w + static_cast<long long>(y) * w + x]);
    out[i] = s / static_cast<float>((ye - ys) * (xe - xs));
  }
}

unsigned grid_for(long long n) {
  long long blocks = (n + 255) / 256;
  const long long cap = 8LL * eovae_num_sms();
  if (blocks > cap) blocks = cap;
  return static_cast<unsigned>(blocks < 1 ? 1 : blocks);
}

}  // namespace

extern "C" {

int eovae_latent_resize_rot(const float* z, long long planes, int h, int w, int nh, int nw, int rot_k, float* out, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  EOVAE_CHECK(planes >= 1 && h >= 1 && w >= 1 && nh >= 1 && nw >= 1 && rot_k >= 0 && rot_k <= 3, "latent_resize_rot: bad arguments");
  resize_rot_kernel<false><<<grid_for(planes * nh * nw), 256, 0, st>>>(z, out, planes, h, w, nh, nw, rot_k,
                                                                      static_cast<float>(h) / nh, static_cast<float>(w) / nw);
  EOVAE_LAUNCH_CHECK();
  return 0;
}

int eovae_latent_resize_rot_backward(const float* grad_out, long long planes, int h, int w, int nh, int nw, int rot_k,
                                     float* grad_z, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  EOVAE_CHECK(planes >= 1 && h >= 1 && w >= 1 && nh >= 1 && nw >= 1 && rot_k >= 0 && rot_k <= 3, "latent_resize_rot_backward: bad arguments");
  EOVAE_CUDA(cudaMemsetAsync(grad_z, 0, sizeof(float) * planes * h * w, st));
  resize_rot_kernel<true><<<grid_for(planes * nh * nw), 256, 0, st>>>(grad_out, grad_z, planes, h, w, nh, nw, rot_k,
                                                                     static_cast<float>(h) / nh, static_cast<float>(w) / nw);
  EOVAE_LAUNCH_CHECK();
  return 0;
}

int eovae_area_resize_rot(const float* x, long long planes, int h, int w, int nh, int nw, int rot_k, float* out, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  EOVAE_CHECK(planes >= 1 && h >= 1 && w >= 1 && nh >= 1 && nw >= 1 && rot_k >= 0 && rot_k <= 3, "area_resize_rot: bad arguments");
  area_rot_kernel<<<grid_for(planes * nh * nw), 256, 0, st>>>(x, out, planes, h, w, nh, nw, rot_k);
  EOVAE_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
