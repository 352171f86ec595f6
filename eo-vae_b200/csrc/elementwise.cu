// HBM-bound kernels of the EOFluxVAE path: GroupNorm statistics / apply (+SiLU), layout edges, softmax,
// latent normalisation, KL-reparameterisation, pixel losses.  All vectorised to 16-byte accesses along the
// NHWC channel axis; reductions use warp shuffles + a few fp64 atomics per block.
#include "../../include/eovae.h"
#include "common.cuh"
#include "bulk_ring.cuh"
#include "fp32_path.cuh"

#include <type_traits>

namespace {

constexpr int kGnThreads = 256;

// ------------------------------------------------------------------------------------------------- GN stats
// grid (blocks_per_image, n). Each thread owns one 8-channel vector column and walks pixels.
template <typename T>
__global__ void __launch_bounds__(kGnThreads) gn_partial_kernel(const T* __restrict__ x, long long hw, int c,
                                                                long long pix_stride, int groups,
                                                                double* __restrict__ ws, int pix_per_block) {
  extern __shared__ float s_part[];  // [rows][2][c] per-thread partials (deterministic: no atomics anywhere)
  const int vpp = c >> 3;
  const int rows = blockDim.x / vpp;
  const int v = threadIdx.x % vpp;
  const int r = threadIdx.x / vpp;
  const int n = blockIdx.y;
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  const long long p0 = static_cast<long long>(blockIdx.x) * pix_per_block;
  long long p1 = p0 + pix_per_block;
  if (p1 > hw) p1 = hw;
  if (r < rows) {
    const T* base = x + (static_cast<long long>(n) * hw) * pix_stride + v * 8;
    auto add = [&](const uint4& u) {
      const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = T16<T>::to_f2(w4[j]);
        s[2 * j] += f.x;
        q[2 * j] += f.x * f.x;
        s[2 * j + 1] += f.y;
        q[2 * j + 1] += f.y * f.y;
      }
    };
    long long p = p0 + r;
    for (; p + 3LL * rows < p1; p += 4LL * rows) {  // four 16-byte loads in flight per thread
      const uint4 u0 = __ldg(reinterpret_cast<const uint4*>(base + p * pix_stride));
      const uint4 u1 = __ldg(reinterpret_cast<const uint4*>(base + (p + rows) * pix_stride));
      const uint4 u2 = __ldg(reinterpret_cast<const uint4*>(base + (p + 2LL * rows) * pix_stride));
      const uint4 u3 = __ldg(reinterpret_cast<const uint4*>(base + (p + 3LL * rows) * pix_stride));
      add(u0); add(u1); add(u2); add(u3);
    }
    for (; p < p1; p += rows) add(__ldg(reinterpret_cast<const uint4*>(base + p * pix_stride)));
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s_part[(r * 2) * c + v * 8 + j] = s[j];
      s_part[(r * 2 + 1) * c + v * 8 + j] = q[j];
    }
  }
  __syncthreads();
  const int cpg = c / groups;
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    double a = 0.0, b = 0.0;
    for (int rr = 0; rr < rows; ++rr)
      for (int j = 0; j < cpg; ++j) {
        a += static_cast<double>(s_part[(rr * 2) * c + g * cpg + j]);
        b += static_cast<double>(s_part[(rr * 2 + 1) * c + g * cpg + j]);
      }
    double* o = ws + ((static_cast<long long>(n) * gridDim.x + blockIdx.x) * groups + g) * 2;
    o[0] = a;
    o[1] = b;
  }
}

__global__ void gn_finalize_kernel(const double* __restrict__ ws, float* __restrict__ stats, int total, int groups,
                                   int bpi, double count, float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // (n, g)
  if (i >= total) return;
  const int n = i / groups, g = i % groups;
  double sa = 0.0, sb = 0.0;
  for (int b = 0; b < bpi; ++b) {
    const double* o = ws + ((static_cast<long long>(n) * bpi + b) * groups + g) * 2;
    sa += o[0];
    sb += o[1];
  }
  const double mean = sa / count;
  double var = sb / count - mean * mean;
  if (var < 0.0) var = 0.0;
  stats[2 * i] = static_cast<float>(mean);
  stats[2 * i + 1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
}

// ------------------------------------------------------------------------------------------------- GN apply
// UNROLL = independent 16-byte loads in flight per thread.  Two launch shapes: 256 threads x 4 loads when the kernel has the
// SMs to itself, and the "co-resident" shape 128 threads x 8 loads at <= 80 registers (10 K registers per CTA) that fits
// beside a resident 320-thread x 168-register implicit-GEMM CTA, so that this HBM-bound pass of one half of a batch can run
// UNDER the tensor-bound convolution of the other half (dual-stream encode, eo_vae/models/new_autoencoder.py).
template <typename TI, typename TO, bool SILU, int UNROLL>
__device__ __forceinline__ void gn_apply_body(const TI* __restrict__ x, long long x_pix_stride,
                                              const float* __restrict__ stats, const float* __restrict__ gamma,
                                              const float* __restrict__ beta, TO* __restrict__ y, long long y_pix_stride,
                                              long long hw, int c, int groups, int pix_per_block) {
  const int vpp = c >> 3;
  const int rows = blockDim.x / vpp;
  const int v = threadIdx.x % vpp;
  const int r = threadIdx.x / vpp;
  const int n = blockIdx.y;
  if (r >= rows) return;
  const int cpg = c / groups;
  float a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = v * 8 + j;
    const int g = ch / cpg;
    const float mean = stats[(n * groups + g) * 2], rstd = stats[(n * groups + g) * 2 + 1];
    const float ga = gamma[ch] * rstd;
    a[j] = ga;
    b[j] = beta[ch] - mean * ga;
  }
  const long long p0 = static_cast<long long>(blockIdx.x) * pix_per_block;
  long long p1 = p0 + pix_per_block;
  if (p1 > hw) p1 = hw;
  const TI* xb = x + (static_cast<long long>(n) * hw) * x_pix_stride + v * 8;
  TO* yb = y + (static_cast<long long>(n) * hw) * y_pix_stride + v * 8;
  // UNROLL independent 16-byte loads in flight per thread (the kernel is pure HBM streaming: 1 read + 1 write)
  for (long long p = p0 + r; p < p1; p += static_cast<long long>(rows) * UNROLL) {
    uint4 u[UNROLL];
#pragma unroll
    for (int k = 0; k < UNROLL; ++k) {
      const long long pk = p + static_cast<long long>(k) * rows;
      if (pk < p1) u[k] = __ldcs(reinterpret_cast<const uint4*>(xb + pk * x_pix_stride));
    }
#pragma unroll
    for (int k = 0; k < UNROLL; ++k) {
      const long long pk = p + static_cast<long long>(k) * rows;
      if (pk >= p1) break;
      const uint32_t w4[4] = {u[k].x, u[k].y, u[k].z, u[k].w};
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = T16<TI>::to_f2(w4[j]);
        float h0 = fmaf(f.x, a[2 * j], b[2 * j]);
        float h1 = fmaf(f.y, a[2 * j + 1], b[2 * j + 1]);
        if (SILU) {
          h0 = silu_f(h0);
          h1 = silu_f(h1);
        }
        o[j] = T16<TO>::from_f2(h0, h1);
      }
      *reinterpret_cast<uint4*>(yb + pk * y_pix_stride) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

template <typename TI, typename TO, bool SILU>
__global__ void __launch_bounds__(kGnThreads) gn_apply_kernel(const TI* __restrict__ x, long long x_pix_stride,
                                                              const float* __restrict__ stats,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, TO* __restrict__ y,
                                                              long long y_pix_stride, long long hw, int c, int groups,
                                                              int pix_per_block) {
  gn_apply_body<TI, TO, SILU, 4>(x, x_pix_stride, stats, gamma, beta, y, y_pix_stride, hw, c, groups, pix_per_block);
}

// co-resident launch shape: the same body at 128 threads x 8 loads, registers capped so that one CTA fits in the ~11.7 K
// registers a resident implicit-GEMM CTA leaves free on the SM
// Bulk-ring variant (dense input): x streams through a cp.async.bulk shared-memory ring (bulk_ring.cuh: the bytes in flight
// per SM are set by the ring - 3 blocks x 3 outstanding 16 KB slots - not by the loads a thread can keep in registers),
// results leave as coalesced 16-byte stores (y may be a channel slice of a wider tensor).  Measured on B200 (batch 64,
// tools/gn_apply_shapes.py): 6.66 vs 5.65 TB/s on the level-0 tensor, 18-22 % less time on every encoder shape.
// grid (blocks_per_image, n), 256 threads.
template <typename TI, typename TO, bool SILU, int VEC, int STAGES>
__global__ void __launch_bounds__(eovae::kRingThreads, 3) gn_apply_bulk_kernel(const TI* __restrict__ x,
                                                                               const float* __restrict__ stats,
                                                                               const float* __restrict__ gamma,
                                                                               const float* __restrict__ beta,
                                                                               TO* __restrict__ y, long long y_pix_stride,
                                                                               long long hw, int c, int groups,
                                                                               int pix_per_block) {
  extern __shared__ __align__(128) unsigned char ring_smem[];
  using Ring = eovae::BulkRing<1, VEC, STAGES>;
  Ring ring;
  const int vpp = c >> 3;
  const int rows = eovae::kRingThreads / vpp;
  const int v = threadIdx.x % vpp, r = threadIdx.x / vpp;
  const int n = blockIdx.y;
  const int cpg = c / groups;
  const long long p0 = static_cast<long long>(blockIdx.x) * pix_per_block;
  long long p1 = p0 + pix_per_block;
  if (p1 > hw) p1 = hw;
  ring.src[0] = reinterpret_cast<const char*>(x + (static_cast<long long>(n) * hw + p0) * c);
  ring.init(ring_smem, (p1 - p0) * c * 2);
  float a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = v * 8 + j;
    const int g = ch / cpg;
    const float mean = stats[(n * groups + g) * 2], rstd = stats[(n * groups + g) * 2 + 1];
    const float ga = gamma[ch] * rstd;
    a[j] = ga;
    b[j] = beta[ch] - mean * ga;
  }
  // vector t + k * 256 of a slot = pixel r + k * rows of the slot, channels [8 v, 8 v + 8)
  TO* yb = y + (static_cast<long long>(n) * hw + p0 + r) * y_pix_stride + v * 8;
  const long long slot_pix = Ring::kSlotVecs / vpp;
  for (int i = 0; i < ring.nchunks; ++i) {
    ring.wait(i);
    uint4 u[VEC];
    const uint4* sx = ring.slot(i % STAGES, 0);
#pragma unroll
    for (int k = 0; k < VEC; ++k) u[k] = sx[threadIdx.x + k * eovae::kRingThreads];
    const int valid = ring.valid_vecs(i);
    ring.release(i);
    TO* o = yb + i * slot_pix * y_pix_stride;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      if (static_cast<int>(threadIdx.x) + k * eovae::kRingThreads >= valid) break;
      const uint32_t w4[4] = {u[k].x, u[k].y, u[k].z, u[k].w};
      uint32_t q[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = T16<TI>::to_f2(w4[j]);
        float h0 = fmaf(f.x, a[2 * j], b[2 * j]);
        float h1 = fmaf(f.y, a[2 * j + 1], b[2 * j + 1]);
        if (SILU) {
          h0 = silu_f(h0);
          h1 = silu_f(h1);
        }
        q[j] = T16<TO>::from_f2(h0, h1);
      }
      *reinterpret_cast<uint4*>(o + static_cast<long long>(k) * rows * y_pix_stride) = make_uint4(q[0], q[1], q[2], q[3]);
    }
  }
}

template <typename T, bool SILU>
__global__ void __launch_bounds__(128, 6) gn_apply_co_kernel(const T* __restrict__ x, long long x_pix_stride,
                                                                           const float* __restrict__ stats,
                                                                           const float* __restrict__ gamma,
                                                                           const float* __restrict__ beta, T* __restrict__ y,
                                                                           long long y_pix_stride, long long hw, int c,
                                                                           int groups, int pix_per_block) {
  gn_apply_body<T, T, SILU, 8>(x, x_pix_stride, stats, gamma, beta, y, y_pix_stride, hw, c, groups, pix_per_block);
}

int g_gn_apply_coresident = 0;  // eovae_set_tuning(EOVAE_TUNE_GN_APPLY_CORESIDENT, mode): see include/eovae.h
long long g_gn_apply_block_elems = 0;  // eovae_set_tuning(EOVAE_TUNE_GN_APPLY_BLOCK_ELEMS, n): bulk-ring kernel, 0 = auto

int gn_block_threads(int c) {
  const int vpp = c / 8;
  int t = (kGnThreads / vpp) * vpp;
  return t < vpp ? 0 : t;
}

// Pixel partition of one image: depends on (hw, c) only - never on the batch size - so that a patch's statistics
// (hence its latents) are bit-identical whatever batch it is encoded in.
void gn_grid(int n, long long hw, int c, int rows, int* bpi, int* ppb) {
  (void)n;
  long long per = 32768 / c;  // ~64 KB of 16-bit data per block
  if (per < rows) per = rows;
  per = (per + rows - 1) / rows * rows;
  if (per > hw) per = (hw + rows - 1) / rows * rows;
  *ppb = static_cast<int>(per);
  *bpi = static_cast<int>((hw + per - 1) / per);
}

// ------------------------------------------------------------------------------------------------- layout edges
template <typename TO>
__global__ void nchw_to_nhwc16_kernel(const float* __restrict__ x, TO* __restrict__ out, int c, long long hw, int c_pad) {
  const long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  if (p >= hw) return;
  const float* xb = x + static_cast<long long>(n) * c * hw + p;
  TO* ob = out + (static_cast<long long>(n) * hw + p) * c_pad;
  for (int c0 = 0; c0 < c_pad; c0 += 8) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (c0 + j < c) ? __ldg(xb + static_cast<long long>(c0 + j) * hw) : 0.f;
    uint4 o;
    o.x = T16<TO>::from_f2(v[0], v[1]);
    o.y = T16<TO>::from_f2(v[2], v[3]);
    o.z = T16<TO>::from_f2(v[4], v[5]);
    o.w = T16<TO>::from_f2(v[6], v[7]);
    *reinterpret_cast<uint4*>(ob + c0) = o;
  }
}

template <typename TI>
__device__ __forceinline__ float load_as_float(const TI* p) {
  return T16<TI>::to_f(*p);
}
template <>
__device__ __forceinline__ float load_as_float<float>(const float* p) {
  return *p;
}

template <typename TI>
__global__ void nhwc_to_nchw_kernel(const TI* __restrict__ x, long long x_pix_stride, float* __restrict__ out, int c,
                                    long long hw) {
  const long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  if (p >= hw) return;
  const TI* xb = x + (static_cast<long long>(n) * hw + p) * x_pix_stride;
  float* ob = out + static_cast<long long>(n) * c * hw + p;
  for (int ch = 0; ch < c; ++ch) ob[static_cast<long long>(ch) * hw] = load_as_float<TI>(xb + ch);
}

__global__ void upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int h, int w, int c8,
                                  long long total) {
  // one thread per output 16-byte vector
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int v = static_cast<int>(i % c8);
  long long t = i / c8;
  const int ox = static_cast<int>(t % (2 * w));
  t /= (2 * w);
  const int oy = static_cast<int>(t % (2 * h));
  const long long n = t / (2 * h);
  out[i] = __ldg(&x[((n * h + (oy >> 1)) * w + (ox >> 1)) * c8 + v]);
}

// ------------------------------------------------------------------------------------------------- softmax
template <typename TI, typename TO>
__global__ void __launch_bounds__(128) softmax_rows_kernel(const TI* __restrict__ s, TO* __restrict__ p, int cols,
                                                           long long s_ld, long long p_ld) {
  constexpr int MAXV = 32;  // cached values per thread (cols <= 4096)
  // fp32 probabilities = the fp32 validation path: exact expf there, the fast ex2-based __expf for 16-bit outputs
  auto ex = [](float v) { return std::is_same<TO, float>::value ? expf(v) : __expf(v); };
  const long long row = blockIdx.x;
  const TI* sr = s + row * s_ld;
  TO* pr = p + row * p_ld;
  for (long long cidx = cols + threadIdx.x; cidx < p_ld; cidx += 128) pr[cidx] = T16<TO>::from_f(0.f);  // K padding
  __shared__ float red[4];
  float v[MAXV];
  float m = -INFINITY;
  const bool cached = cols <= MAXV * 128;
  if (cached) {
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
      const int cidx = threadIdx.x + j * 128;
      v[j] = cidx < cols ? load_as_float<TI>(sr + cidx) : -INFINITY;
      m = fmaxf(m, v[j]);
    }
  } else {
    for (int cidx = threadIdx.x; cidx < cols; cidx += 128) m = fmaxf(m, load_as_float<TI>(sr + cidx));
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  float sum = 0.f;
  if (cached) {
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
      v[j] = ex(v[j] - m);  // exp(-inf) = 0 for the padding lanes
      sum += v[j];
    }
  } else {
    for (int cidx = threadIdx.x; cidx < cols; cidx += 128) sum += ex(load_as_float<TI>(sr + cidx) - m);
  }
  sum = warp_sum(sum);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  const float inv = 1.f / (red[0] + red[1] + red[2] + red[3]);
  if (cached) {
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
      const int cidx = threadIdx.x + j * 128;
      if (cidx < cols) pr[cidx] = T16<TO>::from_f(v[j] * inv);
    }
  } else {
    for (int cidx = threadIdx.x; cidx < cols; cidx += 128)
      pr[cidx] = T16<TO>::from_f(ex(load_as_float<TI>(sr + cidx) - m) * inv);
  }
}

__global__ void transpose16_kernel(const uint16_t* __restrict__ in, long long in_ld, uint16_t* __restrict__ out,
                                   long long out_ld, int rows, int cols) {
  __shared__ uint16_t tile[32][34];
  const long long b = blockIdx.z;
  const uint16_t* ib = in + b * rows * in_ld;
  uint16_t* ob = out + b * static_cast<long long>(cols) * out_ld;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, cc = c0 + threadIdx.x;
    if (r < rows && cc < cols) tile[j][threadIdx.x] = ib[static_cast<long long>(r) * in_ld + cc];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int cc = c0 + j, r = r0 + threadIdx.x;
    if (cc < cols && r < out_ld) ob[static_cast<long long>(cc) * out_ld + r] = r < rows ? tile[threadIdx.x][j] : uint16_t(0);
  }
}

// Channel-major copies for the transposed-operand weight-gradient kernel, image rows padded to w_pad (a multiple of 8 so
// that a vertical tap shift of +-w_pad pixels keeps the TMA box start 16-byte aligned):
//   out[d][b][c][y*w_pad + xp] = in[b][y*w + xp + d - 1][c]  when xp < w and 0 <= xp + d - 1 < w, else 0,
// for the copies d in [d0, d0 + ncopies) (d = 1 is the unshifted copy; TMA cannot shift the innermost coordinate by one
// element, so the horizontal taps come as separate copies).
__global__ void transpose16_xshift_kernel(const uint16_t* __restrict__ in, long long in_ld, uint16_t* __restrict__ out,
                                          int h, int w, int w_pad, int cols, int d0, int ncopies, long long copy_stride) {
  __shared__ uint16_t tile[34][33];
  const long long b = blockIdx.z;
  const int rows_out = h * w_pad;
  const uint16_t* ib = in + b * static_cast<long long>(h) * w * in_ld;
  uint16_t* ob = out + b * static_cast<long long>(cols) * rows_out;
  const int c0 = blockIdx.x * 32, o0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 34; j += blockDim.y) {
    const int o = o0 + j - 1, cc = c0 + threadIdx.x;
    uint16_t v = 0;
    if (o >= 0 && o < rows_out && cc < cols) {
      const int y = o / w_pad, xp = o % w_pad;
      if (xp < w) v = ib[(static_cast<long long>(y) * w + xp) * in_ld + cc];
    }
    tile[j][threadIdx.x] = v;
  }
  __syncthreads();
  const int o = o0 + threadIdx.x;
  const int xp = o % w_pad;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int cc = c0 + j;
    if (cc >= cols || o >= rows_out) continue;
    for (int i = 0; i < ncopies; ++i) {
      const int d = d0 + i;
      const int xs = xp + d - 1;
      ob[i * copy_stride + static_cast<long long>(cc) * rows_out + o] =
          (xp < w && xs >= 0 && xs < w) ? tile[threadIdx.x + d][j] : uint16_t(0);
    }
  }
}

// ------------------------------------------------------------------------------------------------- latent glue
struct Strides4 { long long n, c, y, x; };  // element strides of a logical [N, C, H, W] tensor

__global__ void latent_norm_kernel(const float* __restrict__ moments, Strides4 ms, const float* __restrict__ rm,
                                   const float* __restrict__ rv, float eps, float* __restrict__ z, int h, int w, int zc,
                                   long long total) {
  // one thread per output element z[n][c][y][x]
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = static_cast<int>(i % w);
  long long t = i / w;
  const int y = static_cast<int>(t % h);
  t /= h;
  const int c = static_cast<int>(t % zc);
  const long long n = t / zc;
  const float mean = __ldg(&moments[n * ms.n + c * ms.c + y * ms.y + x * ms.x]);
  const int c4 = c * 4 + (y & 1) * 2 + (x & 1);  // '(c pi pj)' channel of the packed latent
  z[i] = (mean - rm[c4]) * rsqrtf(rv[c4] + eps);
}

template <typename TO>
__global__ void latent_denorm_kernel(const float* __restrict__ z, const float* __restrict__ rm,
                                     const float* __restrict__ rv, float eps, TO* __restrict__ out, int h, int w, int zc,
                                     long long total) {
  // one thread per (n, y, x, c) of the NHWC output
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = static_cast<int>(i % zc);
  long long t = i / zc;
  const int x = static_cast<int>(t % w);
  t /= w;
  const int y = static_cast<int>(t % h);
  const long long n = t / h;
  const int c4 = c * 4 + (y & 1) * 2 + (x & 1);
  const float v = z[((n * zc + c) * h + y) * w + x] * sqrtf(rv[c4] + eps) + rm[c4];
  out[i] = T16<TO>::from_f(v);
}

__global__ void kl_reparam_kernel(const float* __restrict__ moments, Strides4 ms, const float* __restrict__ eps,
                                  float* __restrict__ z, float* __restrict__ kl, int h, int w, int zc) {
  // one block per image; z NCHW
  const int n = blockIdx.x;
  const long long per = static_cast<long long>(zc) * h * w;
  float acc = 0.f;
  for (long long i = threadIdx.x; i < per; i += blockDim.x) {
    const int x = static_cast<int>(i % w);
    long long t = i / w;
    const int y = static_cast<int>(t % h);
    const int c = static_cast<int>(t / h);
    const float* m = moments + n * ms.n + y * ms.y + x * ms.x;
    const float mean = m[c * ms.c];
    const float lv = fminf(fmaxf(m[(zc + c) * ms.c], -30.f), 20.f);
    const float var = expf(lv);
    acc += mean * mean + var - 1.f - lv;
    if (z != nullptr) {
      const float e = eps != nullptr ? eps[n * per + i] : 0.f;
      z[n * per + i] = mean + expf(0.5f * lv) * e;
    }
  }
  __shared__ float red[32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0 && kl != nullptr) kl[n] = 0.5f * v;
  }
}

// ------------------------------------------------------------------------------------------------- pixel losses
__global__ void l1_char_partial_kernel(const float4* __restrict__ a, const float4* __restrict__ b, long long n4,
                                       const float* __restrict__ a_tail, const float* __restrict__ b_tail, int tail,
                                       float eps2, double* __restrict__ ws) {
  float s1 = 0.f, s2 = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 x = __ldg(&a[i]), y = __ldg(&b[i]);
    const float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
    s1 += fabsf(d0) + fabsf(d1) + fabsf(d2) + fabsf(d3);
    s2 += sqrtf(d0 * d0 + eps2) + sqrtf(d1 * d1 + eps2) + sqrtf(d2 * d2 + eps2) + sqrtf(d3 * d3 + eps2);
  }
  if (blockIdx.x == 0 && threadIdx.x < tail) {
    const float d = a_tail[threadIdx.x] - b_tail[threadIdx.x];
    s1 += fabsf(d);
    s2 += sqrtf(d * d + eps2);
  }
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
__global__ void l1_char_finalize_kernel(const double* ws, double count, float* out) {
  out[0] = static_cast<float>(ws[0] / count);
  out[1] = static_cast<float>(ws[1] / count);
}

// ------------------------------------------------------------------------------------------------- latent statistics
// encode_latents.py:36-109 (RunningStatsButFast): per-channel batch mean / unbiased variance / min / max of an NCHW fp32
// latent batch in ONE pass (one block per channel, fp64 sums), then the parallel-variance merge into the running state
// on the device - no host synchronisation per batch.
__global__ void __launch_bounds__(256) channel_stats_kernel(const float* __restrict__ x, int n, int c, long long hw,
                                                            float* __restrict__ tmp /*[c][4]*/) {
  const int ch = blockIdx.x;
  double s = 0.0, q = 0.0;
  float mn = INFINITY, mx = -INFINITY;
  for (int img = 0; img < n; ++img) {
    const float* p = x + (static_cast<long long>(img) * c + ch) * hw;
    for (long long i = threadIdx.x; i < hw; i += blockDim.x) {
      const float v = p[i];
      s += v;
      q += static_cast<double>(v) * v;
      mn = fminf(mn, v);
      mx = fmaxf(mx, v);
    }
  }
  __shared__ double ss[256], sq[256];
  __shared__ float smn[256], smx[256];
  ss[threadIdx.x] = s; sq[threadIdx.x] = q; smn[threadIdx.x] = mn; smx[threadIdx.x] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0.0, tq = 0.0;
    float tn = INFINITY, tx = -INFINITY;
    for (int i = 0; i < 256; ++i) {  // fixed order: deterministic
      ts += ss[i]; tq += sq[i];
      tn = fminf(tn, smn[i]); tx = fmaxf(tx, smx[i]);
    }
    const double cnt = static_cast<double>(n) * hw;
    const double mean = ts / cnt;
    const double var = cnt > 1.0 ? (tq - cnt * mean * mean) / (cnt - 1.0) : 0.0;  // torch.var default: unbiased
    tmp[4 * ch] = static_cast<float>(mean);
    tmp[4 * ch + 1] = static_cast<float>(var > 0.0 ? var : 0.0);
    tmp[4 * ch + 2] = tn;
    tmp[4 * ch + 3] = tx;
  }
}
// the reference's merge, formula for formula (encode_latents.py:78-94), one thread per channel
__global__ void running_stats_merge_kernel(const float* __restrict__ tmp, int c, float batch_count, float* __restrict__ mean,
                                           float* __restrict__ var, float* __restrict__ std, float* __restrict__ count,
                                           float* __restrict__ vmin, float* __restrict__ vmax) {
  const int ch = threadIdx.x;
  const float cnt = count[0];
  if (ch < c) {
    const float bm = tmp[4 * ch], bv = tmp[4 * ch + 1];
    const float n_ab = cnt + batch_count;
    const float m_a = mean[ch] * cnt, m_b = bm * batch_count;
    const float M2_a = var[ch] * cnt, M2_b = bv * batch_count;
    const float delta = bm - mean[ch];
    mean[ch] = (m_a + m_b) / n_ab;
    const float v = (M2_a + M2_b + delta * delta * cnt * batch_count / (n_ab + 1e-8f)) / n_ab;
    var[ch] = v;
    std[ch] = sqrtf(v + 1e-8f);
    vmin[ch] = fminf(vmin[ch], tmp[4 * ch + 2]);
    vmax[ch] = fmaxf(vmax[ch], tmp[4 * ch + 3]);
  }
  __syncthreads();
  if (ch == 0) count[0] = cnt + batch_count;
}

// ------------------------------------------------------------------------------------------------- data-side prologue
// terramesh_datamodule.py:189-197 (clip + z-score), :476-479 (bilinear resize, align_corners = False), :347-369 (D4
// augmentation: horizontal flip, vertical flip, k x rot90) in ONE gather pass: raw NCHW (fp32 or 16-bit integer DNs) ->
// normalised, resized, augmented NCHW fp32.  The affine normalisation commutes with the (convex) bilinear weights, so the
// clip is applied to the four taps and the z-score to the interpolated value.
template <typename TIn>
__global__ void preprocess_kernel(const TIn* __restrict__ in, int c, int hi, int wi, int hr, int wr, int ho, int wo,
                                  const float* __restrict__ mean, const float* __restrict__ std, float std_eps, int do_clip,
                                  float clip_lo, float clip_hi, int flip_h, int flip_v, int rot_k, float* __restrict__ out,
                                  long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = static_cast<int>(i % wo);
  long long t = i / wo;
  const int y = static_cast<int>(t % ho);
  t /= ho;
  const int ch = static_cast<int>(t % c);
  const long long b = t / c;
  // undo rot90 (torch.rot90 over dims [-2, -1], counter-clockwise): output (y, x) -> position (ry, rx) of the h_r x w_r image
  int ry, rx;
  switch (rot_k & 3) {
    case 1: ry = x; rx = wr - 1 - y; break;
    case 2: ry = hr - 1 - y; rx = wr - 1 - x; break;
    case 3: ry = hr - 1 - x; rx = y; break;
    default: ry = y; rx = x; break;
  }
  if (flip_v) ry = hr - 1 - ry;
  if (flip_h) rx = wr - 1 - rx;
  const TIn* plane = in + (b * c + ch) * static_cast<long long>(hi) * wi;
  auto tap = [&](int yy, int xx) {
    float v = static_cast<float>(plane[static_cast<long long>(yy) * wi + xx]);
    if (do_clip) v = fminf(fmaxf(v, clip_lo), clip_hi);
    return v;
  };
  float v;
  if (hr == hi && wr == wi) {
    v = tap(ry, rx);
  } else {  // F.interpolate(mode='bilinear', align_corners=False)
    const float sy = fmaxf((ry + 0.5f) * (static_cast<float>(hi) / hr) - 0.5f, 0.f);
    const float sx = fmaxf((rx + 0.5f) * (static_cast<float>(wi) / wr) - 0.5f, 0.f);
    const int y0 = min(static_cast<int>(sy), hi - 1), x0 = min(static_cast<int>(sx), wi - 1);
    const int y1 = min(y0 + 1, hi - 1), x1 = min(x0 + 1, wi - 1);
    const float ly = sy - y0, lx = sx - x0;
    v = (1.f - ly) * ((1.f - lx) * tap(y0, x0) + lx * tap(y0, x1)) + ly * ((1.f - lx) * tap(y1, x0) + lx * tap(y1, x1));
  }
  out[i] = (v - mean[ch]) / (std[ch] + std_eps);
}

}  // namespace

// launch-shape knobs of backward.cu, set through eovae_set_tuning
extern long long g_bwd_block_elems;
extern int g_gn_bwd_bulk;

extern "C" {

size_t eovae_gn_stats_workspace_bytes(int n, long long hw, int c, int groups) {
  const int threads = gn_block_threads(c);
  if (threads <= 0) return 0;
  int bpi, ppb;
  gn_grid(n, hw, c, threads / (c / 8), &bpi, &ppb);
  return sizeof(double) * 2 * static_cast<size_t>(n) * bpi * groups;
}

int eovae_gn_stats(const void* x, int x_dtype, int n, long long hw, int c, long long pix_stride, int groups, float eps,
                   float* stats, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (x_dtype == EOVAE_F32) {  // fp32 validation path (fp32_path.cu); needs no workspace
    EOVAE_CHECK(groups > 0 && c % groups == 0, "gn_stats: C (%d) must be a multiple of groups (%d)", c, groups);
    return eovae::f32::gn_stats(static_cast<const float*>(x), n, hw, c, pix_stride, groups, eps, stats, stream);
  }
  EOVAE_CHECK(c % 8 == 0 && c % groups == 0, "gn_stats: C (%d) must be a multiple of 8 and of groups (%d)", c, groups);
  EOVAE_CHECK(pix_stride % 8 == 0, "gn_stats: pixel stride must be a multiple of 8");
  EOVAE_CHECK(x_dtype == EOVAE_BF16 || x_dtype == EOVAE_F16, "gn_stats: x must be 16-bit");
  const int threads = gn_block_threads(c);
  EOVAE_CHECK(threads > 0, "gn_stats: C too large (%d)", c);
  EOVAE_CHECK(workspace_bytes >= eovae_gn_stats_workspace_bytes(n, hw, c, groups), "gn_stats: workspace too small");
  const int rows = threads / (c / 8);
  int bpi, ppb;
  gn_grid(n, hw, c, rows, &bpi, &ppb);
  double* ws = static_cast<double*>(workspace);
  dim3 grid(bpi, n);
  const size_t smem = sizeof(float) * 2 * c * rows;
  if (x_dtype == EOVAE_BF16)
    gn_partial_kernel<__nv_bfloat16><<<grid, threads, smem, stream>>>(static_cast<const __nv_bfloat16*>(x), hw, c, pix_stride, groups, ws, ppb);
  else
    gn_partial_kernel<__half><<<grid, threads, smem, stream>>>(static_cast<const __half*>(x), hw, c, pix_stride, groups, ws, ppb);
  EOVAE_LAUNCH_CHECK();
  const int total = n * groups;
  gn_finalize_kernel<<<ceil_div(total, 128), 128, 0, stream>>>(ws, stats, total, groups, bpi, static_cast<double>(hw) * (c / groups), eps);
  EOVAE_LAUNCH_CHECK();
  return 0;
}

void eovae_set_tuning(int key, int value) {
  if (key == EOVAE_TUNE_GN_APPLY_CORESIDENT) g_gn_apply_coresident = value;
  if (key == EOVAE_TUNE_GN_APPLY_BLOCK_ELEMS && (value == 0 || value >= 2048)) g_gn_apply_block_elems = value;
  if (key == EOVAE_TUNE_GN_BWD_BLOCK_ELEMS && (value == 0 || value >= 4096)) g_bwd_block_elems = value;
  if (key == EOVAE_TUNE_GN_BWD_BULK) g_gn_bwd_bulk = value;
}

int eovae_gn_apply(const void* x, int x_dtype, long long x_pix_stride, const float* stats, const float* gamma,
                   const float* beta, void* y, int y_dtype, long long y_pix_stride, int n, long long hw, int c, int groups,
                   int apply_silu, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (x_dtype == EOVAE_F32 || y_dtype == EOVAE_F32) {  // fp32 validation path
    EOVAE_CHECK(x_dtype == EOVAE_F32 && y_dtype == EOVAE_F32 && groups > 0 && c % groups == 0,
                "gn_apply: the fp32 path needs fp32 input and output and C a multiple of groups");
    return eovae::f32::gn_apply(static_cast<const float*>(x), x_pix_stride, stats, gamma, beta, static_cast<float*>(y),
                                y_pix_stride, n, hw, c, groups, apply_silu, stream);
  }
  EOVAE_CHECK(c % 8 == 0 && c % groups == 0, "gn_apply: C (%d) must be a multiple of 8 and of groups (%d)", c, groups);
  EOVAE_CHECK(x_pix_stride % 8 == 0 && y_pix_stride % 8 == 0, "gn_apply: pixel strides must be multiples of 8");
  EOVAE_CHECK((x_dtype == EOVAE_BF16 || x_dtype == EOVAE_F16) && (y_dtype == EOVAE_BF16 || y_dtype == EOVAE_F16),
              "gn_apply: 16-bit tensors only");
  const int threads = gn_block_threads(c);
  EOVAE_CHECK(threads > 0, "gn_apply: C too large (%d)", c);
  int bpi, ppb;
  gn_grid(n, hw, c, threads / (c / 8), &bpi, &ppb);
  dim3 grid(bpi, n);
  // 128 threads x 8 loads in flight measures 5.66 vs 5.28 TB/s on the large tensors (tools/gn_apply_shapes.py) and loses on
  // the small ones: taken from 32 M elements up, or always when the co-resident shape is requested
  const bool big = static_cast<long long>(n) * hw * c >= (32LL << 20);
  // default: bulk-ring kernel wherever the input is dense and its 16-byte vectors per pixel divide the block
  if ((g_gn_apply_coresident == 0 || g_gn_apply_coresident == 3) && x_pix_stride == c &&
      eovae::kRingThreads % (c / 8) == 0) {
    const long long total = static_cast<long long>(n) * hw * c;
    long long elems = g_gn_apply_block_elems;
    if (elems <= 0) {  // ~8 slots per block on the large tensors, at least ~256 blocks on the small ones
      elems = 8192;
      while (elems < 65536 && elems * 2 * 256 <= total) elems *= 2;
    }
    long long per = elems / c;
    if (per < 1) per = 1;
    if (per > hw) per = hw;
    dim3 g3(static_cast<unsigned>((hw + per - 1) / per), n);
    using Ring = eovae::BulkRing<1, 4, 4>;
#define EOVAE_GN_APPLY_BULK(TI, TO, S)                                                                                 \
  do {                                                                                                                 \
    static bool attr_set = false;                                                                                      \
    if (!attr_set) {                                                                                                   \
      EOVAE_CUDA(cudaFuncSetAttribute(gn_apply_bulk_kernel<TI, TO, S, 4, 4>,                                           \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(Ring::kSmemBytes))); \
      attr_set = true;                                                                                                 \
    }                                                                                                                  \
    gn_apply_bulk_kernel<TI, TO, S, 4, 4><<<g3, eovae::kRingThreads, Ring::kSmemBytes, stream>>>(                      \
        static_cast<const TI*>(x), stats, gamma, beta, static_cast<TO*>(y), y_pix_stride, hw, c, groups,               \
        static_cast<int>(per));                                                                                        \
  } while (0)
    const int bkey = (x_dtype << 2) | (y_dtype << 1) | (apply_silu ? 1 : 0);
    switch (bkey) {
      case 0: EOVAE_GN_APPLY_BULK(__nv_bfloat16, __nv_bfloat16, false); break;
      case 1: EOVAE_GN_APPLY_BULK(__nv_bfloat16, __nv_bfloat16, true); break;
      case 2: EOVAE_GN_APPLY_BULK(__nv_bfloat16, __half, false); break;
      case 3: EOVAE_GN_APPLY_BULK(__nv_bfloat16, __half, true); break;
      case 4: EOVAE_GN_APPLY_BULK(__half, __nv_bfloat16, false); break;
      case 5: EOVAE_GN_APPLY_BULK(__half, __nv_bfloat16, true); break;
      case 6: EOVAE_GN_APPLY_BULK(__half, __half, false); break;
      default: EOVAE_GN_APPLY_BULK(__half, __half, true); break;
    }
#undef EOVAE_GN_APPLY_BULK
    EOVAE_LAUNCH_CHECK();
    return 0;
  }
  if ((g_gn_apply_coresident == 1 || g_gn_apply_coresident == 2 || big) && x_dtype == y_dtype && c <= 1024) {
    // co-resident launch shape (see the kernel comment): 128 threads, 8 loads in flight, capped registers
    const int vpp = c / 8;
    const int thr = (128 / vpp) * vpp;
    if (thr >= vpp && thr > 0) {
      gn_grid(n, hw, c, thr / vpp, &bpi, &ppb);
      dim3 g2(bpi, n);
      // An SM runs CTAs of two kernels side by side only under ONE L1 / shared-memory split.  The implicit GEMM needs the
      // maximum shared-memory carve-out, so this kernel asks for the same split (it uses no shared memory and streams
      // through L2; the smaller L1 costs it nothing) - with the default preference the SM would have to drain first.
#define EOVAE_GN_APPLY_CO(T, S)                                                                                     \
  do {                                                                                                              \
    static int carve_set = -2;                                                                                      \
    const int carve = g_gn_apply_coresident == 1 ? static_cast<int>(cudaSharedmemCarveoutMaxShared)                 \
                                                 : static_cast<int>(cudaSharedmemCarveoutDefault);                  \
    if (carve_set != carve) {                                                                                       \
      EOVAE_CUDA(cudaFuncSetAttribute(gn_apply_co_kernel<T, S>, cudaFuncAttributePreferredSharedMemoryCarveout, carve)); \
      carve_set = carve;                                                                                            \
    }                                                                                                               \
    gn_apply_co_kernel<T, S><<<g2, thr, 0, stream>>>(static_cast<const T*>(x), x_pix_stride, stats, gamma, beta,    \
                                                     static_cast<T*>(y), y_pix_stride, hw, c, groups, ppb);         \
  } while (0)
      if (x_dtype == EOVAE_BF16) { if (apply_silu) EOVAE_GN_APPLY_CO(__nv_bfloat16, true); else EOVAE_GN_APPLY_CO(__nv_bfloat16, false); }
      else { if (apply_silu) EOVAE_GN_APPLY_CO(__half, true); else EOVAE_GN_APPLY_CO(__half, false); }
#undef EOVAE_GN_APPLY_CO
      EOVAE_LAUNCH_CHECK();
      return 0;
    }
  }
#define EOVAE_GN_APPLY(TI, TO, S)                                                                                        \
  gn_apply_kernel<TI, TO, S><<<grid, threads, 0, stream>>>(static_cast<const TI*>(x), x_pix_stride, stats, gamma, beta, \
                                                           static_cast<TO*>(y), y_pix_stride, hw, c, groups, ppb)
  const int key = (x_dtype << 2) | (y_dtype << 1) | (apply_silu ? 1 : 0);
  switch (key) {
    case 0: EOVAE_GN_APPLY(__nv_bfloat16, __nv_bfloat16, false); break;
    case 1: EOVAE_GN_APPLY(__nv_bfloat16, __nv_bfloat16, true); break;
    case 2: EOVAE_GN_APPLY(__nv_bfloat16, __half, false); break;
    case 3: EOVAE_GN_APPLY(__nv_bfloat16, __half, true); break;
    case 4: EOVAE_GN_APPLY(__half, __nv_bfloat16, false); break;
    case 5: EOVAE_GN_APPLY(__half, __nv_bfloat16, true); break;
    case 6: EOVAE_GN_APPLY(__half, __half, false); break;
    default: EOVAE_GN_APPLY(__half, __half, true); break;
  }
#undef EOVAE_GN_APPLY
  EOVAE_LAUNCH_CHECK();
  return 0;
}

int eovae_nchw_to_nhwc16(const float* x, void* out, int n, int c, int h, int w, int c_pad, int out_dtype, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (out_dtype == EOVAE_F32) {  // fp32 validation path
    EOVAE_CHECK(c_pad % 4 == 0 && c_pad >= c, "nchw_to_nhwc16: c_pad (%d) must be a multiple of 4 and >= C (%d)", c_pad, c);
    return eovae::f32::nchw_to_nhwc(x, static_cast<float*>(out), n, c, h, w, c_pad, stream);
  }
  EOVAE_CHECK(c_pad % 8 == 0 && c_pad >= c, "nchw_to_nhwc16: c_pad (%d) must be a multiple of 8 and >= C (%d)", c_pad, c);
  const long long hw = static_cast<long long>(h) * w;
  dim3 grid(static_cast<unsigned>((hw + 255) / 256), n);
  if (out_dtype == EOVAE_BF16)
    nchw_to_nhwc16_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(x, static_cast<__nv_bfloat16*>(out), c, hw, c_pad);
  else if (out_dtype == EOVAE_F16)
    nchw_to_nhwc16_kernel<__half><<<grid, 256, 0, stream>>>(x, static_cast<__half*>(out), c, hw, c_pad);
  else
    EOVAE_CHECK(false, "nchw_to_nhwc16: bad dtype");
  EOVAE_LAUNCH_CHECK();
  return 0;
}

int eovae_nhwc_to_nchw_f32(const void* x, int x_dtype, long long x_pix_stride, float* out, int n, int c, int h, int w,
                           void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const long long hw = static_cast<long long>(h) * w;
  dim3 grid(static_cast<unsigned>((hw + 255) / 256), n);
  if (x_dtype == EOVAE_F32)
    nhwc_to_nchw_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(x), x_pix_stride, out, c, hw);
  else if (x_dtype == EOVAE_BF16)
    nhwc_to_nchw_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), x_pix_stride, out, c, hw);
  else
    nhwc_to_nchw_kernel<__half><<<grid, 256, 0, stream>>>(static_cast<const __half*>(x), x_pix_stride, out, c, hw);
  EOVAE_LAUNCH_CHECK();
  return 0;
}

int eovae_upsample2x(const void* x, void* out, int n, int h, int w, int c, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  EOVAE_CHECK(c % 8 == 0, "upsample2x: C must be a multiple of 8");
  const long long total = static_cast<long long>(n) * 4 * h * w * (c / 8);
  upsample2x_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
      static_cast<const uint4*>(x), static_cast<uint4*>(out), h, w, c / 8, total);
  EOVAE_LAUNCH_CHECK();
  return 0;
}

int eovae_softmax_rows(const void* s, int s_dtype, long long s_ld, void* p, int p_dtype, long long p_ld, long long rows,
                       int cols, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  EOVAE_CHECK(rows > 0 && cols > 0 && rows < (1LL << 31) && s_ld >= cols && p_ld >= cols, "softmax_rows: bad shape");
  const unsigned grid = static_cast<unsigned>(rows);
#define EOVAE_SM(TI, TO) softmax_rows_kernel<TI, TO><<<grid, 128, 0, stream>>>(static_cast<const TI*>(s), static_cast<TO*>(p), cols, s_ld, p_ld)
  if (s_dtype == EOVAE_F32 && p_dtype == EOVAE_BF16) EOVAE_SM(float, __nv_bfloat16);
  else if (s_dtype == EOVAE_F32 && p_dtype == EOVAE_F16) EOVAE_SM(float, __half);
  else if (s_dtype == EOVAE_BF16 && p_dtype == EOVAE_BF16) EOVAE_SM(__nv_bfloat16, __nv_bfloat16);
  else if (s_dtype == EOVAE_F16 && p_dtype == EOVAE_F16) EOVAE_SM(__half, __half);
  else if (s_dtype == EOVAE_F32 && p_dtype == EOVAE_F32) EOVAE_SM(float, float);  // fp32 validation path
  else EOVAE_CHECK(false, "softmax_rows: unsupported dtype pair %d -> %d", s_dtype, p_dtype);
#undef EOVAE_SM
  EOVAE_LAUNCH_CHECK();
  return 0;
}

int eovae_transpose16(const void* in, long long in_ld, void* out, long long out_ld, int batch, int rows, int cols,
                      void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  EOVAE_CHECK(out_ld >= rows && in_ld >= cols, "transpose16: pitches smaller than extents");
  dim3 grid(ceil_div(cols, 32), ceil_div(static_cast<int>(out_ld), 32), batch);
  dim3 block(32, 8);
  transpose16_kernel<<<grid, block, 0, stream>>>(static_cast<const uint16_t*>(in), in_ld, static_cast<uint16_t*>(out), out_ld, rows, cols);
  EOVAE_LAUNCH_CHECK();
  return 0;
}

int eovae_transpose16_xshift(const void* in, long long in_ld, void* out, int batch, int h, int w, int w_pad, int cols,
                             int first_shift, int ncopies, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  EOVAE_CHECK(in_ld >= cols && w_pad >= w, "transpose16_xshift: pitch smaller than extent");
  EOVAE_CHECK(first_shift >= -1 && ncopies >= 1 && first_shift + ncopies <= 2, "transpose16_xshift: shifts must lie in [-1, 1]");
  const int rows = h * w_pad;
  dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32), batch);
  dim3 block(32, 8);
  transpose16_xshift_kernel<<<grid, block, 0, stream>>>(static_cast<const uint16_t*>(in), in_ld, static_cast<uint16_t*>(out), h, w,
                                                        w_pad, cols, first_shift + 1, ncopies,
                                                        static_cast<long long>(batch) * cols * rows);
  EOVAE_LAUNCH_CHECK();
  return 0;
}

int eovae_latent_norm(const float* moments, const long long* mstrides, const float* running_mean,
                      const float* running_var, float eps, float* z, int n, int h, int w, int zc, void* stream_) {
  const Strides4 ms{mstrides[0], mstrides[1], mstrides[2], mstrides[3]};
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  EOVAE_CHECK(h % 2 == 0 && w % 2 == 0, "latent_norm: latent H, W must be even (2x2 packing)");
  const long long total = static_cast<long long>(n) * zc * h * w;
  latent_norm_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(moments, ms, running_mean, running_var, eps, z, h, w, zc, total);
  EOVAE_LAUNCH_CHECK();
  return 0;
}

int eovae_latent_denorm(const float* z, const float* running_mean, const float* running_var, float eps, void* out,
                        int out_dtype, int n, int h, int w, int zc, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  EOVAE_CHECK(h % 2 == 0 && w % 2 == 0, "latent_denorm: latent H, W must be even (2x2 packing)");
  const long long total = static_cast<long long>(n) * zc * h * w;
  const unsigned grid = static_cast<unsigned>((total + 255) / 256);
  if (out_dtype == EOVAE_BF16)
    latent_denorm_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(z, running_mean, running_var, eps, static_cast<__nv_bfloat16*>(out), h, w, zc, total);
  else if (out_dtype == EOVAE_F16)
    latent_denorm_kernel<__half><<<grid, 256, 0, stream>>>(z, running_mean, running_var, eps, static_cast<__half*>(out), h, w, zc, total);
  else if (out_dtype == EOVAE_F32)  // fp32 validation path
    latent_denorm_kernel<float><<<grid, 256, 0, stream>>>(z, running_mean, running_var, eps, static_cast<float*>(out), h, w, zc, total);
  else
    EOVAE_CHECK(false, "latent_denorm: bad dtype");
  EOVAE_LAUNCH_CHECK();
  return 0;
}

int eovae_kl_reparam(const float* moments, const long long* mstrides, const float* eps, float* z, float* kl, int n,
                     int h, int w, int zc, void* stream_) {
  const Strides4 ms{mstrides[0], mstrides[1], mstrides[2], mstrides[3]};
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  kl_reparam_kernel<<<n, 256, 0, stream>>>(moments, ms, eps, z, kl, h, w, zc);
  EOVAE_LAUNCH_CHECK();
  return 0;
}

int eovae_l1_charbonnier(const float* a, const float* b, long long count, float eps, float* out, void* workspace,
                         size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  EOVAE_CHECK(workspace_bytes >= 2 * sizeof(double), "l1_charbonnier: workspace too small");
  EOVAE_CHECK(reinterpret_cast<uintptr_t>(a) % 16 == 0 && reinterpret_cast<uintptr_t>(b) % 16 == 0, "l1_charbonnier: alignment");
  double* ws = static_cast<double*>(workspace);
  EOVAE_CUDA(cudaMemsetAsync(ws, 0, 2 * sizeof(double), stream));
  const long long n4 = count / 4;
  const int tail = static_cast<int>(count - n4 * 4);
  long long blocks = (n4 + 255) / 256;
  const long long cap = 8LL * eovae_num_sms();
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  l1_char_partial_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
      reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b), n4, a + n4 * 4, b + n4 * 4, tail, eps * eps, ws);
  EOVAE_LAUNCH_CHECK();
  l1_char_finalize_kernel<<<1, 1, 0, stream>>>(ws, static_cast<double>(count), out);
  EOVAE_LAUNCH_CHECK();
  return 0;
}

int eovae_running_stats_update(const float* x, int n, int c, long long hw, float* mean, float* var, float* std, float* count,
                               float* vmin, float* vmax, float* workspace, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  EOVAE_CHECK(c >= 1 && c <= 1024 && n >= 1 && hw >= 1, "running_stats_update: bad shape");
  channel_stats_kernel<<<c, 256, 0, stream>>>(x, n, c, hw, workspace);
  EOVAE_LAUNCH_CHECK();
  running_stats_merge_kernel<<<1, round_up(c, 32), 0, stream>>>(workspace, c, static_cast<float>(static_cast<double>(n) * hw), mean, var,
                                                              std, count, vmin, vmax);
  EOVAE_LAUNCH_CHECK();
  return 0;
}

int eovae_preprocess(const void* in, int in_dtype, int n, int c, int hi, int wi, int hr, int wr, const float* mean,
                     const float* std, float std_eps, int do_clip, float clip_lo, float clip_hi, int flip_h, int flip_v, int rot_k,
                     float* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int odd = rot_k & 1;
  const int ho = odd ? wr : hr, wo = odd ? hr : wr;
  const long long total = static_cast<long long>(n) * c * ho * wo;
  const unsigned grid = static_cast<unsigned>((total + 255) / 256);
#define EOVAE_PRE(T)                                                                                                   \
  preprocess_kernel<T><<<grid, 256, 0, stream>>>(static_cast<const T*>(in), c, hi, wi, hr, wr, ho, wo, mean, std, std_eps, \
                                                 do_clip, clip_lo, clip_hi, flip_h, flip_v, rot_k, out, total)
  switch (in_dtype) {
    case 2: EOVAE_PRE(float); break;          /* EOVAE_DT_F32 */
    case 3: EOVAE_PRE(int16_t); break;        /* EOVAE_DT_I16 */
    case 4: EOVAE_PRE(uint16_t); break;       /* EOVAE_DT_U16 */
    default: EOVAE_CHECK(false, "preprocess: input dtype must be fp32, int16 or uint16 (got %d)", in_dtype);
  }
#undef EOVAE_PRE
  EOVAE_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
