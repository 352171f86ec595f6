// fp32 VALIDATION path (north_star: "1e-4 on an fp32 validation path").
//
// The reference's native arithmetic is fp32 (model.py:167-197, train.py:66).  16-bit tensor-core operands cannot reproduce
// it tighter than ~1e-3, and TF32 (10 explicit mantissa bits) only gets ~8x closer than bf16, so this path computes every
// contraction with fp32 FMAs on the SIMT pipes: one implicit-GEMM kernel (3x3 / 1x1 / stride-2 convolutions and the
// batched attention GEMMs through general operand strides), GroupNorm with fp64 statistics, fp32 activations end to end.
// It is selected by eo_vae.set_compute_dtype(torch.float32) -> dtype code EOVAE_F32 at the C ABI; it is a correctness
// instrument (forward / eval only), not a fast path: ~10 TFLOP/s, i.e. ~1 % of the tensor-core path.
#include "../../include/eovae.h"
#include "common.cuh"
#include "fp32_path.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, THREADS = 256;

struct ConvArgs {
  const float* x;
  int N, H, W, C;
  long long x_ps;      // elements between pixels
  int mode;            // EOVAE_CONV_*
  const float* w;      // B operand: element (n = cout, k) at w + img * w_bs + n * w_ns + k * w_ks
  long long w_ns, w_ks, w_bs;
  int kpt, taps;       // K = taps * kpt (kpt = channels per tap incl. zero padding, a multiple of BK)
  int cout;
  const float* bias;
  const float* res;
  long long res_ps;
  float* out;
  long long out_ps;
  float scale;
  int Ho, Wo, tiles_per_img;
};

__global__ void __launch_bounds__(THREADS) conv_f32_kernel(const ConvArgs p) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int img = blockIdx.x / p.tiles_per_img;
  const int m0 = (blockIdx.x % p.tiles_per_img) * BM;
  const int n0 = blockIdx.y * BN;
  const int hw_out = p.Ho * p.Wo;
  const int t = threadIdx.x;
  // loader roles: one float4 of A (pixel lm, channels 4*lq..) and four B scalars (cout ln, k 4*lq..) per K slab
  const int lm = t >> 2, lq = t & 3;
  const int pm = m0 + lm;
  const bool m_ok = pm < hw_out;
  const int oy = m_ok ? pm / p.Wo : 0, ox = m_ok ? pm % p.Wo : 0;
  const float* ximg = p.x + static_cast<long long>(img) * p.H * p.W * p.x_ps;
  const float* wimg = p.w + static_cast<long long>(img) * p.w_bs;
  const int bn = n0 + lm;
  // compute roles: 4 x 4 micro-tile
  const int ty = t >> 4, tx = t & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int slabs_per_tap = p.kpt / BK;
  for (int tap = 0; tap < p.taps; ++tap) {
    int iy, ix;
    if (p.mode == EOVAE_CONV_1X1) {
      iy = oy; ix = ox;
    } else if (p.mode == EOVAE_CONV_3X3) {
      iy = oy + tap / 3 - 1; ix = ox + tap % 3 - 1;
    } else {  // pad (0,1,0,1) then 3x3 stride 2, pad 0 (layers.py:33-37)
      iy = 2 * oy + tap / 3; ix = 2 * ox + tap % 3;
    }
    const bool pix_ok = m_ok && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
    const float* xp = ximg + (static_cast<long long>(iy) * p.W + ix) * p.x_ps;
    for (int s = 0; s < slabs_per_tap; ++s) {
      const int c0 = s * BK + lq * 4;
      float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
      if (pix_ok) {
        if (c0 + 3 < p.C) {
          av = __ldg(reinterpret_cast<const float4*>(xp + c0));
        } else {
          if (c0 < p.C) av.x = __ldg(xp + c0);
          if (c0 + 1 < p.C) av.y = __ldg(xp + c0 + 1);
          if (c0 + 2 < p.C) av.z = __ldg(xp + c0 + 2);
        }
      }
      float bv[4] = {0.f, 0.f, 0.f, 0.f};
      if (bn < p.cout) {
        const long long kb = static_cast<long long>(tap) * p.kpt + c0;
        const float* wp = wimg + static_cast<long long>(bn) * p.w_ns + kb * p.w_ks;
#pragma unroll
        for (int j = 0; j < 4; ++j) bv[j] = __ldg(wp + j * p.w_ks);
      }
      __syncthreads();  // previous slab fully consumed
      As[lq * 4 + 0][lm] = av.x; As[lq * 4 + 1][lm] = av.y; As[lq * 4 + 2][lm] = av.z; As[lq * 4 + 3][lm] = av.w;
#pragma unroll
      for (int j = 0; j < 4; ++j) Bs[lq * 4 + j][lm] = bv[j];
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
  }
  // epilogue: scale * acc + bias + residual
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= hw_out) continue;
    const long long pix = static_cast<long long>(img) * hw_out + m;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.cout) continue;
      float v = acc[i][j] * p.scale;
      if (p.bias != nullptr) v += p.bias[n];
      if (p.res != nullptr) v += p.res[pix * p.res_ps + n];
      p.out[pix * p.out_ps + n] = v;
    }
  }
}

// GroupNorm statistics, fp32 NHWC input: one block per (image, group), fp64 sums (mean, then rstd)
__global__ void __launch_bounds__(256) gn_stats_f32_kernel(const float* __restrict__ x, long long hw, int c, long long ps,
                                                           int groups, float eps, float* __restrict__ stats) {
  const int n = blockIdx.x / groups, g = blockIdx.x % groups;
  const int cpg = c / groups;
  const float* base = x + static_cast<long long>(n) * hw * ps + g * cpg;
  double s = 0.0, q = 0.0;
  for (long long p = threadIdx.x; p < hw; p += 256)
    for (int j = 0; j < cpg; ++j) {
      const double v = static_cast<double>(base[p * ps + j]);
      s += v;
      q += v * v;
    }
  __shared__ double rs[256], rq[256];
  rs[threadIdx.x] = s;
  rq[threadIdx.x] = q;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      rs[threadIdx.x] += rs[threadIdx.x + o];
      rq[threadIdx.x] += rq[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double cnt = static_cast<double>(hw) * cpg;
    const double mean = rs[0] / cnt;
    double var = rq[0] / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[2 * blockIdx.x] = static_cast<float>(mean);
    stats[2 * blockIdx.x + 1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  }
}

// y = [silu]((x - mean) * rstd * gamma + beta): the reference's operation order (F.group_norm then x * sigmoid(x)), exact expf
__global__ void gn_apply_f32_kernel(const float* __restrict__ x, long long x_ps, const float* __restrict__ stats,
                                    const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ y,
                                    long long y_ps, long long hw, int c, int groups, int silu, long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int ch = static_cast<int>(i % c);
  const long long pix = i / c;
  const long long n = pix / hw;
  const int g = ch / (c / groups);
  const float mean = stats[(n * groups + g) * 2], rstd = stats[(n * groups + g) * 2 + 1];
  float v = (x[pix * x_ps + ch] - mean) * rstd * gamma[ch] + beta[ch];
  if (silu) v = v / (1.0f + expf(-v));
  y[pix * y_ps + ch] = v;
}

__global__ void nchw_to_nhwc_f32_kernel(const float* __restrict__ x, float* __restrict__ out, int c, long long hw, int c_pad,
                                        long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;  // over n * hw * c_pad
  if (i >= total) return;
  const int ch = static_cast<int>(i % c_pad);
  const long long pix = i / c_pad;
  const long long n = pix / hw, p = pix % hw;
  out[i] = ch < c ? x[(n * c + ch) * hw + p] : 0.f;
}

}  // namespace

namespace eovae {
namespace f32 {

int conv2d(const float* x, int n, int h, int w, int cin, long long x_ps, int mode, const float* wt, long long w_ns,
           long long w_ks, long long w_bs, int kpt, int cout, const float* bias, const float* res, long long res_ps, float* out,
           long long out_ps, float scale, cudaStream_t stream) {
  EOVAE_CHECK(kpt % BK == 0, "fp32 conv: K per tap (%d) must be a multiple of %d", kpt, BK);
  EOVAE_CHECK(x_ps % 4 == 0 && (reinterpret_cast<uintptr_t>(x) % 16) == 0, "fp32 conv: input pixels must be 16-byte aligned");
  ConvArgs p;
  p.x = x; p.N = n; p.H = h; p.W = w; p.C = cin; p.x_ps = x_ps; p.mode = mode;
  p.w = wt; p.w_ns = w_ns; p.w_ks = w_ks; p.w_bs = w_bs; p.kpt = kpt; p.taps = mode == EOVAE_CONV_1X1 ? 1 : 9;
  p.cout = cout; p.bias = bias; p.res = res; p.res_ps = res_ps; p.out = out; p.out_ps = out_ps; p.scale = scale;
  p.Ho = h; p.Wo = w;
  if (mode == EOVAE_CONV_3X3_S2) {
    p.Ho = (h - 2) / 2 + 1;
    p.Wo = (w - 2) / 2 + 1;
  }
  p.tiles_per_img = ceil_div(p.Ho * p.Wo, BM);
  dim3 grid(static_cast<unsigned>(p.tiles_per_img) * n, ceil_div(cout, BN));
  conv_f32_kernel<<<grid, THREADS, 0, stream>>>(p);
  EOVAE_LAUNCH_CHECK();
  return 0;
}

int gn_stats(const float* x, int n, long long hw, int c, long long ps, int groups, float eps, float* stats, cudaStream_t stream) {
  gn_stats_f32_kernel<<<n * groups, 256, 0, stream>>>(x, hw, c, ps, groups, eps, stats);
  EOVAE_LAUNCH_CHECK();
  return 0;
}

int gn_apply(const float* x, long long x_ps, const float* stats, const float* gamma, const float* beta, float* y, long long y_ps,
             int n, long long hw, int c, int groups, int silu, cudaStream_t stream) {
  const long long total = static_cast<long long>(n) * hw * c;
  gn_apply_f32_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(x, x_ps, stats, gamma, beta, y, y_ps, hw, c,
                                                                                     groups, silu, total);
  EOVAE_LAUNCH_CHECK();
  return 0;
}

int nchw_to_nhwc(const float* x, float* out, int n, int c, int h, int w, int c_pad, cudaStream_t stream) {
  const long long hw = static_cast<long long>(h) * w, total = static_cast<long long>(n) * hw * c_pad;
  nchw_to_nhwc_f32_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(x, out, c, hw, c_pad, total);
  EOVAE_LAUNCH_CHECK();
  return 0;
}

}  // namespace f32
}  // namespace eovae
