// Backward-pass kernels of the HBM-bound ops (first part of the training path): GroupNorm (+SiLU) backward,
// stride-2 gradient scatter, 2x2 gradient pooling (nearest-upsample adjoint), per-channel bias gradient.
// Data gradients of the convolutions reuse the implicit-GEMM kernel with transposed / flipped weight operands
// (eovae_pack_conv_weight_dgrad); weight gradients are in wgrad_sm100.cu.
#include "../../include/eovae.h"
#include "common.cuh"
#include "bulk_ring.cuh"

int g_gn_bwd_bulk = 1;                // eovae_set_tuning(EOVAE_TUNE_GN_BWD_BULK, 0/1): cp.async.bulk staged kernels
long long g_bwd_block_elems = 0;      // eovae_set_tuning(EOVAE_TUNE_GN_BWD_BLOCK_ELEMS, n): pixels x channels per block, 0 = auto

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// d/dz [z * sigmoid(z)] = s + z s (1 - s) with s = sigmoid(z) = 0.5 + 0.5 tanh(z / 2); z comes from one FMA on the raw
// input (the per-channel GroupNorm affine is folded into its coefficients).  ONE MUFU op (tanh.approx.f32, relative
// error 2^-11 - below the bf16 / fp16 rounding of the gradient it multiplies) + 5 FP32 instructions: the ex2 + rcp form
// (2 MUFU + 5 FP32) kept both backward kernels instruction-issue bound at 0.3-0.5 of the HBM rate.
__device__ __forceinline__ float silu_grad2(float z) {
  const float s = fmaf(0.5f, tanh_approx(0.5f * z), 0.5f);
  return fmaf(z * (1.0f - s), s, s);
}

// Pass 1: per (image, channel) sums  A = sum dz,  B = sum dz * xhat   with dz = g * silu'(z) (or g when !SILU),
// z = xhat * gamma + beta, xhat = (x - mean) * rstd.  grid (blocks_per_image, n); fixed-slot partials (deterministic).
// T = storage type of the forward activation x, TG = storage type of the gradients (g in, dx out, optional add)
template <typename T, typename TG, bool SILU>
__global__ void __launch_bounds__(kThreads, 3) gn_bwd_reduce_kernel(const T* __restrict__ x, const TG* __restrict__ g,
                                                                  const float* __restrict__ stats,
                                                                  const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta, long long hw, int c,
                                                                  int groups, float* __restrict__ partial,
                                                                  int pix_per_block) {
  extern __shared__ float sm[];  // [rows][2][c]
  const int vpp = c >> 3;
  const int rows = blockDim.x / vpp;
  const int v = threadIdx.x % vpp, r = threadIdx.x / vpp;
  const int n = blockIdx.y;
  const int cpg = c / groups;
  float sa[8], sb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sa[j] = sb[j] = 0.f;
  if (r < rows) {
    // per-channel constants: xhat = x * rs + nm, z = x * za + zb
    float rs[8], nm[8], za[8], zb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ch = v * 8 + j, gi = ch / cpg;
      const float mean = stats[(n * groups + gi) * 2], rstd = stats[(n * groups + gi) * 2 + 1];
      rs[j] = rstd;
      nm[j] = -mean * rstd;
      za[j] = rstd * gamma[ch];
      zb[j] = fmaf(-mean, za[j], beta[ch]);
    }
    const long long p0 = static_cast<long long>(blockIdx.x) * pix_per_block;
    long long p1 = p0 + pix_per_block;
    if (p1 > hw) p1 = hw;
    const T* xb = x + (static_cast<long long>(n) * hw) * c + v * 8;
    const TG* gb = g + (static_cast<long long>(n) * hw) * c + v * 8;
    auto accum = [&](const uint4& ux, const uint4& ug) {
      const uint32_t wx[4] = {ux.x, ux.y, ux.z, ux.w}, wg[4] = {ug.x, ug.y, ug.z, ug.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 fx = T16<T>::to_f2(wx[j]), fg = T16<TG>::to_f2(wg[j]);
        const float xh0 = fmaf(fx.x, rs[2 * j], nm[2 * j]), xh1 = fmaf(fx.y, rs[2 * j + 1], nm[2 * j + 1]);
        float d0 = fg.x, d1 = fg.y;
        if (SILU) {
          d0 *= silu_grad2(fmaf(fx.x, za[2 * j], zb[2 * j]));
          d1 *= silu_grad2(fmaf(fx.y, za[2 * j + 1], zb[2 * j + 1]));
        }
        sa[2 * j] += d0; sb[2 * j] = fmaf(d0, xh0, sb[2 * j]);
        sa[2 * j + 1] += d1; sb[2 * j + 1] = fmaf(d1, xh1, sb[2 * j + 1]);
      }
    };
    long long p = p0 + r;
    for (; p + 3LL * rows < p1; p += 4LL * rows) {  // four pixels per iteration: eight 16-byte loads in flight per thread
      const uint4 ux0 = __ldg(reinterpret_cast<const uint4*>(xb + p * c));
      const uint4 ug0 = __ldg(reinterpret_cast<const uint4*>(gb + p * c));
      const uint4 ux1 = __ldg(reinterpret_cast<const uint4*>(xb + (p + rows) * c));
      const uint4 ug1 = __ldg(reinterpret_cast<const uint4*>(gb + (p + rows) * c));
      const uint4 ux2 = __ldg(reinterpret_cast<const uint4*>(xb + (p + 2LL * rows) * c));
      const uint4 ug2 = __ldg(reinterpret_cast<const uint4*>(gb + (p + 2LL * rows) * c));
      const uint4 ux3 = __ldg(reinterpret_cast<const uint4*>(xb + (p + 3LL * rows) * c));
      const uint4 ug3 = __ldg(reinterpret_cast<const uint4*>(gb + (p + 3LL * rows) * c));
      accum(ux0, ug0);
      accum(ux1, ug1);
      accum(ux2, ug2);
      accum(ux3, ug3);
    }
    for (; p + rows < p1; p += 2LL * rows) {
      const uint4 ux0 = __ldg(reinterpret_cast<const uint4*>(xb + p * c));
      const uint4 ug0 = __ldg(reinterpret_cast<const uint4*>(gb + p * c));
      const uint4 ux1 = __ldg(reinterpret_cast<const uint4*>(xb + (p + rows) * c));
      const uint4 ug1 = __ldg(reinterpret_cast<const uint4*>(gb + (p + rows) * c));
      accum(ux0, ug0);
      accum(ux1, ug1);
    }
    for (; p < p1; p += rows)
      accum(__ldg(reinterpret_cast<const uint4*>(xb + p * c)), __ldg(reinterpret_cast<const uint4*>(gb + p * c)));
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sm[(r * 2) * c + v * 8 + j] = sa[j];
      sm[(r * 2 + 1) * c + v * 8 + j] = sb[j];
    }
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int rr = 0; rr < rows; ++rr) {
      a += sm[(rr * 2) * c + ch];
      b += sm[(rr * 2 + 1) * c + ch];
    }
    float* o = partial + ((static_cast<long long>(n) * gridDim.x + blockIdx.x) * c + ch) * 2;
    o[0] = a;
    o[1] = b;
  }
}

// Pass 2 (tiny): per (image, channel) totals -> per-(image, group) s1 = sum dz*gamma, s2 = sum dz*gamma*xhat and the
// parameter gradients dgamma[c] += sum_n B, dbeta[c] += sum_n A.   One block per image for the group sums.
__global__ void gn_bwd_finalize_kernel(const float* __restrict__ partial, const float* __restrict__ gamma, int n_img, int bpi,
                                       int c, int groups, int cw, float* __restrict__ gsum /*[n][groups][2]*/,
                                       float* __restrict__ chsum /*[n][c][2]*/) {
  // grid (c / cw, images): a block owns the channel window [c0, c0 + cw) of one image - whole groups (the host picks cw = 32
  // channels when the groups tile it, else cw = c) - so that a batch-16 step launches 64-256 blocks instead of 16
  const int n = blockIdx.y;
  const int c0 = blockIdx.x * cw;
  extern __shared__ float tot[];  // [cw][2] totals, then [slices][cw][2] slice sums
  // the bpi partial rows of a channel are split over `slices` threads (fixed assignment), combined in slice order
  const int slices = cw <= static_cast<int>(blockDim.x) ? static_cast<int>(blockDim.x) / cw : 1;
  float* slice_sum = tot + 2 * cw;
  for (int t = threadIdx.x; t < cw * slices; t += blockDim.x) {
    const int chl = t % cw, sl = t / cw;
    double a = 0.0, b = 0.0;
#pragma unroll 8
    for (int k = sl; k < bpi; k += slices) {  // unrolled: the loads of a batch of rows are issued together
      const float2 o = *reinterpret_cast<const float2*>(partial + ((static_cast<long long>(n) * bpi + k) * c + c0 + chl) * 2);
      a += o.x;
      b += o.y;
    }
    slice_sum[(sl * cw + chl) * 2] = static_cast<float>(a);
    slice_sum[(sl * cw + chl) * 2 + 1] = static_cast<float>(b);
  }
  __syncthreads();
  for (int chl = threadIdx.x; chl < cw; chl += blockDim.x) {
    double a = 0.0, b = 0.0;
    for (int sl = 0; sl < slices; ++sl) {
      a += slice_sum[(sl * cw + chl) * 2];
      b += slice_sum[(sl * cw + chl) * 2 + 1];
    }
    tot[2 * chl] = static_cast<float>(a);
    tot[2 * chl + 1] = static_cast<float>(b);
    chsum[(static_cast<long long>(n) * c + c0 + chl) * 2] = static_cast<float>(a);
    chsum[(static_cast<long long>(n) * c + c0 + chl) * 2 + 1] = static_cast<float>(b);
  }
  __syncthreads();
  const int cpg = c / groups;
  for (int gl = threadIdx.x; gl < cw / cpg; gl += blockDim.x) {
    double s1 = 0.0, s2 = 0.0;
    for (int j = 0; j < cpg; ++j) {
      const int chl = gl * cpg + j;
      s1 += static_cast<double>(tot[2 * chl]) * gamma[c0 + chl];
      s2 += static_cast<double>(tot[2 * chl + 1]) * gamma[c0 + chl];
    }
    const int gi = c0 / cpg + gl;
    gsum[(n * groups + gi) * 2] = static_cast<float>(s1);
    gsum[(n * groups + gi) * 2 + 1] = static_cast<float>(s2);
  }
}
__global__ void gn_bwd_param_kernel(const float* __restrict__ chsum, int n_img, int c, float* __restrict__ dgamma,
                                    float* __restrict__ dbeta, int accumulate) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  double a = 0.0, b = 0.0;
  for (int n = 0; n < n_img; ++n) {
    a += chsum[(static_cast<long long>(n) * c + ch) * 2];
    b += chsum[(static_cast<long long>(n) * c + ch) * 2 + 1];
  }
  dbeta[ch] = (accumulate ? dbeta[ch] : 0.f) + static_cast<float>(a);
  dgamma[ch] = (accumulate ? dgamma[ch] : 0.f) + static_cast<float>(b);
}

// Pass 3: dx = rstd * (dz*gamma - (s1 + xhat*s2) / M) (+ optional accumulation into an existing gradient)
template <typename T, typename TG, bool SILU>
__global__ void __launch_bounds__(kThreads, 3) gn_bwd_apply_kernel(const T* __restrict__ x, const TG* __restrict__ g,
                                                                 const float* __restrict__ stats,
                                                                 const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta,
                                                                 const float* __restrict__ gsum, const TG* __restrict__ add,
                                                                 TG* __restrict__ dx, long long hw, int c, int groups,
                                                                 int pix_per_block, float* __restrict__ colpart) {
  // colpart != NULL: also emit this block's per-channel sums of dx (fixed slot [n][block][c]) - the bias gradient of the
  // convolution that produced x's forward input, for free (no extra pass over dx)
  extern __shared__ float sm_cs[];  // [rows][c], only with colpart
  const int vpp = c >> 3;
  const int rows = blockDim.x / vpp;
  const int v = threadIdx.x % vpp, r = threadIdx.x / vpp;
  const int n = blockIdx.y;
  float cs[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) cs[j] = 0.f;
  if (r < rows) {
  const int cpg = c / groups;
  const float inv_m = 1.0f / (static_cast<float>(hw) * cpg);
  // per-channel constants: z = x * za + zb and, with xhat = (x - mean) * rstd, c1 = rstd * s1 / M, c2 = rstd * s2 / M,
  // dx = dz * za - c1 - xhat * c2 = dz * za + x * xa + xb   (xa = -rstd * c2, xb = mean * rstd * c2 - c1)
  float za[8], zb[8], xa[8], xb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = v * 8 + j, gi = ch / cpg;
    const float mean = stats[(n * groups + gi) * 2], rstd = stats[(n * groups + gi) * 2 + 1];
    za[j] = rstd * gamma[ch];
    zb[j] = fmaf(-mean, za[j], beta[ch]);
    const float c1 = rstd * gsum[(n * groups + gi) * 2] * inv_m;
    const float c2 = rstd * gsum[(n * groups + gi) * 2 + 1] * inv_m;
    xa[j] = -rstd * c2;
    xb[j] = fmaf(mean * rstd, c2, -c1);
  }
  const long long p0 = static_cast<long long>(blockIdx.x) * pix_per_block;
  long long p1 = p0 + pix_per_block;
  if (p1 > hw) p1 = hw;
  const long long base = (static_cast<long long>(n) * hw) * c + v * 8;
  auto compute = [&](const uint4& ux, const uint4& ug, const uint4& ua) {
    const uint32_t wx[4] = {ux.x, ux.y, ux.z, ux.w}, wg[4] = {ug.x, ug.y, ug.z, ug.w}, wa[4] = {ua.x, ua.y, ua.z, ua.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 fx = T16<T>::to_f2(wx[j]), fg = T16<TG>::to_f2(wg[j]);
      float2 fa = make_float2(0.f, 0.f);
      if (add != nullptr) fa = T16<TG>::to_f2(wa[j]);
      float d0 = fg.x, d1 = fg.y;
      if (SILU) {
        d0 *= silu_grad2(fmaf(fx.x, za[2 * j], zb[2 * j]));
        d1 *= silu_grad2(fmaf(fx.y, za[2 * j + 1], zb[2 * j + 1]));
      }
      const float r0 = fmaf(fx.x, xa[2 * j], fmaf(d0, za[2 * j], fa.x + xb[2 * j]));
      const float r1 = fmaf(fx.y, xa[2 * j + 1], fmaf(d1, za[2 * j + 1], fa.y + xb[2 * j + 1]));
      cs[2 * j] += r0;
      cs[2 * j + 1] += r1;
      o[j] = T16<TG>::from_f2(r0, r1);
    }
    return make_uint4(o[0], o[1], o[2], o[3]);
  };
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  long long p = p0 + r;
  for (; p + rows < p1; p += 2LL * rows) {  // two pixels per iteration: up to six 16-byte loads in flight per thread
    const long long o0 = base + p * c, o1 = base + (p + rows) * c;
    const uint4 ux0 = __ldg(reinterpret_cast<const uint4*>(x + o0)), ug0 = __ldg(reinterpret_cast<const uint4*>(g + o0));
    const uint4 ux1 = __ldg(reinterpret_cast<const uint4*>(x + o1)), ug1 = __ldg(reinterpret_cast<const uint4*>(g + o1));
    uint4 ua0 = zero4, ua1 = zero4;
    if (add != nullptr) {
      ua0 = __ldg(reinterpret_cast<const uint4*>(add + o0));
      ua1 = __ldg(reinterpret_cast<const uint4*>(add + o1));
    }
    *reinterpret_cast<uint4*>(dx + o0) = compute(ux0, ug0, ua0);
    *reinterpret_cast<uint4*>(dx + o1) = compute(ux1, ug1, ua1);
  }
  for (; p < p1; p += rows) {
    const long long o0 = base + p * c;
    const uint4 ua0 = add != nullptr ? __ldg(reinterpret_cast<const uint4*>(add + o0)) : zero4;
    *reinterpret_cast<uint4*>(dx + o0) =
        compute(__ldg(reinterpret_cast<const uint4*>(x + o0)), __ldg(reinterpret_cast<const uint4*>(g + o0)), ua0);
  }
  }  // r < rows
  if (colpart == nullptr) return;
  if (r < rows) {
#pragma unroll
    for (int j = 0; j < 8; ++j) sm_cs[r * c + v * 8 + j] = cs[j];
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float a = 0.f;
    for (int rr = 0; rr < rows; ++rr) a += sm_cs[rr * c + ch];
    colpart[(static_cast<long long>(n) * gridDim.x + blockIdx.x) * c + ch] = a;
  }
}

void bwd_grid(int n, long long hw, int c, int rows, int* bpi, int* ppb) {
  // elements (pixels x channels) per block: enough work to amortise the per-block parameter loads, the pipeline fill and
  // the shared-memory reduction, small enough that the grid still covers the 148 SMs.  Measured on B200 (batch 16,
  // tools/gn_bwd_grid.py): 128 Ki elements from 32 Mi-element tensors up, 64 Ki at 8 Mi; knob > 0 overrides.
  long long elems = g_bwd_block_elems;
  if (elems <= 0) {
    const long long total = static_cast<long long>(n) * hw * c;
    elems = 16384;
    while (elems < 131072 && elems * 2 * 128 <= total) elems *= 2;
  }
  long long per = elems / c;
  if (per < rows) per = rows;
  per = (per + rows - 1) / rows * rows;
  if (per > hw) per = (hw + rows - 1) / rows * rows;
  *ppb = static_cast<int>(per);
  *bpi = static_cast<int>((hw + per - 1) / per);
}

// dY [n][ho][wo][c] -> Z [n][h][w][c], zero except Z[2i+1][2j+1] = dY[i][j]: turns the data gradient of the
// pad(0,1,0,1) + stride-2 conv into a stride-1 pad-1 conv with the flipped kernel.
__global__ void scatter_s2_kernel(const uint4* __restrict__ dy, uint4* __restrict__ z, int ho, int wo, int h, int w, int c8,
                                  long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int v = static_cast<int>(i % c8);
  long long t = i / c8;
  const int x = static_cast<int>(t % w);
  t /= w;
  const int y = static_cast<int>(t % h);
  const long long n = t / h;
  uint4 val = make_uint4(0, 0, 0, 0);
  if ((x & 1) && (y & 1) && (x >> 1) < wo && (y >> 1) < ho) val = __ldg(&dy[((n * ho + (y >> 1)) * wo + (x >> 1)) * c8 + v]);
  z[i] = val;
}

// adjoint of nearest x2 upsampling: out[n][i][j][c] = sum of the 2x2 block of g
template <typename T>
__global__ void pool2x2_sum_kernel(const T* __restrict__ g, T* __restrict__ out, int h, int w, int c8, long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int v = static_cast<int>(i % c8);
  long long t = i / c8;
  const int x = static_cast<int>(t % w);
  t /= w;
  const int y = static_cast<int>(t % h);
  const long long n = t / h;
  float acc[8] = {};
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(g) + ((n * 2 * h + 2 * y + dy) * 2 * w + 2 * x + dx) * c8 + v);
      const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = T16<T>::to_f2(w4[j]);
        acc[2 * j] += f.x;
        acc[2 * j + 1] += f.y;
      }
    }
  uint4 o;
  o.x = T16<T>::from_f2(acc[0], acc[1]);
  o.y = T16<T>::from_f2(acc[2], acc[3]);
  o.z = T16<T>::from_f2(acc[4], acc[5]);
  o.w = T16<T>::from_f2(acc[6], acc[7]);
  reinterpret_cast<uint4*>(out)[i] = o;
}

// per-channel sum over all pixels of an NHWC 16-bit tensor (bias gradient); fixed-slot partials
template <typename T>
__global__ void __launch_bounds__(kThreads) colsum_partial_kernel(const T* __restrict__ g, long long pixels, int c,
                                                                   float* __restrict__ partial, long long pix_per_block) {
  extern __shared__ float sm[];  // [rows][c]
  const int vpp = c >> 3;
  const int rows = blockDim.x / vpp;
  const int v = threadIdx.x % vpp, r = threadIdx.x / vpp;
  float s[8] = {};
  if (r < rows) {
    const long long p0 = blockIdx.x * pix_per_block;
    long long p1 = p0 + pix_per_block;
    if (p1 > pixels) p1 = pixels;
    auto add = [&](const uint4& u) {
      const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = T16<T>::to_f2(w4[j]);
        s[2 * j] += f.x;
        s[2 * j + 1] += f.y;
      }
    };
    long long p = p0 + r;
    for (; p + 3LL * rows < p1; p += 4LL * rows) {  // four independent 16-byte loads in flight per thread
      const uint4 u0 = __ldg(reinterpret_cast<const uint4*>(g + p * c + v * 8));
      const uint4 u1 = __ldg(reinterpret_cast<const uint4*>(g + (p + rows) * c + v * 8));
      const uint4 u2 = __ldg(reinterpret_cast<const uint4*>(g + (p + 2LL * rows) * c + v * 8));
      const uint4 u3 = __ldg(reinterpret_cast<const uint4*>(g + (p + 3LL * rows) * c + v * 8));
      add(u0); add(u1); add(u2); add(u3);
    }
    for (; p < p1; p += rows) add(__ldg(reinterpret_cast<const uint4*>(g + p * c + v * 8)));
#pragma unroll
    for (int j = 0; j < 8; ++j) sm[r * c + v * 8 + j] = s[j];
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float a = 0.f;
    for (int rr = 0; rr < rows; ++rr) a += sm[rr * c + ch];
    partial[static_cast<long long>(blockIdx.x) * c + ch] = a;
  }
}
// block (32 channels, 32 slices): slice y sums the partial rows b = y, y + 32, ...; fixed-order combine (deterministic)
__global__ void colsum_finalize_kernel(const float* __restrict__ partial, int blocks, int c, float* __restrict__ out,
                                       int accumulate) {
  __shared__ double red[32][32];
  const int ch = blockIdx.x * 32 + threadIdx.x;
  double a = 0.0;
  if (ch < c)
    for (int b = threadIdx.y; b < blocks; b += 32) a += partial[static_cast<long long>(b) * c + ch];
  red[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0 && ch < c) {
    double t = 0.0;
#pragma unroll
    for (int y = 0; y < 32; ++y) t += red[y][threadIdx.x];
    out[ch] = (accumulate ? out[ch] : 0.f) + static_cast<float>(t);
  }
}

// ------------------------------------------------------------------------------------ bulk-async staged GroupNorm backward
// Same two passes, but x and g reach the SM through cp.async.bulk (TMA, 1-D) into a shared-memory ring instead of register
// loads: the bytes in flight per SM no longer depend on the register allocator (at 80 registers / thread ptxas serialised
// the loads of the register version down to two 16-byte requests per thread = 24 KB / SM - long-scoreboard bound at 3.5 of
// 6.5 TB/s; the ring keeps 128-192 KB / SM outstanding).  A block owns the contiguous pixel range [p0, p1) of one image =
// one contiguous byte range of each NHWC tensor, walked in stages of kThreads * VEC 16-byte vectors; thread t owns vectors
// t, t + 256, ... of a stage, which all belong to the same 8 channels because (C / 8) divides 256.
template <typename T, typename TG, bool SILU, int VEC, int STAGES>
__global__ void __launch_bounds__(kThreads, 2) gn_bwd_reduce_bulk_kernel(const T* __restrict__ x, const TG* __restrict__ g,
                                                                         const float* __restrict__ stats,
                                                                         const float* __restrict__ gamma,
                                                                         const float* __restrict__ beta, long long hw, int c,
                                                                         int groups, float* __restrict__ partial,
                                                                         int pix_per_block) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  using Ring = eovae::BulkRing<2, VEC, STAGES>;
  Ring ring;
  const int vpp = c >> 3;
  const int rows = kThreads / vpp;
  const int v = threadIdx.x % vpp, r = threadIdx.x / vpp;
  const int n = blockIdx.y;
  const int cpg = c / groups;
  const long long p0 = static_cast<long long>(blockIdx.x) * pix_per_block;
  long long p1 = p0 + pix_per_block;
  if (p1 > hw) p1 = hw;
  const long long base = (static_cast<long long>(n) * hw + p0) * c;
  ring.src[0] = reinterpret_cast<const char*>(x + base);
  ring.src[1] = reinterpret_cast<const char*>(g + base);
  ring.init(smem_raw, (p1 - p0) * c * 2);
  float sa[8], sb[8], rs[8], nm[8], za[8], zb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = v * 8 + j, gi = ch / cpg;
    const float mean = stats[(n * groups + gi) * 2], rstd = stats[(n * groups + gi) * 2 + 1];
    sa[j] = sb[j] = 0.f;
    rs[j] = rstd;
    nm[j] = -mean * rstd;
    za[j] = rstd * gamma[ch];
    zb[j] = fmaf(-mean, za[j], beta[ch]);
  }
  auto accum = [&](const uint4& ux, const uint4& ug) {
    const uint32_t wx[4] = {ux.x, ux.y, ux.z, ux.w}, wg[4] = {ug.x, ug.y, ug.z, ug.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 fx = T16<T>::to_f2(wx[j]), fg = T16<TG>::to_f2(wg[j]);
      const float xh0 = fmaf(fx.x, rs[2 * j], nm[2 * j]), xh1 = fmaf(fx.y, rs[2 * j + 1], nm[2 * j + 1]);
      float d0 = fg.x, d1 = fg.y;
      if (SILU) {
        d0 *= silu_grad2(fmaf(fx.x, za[2 * j], zb[2 * j]));
        d1 *= silu_grad2(fmaf(fx.y, za[2 * j + 1], zb[2 * j + 1]));
      }
      sa[2 * j] += d0; sb[2 * j] = fmaf(d0, xh0, sb[2 * j]);
      sa[2 * j + 1] += d1; sb[2 * j + 1] = fmaf(d1, xh1, sb[2 * j + 1]);
    }
  };
  for (int i = 0; i < ring.nchunks; ++i) {
    const int s = i % STAGES;
    ring.wait(i);
    uint4 ux[VEC], ug[VEC];
    const uint4 *sx = ring.slot(s, 0), *sg = ring.slot(s, 1);
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      ux[k] = sx[threadIdx.x + k * kThreads];
      ug[k] = sg[threadIdx.x + k * kThreads];
    }
    const int valid = ring.valid_vecs(i);
    ring.release(i);
#pragma unroll
    for (int k = 0; k < VEC; ++k)
      if (static_cast<int>(threadIdx.x) + k * kThreads < valid) accum(ux[k], ug[k]);
  }
  __syncthreads();  // the ring is idle (every issued chunk was consumed): reuse it for the block reduction
  float* sm = reinterpret_cast<float*>(smem_raw);  // [rows][2][c]
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sm[(r * 2) * c + v * 8 + j] = sa[j];
    sm[(r * 2 + 1) * c + v * 8 + j] = sb[j];
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int rr = 0; rr < rows; ++rr) {
      a += sm[(rr * 2) * c + ch];
      b += sm[(rr * 2 + 1) * c + ch];
    }
    float* o = partial + ((static_cast<long long>(n) * gridDim.x + blockIdx.x) * c + ch) * 2;
    o[0] = a;
    o[1] = b;
  }
}

template <typename T, typename TG, bool SILU, bool ADD, int VEC, int STAGES>
__global__ void __launch_bounds__(kThreads, 2) gn_bwd_apply_bulk_kernel(const T* __restrict__ x, const TG* __restrict__ g,
                                                                        const float* __restrict__ stats,
                                                                        const float* __restrict__ gamma,
                                                                        const float* __restrict__ beta,
                                                                        const float* __restrict__ gsum,
                                                                        const TG* __restrict__ add, TG* __restrict__ dx,
                                                                        long long hw, int c, int groups, int pix_per_block,
                                                                        float* __restrict__ colpart) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int NT = ADD ? 3 : 2;
  using Ring = eovae::BulkRing<NT, VEC, STAGES>;
  Ring ring;
  const int vpp = c >> 3;
  const int rows = kThreads / vpp;
  const int v = threadIdx.x % vpp, r = threadIdx.x / vpp;
  const int n = blockIdx.y;
  const int cpg = c / groups;
  const long long p0 = static_cast<long long>(blockIdx.x) * pix_per_block;
  long long p1 = p0 + pix_per_block;
  if (p1 > hw) p1 = hw;
  const long long base = (static_cast<long long>(n) * hw + p0) * c;
  ring.src[0] = reinterpret_cast<const char*>(x + base);
  ring.src[1] = reinterpret_cast<const char*>(g + base);
  if (ADD) ring.src[NT - 1] = reinterpret_cast<const char*>(add + base);
  ring.init(smem_raw, (p1 - p0) * c * 2);
  const float inv_m = 1.0f / (static_cast<float>(hw) * cpg);
  // z = x * za + zb;  dx = dz * za - c1 - xhat * c2 = dz * za + x * xa + xb  (see gn_bwd_apply_kernel)
  float cs[8], za[8], zb[8], xa[8], xb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = v * 8 + j, gi = ch / cpg;
    const float mean = stats[(n * groups + gi) * 2], rstd = stats[(n * groups + gi) * 2 + 1];
    cs[j] = 0.f;
    za[j] = rstd * gamma[ch];
    zb[j] = fmaf(-mean, za[j], beta[ch]);
    const float c1 = rstd * gsum[(n * groups + gi) * 2] * inv_m;
    const float c2 = rstd * gsum[(n * groups + gi) * 2 + 1] * inv_m;
    xa[j] = -rstd * c2;
    xb[j] = fmaf(mean * rstd, c2, -c1);
  }
  auto compute = [&](const uint4& ux, const uint4& ug, const uint4& ua) {
    const uint32_t wx[4] = {ux.x, ux.y, ux.z, ux.w}, wg[4] = {ug.x, ug.y, ug.z, ug.w}, wa[4] = {ua.x, ua.y, ua.z, ua.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 fx = T16<T>::to_f2(wx[j]), fg = T16<TG>::to_f2(wg[j]);
      float2 fa = make_float2(0.f, 0.f);
      if (ADD) fa = T16<TG>::to_f2(wa[j]);
      float d0 = fg.x, d1 = fg.y;
      if (SILU) {
        d0 *= silu_grad2(fmaf(fx.x, za[2 * j], zb[2 * j]));
        d1 *= silu_grad2(fmaf(fx.y, za[2 * j + 1], zb[2 * j + 1]));
      }
      const float r0 = fmaf(fx.x, xa[2 * j], fmaf(d0, za[2 * j], fa.x + xb[2 * j]));
      const float r1 = fmaf(fx.y, xa[2 * j + 1], fmaf(d1, za[2 * j + 1], fa.y + xb[2 * j + 1]));
      cs[2 * j] += r0;
      cs[2 * j + 1] += r1;
      o[j] = T16<TG>::from_f2(r0, r1);
    }
    return make_uint4(o[0], o[1], o[2], o[3]);
  };
  uint4* out = reinterpret_cast<uint4*>(dx + base);
  for (int i = 0; i < ring.nchunks; ++i) {
    const int s = i % STAGES;
    ring.wait(i);
    uint4 ux[VEC], ug[VEC], ua[VEC];
    const uint4 *sx = ring.slot(s, 0), *sg = ring.slot(s, 1), *sa = ring.slot(s, NT - 1);
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      ux[k] = sx[threadIdx.x + k * kThreads];
      ug[k] = sg[threadIdx.x + k * kThreads];
      ua[k] = ADD ? sa[threadIdx.x + k * kThreads] : make_uint4(0, 0, 0, 0);
    }
    const int valid = ring.valid_vecs(i);
    ring.release(i);
    uint4* o = out + static_cast<long long>(i) * Ring::kSlotVecs;
#pragma unroll
    for (int k = 0; k < VEC; ++k)
      if (static_cast<int>(threadIdx.x) + k * kThreads < valid) o[threadIdx.x + k * kThreads] = compute(ux[k], ug[k], ua[k]);
  }
  if (colpart == nullptr) return;
  __syncthreads();
  float* sm_cs = reinterpret_cast<float*>(smem_raw);  // [rows][c]
#pragma unroll
  for (int j = 0; j < 8; ++j) sm_cs[r * c + v * 8 + j] = cs[j];
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float a = 0.f;
    for (int rr = 0; rr < rows; ++rr) a += sm_cs[rr * c + ch];
    colpart[(static_cast<long long>(n) * gridDim.x + blockIdx.x) * c + ch] = a;
  }
}

int block_threads(int c) {
  const int vpp = c / 8;
  const int t = (kThreads / vpp) * vpp;
  return t < vpp ? 0 : t;
}

struct Strides4 { long long n, c, y, x; };

// ------------------------------------------------------------------------------------------------- attention / latent / loss
// dS = scale * P o (dP - rowsum(dP o P)): one warp per row; P 16-bit, dP fp32, dS 16-bit with columns >= cols zeroed.
template <typename T, typename TG>
__global__ void softmax_bwd_kernel(const T* __restrict__ p, long long p_ld, const float* __restrict__ dp, long long dp_ld,
                                   TG* __restrict__ ds, long long ds_ld, long long rows, int cols, int out_cols, float scale) {
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const T* pr = p + row * p_ld;
  const float* dr = dp + row * dp_ld;
  float dot = 0.f;
  for (int j = lane; j < cols; j += 32) dot += T16<T>::to_f(pr[j]) * dr[j];
  dot = warp_sum(dot);
  TG* o = ds + row * ds_ld;
  for (int j = lane; j < out_cols; j += 32)
    o[j] = j < cols ? T16<TG>::from_f(scale * T16<T>::to_f(pr[j]) * (dr[j] - dot)) : T16<TG>::from_f(0.f);
}

// z = mean + exp(0.5 * clamp(logvar)) * eps  ->  dmean = dz, dlogvar = dz * eps * 0.5 * std inside the clamp, else 0.
// dmoments: dense NCHW fp32 [n][2zc][h][w].
__global__ void reparam_bwd_kernel(const float* __restrict__ moments, Strides4 ms, const float* __restrict__ eps,
                                   const float* __restrict__ dz, float* __restrict__ dmoments, int h, int w, int zc,
                                   long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;  // over dz [n][zc][h][w]
  if (i >= total) return;
  const int x = static_cast<int>(i % w);
  long long t = i / w;
  const int y = static_cast<int>(t % h);
  t /= h;
  const int c = static_cast<int>(t % zc);
  const long long n = t / zc;
  const float lv = moments[n * ms.n + (zc + c) * ms.c + y * ms.y + x * ms.x];
  const float g = dz[i];
  const long long plane = static_cast<long long>(h) * w;
  const long long o = (n * 2 * zc + c) * plane + static_cast<long long>(y) * w + x;
  dmoments[o] = g;
  const bool inside = lv >= -30.f && lv <= 20.f;
  dmoments[o + zc * plane] = inside ? g * eps[i] * 0.5f * expf(0.5f * lv) : 0.f;
}

// d/da of mean|a-b| (kind 0) or mean sqrt((a-b)^2 + eps^2) (kind 1), times the upstream scalar *gscale.
__global__ void pixel_loss_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, long long count, float eps2,
                                      int kind, const float* __restrict__ gscale, float* __restrict__ ga) {
  const float k = gscale[0] / static_cast<float>(count);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < count;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float d = a[i] - b[i];
    ga[i] = k * (kind == 0 ? (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) : d * rsqrtf(d * d + eps2));
  }
}

// ------------------------------------------------------------------------------------------------- train-mode latent glue
// new_autoencoder.py:466-469,533-543 in TRAIN mode, one block per packed channel k = (c, pi, pj) of the 2x2-unshuffled
// latent: batch statistics (biased variance, eps_bn) -> running-statistics update (momentum, unbiased variance) ->
// normalise -> inverse normalisation with the UPDATED running statistics (eps_inv) -> pixel shuffle -> NHWC 16-bit decoder
// input.  Everything the reference does with ~8 tiny ATen kernels and three tensor round trips.
template <typename T>
__global__ void __launch_bounds__(256) latent_bn_train_fwd_kernel(const float* __restrict__ z, int n, int zc, int h, int w,
                                                                  float* __restrict__ running_mean,
                                                                  float* __restrict__ running_var, float momentum,
                                                                  float eps_bn, float eps_inv, T* __restrict__ out,
                                                                  long long out_pitch, float* __restrict__ save /*[4zc][3]*/) {
  const int k = blockIdx.x, c = k >> 2, pi = (k >> 1) & 1, pj = k & 1;
  const int h2 = h >> 1, w2 = w >> 1;
  const long long per = static_cast<long long>(n) * h2 * w2;
  __shared__ double red[2][256];
  double s = 0.0, q = 0.0;
  for (long long i = threadIdx.x; i < per; i += blockDim.x) {
    const int j2 = static_cast<int>(i % w2);
    const int i2 = static_cast<int>((i / w2) % h2);
    const long long img = i / (static_cast<long long>(w2) * h2);
    const float v = z[((img * zc + c) * h + 2 * i2 + pi) * w + 2 * j2 + pj];
    s += v;
    q += static_cast<double>(v) * v;
  }
  red[0][threadIdx.x] = s;
  red[1][threadIdx.x] = q;
  __syncthreads();
  __shared__ float sh[4];
  if (threadIdx.x == 0) {
    double ts = 0.0, tq = 0.0;
    for (int i = 0; i < 256; ++i) { ts += red[0][i]; tq += red[1][i]; }
    const double mean = ts / per;
    double var = tq / per - mean * mean;
    if (var < 0.0) var = 0.0;
    const double unbiased = per > 1 ? var * per / (per - 1) : var;
    const float rm = (1.f - momentum) * running_mean[k] + momentum * static_cast<float>(mean);
    const float rv = (1.f - momentum) * running_var[k] + momentum * static_cast<float>(unbiased);
    running_mean[k] = rm;
    running_var[k] = rv;
    const float rstd = rsqrtf(static_cast<float>(var) + eps_bn);
    const float scale = sqrtf(rv + eps_inv);
    sh[0] = static_cast<float>(mean); sh[1] = rstd; sh[2] = scale; sh[3] = rm;
    save[3 * k] = static_cast<float>(mean);
    save[3 * k + 1] = rstd;
    save[3 * k + 2] = scale;
  }
  __syncthreads();
  const float mean = sh[0], rstd = sh[1], scale = sh[2], rm = sh[3];
  for (long long i = threadIdx.x; i < per; i += blockDim.x) {
    const int j2 = static_cast<int>(i % w2);
    const int i2 = static_cast<int>((i / w2) % h2);
    const long long img = i / (static_cast<long long>(w2) * h2);
    const int y = 2 * i2 + pi, x = 2 * j2 + pj;
    const float v = z[((img * zc + c) * h + y) * w + x];
    out[((img * h + y) * w + x) * out_pitch + c] = T16<T>::from_f((v - mean) * rstd * scale + rm);
  }
}

// adjoint: dzn = dout * scale, then the BatchNorm backward dz = rstd * (dzn - mean(dzn) - zn * mean(dzn * zn))
template <typename T>
__global__ void __launch_bounds__(256) latent_bn_train_bwd_kernel(const T* __restrict__ dout, long long dout_pitch,
                                                                  const float* __restrict__ z, int n, int zc, int h, int w,
                                                                  const float* __restrict__ save, float* __restrict__ dz) {
  const int k = blockIdx.x, c = k >> 2, pi = (k >> 1) & 1, pj = k & 1;
  const int h2 = h >> 1, w2 = w >> 1;
  const long long per = static_cast<long long>(n) * h2 * w2;
  const float mean = save[3 * k], rstd = save[3 * k + 1], scale = save[3 * k + 2];
  __shared__ double red[2][256];
  double a = 0.0, b = 0.0;
  for (long long i = threadIdx.x; i < per; i += blockDim.x) {
    const int j2 = static_cast<int>(i % w2);
    const int i2 = static_cast<int>((i / w2) % h2);
    const long long img = i / (static_cast<long long>(w2) * h2);
    const int y = 2 * i2 + pi, x = 2 * j2 + pj;
    const float g = T16<T>::to_f(dout[((img * h + y) * w + x) * dout_pitch + c]) * scale;
    const float zn = (z[((img * zc + c) * h + y) * w + x] - mean) * rstd;
    a += g;
    b += static_cast<double>(g) * zn;
  }
  red[0][threadIdx.x] = a;
  red[1][threadIdx.x] = b;
  __syncthreads();
  __shared__ float sh[2];
  if (threadIdx.x == 0) {
    double ta = 0.0, tb = 0.0;
    for (int i = 0; i < 256; ++i) { ta += red[0][i]; tb += red[1][i]; }
    sh[0] = static_cast<float>(ta / per);
    sh[1] = static_cast<float>(tb / per);
  }
  __syncthreads();
  const float ma = sh[0], mb = sh[1];
  for (long long i = threadIdx.x; i < per; i += blockDim.x) {
    const int j2 = static_cast<int>(i % w2);
    const int i2 = static_cast<int>((i / w2) % h2);
    const long long img = i / (static_cast<long long>(w2) * h2);
    const int y = 2 * i2 + pi, x = 2 * j2 + pj;
    const long long zi = ((img * zc + c) * h + y) * w + x;
    const float g = T16<T>::to_f(dout[((img * h + y) * w + x) * dout_pitch + c]) * scale;
    const float zn = (z[zi] - mean) * rstd;
    dz[zi] = rstd * (g - ma - zn * mb);
  }
}

}  // namespace

extern "C" {

size_t eovae_gn_backward_workspace_bytes(int n, long long hw, int c, int groups) {
  const int threads = block_threads(c);
  if (threads <= 0) return 0;
  int bpi, ppb;
  bwd_grid(n, hw, c, threads / (c / 8), &bpi, &ppb);
  return sizeof(float) * (2 * static_cast<size_t>(n) * bpi * c + 2 * static_cast<size_t>(n) * groups +
                          2 * static_cast<size_t>(n) * c + 64);
}

int eovae_gn_backward(const void* x, const void* grad_out, int dtype, int grad_dtype, const float* stats, const float* gamma,
                      const float* beta, int n, long long hw, int c, int groups, int with_silu, const void* grad_add,
                      void* grad_x, float* dgamma, float* dbeta, int accumulate_params, float* grad_x_colsum,
                      void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  EOVAE_CHECK(c % 8 == 0 && c % groups == 0, "gn_backward: C (%d) must be a multiple of 8 and of groups (%d)", c, groups);
  EOVAE_CHECK((dtype == EOVAE_BF16 || dtype == EOVAE_F16) && (grad_dtype == EOVAE_BF16 || grad_dtype == EOVAE_F16),
              "gn_backward: 16-bit tensors only");
  const int threads = block_threads(c);
  EOVAE_CHECK(threads > 0, "gn_backward: C too large (%d)", c);
  EOVAE_CHECK(workspace_bytes >= eovae_gn_backward_workspace_bytes(n, hw, c, groups), "gn_backward: workspace too small");
  const int rows = threads / (c / 8);
  int bpi, ppb;
  bwd_grid(n, hw, c, rows, &bpi, &ppb);
  float* partial = static_cast<float*>(workspace);
  float* gsum = partial + 2 * static_cast<size_t>(n) * bpi * c;
  float* chsum = gsum + 2 * static_cast<size_t>(n) * groups;
  dim3 grid(bpi, n);
  const size_t smem = sizeof(float) * 2 * c * rows;
#define EOVAE_GNB_R(T, TG, S)                                                                                             \
  gn_bwd_reduce_kernel<T, TG, S><<<grid, threads, smem, stream>>>(static_cast<const T*>(x), static_cast<const TG*>(grad_out), \
                                                                 stats, gamma, beta, hw, c, groups, partial, ppb)
#define EOVAE_GNB_A(T, TG, S)                                                                                            \
  gn_bwd_apply_kernel<T, TG, S><<<grid, threads, grad_x_colsum ? sizeof(float) * c * rows : 0, stream>>>(                \
      static_cast<const T*>(x), static_cast<const TG*>(grad_out), stats, gamma, beta, gsum, static_cast<const TG*>(grad_add), \
      static_cast<TG*>(grad_x), hw, c, groups, ppb, grad_x_colsum ? partial : nullptr)
  // bulk-async staged variants (see BulkRing): channel-vector count must divide the block
  const bool bulk = g_gn_bwd_bulk != 0 && threads == kThreads && kThreads % (c / 8) == 0;
#define EOVAE_GNB_LAUNCH_BULK(KERNEL, BYTES, ...)                                                                    \
  do {                                                                                                                \
    static bool attr_set = false;                                                                                     \
    if (!attr_set) {                                                                                                  \
      EOVAE_CUDA(cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(BYTES))); \
      attr_set = true;                                                                                                \
    }                                                                                                                 \
    KERNEL<<<grid, kThreads, BYTES, stream>>>(__VA_ARGS__);                                                           \
  } while (0)
#define EOVAE_GNB_RB(T, TG, S)                                                                                        \
  EOVAE_GNB_LAUNCH_BULK((gn_bwd_reduce_bulk_kernel<T, TG, S, 4, 3>), (eovae::BulkRing<2, 4, 3>::kSmemBytes),                  \
                        static_cast<const T*>(x), static_cast<const TG*>(grad_out), stats, gamma, beta, hw, c, groups, \
                        partial, ppb)
#define EOVAE_GNB_AB(T, TG, S)                                                                                        \
  do {                                                                                                                \
    if (grad_add != nullptr)                                                                                          \
      EOVAE_GNB_LAUNCH_BULK((gn_bwd_apply_bulk_kernel<T, TG, S, true, 2, 4>), (eovae::BulkRing<3, 2, 4>::kSmemBytes),         \
                            static_cast<const T*>(x), static_cast<const TG*>(grad_out), stats, gamma, beta, gsum,     \
                            static_cast<const TG*>(grad_add), static_cast<TG*>(grad_x), hw, c, groups, ppb,           \
                            grad_x_colsum ? partial : nullptr);                                                       \
    else                                                                                                              \
      EOVAE_GNB_LAUNCH_BULK((gn_bwd_apply_bulk_kernel<T, TG, S, false, 4, 3>), (eovae::BulkRing<2, 4, 3>::kSmemBytes),        \
                            static_cast<const T*>(x), static_cast<const TG*>(grad_out), stats, gamma, beta, gsum,     \
                            static_cast<const TG*>(grad_add), static_cast<TG*>(grad_x), hw, c, groups, ppb,           \
                            grad_x_colsum ? partial : nullptr);                                                       \
  } while (0)
  // (activation type, gradient type): bf16/bf16, f16/f16 and the default training mix f16 activations / bf16 gradients
#define EOVAE_GNB_DISPATCH(M)                                                                                  \
  do {                                                                                                         \
    if (dtype == EOVAE_BF16 && grad_dtype == EOVAE_BF16) {                                                     \
      if (with_silu) M(__nv_bfloat16, __nv_bfloat16, true); else M(__nv_bfloat16, __nv_bfloat16, false);       \
    } else if (dtype == EOVAE_F16 && grad_dtype == EOVAE_F16) {                                                \
      if (with_silu) M(__half, __half, true); else M(__half, __half, false);                                   \
    } else if (dtype == EOVAE_F16 && grad_dtype == EOVAE_BF16) {                                               \
      if (with_silu) M(__half, __nv_bfloat16, true); else M(__half, __nv_bfloat16, false);                     \
    } else {                                                                                                   \
      EOVAE_CHECK(false, "gn_backward: unsupported (activation, gradient) dtype pair (%d, %d)", dtype, grad_dtype); \
    }                                                                                                          \
  } while (0)
  if (bulk) EOVAE_GNB_DISPATCH(EOVAE_GNB_RB); else EOVAE_GNB_DISPATCH(EOVAE_GNB_R);
  EOVAE_LAUNCH_CHECK();
  {
    const int cpg = c / groups;
    const bool window = c % 32 == 0 && cpg <= 32 && 32 % cpg == 0;  // 32-channel windows hold whole groups
    const int cw = window ? 32 : c;
    const int fthreads = window ? 256 : 1024;
    const int slices = cw <= fthreads ? fthreads / cw : 1;
    gn_bwd_finalize_kernel<<<dim3(c / cw, n), fthreads, sizeof(float) * 2 * cw * (1 + slices), stream>>>(partial, gamma, n, bpi, c, groups, cw,
                                                                                                       gsum, chsum);
  }
  EOVAE_LAUNCH_CHECK();
  if (dgamma != nullptr && dbeta != nullptr) {
    gn_bwd_param_kernel<<<ceil_div(c, 128), 128, 0, stream>>>(chsum, n, c, dgamma, dbeta, accumulate_params);
    EOVAE_LAUNCH_CHECK();
  }
  if (grad_x != nullptr) {
    if (bulk) EOVAE_GNB_DISPATCH(EOVAE_GNB_AB); else EOVAE_GNB_DISPATCH(EOVAE_GNB_A);
    EOVAE_LAUNCH_CHECK();
    if (grad_x_colsum != nullptr) {  // the reduce partials are dead by now: their buffer carried the column-sum slots
      colsum_finalize_kernel<<<ceil_div(c, 32), dim3(32, 32), 0, stream>>>(partial, n * bpi, c, grad_x_colsum, 0);
      EOVAE_LAUNCH_CHECK();
    }
  }
#undef EOVAE_GNB_R
#undef EOVAE_GNB_RB
#undef EOVAE_GNB_AB
#undef EOVAE_GNB_LAUNCH_BULK
#undef EOVAE_GNB_A
#undef EOVAE_GNB_DISPATCH
  return 0;
}

int eovae_scatter_stride2(const void* dy, void* z, int n, int ho, int wo, int h, int w, int c, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  EOVAE_CHECK(c % 8 == 0, "scatter_stride2: C must be a multiple of 8");
  const long long total = static_cast<long long>(n) * h * w * (c / 8);
  scatter_s2_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
      static_cast<const uint4*>(dy), static_cast<uint4*>(z), ho, wo, h, w, c / 8, total);
  EOVAE_LAUNCH_CHECK();
  return 0;
}

int eovae_pool2x2_sum(const void* g, void* out, int dtype, int n, int h, int w, int c, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  EOVAE_CHECK(c % 8 == 0, "pool2x2_sum: C must be a multiple of 8");
  const long long total = static_cast<long long>(n) * h * w * (c / 8);
  const unsigned grid = static_cast<unsigned>((total + 255) / 256);
  if (dtype == EOVAE_BF16)
    pool2x2_sum_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(g), static_cast<__nv_bfloat16*>(out), h, w, c / 8, total);
  else
    pool2x2_sum_kernel<__half><<<grid, 256, 0, stream>>>(static_cast<const __half*>(g), static_cast<__half*>(out), h, w, c / 8, total);
  EOVAE_LAUNCH_CHECK();
  return 0;
}

size_t eovae_bias_grad_workspace_bytes(long long pixels, int c) {
  return sizeof(float) * static_cast<size_t>((pixels + 511) / 512 + 1) * c;
}

int eovae_bias_grad(const void* grad_out, int dtype, long long pixels, int c, float* dbias, int accumulate, void* workspace,
                    size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  EOVAE_CHECK(c % 8 == 0, "bias_grad: C must be a multiple of 8");
  EOVAE_CHECK(workspace_bytes >= eovae_bias_grad_workspace_bytes(pixels, c), "bias_grad: workspace too small");
  const int threads = block_threads(c);
  EOVAE_CHECK(threads > 0, "bias_grad: C too large");
  const int rows = threads / (c / 8);
  const long long ppb = 512;
  const int blocks = static_cast<int>((pixels + ppb - 1) / ppb);
  float* partial = static_cast<float*>(workspace);
  const size_t smem = sizeof(float) * rows * c;
  if (dtype == EOVAE_BF16)
    colsum_partial_kernel<__nv_bfloat16><<<blocks, threads, smem, stream>>>(static_cast<const __nv_bfloat16*>(grad_out), pixels, c, partial, ppb);
  else
    colsum_partial_kernel<__half><<<blocks, threads, smem, stream>>>(static_cast<const __half*>(grad_out), pixels, c, partial, ppb);
  EOVAE_LAUNCH_CHECK();
  colsum_finalize_kernel<<<ceil_div(c, 32), dim3(32, 32), 0, stream>>>(partial, blocks, c, dbias, accumulate);
  EOVAE_LAUNCH_CHECK();
  return 0;
}

int eovae_softmax_backward(const void* p, long long p_ld, const float* dp, long long dp_ld, void* ds, long long ds_ld,
                           int dtype, int ds_dtype, long long rows, int cols, int out_cols, float scale, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  EOVAE_CHECK(p_ld >= cols && dp_ld >= cols && ds_ld >= out_cols && out_cols >= cols, "softmax_backward: bad pitches");
  const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
#define EOVAE_SMB(T, TG)                                                                                             \
  softmax_bwd_kernel<T, TG><<<grid, 256, 0, stream>>>(static_cast<const T*>(p), p_ld, dp, dp_ld, static_cast<TG*>(ds), ds_ld, \
                                                      rows, cols, out_cols, scale)
  if (dtype == EOVAE_BF16 && ds_dtype == EOVAE_BF16) EOVAE_SMB(__nv_bfloat16, __nv_bfloat16);
  else if (dtype == EOVAE_F16 && ds_dtype == EOVAE_F16) EOVAE_SMB(__half, __half);
  else if (dtype == EOVAE_F16 && ds_dtype == EOVAE_BF16) EOVAE_SMB(__half, __nv_bfloat16);
  else EOVAE_CHECK(false, "softmax_backward: unsupported (probability, gradient) dtype pair (%d, %d)", dtype, ds_dtype);
#undef EOVAE_SMB
  EOVAE_LAUNCH_CHECK();
  return 0;
}

int eovae_reparam_backward(const float* moments, const long long* mstrides, const float* eps, const float* dz, float* dmoments,
                           int n, int h, int w, int zc, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const Strides4 ms{mstrides[0], mstrides[1], mstrides[2], mstrides[3]};
  const long long total = static_cast<long long>(n) * zc * h * w;
  reparam_bwd_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(moments, ms, eps, dz, dmoments, h, w, zc, total);
  EOVAE_LAUNCH_CHECK();
  return 0;
}

int eovae_pixel_loss_backward(const float* a, const float* b, long long count, float eps, int kind, const float* grad_scale,
                              float* grad_a, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  EOVAE_CHECK(kind == 0 || kind == 1, "pixel_loss_backward: kind must be 0 (L1) or 1 (Charbonnier)");
  long long blocks = (count + 1023) / 1024;
  const long long cap = 8LL * eovae_num_sms();
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  pixel_loss_bwd_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(a, b, count, eps * eps, kind, grad_scale, grad_a);
  EOVAE_LAUNCH_CHECK();
  return 0;
}

int eovae_latent_bn_train_forward(const float* z, int n, int zc, int h, int w, float* running_mean, float* running_var,
                                  float momentum, float eps_bn, float eps_inv, void* out, int out_dtype, long long out_pix_stride,
                                  float* save, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  EOVAE_CHECK(h % 2 == 0 && w % 2 == 0, "latent_bn_train: latent H, W must be even (2x2 packing)");
  if (out_dtype == EOVAE_BF16)
    latent_bn_train_fwd_kernel<__nv_bfloat16><<<4 * zc, 256, 0, stream>>>(z, n, zc, h, w, running_mean, running_var, momentum, eps_bn,
                                                                          eps_inv, static_cast<__nv_bfloat16*>(out), out_pix_stride, save);
  else if (out_dtype == EOVAE_F16)
    latent_bn_train_fwd_kernel<__half><<<4 * zc, 256, 0, stream>>>(z, n, zc, h, w, running_mean, running_var, momentum, eps_bn, eps_inv,
                                                                   static_cast<__half*>(out), out_pix_stride, save);
  else
    EOVAE_CHECK(false, "latent_bn_train: 16-bit output only");
  EOVAE_LAUNCH_CHECK();
  return 0;
}

int eovae_latent_bn_train_backward(const void* dout, int dtype, long long dout_pix_stride, const float* z, int n, int zc, int h, int w,
                                   const float* save, float* dz, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (dtype == EOVAE_BF16)
    latent_bn_train_bwd_kernel<__nv_bfloat16><<<4 * zc, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(dout), dout_pix_stride, z, n,
                                                                          zc, h, w, save, dz);
  else if (dtype == EOVAE_F16)
    latent_bn_train_bwd_kernel<__half><<<4 * zc, 256, 0, stream>>>(static_cast<const __half*>(dout), dout_pix_stride, z, n, zc, h, w,
                                                                   save, dz);
  else
    EOVAE_CHECK(false, "latent_bn_train_backward: 16-bit gradient only");
  EOVAE_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
