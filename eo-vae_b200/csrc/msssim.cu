// Multi-scale SSIM (forward) for EOConsistencyLoss: the reference calls torchmetrics'
// MultiScaleStructuralSimilarityIndexMeasure(data_range=6, kernel_size=5, 5 betas) at consistency_loss.py:24-37.
// torchmetrics is not vendored; the algorithm restated here (and in oracle/eovae_oracle.py::ms_ssim) is: per scale an
// 11-tap (sigma 1.5) separable Gaussian over reflect-padded p, t, p^2, t^2, p*t; ssim / contrast maps cropped by 5
// pixels; per-sample means, relu; 2x2 average pooling between scales; prod_i cs_i^beta_i * ssim_last^beta_last.
//
// One kernel per scale: a block owns a 32x32 output tile of one (sample, channel) plane, stages the 42x42 haloed
// input once in shared memory, runs the horizontal then the vertical pass out of shared memory, reduces its ssim /
// cs sums with warp shuffles into a FIXED partial slot (deterministic) and also emits the 2x2-pooled tile that
// feeds the next scale - each scale reads its input exactly once (~(10/3) N s bytes over the pyramid).
#include "../../include/eovae.h"
#include "common.cuh"

namespace {

constexpr int TS = 32;           // output tile edge
constexpr int HALO = 5;          // (11 - 1) / 2
constexpr int IN = TS + 2 * HALO;
constexpr int NSCALES = 5;

struct Gauss { float w[11]; };

__device__ __forceinline__ int reflect(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i;
}

__global__ void __launch_bounds__(256) ssim_scale_kernel(const float* __restrict__ p, const float* __restrict__ t, int h,
                                                         int w, Gauss g, float c1, float c2, float* __restrict__ partial,
                                                         float* __restrict__ p_next, float* __restrict__ t_next) {
  __shared__ float sp[IN][IN + 1], st[IN][IN + 1];
  __shared__ float hz[5][IN][TS + 1];  // horizontally filtered p, t, pp, tt, pt
  __shared__ float red[2][8];
  const int plane = blockIdx.z;
  const int x0 = blockIdx.x * TS, y0 = blockIdx.y * TS;
  const float* pp = p + static_cast<long long>(plane) * h * w;
  const float* tp = t + static_cast<long long>(plane) * h * w;
  for (int i = threadIdx.x; i < IN * IN; i += 256) {
    const int r = i / IN, c = i % IN;
    const int yy = reflect(y0 + r - HALO, h), xx = reflect(x0 + c - HALO, w);
    // tiles may overhang the plane (h, w not multiples of 32): clamp the reflected index, those outputs are masked
    const int yc = min(max(yy, 0), h - 1), xc = min(max(xx, 0), w - 1);
    sp[r][c] = __ldg(pp + static_cast<long long>(yc) * w + xc);
    st[r][c] = __ldg(tp + static_cast<long long>(yc) * w + xc);
  }
  __syncthreads();
  // 2x2 average pooling of the un-haloed tile -> next scale input
  if (p_next != nullptr) {
    const int hn = h / 2, wn = w / 2;
    for (int i = threadIdx.x; i < (TS / 2) * (TS / 2); i += 256) {
      const int r = i / (TS / 2), c = i % (TS / 2);
      const int oy = y0 / 2 + r, ox = x0 / 2 + c;
      if (oy < hn && ox < wn) {
        const int sr = HALO + 2 * r, sc = HALO + 2 * c;
        p_next[(static_cast<long long>(plane) * hn + oy) * wn + ox] =
            0.25f * (sp[sr][sc] + sp[sr][sc + 1] + sp[sr + 1][sc] + sp[sr + 1][sc + 1]);
        t_next[(static_cast<long long>(plane) * hn + oy) * wn + ox] =
            0.25f * (st[sr][sc] + st[sr][sc + 1] + st[sr + 1][sc] + st[sr + 1][sc + 1]);
      }
    }
  }
  // horizontal pass, register blocked: one item = 4 consecutive output columns of one haloed row (14 p + 14 t shared-memory
  // loads for 4 x 5 outputs instead of 88); 4 rows x 8 column groups per warp -> conflict-free loads and stores
  for (int i = threadIdx.x; i < IN * (TS / 4); i += 256) {
    const int r = i / (TS / 4), c0 = 4 * (i % (TS / 4));
    float pv[14], tv[14];
#pragma unroll
    for (int k = 0; k < 14; ++k) {
      pv[k] = sp[r][c0 + k];
      tv[k] = st[r][c0 + k];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
#pragma unroll
      for (int k = 0; k < 11; ++k) {
        const float p1 = pv[j + k], t1 = tv[j + k], wk = g.w[k];
        a0 = fmaf(wk, p1, a0);
        a1 = fmaf(wk, t1, a1);
        a2 = fmaf(wk, p1 * p1, a2);
        a3 = fmaf(wk, t1 * t1, a3);
        a4 = fmaf(wk, p1 * t1, a4);
      }
      hz[0][r][c0 + j] = a0; hz[1][r][c0 + j] = a1; hz[2][r][c0 + j] = a2; hz[3][r][c0 + j] = a3; hz[4][r][c0 + j] = a4;
    }
  }
  __syncthreads();
  // vertical pass, register blocked: one thread = one column x 4 consecutive rows (14 loads per map for 4 outputs instead of
  // 44), then ssim / cs on the cropped interior
  float s_ssim = 0.f, s_cs = 0.f;
  {
    const int c = threadIdx.x % TS, r0 = 4 * (threadIdx.x / TS);
    float m[5][4];
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      float col[14];
#pragma unroll
      for (int k = 0; k < 14; ++k) col[k] = hz[q][r0 + k][c];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < 11; ++k) a = fmaf(g.w[k], col[j + k], a);
        m[q][j] = a;
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int y = y0 + r0 + j, x = x0 + c;
      if (y < HALO || y >= h - HALO || x < HALO || x >= w - HALO) continue;
      const float m0 = m[0][j], m1 = m[1][j];
      const float mpp = m0 * m0, mtt = m1 * m1, mpt = m0 * m1;
      const float spp = fmaxf(m[2][j] - mpp, 0.f), stt = fmaxf(m[3][j] - mtt, 0.f), spt = m[4][j] - mpt;
      const float upper = 2.f * spt + c2, lower = spp + stt + c2;
      s_cs += upper / lower;
      s_ssim += ((2.f * mpt + c1) * upper) / ((mpp + mtt + c1) * lower);
    }
  }
  s_ssim = warp_sum(s_ssim);
  s_cs = warp_sum(s_cs);
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = s_ssim;
    red[1][threadIdx.x >> 5] = s_cs;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int i = 0; i < 8; ++i) {
      a += red[0][i];
      b += red[1][i];
    }
    const long long slot = (static_cast<long long>(plane) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    partial[2 * slot] = a;
    partial[2 * slot + 1] = b;
  }
}

struct ScaleInfo { long long offset; int slots_per_sample; double count; };
struct Scales { ScaleInfo s[NSCALES]; float beta[NSCALES]; };

// Per-(sample, scale) mean of the ssim / cs partial sums: ONE WARP per pair (lanes stride the slots, double accumulation,
// shuffle tree - a fixed order), results in shared memory.  (One thread per sample walked 768 + 192 + ... slots as a single
// dependent chain: 80 us per launch for 16 samples.)
constexpr int MS_CHUNK = 64;  // samples per pass of a 256-thread block
__device__ __forceinline__ void msssim_scale_means(const float* __restrict__ partial, const Scales& sc, int b, int n0,
                                                   double (*vs)[NSCALES] /* [MS_CHUNK][NSCALES] shared */) {
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int nloc = min(MS_CHUNK, b - n0);
  for (int item = wid; item < nloc * NSCALES; item += nwarps) {
    const int nl = item / NSCALES, s = item % NSCALES;
    const int slots = sc.s[s].slots_per_sample;
    const float2* base = reinterpret_cast<const float2*>(partial + sc.s[s].offset) + static_cast<long long>(n0 + nl) * slots;
    double a = 0.0, c = 0.0;
#pragma unroll 4
    for (int i = lane; i < slots; i += 32) {
      const float2 v = base[i];
      a += v.x;
      c += v.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if (lane == 0) {
      double v = (s == NSCALES - 1 ? a : c) / sc.s[s].count;
      if (v < 0.0) v = 0.0;  // normalize='relu'
      vs[nl][s] = v;
    }
  }
}

// one block: per sample the weighted product of the scale means; then the batch mean
__global__ void __launch_bounds__(256) msssim_finalize_kernel(const float* __restrict__ partial, Scales sc, int b,
                                                              float* __restrict__ per_sample, float* __restrict__ out) {
  __shared__ double acc[256];
  __shared__ double vs[MS_CHUNK][NSCALES];
  double local = 0.0;
  for (int n0 = 0; n0 < b; n0 += MS_CHUNK) {
    __syncthreads();  // the previous chunk's means have been consumed
    msssim_scale_means(partial, sc, b, n0, vs);
    __syncthreads();
    const int n = n0 + static_cast<int>(threadIdx.x);
    if (threadIdx.x < MS_CHUNK && n < b) {
      double prod = 1.0;
      for (int s = 0; s < NSCALES; ++s) prod *= pow(vs[threadIdx.x][s], static_cast<double>(sc.beta[s]));
      if (per_sample != nullptr) per_sample[n] = static_cast<float>(prod);
      local += prod;
    }
  }
  acc[threadIdx.x] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    // sample order: thread t holds samples t, t + 64, ... - summed thread by thread (fixed order)
    double tot = 0.0;
    for (int i = 0; i < MS_CHUNK; ++i) tot += acc[i];
    out[0] = static_cast<float>(tot / b);
  }
}

size_t pyramid_floats(int b, int c, int h, int w) {
  size_t n = 0;
  for (int s = 1; s < NSCALES; ++s) {
    h /= 2;
    w /= 2;
    n += 2 * static_cast<size_t>(b) * c * h * w;
  }
  return n;
}
size_t partial_floats(int b, int c, int h, int w) {
  size_t n = 0;
  for (int s = 0; s < NSCALES; ++s) {
    n += 2 * static_cast<size_t>(b) * c * ceil_div(h, TS) * ceil_div(w, TS);
    h /= 2;
    w /= 2;
  }
  return n;
}


// ------------------------------------------------------------------------------------------------- backward
// d(mean MS-SSIM)/d(pred).  Per scale the map value at an output pixel is f(mu_p, mu_t, E_pp, E_tt, E_pt) with the five
// Gaussian-filtered fields; only outputs inside the 5-pixel crop count, and those never touch the reflect padding, so
// the adjoint of the filter is the same symmetric 11-tap correlation with zeros outside the crop:
//   dp = G*A + 2 p (G*B) + t (G*C),  A = k df/dmu_p, B = k df/dE_pp, C = k df/dE_pt,  k = per-(sample, scale) scalar.
// A block owns a 32x32 tile of dp: it stages the 52x52 haloed inputs, recomputes the filtered fields on the 42x42 outputs
// that reach the tile, and runs the adjoint passes out of shared memory; the 2x2 average-pool adjoint of the next
// (coarser) scale's gradient is added on the way out.  Scales run coarse -> fine.
constexpr int OUTR = TS + 2 * HALO;   // 42: outputs that reach the tile
constexpr int INR = TS + 4 * HALO;    // 52: inputs those outputs read
constexpr size_t kBwdSmemFloats = 2 * INR * (INR + 1) + 5 * INR * (OUTR + 1) + 3 * OUTR * (OUTR + 1) + 3 * OUTR * (TS + 1);

__global__ void __launch_bounds__(256) ssim_scale_bwd_kernel(const float* __restrict__ p, const float* __restrict__ t, int h,
                                                             int w, int c, Gauss g, float c1, float c2,
                                                             const float* __restrict__ coef, int use_ssim,
                                                             const float* __restrict__ dp_next, float* __restrict__ dp) {
  extern __shared__ float smf[];
  float (*sp)[INR + 1] = reinterpret_cast<float (*)[INR + 1]>(smf);
  float (*st)[INR + 1] = reinterpret_cast<float (*)[INR + 1]>(smf + INR * (INR + 1));
  float (*hz)[INR][OUTR + 1] = reinterpret_cast<float (*)[INR][OUTR + 1]>(smf + 2 * INR * (INR + 1));
  float (*abc)[OUTR][OUTR + 1] = reinterpret_cast<float (*)[OUTR][OUTR + 1]>(smf + 2 * INR * (INR + 1) + 5 * INR * (OUTR + 1));
  float (*h2)[OUTR][TS + 1] =
      reinterpret_cast<float (*)[OUTR][TS + 1]>(smf + 2 * INR * (INR + 1) + 5 * INR * (OUTR + 1) + 3 * OUTR * (OUTR + 1));
  const int plane = blockIdx.z;
  const float k = coef[plane / c];
  const int x0 = blockIdx.x * TS, y0 = blockIdx.y * TS;
  const float* pp = p + static_cast<long long>(plane) * h * w;
  const float* tp = t + static_cast<long long>(plane) * h * w;
  for (int i = threadIdx.x; i < INR * INR; i += 256) {
    const int r = i / INR, cc = i % INR;
    const int yc = min(max(y0 + r - 2 * HALO, 0), h - 1), xc = min(max(x0 + cc - 2 * HALO, 0), w - 1);
    sp[r][cc] = __ldg(pp + static_cast<long long>(yc) * w + xc);
    st[r][cc] = __ldg(tp + static_cast<long long>(yc) * w + xc);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < INR * OUTR; i += 256) {
    const int r = i / OUTR, cc = i % OUTR;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
#pragma unroll
    for (int j = 0; j < 11; ++j) {
      const float pv = sp[r][cc + j], tv = st[r][cc + j], wk = g.w[j];
      a0 = fmaf(wk, pv, a0);
      a1 = fmaf(wk, tv, a1);
      a2 = fmaf(wk, pv * pv, a2);
      a3 = fmaf(wk, tv * tv, a3);
      a4 = fmaf(wk, pv * tv, a4);
    }
    hz[0][r][cc] = a0; hz[1][r][cc] = a1; hz[2][r][cc] = a2; hz[3][r][cc] = a3; hz[4][r][cc] = a4;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < OUTR * OUTR; i += 256) {
    const int r = i / OUTR, cc = i % OUTR;
    const int y = y0 + r - HALO, x = x0 + cc - HALO;
    float va = 0.f, vb = 0.f, vc = 0.f;
    if (k != 0.f && y >= HALO && y < h - HALO && x >= HALO && x < w - HALO) {
      float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f, m4 = 0.f;
#pragma unroll
      for (int j = 0; j < 11; ++j) {
        const float wk = g.w[j];
        m0 = fmaf(wk, hz[0][r + j][cc], m0);
        m1 = fmaf(wk, hz[1][r + j][cc], m1);
        m2 = fmaf(wk, hz[2][r + j][cc], m2);
        m3 = fmaf(wk, hz[3][r + j][cc], m3);
        m4 = fmaf(wk, hz[4][r + j][cc], m4);
      }
      const float mpp = m0 * m0, mtt = m1 * m1, mpt = m0 * m1;
      const float vpp = m2 - mpp;
      const float spp = fmaxf(vpp, 0.f), stt = fmaxf(m3 - mtt, 0.f), spt = m4 - mpt;
      const float gate = vpp > 0.f ? 1.f : 0.f;
      const float upper = 2.f * spt + c2, lower = spp + stt + c2;
      const float inv_l = 1.f / lower;
      const float cs = upper * inv_l;
      const float dcs_dspt = 2.f * inv_l, dcs_dspp = -upper * inv_l * inv_l * gate;
      // chain: spt = E_pt - mu_p mu_t ; spp = E_pp - mu_p^2
      float d_ept = dcs_dspt, d_epp = dcs_dspp, d_mu = dcs_dspt * (-m1) + dcs_dspp * (-2.f * m0);
      if (use_ssim) {
        const float ln = 2.f * mpt + c1, ld = mpp + mtt + c1;
        const float lum = ln / ld;
        const float dlum = (2.f * m1 * ld - ln * 2.f * m0) / (ld * ld);
        d_mu = lum * d_mu + cs * dlum;
        d_ept *= lum;
        d_epp *= lum;
      }
      va = k * d_mu;
      vb = k * d_epp;
      vc = k * d_ept;
    }
    abc[0][r][cc] = va; abc[1][r][cc] = vb; abc[2][r][cc] = vc;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < OUTR * TS; i += 256) {
    const int r = i / TS, cc = i % TS;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int j = 0; j < 11; ++j) {
      const float wk = g.w[j];
      a0 = fmaf(wk, abc[0][r][cc + j], a0);
      a1 = fmaf(wk, abc[1][r][cc + j], a1);
      a2 = fmaf(wk, abc[2][r][cc + j], a2);
    }
    h2[0][r][cc] = a0; h2[1][r][cc] = a1; h2[2][r][cc] = a2;
  }
  __syncthreads();
  const int hn = h / 2, wn = w / 2;
  for (int i = threadIdx.x; i < TS * TS; i += 256) {
    const int r = i / TS, cc = i % TS;
    const int y = y0 + r, x = x0 + cc;
    if (y >= h || x >= w) continue;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int j = 0; j < 11; ++j) {
      const float wk = g.w[j];
      a0 = fmaf(wk, h2[0][r + j][cc], a0);
      a1 = fmaf(wk, h2[1][r + j][cc], a1);
      a2 = fmaf(wk, h2[2][r + j][cc], a2);
    }
    float v = a0 + 2.f * sp[r + 2 * HALO][cc + 2 * HALO] * a1 + st[r + 2 * HALO][cc + 2 * HALO] * a2;
    if (dp_next != nullptr && (y >> 1) < hn && (x >> 1) < wn)
      v += 0.25f * dp_next[(static_cast<long long>(plane) * hn + (y >> 1)) * wn + (x >> 1)];
    dp[(static_cast<long long>(plane) * h + y) * w + x] = v;
  }
}

// coef[s][n] = gscale / B * beta_s * prod_n / v_{n,s} / count_s   (0 where the relu clipped v)
__global__ void __launch_bounds__(256) msssim_coef_kernel(const float* __restrict__ partial, Scales sc, int b,
                                                          const float* __restrict__ gscale, float* __restrict__ coef) {
  __shared__ double vs[MS_CHUNK][NSCALES];
  const int n0 = blockIdx.x * MS_CHUNK;
  msssim_scale_means(partial, sc, b, n0, vs);
  __syncthreads();
  const int n = n0 + static_cast<int>(threadIdx.x);
  if (threadIdx.x >= MS_CHUNK || n >= b) return;
  double prod = 1.0;
  for (int s = 0; s < NSCALES; ++s) prod *= pow(vs[threadIdx.x][s], static_cast<double>(sc.beta[s]));
  for (int s = 0; s < NSCALES; ++s) {
    const double v = vs[threadIdx.x][s];
    coef[s * b + n] = v > 0.0 ? static_cast<float>(gscale[0] / b * sc.beta[s] * prod / v / sc.s[s].count) : 0.f;
  }
}

void make_gauss(Gauss* g) {
  double sum = 0.0, tmp[11];
  for (int i = 0; i < 11; ++i) {
    const double d = (i - 5) / 1.5;
    tmp[i] = exp(-d * d / 2.0);
    sum += tmp[i];
  }
  for (int i = 0; i < 11; ++i) g->w[i] = static_cast<float>(tmp[i] / sum);
}

}  // namespace

extern "C" {

size_t eovae_msssim_workspace_bytes(int b, int c, int h, int w) {
  return sizeof(float) * (pyramid_floats(b, c, h, w) + partial_floats(b, c, h, w) + 64);
}

int eovae_msssim(const float* pred, const float* target, int b, int c, int h, int w, float data_range, float* out,
                 float* per_sample, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  EOVAE_CHECK(b > 0 && c > 0, "msssim: empty batch");
  EOVAE_CHECK((h >> 4) > 2 * HALO && (w >> 4) > 2 * HALO,
              "msssim: image %dx%d too small: the coarsest of the 5 scales must exceed the 10-pixel crop", h, w);
  EOVAE_CHECK(h % 16 == 0 && w % 16 == 0, "msssim: H and W must be multiples of 16 (four 2x2 poolings)");
  EOVAE_CHECK(workspace_bytes >= eovae_msssim_workspace_bytes(b, c, h, w), "msssim: workspace too small");
  Gauss g;
  {
    double sum = 0.0, tmp[11];
    for (int i = 0; i < 11; ++i) {
      const double d = (i - 5) / 1.5;
      tmp[i] = exp(-d * d / 2.0);
      sum += tmp[i];
    }
    for (int i = 0; i < 11; ++i) g.w[i] = static_cast<float>(tmp[i] / sum);
  }
  const float c1 = (0.01f * data_range) * (0.01f * data_range), c2 = (0.03f * data_range) * (0.03f * data_range);
  float* ws = static_cast<float*>(workspace);
  float* pyr = ws;
  float* partial = ws + pyramid_floats(b, c, h, w);
  Scales sc;
  const float betas[NSCALES] = {0.0448f, 0.2856f, 0.3001f, 0.2363f, 0.1333f};
  const float* ps = pred;
  const float* ts = target;
  long long poff = 0, pyoff = 0;
  int hs = h, wsz = w;
  for (int s = 0; s < NSCALES; ++s) {
    const int tx = ceil_div(wsz, TS), ty = ceil_div(hs, TS);
    float* pn = nullptr;
    float* tn = nullptr;
    if (s + 1 < NSCALES) {
      const long long plane_next = static_cast<long long>(b) * c * (hs / 2) * (wsz / 2);
      pn = pyr + pyoff;
      tn = pyr + pyoff + plane_next;
      pyoff += 2 * plane_next;
    }
    dim3 grid(tx, ty, b * c);
    ssim_scale_kernel<<<grid, 256, 0, stream>>>(ps, ts, hs, wsz, g, c1, c2, partial + poff, pn, tn);
    EOVAE_LAUNCH_CHECK();
    sc.s[s].offset = poff;
    sc.s[s].slots_per_sample = c * tx * ty;
    sc.s[s].count = static_cast<double>(c) * (hs - 2 * HALO) * (wsz - 2 * HALO);
    sc.beta[s] = betas[s];
    poff += 2LL * b * c * tx * ty;
    ps = pn;
    ts = tn;
    hs /= 2;
    wsz /= 2;
  }
  msssim_finalize_kernel<<<1, 256, 0, stream>>>(partial, sc, b, per_sample, out);
  EOVAE_LAUNCH_CHECK();
  return 0;
}

size_t eovae_msssim_backward_workspace_bytes(int b, int c, int h, int w) {
  // forward workspace (pyramid + partial sums) + gradient pyramid (pred half only) + coefficients
  return eovae_msssim_workspace_bytes(b, c, h, w) + sizeof(float) * (pyramid_floats(b, c, h, w) / 2 + NSCALES * static_cast<size_t>(b) + 64);
}

/* grad_pred = *grad_scale * d(mean MS-SSIM)/d(pred); re-runs the forward pyramid inside. */
int eovae_msssim_backward(const float* pred, const float* target, int b, int c, int h, int w, float data_range,
                          const float* grad_scale, float* grad_pred, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  EOVAE_CHECK(workspace_bytes >= eovae_msssim_backward_workspace_bytes(b, c, h, w), "msssim_backward: workspace too small");
  float* ws = static_cast<float*>(workspace);
  const size_t fwd_floats = eovae_msssim_workspace_bytes(b, c, h, w) / sizeof(float);
  float* out_scratch = ws + fwd_floats - 32;  // inside the forward workspace's 64-float slack
  if (int rc = eovae_msssim(pred, target, b, c, h, w, data_range, out_scratch, nullptr, workspace,
                            eovae_msssim_workspace_bytes(b, c, h, w), stream_))
    return rc;
  float* pyr = ws;
  float* partial = ws + pyramid_floats(b, c, h, w);
  float* gpyr = ws + fwd_floats;
  float* coef = gpyr + pyramid_floats(b, c, h, w) / 2;
  Gauss g;
  make_gauss(&g);
  const float c1 = (0.01f * data_range) * (0.01f * data_range), c2 = (0.03f * data_range) * (0.03f * data_range);
  const float betas[NSCALES] = {0.0448f, 0.2856f, 0.3001f, 0.2363f, 0.1333f};
  Scales sc;
  const float* ps[NSCALES];
  const float* ts[NSCALES];
  float* gs[NSCALES];
  int hs[NSCALES], wsz[NSCALES];
  {
    long long poff = 0, pyoff = 0, goff = 0;
    int hh = h, ww = w;
    ps[0] = pred; ts[0] = target; gs[0] = grad_pred;
    for (int s = 0; s < NSCALES; ++s) {
      hs[s] = hh; wsz[s] = ww;
      const int tx = ceil_div(ww, TS), ty = ceil_div(hh, TS);
      sc.s[s].offset = poff;
      sc.s[s].slots_per_sample = c * tx * ty;
      sc.s[s].count = static_cast<double>(c) * (hh - 2 * HALO) * (ww - 2 * HALO);
      sc.beta[s] = betas[s];
      poff += 2LL * b * c * tx * ty;
      if (s + 1 < NSCALES) {
        const long long plane_next = static_cast<long long>(b) * c * (hh / 2) * (ww / 2);
        ps[s + 1] = pyr + pyoff;
        ts[s + 1] = pyr + pyoff + plane_next;
        pyoff += 2 * plane_next;
        gs[s + 1] = gpyr + goff;
        goff += plane_next;
      }
      hh /= 2;
      ww /= 2;
    }
  }
  msssim_coef_kernel<<<ceil_div(b, MS_CHUNK), 256, 0, stream>>>(partial, sc, b, grad_scale, coef);
  EOVAE_LAUNCH_CHECK();
  static bool attr = false;
  const size_t smem = kBwdSmemFloats * sizeof(float);
  if (!attr) {
    EOVAE_CUDA(cudaFuncSetAttribute(ssim_scale_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr = true;
  }
  for (int s = NSCALES - 1; s >= 0; --s) {
    dim3 grid(ceil_div(wsz[s], TS), ceil_div(hs[s], TS), b * c);
    ssim_scale_bwd_kernel<<<grid, 256, smem, stream>>>(ps[s], ts[s], hs[s], wsz[s], c, g, c1, c2, coef + s * b,
                                                        s == NSCALES - 1 ? 1 : 0, s + 1 < NSCALES ? gs[s + 1] : nullptr, gs[s]);
    EOVAE_LAUNCH_CHECK();
  }
  return 0;
}

}  // extern "C"
