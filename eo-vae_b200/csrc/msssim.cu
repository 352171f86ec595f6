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
  // horizontal pass: IN rows x TS columns
  for (int i = threadIdx.x; i < IN * TS; i += 256) {
    const int r = i / TS, c = i % TS;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float pv = sp[r][c + k], tv = st[r][c + k], wk = g.w[k];
      a0 = fmaf(wk, pv, a0);
      a1 = fmaf(wk, tv, a1);
      a2 = fmaf(wk, pv * pv, a2);
      a3 = fmaf(wk, tv * tv, a3);
      a4 = fmaf(wk, pv * tv, a4);
    }
    hz[0][r][c] = a0; hz[1][r][c] = a1; hz[2][r][c] = a2; hz[3][r][c] = a3; hz[4][r][c] = a4;
  }
  __syncthreads();
  // vertical pass + ssim / cs on the cropped interior
  float s_ssim = 0.f, s_cs = 0.f;
  for (int i = threadIdx.x; i < TS * TS; i += 256) {
    const int r = i / TS, c = i % TS;
    const int y = y0 + r, x = x0 + c;
    if (y < HALO || y >= h - HALO || x < HALO || x >= w - HALO) continue;
    float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f, m4 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float wk = g.w[k];
      m0 = fmaf(wk, hz[0][r + k][c], m0);
      m1 = fmaf(wk, hz[1][r + k][c], m1);
      m2 = fmaf(wk, hz[2][r + k][c], m2);
      m3 = fmaf(wk, hz[3][r + k][c], m3);
      m4 = fmaf(wk, hz[4][r + k][c], m4);
    }
    const float mpp = m0 * m0, mtt = m1 * m1, mpt = m0 * m1;
    const float spp = fmaxf(m2 - mpp, 0.f), stt = fmaxf(m3 - mtt, 0.f), spt = m4 - mpt;
    const float upper = 2.f * spt + c2, lower = spp + stt + c2;
    s_cs += upper / lower;
    s_ssim += ((2.f * mpt + c1) * upper) / ((mpp + mtt + c1) * lower);
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

// one block: per sample, per scale: mean ssim / cs (fixed order), relu, weighted product; then the batch mean
__global__ void msssim_finalize_kernel(const float* __restrict__ partial, Scales sc, int b, float* __restrict__ per_sample,
                                       float* __restrict__ out) {
  __shared__ double acc[256];
  double local = 0.0;
  for (int n = threadIdx.x; n < b; n += blockDim.x) {
    double prod = 1.0;
    for (int s = 0; s < NSCALES; ++s) {
      const float* base = partial + sc.s[s].offset + static_cast<long long>(n) * sc.s[s].slots_per_sample * 2;
      double a = 0.0, c = 0.0;
      for (int i = 0; i < sc.s[s].slots_per_sample; ++i) {
        a += base[2 * i];
        c += base[2 * i + 1];
      }
      double v = (s == NSCALES - 1 ? a : c) / sc.s[s].count;
      if (v < 0.0) v = 0.0;  // normalize='relu'
      prod *= pow(v, static_cast<double>(sc.beta[s]));
    }
    if (per_sample != nullptr) per_sample[n] = static_cast<float>(prod);
    local += prod;
  }
  acc[threadIdx.x] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int i = 0; i < static_cast<int>(blockDim.x); ++i) tot += acc[i];
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

}  // extern "C"
