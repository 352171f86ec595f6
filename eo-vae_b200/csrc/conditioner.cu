// AdaIN conditioning of the ResnetBlocks (use_adain=True): WavelengthConditioner (model.py:35-64) and the per-block
// style projection folded into the second GroupNorm's affine (layers.py:68-76, 96-104).  Everything here is a function of
// the wavelength vector only (one row, <= 1024 wide): fp32 GEMV-sized kernels, latency bound, deterministic.
#include "../../include/eovae.h"
#include "common.cuh"

namespace {

__device__ __forceinline__ float sigmoid_exact(float v) { return 1.f / (1.f + expf(-v)); }

// emb[k] = mean_b sin(wvs[b] * omega[k]) | emb[half + k] = mean_b cos(...)    (model.py:17-32, 56-60; wvs in micrometres,
// NOT scaled by 1000 - unlike the dynamic-conv embedding)
__global__ void style_embed_kernel(const float* __restrict__ wvs, int c, const float* __restrict__ omega, int d,
                                   float* __restrict__ emb) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = d / 2;
  if (k >= half) return;
  float s = 0.f, co = 0.f;
  for (int b = 0; b < c; ++b) {
    const float ang = __fmul_rn(wvs[b], omega[k]);
    s += sinf(ang);
    co += cosf(ang);
  }
  emb[k] = s / c;
  emb[half + k] = co / c;
}

// one warp per output: z[n] = x . w[n] + b[n] ; a[n] = silu(z[n]) (optional)
__global__ void gemv_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                            float* __restrict__ z, float* __restrict__ a, int n, int k) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* wr = w + static_cast<long long>(row) * k;
  float acc = 0.f;
  for (int i = lane; i < k; i += 32) acc = fmaf(wr[i], x[i], acc);
  acc = warp_sum(acc);
  if (lane == 0) {
    const float v = acc + (b != nullptr ? b[row] : 0.f);
    z[row] = v;
    if (a != nullptr) a[row] = v * sigmoid_exact(v);
  }
}

// dx[j] (+)= sum_n dy[n] * w[n][j]   (coalesced over j; fixed summation order)
__global__ void gemv_t_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx, int n, int k,
                              int accumulate) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= k) return;
  float acc = 0.f;
  for (int r = 0; r < n; ++r) acc = fmaf(dy[r], w[static_cast<long long>(r) * k + j], acc);
  dx[j] = (accumulate ? dx[j] : 0.f) + acc;
}

// dw[n][j] = dy[n] * x[j] ; db[n] = dy[n]
__global__ void outer_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dw,
                             float* __restrict__ db, int n, int k) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(n) * k) return;
  const int r = static_cast<int>(i / k), j = static_cast<int>(i % k);
  dw[i] = dy[r] * x[j];
  if (j == 0) db[r] = dy[r];
}

// dz = da * silu'(z)
__global__ void silu_bwd_kernel(const float* __restrict__ z, const float* __restrict__ da, float* __restrict__ dz, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = z[i], sg = sigmoid_exact(v);
  dz[i] = da[i] * sg * (1.f + v * (1.f - sg));
}

// gamma' = gamma * scale ; beta' = beta * scale + shift ; style2 = [scale | shift]
__global__ void adain_fold_kernel(const float* __restrict__ style2, const float* __restrict__ gamma,
                                  const float* __restrict__ beta, float* __restrict__ g_out, float* __restrict__ b_out, int c) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  const float sc = style2[i], sh = style2[c + i];
  g_out[i] = gamma[i] * sc;
  b_out[i] = fmaf(beta[i], sc, sh);
}
__global__ void adain_fold_bwd_kernel(const float* __restrict__ style2, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, const float* __restrict__ dg_out,
                                      const float* __restrict__ db_out, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                      float* __restrict__ dstyle2, int c) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  const float sc = style2[i];
  dgamma[i] = dg_out[i] * sc;
  dbeta[i] = db_out[i] * sc;
  dstyle2[i] = fmaf(dg_out[i], gamma[i], db_out[i] * beta[i]);
  dstyle2[c + i] = db_out[i];
}

inline unsigned blocks(long long n, int per = 256) { return static_cast<unsigned>((n + per - 1) / per); }

int gemv(const float* x, const float* w, const float* b, float* z, float* a, int n, int k, cudaStream_t st) {
  gemv_kernel<<<ceil_div(n, 8), 256, 0, st>>>(x, w, b, z, a, n, k);
  EOVAE_LAUNCH_CHECK();
  return 0;
}

}  // namespace

extern "C" {

// tape layout (floats): emb[d] | z1[2d] a1[2d] | z2[d] a2[d] | scratch: g1[2d] g2[2d] (gradients)
size_t eovae_wavelength_style_workspace_bytes(int d) { return sizeof(float) * static_cast<size_t>(11) * d; }

int eovae_wavelength_style_forward(const float* wvs_um, int c, const float* const* params, int d, float* style,
                                   void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  EOVAE_CHECK(c >= 1 && d >= 2 && d % 2 == 0, "wavelength style: bad band count / width (%d, %d)", c, d);
  EOVAE_CHECK(workspace_bytes >= eovae_wavelength_style_workspace_bytes(d), "wavelength style: workspace too small");
  float* ws = static_cast<float*>(workspace);
  float *emb = ws, *z1 = ws + d, *a1 = ws + 3 * d, *z2 = ws + 5 * d, *a2 = ws + 6 * d;
  style_embed_kernel<<<ceil_div(d / 2, 128), 128, 0, st>>>(wvs_um, c, params[0], d, emb);
  EOVAE_LAUNCH_CHECK();
  if (gemv(emb, params[1], params[2], z1, a1, 2 * d, d, st)) return -1;
  if (gemv(a1, params[3], params[4], z2, a2, d, 2 * d, st)) return -1;
  if (gemv(a2, params[5], params[6], style, nullptr, d, d, st)) return -1;
  return 0;
}

int eovae_wavelength_style_backward(const float* const* params, int d, const float* dstyle, float* const* grads,
                                    void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  EOVAE_CHECK(workspace_bytes >= eovae_wavelength_style_workspace_bytes(d), "wavelength style backward: workspace too small");
  float* ws = static_cast<float*>(workspace);
  float *emb = ws, *z1 = ws + d, *a1 = ws + 3 * d, *z2 = ws + 5 * d, *a2 = ws + 6 * d, *g1 = ws + 7 * d, *g2 = ws + 9 * d;
  // style = a2 W3^T + b3
  outer_kernel<<<blocks(static_cast<long long>(d) * d), 256, 0, st>>>(dstyle, a2, grads[5], grads[6], d, d);
  EOVAE_LAUNCH_CHECK();
  gemv_t_kernel<<<blocks(d), 256, 0, st>>>(dstyle, params[5], g1, d, d, 0);            // da2
  EOVAE_LAUNCH_CHECK();
  silu_bwd_kernel<<<blocks(d), 256, 0, st>>>(z2, g1, g2, d);                            // dz2
  EOVAE_LAUNCH_CHECK();
  outer_kernel<<<blocks(static_cast<long long>(d) * 2 * d), 256, 0, st>>>(g2, a1, grads[3], grads[4], d, 2 * d);
  EOVAE_LAUNCH_CHECK();
  gemv_t_kernel<<<blocks(2 * d), 256, 0, st>>>(g2, params[3], g1, d, 2 * d, 0);        // da1 [2d]
  EOVAE_LAUNCH_CHECK();
  silu_bwd_kernel<<<blocks(2 * d), 256, 0, st>>>(z1, g1, g2, 2 * d);                    // dz1 [2d]
  EOVAE_LAUNCH_CHECK();
  outer_kernel<<<blocks(static_cast<long long>(2 * d) * d), 256, 0, st>>>(g2, emb, grads[1], grads[2], 2 * d, d);
  EOVAE_LAUNCH_CHECK();
  return 0;
}

int eovae_adain_affine_forward(const float* style, int d, const float* wproj, const float* bproj, const float* gamma,
                               const float* beta, int cout, float* gamma_out, float* beta_out, float* style2, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  EOVAE_CHECK(d >= 1 && cout >= 1, "adain affine: bad sizes (%d, %d)", d, cout);
  if (gemv(style, wproj, bproj, style2, nullptr, 2 * cout, d, st)) return -1;
  adain_fold_kernel<<<blocks(cout), 256, 0, st>>>(style2, gamma, beta, gamma_out, beta_out, cout);
  EOVAE_LAUNCH_CHECK();
  return 0;
}

int eovae_adain_affine_backward(const float* style, int d, const float* wproj, const float* gamma, const float* beta,
                                const float* style2, int cout, const float* dgamma_out, const float* dbeta_out,
                                float* dgamma, float* dbeta, float* dwproj, float* dbproj, float* dstyle, float* dstyle2,
                                void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  EOVAE_CHECK(d >= 1 && cout >= 1, "adain affine backward: bad sizes (%d, %d)", d, cout);
  adain_fold_bwd_kernel<<<blocks(cout), 256, 0, st>>>(style2, gamma, beta, dgamma_out, dbeta_out, dgamma, dbeta, dstyle2, cout);
  EOVAE_LAUNCH_CHECK();
  outer_kernel<<<blocks(static_cast<long long>(2 * cout) * d), 256, 0, st>>>(dstyle2, style, dwproj, dbproj, 2 * cout, d);
  EOVAE_LAUNCH_CHECK();
  gemv_t_kernel<<<blocks(d), 256, 0, st>>>(dstyle2, wproj, dstyle, 2 * cout, d, 0);
  EOVAE_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
