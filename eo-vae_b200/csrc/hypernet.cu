// Wavelength hypernetwork (TransformerWeightGenerator + FCResLayer + sincos embedding), fp32 SIMT kernels.
// ~1.5 GFLOP per call, batch independent and latency bound (sequence of 128 + C + 1 tokens, d_model 256):
// not a tensor-core problem, and it must stay fp32 (sin/cos of arguments up to ~1e4 rad).
#include "../../include/eovae.h"
#include "common.cuh"

namespace {

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2 };

__global__ void sincos_kernel(const float* __restrict__ wvs_um, const float* __restrict__ omega, float* __restrict__ emb,
                              int c, int d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = d / 2;
  if (i >= c * half) return;
  const int m = i / half, k = i % half;
  const float pos = __fmul_rn(wvs_um[m], 1000.0f);
  const float ang = __fmul_rn(pos, omega[k]);
  emb[m * d + k] = sinf(ang);
  emb[m * d + half + k] = cosf(ang);
}

// Y[s][n] = act(sum_k X[s][k] * W[n][k] + b[n]) + R[s][n]
// Skinny GEMM (s <= 142 rows): 64 x 64 output tiles, 4 x 4 outputs per thread, K split across blockIdx.z so that
// even the N = 256 layers fill the 148 SMs; split partials go to a workspace and are reduced in FIXED order by
// linear_reduce_kernel (deterministic - no atomics), which also applies bias / activation / residual.
constexpr int LBM = 64, LBN = 64, LBK = 16;
__global__ void __launch_bounds__(256) linear_partial_kernel(const float* __restrict__ x, int ldx,
                                                             const float* __restrict__ w, float* __restrict__ part,
                                                             int s, int n, int k, int k_per_split) {
  __shared__ float xs[LBK][LBM + 4];
  __shared__ float ws[LBK][LBN + 4];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int row0 = blockIdx.y * LBM, col0 = blockIdx.x * LBN;
  const int kb = blockIdx.z * k_per_split;
  const int ke = min(k, kb + k_per_split);
  float acc[4][4] = {};
  for (int k0 = kb; k0 < ke; k0 += LBK) {
    {  // 64 rows x 16 k: one float4 per thread for each operand
      const int rr = threadIdx.x / 4, kk = (threadIdx.x % 4) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row0 + rr < s && k0 + kk < ke) v = *reinterpret_cast<const float4*>(&x[static_cast<long long>(row0 + rr) * ldx + k0 + kk]);
      xs[kk][rr] = v.x; xs[kk + 1][rr] = v.y; xs[kk + 2][rr] = v.z; xs[kk + 3][rr] = v.w;
      float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
      if (col0 + rr < n && k0 + kk < ke) u = __ldg(reinterpret_cast<const float4*>(&w[static_cast<long long>(col0 + rr) * k + k0 + kk]));
      ws[kk][rr] = u.x; ws[kk + 1][rr] = u.y; ws[kk + 2][rr] = u.z; ws[kk + 3][rr] = u.w;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < LBK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&xs[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&ws[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* pz = part + static_cast<long long>(blockIdx.z) * s * n;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int rr = row0 + ty * 4 + i;
    if (rr >= s) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int cc = col0 + tx * 4 + j;
      if (cc < n) pz[static_cast<long long>(rr) * n + cc] = acc[i][j];
    }
  }
}

__global__ void linear_reduce_kernel(const float* __restrict__ part, int splits, const float* __restrict__ b,
                                     const float* __restrict__ r, int ldr, float* __restrict__ y, int ldy, int s, int n,
                                     int act) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= s * n) return;
  const int rr = i / n, cc = i % n;
  float v = 0.f;
  for (int z = 0; z < splits; ++z) v += part[static_cast<long long>(z) * s * n + i];
  if (b != nullptr) v += b[cc];
  if (act == ACT_RELU) v = fmaxf(v, 0.f);
  if (act == ACT_GELU) v = 0.5f * v * (1.f + erff(v * 0.70710678118654752f));
  if (r != nullptr) v += r[static_cast<long long>(rr) * ldr + cc];
  y[static_cast<long long>(rr) * ldy + cc] = v;
}

// y[row] = LayerNorm(x[row]) * g + b ; one warp per row
__global__ void layernorm_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ b,
                                 float* __restrict__ y, int rows, int d, float eps) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + static_cast<long long>(row) * d;
  float s = 0.f;
  for (int i = lane; i < d; i += 32) s += xr[i];
  const float mean = warp_sum(s) / d;
  float q = 0.f;
  for (int i = lane; i < d; i += 32) {
    const float t = xr[i] - mean;
    q += t * t;
  }
  const float rstd = rsqrtf(warp_sum(q) / d + eps);
  for (int i = lane; i < d; i += 32) y[static_cast<long long>(row) * d + i] = (xr[i] - mean) * rstd * g[i] + b[i];
}

// qkv [s][3d] -> out [s][d]; one warp per (head, query)
constexpr int MHA_MAX_S = 192;
__global__ void __launch_bounds__(128) mha_kernel(const float* __restrict__ qkv, float* __restrict__ out, int s, int d,
                                                  int heads) {
  __shared__ float probs[4][MHA_MAX_S];
  __shared__ float qs[4][128];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * 4 + wid;
  const int hd = d / heads;
  if (item >= s * heads) return;
  const int h = item / s, qi = item % s;
  const float scale = rsqrtf(static_cast<float>(hd));
  const float* qp = qkv + static_cast<long long>(qi) * 3 * d + h * hd;
  for (int i = lane; i < hd; i += 32) qs[wid][i] = qp[i] * scale;
  __syncwarp();
  float m = -INFINITY;
  for (int j = lane; j < s; j += 32) {
    const float* kp = qkv + static_cast<long long>(j) * 3 * d + d + h * hd;
    float acc = 0.f;
    for (int i = 0; i < hd; ++i) acc = fmaf(qs[wid][i], kp[i], acc);
    probs[wid][j] = acc;
    m = fmaxf(m, acc);
  }
  m = warp_max(m);
  float sum = 0.f;
  for (int j = lane; j < s; j += 32) {
    const float e = expf(probs[wid][j] - m);
    probs[wid][j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  __syncwarp();
  const float inv = 1.f / sum;
  for (int i = lane; i < hd; i += 32) {
    float acc = 0.f;
    for (int j = 0; j < s; ++j) acc = fmaf(probs[wid][j], qkv[static_cast<long long>(j) * 3 * d + 2 * d + h * hd + i], acc);
    out[static_cast<long long>(qi) * d + h * hd + i] = acc * inv;
  }
}

// y[r][:] = a[r][:] + b[(bcast ? 0 : r)][:]
__global__ void add_rows_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y, int rows,
                                int d, int bcast) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * d) return;
  y[i] = a[i] + b[bcast ? (i % d) : i];
}

constexpr size_t kLinearPartFloats = static_cast<size_t>(32) * 192 * 2048;  // upper bound used by the workspace query

int linear(const float* x, int ldx, const float* w, const float* b, const float* r, int ldr, float* y, int ldy, int s,
           int n, int k, int act, float* part, cudaStream_t st) {
  EOVAE_CHECK(k % 4 == 0 && ldx % 4 == 0, "hypernet linear: K and ldx must be multiples of 4");
  const int tiles = ceil_div(n, LBN) * ceil_div(s, LBM);
  int splits = ceil_div(2 * eovae_num_sms(), tiles);          // aim at ~2 blocks per SM
  const int max_splits = ceil_div(k, 4 * LBK);                // at least 64 k per split
  if (splits > max_splits) splits = max_splits;
  if (splits > 32) splits = 32;
  if (splits < 1) splits = 1;
  int kps = round_up(ceil_div(k, splits), LBK);
  splits = ceil_div(k, kps);
  EOVAE_CHECK(static_cast<size_t>(splits) * s * n <= kLinearPartFloats, "hypernet linear: partial buffer too small");
  dim3 grid(ceil_div(n, LBN), ceil_div(s, LBM), splits);
  linear_partial_kernel<<<grid, 256, 0, st>>>(x, ldx, w, part, s, n, k, kps);
  EOVAE_LAUNCH_CHECK();
  linear_reduce_kernel<<<ceil_div(s * n, 256), 256, 0, st>>>(part, splits, b, r, ldr, y, ldy, s, n, act);
  EOVAE_LAUNCH_CHECK();
  return 0;
}

}  // namespace

extern "C" {

size_t eovae_hypernet_workspace_bytes(int c, int d, int ff, int embed) {
  const size_t s = 128 + c + 1;
  (void)embed;
  const size_t floats = 5 * static_cast<size_t>(c) * d + 3 * s * d + s * 3 * d + s * ff + 64 + kLinearPartFloats;
  return floats * sizeof(float);
}

// params order: 0 omega[d/2] | 1 weight_tokens | 2 bias_token | 3,4 fclayer.w1 (w,b) | 5,6 fclayer.w2 | 7,8 fc_weight |
//               9,10 fc_bias | then per layer: in_proj (w,b), out_proj (w,b), linear1 (w,b), linear2 (w,b), norm1 (g,b), norm2 (g,b)
int eovae_hypernet_forward(const float* wvs_um, int c, const float* const* params, int num_layers, int d, int heads,
                           int ff, int embed, int decoder, float* wk_out, float* bias_out, void* workspace,
                           size_t workspace_bytes, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  EOVAE_CHECK(c >= 1 && c <= 62, "hypernet: band count %d out of range [1, 62]", c);
  EOVAE_CHECK(d % heads == 0 && d / heads <= 128 && d % 2 == 0, "hypernet: bad d_model/heads (%d/%d)", d, heads);
  EOVAE_CHECK(workspace_bytes >= eovae_hypernet_workspace_bytes(c, d, ff, embed), "hypernet: workspace too small");
  const int s = 128 + c + 1;
  EOVAE_CHECK(s <= MHA_MAX_S, "hypernet: sequence too long");
  float* ws = static_cast<float*>(workspace);
  float* emb = ws; ws += c * d;
  float* t1 = ws; ws += c * d;
  float* waves = ws; ws += c * d;
  float* headin = ws; ws += c * d;
  float* headin2 = ws; ws += c * d;
  float* x = ws; ws += s * d;
  float* att = ws; ws += s * d;
  float* tmp = ws; ws += s * d;
  float* qkv = ws; ws += s * 3 * d;
  float* ffh = ws; ws += static_cast<size_t>(s) * ff;
  float* part = ws;
  const float* omega = params[0];
  const float* wtok = params[1];
  const float* btok = params[2];

  sincos_kernel<<<ceil_div(c * d / 2, 128), 128, 0, st>>>(wvs_um, omega, emb, c, d);
  EOVAE_LAUNCH_CHECK();
  // FCResLayer: waves = emb + relu(W2 relu(W1 emb + b1) + b2)
  if (linear(emb, d, params[3], params[4], nullptr, 0, t1, d, c, d, d, ACT_RELU, part, st)) return -1;
  if (linear(t1, d, params[5], params[6], emb, d, waves, d, c, d, d, ACT_RELU, part, st)) return -1;
  // token sequence [weight_tokens; waves; bias_token]
  EOVAE_CUDA(cudaMemcpyAsync(x, wtok, sizeof(float) * 128 * d, cudaMemcpyDeviceToDevice, st));
  EOVAE_CUDA(cudaMemcpyAsync(x + 128 * d, waves, sizeof(float) * c * d, cudaMemcpyDeviceToDevice, st));
  EOVAE_CUDA(cudaMemcpyAsync(x + (128 + c) * d, btok, sizeof(float) * d, cudaMemcpyDeviceToDevice, st));
  for (int l = 0; l < num_layers; ++l) {
    const float* const* lp = params + 11 + 12 * l;
    if (linear(x, d, lp[0], lp[1], nullptr, 0, qkv, 3 * d, s, 3 * d, d, ACT_NONE, part, st)) return -1;
    mha_kernel<<<ceil_div(s * heads, 4), 128, 0, st>>>(qkv, att, s, d, heads);
    EOVAE_LAUNCH_CHECK();
    if (linear(att, d, lp[2], lp[3], x, d, tmp, d, s, d, d, ACT_NONE, part, st)) return -1;
    layernorm_kernel<<<ceil_div(s, 4), 128, 0, st>>>(tmp, lp[8], lp[9], x, s, d, 1e-5f);
    EOVAE_LAUNCH_CHECK();
    if (linear(x, d, lp[4], lp[5], nullptr, 0, ffh, ff, s, ff, d, ACT_GELU, part, st)) return -1;
    if (linear(ffh, ff, lp[6], lp[7], x, d, tmp, d, s, d, ff, ACT_NONE, part, st)) return -1;
    layernorm_kernel<<<ceil_div(s, 4), 128, 0, st>>>(tmp, lp[10], lp[11], x, s, d, 1e-5f);
    EOVAE_LAUNCH_CHECK();
  }
  // heads
  add_rows_kernel<<<ceil_div(c * d, 256), 256, 0, st>>>(x + 128 * d, waves, headin, c, d, 0);
  EOVAE_LAUNCH_CHECK();
  if (linear(headin, d, params[7], params[8], nullptr, 0, wk_out, 9 * embed, c, 9 * embed, d, ACT_NONE, part, st)) return -1;
  if (decoder) {
    add_rows_kernel<<<ceil_div(c * d, 256), 256, 0, st>>>(x + 128 * d, btok, headin2, c, d, 1);
    EOVAE_LAUNCH_CHECK();
    if (linear(headin2, d, params[9], params[10], nullptr, 0, bias_out, 1, c, 1, d, ACT_NONE, part, st)) return -1;
  } else {
    if (linear(x + (128 + c) * d, d, params[9], params[10], nullptr, 0, bias_out, embed, 1, embed, d, ACT_NONE, part, st)) return -1;
  }
  return 0;
}

}  // extern "C"
