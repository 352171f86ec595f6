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
  // 64 rows x 16 k per operand and step: one float4 per thread each, fetched one step ahead into registers so the
  // global-load latency hides behind the previous step's FMAs (these GEMMs run 10-30 blocks: latency, not throughput)
  const int rr = threadIdx.x / 4, kk4 = (threadIdx.x % 4) * 4;
  auto fetch = [&](int k0, float4& v, float4& u) {
    v = make_float4(0.f, 0.f, 0.f, 0.f);
    u = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + rr < s && k0 + kk4 < ke) v = *reinterpret_cast<const float4*>(&x[static_cast<long long>(row0 + rr) * ldx + k0 + kk4]);
    if (col0 + rr < n && k0 + kk4 < ke) u = __ldg(reinterpret_cast<const float4*>(&w[static_cast<long long>(col0 + rr) * k + k0 + kk4]));
  };
  float4 v, u;
  if (kb < ke) fetch(kb, v, u);
  for (int k0 = kb; k0 < ke; k0 += LBK) {
    xs[kk4][rr] = v.x; xs[kk4 + 1][rr] = v.y; xs[kk4 + 2][rr] = v.z; xs[kk4 + 3][rr] = v.w;
    ws[kk4][rr] = u.x; ws[kk4 + 1][rr] = u.y; ws[kk4 + 2][rr] = u.z; ws[kk4 + 3][rr] = u.w;
    __syncthreads();
    if (k0 + LBK < ke) fetch(k0 + LBK, v, u);
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
// One block = MHA_QPB queries (one warp each) of ONE head.  K_h and V_h (s x hd floats each) are staged in shared memory once
// per block: with one warp per (head, query) reading K / V rows straight from global memory, every SM asked for the same
// lines at the same time and the kernel sat in load latency (46 us per launch at 4 warps per SM for 0.1 GFLOP).  The sums
// run in the original element order (same bits).
constexpr int MHA_QPB = 8;
__global__ void __launch_bounds__(32 * MHA_QPB) mha_kernel(const float* __restrict__ qkv, float* __restrict__ out, int s, int d,
                                                           int heads) {
  extern __shared__ float mha_sm[];
  const int hd = d / heads;
  const int kld = hd + 1;                       // padded K rows: lane j reads row j without bank conflicts
  float* ks = mha_sm;                           // [s][hd + 1]
  float* vs = ks + s * kld;                     // [s][hd]
  float* probs_all = vs + s * hd;               // [MHA_QPB][MHA_MAX_S]
  float* qs_all = probs_all + MHA_QPB * MHA_MAX_S;  // [MHA_QPB][128]
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, qi = blockIdx.x * MHA_QPB + wid;
  // stage K_h, V_h: warp w takes rows w, w + 8, ...; four rows (up to 16 loads per lane) in flight
  for (int j0 = wid; j0 < s; j0 += 4 * MHA_QPB) {
    for (int i = lane; i < hd; i += 32) {
      float kv[4], vv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + u * MHA_QPB;
        const float* row = qkv + static_cast<long long>(j < s ? j : 0) * 3 * d + h * hd + i;
        kv[u] = __ldg(row + d);
        vv[u] = __ldg(row + 2 * d);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + u * MHA_QPB;
        if (j < s) {
          ks[j * kld + i] = kv[u];
          vs[j * hd + i] = vv[u];
        }
      }
    }
  }
  float* probs = probs_all + wid * MHA_MAX_S;
  float* qs = qs_all + wid * 128;
  const float scale = rsqrtf(static_cast<float>(hd));
  if (qi < s) {
    const float* qp = qkv + static_cast<long long>(qi) * 3 * d + h * hd;
    for (int i = lane; i < hd; i += 32) qs[i] = qp[i] * scale;
  }
  __syncthreads();
  if (qi >= s) return;
  float m = -INFINITY;
  for (int j = lane; j < s; j += 32) {
    const float* kp = ks + j * kld;
    float acc = 0.f;
#pragma unroll 8
    for (int i = 0; i < hd; ++i) acc = fmaf(qs[i], kp[i], acc);
    probs[j] = acc;
    m = fmaxf(m, acc);
  }
  m = warp_max(m);
  float sum = 0.f;
  for (int j = lane; j < s; j += 32) {
    const float e = expf(probs[j] - m);
    probs[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  __syncwarp();
  const float inv = 1.f / sum;
  for (int i = lane; i < hd; i += 32) {
    float acc = 0.f;
#pragma unroll 8
    for (int j = 0; j < s; ++j) acc = fmaf(probs[j], vs[j * hd + i], acc);
    out[static_cast<long long>(qi) * d + h * hd + i] = acc * inv;
  }
}

int launch_mha(const float* qkv, float* out, int s, int d, int heads, cudaStream_t st) {
  const int hd = d / heads;
  const size_t smem = sizeof(float) * (static_cast<size_t>(s) * (2 * hd + 1) + MHA_QPB * (MHA_MAX_S + 128));
  static size_t attr_bytes = 0;
  if (smem > attr_bytes) {
    EOVAE_CUDA(cudaFuncSetAttribute(mha_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr_bytes = smem;
  }
  mha_kernel<<<dim3(ceil_div(s, MHA_QPB), heads), 32 * MHA_QPB, smem, st>>>(qkv, out, s, d, heads);
  EOVAE_LAUNCH_CHECK();
  return 0;
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


// ------------------------------------------------------------------------------------------------- backward pieces
// C[m][n] (+)= sum_k A[i*a_rs + k*a_cs] * B[k*b_rs + j*b_cs]: generic-stride fp32 GEMM for the (tiny, <= 142-row)
// hypernetwork gradients; 64 x 64 tiles, 4 x 4 outputs per thread.  blockIdx.z = batch index (per-head products, batch
// strides a_bs / b_bs / c_bs) x K split: with `splits` > 1 every z-slice owns a K range and writes its tile to
// part[split][batch][m][n]; sgemm_reduce_kernel sums the slices in fixed order (deterministic).
__global__ void __launch_bounds__(256) sgemm_strided_kernel(const float* __restrict__ a, long long a_rs, long long a_cs,
                                                            long long a_bs, const float* __restrict__ b, long long b_rs,
                                                            long long b_cs, long long b_bs, float* __restrict__ c,
                                                            long long ldc, long long c_bs, int m, int n, int k,
                                                            int accumulate, int splits, int k_per_split,
                                                            float* __restrict__ part) {
  __shared__ float as[LBK][LBM + 4];
  __shared__ float bs[LBK][LBN + 4];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int row0 = blockIdx.y * LBM, col0 = blockIdx.x * LBN;
  const int batch = blockIdx.z / splits, split = blockIdx.z % splits;
  a += batch * a_bs;
  b += batch * b_bs;
  const int kb = split * k_per_split;
  const int ke = min(k, kb + k_per_split);
  float acc[4][4] = {};
  // every thread owns 4 elements of each 64 x 16 operand tile, fetched one k-step ahead into registers
  float ra[4], rb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e = threadIdx.x + q * 256;
      // consecutive threads walk the unit-stride direction of A (rows when A is read transposed, a_rs == 1)
      const int kk = a_rs == 1 ? e / LBM : e % LBK, rr = a_rs == 1 ? e % LBM : e / LBK;
      ra[q] = (row0 + rr < m && k0 + kk < ke) ? a[(row0 + rr) * a_rs + (k0 + kk) * a_cs] : 0.f;
      const int cc = e % LBN, k2 = e / LBN;
      rb[q] = (col0 + cc < n && k0 + k2 < ke) ? b[(k0 + k2) * b_rs + (col0 + cc) * b_cs] : 0.f;
    }
  };
  if (kb < ke) fetch(kb);
  for (int k0 = kb; k0 < ke; k0 += LBK) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e = threadIdx.x + q * 256;
      const int kk = a_rs == 1 ? e / LBM : e % LBK, rr = a_rs == 1 ? e % LBM : e / LBK;
      as[kk][rr] = ra[q];
      bs[e / LBN][e % LBN] = rb[q];
    }
    __syncthreads();
    if (k0 + LBK < ke) fetch(k0 + LBK);
#pragma unroll
    for (int kk = 0; kk < LBK; ++kk) {
      const float4 av4 = *reinterpret_cast<const float4*>(&as[kk][ty * 4]);
      const float4 bv4 = *reinterpret_cast<const float4*>(&bs[kk][tx * 4]);
      const float av[4] = {av4.x, av4.y, av4.z, av4.w}, bv[4] = {bv4.x, bv4.y, bv4.z, bv4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* out = splits > 1 ? part + (static_cast<long long>(split) * gridDim.z / splits + batch) * m * n : c + batch * c_bs;
  const long long ld = splits > 1 ? n : ldc;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int rr = row0 + ty * 4 + i;
    if (rr >= m) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int cc = col0 + tx * 4 + j;
      if (cc < n) {
        float* o = out + rr * ld + cc;
        *o = (splits == 1 && accumulate ? *o : 0.f) + acc[i][j];
      }
    }
  }
}
__global__ void sgemm_reduce_kernel(const float* __restrict__ part, int splits, int batches, float* __restrict__ c, long long ldc,
                                    long long c_bs, int m, int n, int accumulate) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;  // over (batch, m, n)
  const long long per = static_cast<long long>(m) * n;
  if (i >= per * batches) return;
  const int batch = static_cast<int>(i / per);
  const long long r = i % per;
  float v = 0.f;
#pragma unroll 8
  for (int z = 0; z < splits; ++z) v += part[(static_cast<long long>(z) * batches + batch) * per + r];
  float* o = c + batch * c_bs + (r / n) * ldc + (r % n);
  *o = (accumulate ? *o : 0.f) + v;
}

// `part`: split-K scratch (kLinearPartFloats floats) or nullptr to forbid splitting
int sgemm_b(const float* a, long long a_rs, long long a_cs, long long a_bs, const float* b, long long b_rs, long long b_cs,
            long long b_bs, float* c, long long ldc, long long c_bs, int batches, int m, int n, int k, int accumulate, float* part,
            cudaStream_t st) {
  const int tiles = ceil_div(n, LBN) * ceil_div(m, LBM) * batches;
  int splits = 1;
  if (part != nullptr && tiles < eovae_num_sms() && k >= 8 * LBK) {
    splits = ceil_div(2 * eovae_num_sms(), tiles);
    const int max_splits = k / (4 * LBK);  // at least 64 k per split
    if (splits > max_splits) splits = max_splits;
    while (splits > 1 && static_cast<size_t>(splits) * batches * m * n > static_cast<size_t>(32) * 192 * 2048) --splits;
    if (splits < 1) splits = 1;
  }
  int kps = round_up(ceil_div(k, splits), LBK);
  splits = ceil_div(k, kps);
  dim3 grid(ceil_div(n, LBN), ceil_div(m, LBM), batches * splits);
  sgemm_strided_kernel<<<grid, 256, 0, st>>>(a, a_rs, a_cs, a_bs, b, b_rs, b_cs, b_bs, c, ldc, c_bs, m, n, k, accumulate, splits, kps, part);
  EOVAE_LAUNCH_CHECK();
  if (splits > 1) {
    const long long total = static_cast<long long>(batches) * m * n;
    sgemm_reduce_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(part, splits, batches, c, ldc, c_bs, m, n, accumulate);
    EOVAE_LAUNCH_CHECK();
  }
  return 0;
}

thread_local float* g_sgemm_part = nullptr;  // split-K scratch of the eovae_hypernet_backward call running on this host thread
int sgemm(const float* a, long long a_rs, long long a_cs, const float* b, long long b_rs, long long b_cs, float* c,
          long long ldc, int m, int n, int k, int accumulate, cudaStream_t st) {
  return sgemm_b(a, a_rs, a_cs, 0, b, b_rs, b_cs, 0, c, ldc, 0, 1, m, n, k, accumulate, g_sgemm_part, st);
}

// out[j] (+)= sum_r x[r][j].  Block = 32 columns x 8 row lanes (one thread per column walked the rows as one dependent
// chain: 10-40 us per launch at the hypernetwork's 140 x 128..2048 shapes); fixed combination order (deterministic).
constexpr int COLSUM_LANES = 8;
__global__ void __launch_bounds__(32 * COLSUM_LANES) colsum_f32_kernel(const float* __restrict__ x, long long ld, int rows,
                                                                       int cols, float* __restrict__ out, int accumulate) {
  __shared__ float part[COLSUM_LANES][32];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + cx;
  float a = 0.f;
  if (j < cols) {
#pragma unroll 4
    for (int r = ry; r < rows; r += COLSUM_LANES) a += x[r * ld + j];
  }
  part[ry][cx] = a;
  __syncthreads();
  if (ry == 0 && j < cols) {
    float t = 0.f;
#pragma unroll
    for (int l = 0; l < COLSUM_LANES; ++l) t += part[l][cx];
    out[j] = (accumulate ? out[j] : 0.f) + t;
  }
}

// y[i] = act(z[i]) (forward with a saved pre-activation) / dz[i] = dy[i] * act'(z[i])
__global__ void gelu_fwd_kernel(const float* __restrict__ z, float* __restrict__ y, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) y[i] = 0.5f * z[i] * (1.f + erff(z[i] * 0.70710678118654752f));
}
__global__ void gelu_bwd_kernel(const float* __restrict__ z, const float* __restrict__ dy, float* __restrict__ dz, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = z[i];
  const float cdf = 0.5f * (1.f + erff(v * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * expf(-0.5f * v * v);
  dz[i] = dy[i] * (cdf + v * pdf);
}
// dz = dy where pos > 0 (ReLU mask taken from the activation's OUTPUT, optionally output - base), else 0
__global__ void relu_bwd_kernel(const float* __restrict__ out, const float* __restrict__ base, const float* __restrict__ dy,
                                float* __restrict__ dz, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = out[i] - (base != nullptr ? base[i] : 0.f);
  dz[i] = v > 0.f ? dy[i] : 0.f;
}
// y[i] (+)= alpha * x[i]
__global__ void axpy_kernel(const float* __restrict__ x, float alpha, float* __restrict__ y, long long n, int accumulate) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) y[i] = (accumulate ? y[i] : 0.f) + alpha * x[i];
}

// LayerNorm backward, one warp per row: dx; row statistics go to rowstat[row] = (mean, rstd) for the parameter pass
__global__ void layernorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ dy,
                                     float* __restrict__ dx, float* __restrict__ rowstat, int rows, int d, float eps) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + static_cast<long long>(row) * d;
  const float* dr = dy + static_cast<long long>(row) * d;
  float s = 0.f;
  for (int i = lane; i < d; i += 32) s += xr[i];
  const float mean = warp_sum(s) / d;
  float q = 0.f;
  for (int i = lane; i < d; i += 32) {
    const float t = xr[i] - mean;
    q += t * t;
  }
  const float rstd = rsqrtf(warp_sum(q) / d + eps);
  float a = 0.f, b = 0.f;
  for (int i = lane; i < d; i += 32) {
    const float dg = dr[i] * g[i];
    a += dg;
    b += dg * (xr[i] - mean) * rstd;
  }
  a = warp_sum(a) / d;
  b = warp_sum(b) / d;
  for (int i = lane; i < d; i += 32) {
    const float xh = (xr[i] - mean) * rstd;
    dx[static_cast<long long>(row) * d + i] = rstd * (dr[i] * g[i] - a - xh * b);
  }
  if (lane == 0) {
    rowstat[2 * row] = mean;
    rowstat[2 * row + 1] = rstd;
  }
}
__global__ void __launch_bounds__(32 * COLSUM_LANES) layernorm_bwd_param_kernel(const float* __restrict__ x,
                                                                                const float* __restrict__ dy,
                                                                                const float* __restrict__ rowstat,
                                                                                float* __restrict__ dg, float* __restrict__ db,
                                                                                int rows, int d) {
  __shared__ float pa[COLSUM_LANES][32], pb[COLSUM_LANES][32];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + cx;
  float a = 0.f, b = 0.f;
  if (j < d) {
#pragma unroll 4
    for (int r = ry; r < rows; r += COLSUM_LANES) {
      const float g = dy[static_cast<long long>(r) * d + j];
      a += g * (x[static_cast<long long>(r) * d + j] - rowstat[2 * r]) * rowstat[2 * r + 1];
      b += g;
    }
  }
  pa[ry][cx] = a;
  pb[ry][cx] = b;
  __syncthreads();
  if (ry == 0 && j < d) {
    float ta = 0.f, tb = 0.f;
#pragma unroll
    for (int l = 0; l < COLSUM_LANES; ++l) {
      ta += pa[l][cx];
      tb += pb[l][cx];
    }
    dg[j] = ta;
    db[j] = tb;
  }
}

// attention backward, pass 1: one warp per (head, query): recompute the probability row, dP = dO V^T,
// dS = P o (dP - sum(P o dP)) * scale; writes P and dS rows ([heads][s][s]) and dQ.  (dK = dS^T Q and dV = P^T dO are two
// head-batched GEMMs on these buffers.)
__global__ void __launch_bounds__(128) mha_bwd_q_kernel(const float* __restrict__ qkv, const float* __restrict__ dout,
                                                        float* __restrict__ pbuf, float* __restrict__ dsbuf,
                                                        float* __restrict__ dqkv, int s, int d, int heads) {
  __shared__ float probs[4][MHA_MAX_S];
  __shared__ float dsr[4][MHA_MAX_S];
  __shared__ __align__(16) float qs[4][128];
  __shared__ __align__(16) float dos[4][128];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * 4 + wid;
  const int hd = d / heads;
  if (item >= s * heads) return;
  const int h = item / s, qi = item % s;
  const float scale = rsqrtf(static_cast<float>(hd));
  const float* qp = qkv + static_cast<long long>(qi) * 3 * d + h * hd;
  const float* dop = dout + static_cast<long long>(qi) * d + h * hd;
  for (int i = lane; i < hd; i += 32) {
    qs[wid][i] = qp[i] * scale;
    dos[wid][i] = dop[i];
  }
  __syncwarp();
  float m = -INFINITY;
  for (int j = lane; j < s; j += 32) {
    const float* kp = qkv + static_cast<long long>(j) * 3 * d + d + h * hd;
    const float* vp = kp + d;
    float acc = 0.f, dp = 0.f;
    if ((hd & 3) == 0 && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0) {  // see mha_kernel
      const float4 *kp4 = reinterpret_cast<const float4*>(kp), *vp4 = reinterpret_cast<const float4*>(vp);
      const float4 *q4 = reinterpret_cast<const float4*>(qs[wid]), *do4 = reinterpret_cast<const float4*>(dos[wid]);
#pragma unroll 4
      for (int i = 0; i < (hd >> 2); ++i) {  // element order kept: the same bits as the scalar loop
        const float4 kv = __ldg(kp4 + i), vv = __ldg(vp4 + i), qv = q4[i], dv = do4[i];
        acc = fmaf(qv.x, kv.x, acc); acc = fmaf(qv.y, kv.y, acc); acc = fmaf(qv.z, kv.z, acc); acc = fmaf(qv.w, kv.w, acc);
        dp = fmaf(dv.x, vv.x, dp); dp = fmaf(dv.y, vv.y, dp); dp = fmaf(dv.z, vv.z, dp); dp = fmaf(dv.w, vv.w, dp);
      }
    } else {
      for (int i = 0; i < hd; ++i) {
        acc = fmaf(qs[wid][i], kp[i], acc);
        dp = fmaf(dos[wid][i], vp[i], dp);
      }
    }
    probs[wid][j] = acc;
    dsr[wid][j] = dp;
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
  const float inv = 1.f / sum;
  float dot = 0.f;
  for (int j = lane; j < s; j += 32) {
    const float p = probs[wid][j] * inv;
    probs[wid][j] = p;
    dot += p * dsr[wid][j];
  }
  dot = warp_sum(dot);
  float* prow = pbuf + (static_cast<long long>(h) * s + qi) * s;
  float* dsrow = dsbuf + (static_cast<long long>(h) * s + qi) * s;
  for (int j = lane; j < s; j += 32) {
    const float p = probs[wid][j];
    const float ds = p * (dsr[wid][j] - dot) * scale;
    dsr[wid][j] = ds;
    prow[j] = p;
    dsrow[j] = ds;
  }
  __syncwarp();
  if ((hd & 3) == 0 && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(dqkv) & 15) == 0) {
    for (int i4 = lane; i4 < (hd >> 2); i4 += 32) {  // dQ_i = sum_j dS_ij K_j, four columns per lane (see mha_kernel)
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4* kp4 = reinterpret_cast<const float4*>(qkv + d + h * hd) + i4;
      const long long ld4 = (3LL * d) >> 2;
#pragma unroll 16
      for (int j = 0; j < s; ++j) {
        const float4 kv = __ldg(kp4 + j * ld4);
        const float dj = dsr[wid][j];
        acc.x = fmaf(dj, kv.x, acc.x);
        acc.y = fmaf(dj, kv.y, acc.y);
        acc.z = fmaf(dj, kv.z, acc.z);
        acc.w = fmaf(dj, kv.w, acc.w);
      }
      reinterpret_cast<float4*>(dqkv + static_cast<long long>(qi) * 3 * d + h * hd)[i4] = acc;
    }
  } else {
    for (int i = lane; i < hd; i += 32) {  // dQ_i = sum_j dS_ij K_j
      float acc = 0.f;
      const float* kcol = qkv + d + h * hd + i;
#pragma unroll 8
      for (int j = 0; j < s; ++j) acc = fmaf(dsr[wid][j], __ldg(kcol + static_cast<long long>(j) * 3 * d), acc);
      dqkv[static_cast<long long>(qi) * 3 * d + h * hd + i] = acc;
    }
  }
}
// d_wk[band][tap*E + e] = scale * dW[o][ci][tap], (o, ci) = decoder ? (band, e) : (e, band); dW rows have cin_ld inputs
__global__ void dyn_weight_grad_kernel(const float* __restrict__ dw, int cin_ld, int c, int embed, int decoder, float scale,
                                       float* __restrict__ dwk) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c * 9 * embed) return;
  const int e = i % embed;
  const int tap = (i / embed) % 9;
  const int band = i / (9 * embed);
  const int o = decoder ? band : e, ci = decoder ? e : band;
  dwk[i] = scale * dw[(static_cast<long long>(o) * cin_ld + ci) * 9 + tap];
}

inline unsigned blocks_for(long long n) { return static_cast<unsigned>((n + 255) / 256); }

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
    if (launch_mha(qkv, att, s, d, heads, st)) return -1;
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

// ---- backward of eovae_hypernet_forward (+ eovae_pack_dyn_weight): the forward is re-run keeping every activation in
// the workspace, then walked backwards.  grads[i] receives the gradient of params[i] (written, not accumulated; index 0,
// the sincos table, is not a parameter and may be NULL).
size_t eovae_hypernet_backward_workspace_bytes(int c, int d, int ff, int embed, int num_layers) {
  const size_t s = 128 + c + 1;
  const size_t per_layer = s * d * 5 + s * 3 * d + 2 * s * ff;
  const size_t heads_max = 8;
  const size_t floats = 5 * static_cast<size_t>(c) * d + s * d + num_layers * per_layer  // tape
                        + 4 * s * d + s * 3 * d + 2 * s * ff                               // gradient scratch
                        + 2 * heads_max * s * s + 2 * s + static_cast<size_t>(c) * 9 * embed + 64 + kLinearPartFloats;
  return floats * sizeof(float);
}

}  // extern "C"

namespace {

struct TapeLayer { float *qkv, *att, *tmp1, *x1, *z, *ffh, *tmp2, *xout; };
struct Tape {
  float *emb, *t1, *waves, *headin, *headin2, *x0;
  TapeLayer L[16];
  float *dx, *da, *db_, *dc, *dqkv, *dffh, *dz, *pbuf, *dsbuf, *rowstat, *dwk, *part;
};

// one layout for the taped forward and the backward (eovae_hypernet_backward_workspace_bytes covers it)
void carve_tape(Tape& t, void* workspace, int c, int d, int ff, int embed, int num_layers) {
  const int s = 128 + c + 1;
  const long long sd = static_cast<long long>(s) * d, sf = static_cast<long long>(s) * ff, cd = static_cast<long long>(c) * d;
  float* ws = static_cast<float*>(workspace);
  auto take = [&](long long n) { float* p = ws; ws += n; return p; };
  t.emb = take(cd); t.t1 = take(cd); t.waves = take(cd); t.headin = take(cd); t.headin2 = take(cd);
  t.x0 = take(sd);
  for (int l = 0; l < num_layers; ++l) {
    t.L[l].qkv = take(3 * sd); t.L[l].att = take(sd); t.L[l].tmp1 = take(sd); t.L[l].x1 = take(sd); t.L[l].z = take(sf);
    t.L[l].ffh = take(sf); t.L[l].tmp2 = take(sd); t.L[l].xout = take(sd);
  }
  t.dx = take(sd); t.da = take(sd); t.db_ = take(sd); t.dc = take(sd);
  t.dqkv = take(3 * sd); t.dffh = take(sf); t.dz = take(sf);
  t.pbuf = take(8LL * s * s); t.dsbuf = take(8LL * s * s);
  t.rowstat = take(2 * s);
  t.dwk = take(static_cast<long long>(c) * 9 * embed);
  take(64);
  t.part = ws;
}

// forward keeping every activation (the backward's inputs) in the tape
int tape_forward(Tape& t, const float* wvs_um, int c, const float* const* params, int num_layers, int d, int heads, int ff,
                 int decoder, cudaStream_t st) {
  const int s = 128 + c + 1;
  const long long sf = static_cast<long long>(s) * ff, cd = static_cast<long long>(c) * d;
  float *emb = t.emb, *t1 = t.t1, *waves = t.waves, *headin = t.headin, *headin2 = t.headin2, *x0 = t.x0, *part = t.part;
  TapeLayer* L = t.L;
  const float* omega = params[0];
  const float* wtok = params[1];
  const float* btok = params[2];
  sincos_kernel<<<ceil_div(c * d / 2, 128), 128, 0, st>>>(wvs_um, omega, emb, c, d);
  EOVAE_LAUNCH_CHECK();
  if (linear(emb, d, params[3], params[4], nullptr, 0, t1, d, c, d, d, ACT_RELU, part, st)) return -1;
  if (linear(t1, d, params[5], params[6], emb, d, waves, d, c, d, d, ACT_RELU, part, st)) return -1;
  EOVAE_CUDA(cudaMemcpyAsync(x0, wtok, sizeof(float) * 128 * d, cudaMemcpyDeviceToDevice, st));
  EOVAE_CUDA(cudaMemcpyAsync(x0 + 128 * d, waves, sizeof(float) * cd, cudaMemcpyDeviceToDevice, st));
  EOVAE_CUDA(cudaMemcpyAsync(x0 + (128 + c) * d, btok, sizeof(float) * d, cudaMemcpyDeviceToDevice, st));
  const float* xin = x0;
  for (int l = 0; l < num_layers; ++l) {
    const float* const* lp = params + 11 + 12 * l;
    if (linear(xin, d, lp[0], lp[1], nullptr, 0, L[l].qkv, 3 * d, s, 3 * d, d, ACT_NONE, part, st)) return -1;
    if (launch_mha(L[l].qkv, L[l].att, s, d, heads, st)) return -1;
    if (linear(L[l].att, d, lp[2], lp[3], xin, d, L[l].tmp1, d, s, d, d, ACT_NONE, part, st)) return -1;
    layernorm_kernel<<<ceil_div(s, 4), 128, 0, st>>>(L[l].tmp1, lp[8], lp[9], L[l].x1, s, d, 1e-5f);
    EOVAE_LAUNCH_CHECK();
    if (linear(L[l].x1, d, lp[4], lp[5], nullptr, 0, L[l].z, ff, s, ff, d, ACT_NONE, part, st)) return -1;
    gelu_fwd_kernel<<<blocks_for(sf), 256, 0, st>>>(L[l].z, L[l].ffh, sf);
    EOVAE_LAUNCH_CHECK();
    if (linear(L[l].ffh, ff, lp[6], lp[7], L[l].x1, d, L[l].tmp2, d, s, d, ff, ACT_NONE, part, st)) return -1;
    layernorm_kernel<<<ceil_div(s, 4), 128, 0, st>>>(L[l].tmp2, lp[10], lp[11], L[l].xout, s, d, 1e-5f);
    EOVAE_LAUNCH_CHECK();
    xin = L[l].xout;
  }
  add_rows_kernel<<<ceil_div(c * d, 256), 256, 0, st>>>(xin + 128 * d, waves, headin, c, d, 0);
  EOVAE_LAUNCH_CHECK();
  if (decoder) {
    add_rows_kernel<<<ceil_div(c * d, 256), 256, 0, st>>>(xin + 128 * d, btok, headin2, c, d, 1);
    EOVAE_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace

extern "C" {

/* eovae_hypernet_forward that also leaves the tape of activations in `workspace` (layout and size of
 * eovae_hypernet_backward_workspace_bytes) for a following eovae_hypernet_backward(..., tape_valid = 1) */
int eovae_hypernet_forward_taped(const float* wvs_um, int c, const float* const* params, int num_layers, int d, int heads,
                                 int ff, int embed, int decoder, float* wk_out, float* bias_out, void* workspace,
                                 size_t workspace_bytes, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  EOVAE_CHECK(c >= 1 && c <= 62, "hypernet: band count %d out of range [1, 62]", c);
  EOVAE_CHECK(d % heads == 0 && d / heads <= 128 && d % 2 == 0 && heads <= 8, "hypernet: bad d_model/heads (%d/%d)", d, heads);
  EOVAE_CHECK(num_layers <= 16 && 128 + c + 1 <= MHA_MAX_S, "hypernet: too many layers / tokens");
  EOVAE_CHECK(workspace_bytes >= eovae_hypernet_backward_workspace_bytes(c, d, ff, embed, num_layers), "hypernet taped forward: workspace too small");
  Tape t;
  carve_tape(t, workspace, c, d, ff, embed, num_layers);
  if (tape_forward(t, wvs_um, c, params, num_layers, d, heads, ff, decoder, st)) return -1;
  const float* xl = num_layers > 0 ? t.L[num_layers - 1].xout : t.x0;
  if (linear(t.headin, d, params[7], params[8], nullptr, 0, wk_out, 9 * embed, c, 9 * embed, d, ACT_NONE, t.part, st)) return -1;
  if (decoder) {
    if (linear(t.headin2, d, params[9], params[10], nullptr, 0, bias_out, 1, c, 1, d, ACT_NONE, t.part, st)) return -1;
  } else {
    if (linear(xl + (128 + c) * d, d, params[9], params[10], nullptr, 0, bias_out, embed, 1, embed, d, ACT_NONE, t.part, st)) return -1;
  }
  return 0;
}

int eovae_hypernet_backward(const float* wvs_um, int c, const float* const* params, int num_layers, int d, int heads,
                            int ff, int embed, int decoder, const float* dw_oihw, int dw_cin_ld, float w_scale,
                            const float* dbias, float bias_scale, float* const* grads, int tape_valid, void* workspace,
                            size_t workspace_bytes, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  EOVAE_CHECK(c >= 1 && c <= 62, "hypernet: band count %d out of range [1, 62]", c);
  EOVAE_CHECK(d % heads == 0 && d / heads <= 128 && d % 2 == 0 && heads <= 8, "hypernet: bad d_model/heads (%d/%d)", d, heads);
  EOVAE_CHECK(workspace_bytes >= eovae_hypernet_backward_workspace_bytes(c, d, ff, embed, num_layers),
              "hypernet backward: workspace too small");
  const int s = 128 + c + 1;
  EOVAE_CHECK(s <= MHA_MAX_S && num_layers <= 16, "hypernet: sequence too long / too many layers");
  const long long sd = static_cast<long long>(s) * d, sf = static_cast<long long>(s) * ff, cd = static_cast<long long>(c) * d;
  Tape t;
  carve_tape(t, workspace, c, d, ff, embed, num_layers);
  if (!tape_valid && tape_forward(t, wvs_um, c, params, num_layers, d, heads, ff, decoder, st)) return -1;
  float *waves = t.waves, *emb = t.emb, *t1 = t.t1, *headin = t.headin, *headin2 = t.headin2, *x0 = t.x0;
  float *dx = t.dx, *da = t.da, *db_ = t.db_, *dc = t.dc, *dqkv = t.dqkv, *dffh = t.dffh, *dz = t.dz, *pbuf = t.pbuf,
        *dsbuf = t.dsbuf, *rowstat = t.rowstat, *dwk = t.dwk;
  TapeLayer* L = t.L;
  g_sgemm_part = t.part;
  const float* btok = params[2];
  (void)btok;
  const float* xl = num_layers > 0 ? L[num_layers - 1].xout : x0;

  // ---------------- heads
  const int ne = 9 * embed;
  dyn_weight_grad_kernel<<<blocks_for(static_cast<long long>(c) * ne), 256, 0, st>>>(dw_oihw, dw_cin_ld, c, embed, decoder, w_scale, dwk);
  EOVAE_LAUNCH_CHECK();
  EOVAE_CUDA(cudaMemsetAsync(dx, 0, sizeof(float) * sd, st));
  float* dwaves = da;  // [c][d]
  // wk = headin Wfw^T + bfw
  if (sgemm(dwk, 1, ne, headin, d, 1, grads[7], d, ne, d, c, 0, st)) return -1;            // dWfw [9E][d] = dwk^T headin
  colsum_f32_kernel<<<ceil_div(ne, 32), 32 * COLSUM_LANES, 0, st>>>(dwk, ne, c, ne, grads[8], 0);
  EOVAE_LAUNCH_CHECK();
  if (sgemm(dwk, ne, 1, params[7], d, 1, dwaves, d, c, d, ne, 0, st)) return -1;            // dheadin [c][d] = dwk Wfw
  axpy_kernel<<<blocks_for(cd), 256, 0, st>>>(dwaves, 1.f, dx + 128 * d, cd, 1);
  EOVAE_LAUNCH_CHECK();
  EOVAE_CUDA(cudaMemsetAsync(grads[2], 0, sizeof(float) * d, st));
  if (decoder) {
    // bias[c] = bias_scale * (headin2 wfb^T + bfb), wfb [1][d]
    axpy_kernel<<<blocks_for(c), 256, 0, st>>>(dbias, bias_scale, rowstat, c, 0);           // d(bias_raw) [c]
    EOVAE_LAUNCH_CHECK();
    if (sgemm(rowstat, 0, 1, headin2, d, 1, grads[9], d, 1, d, c, 0, st)) return -1;        // dwfb [1][d]
    colsum_f32_kernel<<<1, 32 * COLSUM_LANES, 0, st>>>(rowstat, 1, c, 1, grads[10], 0);
    EOVAE_LAUNCH_CHECK();
    if (sgemm(rowstat, 1, 0, params[9], 0, 1, db_, d, c, d, 1, 0, st)) return -1;           // dheadin2 [c][d] = db (x) wfb
    axpy_kernel<<<blocks_for(cd), 256, 0, st>>>(db_, 1.f, dx + 128 * d, cd, 1);
    EOVAE_LAUNCH_CHECK();
    colsum_f32_kernel<<<ceil_div(d, 32), 32 * COLSUM_LANES, 0, st>>>(db_, d, c, d, grads[2], 1);         // bias_token (broadcast add)
    EOVAE_LAUNCH_CHECK();
  } else {
    // bias[E] = bias_scale * (x_last Wfb^T + bfb), Wfb [E][d]
    axpy_kernel<<<blocks_for(embed), 256, 0, st>>>(dbias, bias_scale, grads[10], embed, 0);
    EOVAE_LAUNCH_CHECK();
    if (sgemm(grads[10], 1, 0, xl + (128 + c) * d, 0, 1, grads[9], d, embed, d, 1, 0, st)) return -1;   // dWfb = db (x) x_last
    if (sgemm(grads[10], 0, 1, params[9], d, 1, dx + (128 + c) * d, d, 1, d, embed, 1, st)) return -1;  // dx_last += db Wfb
  }

  // ---------------- transformer layers, last to first (dx = gradient of the layer output)
  for (int l = num_layers - 1; l >= 0; --l) {
    const float* const* lp = params + 11 + 12 * l;
    float* const* lg = grads + 11 + 12 * l;
    const float* lin = l == 0 ? x0 : L[l - 1].xout;
    float* dtmp2 = db_;
    layernorm_bwd_kernel<<<ceil_div(s, 4), 128, 0, st>>>(L[l].tmp2, lp[10], dx, dtmp2, rowstat, s, d, 1e-5f);
    EOVAE_LAUNCH_CHECK();
    layernorm_bwd_param_kernel<<<ceil_div(d, 32), 32 * COLSUM_LANES, 0, st>>>(L[l].tmp2, dx, rowstat, lg[10], lg[11], s, d);
    EOVAE_LAUNCH_CHECK();
    // tmp2 = ffh W2^T + b2 + x1
    if (sgemm(dtmp2, d, 1, lp[6], ff, 1, dffh, ff, s, ff, d, 0, st)) return -1;             // dffh = dtmp2 W2
    if (sgemm(dtmp2, 1, d, L[l].ffh, ff, 1, lg[6], ff, d, ff, s, 0, st)) return -1;          // dW2 [d][ff]
    colsum_f32_kernel<<<ceil_div(d, 32), 32 * COLSUM_LANES, 0, st>>>(dtmp2, d, s, d, lg[7], 0);
    EOVAE_LAUNCH_CHECK();
    gelu_bwd_kernel<<<blocks_for(sf), 256, 0, st>>>(L[l].z, dffh, dz, sf);
    EOVAE_LAUNCH_CHECK();
    // z = x1 W1^T + b1 ; dx1 = dtmp2 + dz W1
    float* dx1 = dc;
    EOVAE_CUDA(cudaMemcpyAsync(dx1, dtmp2, sizeof(float) * sd, cudaMemcpyDeviceToDevice, st));
    if (sgemm(dz, ff, 1, lp[4], d, 1, dx1, d, s, d, ff, 1, st)) return -1;
    if (sgemm(dz, 1, ff, L[l].x1, d, 1, lg[4], d, ff, d, s, 0, st)) return -1;               // dW1 [ff][d]
    colsum_f32_kernel<<<ceil_div(ff, 32), 32 * COLSUM_LANES, 0, st>>>(dz, ff, s, ff, lg[5], 0);
    EOVAE_LAUNCH_CHECK();
    float* dtmp1 = db_;
    layernorm_bwd_kernel<<<ceil_div(s, 4), 128, 0, st>>>(L[l].tmp1, lp[8], dx1, dtmp1, rowstat, s, d, 1e-5f);
    EOVAE_LAUNCH_CHECK();
    layernorm_bwd_param_kernel<<<ceil_div(d, 32), 32 * COLSUM_LANES, 0, st>>>(L[l].tmp1, dx1, rowstat, lg[8], lg[9], s, d);
    EOVAE_LAUNCH_CHECK();
    // tmp1 = att Wo^T + bo + xin
    float* datt = dc;
    if (sgemm(dtmp1, d, 1, lp[2], d, 1, datt, d, s, d, d, 0, st)) return -1;
    if (sgemm(dtmp1, 1, d, L[l].att, d, 1, lg[2], d, d, d, s, 0, st)) return -1;             // dWo [d][d]
    colsum_f32_kernel<<<ceil_div(d, 32), 32 * COLSUM_LANES, 0, st>>>(dtmp1, d, s, d, lg[3], 0);
    EOVAE_LAUNCH_CHECK();
    mha_bwd_q_kernel<<<ceil_div(s * heads, 4), 128, 0, st>>>(L[l].qkv, datt, pbuf, dsbuf, dqkv, s, d, heads);
    EOVAE_LAUNCH_CHECK();
    {  // dK_h = dS_h^T Q_h, dV_h = P_h^T dO_h: two head-batched GEMMs over the query index
      const int hd = d / heads;
      if (sgemm_b(dsbuf, 1, s, static_cast<long long>(s) * s, L[l].qkv, 3 * d, 1, hd, dqkv + d, 3 * d, hd, heads, s, hd, s, 0, nullptr, st)) return -1;
      if (sgemm_b(pbuf, 1, s, static_cast<long long>(s) * s, datt, d, 1, hd, dqkv + 2 * d, 3 * d, hd, heads, s, hd, s, 0, nullptr, st)) return -1;
    }
    // qkv = xin Win^T + bin ; dxin = dtmp1 + dqkv Win
    EOVAE_CUDA(cudaMemcpyAsync(dx, dtmp1, sizeof(float) * sd, cudaMemcpyDeviceToDevice, st));
    if (sgemm(dqkv, 3 * d, 1, lp[0], d, 1, dx, d, s, d, 3 * d, 1, st)) return -1;
    if (sgemm(dqkv, 1, 3 * d, lin, d, 1, lg[0], d, 3 * d, d, s, 0, st)) return -1;           // dWin [3d][d]
    colsum_f32_kernel<<<ceil_div(3 * d, 32), 32 * COLSUM_LANES, 0, st>>>(dqkv, 3 * d, s, 3 * d, lg[1], 0);
    EOVAE_LAUNCH_CHECK();
  }

  // ---------------- tokens and FCResLayer
  EOVAE_CUDA(cudaMemcpyAsync(grads[1], dx, sizeof(float) * 128 * d, cudaMemcpyDeviceToDevice, st));
  axpy_kernel<<<blocks_for(d), 256, 0, st>>>(dx + (128 + c) * d, 1.f, grads[2], d, 1);
  EOVAE_LAUNCH_CHECK();
  axpy_kernel<<<blocks_for(cd), 256, 0, st>>>(dx + 128 * d, 1.f, dwaves, cd, 1);
  EOVAE_LAUNCH_CHECK();
  // waves = emb + relu(u2), u2 = t1 W2^T + b2, t1 = relu(emb W1^T + b1)
  float* du2 = db_;
  relu_bwd_kernel<<<blocks_for(cd), 256, 0, st>>>(waves, emb, dwaves, du2, cd);
  EOVAE_LAUNCH_CHECK();
  if (sgemm(du2, 1, d, t1, d, 1, grads[5], d, d, d, c, 0, st)) return -1;
  colsum_f32_kernel<<<ceil_div(d, 32), 32 * COLSUM_LANES, 0, st>>>(du2, d, c, d, grads[6], 0);
  EOVAE_LAUNCH_CHECK();
  float* dt1 = dc;
  if (sgemm(du2, d, 1, params[5], d, 1, dt1, d, c, d, d, 0, st)) return -1;
  relu_bwd_kernel<<<blocks_for(cd), 256, 0, st>>>(t1, nullptr, dt1, dt1, cd);
  EOVAE_LAUNCH_CHECK();
  if (sgemm(dt1, 1, d, emb, d, 1, grads[3], d, d, d, c, 0, st)) return -1;
  colsum_f32_kernel<<<ceil_div(d, 32), 32 * COLSUM_LANES, 0, st>>>(dt1, d, c, d, grads[4], 0);
  EOVAE_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"

// batched strided fp32 GEMM for the other translation units (focal-frequency loss DFTs), no split-K
namespace eovae {
int sgemm_batched(const float* a, long long a_rs, long long a_cs, long long a_bs, const float* b, long long b_rs, long long b_cs,
                  long long b_bs, float* c, long long ldc, long long c_bs, int batches, int m, int n, int k, int accumulate,
                  cudaStream_t st) {
  for (int b0 = 0; b0 < batches; b0 += 32768) {  // gridDim.z limit
    const int nb = batches - b0 < 32768 ? batches - b0 : 32768;
    if (sgemm_b(a + b0 * a_bs, a_rs, a_cs, a_bs, b + b0 * b_bs, b_rs, b_cs, b_bs, c + b0 * c_bs, ldc, c_bs, nb, m, n, k, accumulate,
                nullptr, st))
      return -1;
  }
  return 0;
}
}  // namespace eovae

// =====================================================================================================================
// FactorizedWeightGenerator(_decoder) (dynamic_conv.py:186-302): PRE-norm transformer layers (norm_first=True, ff = 4 d)
// and a low-rank head Linear(d, rank) -> GELU -> Linear(rank, 9E).  Same building blocks as above; the forward always
// keeps its activations (the tape is ~1 MB), so one entry point serves inference and training.  Dropout (p = 0.1 inside
// the reference's train-mode transformer) is not applied: the generated kernel is deterministic in both modes.
// params: 0 omega | 1 weight_tokens | 2 bias_token | 3,4 fclayer.w1 | 5,6 fclayer.w2 | 7,8 fc_weight.0 [rank][d] |
//         9,10 fc_weight.2 [9E][rank] | 11,12 fc_bias | per layer (12): in_proj, out_proj, linear1, linear2, norm1, norm2
namespace {

struct FTapeLayer { float *ln1, *qkv, *att, *x1, *ln2, *z, *ffh, *xout; };
struct FTape {
  float *emb, *t1, *waves, *headin, *headin2, *hr, *hg, *x0;
  FTapeLayer L[16];
  float *dx, *da, *db_, *dc, *dqkv, *dffh, *dz, *pbuf, *dsbuf, *rowstat, *dwk, *dhg, *dhr, *part;
};

size_t carve_ftape(FTape* t, void* workspace, int c, int d, int ff, int embed, int rank, int num_layers) {
  const int s = 128 + c + 1;
  const long long sd = static_cast<long long>(s) * d, sf = static_cast<long long>(s) * ff, cd = static_cast<long long>(c) * d;
  const long long cr = static_cast<long long>(c) * rank;
  float* base = static_cast<float*>(workspace);
  float* ws = base;
  auto take = [&](long long n) { float* p = ws; ws += (n + 3) / 4 * 4; return p; };
  FTape tmp;
  FTape& T = t != nullptr ? *t : tmp;
  T.emb = take(cd); T.t1 = take(cd); T.waves = take(cd); T.headin = take(cd); T.headin2 = take(cd);
  T.hr = take(cr); T.hg = take(cr); T.x0 = take(sd);
  for (int l = 0; l < num_layers; ++l) {
    T.L[l].ln1 = take(sd); T.L[l].qkv = take(3 * sd); T.L[l].att = take(sd); T.L[l].x1 = take(sd); T.L[l].ln2 = take(sd);
    T.L[l].z = take(sf); T.L[l].ffh = take(sf); T.L[l].xout = take(sd);
  }
  T.dx = take(sd); T.da = take(sd); T.db_ = take(sd); T.dc = take(sd);
  T.dqkv = take(3 * sd); T.dffh = take(sf); T.dz = take(sf);
  T.pbuf = take(8LL * s * s); T.dsbuf = take(8LL * s * s);
  T.rowstat = take(2 * s);
  T.dwk = take(static_cast<long long>(c) * 9 * embed);
  T.dhg = take(cr); T.dhr = take(cr);
  take(64);
  T.part = ws;
  return static_cast<size_t>(ws - base) + kLinearPartFloats;
}

int factorized_check(int c, int d, int heads, int ff, int embed, int rank, int num_layers) {
  EOVAE_CHECK(c >= 1 && c <= 62, "hypernet: band count %d out of range [1, 62]", c);
  EOVAE_CHECK(d % heads == 0 && d / heads <= 128 && d % 4 == 0 && heads <= 8, "hypernet: bad d_model/heads (%d/%d)", d, heads);
  EOVAE_CHECK(num_layers >= 0 && num_layers <= 16 && 128 + c + 1 <= MHA_MAX_S, "hypernet: too many layers / tokens");
  EOVAE_CHECK(rank >= 4 && rank % 4 == 0 && ff % 4 == 0 && embed >= 1, "factorized hypernet: rank (%d) and ff (%d) must be multiples of 4", rank, ff);
  return 0;
}

}  // namespace

extern "C" {

size_t eovae_hypernet_factorized_workspace_bytes(int c, int d, int ff, int embed, int rank, int num_layers) {
  if (num_layers < 0 || num_layers > 16) return 0;
  return carve_ftape(nullptr, nullptr, c, d, ff, embed, rank, num_layers) * sizeof(float);
}

int eovae_hypernet_factorized_forward(const float* wvs_um, int c, const float* const* params, int num_layers, int d,
                                      int heads, int ff, int embed, int rank, int decoder, float* wk_out, float* bias_out,
                                      void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  if (factorized_check(c, d, heads, ff, embed, rank, num_layers)) return -1;
  EOVAE_CHECK(workspace_bytes >= eovae_hypernet_factorized_workspace_bytes(c, d, ff, embed, rank, num_layers),
              "factorized hypernet: workspace too small");
  FTape t;
  carve_ftape(&t, workspace, c, d, ff, embed, rank, num_layers);
  const int s = 128 + c + 1;
  const long long sf = static_cast<long long>(s) * ff, cd = static_cast<long long>(c) * d, cr = static_cast<long long>(c) * rank;
  float* part = t.part;
  const float* wtok = params[1];
  const float* btok = params[2];
  sincos_kernel<<<ceil_div(c * d / 2, 128), 128, 0, st>>>(wvs_um, params[0], t.emb, c, d);
  EOVAE_LAUNCH_CHECK();
  if (linear(t.emb, d, params[3], params[4], nullptr, 0, t.t1, d, c, d, d, ACT_RELU, part, st)) return -1;
  if (linear(t.t1, d, params[5], params[6], t.emb, d, t.waves, d, c, d, d, ACT_RELU, part, st)) return -1;
  EOVAE_CUDA(cudaMemcpyAsync(t.x0, wtok, sizeof(float) * 128 * d, cudaMemcpyDeviceToDevice, st));
  EOVAE_CUDA(cudaMemcpyAsync(t.x0 + 128 * d, t.waves, sizeof(float) * cd, cudaMemcpyDeviceToDevice, st));
  EOVAE_CUDA(cudaMemcpyAsync(t.x0 + (128 + c) * d, btok, sizeof(float) * d, cudaMemcpyDeviceToDevice, st));
  const float* xin = t.x0;
  for (int l = 0; l < num_layers; ++l) {
    const float* const* lp = params + 13 + 12 * l;
    FTapeLayer& L = t.L[l];
    // x1 = x + out_proj(MHA(LN1(x)))
    layernorm_kernel<<<ceil_div(s, 4), 128, 0, st>>>(xin, lp[8], lp[9], L.ln1, s, d, 1e-5f);
    EOVAE_LAUNCH_CHECK();
    if (linear(L.ln1, d, lp[0], lp[1], nullptr, 0, L.qkv, 3 * d, s, 3 * d, d, ACT_NONE, part, st)) return -1;
    if (launch_mha(L.qkv, L.att, s, d, heads, st)) return -1;
    if (linear(L.att, d, lp[2], lp[3], xin, d, L.x1, d, s, d, d, ACT_NONE, part, st)) return -1;
    // xout = x1 + W2 gelu(W1 LN2(x1))
    layernorm_kernel<<<ceil_div(s, 4), 128, 0, st>>>(L.x1, lp[10], lp[11], L.ln2, s, d, 1e-5f);
    EOVAE_LAUNCH_CHECK();
    if (linear(L.ln2, d, lp[4], lp[5], nullptr, 0, L.z, ff, s, ff, d, ACT_NONE, part, st)) return -1;
    gelu_fwd_kernel<<<blocks_for(sf), 256, 0, st>>>(L.z, L.ffh, sf);
    EOVAE_LAUNCH_CHECK();
    if (linear(L.ffh, ff, lp[6], lp[7], L.x1, d, L.xout, d, s, d, ff, ACT_NONE, part, st)) return -1;
    xin = L.xout;
  }
  // features = T[128:128+C] + waves ; wk = W2h gelu(W0 features + b0) + b2h
  add_rows_kernel<<<ceil_div(c * d, 256), 256, 0, st>>>(xin + 128 * d, t.waves, t.headin, c, d, 0);
  EOVAE_LAUNCH_CHECK();
  if (linear(t.headin, d, params[7], params[8], nullptr, 0, t.hr, rank, c, rank, d, ACT_NONE, part, st)) return -1;
  gelu_fwd_kernel<<<blocks_for(cr), 256, 0, st>>>(t.hr, t.hg, cr);
  EOVAE_LAUNCH_CHECK();
  if (linear(t.hg, rank, params[9], params[10], nullptr, 0, wk_out, 9 * embed, c, 9 * embed, rank, ACT_NONE, part, st)) return -1;
  if (decoder) {  // bias[c] = fc_bias(features + bias_token)  (dynamic_conv.py:296-299)
    add_rows_kernel<<<ceil_div(c * d, 256), 256, 0, st>>>(t.headin, btok, t.headin2, c, d, 1);
    EOVAE_LAUNCH_CHECK();
    if (linear(t.headin2, d, params[11], params[12], nullptr, 0, bias_out, 1, c, 1, d, ACT_NONE, part, st)) return -1;
  } else {        // bias[E] = fc_bias(T[-1])
    if (linear(xin + (128 + c) * d, d, params[11], params[12], nullptr, 0, bias_out, embed, 1, embed, d, ACT_NONE, part, st)) return -1;
  }
  return 0;
}

/* adjoint of eovae_hypernet_factorized_forward (+ eovae_pack_dyn_weight); `workspace` must still hold that forward's
 * activations (same arguments).  grads[i] <- d/d params[i] (written; grads[0] unused). */
int eovae_hypernet_factorized_backward(const float* wvs_um, int c, const float* const* params, int num_layers, int d,
                                       int heads, int ff, int embed, int rank, int decoder, const float* dw_oihw,
                                       int dw_cin_ld, float w_scale, const float* dbias, float bias_scale,
                                       float* const* grads, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  (void)wvs_um;
  if (factorized_check(c, d, heads, ff, embed, rank, num_layers)) return -1;
  EOVAE_CHECK(workspace_bytes >= eovae_hypernet_factorized_workspace_bytes(c, d, ff, embed, rank, num_layers),
              "factorized hypernet backward: workspace too small");
  FTape t;
  carve_ftape(&t, workspace, c, d, ff, embed, rank, num_layers);
  const int s = 128 + c + 1;
  const long long sd = static_cast<long long>(s) * d, sf = static_cast<long long>(s) * ff, cd = static_cast<long long>(c) * d;
  const long long cr = static_cast<long long>(c) * rank;
  g_sgemm_part = t.part;
  float *dx = t.dx, *da = t.da, *db_ = t.db_, *dc = t.dc, *dqkv = t.dqkv, *dffh = t.dffh, *dz = t.dz, *rowstat = t.rowstat;
  const float* xl = num_layers > 0 ? t.L[num_layers - 1].xout : t.x0;
  const int ne = 9 * embed;

  // ---------------- heads
  dyn_weight_grad_kernel<<<blocks_for(static_cast<long long>(c) * ne), 256, 0, st>>>(dw_oihw, dw_cin_ld, c, embed, decoder, w_scale, t.dwk);
  EOVAE_LAUNCH_CHECK();
  EOVAE_CUDA(cudaMemsetAsync(dx, 0, sizeof(float) * sd, st));
  // wk = hg W2h^T + b2h
  if (sgemm(t.dwk, 1, ne, t.hg, rank, 1, grads[9], rank, ne, rank, c, 0, st)) return -1;      // dW2h [9E][rank]
  colsum_f32_kernel<<<ceil_div(ne, 32), 32 * COLSUM_LANES, 0, st>>>(t.dwk, ne, c, ne, grads[10], 0);
  EOVAE_LAUNCH_CHECK();
  if (sgemm(t.dwk, ne, 1, params[9], rank, 1, t.dhg, rank, c, rank, ne, 0, st)) return -1;    // dhg [c][rank]
  gelu_bwd_kernel<<<blocks_for(cr), 256, 0, st>>>(t.hr, t.dhg, t.dhr, cr);
  EOVAE_LAUNCH_CHECK();
  if (sgemm(t.dhr, 1, rank, t.headin, d, 1, grads[7], d, rank, d, c, 0, st)) return -1;       // dW0 [rank][d]
  colsum_f32_kernel<<<ceil_div(rank, 32), 32 * COLSUM_LANES, 0, st>>>(t.dhr, rank, c, rank, grads[8], 0);
  EOVAE_LAUNCH_CHECK();
  float* dfeat = da;  // [c][d] gradient of features = T[128:128+C] + waves
  if (sgemm(t.dhr, rank, 1, params[7], d, 1, dfeat, d, c, d, rank, 0, st)) return -1;
  EOVAE_CUDA(cudaMemsetAsync(grads[2], 0, sizeof(float) * d, st));
  if (decoder) {
    axpy_kernel<<<blocks_for(c), 256, 0, st>>>(dbias, bias_scale, rowstat, c, 0);             // d(bias_raw) [c]
    EOVAE_LAUNCH_CHECK();
    if (sgemm(rowstat, 0, 1, t.headin2, d, 1, grads[11], d, 1, d, c, 0, st)) return -1;       // dwfb [1][d]
    colsum_f32_kernel<<<1, 32 * COLSUM_LANES, 0, st>>>(rowstat, 1, c, 1, grads[12], 0);
    EOVAE_LAUNCH_CHECK();
    if (sgemm(rowstat, 1, 0, params[11], 0, 1, db_, d, c, d, 1, 0, st)) return -1;            // dheadin2 [c][d]
    axpy_kernel<<<blocks_for(cd), 256, 0, st>>>(db_, 1.f, dfeat, cd, 1);                      // headin2 = features + btok
    EOVAE_LAUNCH_CHECK();
    colsum_f32_kernel<<<ceil_div(d, 32), 32 * COLSUM_LANES, 0, st>>>(db_, d, c, d, grads[2], 1);
    EOVAE_LAUNCH_CHECK();
  } else {
    axpy_kernel<<<blocks_for(embed), 256, 0, st>>>(dbias, bias_scale, grads[12], embed, 0);
    EOVAE_LAUNCH_CHECK();
    if (sgemm(grads[12], 1, 0, xl + (128 + c) * d, 0, 1, grads[11], d, embed, d, 1, 0, st)) return -1;
    if (sgemm(grads[12], 0, 1, params[11], d, 1, dx + (128 + c) * d, d, 1, d, embed, 1, st)) return -1;
  }
  axpy_kernel<<<blocks_for(cd), 256, 0, st>>>(dfeat, 1.f, dx + 128 * d, cd, 1);
  EOVAE_LAUNCH_CHECK();
  float* dwaves = da;  // the `waves` summand of features keeps dfeat until the token gradient is added below

  // ---------------- pre-norm layers, last to first (dx = gradient of the layer output)
  for (int l = num_layers - 1; l >= 0; --l) {
    const float* const* lp = params + 13 + 12 * l;
    float* const* lg = grads + 13 + 12 * l;
    FTapeLayer& L = t.L[l];
    const float* lin = l == 0 ? t.x0 : t.L[l - 1].xout;
    // xout = x1 + ffh W2^T + b2
    if (sgemm(dx, d, 1, lp[6], ff, 1, dffh, ff, s, ff, d, 0, st)) return -1;                  // dffh = dx W2
    if (sgemm(dx, 1, d, L.ffh, ff, 1, lg[6], ff, d, ff, s, 0, st)) return -1;                 // dW2 [d][ff]
    colsum_f32_kernel<<<ceil_div(d, 32), 32 * COLSUM_LANES, 0, st>>>(dx, d, s, d, lg[7], 0);
    EOVAE_LAUNCH_CHECK();
    gelu_bwd_kernel<<<blocks_for(sf), 256, 0, st>>>(L.z, dffh, dz, sf);
    EOVAE_LAUNCH_CHECK();
    // z = ln2 W1^T + b1
    float* dln2 = db_;
    if (sgemm(dz, ff, 1, lp[4], d, 1, dln2, d, s, d, ff, 0, st)) return -1;
    if (sgemm(dz, 1, ff, L.ln2, d, 1, lg[4], d, ff, d, s, 0, st)) return -1;                  // dW1 [ff][d]
    colsum_f32_kernel<<<ceil_div(ff, 32), 32 * COLSUM_LANES, 0, st>>>(dz, ff, s, ff, lg[5], 0);
    EOVAE_LAUNCH_CHECK();
    // ln2 = LN(x1; norm2): dx1 = dx + LN'(dln2)
    float* dx1 = dc;
    layernorm_bwd_kernel<<<ceil_div(s, 4), 128, 0, st>>>(L.x1, lp[10], dln2, dx1, rowstat, s, d, 1e-5f);
    EOVAE_LAUNCH_CHECK();
    layernorm_bwd_param_kernel<<<ceil_div(d, 32), 32 * COLSUM_LANES, 0, st>>>(L.x1, dln2, rowstat, lg[10], lg[11], s, d);
    EOVAE_LAUNCH_CHECK();
    axpy_kernel<<<blocks_for(sd), 256, 0, st>>>(dx, 1.f, dx1, sd, 1);
    EOVAE_LAUNCH_CHECK();
    // x1 = xin + att Wo^T + bo
    float* datt = db_;
    if (sgemm(dx1, d, 1, lp[2], d, 1, datt, d, s, d, d, 0, st)) return -1;
    if (sgemm(dx1, 1, d, L.att, d, 1, lg[2], d, d, d, s, 0, st)) return -1;                   // dWo [d][d]
    colsum_f32_kernel<<<ceil_div(d, 32), 32 * COLSUM_LANES, 0, st>>>(dx1, d, s, d, lg[3], 0);
    EOVAE_LAUNCH_CHECK();
    mha_bwd_q_kernel<<<ceil_div(s * heads, 4), 128, 0, st>>>(L.qkv, datt, t.pbuf, t.dsbuf, dqkv, s, d, heads);
    EOVAE_LAUNCH_CHECK();
    {
      const int hd = d / heads;
      if (sgemm_b(t.dsbuf, 1, s, static_cast<long long>(s) * s, L.qkv, 3 * d, 1, hd, dqkv + d, 3 * d, hd, heads, s, hd, s, 0, nullptr, st)) return -1;
      if (sgemm_b(t.pbuf, 1, s, static_cast<long long>(s) * s, datt, d, 1, hd, dqkv + 2 * d, 3 * d, hd, heads, s, hd, s, 0, nullptr, st)) return -1;
    }
    // qkv = ln1 Win^T + bin ; ln1 = LN(xin; norm1): dxin = dx1 + LN'(dqkv Win)
    float* dln1 = db_;
    if (sgemm(dqkv, 3 * d, 1, lp[0], d, 1, dln1, d, s, d, 3 * d, 0, st)) return -1;
    if (sgemm(dqkv, 1, 3 * d, L.ln1, d, 1, lg[0], d, 3 * d, d, s, 0, st)) return -1;           // dWin [3d][d]
    colsum_f32_kernel<<<ceil_div(3 * d, 32), 32 * COLSUM_LANES, 0, st>>>(dqkv, 3 * d, s, 3 * d, lg[1], 0);
    EOVAE_LAUNCH_CHECK();
    layernorm_bwd_kernel<<<ceil_div(s, 4), 128, 0, st>>>(lin, lp[8], dln1, dx, rowstat, s, d, 1e-5f);
    EOVAE_LAUNCH_CHECK();
    layernorm_bwd_param_kernel<<<ceil_div(d, 32), 32 * COLSUM_LANES, 0, st>>>(lin, dln1, rowstat, lg[8], lg[9], s, d);
    EOVAE_LAUNCH_CHECK();
    axpy_kernel<<<blocks_for(sd), 256, 0, st>>>(dx1, 1.f, dx, sd, 1);
    EOVAE_LAUNCH_CHECK();
  }

  // ---------------- tokens and FCResLayer
  EOVAE_CUDA(cudaMemcpyAsync(grads[1], dx, sizeof(float) * 128 * d, cudaMemcpyDeviceToDevice, st));
  axpy_kernel<<<blocks_for(d), 256, 0, st>>>(dx + (128 + c) * d, 1.f, grads[2], d, 1);
  EOVAE_LAUNCH_CHECK();
  axpy_kernel<<<blocks_for(cd), 256, 0, st>>>(dx + 128 * d, 1.f, dwaves, cd, 1);
  EOVAE_LAUNCH_CHECK();
  float* du2 = db_;
  relu_bwd_kernel<<<blocks_for(cd), 256, 0, st>>>(t.waves, t.emb, dwaves, du2, cd);
  EOVAE_LAUNCH_CHECK();
  if (sgemm(du2, 1, d, t.t1, d, 1, grads[5], d, d, d, c, 0, st)) return -1;
  colsum_f32_kernel<<<ceil_div(d, 32), 32 * COLSUM_LANES, 0, st>>>(du2, d, c, d, grads[6], 0);
  EOVAE_LAUNCH_CHECK();
  float* dt1 = dc;
  if (sgemm(du2, d, 1, params[5], d, 1, dt1, d, c, d, d, 0, st)) return -1;
  relu_bwd_kernel<<<blocks_for(cd), 256, 0, st>>>(t.t1, nullptr, dt1, dt1, cd);
  EOVAE_LAUNCH_CHECK();
  if (sgemm(dt1, 1, d, t.emb, d, 1, grads[3], d, d, d, c, 0, st)) return -1;
  colsum_f32_kernel<<<ceil_div(d, 32), 32 * COLSUM_LANES, 0, st>>>(dt1, d, c, d, grads[4], 0);
  EOVAE_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
