// Focal frequency loss branch of EOConsistencyLoss (freq_weight > 0; ffl.py:17-104 with the options consistency_loss.py:
// 388-395 fixes: ave_spectrum = False, batch_matrix = True, log_matrix = True, matrix = None).
//
// The 2-D transform is linear, so F(pred) - F(target) = F(pred - target): ONE orthonormal 2-D DFT of the difference
// per patch.  Patches are H/pf x W/pf (128 x 128 at the shipped patch_factor 2; any size, not only powers of two), so the
// DFT runs as dense fp32 matrix products with the (symmetric) DFT matrices W_h, W_w on the batched SIMT GEMM:
//   Y = W_h d (2 real GEMMs), Z = Y W_w (4 real GEMMs).
// weight = clamp(log1p(|Z|^alpha) / max over the whole batch, 0, 1) (detached), loss = mean(weight |Z|^2).
// Backward: d loss / d pred = (2 / N) Re(IDFT2(weight Z)) scattered back to the image - 6 more real GEMMs.
// (The reference's nan_to_num on the two spectra is the identity for finite inputs and is not reproduced.)
#include "../../include/eovae.h"
#include "common.cuh"

namespace {

// DFT matrix of size p, orthonormal: w[j][k] = exp(-2 pi i jk / p) / sqrt(p); (jk mod p) keeps the angle exact
__global__ void dft_matrix_kernel(int p, float* __restrict__ wr, float* __restrict__ wi, float* __restrict__ wi_neg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p * p) return;
  const int j = i / p, k = i % p;
  const int r = static_cast<int>((static_cast<long long>(j) * k) % p);
  float s, c;
  sincospif(2.0f * static_cast<float>(r) / static_cast<float>(p), &s, &c);
  const float inv = rsqrtf(static_cast<float>(p));
  wr[i] = c * inv;
  wi[i] = -s * inv;
  wi_neg[i] = s * inv;
}

// d[(b, c, py, px)][y][x] = pred - target of that patch
__global__ void ffl_diff_kernel(const float* __restrict__ a, const float* __restrict__ b, int h, int w, int ph, int pw, int pf,
                                long long planes, float* __restrict__ d) {
  const long long total = planes * pf * pf * ph * pw;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % pw), y = static_cast<int>((i / pw) % ph);
    const long long patch = i / (static_cast<long long>(ph) * pw);
    const int px = static_cast<int>(patch % pf), py = static_cast<int>((patch / pf) % pf);
    const long long plane = patch / (pf * pf);
    const long long src = (plane * h + py * ph + y) * w + px * pw + x;
    d[i] = __ldg(&a[src]) - __ldg(&b[src]);
  }
}

__device__ __forceinline__ float ffl_metric(float zr, float zi, float alpha) {
  const float r = sqrtf(zr * zr + zi * zi + 1e-8f);
  return log1pf(alpha == 1.f ? r : powf(r, alpha));
}

__global__ void __launch_bounds__(256) ffl_max_kernel(const float* __restrict__ zr, const float* __restrict__ zi, long long n,
                                                      float alpha, unsigned int* __restrict__ max_bits) {
  float m = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    m = fmaxf(m, ffl_metric(zr[i], zi[i], alpha));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(max_bits, __float_as_uint(m));  // non-negative floats order like their bits
}

// loss partial sums; (zr, zi) <- weight * (zr, zi) for the backward when `keep` is set
__global__ void __launch_bounds__(256) ffl_sum_kernel(float* __restrict__ zr, float* __restrict__ zi, long long n, float alpha,
                                                      const unsigned int* __restrict__ max_bits, int keep,
                                                      double* __restrict__ ws) {
  float mx = __uint_as_float(max_bits[0]);
  if (!(isfinite(mx) && mx > 0.f)) mx = 1.f;
  float s = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float r = zr[i], q = zi[i];
    const float wgt = fminf(fmaxf(ffl_metric(r, q, alpha) / mx, 0.f), 1.f);
    s += wgt * (r * r + q * q);
    if (keep) {
      zr[i] = wgt * r;
      zi[i] = wgt * q;
    }
  }
  __shared__ float red[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int i = 0; i < (blockDim.x >> 5); ++i) t += red[i];
    atomicAdd(&ws[0], t);
  }
}
__global__ void ffl_finalize_kernel(const double* ws, double n, float* out) { out[0] = static_cast<float>(ws[0] / n); }

// grad[image] = scale * g[(patch)][y][x]
__global__ void ffl_scatter_kernel(const float* __restrict__ g, int h, int w, int ph, int pw, int pf, long long planes,
                                   const float* __restrict__ gscale, float k, float* __restrict__ grad) {
  const float sc = gscale[0] * k;
  const long long total = planes * pf * pf * ph * pw;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % pw), y = static_cast<int>((i / pw) % ph);
    const long long patch = i / (static_cast<long long>(ph) * pw);
    const int px = static_cast<int>(patch % pf), py = static_cast<int>((patch / pf) % pf);
    const long long plane = patch / (pf * pf);
    grad[(plane * h + py * ph + y) * w + px * pw + x] = sc * g[i];
  }
}

unsigned grid_for(long long n) {
  long long blocks = (n + 255) / 256;
  const long long cap = 8LL * eovae_num_sms();
  if (blocks > cap) blocks = cap;
  return static_cast<unsigned>(blocks < 1 ? 1 : blocks);
}

struct FflBuffers {
  float *whr, *whi, *whn, *wwr, *wwi, *wwn, *d, *yr, *yi, *zr, *zi;
  double* sum;
  unsigned int* max_bits;
};

size_t carve(FflBuffers* f, void* workspace, long long n, int ph, int pw) {
  char* base = static_cast<char*>(workspace);
  char* p = base;
  auto take = [&](size_t bytes) { char* q = p; p += (bytes + 255) / 256 * 256; return q; };
  FflBuffers tmp;
  FflBuffers& F = f != nullptr ? *f : tmp;
  F.sum = reinterpret_cast<double*>(take(16));
  F.max_bits = reinterpret_cast<unsigned int*>(take(16));
  F.whr = reinterpret_cast<float*>(take(sizeof(float) * ph * ph));
  F.whi = reinterpret_cast<float*>(take(sizeof(float) * ph * ph));
  F.whn = reinterpret_cast<float*>(take(sizeof(float) * ph * ph));
  F.wwr = reinterpret_cast<float*>(take(sizeof(float) * pw * pw));
  F.wwi = reinterpret_cast<float*>(take(sizeof(float) * pw * pw));
  F.wwn = reinterpret_cast<float*>(take(sizeof(float) * pw * pw));
  F.d = reinterpret_cast<float*>(take(sizeof(float) * n));
  F.yr = reinterpret_cast<float*>(take(sizeof(float) * n));
  F.yi = reinterpret_cast<float*>(take(sizeof(float) * n));
  F.zr = reinterpret_cast<float*>(take(sizeof(float) * n));
  F.zi = reinterpret_cast<float*>(take(sizeof(float) * n));
  return static_cast<size_t>(p - base);
}

int ffl_check(int b, int c, int h, int w, int pf) {
  EOVAE_CHECK(b >= 1 && c >= 1 && pf >= 1 && h >= pf && w >= pf, "focal_freq_loss: bad shape (%d, %d, %d, %d) / patch factor %d", b, c, h, w, pf);
  EOVAE_CHECK(h % pf == 0 && w % pf == 0, "focal_freq_loss: H (%d) and W (%d) must be divisible by the patch factor %d", h, w, pf);
  return 0;
}

}  // namespace

extern "C" {

size_t eovae_focal_freq_loss_workspace_bytes(int b, int c, int h, int w, int patch_factor) {
  if (patch_factor < 1 || h < patch_factor || w < patch_factor) return 0;
  return carve(nullptr, nullptr, static_cast<long long>(b) * c * h * w, h / patch_factor, w / patch_factor);
}

/* forward; with keep_for_backward the workspace afterwards holds weight * spectrum for eovae_focal_freq_loss_backward */
int eovae_focal_freq_loss(const float* pred, const float* target, int b, int c, int h, int w, int patch_factor, float alpha,
                          int keep_for_backward, float* out, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  if (ffl_check(b, c, h, w, patch_factor)) return -1;
  EOVAE_CHECK(workspace_bytes >= eovae_focal_freq_loss_workspace_bytes(b, c, h, w, patch_factor), "focal_freq_loss: workspace too small");
  const int pf = patch_factor, ph = h / pf, pw = w / pf;
  const long long planes = static_cast<long long>(b) * c, n = planes * h * w;
  const long long pp = static_cast<long long>(ph) * pw;
  const int patches = static_cast<int>(planes * pf * pf);
  FflBuffers f;
  carve(&f, workspace, n, ph, pw);
  EOVAE_CUDA(cudaMemsetAsync(f.sum, 0, 16, st));
  EOVAE_CUDA(cudaMemsetAsync(f.max_bits, 0, 16, st));
  dft_matrix_kernel<<<ceil_div(ph * ph, 256), 256, 0, st>>>(ph, f.whr, f.whi, f.whn);
  EOVAE_LAUNCH_CHECK();
  dft_matrix_kernel<<<ceil_div(pw * pw, 256), 256, 0, st>>>(pw, f.wwr, f.wwi, f.wwn);
  EOVAE_LAUNCH_CHECK();
  ffl_diff_kernel<<<grid_for(n), 256, 0, st>>>(pred, target, h, w, ph, pw, pf, planes, f.d);
  EOVAE_LAUNCH_CHECK();
  // Y = W_h d
  if (eovae::sgemm_batched(f.whr, ph, 1, 0, f.d, pw, 1, pp, f.yr, pw, pp, patches, ph, pw, ph, 0, st)) return -1;
  if (eovae::sgemm_batched(f.whi, ph, 1, 0, f.d, pw, 1, pp, f.yi, pw, pp, patches, ph, pw, ph, 0, st)) return -1;
  // Z = Y W_w : Zr = Yr Wr - Yi Wi, Zi = Yr Wi + Yi Wr
  if (eovae::sgemm_batched(f.yr, pw, 1, pp, f.wwr, pw, 1, 0, f.zr, pw, pp, patches, ph, pw, pw, 0, st)) return -1;
  if (eovae::sgemm_batched(f.yi, pw, 1, pp, f.wwn, pw, 1, 0, f.zr, pw, pp, patches, ph, pw, pw, 1, st)) return -1;
  if (eovae::sgemm_batched(f.yr, pw, 1, pp, f.wwi, pw, 1, 0, f.zi, pw, pp, patches, ph, pw, pw, 0, st)) return -1;
  if (eovae::sgemm_batched(f.yi, pw, 1, pp, f.wwr, pw, 1, 0, f.zi, pw, pp, patches, ph, pw, pw, 1, st)) return -1;
  ffl_max_kernel<<<grid_for(n), 256, 0, st>>>(f.zr, f.zi, n, alpha, f.max_bits);
  EOVAE_LAUNCH_CHECK();
  ffl_sum_kernel<<<grid_for(n), 256, 0, st>>>(f.zr, f.zi, n, alpha, f.max_bits, keep_for_backward, f.sum);
  EOVAE_LAUNCH_CHECK();
  ffl_finalize_kernel<<<1, 1, 0, st>>>(f.sum, static_cast<double>(n), out);
  EOVAE_LAUNCH_CHECK();
  return 0;
}

/* gradient wrt pred times the device scalar *grad_scale; `workspace` = the one eovae_focal_freq_loss(..., keep_for_backward
 * = 1) ran in (same shape arguments) */
int eovae_focal_freq_loss_backward(int b, int c, int h, int w, int patch_factor, const float* grad_scale, float* grad_pred,
                                   void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  if (ffl_check(b, c, h, w, patch_factor)) return -1;
  EOVAE_CHECK(workspace_bytes >= eovae_focal_freq_loss_workspace_bytes(b, c, h, w, patch_factor), "focal_freq_loss_backward: workspace too small");
  const int pf = patch_factor, ph = h / pf, pw = w / pf;
  const long long planes = static_cast<long long>(b) * c, n = planes * h * w;
  const long long pp = static_cast<long long>(ph) * pw;
  const int patches = static_cast<int>(planes * pf * pf);
  FflBuffers f;
  carve(&f, workspace, n, ph, pw);
  // T = G conj(W_w): Tr = Gr Wr + Gi Wi, Ti = Gi Wr - Gr Wi      (G = weight * Z in zr / zi; T in yr / yi)
  if (eovae::sgemm_batched(f.zr, pw, 1, pp, f.wwr, pw, 1, 0, f.yr, pw, pp, patches, ph, pw, pw, 0, st)) return -1;
  if (eovae::sgemm_batched(f.zi, pw, 1, pp, f.wwi, pw, 1, 0, f.yr, pw, pp, patches, ph, pw, pw, 1, st)) return -1;
  if (eovae::sgemm_batched(f.zi, pw, 1, pp, f.wwr, pw, 1, 0, f.yi, pw, pp, patches, ph, pw, pw, 0, st)) return -1;
  if (eovae::sgemm_batched(f.zr, pw, 1, pp, f.wwn, pw, 1, 0, f.yi, pw, pp, patches, ph, pw, pw, 1, st)) return -1;
  // Re(conj(W_h) T) = Wr Tr + Wi Ti
  if (eovae::sgemm_batched(f.whr, ph, 1, 0, f.yr, pw, 1, pp, f.d, pw, pp, patches, ph, pw, ph, 0, st)) return -1;
  if (eovae::sgemm_batched(f.whi, ph, 1, 0, f.yi, pw, 1, pp, f.d, pw, pp, patches, ph, pw, ph, 1, st)) return -1;
  ffl_scatter_kernel<<<grid_for(n), 256, 0, st>>>(f.d, h, w, ph, pw, pf, planes, grad_scale, static_cast<float>(2.0 / static_cast<double>(n)),
                                                 grad_pred);
  EOVAE_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
