// Internal interface of the fp32 validation path (fp32_path.cu); reached through the public C-ABI entry points when a
// tensor's dtype code is EOVAE_F32.
#pragma once
#include <cuda_runtime.h>

namespace eovae {
namespace f32 {
// out[img][pix][n] = scale * sum_{tap, c} x[img][pix + tap][c] * B[img](n, tap * kpt + c) + bias[n] + res[img][pix][n]
// B element (n, k) at w + img * w_bs + n * w_ns + k * w_ks (w_bs = 0: one weight matrix shared by all images)
int conv2d(const float* x, int n, int h, int w, int cin, long long x_ps, int mode, const float* wt, long long w_ns,
           long long w_ks, long long w_bs, int kpt, int cout, const float* bias, const float* res, long long res_ps, float* out,
           long long out_ps, float scale, cudaStream_t stream);
int gn_stats(const float* x, int n, long long hw, int c, long long ps, int groups, float eps, float* stats, cudaStream_t stream);
int gn_apply(const float* x, long long x_ps, const float* stats, const float* gamma, const float* beta, float* y, long long y_ps,
             int n, long long hw, int c, int groups, int silu, cudaStream_t stream);
int nchw_to_nhwc(const float* x, float* out, int n, int c, int h, int w, int c_pad, cudaStream_t stream);
}  // namespace f32
}  // namespace eovae
