/*
 * eovae.h - C ABI of libeovae_sm100.so: the sm_100a (B200) kernels behind the EOFluxVAE hot path.
 *
 * The reference (nilsleh/eo-vae) is pure Python/PyTorch and has NO FFI of its own: its "operator API" for this
 * path is the set of torch library calls made by the nn.Modules listed below.  Each entry point here replaces
 * one of those call sites (paths relative to the reference root); the Python package eo-vae_b200/eo_vae binds
 * them with ctypes (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (torch caching allocator); nothing is retained
 *   - all launches are asynchronous on `stream` (a cudaStream_t passed as void*); no internal device sync
 *   - return 0 on success, negative on error; eovae_last_error() returns a thread-local message
 *   - activations are NHWC ("pixel-major"), 16-bit (EOVAE_BF16 / EOVAE_F16) unless stated; `*_pix_stride`
 *     is the distance between consecutive pixels in ELEMENTS (>= channels) so ops can read or write a channel
 *     slice of a wider tensor
 *   - unsupported shapes / dtypes are errors: there is no CPU or library fallback
 */
#ifndef EOVAE_H_
#define EOVAE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EOVAE_ABI_VERSION 2

/* dtype codes */
#define EOVAE_DT_BF16 0
#define EOVAE_DT_F16 1
#define EOVAE_DT_F32 2

/* conv modes */
#define EOVAE_CONV_3X3 0    /* 3x3, stride 1, zero pad 1                  (nn.Conv2d, layers.py:64,81)            */
#define EOVAE_CONV_1X1 1    /* 1x1                                        (layers.py:85,123-126; model.py:165,236) */
#define EOVAE_CONV_3X3_S2 2 /* F.pad(0,1,0,1) + 3x3 stride 2 pad 0        (Downsample, layers.py:33-37)            */

int eovae_version(void);
const char* eovae_last_error(void);
int eovae_num_sms(void);
/* number of kernels this library has launched in this process (bench.py reports the delta as gpu_launches) */
unsigned long long eovae_launch_count(void);
/* test / profiling aid: low byte = bit mask 1 skip epilogue work | 2 skip MMA issue | 4 skip TMA loads (these produce
 * garbage by design, tools/igemm_bench.py only); bits 8-9: 0 automatic | 1 force single-CTA | 2 force CTA-pair
 * (cta_group::2) implicit GEMM; bit 10: force one K-chunk per stage; bit 11: per-thread stores instead of the TMA-store epilogue (tests run all);
 * bit 12: no halo mainloop; bit 13: three taps per stage in the 16-channel conv; bit 14: 64-byte store rows everywhere;
 * bit 15: residual through registers; low byte also 32 = no bulk store issue, 64 = no tcgen05.ld (tools/epilogue_probe.py).
 * The product path never sets it.                         */
void eovae_set_debug_mode(int mode);
/* launch-shape knobs (process-wide; results never depend on them).
 * EOVAE_TUNE_GN_APPLY_CORESIDENT selects the kernel of eovae_gn_apply: 0 (default) = cp.async.bulk shared-memory ring
 * wherever the input is dense (else the register-load kernels); 1 = 128-thread CTAs with 8 loads in flight, <= 80 registers
 * and the maximum shared-memory carve-out, a shape that fits on an SM beside a resident implicit-GEMM CTA, so that the pass
 * overlaps a convolution running on another stream (the dual-stream encode); 2 = that launch shape, default carve-out;
 * 4 = register-load kernels with the size heuristic of round 1 (128 x 8 from 32 Mi elements up, else 256 x 4). */
#define EOVAE_TUNE_GN_APPLY_CORESIDENT 1
/* EOVAE_TUNE_GN_BWD_BLOCK_ELEMS: elements (pixels x channels) handled by one block of the GroupNorm backward kernels;
   0 (default) = chosen from the tensor size */
#define EOVAE_TUNE_GN_BWD_BLOCK_ELEMS 3
/* EOVAE_TUNE_GN_BWD_BULK: 1 (default) = cp.async.bulk shared-memory ring in the GroupNorm backward kernels, 0 = register loads */
#define EOVAE_TUNE_GN_BWD_BULK 4
/* EOVAE_TUNE_GN_APPLY_BLOCK_ELEMS: elements per block of the bulk-ring eovae_gn_apply kernel; 0 (default) = from the size */
#define EOVAE_TUNE_GN_APPLY_BLOCK_ELEMS 5
void eovae_set_tuning(int key, int value);

/* ---- weight packing (derived, non-persistent caches of the OIHW fp32 master parameters) -------------------- */
/* channels per K-chunk (in bytes: 32/64/128) and padded channels per tap chosen for a given Cin */
int eovae_conv_chunk_bytes(int cin);
int eovae_conv_k_per_tap(int cin);
/* OIHW fp32 [cout][cin][kh][kw] -> K-major 16-bit [round_up(cout,16)][kh*kw][k_per_tap(cin)], zero padded */
int eovae_pack_conv_weight(const float* w_oihw, void* out, int cout, int cin, int kh, int kw, int dtype, void* stream);

/* ---- tcgen05 implicit-GEMM convolution: replaces cuDNN conv2d at layers.py:64,81,85,123-126,33-37 and
 *      model.py:162,165,236,260.   out = scale * conv(x, w) + bias + residual                                   */
/*      Optional fused GroupNorm statistics of the OUTPUT (what the next GroupNorm of the network needs): when
 *      gn_stats != NULL the epilogue also emits per-tile partial sums (fixed slots, no atomics -> deterministic) and a
 *      tiny second kernel reduces them to gn_stats[n][gn_groups][2] = (mean, rstd).  eovae_conv2d_gn_workspace_bytes
 *      returns the workspace size, or 0 when the shape does not support the fusion (use eovae_gn_stats instead).     */
size_t eovae_conv2d_gn_workspace_bytes(int n, int h, int w, int mode, int cout, int groups);
int eovae_conv2d_gn_prologue_ok(int n, int h, int w, int cin, int cout, int mode, int groups);
int eovae_conv2d(const void* x, int n, int h, int w, int cin, long long x_pix_stride, int mode, const void* w_packed,
                 int cout, const float* bias, const void* residual, int res_dtype, long long res_pix_stride, void* out,
                 int out_dtype, long long out_pix_stride, int act_dtype, float scale, float* gn_stats, int gn_groups,
                 float gn_eps, void* gn_workspace, size_t gn_workspace_bytes, const void* x2, int cin2,
                 long long x2_pix_stride, const float* in_gn_stats, const float* in_gn_gamma, const float* in_gn_beta,
                 int in_gn_groups, void* in_gn_workspace, size_t in_gn_workspace_bytes, void* stream);
/*      Optional fused GroupNorm + SiLU of the INPUT (in_gn_stats != NULL): x is the RAW tensor and the mainloop
 *      normalises every operand tile in shared memory before the MMA reads it, i.e. out = conv(silu(GN(x))) without the
 *      normalised tensor ever touching HBM (layers.py:93-95,106-108).  in_gn_stats = [n][groups][2] (mean, rstd) of x,
 *      gamma / beta = the GroupNorm affine, workspace >= n*cin*8 bytes.  Only for shapes where
 *      eovae_conv2d_gn_prologue_ok(...) returns 1 (3x3 stride 1, image rows of >= 128 pixels, Cin % 64 == 0).
 *      Optional fused 1x1 operand (x2 != NULL, stride-1 modes): out += conv1x1(x2, w2) computed in the SAME mainloop -
 *      the ResnetBlock nin_shortcut (layers.py:85,111-112) folded into conv2.  w_packed then holds, per output row,
 *      the taps*k_per_tap(cin) columns of the main kernel followed by cin2 columns of the 1x1 kernel, and `bias` the
 *      sum of both biases.                                                                                           */

/* ---- batched C[b] = scale * A[b] (m x k) * B[b]^T (n x k): the q k^T and p v products of AttnBlock
 *      (F.scaled_dot_product_attention, layers.py:134-141)                                                      */
int eovae_gemm_tn_batched(const void* a, long long lda, long long a_batch_stride, const void* b, long long ldb,
                          long long b_batch_stride, void* c, int c_dtype, long long ldc, int batch, int m, int n, int k,
                          int a_dtype, int b_dtype, float scale, void* stream);

/* ---- Upsample (layers.py:40-50: nearest x2 then conv3x3) in sub-pixel form: four 2x2 convolutions on the LOW-resolution
 *      input, one per output parity (py, px), with the 3x3 taps that fall on the same source pixel summed (16 instead of 36
 *      MACs per output pixel pair; the upsampled tensor is never materialised).  x [n][h][w][cin] -> out [n][2h][2w][cout].
 *      eovae_pack_conv_weight_up2x: OIHW fp32 -> dgrad = 0: [4 phases][round_up(cout,16)][4 taps][k_per_tap(cin)]
 *                                                 dgrad = 1: [round_up(cin,16)][16 = phase*4+tap][k_per_tap(round_up(cout,8))]
 *      eovae_conv2d_up2x(_ok): forward (+ bias, + GroupNorm statistics of the output like eovae_conv2d);
 *      eovae_conv2d_up2x_dgrad: dy [n][h2][w2][cout] -> dx [n][h2/2][w2/2][cin], ONE launch reading the four parity
 *      sub-lattices of dy; eovae_conv2d_up2x_wgrad: dW (3x3 OIHW fp32) from the low-resolution input and dy.               */
int eovae_conv2d_up2x_ok(int n, int h, int w, int cin, int cout);
size_t eovae_conv2d_up2x_gn_workspace_bytes(int n, int h, int w, int cout, int groups);
int eovae_pack_conv_weight_up2x(const float* w_oihw, void* out, int cout, int cin, int dtype, int dgrad, void* stream);
int eovae_conv2d_up2x(const void* x, int n, int h, int w, int cin, long long x_pix_stride, const void* w_packed, int cout,
                      const float* bias, void* out, int out_dtype, long long out_pix_stride, int act_dtype, float* gn_stats,
                      int gn_groups, float gn_eps, void* gn_workspace, size_t gn_workspace_bytes, void* stream);
int eovae_conv2d_up2x_dgrad(const void* dy, int n, int h2, int w2, int cout, long long dy_pix_stride, const void* w_packed, int cin,
                            void* dx, int dx_dtype, long long dx_pix_stride, int act_dtype, void* stream);
/* data gradient of the Downsample conv (EOVAE_CONV_3X3_S2) in the same sub-pixel form: dy [n][ho][wo][cout] ->
 * dx [n][2ho][2wo][cin]; operand = eovae_pack_conv_weight_up2x(..., dgrad = 2): [4][round_up(cin,16)][4][k_per_tap(cout)] */
int eovae_conv2d_s2_dgrad(const void* dy, int n, int ho, int wo, int cout, long long dy_pix_stride, const void* w_packed, int cin,
                          void* dx, int dx_dtype, long long dx_pix_stride, int act_dtype, void* stream);
/* weight gradient of the Downsample conv: x [n][2ho][2wo][cin] read on its four parity sub-lattices, dy [n][ho][wo][cout] */
size_t eovae_conv2d_s2_wgrad_workspace_bytes(int n, int ho, int wo, int cin, int cout);
int eovae_conv2d_s2_wgrad(const void* x, long long x_pix_stride, const void* dy, long long dy_pix_stride, int dtype, int n, int ho,
                          int wo, int cin, int cout, float* dw_oihw, int accumulate, void* workspace, size_t workspace_bytes,
                          void* stream);
size_t eovae_conv2d_up2x_wgrad_workspace_bytes(int n, int h, int w, int cin, int cout);
int eovae_conv2d_up2x_wgrad(const void* x, long long x_pix_stride, const void* dy, long long dy_pix_stride, int dtype, int n, int h,
                            int w, int cin, int cout, float* dw_oihw, int accumulate, void* workspace, size_t workspace_bytes,
                            void* stream);

/* fp32 validation path (north_star: 1e-4 against the fp32 reference; selected by dtype code EOVAE_DT_F32 in eovae_conv2d,
 * eovae_gemm_tn_batched, eovae_gn_stats, eovae_gn_apply, eovae_nchw_to_nhwc16, eovae_softmax_rows, eovae_latent_denorm,
 * eovae_pack_conv_weight, eovae_pack_dyn_weight): SIMT fp32 FMAs, no tensor cores.  This entry is its general batched GEMM
 * c[i][r][j] = scale * sum_k a[i][r][k] * b[i*b_batch_stride + j*b_n_stride + k*b_k_stride]  (K a multiple of 16).       */
int eovae_gemm_strided_f32(const float* a, long long lda, long long a_batch_stride, const float* b, long long b_n_stride,
                           long long b_k_stride, long long b_batch_stride, float* c, long long ldc, int batch, int m, int n, int k,
                           float scale, void* stream);

/* ---- fused flash-style attention of AttnBlock (layers.py:134-141): out[n][q][:] = softmax_k(q . k / sqrt(c)) v for
 *      qkv [n][l][qkv_ld >= 3c] 16-bit (q | k | v channel slices of one tensor), out [n][l][out_ld >= c] 16-bit.  Scores
 *      and probabilities stay in TMEM / shared memory.  eovae_attention_fused_ok(l, c): c % 64 == 0, c <= 512.        */
int eovae_attention_fused_ok(int l, int c);
int eovae_attention_fused(const void* qkv, long long qkv_ld, int n, int l, int c, void* out, long long out_ld, int dtype,
                          void* stream);

/* ---- GroupNorm(32, eps) statistics + normalise/affine(/SiLU): replaces ATen group_norm + x*sigmoid(x)
 *      (layers.py:61,78,120; layers.py:21-22; model.py:159,193-194,295,349-350)
 *      stats: float [n][groups][2] = (mean, rstd);  workspace: eovae_gn_stats_workspace_bytes(...) bytes.
 *      Deterministic (fixed-order partial sums, no atomics).                                                    */
size_t eovae_gn_stats_workspace_bytes(int n, long long hw, int c, int groups);
int eovae_gn_stats(const void* x, int x_dtype, int n, long long hw, int c, long long pix_stride, int groups, float eps,
                   float* stats, void* workspace, size_t workspace_bytes, void* stream);
int eovae_gn_apply(const void* x, int x_dtype, long long x_pix_stride, const float* stats, const float* gamma,
                   const float* beta, void* y, int y_dtype, long long y_pix_stride, int n, long long hw, int c, int groups,
                   int apply_silu, void* stream);

/* ---- layout / edge kernels ----------------------------------------------------------------------------------- */
/* NCHW fp32 image batch -> NHWC 16-bit with channels zero-padded to c_pad (input edge of DynamicConv) */
int eovae_nchw_to_nhwc16(const float* x, void* out, int n, int c, int h, int w, int c_pad, int out_dtype, void* stream);
/* NHWC (fp32 or 16-bit) -> NCHW fp32 (output edge: moments / reconstructions) */
int eovae_nhwc_to_nchw_f32(const void* x, int x_dtype, long long x_pix_stride, float* out, int n, int c, int h, int w,
                           void* stream);
/* nearest-neighbour x2 upsample, NHWC 16-bit (F.interpolate, layers.py:48) */
int eovae_upsample2x(const void* x, void* out, int n, int h, int w, int c, void* stream);
/* row softmax: s [rows][s_ld >= cols] (fp32 or 16-bit) -> p [rows][p_ld >= cols] 16-bit, columns >= cols zeroed
 * (lets the p*v GEMM run on a K extent padded to the 16-element MMA step) */
int eovae_softmax_rows(const void* s, int s_dtype, long long s_ld, void* p, int p_dtype, long long p_ld, long long rows,
                       int cols, void* stream);
/* batched transpose of 16-bit matrices: in [batch][rows][in_ld>=cols] -> out [batch][cols][out_ld>=rows], zero padded */
int eovae_transpose16(const void* in, long long in_ld, void* out, long long out_ld, int batch, int rows, int cols,
                      void* stream);
/* in [batch][h*w][in_ld>=cols] (NHWC rows) -> out [ncopies][batch][cols][h*w_pad]: channel-major copies with the image
 * rows padded to w_pad; copy i holds the pixels shifted by first_shift + i along x (zero where the shift leaves the row and
 * in the pad columns).  Operands of eovae_conv2d_wgrad: the gradient as one unshifted copy, the 3x3 input as shifts -1..1 */
int eovae_transpose16_xshift(const void* in, long long in_ld, void* out, int batch, int h, int w, int w_pad, int cols,
                             int first_shift, int ncopies, void* stream);

/* ---- posterior + latent glue (distributions.py:20-67, new_autoencoder.py:466-469,533-543,730-738) ------------ */
/* moments fp32, logical [n][2*zc][h][w] with HOST array mstrides[4] = element strides (n, c, y, x)
 * -> z_norm NCHW fp32 [n][zc][h][w]:
 *   pixel-unshuffle(2) -> BatchNorm2d(eval, running stats, eps) -> pixel-shuffle(2) collapses to a per
 *   (channel, row parity, col parity) affine on the posterior mean                                               */
int eovae_latent_norm(const float* moments, const long long* mstrides, const float* running_mean,
                      const float* running_var, float eps, float* z, int n, int h, int w, int zc, void* stream);
/* z (spatial, normalised) NCHW fp32 -> inverse BN (running stats, eps) -> NHWC 16-bit decoder input */
int eovae_latent_denorm(const float* z, const float* running_mean, const float* running_var, float eps, void* out,
                        int out_dtype, int n, int h, int w, int zc, void* stream);
/* fused reparameterisation + KL: z = mean + exp(0.5*clamp(logvar)) * eps ; kl[n] = 0.5*sum(mean^2+var-1-logvar)
 * moments fp32 with host strides as above; eps, z NCHW fp32 (eps NULL -> z = mean; z NULL -> KL only)          */
int eovae_kl_reparam(const float* moments, const long long* mstrides, const float* eps, float* z, float* kl, int n,
                     int h, int w, int zc, void* stream);

/* ---- wavelength hypernetwork (dynamic_conv.py:37-59,110-130,162-183,352-366,499-525,666-697), all fp32 -------- */
/* params: host array of device pointers: 0 omega[d/2] (sincos frequencies) | 1 weight_tokens | 2 bias_token |
 *   3,4 fclayer.w1 (w,b) | 5,6 fclayer.w2 | 7,8 fc_weight | 9,10 fc_bias | then 12 per transformer layer:
 *   in_proj (w,b), out_proj (w,b), linear1 (w,b), linear2 (w,b), norm1 (g,b), norm2 (g,b)
 * wk_out  : [C][9*embed] raw fc_weight output (unscaled)
 * bias_out: encoder [embed], decoder [C]  (raw fc_bias output, unscaled)
 * workspace: eovae_hypernet_workspace_bytes(...) bytes                                                           */
size_t eovae_hypernet_workspace_bytes(int c, int d, int ff, int embed);
int eovae_hypernet_forward(const float* wvs_um, int c, const float* const* params, int num_layers, int d, int heads,
                           int ff, int embed, int decoder, float* wk_out, float* bias_out, void* workspace,
                           size_t workspace_bytes, void* stream);
/* generated kernel -> igemm B operand (16-bit, K-major, zero padded), optional fp32 OIHW copy (both scaled by
 * `scale`), and bias_out[i] = bias_raw[i] * bias_scale (nbias entries)                                          */
int eovae_pack_dyn_weight(const float* wk, int c, int embed, int decoder, float scale, void* packed, int dtype,
                          int k_per_tap, int rows_pad, float* oihw_out, const float* bias_raw, float bias_scale,
                          float* bias_out, int nbias, void* stream);

/* ---- losses (consistency_loss.py:12-21,418; 24-37 via torchmetrics MS-SSIM) ---------------------------------- */
/* out[0] = mean |a-b| , out[1] = mean sqrt((a-b)^2 + eps^2); a, b fp32, any layout (elementwise) */
int eovae_l1_charbonnier(const float* a, const float* b, long long count, float eps, float* out, void* workspace,
                         size_t workspace_bytes, void* stream);

/* ----  latent statistics of the encode_latents pipeline (encode_latents.py:36-109, RunningStatsButFast):
 * per-channel mean / unbiased variance / min / max of x [n][c][hw] fp32 merged into the running state (all device
 * buffers: mean, var, std, vmin, vmax [c], count [1]; workspace 4*c floats) without a host round trip */
int eovae_running_stats_update(const float* x, int n, int c, long long hw, float* mean, float* var, float* std, float* count,
                               float* vmin, float* vmax, float* workspace, void* stream);

/* ----  data-side prologue in one gather pass: raw batch [n][c][hi][wi] (EOVAE_DT_F32, or 16-bit integer
 * digital numbers: 3 = int16, 4 = uint16) -> optional clip -> bilinear resize to hr x wr (align_corners = False) ->
 * D4 augmentation (horizontal flip, vertical flip, rot_k x 90 degrees counter-clockwise) -> (v - mean[c]) / (std[c] + eps),
 * NCHW fp32 [n][c][ho][wo] with (ho, wo) = (hr, wr) swapped for odd rot_k.  terramesh_datamodule.py:189-197,347-369,476-482 */
int eovae_preprocess(const void* in, int in_dtype, int n, int c, int hi, int wi, int hr, int wr, const float* mean,
                     const float* std, float std_eps, int do_clip, float clip_lo, float clip_hi, int flip_h, int flip_v, int rot_k,
                     float* out, void* stream);

/* mean multi-scale SSIM over the batch (out[0]) and per sample (per_sample[b], may be NULL); pred, target fp32 NCHW
 * [b][c][h][w], h and w multiples of 16 and >= 176; 5 scales, 11-tap sigma-1.5 Gaussian, reflect padding, relu
 * normalisation, betas (0.0448, 0.2856, 0.3001, 0.2363, 0.1333): torchmetrics' algorithm as called by the reference
 * (consistency_loss.py:24-37).  loss = 1 - out[0].                                                                  */
size_t eovae_msssim_workspace_bytes(int b, int c, int h, int w);
int eovae_msssim(const float* pred, const float* target, int b, int c, int h, int w, float data_range, float* out,
                 float* per_sample, void* workspace, size_t workspace_bytes, void* stream);

/* ---- backward pass (first kernels of the training path, new_autoencoder.py:648 manual_backward) ------------------ */
/* dgrad operand: conv data gradient = eovae_conv2d(dy, pack_dgrad(W), Cout' = Cin) with the flipped / transposed kernel:
 * OIHW fp32 [cout][cin][kh][kw] -> [round_up(cin,16)][kh*kw][k_per_tap(round_up(cout,8))]                            */
int eovae_pack_conv_weight_dgrad(const float* w_oihw, void* out, int cout, int cin, int kh, int kw, int dtype,
                                 void* stream);
/* GroupNorm(+SiLU) backward (dense NHWC 16-bit x, grad_out): grad_x = dL/dx (+ grad_add if not NULL),
 * dgamma/dbeta fp32 [c] (optionally accumulated); grad_x or the parameter outputs may be NULL; grad_x_colsum (fp32 [c],
 * may be NULL) receives the per-channel sum of grad_x = the bias gradient of the conv that produced x, from the same pass */
size_t eovae_gn_backward_workspace_bytes(int n, long long hw, int c, int groups);
/* dtype = storage type of the forward activation x; grad_dtype = storage type of grad_out / grad_add / grad_x (the
 * training default mixes f16 activations with bf16 gradients) */
int eovae_gn_backward(const void* x, const void* grad_out, int dtype, int grad_dtype, const float* stats, const float* gamma,
                      const float* beta, int n, long long hw, int c, int groups, int with_silu, const void* grad_add,
                      void* grad_x, float* dgamma, float* dbeta, int accumulate_params, float* grad_x_colsum,
                      void* workspace, size_t workspace_bytes, void* stream);
/* dy [n][ho][wo][c] -> z [n][h][w][c] = 0 except z[2i+1][2j+1] = dy[i][j] (adjoint of the Downsample stride-2 gather) */
int eovae_scatter_stride2(const void* dy, void* z, int n, int ho, int wo, int h, int w, int c, void* stream);
/* out [n][h][w][c] = sum over the 2x2 blocks of g [n][2h][2w][c] (adjoint of nearest x2 upsampling) */
int eovae_pool2x2_sum(const void* g, void* out, int dtype, int n, int h, int w, int c, void* stream);
/* weight gradient of a 3x3 (stride 1, pad 1) or 1x1 convolution on tcgen05, contracted over the pixels:
 * dy_t: CHANNEL-MAJOR 16-bit copy [n][cout][h*w] (eovae_transpose16_xshift, one unshifted copy); x_t: [n][cin][h*w] for
 * 1x1, the three x-shifted copies [3][n][cin][h*w] for 3x3; here `w` is the PADDED row length (a multiple of 8);
 * dw_oihw fp32 [cout][cin][k][k] (optionally accumulated).  Deterministic (fixed-order split-K reduction).          */
size_t eovae_conv2d_wgrad_workspace_bytes(int n, int h, int w, int cin, int cout, int ksize);
int eovae_conv2d_wgrad(const void* x_t, const void* dy_t, int dtype, int dy_dtype, int n, int h, int w, int cin, int cout, int ksize,
                       float* dw_oihw, int accumulate, void* workspace, size_t workspace_bytes, void* stream);
/* attention backward: ds = scale * p o (dp - rowsum(dp o p)); p 16-bit [rows][p_ld], dp fp32, ds 16-bit zero padded to
 * out_cols (layers.py:134-141 adjoint) */
int eovae_softmax_backward(const void* p, long long p_ld, const float* dp, long long dp_ld, void* ds, long long ds_ld,
                           int dtype, int ds_dtype, long long rows, int cols, int out_cols, float scale, void* stream);
/* adjoint of the reparameterisation in eovae_kl_reparam (distributions.py:44-46): dz NCHW fp32 [n][zc][h][w] ->
 * dmoments dense NCHW fp32 [n][2zc][h][w] (mean half = dz, logvar half = dz*eps*std/2 inside the clamp) */
int eovae_reparam_backward(const float* moments, const long long* mstrides, const float* eps, const float* dz, float* dmoments,
                           int n, int h, int w, int zc, void* stream);
/* gradient of eovae_l1_charbonnier wrt a: kind 0 = L1, 1 = Charbonnier; *grad_scale = upstream scalar (device) */
int eovae_pixel_loss_backward(const float* a, const float* b, long long count, float eps, int kind, const float* grad_scale,
                              float* grad_a, void* stream);
/* backward of eovae_hypernet_forward + eovae_pack_dyn_weight: dw_oihw = gradient of the generated conv kernel
 * ([embed][dw_cin_ld >= c][3][3] for the encoder layer, [c][dw_cin_ld >= embed][3][3] for the decoder layer), dbias the
 * gradient of the scaled bias; w_scale / bias_scale the factors eovae_pack_dyn_weight applied.  grads[i] receives the
 * gradient of params[i] (same order as the forward; written, not accumulated; grads[0] is ignored).  tape_valid = 0: the
 * forward is re-run inside (activations kept in the workspace); 1: the workspace still holds eovae_hypernet_forward_taped's.  dynamic_conv.py:110-130,162-183,352-366,511-525,684-697 adjoint. */
size_t eovae_hypernet_backward_workspace_bytes(int c, int d, int ff, int embed, int num_layers);
int eovae_hypernet_backward(const float* wvs_um, int c, const float* const* params, int num_layers, int d, int heads,
                            int ff, int embed, int decoder, const float* dw_oihw, int dw_cin_ld, float w_scale,
                            const float* dbias, float bias_scale, float* const* grads, int tape_valid, void* workspace,
                            size_t workspace_bytes, void* stream);
/* forward that ALSO leaves its activations in `workspace` (size eovae_hypernet_backward_workspace_bytes) so that a
 * following eovae_hypernet_backward(..., tape_valid = 1, same workspace) does not re-run it */
int eovae_hypernet_forward_taped(const float* wvs_um, int c, const float* const* params, int num_layers, int d, int heads,
                                 int ff, int embed, int decoder, float* wk_out, float* bias_out, void* workspace,
                                 size_t workspace_bytes, void* stream);
/* ---- FactorizedWeightGenerator(_decoder) variant of the wavelength hypernetwork (dynamic_conv.py:186-302; selected by
 * generator_type='factorized', configs/finetune_consistency_factor.yaml:55-73): pre-norm transformer layers
 * (norm_first=True, ff = 4 d) and a low-rank head Linear(d, rank) -> GELU -> Linear(rank, 9 embed).  Outputs as
 * eovae_hypernet_forward.  The forward always leaves its activations in `workspace`
 * (eovae_hypernet_factorized_workspace_bytes) for eovae_hypernet_factorized_backward.  Dropout is not applied.
 * params: 0 omega[d/2] | 1 weight_tokens | 2 bias_token | 3,4 fclayer.w1 (w,b) | 5,6 fclayer.w2 | 7,8 fc_weight.0
 * [rank][d] | 9,10 fc_weight.2 [9 embed][rank] | 11,12 fc_bias | per layer (12): in_proj, out_proj, linear1, linear2,
 * norm1, norm2 (w,b each) */
size_t eovae_hypernet_factorized_workspace_bytes(int c, int d, int ff, int embed, int rank, int num_layers);
int eovae_hypernet_factorized_forward(const float* wvs_um, int c, const float* const* params, int num_layers, int d,
                                      int heads, int ff, int embed, int rank, int decoder, float* wk_out, float* bias_out,
                                      void* workspace, size_t workspace_bytes, void* stream);
/* adjoint of the above + eovae_pack_dyn_weight (arguments as eovae_hypernet_backward); `workspace` must still hold the
 * activations of eovae_hypernet_factorized_forward for the same inputs */
int eovae_hypernet_factorized_backward(const float* wvs_um, int c, const float* const* params, int num_layers, int d,
                                       int heads, int ff, int embed, int rank, int decoder, const float* dw_oihw,
                                       int dw_cin_ld, float w_scale, const float* dbias, float bias_scale,
                                       float* const* grads, void* workspace, size_t workspace_bytes, void* stream);
/* ---- AdaIN conditioning of the ResnetBlocks (use_adain=True in dynamic_conv_kwargs; model.py:96-100, 173-191, 331-343) --
 * WavelengthConditioner.forward (model.py:35-64): style[d] = mlp(mean over bands of sincos(wvs)), mlp = Linear(d, 2d) ->
 * SiLU -> Linear(2d, d) -> SiLU -> Linear(d, d); wvs in micrometres (no x1000 here).  The style depends on the wavelength
 * vector only, so one row is computed (the reference repeats it over the batch).
 * params: 0 omega[d/2] | 1,2 mlp.0 (w,b) | 3,4 mlp.2 | 5,6 mlp.4.  The forward leaves its activations in `workspace`
 * (eovae_wavelength_style_workspace_bytes) for the backward; grads[i] <- d/d params[i] (written; grads[0] unused). */
size_t eovae_wavelength_style_workspace_bytes(int d);
int eovae_wavelength_style_forward(const float* wvs_um, int c, const float* const* params, int d, float* style,
                                   void* workspace, size_t workspace_bytes, void* stream);
int eovae_wavelength_style_backward(const float* const* params, int d, const float* dstyle, float* const* grads,
                                    void* workspace, size_t workspace_bytes, void* stream);
/* ResnetBlock AdaIN (layers.py:68-76, 96-104): [scale | shift] = emb_proj(style); GroupNorm(h)*gamma+beta followed by
 * *scale + shift is the same GroupNorm with gamma' = gamma*scale, beta' = beta*scale + shift, which is what the
 * GroupNorm-apply / conv-prologue kernels then consume.  style2 [2 cout] keeps emb_proj's output for the backward.
 * Backward: dgamma_out / dbeta_out are the (batch-summed) affine gradients of the GroupNorm backward kernel. */
int eovae_adain_affine_forward(const float* style, int d, const float* wproj, const float* bproj, const float* gamma,
                               const float* beta, int cout, float* gamma_out, float* beta_out, float* style2, void* stream);
int eovae_adain_affine_backward(const float* style, int d, const float* wproj, const float* gamma, const float* beta,
                                const float* style2, int cout, const float* dgamma_out, const float* dbeta_out,
                                float* dgamma, float* dbeta, float* dwproj, float* dbproj, float* dstyle, float* dstyle2,
                                void* stream);
/* ---- optional EOConsistencyLoss branches, fp32 NCHW contiguous, out = one device scalar ------------------
 * SAMLoss.forward (consistency_loss.py:186-210): mean over (b, pixel) of 1 - <p, t> / (|p| |t| + eps), channel axis = 1.
 * workspace: 16 bytes.  Backward: gradient wrt pred times the device scalar *grad_scale. */
int eovae_sam_loss(const float* pred, const float* target, int b, int c, long long hw, float eps, float* out, void* workspace,
                   size_t workspace_bytes, void* stream);
int eovae_sam_loss_backward(const float* pred, const float* target, int b, int c, long long hw, float eps,
                            const float* grad_scale, float* grad_pred, void* stream);
/* GradientDifferenceLoss.forward with alpha = 1 (consistency_loss.py:241-269): mean | |dx p| - |dx t| | + mean | |dy p| - |dy t| |
 * over `planes` = B*C images of h x w (forward differences).  workspace: 16 bytes. */
int eovae_grad_diff_loss(const float* pred, const float* target, long long planes, int h, int w, float* out, void* workspace,
                         size_t workspace_bytes, void* stream);
int eovae_grad_diff_loss_backward(const float* pred, const float* target, long long planes, int h, int w,
                                  const float* grad_scale, float* grad_pred, void* stream);
/* FocalFrequencyLoss.forward (ffl.py:17-104) as EOConsistencyLoss builds it (consistency_loss.py:388-395: patch_factor,
 * alpha, ave_spectrum = False, batch_matrix = True, log_matrix = True): orthonormal 2-D DFT of (pred - target) per
 * H/pf x W/pf patch (dense fp32 DFT-matrix products, any patch size), weight = clamp(log1p(|Z|^alpha) / batch max, 0, 1)
 * (detached), out = mean(weight |Z|^2).  keep_for_backward = 1 leaves weight * Z in `workspace` for the backward, which
 * returns d out / d pred times the device scalar *grad_scale. */
size_t eovae_focal_freq_loss_workspace_bytes(int b, int c, int h, int w, int patch_factor);
int eovae_focal_freq_loss(const float* pred, const float* target, int b, int c, int h, int w, int patch_factor, float alpha,
                          int keep_for_backward, float* out, void* workspace, size_t workspace_bytes, void* stream);
int eovae_focal_freq_loss_backward(int b, int c, int h, int w, int patch_factor, const float* grad_scale, float* grad_pred,
                                   void* workspace, size_t workspace_bytes, void* stream);
/* ---- EQ-VAE transforms of the training step (new_autoencoder.py:460-464, 519-531, 611-636), fp32 NCHW contiguous ------
 * out = torch.rot90(F.interpolate(z, size=(nh, nw), mode='bilinear', align_corners=False), k=rot_k, dims=[-1, -2]);
 * out is [planes][nw][nh] when rot_k is odd, else [planes][nh][nw].  nh = h, nw = w: rotation only. */
int eovae_latent_resize_rot(const float* z, long long planes, int h, int w, int nh, int nw, int rot_k, float* out, void* stream);
int eovae_latent_resize_rot_backward(const float* grad_out, long long planes, int h, int w, int nh, int nw, int rot_k,
                                     float* grad_z, void* stream);
/* reconstruction target: torch.rot90(F.interpolate(x, size=(nh, nw), mode='area'), k=rot_k, dims=[-1, -2]) (no gradient) */
int eovae_area_resize_rot(const float* x, long long planes, int h, int w, int nh, int nw, int rot_k, float* out, void* stream);
/* ---- optimiser half of the training step (new_autoencoder.py:549-557 torch.optim.Adam(lr); :650-657 clip_grad_norm_ +
 * step), multi-tensor, fp32.  The tensors are described by DEVICE tables the caller builds: pointer arrays [T], sizes [T]
 * (elements) and a chunk table (chunk i covers elements [chunk_offset[i], + chunk_elems) of tensor chunk_tensor[i]).
 * eovae_grad_norm: *out_norm = sqrt(sum over all tensors of g^2) (fixed-slot partials [num_chunks], deterministic).
 * eovae_adam_step: torch.optim.Adam's update (no weight decay / amsgrad / maximize) at step number `step` (1-based) with
 * g scaled by min(max_norm / (*grad_norm + 1e-6), 1) when grad_norm != NULL and max_norm > 0 (clip_grad_norm_ folded in). */
int eovae_grad_norm(const float* const* grads, const long long* sizes, const int* chunk_tensor, const long long* chunk_offset,
                    int num_chunks, int chunk_elems, float* partial, float* out_norm, void* stream);
int eovae_adam_step(float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                    const long long* sizes, const int* chunk_tensor, const long long* chunk_offset, int num_chunks,
                    int chunk_elems, float lr, float beta1, float beta2, float eps, int step, const float* grad_norm,
                    float max_norm, void* stream);
/* gradient of eovae_msssim's batch-mean value wrt pred (fp32 NCHW), times the device scalar *grad_scale; the forward
 * pyramid is rebuilt inside the workspace (consistency_loss.py:24-37 adjoint) */
size_t eovae_msssim_backward_workspace_bytes(int b, int c, int h, int w);
int eovae_msssim_backward(const float* pred, const float* target, int b, int c, int h, int w, float data_range,
                          const float* grad_scale, float* grad_pred, void* workspace, size_t workspace_bytes, void* stream);
/* the same weight gradient read straight from the NHWC tensors (x: conv input, dy: output gradient, pixel pitches in
 * elements) as MN-major tcgen05 operands - no transposed copies.  Usable when eovae_conv2d_wgrad_nhwc_ok(h, w) returns 1
 * (W divides 64 or is a multiple of 64, H a multiple of 64 / min(W, 64)); otherwise use eovae_conv2d_wgrad. */
int eovae_conv2d_wgrad_nhwc_ok(int h, int w);
size_t eovae_conv2d_wgrad_nhwc_workspace_bytes(int n, int h, int w, int cin, int cout, int ksize);
int eovae_conv2d_wgrad_nhwc(const void* x, long long x_pix_stride, const void* dy, long long dy_pix_stride, int dtype, int dy_dtype, int n, int h,
                            int w, int cin, int cout, int ksize, float* dw_oihw, int accumulate, void* workspace,
                            size_t workspace_bytes, void* stream);
/* TRAIN-mode latent glue (new_autoencoder.py:466-469 with self.training, :533-543): z NCHW fp32 [n][zc][h][w] ->
 * pixel-unshuffle(2) -> BatchNorm2d batch statistics (biased variance, eps_bn) with the running statistics [4*zc] updated
 * in place (momentum, unbiased variance) -> inverse normalisation with the UPDATED running statistics (eps_inv) ->
 * pixel-shuffle -> decoder input, NHWC 16-bit.  save [4*zc][3] = (batch mean, rstd, inverse scale) for the backward. */
int eovae_latent_bn_train_forward(const float* z, int n, int zc, int h, int w, float* running_mean, float* running_var,
                                  float momentum, float eps_bn, float eps_inv, void* out, int out_dtype, long long out_pix_stride,
                                  float* save, void* stream);
int eovae_latent_bn_train_backward(const void* dout, int dtype, long long dout_pix_stride, const float* z, int n, int zc, int h, int w,
                                   const float* save, float* dz, void* stream);
/* dbias[c] (+)= sum over pixels of grad_out [pixels][c] (16-bit) */
size_t eovae_bias_grad_workspace_bytes(long long pixels, int c);
int eovae_bias_grad(const void* grad_out, int dtype, long long pixels, int c, float* dbias, int accumulate, void* workspace,
                    size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EOVAE_H_ */
