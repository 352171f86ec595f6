// Library plumbing (error slot, device query) and weight-packing kernels.
#include "../../include/eovae.h"
#include "common.cuh"

namespace eovae {
static thread_local char g_err[1024] = "";
unsigned long long g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  set_error("CUDA error %d (%s) at %s", static_cast<int>(e), cudaGetErrorString(e), what);
  return -2;
}
}  // namespace eovae

int eovae_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
  }
  return sms;
}

namespace {

template <typename T>
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, T* __restrict__ out, int cout, int cin, int taps,
                                        int kpt, long long total) {
  // out[o][tap][c] (c < kpt), zero for o >= cout or c >= cin
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = static_cast<int>(i % kpt);
  long long t = i / kpt;
  const int tap = static_cast<int>(t % taps);
  const int o = static_cast<int>(t / taps);
  float v = 0.f;
  if (o < cout && c < cin) v = w[(static_cast<long long>(o) * cin + c) * taps + tap];
  out[i] = T16<T>::from_f(v);
}

// dgrad operand: the data gradient of y = conv(x, W) is conv(dy, W') with W'[ci][co][kh][kw] = W[co][ci][K-1-kh][K-1-kw]
template <typename T>
__global__ void pack_conv_weight_dgrad_kernel(const float* __restrict__ w, T* __restrict__ out, int cout, int cin, int taps,
                                              int kpt, long long total) {
  // out[ci][tap][co] (co < kpt = padded Cout), rows padded to round_up(cin, 16)
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int co = static_cast<int>(i % kpt);
  long long t = i / kpt;
  const int tap = static_cast<int>(t % taps);
  const int ci = static_cast<int>(t / taps);
  float v = 0.f;
  if (ci < cin && co < cout) v = w[(static_cast<long long>(co) * cin + ci) * taps + (taps - 1 - tap)];
  out[i] = T16<T>::from_f(v);
}

// Coalesced versions of the two packs (every conv weight is re-packed twice per training step: forward operand and data-
// gradient operand).  Forward: one block per output channel stages w[o] ([cin][taps] fp32, contiguous) in shared memory and
// writes out[o][tap][0..kpt) as 32-bit pairs.
constexpr int PF_THREADS = 512;
template <typename T>
__global__ void __launch_bounds__(PF_THREADS) pack_conv_weight_tiled_kernel(const float* __restrict__ w, T* __restrict__ out, int cout,
                                                                     int cin, int taps, int kpt) {
  extern __shared__ float sm[];
  const int o = blockIdx.x;
  const int n = cin * taps;
  if (o < cout) {
    constexpr int LB = 9;  // loads in batches (see the data-gradient pack): 512 x 9 = one batch for a 512-channel 3x3 row
    for (int i0 = threadIdx.x; i0 < n; i0 += PF_THREADS * LB) {
      float v[LB];
#pragma unroll
      for (int u = 0; u < LB; ++u) v[u] = i0 + u * PF_THREADS < n ? __ldg(&w[static_cast<long long>(o) * n + i0 + u * PF_THREADS]) : 0.f;
#pragma unroll
      for (int u = 0; u < LB; ++u)
        if (i0 + u * PF_THREADS < n) sm[i0 + u * PF_THREADS] = v[u];
    }
  }
  __syncthreads();
  const int half = kpt >> 1;
  uint32_t* orow = reinterpret_cast<uint32_t*>(out + static_cast<long long>(o) * taps * kpt);
  // thread -> (tap lane, channel pair) once, then strided loops: no per-element division by the runtime extents
  const int hw_ = half < PF_THREADS ? half : PF_THREADS;  // channel pairs covered per pass
  const int tstep = PF_THREADS / hw_;                     // taps covered per pass
  const int h0 = threadIdx.x % hw_, t0 = threadIdx.x / hw_;
  if (t0 < tstep) {
    for (int tap = t0; tap < taps; tap += tstep) {
      for (int h = h0; h < half; h += hw_) {
        const int c = 2 * h;
        const float v0 = (o < cout && c < cin) ? sm[c * taps + tap] : 0.f;
        const float v1 = (o < cout && c + 1 < cin) ? sm[(c + 1) * taps + tap] : 0.f;
        orow[tap * half + h] = T16<T>::from_f2(v0, v1);
      }
    }
  }
}

// Data-gradient operand out[ci][tap][co] = w[co][ci][taps-1-tap]: a (64 co) x (16 ci) tile goes through shared memory so that
// both the fp32 reads (16 * taps contiguous floats per co) and the 16-bit writes (64 contiguous co per (ci, tap)) coalesce.
constexpr int PD_CO = 64, PD_CI = 16, PD_THREADS = 1024;
template <typename T, int TAPS>
__global__ void __launch_bounds__(PD_THREADS) pack_conv_weight_dgrad_tiled_kernel(const float* __restrict__ w, T* __restrict__ out,
                                                                           int cout, int cin, int cin_pad, int kpt) {
  // TAPS is a template parameter: the index arithmetic below divides by it per element (runtime divisions made this
  // kernel 5x slower than its 14 MB of traffic)
  __shared__ float sm[PD_CO][PD_CI * TAPS + 1];
  const int ci0 = blockIdx.x * PD_CI, co0 = blockIdx.y * PD_CO;
  constexpr int seg = PD_CI * TAPS;
  // all loads of a thread in ONE batch (in-order issue: a load consumed by the next instruction serialises the loop at the
  // memory latency - it was 36 round trips per block): 1024 threads x 9 elements cover the 64 x 144 tile
  constexpr int LB = (PD_CO * seg + PD_THREADS - 1) / PD_THREADS;
  for (int i0 = threadIdx.x; i0 < PD_CO * seg; i0 += PD_THREADS * LB) {
    float v[LB];
#pragma unroll
    for (int u = 0; u < LB; ++u) {
      const int i = i0 + u * PD_THREADS;
      const int r = i / seg, j = i % seg;
      const int co = co0 + r, ci = ci0 + j / TAPS;
      v[u] = (i < PD_CO * seg && co < cout && ci < cin) ? __ldg(&w[(static_cast<long long>(co) * cin + ci0) * TAPS + j]) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < LB; ++u) {
      const int i = i0 + u * PD_THREADS;
      if (i < PD_CO * seg) sm[i / seg][i % seg] = v[u];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < seg * (PD_CO / 2); i += PD_THREADS) {
    const int pair = i % (PD_CO / 2), t = i / (PD_CO / 2);
    const int tap = t % TAPS, cil = t / TAPS;
    const int ci = ci0 + cil, co = co0 + 2 * pair;
    if (ci < cin_pad && co < kpt) {
      const int src = cil * TAPS + (TAPS - 1 - tap);
      *reinterpret_cast<uint32_t*>(out + (static_cast<long long>(ci) * TAPS + tap) * kpt + co) =
          T16<T>::from_f2(sm[2 * pair][src], sm[2 * pair + 1][src]);
    }
  }
}

template <typename T>
__global__ void pack_dyn_weight_kernel(const float* __restrict__ wk, int c, int embed, int decoder, float scale,
                                       T* __restrict__ packed, int kpt, int rows_pad, float* __restrict__ oihw,
                                       const float* __restrict__ bias_raw, float bias_scale,
                                       float* __restrict__ bias_out, int nbias) {
  if (blockIdx.x == 0 && bias_out != nullptr)
    for (int i = threadIdx.x; i < nbias; i += blockDim.x) bias_out[i] = bias_raw[i] * bias_scale;
  // wk: [c][9*embed], flat (tap*embed + e).  Encoder conv weight W[e][band][tap]; decoder W[band][e][tap].
  const int rows = decoder ? c : embed;  // Cout
  const int cin = decoder ? embed : c;
  const long long total = static_cast<long long>(rows_pad) * 9 * kpt;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(i % kpt);
    long long t = i / kpt;
    const int tap = static_cast<int>(t % 9);
    const int o = static_cast<int>(t / 9);
    float v = 0.f;
    if (o < rows && ci < cin) {
      const int band = decoder ? o : ci;
      const int e = decoder ? ci : o;
      v = wk[static_cast<long long>(band) * 9 * embed + tap * embed + e] * scale;
      if (oihw != nullptr) oihw[(static_cast<long long>(o) * cin + ci) * 9 + tap] = v;
    }
    packed[i] = T16<T>::from_f(v);
  }
}

}  // namespace

extern "C" {

int eovae_version(void) { return EOVAE_ABI_VERSION; }
const char* eovae_last_error(void) { return eovae::g_err; }
unsigned long long eovae_launch_count(void) { return eovae::g_launches; }

int eovae_pack_conv_weight(const float* w_oihw, void* out, int cout, int cin, int kh, int kw, int dtype, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  EOVAE_CHECK((kh == 3 && kw == 3) || (kh == 1 && kw == 1), "pack_conv_weight: only 3x3 and 1x1 kernels");
  const int taps = kh * kw;
  const int kpt = eovae_conv_k_per_tap(cin);
  const long long total = static_cast<long long>(round_up(cout, 16)) * taps * kpt;
  const unsigned grid = static_cast<unsigned>((total + 255) / 256);
  const size_t stage = sizeof(float) * static_cast<size_t>(cin) * taps;
  if (stage <= 40 * 1024 && (dtype == EOVAE_BF16 || dtype == EOVAE_F16)) {  // coalesced path: one block per output channel
    const int rows = round_up(cout, 16);
    if (dtype == EOVAE_BF16)
      pack_conv_weight_tiled_kernel<__nv_bfloat16><<<rows, PF_THREADS, stage, stream>>>(w_oihw, static_cast<__nv_bfloat16*>(out), cout, cin, taps, kpt);
    else
      pack_conv_weight_tiled_kernel<__half><<<rows, PF_THREADS, stage, stream>>>(w_oihw, static_cast<__half*>(out), cout, cin, taps, kpt);
    EOVAE_LAUNCH_CHECK();
    return 0;
  }
  if (dtype == EOVAE_BF16)
    pack_conv_weight_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(w_oihw, static_cast<__nv_bfloat16*>(out), cout, cin, taps, kpt, total);
  else if (dtype == EOVAE_F16)
    pack_conv_weight_kernel<__half><<<grid, 256, 0, stream>>>(w_oihw, static_cast<__half*>(out), cout, cin, taps, kpt, total);
  else if (dtype == EOVAE_F32)  // fp32 validation path: same [cout][tap][k_per_tap] layout, unrounded
    pack_conv_weight_kernel<float><<<grid, 256, 0, stream>>>(w_oihw, static_cast<float*>(out), cout, cin, taps, kpt, total);
  else
    EOVAE_CHECK(false, "pack_conv_weight: bad dtype %d", dtype);
  EOVAE_LAUNCH_CHECK();
  return 0;
}

int eovae_pack_conv_weight_dgrad(const float* w_oihw, void* out, int cout, int cin, int kh, int kw, int dtype,
                                 void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  EOVAE_CHECK((kh == 3 && kw == 3) || (kh == 1 && kw == 1), "pack_conv_weight_dgrad: only 3x3 and 1x1 kernels");
  const int taps = kh * kw;
  const int kpt = eovae_conv_k_per_tap(round_up(cout, 8));
  const long long total = static_cast<long long>(round_up(cin, 16)) * taps * kpt;
  const unsigned grid = static_cast<unsigned>((total + 255) / 256);
  if (dtype == EOVAE_BF16 || dtype == EOVAE_F16) {  // coalesced path: (64 co) x (16 ci) tiles through shared memory
    const int cin_pad = round_up(cin, 16);
    dim3 tg(cin_pad / PD_CI, ceil_div(kpt, PD_CO));
    if (dtype == EOVAE_BF16) {
      if (taps == 9) pack_conv_weight_dgrad_tiled_kernel<__nv_bfloat16, 9><<<tg, PD_THREADS, 0, stream>>>(w_oihw, static_cast<__nv_bfloat16*>(out), cout, cin, cin_pad, kpt);
      else pack_conv_weight_dgrad_tiled_kernel<__nv_bfloat16, 1><<<tg, PD_THREADS, 0, stream>>>(w_oihw, static_cast<__nv_bfloat16*>(out), cout, cin, cin_pad, kpt);
    } else {
      if (taps == 9) pack_conv_weight_dgrad_tiled_kernel<__half, 9><<<tg, PD_THREADS, 0, stream>>>(w_oihw, static_cast<__half*>(out), cout, cin, cin_pad, kpt);
      else pack_conv_weight_dgrad_tiled_kernel<__half, 1><<<tg, PD_THREADS, 0, stream>>>(w_oihw, static_cast<__half*>(out), cout, cin, cin_pad, kpt);
    }
    EOVAE_LAUNCH_CHECK();
    return 0;
  }
  if (dtype == EOVAE_BF16)
    pack_conv_weight_dgrad_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(w_oihw, static_cast<__nv_bfloat16*>(out), cout, cin, taps, kpt, total);
  else if (dtype == EOVAE_F16)
    pack_conv_weight_dgrad_kernel<__half><<<grid, 256, 0, stream>>>(w_oihw, static_cast<__half*>(out), cout, cin, taps, kpt, total);
  else
    EOVAE_CHECK(false, "pack_conv_weight_dgrad: bad dtype %d", dtype);
  EOVAE_LAUNCH_CHECK();
  return 0;
}

int eovae_pack_dyn_weight(const float* wk, int c, int embed, int decoder, float scale, void* packed, int dtype,
                          int k_per_tap, int rows_pad, float* oihw_out, const float* bias_raw, float bias_scale,
                          float* bias_out, int nbias, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int rows = decoder ? c : embed;
  const int cin = decoder ? embed : c;
  EOVAE_CHECK(rows_pad >= rows && k_per_tap >= cin, "pack_dyn_weight: padding smaller than extent");
  const long long total = static_cast<long long>(rows_pad) * 9 * k_per_tap;
  long long blocks = (total + 255) / 256;
  if (blocks > 1024) blocks = 1024;
  if (dtype == EOVAE_BF16)
    pack_dyn_weight_kernel<__nv_bfloat16><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(wk, c, embed, decoder, scale, static_cast<__nv_bfloat16*>(packed), k_per_tap, rows_pad, oihw_out, bias_raw, bias_scale, bias_out, nbias);
  else if (dtype == EOVAE_F16)
    pack_dyn_weight_kernel<__half><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(wk, c, embed, decoder, scale, static_cast<__half*>(packed), k_per_tap, rows_pad, oihw_out, bias_raw, bias_scale, bias_out, nbias);
  else if (dtype == EOVAE_F32)
    pack_dyn_weight_kernel<float><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(wk, c, embed, decoder, scale, static_cast<float*>(packed), k_per_tap, rows_pad, oihw_out, bias_raw, bias_scale, bias_out, nbias);
  else
    EOVAE_CHECK(false, "pack_dyn_weight: bad dtype %d", dtype);
  EOVAE_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
