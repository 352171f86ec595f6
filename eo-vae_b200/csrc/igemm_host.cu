// Host side of the tcgen05 implicit-GEMM: TMA descriptor construction, tile-shape selection and launch.
// Public C-ABI entry points are declared in include/eovae.h.
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "../../include/eovae.h"
#include "fp32_path.cuh"
#include "igemm_sm100.cuh"

namespace {

int g_debug_mode = 0;
int g_force_ctas = 0;  // 0 = auto, 1 / 2 = forced (tests, tools/igemm_bench.py)
int g_force_kch1 = 0;  // 1 = always one K-chunk per stage
int g_no_res_tma = 0;     // 1 = residual rows through registers even where the TMA path applies
int g_no_wide_store = 0;  // 1 = 64-byte-row output boxes everywhere
int g_no_kch9 = 0;     // 1 = narrow-channel 3x3 convs stage three taps (not all nine) per pipeline stage
int g_no_tma_store = 0;  // 1 = epilogue writes with per-thread 16-byte stores instead of bulk tensor stores
int g_no_halo = 0;       // 1 = never use the halo-reuse mainloop

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                              const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode_fn() {
  static EncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(p);
  });
  return fn;
}

// rank-`rank` tiled map over 16-bit elements; strides in BYTES for dims 1..rank-1.
int encode_map(CUtensorMap* map, int dtype, int rank, const void* base, const uint64_t* dims, const uint64_t* strides,
               const uint32_t* box, int chunk_bytes, bool promote = true) {
  EncodeFn fn = get_encode_fn();
  EOVAE_CHECK(fn != nullptr, "cuTensorMapEncodeTiled entry point not available (driver too old?)");
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
  }
  for (int i = 0; i < rank - 1; ++i) gstr[i] = strides[i];
  CUtensorMapSwizzle sw = chunk_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                             : (chunk_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUtensorMapDataType dt = dtype == EOVAE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = fn(map, dt, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim, gstr, bdim, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, promote ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    eovae::set_error(
        "cuTensorMapEncodeTiled failed (%d): rank %d base %p dims [%llu %llu %llu %llu] strides [%llu %llu %llu] box [%u %u %u %u] chunk %d",
        static_cast<int>(r), rank, base, (unsigned long long)dims[0], (unsigned long long)dims[1],
        (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
        (unsigned long long)strides[0], (unsigned long long)(rank > 2 ? strides[1] : 0),
        (unsigned long long)(rank > 3 ? strides[2] : 0), box[0], box[1], rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0,
        chunk_bytes);
    return -3;
  }
  return 0;
}

template <int BLOCK_N, int CHUNK_BYTES, int CTAS, int KCH = 1, bool HALO = false, bool GNP = false>
int launch_t(const igemm::Params& p, int total_work, cudaStream_t stream) {
  using Cfg = igemm::Config<BLOCK_N, CHUNK_BYTES, CTAS, KCH, HALO>;
  static_assert(Cfg::STAGES >= 2, "pipeline needs at least two stages");
  auto kern = igemm::igemm_kernel<BLOCK_N, CHUNK_BYTES, CTAS, KCH, HALO, GNP>;
  static bool attr_set = false;  // benign race: idempotent
  if (!attr_set) {
    EOVAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES + (GNP ? igemm::GNP_COEF_BYTES : 0)));
    attr_set = true;
  }
  const int max_groups = eovae_num_sms() / CTAS;  // one CTA (pair) per SM (pair), persistent over the work items
  const int groups = total_work < max_groups ? total_work : max_groups;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(groups * CTAS);
  cfg.blockDim = dim3(GNP ? igemm::NUM_THREADS_GNP : igemm::NUM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES + (GNP ? igemm::GNP_COEF_BYTES : 0);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTAS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  EOVAE_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  EOVAE_LAUNCH_CHECK();
  return 0;
}

template <int CHUNK_BYTES, int CTAS>
int launch_n(int block_n, const igemm::Params& p, int total_work, cudaStream_t stream) {
  switch (block_n) {
    case 16: return launch_t<16, CHUNK_BYTES, CTAS>(p, total_work, stream);
    case 32: return launch_t<32, CHUNK_BYTES, CTAS>(p, total_work, stream);
    case 64: return launch_t<64, CHUNK_BYTES, CTAS>(p, total_work, stream);
    case 128: return launch_t<128, CHUNK_BYTES, CTAS>(p, total_work, stream);
    case 256: return launch_t<256, CHUNK_BYTES, CTAS>(p, total_work, stream);
  }
  eovae::set_error("igemm: unsupported BLOCK_N %d for a %d-CTA group", block_n, CTAS);
  return -1;
}

int pick_block_n(int cout_pad) {
  if (cout_pad >= 256 && cout_pad % 256 == 0) return 256;
  if (cout_pad >= 128 && cout_pad % 128 == 0) return 128;
  if (cout_pad >= 256) return 256;  // ragged last tile handled by TMA OOB fill + epilogue mask
  if (cout_pad >= 128) return 128;
  if (cout_pad >= 64) return 64;
  if (cout_pad >= 32) return 32;
  return 16;
}

// stats[n][g] = (mean, rstd) from the per-tile partial sums written by the igemm epilogue ([image][slot][group][2] floats).
// grid (groups / gsub, images): a block owns `gsub` (<= 4) adjacent groups of one image = one 32-byte sector per slot;
// thread (slot lane, group) walks the slots with a stride of `lanes`.  (One warp per (n, g) read one 8-byte pair per sector
// and took 10 us on the 2048-slot level-0 tensors; one block per image left a batch-16 training step with 16 blocks.)
// Fixed summation order (deterministic), double accumulation.
__global__ void __launch_bounds__(256) gn_tiles_finalize_kernel(const float* __restrict__ partial, float* __restrict__ stats,
                                                                int groups, int gsub, int slots_per_img, double count,
                                                                float eps, int regions, long long region_stride) {
  __shared__ double fin_s[256], fin_q[256];
  const int n = blockIdx.y;
  const int lanes = blockDim.x / gsub;
  const int gl = threadIdx.x % gsub, sl = threadIdx.x / gsub;
  const int g = blockIdx.x * gsub + gl;
  double s = 0.0, q = 0.0;
  if (sl < lanes) {
    for (int r = 0; r < regions; ++r) {  // one region per sub-pixel phase (1 for an ordinary convolution)
      const float2* base = reinterpret_cast<const float2*>(partial + r * region_stride) +
                           static_cast<long long>(n) * slots_per_img * groups + g;
#pragma unroll 4
      for (int t = sl; t < slots_per_img; t += lanes) {
        const float2 v = base[static_cast<long long>(t) * groups];
        s += v.x;
        q += v.y;
      }
    }
  }
  fin_s[threadIdx.x] = s;
  fin_q[threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.x < gsub) {
    s = 0.0;
    q = 0.0;
    for (int l = 0; l < lanes; ++l) {
      s += fin_s[l * gsub + gl];
      q += fin_q[l * gsub + gl];
    }
    const double mean = s / count;
    double var = q / count - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[2 * (n * groups + g)] = static_cast<float>(mean);
    stats[2 * (n * groups + g) + 1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  }
}

// per-(image, channel) affine of GroupNorm for the fused prologue: z = x * a + b
// packed16 (fp16 operands): .x = bits of half2(m16, a / 2) with m16 = the fp16 value nearest the mean, .y = bh =
// (beta + (m16 - mean) a) / 2 in fp32 - the coefficients of h = z / 2 = (x - m16) (a / 2) + bh used by the half2 transform
__global__ void gn_ab_kernel(const float* __restrict__ stats, const float* __restrict__ gamma,
                             const float* __restrict__ beta, float2* __restrict__ ab, int c, int groups, int total, int packed16) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int n = i / c, ch = i % c;
  const int g = ch / (c / groups);
  const float mean = stats[(n * groups + g) * 2], rstd = stats[(n * groups + g) * 2 + 1];
  const float a = gamma[ch] * rstd;
  if (!packed16) {
    ab[i] = make_float2(a, beta[ch] - mean * a);
    return;
  }
  const __half m16 = __float2half_rn(mean);
  const __half ah = __float2half_rn(0.5f * a);
  const uint32_t bits = static_cast<uint32_t>(__half_as_ushort(m16)) | (static_cast<uint32_t>(__half_as_ushort(ah)) << 16);
  ab[i] = make_float2(__uint_as_float(bits), 0.5f * (beta[ch] + (__half2float(m16) - mean) * a));
}

void m_tiling(int n, int ho, int wo, bool batched, int* bw, int* bh, int* bn);

struct GnPrologue {  // GroupNorm(+SiLU) of the conv INPUT, applied inside the mainloop (halo kernels only)
  const float* stats;
  const float* gamma;
  const float* beta;
  int groups;
  void* workspace;  // >= N * Cin * 8 bytes
};

// Custom tap geometry (mode = EOVAE_CONV_CUSTOM): the two sub-pixel forms of the nearest-x2 upsample convolution.
//  * phases = 4, parity_in = false: FORWARD.  Four 2x2 convolutions on the low-resolution input; phase q = (py, px) writes
//    out[:, py::2, px::2, :] of the high-resolution output (out_h2 x out_w2) and uses weight rows [q * rows_pad, ...).
//  * phases = 1, parity_in = true : DATA GRADIENT.  16 taps, tap t reads the parity sub-lattice tap_map[t] of the
//    high-resolution gradient (the tensor given as the A operand, extent 2H x 2W) at offset (tap_dx, tap_dy); output low-res.
struct TapSpec {
  int num_taps;
  int tap_map[igemm::MAX_TAPS], tap_dx[igemm::MAX_TAPS], tap_dy[igemm::MAX_TAPS];
  bool parity_in;
  int phases;
  int phase_dx[4], phase_dy[4];
  int b_phase_rows;
};
constexpr int EOVAE_CONV_CUSTOM = 99;

struct ASpec {       // activation-side operand: NHWC tensor view
  const void* ptr;
  int N, H, W, C;    // logical extent (C = channels visible to the contraction)
  long long pix_stride;  // elements between pixels (>= C; lets a conv read a channel slice of a wider tensor)
};

int launch_igemm(const ASpec& a, int mode, const void* w, int k_per_tap, int chunk_bytes, int cout, long long w_rows,
                 long long w_row_stride, long long w_batch_stride, int w_batches, const float* bias, const void* res, int res_dtype,
                 long long res_pix_stride, void* out, int out_dtype, long long out_pix_stride, int act_dtype,
                 float scale, cudaStream_t stream, float* gn_stats = nullptr, int gn_groups = 0, float gn_eps = 0.f,
                 void* gn_ws = nullptr, size_t gn_ws_bytes = 0, const ASpec* extra = nullptr,
                 const GnPrologue* gnp = nullptr, int w_dtype = -1, const TapSpec* taps = nullptr) {
  if (w_dtype < 0) w_dtype = act_dtype;
  EOVAE_CHECK((act_dtype == EOVAE_BF16 || act_dtype == EOVAE_F16) && (w_dtype == EOVAE_BF16 || w_dtype == EOVAE_F16),
              "igemm: operand dtypes must be bf16/f16");
  // the instruction descriptor has separate A / B format fields, but sm_100a traps (illegal instruction, measured) on
  // kind::f16 with f16 x bf16 operands: refuse instead of faulting.  EOVAE_ALLOW_MIXED_MMA=1 lifts the check (probe only).
  EOVAE_CHECK(act_dtype == w_dtype || getenv("EOVAE_ALLOW_MIXED_MMA") != nullptr,
              "igemm: A (%d) and B (%d) operand formats must be equal (tcgen05 kind::f16 rejects mixed f16/bf16)", act_dtype, w_dtype);
  EOVAE_CHECK(chunk_bytes == 32 || chunk_bytes == 64 || chunk_bytes == 128, "igemm: bad chunk bytes %d", chunk_bytes);
  EOVAE_CHECK(a.pix_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(a.ptr) % 16) == 0,
              "igemm: activation pixel stride (%lld) must be a multiple of 8 elements and base 16B aligned", a.pix_stride);
  EOVAE_CHECK(k_per_tap % (chunk_bytes / 2) == 0, "igemm: k_per_tap %d not a multiple of chunk", k_per_tap);
  EOVAE_CHECK(out_pix_stride % (out_dtype == EOVAE_F32 ? 4 : 8) == 0 && (reinterpret_cast<uintptr_t>(out) % 16) == 0,
              "igemm: output pixel stride (%lld) / base not 16-byte aligned", out_pix_stride);
  EOVAE_CHECK(res == nullptr || (res_pix_stride % (res_dtype == EOVAE_F32 ? 4 : 8) == 0 &&
                                 (reinterpret_cast<uintptr_t>(res) % 16) == 0),
              "igemm: residual pixel stride (%lld) / base not 16-byte aligned", res_pix_stride);
  igemm::Params p;
  memset(&p, 0, sizeof(p));
  p.phases = 1;
  const int ch = chunk_bytes / 2;
  int Ho = a.H, Wo = a.W;
  p.num_taps = (mode == EOVAE_CONV_1X1) ? 1 : 9;
  if (mode == EOVAE_CONV_3X3_S2) {
    Ho = (a.H - 2) / 2 + 1;  // pad (0,1,0,1) then 3x3 stride 2, pad 0
    Wo = (a.W - 2) / 2 + 1;
  }
  if (mode == EOVAE_CONV_CUSTOM) {
    EOVAE_CHECK(taps != nullptr && taps->num_taps >= 1 && taps->num_taps <= igemm::MAX_TAPS && extra == nullptr && gnp == nullptr,
                "igemm: bad custom tap geometry");
    p.num_taps = taps->num_taps;
    if (taps->parity_in) {   // A = high-resolution tensor read through its four parity sub-lattices; output = low resolution
      EOVAE_CHECK(a.H % 2 == 0 && a.W % 2 == 0, "igemm: parity sub-lattices need even extents");
      Ho = a.H / 2;
      Wo = a.W / 2;
    }
    p.phases = taps->phases;
    p.b_phase_rows = taps->b_phase_rows;
    for (int q = 0; q < taps->phases; ++q) {
      p.phase_dx[q] = taps->phase_dx[q];
      p.phase_dy[q] = taps->phase_dy[q];
    }
  }
  // --- M tile = one TMA box of pixels
  m_tiling(a.N, Ho, Wo, w_batches > 1, &p.box_w, &p.box_h, &p.box_n);
  p.tiles_w = ceil_div(Wo, p.box_w);
  p.tiles_h = ceil_div(Ho, p.box_h);
  p.tiles_n = ceil_div(a.N, p.box_n);
  p.chunks_per_tap = k_per_tap / ch;
  p.k_per_tap = k_per_tap;
  int extra_k = 0;
  if (extra != nullptr) {
    EOVAE_CHECK(mode != EOVAE_CONV_3X3_S2, "igemm: a fused 1x1 operand needs a stride-1 convolution");
    EOVAE_CHECK(extra->N == a.N && extra->H == a.H && extra->W == a.W, "igemm: fused 1x1 operand shape mismatch");
    EOVAE_CHECK(extra->C % ch == 0 && extra->pix_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(extra->ptr) % 16) == 0,
                "igemm: fused 1x1 operand must have a multiple of %d channels and 16-byte aligned pixels", ch);
    p.extra_chunks = extra->C / ch;
    p.extra_map = 1;
    extra_k = extra->C;
  }
  p.Wo = Wo;
  p.Ho = Ho;
  p.Nimg = a.N;
  p.Cout = cout;
  p.b_batched = w_batches > 1;
  p.out = out;
  p.out_dtype = out_dtype;
  p.out_pix_stride = out_pix_stride;
  p.res = res;
  p.res_dtype = res_dtype;
  p.res_pix_stride = res_pix_stride;
  p.bias = bias;
  p.out_scale = scale;
  p.debug_mode = g_debug_mode;
  const int block_n = pick_block_n(round_up(cout, 16));
  p.n_tiles = ceil_div(cout, block_n);
  // CTA pairs (cta_group::2, UMMA M = 256) whenever there are at least two m-tiles to pair; a batched B operand
  // additionally needs both tiles of a pair inside one image.
  const int m_tiles_total = p.tiles_w * p.tiles_h * p.tiles_n;
  int ctas = (m_tiles_total >= 2) ? 2 : 1;
  if (w_batches > 1 && (p.tiles_w * p.tiles_h) % 2 != 0) ctas = 1;
  if (g_force_ctas == 1) ctas = 1;
  if (g_force_ctas == 2 && !(w_batches > 1 && (p.tiles_w * p.tiles_h) % 2 != 0)) ctas = 2;
  // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A/B = bf16|f16, K-major both, N>>3, M>>4
  const uint32_t fmt = act_dtype == EOVAE_BF16 ? 1u : 0u, fmt_b = w_dtype == EOVAE_BF16 ? 1u : 0u;
  p.idesc = (1u << 4) | (fmt << 7) | (fmt_b << 10) | (static_cast<uint32_t>(block_n >> 3) << 17) |
            (static_cast<uint32_t>((igemm::BLOCK_M * ctas) >> 4) << 24);

  // --- A maps
  const uint32_t box[4] = {static_cast<uint32_t>(ch), static_cast<uint32_t>(p.box_w), static_cast<uint32_t>(p.box_h),
                           static_cast<uint32_t>(p.box_n)};
  const uint64_t es = 2;
  if (mode == EOVAE_CONV_3X3_S2 || (mode == EOVAE_CONV_CUSTOM && taps->parity_in)) {
    // four parity sub-lattices x[:, ph::2, pw::2, :]; tap (kh,kw) -> lattice (kh&1, kw&1), offset (kh>>1, kw>>1)
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw) {
        uint64_t dims[4] = {static_cast<uint64_t>(a.C), static_cast<uint64_t>((a.W - pw + 1) / 2),
                            static_cast<uint64_t>((a.H - ph + 1) / 2), static_cast<uint64_t>(a.N)};
        uint64_t strides[3] = {2 * a.pix_stride * es, 2 * static_cast<uint64_t>(a.W) * a.pix_stride * es,
                               static_cast<uint64_t>(a.H) * a.W * a.pix_stride * es};
        const uint8_t* base = reinterpret_cast<const uint8_t*>(a.ptr) + (static_cast<uint64_t>(ph) * a.W + pw) * a.pix_stride * es;
        if (dims[1] == 0 || dims[2] == 0) { dims[1] = dims[1] ? dims[1] : 1; dims[2] = dims[2] ? dims[2] : 1; }
        int rc = encode_map(&p.a_map[ph * 2 + pw], act_dtype, 4, base, dims, strides, box, chunk_bytes);
        if (rc) return rc;
      }
    if (mode == EOVAE_CONV_CUSTOM) {
      for (int t = 0; t < p.num_taps; ++t) {
        p.tap_map[t] = taps->tap_map[t];
        p.tap_dy[t] = taps->tap_dy[t];
        p.tap_dx[t] = taps->tap_dx[t];
      }
    } else {
      for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw) {
          p.tap_map[kh * 3 + kw] = (kh & 1) * 2 + (kw & 1);
          p.tap_dy[kh * 3 + kw] = kh >> 1;
          p.tap_dx[kh * 3 + kw] = kw >> 1;
        }
    }
  } else {
    uint64_t dims[4] = {static_cast<uint64_t>(a.C), static_cast<uint64_t>(a.W), static_cast<uint64_t>(a.H),
                        static_cast<uint64_t>(a.N)};
    uint64_t strides[3] = {static_cast<uint64_t>(a.pix_stride) * es, static_cast<uint64_t>(a.W) * a.pix_stride * es,
                           static_cast<uint64_t>(a.H) * a.W * a.pix_stride * es};
    int rc = encode_map(&p.a_map[0], act_dtype, 4, a.ptr, dims, strides, box, chunk_bytes);
    if (rc) return rc;
    p.a_map[1] = p.a_map[2] = p.a_map[3] = p.a_map[0];
    if (extra != nullptr) {
      uint64_t edims[4] = {static_cast<uint64_t>(extra->C), static_cast<uint64_t>(extra->W),
                           static_cast<uint64_t>(extra->H), static_cast<uint64_t>(extra->N)};
      uint64_t estr[3] = {static_cast<uint64_t>(extra->pix_stride) * es,
                          static_cast<uint64_t>(extra->W) * extra->pix_stride * es,
                          static_cast<uint64_t>(extra->H) * extra->W * extra->pix_stride * es};
      rc = encode_map(&p.a_map[1], act_dtype, 4, extra->ptr, edims, estr, box, chunk_bytes);
      if (rc) return rc;
    }
    for (int t = 0; t < p.num_taps; ++t) {
      p.tap_map[t] = 0;
      if (mode == EOVAE_CONV_CUSTOM) {
        p.tap_dy[t] = taps->tap_dy[t];
        p.tap_dx[t] = taps->tap_dx[t];
      } else {
        p.tap_dy[t] = (mode == EOVAE_CONV_1X1) ? 0 : t / 3 - 1;
        p.tap_dx[t] = (mode == EOVAE_CONV_1X1) ? 0 : t % 3 - 1;
      }
    }
  }
  // --- B map: [batch][rows][K] K-major
  {
    const uint64_t ktot = static_cast<uint64_t>(p.num_taps) * k_per_tap + extra_k;
    uint64_t dims[3] = {ktot, static_cast<uint64_t>(w_rows), static_cast<uint64_t>(w_batches)};
    uint64_t strides[2] = {static_cast<uint64_t>(w_row_stride) * es, static_cast<uint64_t>(w_batch_stride) * es};
    if (w_batches <= 1) strides[1] = static_cast<uint64_t>(w_row_stride) * static_cast<uint64_t>(w_rows) * es;
    const uint32_t bbox[3] = {static_cast<uint32_t>(ch), static_cast<uint32_t>(block_n / ctas), 1};
    int rc = encode_map(&p.b_map, w_dtype, 3, w, dims, strides, bbox, chunk_bytes);
    if (rc) return rc;
  }
  // --- output map for the TMA-store epilogue: each epilogue warp owns 32 consecutive tile rows; they must form a
  //     rectangular (w, h, image) sub-box of the tile
  if (out_dtype != EOVAE_F32 && block_n >= 32 && !g_no_tma_store) {
    int sw = 0, sh = 0, sn = 0;
    if (p.box_w >= 32) {
      if (p.box_w % 32 == 0) { sw = 32; sh = 1; sn = 1; }
    } else if (32 % p.box_w == 0) {
      const int rows = 32 / p.box_w;
      if (p.box_h >= rows) {
        if (p.box_h % rows == 0) { sw = p.box_w; sh = rows; sn = 1; }
      } else if (rows % p.box_h == 0 && p.box_n >= rows / p.box_h && p.box_n % (rows / p.box_h) == 0) {
        sw = p.box_w; sh = p.box_h; sn = rows / p.box_h;
      }
    }
    if ((p.box_w * p.box_h * p.box_n) % 32 != 0) sw = 0;  // every warp entirely inside or outside the pixel box
    if (sw > 0) {
      uint64_t odims[4] = {static_cast<uint64_t>(cout), static_cast<uint64_t>(Wo), static_cast<uint64_t>(Ho),
                           static_cast<uint64_t>(a.N)};
      // BLOCK_N = 128: every epilogue warp owns 64 columns -> one 128-byte-row box per warp and tile
      const bool wide = block_n == 128 && cout % 64 == 0 && !g_no_wide_store;
      const uint32_t obox[4] = {wide ? 64u : 32u, static_cast<uint32_t>(sw), static_cast<uint32_t>(sh), static_cast<uint32_t>(sn)};
      const int osw = wide ? 128 : 64;
      if (p.phases == 1) {
        uint64_t ostr[3] = {static_cast<uint64_t>(out_pix_stride) * es, static_cast<uint64_t>(Wo) * out_pix_stride * es,
                            static_cast<uint64_t>(Ho) * Wo * out_pix_stride * es};
        int rc0 = encode_map(&p.out_map, out_dtype, 4, out, odims, ostr, obox, osw, false);
        if (rc0) return rc0;
      } else {
        // phase (py, px) stores the parity sub-lattice out[:, py::2, px::2, :] of the (2 Ho) x (2 Wo) output
        const uint64_t W2 = 2ull * Wo, H2 = 2ull * Ho;
        uint64_t ostr[3] = {2ull * out_pix_stride * es, 2ull * W2 * out_pix_stride * es, H2 * W2 * out_pix_stride * es};
        for (int q = 0; q < 4; ++q) {
          uint8_t* base = reinterpret_cast<uint8_t*>(out) + (static_cast<uint64_t>(q >> 1) * W2 + (q & 1)) * out_pix_stride * es;
          int rc0 = encode_map(q == 0 ? &p.out_map : &p.out_map_ph[q - 1], out_dtype, 4, base, odims, ostr, obox, osw, false);
          if (rc0) return rc0;
        }
      }
      p.out_tma = wide ? 2 : 1;
      // residual of the same geometry: pulled by TMA into the store staging buffer (coalesced) instead of per-thread rows
      if (wide && res != nullptr && res_dtype != EOVAE_F32 && p.phases == 1 && !g_no_res_tma) {
        uint64_t rstr[3] = {static_cast<uint64_t>(res_pix_stride) * es, static_cast<uint64_t>(Wo) * res_pix_stride * es,
                            static_cast<uint64_t>(Ho) * Wo * res_pix_stride * es};
        int rc0 = encode_map(&p.res_map, res_dtype, 4, res, odims, rstr, obox, 128, false);
        if (rc0) return rc0;
        p.res_tma = 1;
      }
    }
  }
  EOVAE_CHECK(p.phases == 1 || (p.out_tma != 0 && res == nullptr && p.box_n == 1),
              "igemm: the sub-pixel upsample convolution needs the TMA-store epilogue (16-bit output, Cout >= 32, whole warps "
              "inside the pixel box), one image per tile and no residual");
  const int total_tiles = ceil_div(m_tiles_total, ctas) * p.n_tiles * p.phases;  // work items
  p.fd_phases.set(p.phases);
  p.fd_ntiles.set(p.n_tiles);
  p.fd_tw.set(p.tiles_w);
  p.fd_th.set(p.tiles_h);
  if (total_tiles == 0) return 0;
  if (gn_stats != nullptr) {
    EOVAE_CHECK(p.box_n == 1 && cout % 32 == 0 && gn_groups > 0 && cout % gn_groups == 0 && 32 % (cout / gn_groups) == 0 &&
                    block_n >= 32,
                "igemm: fused GroupNorm statistics unsupported for this shape (query eovae_conv2d_gn_workspace_bytes)");
    const size_t region = 2 * 4 * static_cast<size_t>(p.tiles_w) * p.tiles_h * p.tiles_n * gn_groups;
    const size_t need = sizeof(float) * region * p.phases;
    EOVAE_CHECK(gn_ws != nullptr && gn_ws_bytes >= need, "igemm: GroupNorm workspace too small");
    p.gn_partial = static_cast<float*>(gn_ws);
    p.gn_phase_stride = static_cast<long long>(region);
    p.gn_groups = gn_groups;
    p.gn_cpg = cout / gn_groups;
    p.gn_cpg_log2 = 0;
    while ((1 << p.gn_cpg_log2) < p.gn_cpg) ++p.gn_cpg_log2;
  }
  int rc;
  // two K-chunks (128 channels) per pipeline stage where the shapes allow >= 3 stages of shared memory
  const int k_items = p.num_taps * p.chunks_per_tap + p.extra_chunks;
  const bool kch2 = chunk_bytes == 128 && k_items % 2 == 0 && !g_force_kch1 &&
                    (block_n == 128 || (block_n == 256 && ctas == 2));
  // narrow-channel inputs (the dynamic input conv: 16 channels = one 32-byte chunk per tap): three taps per stage
  const bool kch3 = chunk_bytes == 32 && k_items % 3 == 0 && !g_force_kch1 && block_n == 128 && ctas == 2;
  // ... all nine taps of a 3x3 kernel in ONE stage (one producer / MMA handshake per tile; 3 stages of 54 KB)
  const bool kch9 = kch3 && k_items == 9 && !g_no_kch9;
  // halo reuse of the A tile across the three horizontal taps: m-tile = 128 consecutive pixels of one image row
  const bool halo = mode == EOVAE_CONV_3X3 && chunk_bytes == 128 && ctas == 2 && p.extra_chunks == 0 && p.box_w == 128 &&
                    p.box_h == 1 && p.box_n == 1 && (block_n == 16 || block_n == 128 || block_n == 256) && !g_no_halo;
  EOVAE_CHECK(gnp == nullptr || (halo && block_n != 16 && a.C % 64 == 0 && a.C % gnp->groups == 0 && a.C * 8 <= igemm::GNP_COEF_BYTES),
              "igemm: fused GroupNorm prologue unsupported for this shape (query eovae_conv2d_gn_prologue_ok)");
  if (gnp != nullptr) {
    const int total = a.N * a.C;
    gn_ab_kernel<<<ceil_div(total, 256), 256, 0, stream>>>(gnp->stats, gnp->gamma, gnp->beta,
                                                           static_cast<float2*>(gnp->workspace), a.C, gnp->groups, total,
                                                           act_dtype == EOVAE_F16 ? 1 : 0);
    EOVAE_LAUNCH_CHECK();
    p.gnp_ab = static_cast<const float2*>(gnp->workspace);
    p.gnp_cin = a.C;
    p.gnp_h = a.H;
    p.gnp_w = a.W;
    p.gnp_bf16 = act_dtype == EOVAE_BF16;
  }
  if (halo) {
    uint64_t dims[4] = {static_cast<uint64_t>(a.C), static_cast<uint64_t>(a.W), static_cast<uint64_t>(a.H),
                        static_cast<uint64_t>(a.N)};
    uint64_t strides[3] = {static_cast<uint64_t>(a.pix_stride) * es, static_cast<uint64_t>(a.W) * a.pix_stride * es,
                           static_cast<uint64_t>(a.H) * a.W * a.pix_stride * es};
    const uint32_t hbox[4] = {static_cast<uint32_t>(ch), static_cast<uint32_t>(igemm::HALO_ROWS), 1u, 1u};
    rc = encode_map(&p.a_map[3], act_dtype, 4, a.ptr, dims, strides, hbox, chunk_bytes);
    if (rc) return rc;
    if (gnp != nullptr && block_n == 128) rc = launch_t<128, 128, 2, 1, true, true>(p, total_tiles, stream);
    else if (gnp != nullptr) rc = launch_t<256, 128, 2, 1, true, true>(p, total_tiles, stream);
    else if (block_n == 128) rc = launch_t<128, 128, 2, 1, true>(p, total_tiles, stream);
    else if (block_n == 16) rc = launch_t<16, 128, 2, 1, true>(p, total_tiles, stream);  // skinny-N: dynamic output conv
    else rc = launch_t<256, 128, 2, 1, true>(p, total_tiles, stream);
  } else if (kch9) {
    rc = launch_t<128, 32, 2, 9>(p, total_tiles, stream);
  } else if (kch3) {
    rc = launch_t<128, 32, 2, 3>(p, total_tiles, stream);
  } else if (kch2) {
    if (block_n == 128 && ctas == 2) rc = launch_t<128, 128, 2, 2>(p, total_tiles, stream);
    else if (block_n == 128) rc = launch_t<128, 128, 1, 2>(p, total_tiles, stream);
    else rc = launch_t<256, 128, 2, 2>(p, total_tiles, stream);
  } else if (ctas == 2) {
    switch (chunk_bytes) {
      case 128: rc = launch_n<128, 2>(block_n, p, total_tiles, stream); break;
      case 64: rc = launch_n<64, 2>(block_n, p, total_tiles, stream); break;
      default: rc = launch_n<32, 2>(block_n, p, total_tiles, stream); break;
    }
  } else {
    switch (chunk_bytes) {
      case 128: rc = launch_n<128, 1>(block_n, p, total_tiles, stream); break;
      case 64: rc = launch_n<64, 1>(block_n, p, total_tiles, stream); break;
      default: rc = launch_n<32, 1>(block_n, p, total_tiles, stream); break;
    }
  }
  if (rc != 0 || gn_stats == nullptr) return rc;
  {
    const int slots = p.tiles_w * p.tiles_h * 4;
    const int gsub = gn_groups % 4 == 0 ? 4 : (gn_groups % 2 == 0 ? 2 : 1);
    int threads = 256;
    while (threads > 32 && threads / gsub >= 2 * slots) threads >>= 1;  // few slots: fewer idle lanes
    gn_tiles_finalize_kernel<<<dim3(gn_groups / gsub, p.Nimg), threads, 0, stream>>>(
        p.gn_partial, gn_stats, gn_groups, gsub, slots, static_cast<double>(Ho) * Wo * p.gn_cpg * p.phases, gn_eps, p.phases,
        p.gn_phase_stride);
  }
  EOVAE_LAUNCH_CHECK();
  return 0;
}

// shape rule shared by launch_igemm and the workspace query
void m_tiling(int n, int ho, int wo, bool batched, int* bw, int* bh, int* bn) {
  *bw = wo < 128 ? wo : 128;
  *bh = 128 / *bw;
  if (*bh > ho) *bh = ho;
  *bn = 128 / (*bw * *bh);
  if (*bn > n) *bn = n;
  if (batched || *bn < 1) *bn = 1;
}

}  // namespace

extern "C" {

void eovae_set_debug_mode(int mode) {
  g_debug_mode = mode & 0xFF;       // low byte: pipeline actor switched off (igemm_sm100.cuh)
  g_force_ctas = (mode >> 8) & 3;   // bits 8-9: force 1- or 2-CTA groups (0 = automatic)
  g_force_kch1 = (mode >> 10) & 1;  // bit 10: force 64-channel pipeline stages
  g_no_tma_store = (mode >> 11) & 1;  // bit 11: disable the TMA-store epilogue
  g_no_halo = (mode >> 12) & 1;       // bit 12: disable the halo-reuse mainloop
  g_no_kch9 = (mode >> 13) & 1;       // bit 13: three (not nine) taps per stage in the narrow-channel 3x3 conv
  g_no_wide_store = (mode >> 14) & 1; // bit 14: 64-byte-row output boxes in the BLOCK_N = 128 kernels too
  g_no_res_tma = (mode >> 15) & 1;    // bit 15: residual through per-thread loads even where the TMA path applies
}

int eovae_conv_chunk_bytes(int cin) {
  if (cin % 64 == 0) return 128;
  if (cin % 32 == 0) return 64;
  return 32;
}

int eovae_conv_k_per_tap(int cin) {
  const int ch = eovae_conv_chunk_bytes(cin) / 2;
  return round_up(cin, ch);
}

size_t eovae_conv2d_gn_workspace_bytes(int n, int h, int w, int mode, int cout, int groups) {
  int ho = h, wo = w;
  if (mode == EOVAE_CONV_3X3_S2) {
    ho = (h - 2) / 2 + 1;
    wo = (w - 2) / 2 + 1;
  }
  int bw, bh, bn;
  m_tiling(n, ho, wo, false, &bw, &bh, &bn);
  if (bn != 1 || groups <= 0 || cout % 32 != 0 || cout % groups != 0 || 32 % (cout / groups) != 0) return 0;
  return sizeof(float) * 2 * 4 * static_cast<size_t>(ceil_div(wo, bw)) * ceil_div(ho, bh) * n * groups;
}

int eovae_conv2d_gn_prologue_ok(int n, int h, int w, int cin, int cout, int mode, int groups) {
  if (mode != EOVAE_CONV_3X3 || g_no_halo || g_force_ctas == 1 || groups <= 0) return 0;
  if (cin % 64 != 0 || cin % groups != 0) return 0;
  int bw, bh, bn;
  m_tiling(n, h, w, false, &bw, &bh, &bn);
  if (bw != 128 || bh != 1 || bn != 1) return 0;
  const int block_n = pick_block_n(round_up(cout, 16));
  if (block_n != 128 && block_n != 256) return 0;
  return ceil_div(w, bw) * h * n >= 2 ? 1 : 0;
}

int eovae_conv2d(const void* x, int n, int h, int w, int cin, long long x_pix_stride, int mode, const void* w_packed,
                 int cout, const float* bias, const void* residual, int res_dtype, long long res_pix_stride, void* out,
                 int out_dtype, long long out_pix_stride, int act_dtype, float scale, float* gn_stats, int gn_groups,
                 float gn_eps, void* gn_workspace, size_t gn_workspace_bytes, const void* x2, int cin2,
                 long long x2_pix_stride, const float* in_gn_stats, const float* in_gn_gamma, const float* in_gn_beta,
                 int in_gn_groups, void* in_gn_workspace, size_t in_gn_workspace_bytes, void* stream) {
  EOVAE_CHECK(mode == EOVAE_CONV_3X3 || mode == EOVAE_CONV_1X1 || mode == EOVAE_CONV_3X3_S2, "conv2d: bad mode %d", mode);
  EOVAE_CHECK(n > 0 && h > 0 && w > 0 && cin > 0 && cout > 0, "conv2d: empty shape");
  if (act_dtype == EOVAE_F32) {
    // fp32 validation path (fp32_path.cu): fp32 NHWC activations, fp32 [cout][tap][k_per_tap] weights, SIMT FMAs
    EOVAE_CHECK(out_dtype == EOVAE_F32 && (residual == nullptr || res_dtype == EOVAE_F32),
                "conv2d: the fp32 path needs fp32 output and residual");
    EOVAE_CHECK(x2 == nullptr && in_gn_stats == nullptr && gn_stats == nullptr,
                "conv2d: fused shortcut / GroupNorm prologue / statistics epilogue are tensor-core-path features");
    EOVAE_CHECK(cin % 4 == 0, "conv2d: Cin (%d) must be a multiple of 4 on the fp32 path", cin);
    const int kpt32 = eovae_conv_k_per_tap(cin);
    const int taps32 = mode == EOVAE_CONV_1X1 ? 1 : 9;
    return eovae::f32::conv2d(static_cast<const float*>(x), n, h, w, cin, x_pix_stride, mode, static_cast<const float*>(w_packed),
                              static_cast<long long>(taps32) * kpt32, 1, 0, kpt32, cout, bias, static_cast<const float*>(residual),
                              res_pix_stride, static_cast<float*>(out), out_pix_stride, scale, static_cast<cudaStream_t>(stream));
  }
  EOVAE_CHECK(cin % 8 == 0, "conv2d: Cin (%d) must be a multiple of 8", cin);
  ASpec a{x, n, h, w, cin, x_pix_stride};
  ASpec e{x2, n, h, w, cin2, x2_pix_stride};
  const int cb = eovae_conv_chunk_bytes(cin);
  const int kpt = eovae_conv_k_per_tap(cin);
  const int taps = mode == EOVAE_CONV_1X1 ? 1 : 9;
  if (x2 != nullptr)
    EOVAE_CHECK(eovae_conv_chunk_bytes(cin2) == cb && eovae_conv_k_per_tap(cin2) == cin2,
                "conv2d: fused 1x1 operand needs Cin2 (%d) compatible with the %d-byte K chunks of Cin (%d)", cin2, cb, cin);
  GnPrologue gp{in_gn_stats, in_gn_gamma, in_gn_beta, in_gn_groups, in_gn_workspace};
  if (in_gn_stats != nullptr) {
    EOVAE_CHECK(x2 == nullptr, "conv2d: the fused GroupNorm prologue cannot be combined with a fused 1x1 operand");
    EOVAE_CHECK(in_gn_gamma != nullptr && in_gn_beta != nullptr && in_gn_workspace != nullptr &&
                    in_gn_workspace_bytes >= sizeof(float) * 2 * static_cast<size_t>(n) * cin,
                "conv2d: GroupNorm prologue needs gamma, beta and a workspace of N*Cin*8 bytes");
    EOVAE_CHECK(x_pix_stride == cin, "conv2d: GroupNorm prologue needs a dense NHWC input");
  }
  return launch_igemm(a, mode, w_packed, kpt, cb, cout, round_up(cout, 16),
                      static_cast<long long>(taps) * kpt + (x2 != nullptr ? cin2 : 0), 0, 1, bias, residual, res_dtype,
                      res_pix_stride, out, out_dtype, out_pix_stride, act_dtype, scale, static_cast<cudaStream_t>(stream),
                      gn_stats, gn_groups, gn_eps, gn_workspace, gn_workspace_bytes, x2 != nullptr ? &e : nullptr,
                      in_gn_stats != nullptr ? &gp : nullptr);
}

int eovae_gemm_tn_batched(const void* a, long long lda, long long a_batch_stride, const void* b, long long ldb,
                          long long b_batch_stride, void* c, int c_dtype, long long ldc, int batch, int m, int n, int k,
                          int a_dtype, int b_dtype, float scale, void* stream) {
  EOVAE_CHECK(batch > 0 && m > 0 && n > 0 && k > 0, "gemm_tn_batched: empty shape");
  if (a_dtype == EOVAE_F32) {  // fp32 validation path: a 1x1 "conv" over [batch][1][m] pixels with a per-batch B operand
    EOVAE_CHECK(b_dtype == EOVAE_F32 && c_dtype == EOVAE_F32 && k % 16 == 0 && lda % 4 == 0,
                "gemm_tn_batched: the fp32 path needs fp32 A, B, C, K a multiple of 16 and lda a multiple of 4");
    EOVAE_CHECK(a_batch_stride == static_cast<long long>(m) * lda, "gemm_tn_batched: A batches must be contiguous");
    return eovae::f32::conv2d(static_cast<const float*>(a), batch, 1, m, k, lda, EOVAE_CONV_1X1, static_cast<const float*>(b), ldb, 1,
                              batch == 1 ? 0 : b_batch_stride, k, n, nullptr, nullptr, 0, static_cast<float*>(c), ldc, scale,
                              static_cast<cudaStream_t>(stream));
  }
  EOVAE_CHECK(k % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0, "gemm_tn_batched: K, lda, ldb must be multiples of 8");
  EOVAE_CHECK(a_batch_stride == static_cast<long long>(m) * lda, "gemm_tn_batched: A batches must be contiguous");
  EOVAE_CHECK(ldb >= k, "gemm_tn_batched: ldb < K");
  ASpec as{a, batch, 1, m, k, lda};
  const int cb = eovae_conv_chunk_bytes(k);
  const int kpt = eovae_conv_k_per_tap(k);
  EOVAE_CHECK(kpt == k, "gemm_tn_batched: K (%d) must be a multiple of 16", k);
  return launch_igemm(as, EOVAE_CONV_1X1, b, kpt, cb, n, n, ldb, b_batch_stride, batch == 1 ? 1 : batch, nullptr, nullptr, 0, 0, c,
                      c_dtype, ldc, a_dtype, scale, static_cast<cudaStream_t>(stream), nullptr, 0, 0.f, nullptr, 0, nullptr, nullptr,
                      b_dtype);
}

// ---------------------------------------------------------------------------------------------------------------------
// Upsample (layers.py:40-50: nearest x2, then 3x3 conv) in sub-pixel form.  Output pixel (2i + py, 2j + px) only sees the
// low-resolution pixels (i + a - 1 + py, j + b - 1 + px), a, b in {0, 1}, through the SUMS of the 3x3 taps that fall on the
// same source pixel:  rows  py = 0: a = 0 <- kh 0, a = 1 <- kh 1 + 2;   py = 1: a = 0 <- kh 0 + 1, a = 1 <- kh 2  (columns alike).
// Four 2x2 convolutions on the low-resolution tensor replace one 3x3 convolution on the 4x larger one: 16 instead of 36
// MACs per output pixel pair, and the upsampled tensor is never materialised.
}  // extern "C"

namespace {

__device__ __forceinline__ bool up2x_in_set(int parity, int a, int k) {  // does 3x3 index k fold onto 2x2 index a?
  return parity == 0 ? (a == 0 ? k == 0 : k >= 1) : (a == 0 ? k <= 1 : k == 2);
}

// forward operand: out[q][o][t][c] (q = py*2+px, t = a*2+b; rows padded to rows_pad, channels to kpt), fp32 sums rounded once
template <typename T>
__global__ void pack_up2x_kernel(const float* __restrict__ w, T* __restrict__ out, int cout, int cin, int rows_pad, int kpt,
                                 long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = static_cast<int>(i % kpt);
  long long r = i / kpt;
  const int t = static_cast<int>(r % 4); r /= 4;
  const int o = static_cast<int>(r % rows_pad);
  const int q = static_cast<int>(r / rows_pad);
  float v = 0.f;
  if (o < cout && c < cin) {
    const float* wp = w + (static_cast<long long>(o) * cin + c) * 9;
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw)
        if (up2x_in_set(q >> 1, t >> 1, kh) && up2x_in_set(q & 1, t & 1, kw)) v += wp[kh * 3 + kw];
  }
  out[i] = T16<T>::from_f(v);
}

// data-gradient operand: out[ci][q*4 + t][co] = W'_q[t][co][ci] (rows padded to round_up(cin, 16), co to kpt)
template <typename T>
__global__ void pack_up2x_dgrad_kernel(const float* __restrict__ w, T* __restrict__ out, int cout, int cin, int kpt,
                                       long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int co = static_cast<int>(i % kpt);
  long long r = i / kpt;
  const int tap = static_cast<int>(r % 16);
  const int ci = static_cast<int>(r / 16);
  const int q = tap >> 2, t = tap & 3;
  float v = 0.f;
  if (co < cout && ci < cin) {
    const float* wp = w + (static_cast<long long>(co) * cin + ci) * 9;
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw)
        if (up2x_in_set(q >> 1, t >> 1, kh) && up2x_in_set(q & 1, t & 1, kw)) v += wp[kh * 3 + kw];
  }
  out[i] = T16<T>::from_f(v);
}

// data-gradient operand of the Downsample conv (pad (0,1,0,1) + 3x3 stride 2, layers.py:33-37) in sub-pixel form: input pixel
// (2i + pu, 2j + pv) receives dy[i - a, j - b] through W[kh, kw]^T with kh = 2a for pu = 0 (a in {0, 1}) and kh = 1 (a = 0 only)
// for pu = 1; columns alike.  out[q = pu*2+pv][ci][t = a*2+b][co], zero where the phase has no such tap.
template <typename T>
__global__ void pack_s2_dgrad_kernel(const float* __restrict__ w, T* __restrict__ out, int cout, int cin, int rows_pad, int kpt,
                                     long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int co = static_cast<int>(i % kpt);
  long long r = i / kpt;
  const int t = static_cast<int>(r % 4); r /= 4;
  const int ci = static_cast<int>(r % rows_pad);
  const int q = static_cast<int>(r / rows_pad);
  const int pu = q >> 1, pv = q & 1, a = t >> 1, b = t & 1;
  const int kh = pu == 0 ? 2 * a : (a == 0 ? 1 : -1), kw = pv == 0 ? 2 * b : (b == 0 ? 1 : -1);
  float v = 0.f;
  if (ci < cin && co < cout && kh >= 0 && kw >= 0) v = w[(static_cast<long long>(co) * cin + ci) * 9 + kh * 3 + kw];
  out[i] = T16<T>::from_f(v);
}

}  // namespace

extern "C" {

int eovae_conv2d_up2x_ok(int n, int h, int w, int cin, int cout) {
  // what the phase epilogue needs (see launch_igemm): 128-byte K chunks, whole 32-pixel warps inside one image's tile
  if (cin % 64 != 0 || cout % 32 != 0 || n < 1) return 0;
  int bw, bh, bn;
  m_tiling(n, h, w, false, &bw, &bh, &bn);
  if (bn != 1 || (bw * bh) % 32 != 0) return 0;
  if (bw >= 32) return bw % 32 == 0 ? 1 : 0;
  return (32 % bw == 0 && bh % (32 / bw) == 0) ? 1 : 0;
}

size_t eovae_conv2d_up2x_gn_workspace_bytes(int n, int h, int w, int cout, int groups) {
  return 4 * eovae_conv2d_gn_workspace_bytes(n, h, w, EOVAE_CONV_1X1, cout, groups);
}

int eovae_pack_conv_weight_up2x(const float* w_oihw, void* out, int cout, int cin, int dtype, int dgrad, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  EOVAE_CHECK(dtype == EOVAE_BF16 || dtype == EOVAE_F16, "pack_conv_weight_up2x: 16-bit operands only");
  if (dgrad == 2) {  // data-gradient operand of the STRIDE-2 conv: [4 phases][round_up(cin,16)][4 taps][k_per_tap(cout)]
    const int kpt = eovae_conv_k_per_tap(round_up(cout, 8)), rows_pad = round_up(cin, 16);
    const long long total = 4LL * rows_pad * 4 * kpt;
    const unsigned grid = static_cast<unsigned>((total + 255) / 256);
    if (dtype == EOVAE_BF16) pack_s2_dgrad_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(w_oihw, static_cast<__nv_bfloat16*>(out), cout, cin, rows_pad, kpt, total);
    else pack_s2_dgrad_kernel<__half><<<grid, 256, 0, stream>>>(w_oihw, static_cast<__half*>(out), cout, cin, rows_pad, kpt, total);
    EOVAE_LAUNCH_CHECK();
    return 0;
  }
  if (!dgrad) {
    const int kpt = eovae_conv_k_per_tap(cin), rows_pad = round_up(cout, 16);
    const long long total = 4LL * rows_pad * 4 * kpt;
    const unsigned grid = static_cast<unsigned>((total + 255) / 256);
    if (dtype == EOVAE_BF16) pack_up2x_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(w_oihw, static_cast<__nv_bfloat16*>(out), cout, cin, rows_pad, kpt, total);
    else pack_up2x_kernel<__half><<<grid, 256, 0, stream>>>(w_oihw, static_cast<__half*>(out), cout, cin, rows_pad, kpt, total);
  } else {
    const int kpt = eovae_conv_k_per_tap(round_up(cout, 8));
    const long long total = static_cast<long long>(round_up(cin, 16)) * 16 * kpt;
    const unsigned grid = static_cast<unsigned>((total + 255) / 256);
    if (dtype == EOVAE_BF16) pack_up2x_dgrad_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(w_oihw, static_cast<__nv_bfloat16*>(out), cout, cin, kpt, total);
    else pack_up2x_dgrad_kernel<__half><<<grid, 256, 0, stream>>>(w_oihw, static_cast<__half*>(out), cout, cin, kpt, total);
  }
  EOVAE_LAUNCH_CHECK();
  return 0;
}

int eovae_conv2d_up2x(const void* x, int n, int h, int w, int cin, long long x_pix_stride, const void* w_packed, int cout,
                      const float* bias, void* out, int out_dtype, long long out_pix_stride, int act_dtype, float* gn_stats,
                      int gn_groups, float gn_eps, void* gn_workspace, size_t gn_workspace_bytes, void* stream) {
  EOVAE_CHECK(eovae_conv2d_up2x_ok(n, h, w, cin, cout), "conv2d_up2x: unsupported shape (query eovae_conv2d_up2x_ok)");
  EOVAE_CHECK(out_dtype == EOVAE_BF16 || out_dtype == EOVAE_F16, "conv2d_up2x: 16-bit output only");
  TapSpec ts;
  memset(&ts, 0, sizeof(ts));
  ts.num_taps = 4;
  for (int t = 0; t < 4; ++t) {
    ts.tap_dy[t] = (t >> 1) - 1;
    ts.tap_dx[t] = (t & 1) - 1;
  }
  ts.parity_in = false;
  ts.phases = 4;
  for (int q = 0; q < 4; ++q) {
    ts.phase_dy[q] = q >> 1;
    ts.phase_dx[q] = q & 1;
  }
  ts.b_phase_rows = round_up(cout, 16);
  ASpec a{x, n, h, w, cin, x_pix_stride};
  const int kpt = eovae_conv_k_per_tap(cin);
  return launch_igemm(a, EOVAE_CONV_CUSTOM, w_packed, kpt, eovae_conv_chunk_bytes(cin), cout, 4LL * round_up(cout, 16), 4LL * kpt, 0, 1,
                      bias, nullptr, 0, 0, out, out_dtype, out_pix_stride, act_dtype, 1.0f, static_cast<cudaStream_t>(stream), gn_stats,
                      gn_groups, gn_eps, gn_workspace, gn_workspace_bytes, nullptr, nullptr, -1, &ts);
}

int eovae_conv2d_s2_dgrad(const void* dy, int n, int ho, int wo, int cout, long long dy_pix_stride, const void* w_packed, int cin,
                          void* dx, int dx_dtype, long long dx_pix_stride, int act_dtype, void* stream) {
  // dx [n][2 ho][2 wo][cin]: the four input parities are four 2x2 (zero-padded) convolutions over dy, each stored on its
  // parity sub-lattice - 16 tap units per dy pixel instead of the 36 of "scatter dy into a zero-interleaved image + full 3x3"
  EOVAE_CHECK(eovae_conv2d_up2x_ok(n, ho, wo, round_up(cout, 64), cin), "conv2d_s2_dgrad: unsupported shape");
  EOVAE_CHECK(dx_dtype == EOVAE_BF16 || dx_dtype == EOVAE_F16, "conv2d_s2_dgrad: 16-bit output only");
  TapSpec ts;
  memset(&ts, 0, sizeof(ts));
  ts.num_taps = 4;
  for (int t = 0; t < 4; ++t) {
    ts.tap_dy[t] = -(t >> 1);
    ts.tap_dx[t] = -(t & 1);
  }
  ts.parity_in = false;
  ts.phases = 4;
  ts.b_phase_rows = round_up(cin, 16);
  ASpec a{dy, n, ho, wo, cout, dy_pix_stride};
  const int kpt = eovae_conv_k_per_tap(round_up(cout, 8));
  return launch_igemm(a, EOVAE_CONV_CUSTOM, w_packed, kpt, eovae_conv_chunk_bytes(cout), cin, 4LL * round_up(cin, 16), 4LL * kpt, 0, 1,
                      nullptr, nullptr, 0, 0, dx, dx_dtype, dx_pix_stride, act_dtype, 1.0f, static_cast<cudaStream_t>(stream), nullptr, 0,
                      0.f, nullptr, 0, nullptr, nullptr, -1, &ts);
}

int eovae_conv2d_up2x_dgrad(const void* dy, int n, int h2, int w2, int cout, long long dy_pix_stride, const void* w_packed, int cin,
                            void* dx, int dx_dtype, long long dx_pix_stride, int act_dtype, void* stream) {
  // dx [n][h2/2][w2/2][cin] = sum over the 4 parity sub-lattices of dy and their 2x2 taps (adjoint of eovae_conv2d_up2x)
  EOVAE_CHECK(h2 % 2 == 0 && w2 % 2 == 0 && cout % 8 == 0, "conv2d_up2x_dgrad: even extents and Cout %% 8 required");
  TapSpec ts;
  memset(&ts, 0, sizeof(ts));
  ts.num_taps = 16;
  for (int q = 0; q < 4; ++q)
    for (int t = 0; t < 4; ++t) {
      ts.tap_map[q * 4 + t] = q;
      ts.tap_dy[q * 4 + t] = -((t >> 1) - 1 + (q >> 1));
      ts.tap_dx[q * 4 + t] = -((t & 1) - 1 + (q & 1));
    }
  ts.parity_in = true;
  ts.phases = 1;
  ASpec a{dy, n, h2, w2, cout, dy_pix_stride};
  const int kpt = eovae_conv_k_per_tap(round_up(cout, 8));
  return launch_igemm(a, EOVAE_CONV_CUSTOM, w_packed, kpt, eovae_conv_chunk_bytes(cout), cin, round_up(cin, 16), 16LL * kpt, 0, 1, nullptr,
                      nullptr, 0, 0, dx, dx_dtype, dx_pix_stride, act_dtype, 1.0f, static_cast<cudaStream_t>(stream), nullptr, 0, 0.f,
                      nullptr, 0, nullptr, nullptr, -1, &ts);
}

int eovae_gemm_strided_f32(const float* a, long long lda, long long a_batch_stride, const float* b, long long b_n_stride,
                           long long b_k_stride, long long b_batch_stride, float* c, long long ldc, int batch, int m, int n, int k,
                           float scale, void* stream) {
  // fp32 validation path: c[i][r][j] = scale * sum_k a[i][r][k] * B_i(j, k), B_i(j, k) = b[i * b_batch_stride + j * b_n_stride +
  // k * b_k_stride] (any operand orientation: P V of the attention reads V in place, keys along b_k_stride)
  EOVAE_CHECK(batch > 0 && m > 0 && n > 0 && k > 0 && k % 16 == 0 && lda % 4 == 0, "gemm_strided_f32: bad shape (K %% 16, lda %% 4)");
  EOVAE_CHECK(a_batch_stride == static_cast<long long>(m) * lda, "gemm_strided_f32: A batches must be contiguous");
  return eovae::f32::conv2d(a, batch, 1, m, k, lda, EOVAE_CONV_1X1, b, b_n_stride, b_k_stride, batch == 1 ? 0 : b_batch_stride, k, n,
                            nullptr, nullptr, 0, c, ldc, scale, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
