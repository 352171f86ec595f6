// Fused flash-style single-head attention for the mid block (layers.py:128-142: softmax(q k^T / sqrt(C)) v, one head,
// d = C up to 512): scores and probabilities never touch HBM.
//
// One CTA = 128 queries of one image x one half of the head dimension (C > 256: two CTAs share a query tile, each owns
// 256 output columns and recomputes the cheap-in-context score tile; attention is < 1 % of the encoder's FLOPs and this
// keeps the fp32 O accumulator (128 x 256) + the S tile (128 x 64) inside the 512 TMEM columns).
//
//   for every 64-key block:  S = Q K^T (tcgen05, fp32 in TMEM, two S buffers so the next block's Q K^T runs under this
//   block's softmax) -> online softmax by the 4 softmax warps (thread = query row): running max m and sum l in registers,
//   P = exp((S - m) / sqrt(C)) as bf16/fp16 into a 128-byte-swizzled K-major shared-memory tile -> O += P V (V read in
//   place from the NHWC qkv tensor as an MN-major operand).  When a row's maximum moves, its O row in TMEM is rescaled
//   (tcgen05.ld -> multiply -> tcgen05.st; skipped warp-wide when no row of the warp changed).  O / l in the epilogue.
//
// warp 0: TMA producer (Q once; K chunks through a 4-stage ring)      warp 6: V producer (one tile per key block)
// warp 1: MMA issuer      warps 2-5: softmax / epilogue (TMEM lane quarter = warp & 3)
#include "../../include/eovae.h"
#include "igemm_sm100.cuh"

namespace {

using namespace igemm;

constexpr int AT_THREADS = 224;  // warps: 0 K producer, 1 MMA issuer, 2-5 softmax / epilogue, 6 V producer
constexpr int KSTAGES = 4;
constexpr int QROWS = 128;  // queries per CTA
constexpr int KB = 64;      // keys per block

struct AttnParams {
  CUtensorMap q_map;   // qkv as (3C, L, N), box (64, 128, 1)
  CUtensorMap kv_map;  // same tensor, box (64, 64, 1)
  int L, C, N;
  int dsplit;          // 1 or 2 CTAs per query tile
  int npv;             // output columns per CTA = C / dsplit (multiple of 64, <= 256)
  void* out;           // [N][L][out_ld] 16-bit
  long long out_ld;
  float scale_log2e;   // log2(e) / sqrt(C)
  uint32_t idesc_qk;   // M 128, N 64, both operands K-major
  uint32_t idesc_pv;   // M 128, N npv, A K-major, B MN-major
  int bf16;
};

__device__ __forceinline__ uint64_t desc_mn128(uint32_t smem_addr) {  // MN-major SW128: 64-key x 64-channel atoms
  uint64_t d = static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(8192 >> 4) << 16;   // next 64 channels (N)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;   // next 8 keys (K)
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(AT_THREADS, 1) attn_fwd_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int kchunks = p.C / 64;                       // 64-channel chunks of the head dimension
  uint8_t* sq = smem;                                 // Q: kchunks x [128 rows x 128 B]
  uint8_t* sk = sq + kchunks * 16384;                 // K ring: KSTAGES x [64 rows x 128 B]
  uint8_t* sp = sk + KSTAGES * 8192;                  // P: [128 rows x 128 B]
  uint8_t* sv = sp + 16384;                           // V: (npv / 64) x [64 keys x 128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sv + (p.npv / 64) * 8192);
  uint64_t* k_full = bars;                            // [KSTAGES]
  uint64_t* k_empty = bars + KSTAGES;                 // [KSTAGES]
  uint64_t* q_full = bars + 2 * KSTAGES;
  uint64_t* s_full = q_full + 1;    // [2]
  uint64_t* s_free = q_full + 3;    // [2]
  uint64_t* p_ready = q_full + 5;
  uint64_t* pv_done = q_full + 6;
  uint64_t* v_full = q_full + 7;
  uint64_t* o_full = q_full + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_full + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int dhalf = blockIdx.x % p.dsplit;
  const int qt = blockIdx.x / p.dsplit;
  const int img = blockIdx.y;
  const int q0 = qt * QROWS;
  const int nblocks = (p.L + KB - 1) / KB;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.q_map);
    prefetch_tmap(&p.kv_map);
    for (int i = 0; i < KSTAGES; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
    }
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 4);   // one arrive per softmax warp
    }
    mbar_init(p_ready, 4);
    mbar_init(pv_done, 1);
    mbar_init(v_full, 1);
    mbar_init(o_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base;        // two score tiles: columns [0, 64) and [64, 128)
  const uint32_t tmem_o = tmem_base + 128;  // columns [128, 128 + npv)

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, kchunks * 16384);
      for (int kc = 0; kc < kchunks; ++kc) tma_load_3d(&p.q_map, q_full, sq + kc * 16384, kc * 64, q0, img);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < nblocks; ++j) {
        for (int kc = 0; kc < kchunks; ++kc) {
          mbar_wait(&k_empty[stage], phase ^ 1);
          mbar_expect_tx(&k_full[stage], 8192);
          tma_load_3d(&p.kv_map, &k_full[stage], sk + stage * 8192, p.C + kc * 64, j * KB, img);
          if (++stage == KSTAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 6) {
    // V producer, a thread of its own: the single V tile is only free once the previous block's P V MMAs have retired, and
    // waiting for that inside the K producer's loop held back the K chunks of the NEXT block (and with them its Q K^T):
    // 0.44 -> 0.365 ms at batch 64
    if (lane == 0) {
      for (int j = 0; j < nblocks; ++j) {
        if (j > 0) mbar_wait(pv_done, (j - 1) & 1);
        mbar_expect_tx(v_full, (p.npv / 64) * 8192);
        for (int b = 0; b < p.npv / 64; ++b)
          tma_load_3d(&p.kv_map, v_full, sv + b * 8192, 2 * p.C + dhalf * p.npv + b * 64, j * KB, img);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      mbar_wait(q_full, 0);
      int stage = 0;
      uint32_t phase = 0;
      auto issue_qk = [&](int t) {   // score tile t into S buffer t & 1
        const int b = t & 1;
        if (t >= 2) mbar_wait(&s_free[b], ((t >> 1) - 1) & 1);   // softmax warps have drained the buffer's previous tile
        tc_fence_after();
        for (int kc = 0; kc < kchunks; ++kc) {
          mbar_wait(&k_full[stage], phase);
          tc_fence_after();
          const uint64_t da = make_smem_desc<128>(smem_u32(sq + kc * 16384));
          const uint64_t db = make_smem_desc<128>(smem_u32(sk + stage * 8192));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_f16(tmem_s + b * KB, da + static_cast<uint64_t>(k * 2), db + static_cast<uint64_t>(k * 2), p.idesc_qk,
                       (kc | k) != 0 ? 1u : 0u);
          tc_commit(&k_empty[stage]);
          if (++stage == KSTAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(&s_full[b]);
      };
      issue_qk(0);
      for (int j = 0; j < nblocks; ++j) {
        if (j + 1 < nblocks) issue_qk(j + 1);          // runs on the tensor pipe while the softmax warps work on tile j
        mbar_wait(p_ready, j & 1);
        mbar_wait(v_full, j & 1);
        tc_fence_after();
        const uint64_t dp = make_smem_desc<128>(smem_u32(sp));
        const uint32_t svb = smem_u32(sv);
#pragma unroll
        for (int k = 0; k < 4; ++k)   // 16 keys per MMA: P advances 32 B inside its 128-B rows, V two 8-key groups
          tc_mma_f16(tmem_o, dp + static_cast<uint64_t>(k * 2), desc_mn128(svb + k * 2048), p.idesc_pv, (j | k) != 0 ? 1u : 0u);
        tc_commit(pv_done);
      }
      tc_commit(o_full);
    }
  } else {
    const int sub = warp & 3;
    const int row = sub * 32 + lane;            // query row inside the tile = TMEM lane
    const int q = q0 + row;
    const uint32_t lane_addr = static_cast<uint32_t>(sub * 32) << 16;
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < nblocks; ++j) {
      const int b = j & 1;
      mbar_wait_long(&s_full[b], (j >> 1) & 1);
      tc_fence_after();
      float s[KB];
#pragma unroll
      for (int c = 0; c < KB; c += 16) tc_ld16(tmem_s + b * KB + lane_addr + c, reinterpret_cast<uint32_t*>(s + c));
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[b]);
      const int valid = p.L - j * KB;             // keys of this block inside the sequence
      float bm = -INFINITY;
#pragma unroll
      for (int c = 0; c < KB; ++c) {
        if (c >= valid) s[c] = -INFINITY;
        bm = fmaxf(bm, s[c]);
      }
      // LAZY reference maximum: the exponent reference only moves when the block maximum exceeds it by more than 8 in the
      // log2 domain (always on the first block, m = -inf).  Probabilities then reach at most 2^8 instead of 1 - exact in
      // fp32 / 16-bit floating point, O and l carry the same factor and it cancels in O / l - and the O rows in TMEM are
      // rescaled on a few blocks instead of whenever any of a warp's 32 rows sees a new maximum (most blocks).
      const bool upd = (bm - m) * p.scale_log2e > 8.0f;
      const float mn = upd ? bm : m;
      const float alpha = upd ? ex2f((m - mn) * p.scale_log2e) : 1.f;   // 0 on the first block
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < KB; ++c) {
        s[c] = ex2f((s[c] - mn) * p.scale_log2e);
        sum += s[c];
      }
      l = l * alpha + sum;
      m = mn;
      if (j > 0) {
        mbar_wait(pv_done, (j - 1) & 1);           // O holds blocks < j, the P tile and the V tile are free again
        tc_fence_after();
        if (__any_sync(0xffffffffu, alpha != 1.f)) {   // some row of this warp moved its maximum: rescale the warp's O rows
#pragma unroll 1
          for (int c = 0; c < p.npv; c += 64) {   // npv is a multiple of 64: four TMEM loads in flight per wait
            uint32_t raw[64];
#pragma unroll
            for (int u = 0; u < 4; ++u) tc_ld16(tmem_o + lane_addr + c + 16 * u, raw + 16 * u);
            tc_wait_ld();
#pragma unroll
            for (int h = 0; h < 64; ++h) raw[h] = __float_as_uint(__uint_as_float(raw[h]) * alpha);
#pragma unroll
            for (int u = 0; u < 4; ++u) tc_st16(tmem_o + lane_addr + c + 16 * u, raw + 16 * u);
          }
          tc_wait_st();
        }
      }
      uint8_t* prow = sp + row * 128;
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) {              // 8 keys = one 16-byte chunk; chunk index XOR (row & 7) = 128-B swizzle
        uint32_t w[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) w[h] = pack16(s[c8 * 8 + 2 * h], s[c8 * 8 + 2 * h + 1], p.bf16 ? EOVAE_BF16 : EOVAE_F16);
        *reinterpret_cast<uint4*>(prow + ((c8 ^ (row & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
    }
    const float inv_l = 1.f / l;
    // ---- epilogue: O / l -> 16-bit rows of the output
    mbar_wait_long(o_full, 0);
    tc_fence_after();
    uint16_t* orow = static_cast<uint16_t*>(p.out) + (static_cast<long long>(img) * p.L + q) * p.out_ld + dhalf * p.npv;
#pragma unroll 1
    for (int c = 0; c < p.npv; c += 64) {   // four TMEM loads in flight per wait
      uint32_t raw[64];
#pragma unroll
      for (int u = 0; u < 4; ++u) tc_ld16(tmem_o + lane_addr + c + 16 * u, raw + 16 * u);
      tc_wait_ld();
      if (q < p.L) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          uint32_t w[4];
#pragma unroll
          for (int h = 0; h < 4; ++h)
            w[h] = pack16(__uint_as_float(raw[8 * u + 2 * h]) * inv_l, __uint_as_float(raw[8 * u + 2 * h + 1]) * inv_l,
                          p.bf16 ? EOVAE_BF16 : EOVAE_F16);
          *reinterpret_cast<uint4*>(orow + c + 8 * u) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn attn_encode_fn() {
  static EncodeFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(ptr);
  }
  return fn;
}

int make_qkv_map(CUtensorMap* m, int dtype, const void* base, int channels, long long ld, int L, int n, int box_rows) {
  EncodeFn fn = attn_encode_fn();
  EOVAE_CHECK(fn != nullptr, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(channels), static_cast<cuuint64_t>(L), static_cast<cuuint64_t>(n)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(ld) * 2 * L};
  cuuint32_t box[3] = {64, static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = fn(m, dtype == EOVAE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3,
                  const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EOVAE_CHECK(r == CUDA_SUCCESS, "attention: cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r));
  return 0;
}

}  // namespace

extern "C" {

int eovae_attention_fused_ok(int l, int c) { return (c % 64 == 0 && c >= 64 && c <= 512 && (c <= 256 || (c / 2) % 64 == 0) && l >= 1) ? 1 : 0; }

int eovae_attention_fused(const void* qkv, long long qkv_ld, int n, int l, int c, void* out, long long out_ld, int dtype,
                          void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  EOVAE_CHECK(dtype == EOVAE_BF16 || dtype == EOVAE_F16, "attention_fused: 16-bit tensors only");
  EOVAE_CHECK(eovae_attention_fused_ok(l, c), "attention_fused: unsupported shape (L %d, C %d)", l, c);
  EOVAE_CHECK(qkv_ld % 8 == 0 && qkv_ld >= 3 * c && out_ld % 8 == 0 && out_ld >= c, "attention_fused: bad pitches");
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.L = l; p.C = c; p.N = n;
  p.dsplit = c > 256 ? 2 : 1;
  p.npv = c / p.dsplit;
  p.out = out;
  p.out_ld = out_ld;
  p.scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(c));
  p.bf16 = dtype == EOVAE_BF16 ? 1 : 0;
  const uint32_t fmt = dtype == EOVAE_BF16 ? 1u : 0u;
  p.idesc_qk = (1u << 4) | (fmt << 7) | (fmt << 10) | (static_cast<uint32_t>(KB >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
  p.idesc_pv = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 16) | (static_cast<uint32_t>(p.npv >> 3) << 17) |
               (static_cast<uint32_t>(128 >> 4) << 24);
  if (make_qkv_map(&p.q_map, dtype, qkv, 3 * c, qkv_ld, l, n, QROWS)) return -3;
  if (make_qkv_map(&p.kv_map, dtype, qkv, 3 * c, qkv_ld, l, n, KB)) return -3;
  const int smem = (c / 64) * 16384 + KSTAGES * 8192 + 16384 + (p.npv / 64) * 8192 + 256 + 1024;
  static int smem_set = 0;
  if (smem > smem_set) {
    EOVAE_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    smem_set = smem;
  }
  dim3 grid(ceil_div(l, QROWS) * p.dsplit, n);
  attn_fwd_kernel<<<grid, AT_THREADS, smem, stream>>>(p);
  EOVAE_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
