// tcgen05 / TMEM / TMA implicit-GEMM kernel for sm_100a.
//
//   D[m, n] = scale * sum_{tap, c} A_tap[m, c] * Wp[n, tap*Kc + c]   (+ bias[n]) (+ residual[m, n])
//
// m runs over output pixels (image n, row y, col x) of an NHWC activation tensor, A_tap is the same tensor
// shifted by the filter tap; every A tile is ONE 4-D TMA box (channels, x, y, image) whose out-of-bounds
// part (conv zero padding, ragged right/bottom edges, channel padding) is zero-filled by the TMA unit, so
// nothing is ever im2col'ed or padded in HBM.  B tiles are 3-D TMA boxes (k, cout, batch) of the packed
// K-major weight matrix (batch = 0 for convolutions, = image for the attention batched GEMMs).
//
// CTAS = 2 runs the mainloop on CTA PAIRS (tcgen05 cta_group::2, cluster of 2): one UMMA covers 256 pixels x BLOCK_N,
// each CTA stages its own 128-pixel A tile and HALF of the B tile, so the shared-memory operand traffic per MMA
// drops from (128 + N) to (128 + N/2) rows - the single-CTA SS-mode MMA is smem-read bound (measured: 157 / 187
// cycles per 128xNx16 MMA at N = 128 / 256 against 64 / 128 cycles of math).
//
// Warp roles (320 threads, 1 CTA / SM, persistent over tiles):
//   warp 0      : TMA producer (one elected lane), STAGES-deep smem ring, mbarrier full/empty
//   warp 1      : TMEM allocator + tcgen05.mma issuer (one lane); accumulators double-buffered in TMEM
//   warps 2..9  : epilogue - two warps per TMEM sub-partition, each owning half of the tile's columns:
//                 tcgen05.ld 32 lanes x 32 columns, bias from shared memory, residual (L2-prefetched one mainloop
//                 ahead), 16-byte stores, and optionally the GroupNorm partial sums of the tile (warp-shuffle
//                 reduce-scatter over the 32 rows, fixed-slot writes -> deterministic, no atomics)
#pragma once
#include "common.cuh"

namespace igemm {

constexpr int BLOCK_M = 128;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + 32 * NUM_EPI_WARPS;
// GroupNorm-prologue variant: 5 warpgroups = {TMA, MMA, 2 idle} | 4 + 4 epilogue warps | 4 + 4 transform warps
constexpr int NUM_XF_WARPS = 8;
constexpr int NUM_THREADS_GNP = 128 + 32 * NUM_EPI_WARPS + 32 * NUM_XF_WARPS;
constexpr int GNP_COEF_BYTES = 4096;  // GroupNorm-prologue kernels: per-image coefficient table [Cin <= 512][8 bytes] in shared memory
constexpr int MAX_TAPS = 16;  // 9 for a 3x3 kernel; 16 = 4 sub-pixel phases x 2x2 taps (data gradient of the upsample conv)

// Division by a launch-time constant: q = umulhi(n, mul) >> (shift - 1), exact for 0 <= n < 2^31 with
// shift = ceil(log2 d), mul = ceil(2^(31 + shift) / d) (mul == 0 encodes d == 1).  The persistent kernels decode a work
// index with seven divisions per tile in every epilogue warp (~25 instructions each as runtime divisions).
struct FastDiv {
  uint32_t mul, shift, d;
  __host__ void set(int div) {
    d = static_cast<uint32_t>(div);
    if (div <= 1) { mul = 0; shift = 0; return; }
    shift = 0;
    while ((1u << shift) < d) ++shift;
    mul = static_cast<uint32_t>(((1ull << (31 + shift)) + d - 1) / d);
  }
  __device__ __forceinline__ int div(int n) const {
    return mul == 0 ? n : static_cast<int>(__umulhi(static_cast<uint32_t>(n), mul) >> (shift - 1));
  }
  __device__ __forceinline__ void divmod(int n, int& q, int& r) const {
    q = div(n);
    r = n - q * static_cast<int>(d);
  }
};

struct Params {
  CUtensorMap a_map[4];
  CUtensorMap b_map;
  int box_w, box_h, box_n;        // pixels per A box (box_w*box_h*box_n <= 128)
  int tiles_w, tiles_h, tiles_n;  // m-tile grid
  int n_tiles;                    // tiles along Cout
  int num_taps, chunks_per_tap;   // main K items = num_taps * chunks_per_tap
  int k_per_tap;                  // padded channels per tap in the packed weight matrix
  int extra_chunks;               // K items of a fused 1x1 operand read through a_map[extra_map] at the output pixel
  int extra_map;                  //   (ResnetBlock nin_shortcut folded into conv2); its weights follow the taps in K
  int tap_map[MAX_TAPS], tap_dx[MAX_TAPS], tap_dy[MAX_TAPS];
  int Wo, Ho, Nimg;               // output extent
  int Cout;                       // valid output channels
  int b_batched;                  // B tile batch coordinate = image index (needs box_n == 1)
  void* out;
  int out_dtype;
  long long out_pix_stride;       // elements between consecutive output pixels
  const void* res;
  int res_dtype;
  long long res_pix_stride;
  const float* bias;
  float out_scale;
  uint32_t idesc;
  // fused GroupNorm statistics of the OUTPUT tensor (nullptr = off; needs box_n == 1, Cout % 32 == 0):
  // gn_partial[((m_tile * 4 + sub) * groups + g) * 2 + {0,1}] = (sum, sum of squares) over the 32 rows of the warp
  float* gn_partial;
  int gn_cpg;                     // channels per group (1, 2, 4, 8, 16 or 32)
  int gn_groups;                  // Cout / gn_cpg
  int gn_cpg_log2;                // log2(gn_cpg) (the supported group widths are powers of two)
  FastDiv fd_phases, fd_ntiles, fd_tw, fd_th;  // work index -> (phase, n-tile, tile x, tile y, tile image)
  int debug_mode;                 // bit mask: 1 no epilogue work | 2 no MMA issue | 4 no TMA loads | 32 no bulk store issue | 64 no tcgen05.ld (tools/igemm_bench.py, epilogue_probe.py)
  // fused GroupNorm + SiLU of the INPUT (GNP kernels): A operand = silu(x * a[n][c] + b[n][c]), zero outside the image
  const float2* gnp_ab;           // [Nimg][Cin] (gamma * rstd, beta - mean * gamma * rstd)
  int gnp_cin;                    // channels of the input tensor
  int gnp_h, gnp_w;               // input extent (padding mask)
  int gnp_bf16;                   // activation element type of the A operand
  CUtensorMap out_map;            // 16-bit output, box = (32 channels, the 32 pixels of one epilogue warp), 64 B swizzle
  int out_tma;                    // 1: epilogue stores through out_map (shared-memory staging + bulk tensor store);
                                  // 2: the same with 64-channel boxes / 128-byte swizzle (BLOCK_N = 128 kernels)
  // Sub-pixel phases (nearest-x2 upsample + 3x3 conv as four 2x2 convs on the LOW-resolution input, layers.py:47-50): the
  // work item gains a phase index; phase q reads the taps shifted by (phase_dx, phase_dy)[q], uses the weight rows
  // [q * b_phase_rows, ...) of the packed matrix, stores through the q-th output map (the parity sub-lattice
  // out[:, py::2, px::2, :] of the high-resolution tensor) and writes its GroupNorm partials into the q-th region.
  int phases;                     // 1 (ordinary convolution) or 4
  int phase_dx[4], phase_dy[4];
  int b_phase_rows;
  long long gn_phase_stride;      // floats between the phases' regions of gn_partial
  CUtensorMap out_map_ph[3];      // output maps of phases 1..3 (phase 0 = out_map)
  CUtensorMap res_map;            // residual with the geometry of the wide out_map (res_tma = 1)
  int res_tma;                    // 1: the epilogue warps pull their residual box into the store staging buffer by TMA
};

// ------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// HINT_NS > 0: the thread may be suspended up to that long while the phase is pending (it is woken when the barrier
// completes) instead of re-polling at once.  Measured on B200, sustained (power-capped) runs, two boxes, builds compared in one
// job: every wait hinted with 500 / 2000 / 20000 ns = encode step -1.4 ... -2.5 %, training step -0.3 ... -1.8 %; hinting
// only the long epilogue-side waits = no change.  The default hints every wait with 1000 ns.
template <uint32_t HINT_NS = 0>
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  if constexpr (HINT_NS > 0) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(HINT_NS)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
  return ok != 0;
}
#ifndef EOVAE_WAIT_HINT_NS
#define EOVAE_WAIT_HINT_NS 1000
#endif
#ifndef EOVAE_LONG_WAIT_HINT_NS
#define EOVAE_LONG_WAIT_HINT_NS 0
#endif
// Bounded wait: a protocol bug traps (launch error) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait<EOVAE_WAIT_HINT_NS>(bar, parity)) {
    if (++spins > (1u << 21)) __trap();  // ~10 s of polling: far beyond any legitimate wait
  }
}
// waits of many threads for a whole mainloop (epilogue side); EOVAE_LONG_WAIT_HINT_NS gives them a longer hint
__device__ __forceinline__ void mbar_wait_long(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait<(EOVAE_LONG_WAIT_HINT_NS > EOVAE_WAIT_HINT_NS ? EOVAE_LONG_WAIT_HINT_NS : EOVAE_WAIT_HINT_NS)>(bar, parity)) {
    if (++spins > (1u << 21)) __trap();
  }
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// ---- CTA-pair (cta_group::2) variants.  A shared::cta address is a valid shared::cluster address of the executing
// CTA; clearing bit 24 (cute::Sm100MmaPeerBitMask) addresses the same offset in the even (leader) CTA of the pair.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {  // arrive on the leader CTA's copy of `bar`
  // default (.release.cta) semantics as in cutlass::arch::ClusterBarrier::arrive: a cluster-scope release here costs
  // an L1 invalidate + ~1e3 cycles per call and serialised the peer's producer (measured: 1800 cycles per k-block)
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}
__device__ __forceinline__ void tma2_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                             int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1),
      "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma2_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1),
      "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc2_commit_mc(uint64_t* bar) {  // arrive on `bar` in BOTH CTAs of the pair
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}
__device__ __forceinline__ void tc2_mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand tile in shared memory: rows of CHUNK_BYTES (32/64/128) bytes, hardware swizzle of the same
// width, 8-row groups CHUNK_BYTES*8 apart (cute::UMMA::SmemDescriptor, version 1 = Blackwell).
template <int CHUNK_BYTES>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t base_offset = 0) {
  constexpr uint64_t layout = CHUNK_BYTES == 128 ? 2 : (CHUNK_BYTES == 64 ? 4 : 6);
  uint64_t d = static_cast<uint64_t>(base_offset & 7) << 49;        // matrix base offset  bits [49,52)
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);            // start address      bits [0,14)
  d |= static_cast<uint64_t>(1) << 16;                               // leading byte off.  bits [16,30) (unused for swizzled K-major)
  d |= static_cast<uint64_t>((CHUNK_BYTES * 8) >> 4) << 32;          // stride byte offset bits [32,46)
  d |= static_cast<uint64_t>(1) << 46;                               // descriptor version bits [46,48)
  d |= layout << 61;                                                 // swizzle mode       bits [61,64)
  return d;
}

// KCH = K-chunks (of CHUNK_BYTES) per pipeline stage.  The single producer / MMA threads pay ~300 cycles of serial
// latency per stage (mbarrier try_wait, tcgen05.commit, TMA issue): measured 0.65 ms of pure handshake on a 1.3 ms
// 128->128 conv with 64-channel stages, so wide-channel layers use two chunks (K = 128) per stage.
//
// HALO (3x3 stride-1 convs whose m-tile is 128 consecutive pixels of ONE image row, 128-byte chunks): a stage holds
// one (kernel row, 64-channel chunk): the input row segment is loaded ONCE with a one-pixel halo on each side (130
// pixels) and the three horizontal taps are three UMMA views of the same shared-memory tile, shifted by one 128-byte
// row each (descriptor start address + matrix base offset) - A traffic through L2 drops 3x.
constexpr int HALO_ROWS = 130;
template <int BLOCK_N, int CHUNK_BYTES, int CTAS, int KCH, bool HALO = false>
struct Config {
  static constexpr int A_CHUNK_BYTES = HALO ? ((HALO_ROWS * CHUNK_BYTES + 1023) / 1024 * 1024) : BLOCK_M * CHUNK_BYTES;
  static constexpr int A_BYTES = A_CHUNK_BYTES * (HALO ? 1 : KCH);
  static constexpr int B_ROWS = BLOCK_N / CTAS;  // rows of the B tile staged by each CTA
  static constexpr int B_BYTES_RAW = B_ROWS * CHUNK_BYTES;
  static constexpr int B_CHUNK_BYTES = (B_BYTES_RAW + 1023) / 1024 * 1024;
  static constexpr int B_PER_STAGE = HALO ? 3 : KCH;
  static constexpr int B_BYTES = B_CHUNK_BYTES * B_PER_STAGE;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_BYTES = NUM_EPI_WARPS * 2 * 32 * 64;  // per-warp double-buffered TMA-store staging
  static constexpr int AUX_BYTES = 1024;                          // barriers + tmem slot (keeps the staging 1 KB aligned)
  static constexpr int MAX_SMEM = 232448;                         // 227 KB per CTA on sm_100
  static constexpr int BUDGET = MAX_SMEM - 1024 /*align slack*/ - AUX_BYTES - EPI_BYTES;
  static constexpr int STAGES_RAW = BUDGET / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + AUX_BYTES + EPI_BYTES;
  static constexpr int TMEM_COLS = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64 ? 64 : (2 * BLOCK_N <= 128 ? 128 : (2 * BLOCK_N <= 256 ? 256 : 512)));
};

// Warp-level reduce-scatter of NV per-lane values over the 32 lanes: afterwards lane l holds in vals[0] the
// total of value index (l >> (5 - log2 NV))... see gn_reduce() for the index formula.  NV shuffles instead of 5*NV.
template <int NV>
__device__ __forceinline__ void reduce_scatter(float (&vals)[NV], int lane) {
  int nv = NV;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    if (nv > 1) {
      const int half = nv >> 1;
      const bool upper = (lane & off) != 0;
#pragma unroll
      for (int i = 0; i < NV / 2; ++i) {
        if (i < half) {
          const float send = upper ? vals[i] : vals[i + half];
          const float keep = upper ? vals[i + half] : vals[i];
          vals[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
      }
      nv = half;
    } else {
      vals[0] += __shfl_xor_sync(0xffffffffu, vals[0], off);
    }
  }
}

// GroupNorm partial sums of one 32-column chunk: v = final fp32 outputs of this thread's row (zero for invalid rows).
template <int CPG>
__device__ __forceinline__ void gn_chunk(const float (&v)[32], int lane, float* dst /* [32/CPG groups][2] */) {
  constexpr int G = 32 / CPG;
  constexpr int NV = 2 * G;
  float vals[NV];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int j = 0; j < CPG; ++j) {
      const float t = v[g * CPG + j];
      s += t;
      q = fmaf(t, t, q);
    }
    vals[g] = s;          // sums in the lower half, squares in the upper half
    vals[G + g] = q;
  }
  reduce_scatter<NV>(vals, lane);
  // after log2(NV) halving steps (offsets 16, 8, ...) lane l owns value index = top log2(NV) bits of l
  constexpr int LOG = (NV == 64) ? 6 : (NV == 32) ? 5 : (NV == 16) ? 4 : (NV == 8) ? 3 : (NV == 4) ? 2 : 1;
  // value index i < G is the sum of group i, i >= G the sum of squares of group i - G
  auto put = [&](int idx, float val) { dst[(idx < G ? idx : idx - G) * 2 + (idx < G ? 0 : 1)] = val; };
  if constexpr (NV <= 32) {
    if ((lane & ((1 << (5 - LOG)) - 1)) == 0) put(lane >> (5 - LOG), vals[0]);
  } else {  // NV == 64: five halvings leave two consecutive indices per lane
    put(2 * lane, vals[0]);
    put(2 * lane + 1, vals[1]);
  }
}

// Two adjacent 32-column chunks of a warp reduced TOGETHER: the shuffle tree of a chunk is a dependent chain of five
// exchange levels (~150 cycles of latency for ~50 instructions) and the epilogue warps have little else to overlap it with,
// so the per-thread sums of the first chunk wait in registers and one reduce-scatter over both chunks' values runs the two
// trees side by side.  Every value still follows the same butterfly (lane l with l ^ 16, ^ 8, ...): bit-identical sums.
template <int CPG>
__device__ __forceinline__ void gn_thread_sums(const float (&v)[32], float* vals /* [2 * 32 / CPG]: sums, then squares */) {
  constexpr int G = 32 / CPG;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int j = 0; j < CPG; ++j) {
      const float t = v[g * CPG + j];
      s += t;
      q = fmaf(t, t, q);
    }
    vals[g] = s;
    vals[G + g] = q;
  }
}
template <int CPG>
__device__ __forceinline__ void gn_pair_finish(const float* lo /* stashed sums of the first chunk */, const float (&v)[32],
                                               int lane, float* dst /* of the FIRST chunk: [2 * 32 / CPG groups][2] */) {
  constexpr int G = 32 / CPG;
  constexpr int NV = 4 * G;  // <= 32 for CPG >= 4
  float vals[NV];
#pragma unroll
  for (int i = 0; i < 2 * G; ++i) vals[i] = lo[i];
  gn_thread_sums<CPG>(v, vals + 2 * G);
  reduce_scatter<NV>(vals, lane);
  constexpr int LOG = (NV == 32) ? 5 : (NV == 16) ? 4 : (NV == 8) ? 3 : 2;
  if ((lane & ((1 << (5 - LOG)) - 1)) == 0) {
    const int idx = lane >> (5 - LOG);         // chunk * 2G + stat * G + g
    const int chunk = idx / (2 * G), r = idx % (2 * G);
    dst[((chunk * G) + (r % G)) * 2 + r / G] = vals[0];
  }
}

// work index -> (phase, n-tile, m-tile of this CTA, tile x / y / image)
struct WorkItem { int ph, nt, mt, tw, th, tn; };
__device__ __forceinline__ WorkItem decode_work(const Params& p, int work, int ctas, int cta_rank) {
  WorkItem w;
  int wk, q;
  p.fd_phases.divmod(work, wk, w.ph);
  p.fd_ntiles.divmod(wk, q, w.nt);
  w.mt = q * ctas + cta_rank;
  int t;
  p.fd_tw.divmod(w.mt, t, w.tw);
  p.fd_th.divmod(t, w.tn, w.th);
  return w;
}

template <int BLOCK_N, int CHUNK_BYTES, int CTAS, int KCH, bool HALO = false, bool GNP = false>
__global__ void __launch_bounds__(GNP ? NUM_THREADS_GNP : NUM_THREADS, 1) igemm_kernel(const __grid_constant__ Params p) {
  static_assert(!GNP || (HALO && CTAS == 2), "the GroupNorm prologue lives in the CTA-pair halo mainloop");
  using Cfg = Config<BLOCK_N, CHUNK_BYTES, CTAS, KCH, HALO>;
  constexpr int EPI_WARP0 = GNP ? 4 : 2;   // first epilogue warp
  constexpr int XF_WARP0 = 12;             // first transform warp (GNP)
  const uint32_t cta_rank = CTAS == 2 ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  const int group_id = CTAS == 2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int num_groups = CTAS == 2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  constexpr int STAGES = Cfg::STAGES;
  constexpr int CH_ELEMS = CHUNK_BYTES / 2;  // 16-bit elements per k-chunk
  constexpr int K_STEPS = CHUNK_BYTES / 32;  // UMMA_K = 16 elements = 32 bytes

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  uint64_t* a_full = bars + 2 * STAGES + 6;       // GNP: this CTA's raw A tile has landed (local)
  uint64_t* ready_bar = bars + 3 * STAGES + 6;    // GNP: A tiles of the whole group are transformed (leader's copy)
  uint64_t* res_bar = bars + 4 * STAGES + 6;      // per epilogue warp: its residual box has landed in its staging buffer
  uint8_t* epi_smem = smem + STAGES * Cfg::STAGE_BYTES + 1024;  // TMA-store staging: 8 warps x 2 x 2 KB, 1 KB aligned
  constexpr int EPI_BUF_BYTES = 32 * 64;
  uint8_t* coef_smem = epi_smem + Cfg::EPI_BYTES;  // GNP only (the launch adds GNP_COEF_BYTES)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  // work item = CTAS adjacent m-tiles x one n-tile (x one sub-pixel phase, fastest: the phases of a tile share its A rows in L2)
  const int total_work = ((m_tiles + CTAS - 1) / CTAS) * p.n_tiles * p.phases;
  const int main_items = p.num_taps * p.chunks_per_tap;
  // pipeline stages per tile; HALO: one stage per (kernel row, channel chunk), else KCH K-items per stage
  const int num_kb = HALO ? 3 * p.chunks_per_tap : (main_items + p.extra_chunks) / KCH;  // host guarantees divisibility

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) prefetch_tmap(&p.a_map[i]);
    prefetch_tmap(&p.b_map);
    if (p.out_tma) prefetch_tmap(&p.out_map);
    for (int i = 1; i < p.phases; ++i) prefetch_tmap(&p.out_map_ph[i - 1]);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], CTAS);   // one producer arrive per CTA of the group (the leader's copy is the live one)
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < NUM_EPI_WARPS; ++i) mbar_init(&res_bar[i], 1);
    if (p.res_tma) prefetch_tmap(&p.res_map);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], NUM_EPI_WARPS * CTAS);  // one arrive per epilogue warp of every CTA of the group
    }
    if constexpr (GNP) {
      for (int i = 0; i < STAGES; ++i) {
        mbar_init(&a_full[i], 1);
        mbar_init(&ready_bar[i], NUM_XF_WARPS * CTAS);  // one arrive per transform warp of every CTA of the group
      }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if constexpr (CTAS == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"(Cfg::TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"(Cfg::TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if constexpr (CTAS == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // GNP: 640 threads start with 96 registers each (61440 in the CTA's pool - setmaxnreg.inc can only take what the
  // CTA's own warps released, so the new budgets must sum to <= 61440 or the kernel deadlocks); every warpgroup
  // re-sizes at the top of ITS branch (ptxas allocates per region dominated by a setmaxnreg):
  // 40 | 168 | 168 | 48 | 48 registers x 128 threads = 60416.
  if (warp < EPI_WARP0) {
  if constexpr (GNP) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (every CTA of the group)
    if (lane == 0) {
      const uint32_t a_box_bytes = static_cast<uint32_t>(p.box_w * p.box_h * p.box_n) * CHUNK_BYTES;
      const uint32_t tx_bytes = HALO ? (HALO_ROWS * CHUNK_BYTES + 3 * Cfg::B_BYTES_RAW) * CTAS
                                     : (a_box_bytes + Cfg::B_BYTES_RAW) * CTAS * KCH;  // bytes landing in ALL CTAs of the group
      int stage = 0;
      uint32_t phase = 0;
      for (int work = group_id; work < total_work; work += num_groups) {
        const WorkItem wi_ = decode_work(p, work, CTAS, static_cast<int>(cta_rank));
        const int ph = wi_.ph, nt = wi_.nt, mt = wi_.mt, tw = wi_.tw, th = wi_.th;
        const int tn = wi_.tn;  // == tiles_n for the padding tile of an odd tail: fully OOB -> zeros
        const int x0 = tw * p.box_w + p.phase_dx[ph], y0 = th * p.box_h + p.phase_dy[ph], img0 = tn * p.box_n;
        const int bb = p.b_batched ? img0 : 0;
        const int b_row0 = nt * BLOCK_N + static_cast<int>(cta_rank) * Cfg::B_ROWS + ph * p.b_phase_rows;
        int h_kh = 0, h_cc = 0;  // K item -> (kernel row | tap, channel chunk) by increments, no division per stage
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (p.debug_mode & 4) {
            if (leader) mbar_arrive(&full_bar[stage]); else mbar_arrive_leader(&full_bar[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          uint8_t* sa = smem_a + stage * Cfg::A_BYTES;
          uint8_t* sb = smem_b + stage * Cfg::B_BYTES;
          if constexpr (GNP) {
            // raw A tile -> this CTA's own barrier (its transform warps wait there); B tiles -> the leader's barrier
            mbar_expect_tx(&a_full[stage], HALO_ROWS * CHUNK_BYTES);
            if (leader) mbar_expect_tx(&full_bar[stage], 3 * Cfg::B_BYTES_RAW * CTAS); else mbar_arrive_leader(&full_bar[stage]);
          } else if constexpr (CTAS == 2) {
            if (leader) mbar_expect_tx(&full_bar[stage], tx_bytes); else mbar_arrive_leader(&full_bar[stage]);
          } else {
            mbar_expect_tx(&full_bar[stage], tx_bytes);
          }
          if constexpr (HALO) {
            const int kh = h_kh;
            const int c0 = h_cc * CH_ELEMS;
            if (++h_cc == p.chunks_per_tap) { h_cc = 0; ++h_kh; }
            // one input row segment with its halo: pixels [x0 - 1, x0 + 128], row y0 + kh - 1 (OOB -> zeros)
            if constexpr (GNP) tma_load_4d(&p.a_map[3], &a_full[stage], sa, c0, x0 - 1, y0 + kh - 1, img0);
            else tma2_load_4d(&p.a_map[3], &full_bar[stage], sa, c0, x0 - 1, y0 + kh - 1, img0);
#pragma unroll
            for (int kw = 0; kw < 3; ++kw)
              tma2_load_3d(&p.b_map, &full_bar[stage], sb + kw * Cfg::B_CHUNK_BYTES, (kh * 3 + kw) * p.k_per_tap + c0,
                           b_row0, bb);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
#pragma unroll
          for (int j = 0; j < KCH; ++j) {
            // K item -> (tensor map, pixel shift, channel chunk, weight column)
            const int item = kb * KCH + j;
            const CUtensorMap* amap;
            int ax, ay, c0, k0;
            if (item < main_items) {
              const int tap = h_kh;  // item -> (tap, channel chunk) by increments
              c0 = h_cc * CH_ELEMS;
              if (++h_cc == p.chunks_per_tap) { h_cc = 0; ++h_kh; }
              amap = &p.a_map[p.tap_map[tap]];
              ax = x0 + p.tap_dx[tap];
              ay = y0 + p.tap_dy[tap];
              k0 = tap * p.k_per_tap + c0;
            } else {
              c0 = (item - main_items) * CH_ELEMS;
              amap = &p.a_map[p.extra_map];
              ax = x0;
              ay = y0;
              k0 = p.num_taps * p.k_per_tap + c0;
            }
            if constexpr (CTAS == 2) {
              tma2_load_4d(amap, &full_bar[stage], sa + j * Cfg::A_CHUNK_BYTES, c0, ax, ay, img0);
              tma2_load_3d(&p.b_map, &full_bar[stage], sb + j * Cfg::B_CHUNK_BYTES, k0, b_row0, bb);
            } else {
              tma_load_4d(amap, &full_bar[stage], sa + j * Cfg::A_CHUNK_BYTES, c0, ax, ay, img0);
              tma_load_3d(&p.b_map, &full_bar[stage], sb + j * Cfg::B_CHUNK_BYTES, k0, b_row0, bb);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA of the group only)
    if (lane == 0 && leader) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int work = group_id; work < total_work; work += num_groups, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          if constexpr (GNP) mbar_wait(&ready_bar[stage], phase);
          tc_fence_after();
          if constexpr (HALO) {
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              // tap kw = the halo tile shifted by kw pixels = kw 128-byte rows.  The 128-byte swizzle is a function of
              // the absolute shared-memory address bits (TMA wrote it that way), so a row-shifted start address needs
              // NO matrix-base-offset (verified on B200: base offset = kw gives garbage, 0 is bit-exact)
              const uint64_t da = make_smem_desc<CHUNK_BYTES>(smem_u32(smem_a + stage * Cfg::A_BYTES + kw * CHUNK_BYTES));
              const uint64_t db = make_smem_desc<CHUNK_BYTES>(smem_u32(smem_b + stage * Cfg::B_BYTES + kw * Cfg::B_CHUNK_BYTES));
#pragma unroll
              for (int k = 0; k < K_STEPS; ++k) {
                if (p.debug_mode & 2) break;
                tc2_mma_f16(d_tmem, da + static_cast<uint64_t>(k * 2), db + static_cast<uint64_t>(k * 2), p.idesc,
                            (kb | kw | k) != 0 ? 1u : 0u);
              }
            }
            tc2_commit_mc(&empty_bar[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
#pragma unroll
          for (int j = 0; j < KCH; ++j) {
            const uint64_t da = make_smem_desc<CHUNK_BYTES>(smem_u32(smem_a + stage * Cfg::A_BYTES + j * Cfg::A_CHUNK_BYTES));
            const uint64_t db = make_smem_desc<CHUNK_BYTES>(smem_u32(smem_b + stage * Cfg::B_BYTES + j * Cfg::B_CHUNK_BYTES));
#pragma unroll
            for (int k = 0; k < K_STEPS; ++k) {
              if (p.debug_mode & 2) break;
              // advance the 14-bit start-address field by k*32 bytes (>>4)
              if constexpr (CTAS == 2)
                tc2_mma_f16(d_tmem, da + static_cast<uint64_t>(k * 2), db + static_cast<uint64_t>(k * 2), p.idesc,
                            (kb | j | k) != 0 ? 1u : 0u);
              else
                tc_mma_f16(d_tmem, da + static_cast<uint64_t>(k * 2), db + static_cast<uint64_t>(k * 2), p.idesc,
                           (kb | j | k) != 0 ? 1u : 0u);
            }
          }
          // frees the smem slot (in every CTA of the group) when these MMAs retire
          if constexpr (CTAS == 2) tc2_commit_mc(&empty_bar[stage]); else tc_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if constexpr (CTAS == 2) tc2_commit_mc(&tmem_full[acc]); else tc_commit(&tmem_full[acc]);  // accumulator complete
      }
    }
  }
  } else if (GNP && warp >= XF_WARP0) {
    // ------------------------------------------------------------------ transform warps (GroupNorm + SiLU prologue)
    if constexpr (GNP) asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
    // STATUS: correct (tests/test_kernels_gpu.py::test_conv2d_gn_prologue) but NOT used by default: with the 48
    // registers the CTA's register pool leaves them, the 8 transform warps reach an IPC of ~0.75 and need ~3400 cycles
    // per (kernel row, chunk) stage against ~1150 cycles of MMA - the fused conv is 1.95 ms where gn_apply (0.41 ms,
    // 80 % of HBM peak) + conv (1.10 ms) take 1.51 ms (profiles/r1_gn_prologue_experiment.md).
    // The raw halo tile is normalised IN PLACE before the MMA reads it: y = silu(x * a + b) per (image, channel),
    // pixels outside the image stay zero (conv padding applies to the normalised tensor).  Each thread owns one
    // 16-byte piece column (8 channels) and walks the 130 rows in steps of 32.
    if constexpr (GNP) {
      const int tl = threadIdx.x - XF_WARP0 * 32;  // 0..255
      const int j = tl & 7, rg = tl >> 3;
      int stage = 0;
      uint32_t phase = 0;
      int coef_img = -1;  // image whose coefficients sit in coef_smem
      for (int work = group_id; work < total_work; work += num_groups) {
        const WorkItem wi_ = decode_work(p, work, CTAS, static_cast<int>(cta_rank));  // phases == 1 here
        const int mt = wi_.mt, tw = wi_.tw, th = wi_.th, tn = wi_.tn;
        const int x0 = tw * p.box_w, y0 = th * p.box_h;
        if (tn != coef_img && mt < m_tiles) {
          // The per-(image, channel) coefficients come from global memory once per IMAGE, not once per stage: a stage gives a
          // thread four 16-byte vectors of work, so a dependent global load in front of them (measured: the transform alone
          // took 3000 cycles per stage, whatever its arithmetic) was the whole cost of the fused prologue.
          asm volatile("bar.sync 1, %0;" ::"n"(NUM_XF_WARPS * 32) : "memory");  // nobody still reads the old table
          const float2* src = p.gnp_ab + static_cast<long long>(tn) * p.gnp_cin;
          for (int i = tl; i < p.gnp_cin; i += NUM_XF_WARPS * 32) reinterpret_cast<float2*>(coef_smem)[i] = __ldg(src + i);
          asm volatile("bar.sync 1, %0;" ::"n"(NUM_XF_WARPS * 32) : "memory");
          coef_img = tn;
        }
        int x_kh = 0, x_cc = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          const int kh = x_kh;
          const int c0 = x_cc * CH_ELEMS + j * 8;
          if (++x_cc == p.chunks_per_tap) { x_cc = 0; ++x_kh; }
          const int y = y0 + kh - 1;
          mbar_wait(&a_full[stage], phase);
          if (mt < m_tiles && y >= 0 && y < p.gnp_h && !(p.debug_mode & 16)) {
            const float4* abp = reinterpret_cast<const float4*>(coef_smem) + (c0 >> 1);   // two channels per float4
            const uint32_t tile = smem_u32(smem_a + stage * Cfg::A_BYTES);
            if (!p.gnp_bf16) {
              // fp16 operands: PACKED half2 math, 4 instructions per channel pair (sub, fma, tanh.approx.f16x2, fma) instead of
              // ~11 fp32 ones per element - the fp32 form kept these 8 warps at ~3400 cycles per stage against ~1150 of MMA.
              // silu(z) = h + h tanh(h), h = z / 2 = (x - m16) * (a / 2) + bh: the mean is subtracted FIRST (m16 = the fp16
              // value nearest the mean, so x - m16 is exact or one rounding of a small number) and the coefficient rounding
              // then acts on the centred value only; bh carries the fp32 remainder (beta + (m16 - mean) a) / 2.
              uint32_t m2[4], a2[4], b2[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float4 f = abp[q];   // two channels: (bits of half2(m16, a/2), bh) each
                const uint32_t e0 = __float_as_uint(f.x), e1 = __float_as_uint(f.z);
                m2[q] = (e0 & 0xFFFFu) | (e1 << 16);
                a2[q] = (e0 >> 16) | (e1 & 0xFFFF0000u);
                b2[q] = T16<__half>::from_f2(f.y, f.w);
              }
              auto xf = [&](uint32_t x, int q) {
                uint32_t d, h, t, o;
                asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(x), "r"(m2[q]));
                asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(h) : "r"(d), "r"(a2[q]), "r"(b2[q]));
                asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(h));
                asm("fma.rn.f16x2 %0, %1, %2, %1;" : "=r"(o) : "r"(h), "r"(t));
                return o;
              };
              // four rows per trip: their loads are issued together (ILP for the 48-register transform warps)
              for (int r0 = rg; r0 < HALO_ROWS; r0 += 4 * 4 * NUM_XF_WARPS) {
                uint32_t v[4][4];
                uint32_t addr[4];
                bool ok[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const int r = r0 + u * 4 * NUM_XF_WARPS;
                  const int x = x0 - 1 + r;
                  ok[u] = r < HALO_ROWS && x >= 0 && x < p.gnp_w;
                  addr[u] = tile + r * 128 + ((j ^ (r & 7)) << 4);
                  if (ok[u])
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[u][0]), "=r"(v[u][1]), "=r"(v[u][2]), "=r"(v[u][3]) : "r"(addr[u]));
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  if (!ok[u]) continue;
#pragma unroll
                  for (int q = 0; q < 4; ++q) v[u][q] = xf(v[u][q], q);
                  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr[u]), "r"(v[u][0]), "r"(v[u][1]), "r"(v[u][2]), "r"(v[u][3]) : "memory");
                }
              }
            } else {
            float ca[8], cb[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 f = abp[q];
              ca[2 * q] = f.x; cb[2 * q] = f.y; ca[2 * q + 1] = f.z; cb[2 * q + 1] = f.w;
            }
#pragma unroll 2
            for (int r = rg; r < HALO_ROWS; r += 4 * NUM_XF_WARPS) {
              const int x = x0 - 1 + r;
              if (x < 0 || x >= p.gnp_w) continue;
              const uint32_t addr = tile + r * 128 + ((j ^ (r & 7)) << 4);
              uint32_t v[4];
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(addr));
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float2 f = T16<__nv_bfloat16>::to_f2(v[q]);
                // silu(z) = z * sigmoid(z) = h + h * tanh(h) with h = z / 2: one MUFU op per element
                const float h0 = 0.5f * fmaf(f.x, ca[2 * q], cb[2 * q]), h1 = 0.5f * fmaf(f.y, ca[2 * q + 1], cb[2 * q + 1]);
                float t0, t1;
                asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(h0));
                asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(h1));
                const float o0 = fmaf(h0, t0, h0), o1 = fmaf(h1, t1, h1);
                v[q] = T16<__nv_bfloat16>::from_f2(o0, o1);
              }
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
            }
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA
          __syncwarp();
          if (lane == 0) {
            if (leader) mbar_arrive(&ready_bar[stage]); else mbar_arrive_leader(&ready_bar[stage]);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= EPI_WARP0 && warp < EPI_WARP0 + NUM_EPI_WARPS) {
    // ------------------------------------------------------------------ epilogue warps
    if constexpr (GNP) asm volatile("setmaxnreg.inc.sync.aligned.u32 168;");
    const int ew = warp - EPI_WARP0;
    const int sub = warp & 3;    // TMEM sub-partition this warp may access: lanes [32*sub, 32*sub+32)
    const int half = ew >> 2;    // which half of the tile's columns this warp owns
    constexpr int HALF_N = BLOCK_N >= 64 ? BLOCK_N / 2 : BLOCK_N;  // narrow tiles: only half 0 works
    constexpr int CW = BLOCK_N >= 32 ? 32 : 16;                    // chunk width (columns per tcgen05.ld round)
    const bool has_cols = (BLOCK_N >= 64) || (half == 0);
    const int col_begin = (BLOCK_N >= 64) ? half * HALF_N : 0;
    const int row = sub * 32 + lane;
    const int box_pix = p.box_w * p.box_h * p.box_n;
    const int wi = row % p.box_w;
    const int hi = (row / p.box_w) % p.box_h;
    const int ni = row / (p.box_w * p.box_h);
    // first pixel of this warp's 32 rows inside the tile (origin of its TMA-store sub-box)
    const int w_wi = (sub * 32) % p.box_w, w_hi = ((sub * 32) / p.box_w) % p.box_h, w_ni = (sub * 32) / (p.box_w * p.box_h);
    const bool out16 = p.out_dtype != EOVAE_F32;
    const bool out_bf16 = p.out_dtype == EOVAE_BF16;
    const bool use_tma_store = (CW == 32) && p.out_tma != 0;
    const bool wide_store = (HALF_N == 64) && p.out_tma == 2;  // out_map boxes are 64 channels wide, 128-byte swizzle
    // residual: either the warp's 32-pixel x 64-channel box arrives by TMA in the store staging buffer (coalesced,
    // asynchronous, under the mainloop of the tile), or every thread reads its own pixel row through registers
    const bool res_tma = wide_store && p.res_tma != 0;
    const bool has_res = p.res != nullptr && !res_tma;
    const bool res16 = has_res && p.res_dtype != EOVAE_F32;
    const bool res_bf16 = p.res_dtype == EOVAE_BF16;
    uint32_t res_phase = 0;
    uint8_t* stage_buf = epi_smem + ew * (2 * EPI_BUF_BYTES);  // two 32-row x 32-column 16-bit buffers per warp
    int sbuf = 0;
    int it = 0;
    for (int work = group_id; work < total_work; work += num_groups, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const WorkItem wi_ = decode_work(p, work, CTAS, static_cast<int>(cta_rank));
      const int ph = wi_.ph, nt = wi_.nt, mt = wi_.mt, tw = wi_.tw, th = wi_.th, tn = wi_.tn;
      const CUtensorMap* omap = ph == 0 ? &p.out_map : &p.out_map_ph[ph - 1];
      const int ox = tw * p.box_w + wi, oy = th * p.box_h + hi, on = tn * p.box_n + ni;
      const bool valid = mt < m_tiles && row < box_pix && ox < p.Wo && oy < p.Ho && on < p.Nimg;
      const long long pix = (static_cast<long long>(on) * p.Ho + oy) * p.Wo + ox;
      const int n_tile0 = nt * BLOCK_N;
      if constexpr (CW == 32 && HALF_N == 64) {
        if (res_tma && lane == 0) {
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the buffer's previous store has been read out
          if (mt < m_tiles && sub * 32 < box_pix && n_tile0 + col_begin < p.Cout) {
            mbar_expect_tx(&res_bar[ew], 32 * 128);
            tma_load_4d(&p.res_map, &res_bar[ew], stage_buf, n_tile0 + col_begin, tw * p.box_w + w_wi, th * p.box_h + w_hi,
                        tn * p.box_n + w_ni);
          } else {
            mbar_arrive(&res_bar[ew]);  // keeps the phase in step with the tile count
          }
        }
      }
      // pull this thread's residual row segment into L2 while the mainloop of this tile is still running
      if (has_res && valid && has_cols) {
        const int esz = p.res_dtype == EOVAE_F32 ? 4 : 2;
        const uint8_t* rp = reinterpret_cast<const uint8_t*>(p.res) + (pix * p.res_pix_stride + n_tile0 + col_begin) * esz;
        for (int b = 0; b < HALF_N * esz; b += 128) prefetch_l2(rp + b);
      }
      // 16-bit residual rows travel through registers one chunk ahead of their use; the first chunk is requested
      // here, before the wait for the accumulator, so its latency hides behind the mainloop
      const uint16_t* res_row = reinterpret_cast<const uint16_t*>(p.res) + pix * p.res_pix_stride;
      uint4 rbuf_a[4], rbuf_b[4];
      auto res_fetch = [&](uint4 (&rb)[4], int c) {  // residual columns [n_tile0 + c, + CW) of this thread's pixel
        if (res16 && valid && c < col_begin + HALF_N && n_tile0 + c + CW <= p.Cout) {
          const uint4* rp = reinterpret_cast<const uint4*>(res_row + n_tile0 + c);
#pragma unroll
          for (int j = 0; j < CW / 8; ++j) rb[j] = __ldg(rp + j);
        }
      };
      if (has_cols) {
        res_fetch(rbuf_a, col_begin);
        res_fetch(rbuf_b, col_begin + CW);
      }
      mbar_wait_long(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(sub * 32) << 16) + acc * BLOCK_N;
      if (has_cols && !(p.debug_mode & 1)) {
        float gn_stash[16];            // per-thread statistics sums of the first chunk of a pair (see gn_pair_finish)
        float* gn_stash_dst = nullptr;
        auto process = [&](const int c, uint4 (&r16)[4]) {
          uint32_t raw[32];
          if (p.debug_mode & 64) {
#pragma unroll
            for (int j = 0; j < 32; ++j) raw[j] = 0u;
          } else {
            tc_ld16(taddr + c, raw);
            if constexpr (CW == 32) tc_ld16(taddr + c + 16, raw + 16);
          }
          const int n0 = n_tile0 + c;
          const bool full = n0 + CW <= p.Cout;
          // bias: warp-uniform 16-byte loads (one L1 line per instruction), in flight together with the TMEM load
          float4 bq[CW / 4];
          if (p.bias != nullptr && full) {
            const float4* bp = reinterpret_cast<const float4*>(p.bias + n0);
#pragma unroll
            for (int j = 0; j < CW / 4; ++j) bq[j] = __ldg(bp + j);
          } else {
#pragma unroll
            for (int j = 0; j < CW / 4; ++j) {
              bq[j].x = (p.bias != nullptr && n0 + 4 * j < p.Cout) ? __ldg(p.bias + n0 + 4 * j) : 0.f;
              bq[j].y = (p.bias != nullptr && n0 + 4 * j + 1 < p.Cout) ? __ldg(p.bias + n0 + 4 * j + 1) : 0.f;
              bq[j].z = (p.bias != nullptr && n0 + 4 * j + 2 < p.Cout) ? __ldg(p.bias + n0 + 4 * j + 2) : 0.f;
              bq[j].w = (p.bias != nullptr && n0 + 4 * j + 3 < p.Cout) ? __ldg(p.bias + n0 + 4 * j + 3) : 0.f;
            }
          }
          tc_wait_ld();
          float v[32];
#pragma unroll
          for (int j = 0; j < CW / 4; ++j) {
            v[4 * j] = fmaf(__uint_as_float(raw[4 * j]), p.out_scale, bq[j].x);
            v[4 * j + 1] = fmaf(__uint_as_float(raw[4 * j + 1]), p.out_scale, bq[j].y);
            v[4 * j + 2] = fmaf(__uint_as_float(raw[4 * j + 2]), p.out_scale, bq[j].z);
            v[4 * j + 3] = fmaf(__uint_as_float(raw[4 * j + 3]), p.out_scale, bq[j].w);
          }
          if constexpr (CW == 16) {
#pragma unroll
            for (int j = 16; j < 32; ++j) v[j] = 0.f;
          }
          if (has_res && valid) {
            if (full && res16) {
              if (res_bf16) {
#pragma unroll
                for (int j = 0; j < CW / 8; ++j) {
                  const float2 a = T16<__nv_bfloat16>::to_f2(r16[j].x), b = T16<__nv_bfloat16>::to_f2(r16[j].y);
                  const float2 cc = T16<__nv_bfloat16>::to_f2(r16[j].z), d = T16<__nv_bfloat16>::to_f2(r16[j].w);
                  v[8 * j] += a.x; v[8 * j + 1] += a.y; v[8 * j + 2] += b.x; v[8 * j + 3] += b.y;
                  v[8 * j + 4] += cc.x; v[8 * j + 5] += cc.y; v[8 * j + 6] += d.x; v[8 * j + 7] += d.y;
                }
              } else {
#pragma unroll
                for (int j = 0; j < CW / 8; ++j) {
                  const float2 a = T16<__half>::to_f2(r16[j].x), b = T16<__half>::to_f2(r16[j].y);
                  const float2 cc = T16<__half>::to_f2(r16[j].z), d = T16<__half>::to_f2(r16[j].w);
                  v[8 * j] += a.x; v[8 * j + 1] += a.y; v[8 * j + 2] += b.x; v[8 * j + 3] += b.y;
                  v[8 * j + 4] += cc.x; v[8 * j + 5] += cc.y; v[8 * j + 6] += d.x; v[8 * j + 7] += d.y;
                }
              }
            } else {
              for (int j = 0; j < CW && n0 + j < p.Cout; ++j) {
                if (p.res_dtype == EOVAE_F32) {
                  v[j] += reinterpret_cast<const float*>(p.res)[pix * p.res_pix_stride + n0 + j];
                } else {
                  const uint32_t u = reinterpret_cast<const uint16_t*>(p.res)[pix * p.res_pix_stride + n0 + j];
                  v[j] += unpack16(u, p.res_dtype).x;
                }
              }
            }
          }
          if (n0 < p.Cout) {
            if (out16 && (use_tma_store || (valid && full))) {
              uint32_t pk[CW / 2];
              if (out_bf16) {
#pragma unroll
                for (int j = 0; j < CW / 2; ++j) pk[j] = T16<__nv_bfloat16>::from_f2(v[2 * j], v[2 * j + 1]);
              } else {
#pragma unroll
                for (int j = 0; j < CW / 2; ++j) pk[j] = T16<__half>::from_f2(v[2 * j], v[2 * j + 1]);
              }
              if (use_tma_store && wide_store) {
                if constexpr (CW == 32 && HALF_N == 64) {
                  // the warp's 64 columns leave as ONE 32-row x 128-byte box (128-byte swizzle: 16-byte piece j of row r at
                  // piece j ^ (r & 7)): the bulk-store engine's cost is per ROW, so 128-byte rows halve it (measured on
                  // the 16 -> 128 input conv, which is bound by it)
                  const int q = (c - col_begin) / CW;  // which 64-byte half of the rows this chunk fills
                  if (q == 0) {
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // previous tile's store drained
                    __syncwarp();
                  }
                  const uint32_t rbase = smem_u32(stage_buf) + lane * 128;
                  const int sw = lane & 7;
                  if (res_tma) {
                    // this lane's residual row pieces sit where its output pieces go (same box, same swizzle): add in
                    // fp32, round once, write back in place
                    if (q == 0) mbar_wait(&res_bar[ew], res_phase);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                      uint32_t rr[4];
                      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                   : "=r"(rr[0]), "=r"(rr[1]), "=r"(rr[2]), "=r"(rr[3])
                                   : "r"(rbase + (((q * 4 + j) ^ sw) << 4))
                                   : "memory");
#pragma unroll
                      for (int e = 0; e < 4; ++e) {
                        const float2 f = p.res_dtype == EOVAE_BF16 ? T16<__nv_bfloat16>::to_f2(rr[e]) : T16<__half>::to_f2(rr[e]);
                        v[8 * j + 2 * e] += f.x;
                        v[8 * j + 2 * e + 1] += f.y;
                      }
                    }
                    if (out_bf16) {
#pragma unroll
                      for (int j = 0; j < CW / 2; ++j) pk[j] = T16<__nv_bfloat16>::from_f2(v[2 * j], v[2 * j + 1]);
                    } else {
#pragma unroll
                      for (int j = 0; j < CW / 2; ++j) pk[j] = T16<__half>::from_f2(v[2 * j], v[2 * j + 1]);
                    }
                  }
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rbase + (((q * 4 + j) ^ sw) << 4)),
                                 "r"(pk[4 * j]), "r"(pk[4 * j + 1]), "r"(pk[4 * j + 2]), "r"(pk[4 * j + 3])
                                 : "memory");
                  if (q == 1) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0 && mt < m_tiles && sub * 32 < box_pix && !(p.debug_mode & 32)) {
                      asm volatile(
                          "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                              reinterpret_cast<uint64_t>(omap)),
                          "r"(smem_u32(stage_buf)), "r"(n_tile0 + col_begin), "r"(tw * p.box_w + w_wi),
                          "r"(th * p.box_h + w_hi), "r"(tn * p.box_n + w_ni)
                          : "memory");
                    }
                    if (lane == 0) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                  }
                }
              } else if (use_tma_store) {
                if constexpr (CW == 32) {
                  // 32 rows x 64 bytes, 64-byte swizzle (16-byte piece j of row r lives at piece j ^ ((r >> 1) & 3)):
                  // conflict-free st.shared, then ONE bulk tensor store per warp and chunk - coalesced by the TMA
                  // unit, clipped at the tensor bounds (ragged tiles, channel tails), asynchronous
                  uint8_t* buf = stage_buf + sbuf * EPI_BUF_BYTES;
                  if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");  // buffer's last store drained
                  __syncwarp();
                  const uint32_t rbase = smem_u32(buf) + lane * 64;
                  const int sw = (lane >> 1) & 3;
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rbase + ((j ^ sw) << 4)), "r"(pk[4 * j]),
                                 "r"(pk[4 * j + 1]), "r"(pk[4 * j + 2]), "r"(pk[4 * j + 3])
                                 : "memory");
                  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                  __syncwarp();
                  if (lane == 0 && mt < m_tiles && sub * 32 < box_pix) {  // warps past the box hold no pixels
                    asm volatile(
                        "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                            reinterpret_cast<uint64_t>(omap)),
                        "r"(smem_u32(buf)), "r"(n0), "r"(tw * p.box_w + w_wi), "r"(th * p.box_h + w_hi),
                        "r"(tn * p.box_n + w_ni)
                        : "memory");
                  }
                  if (lane == 0) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                  sbuf ^= 1;
                }
              } else {
                uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.out) + pix * p.out_pix_stride + n0);
#pragma unroll
                for (int j = 0; j < CW / 8; ++j) o[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
              }
            } else if (valid) {
              if (full) {  // fp32 output
                float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + pix * p.out_pix_stride + n0);
#pragma unroll
                for (int j = 0; j < CW / 4; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
              } else {
                for (int j = 0; j < CW && n0 + j < p.Cout; ++j) {
                  if (out16)
                    reinterpret_cast<uint16_t*>(p.out)[pix * p.out_pix_stride + n0 + j] =
                        static_cast<uint16_t>(pack16(v[j], 0.f, p.out_dtype) & 0xFFFF);
                  else
                    reinterpret_cast<float*>(p.out)[pix * p.out_pix_stride + n0 + j] = v[j];
                }
              }
            }
          }
          if constexpr (CW == 32) {
            if (p.gn_partial != nullptr && full && mt < m_tiles) {  // warp-uniform
              if (!valid) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0.f;
              }
              float* dst = p.gn_partial + ph * p.gn_phase_stride +
                           ((static_cast<long long>(mt) * 4 + sub) * p.gn_groups + (n0 >> p.gn_cpg_log2)) * 2;
              const bool lo_chunk = ((c - col_begin) & CW) == 0;
              if (lo_chunk && p.gn_cpg >= 4 && c + CW < col_begin + HALF_N && n0 + 2 * CW <= p.Cout) {
                // first chunk of a pair: keep the per-thread sums, the partner chunk reduces both
                switch (p.gn_cpg) {
                  case 4: gn_thread_sums<4>(v, gn_stash); break;
                  case 8: gn_thread_sums<8>(v, gn_stash); break;
                  case 16: gn_thread_sums<16>(v, gn_stash); break;
                  default: gn_thread_sums<32>(v, gn_stash); break;
                }
                gn_stash_dst = dst;
              } else if (!lo_chunk && gn_stash_dst != nullptr) {
                switch (p.gn_cpg) {
                  case 4: gn_pair_finish<4>(gn_stash, v, lane, gn_stash_dst); break;
                  case 8: gn_pair_finish<8>(gn_stash, v, lane, gn_stash_dst); break;
                  case 16: gn_pair_finish<16>(gn_stash, v, lane, gn_stash_dst); break;
                  default: gn_pair_finish<32>(gn_stash, v, lane, gn_stash_dst); break;
                }
                gn_stash_dst = nullptr;
              } else {
                switch (p.gn_cpg) {
                  case 1: gn_chunk<1>(v, lane, dst); break;
                  case 2: gn_chunk<2>(v, lane, dst); break;
                  case 4: gn_chunk<4>(v, lane, dst); break;
                  case 8: gn_chunk<8>(v, lane, dst); break;
                  case 16: gn_chunk<16>(v, lane, dst); break;
                  default: gn_chunk<32>(v, lane, dst); break;
                }
              }
            }
          }
          res_fetch(r16, c + 2 * CW);  // this buffer's next use is two chunks ahead
        };
#pragma unroll 1
        for (int c = col_begin; c < col_begin + HALF_N; c += 2 * CW) {
          process(c, rbuf_a);
          if (c + CW < col_begin + HALF_N) process(c + CW, rbuf_b);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&tmem_empty[acc]); else mbar_arrive_leader(&tmem_empty[acc]);
      }
      if (res_tma) res_phase ^= 1;
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all bulk stores of this warp complete
  }

  tc_fence_before();
  if constexpr (CTAS == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CTAS == 2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::TMEM_COLS) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::TMEM_COLS) : "memory");
  }
}

}  // namespace igemm
