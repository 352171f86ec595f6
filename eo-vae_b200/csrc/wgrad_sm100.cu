// Weight gradient of the 3x3 (stride 1, pad 1) and 1x1 convolutions on tcgen05:
//
//   dW[co][ci][kh][kw] = sum_{n, y, x} dY[n, y, x, co] * X[n, y + kh - 1, x + kw - 1, ci]
//
// i.e. per tap a GEMM  D[co, ci] = A[co, p] * B[ci, p]^T  contracted over the pixels p.  Both operands are read from
// CHANNEL-MAJOR copies (dY^T, X^T: [n][channels][h*w]) so that a 64-pixel K chunk of 128 output channels / up to 256 input
// channels is ONE 3-D TMA box (pixel, channel, image) - K-major, 128-byte swizzled.  The vertical tap shift of X is a box
// shifted by +-W pixels whose out-of-image part the TMA unit zero-fills (= the conv padding); the horizontal shift cannot be
// a TMA coordinate (the innermost start must stay 16-byte aligned - a one-element shift faults), so X^T comes as three
// x-shifted copies with the row borders already zeroed (eovae_transpose16_xshift3), stacked along the image dimension.
// Work item = (tap, 128-row Cout tile, Cin tile, K split); every CTA accumulates its pixel range in TMEM and writes one
// fp32 partial tile, a second kernel reduces the K splits in fixed order (deterministic) into OIHW fp32.
#include "../../include/eovae.h"
#include <cstdlib>

#include "igemm_sm100.cuh"

namespace {

using namespace igemm;

constexpr int WG_THREADS = 192;  // TMA warp, MMA warp, 4 epilogue warps
constexpr int WG_STAGES = 4;

struct WgradParams {
  CUtensorMap a_map;   // dY^T: dims (H*W, Cout, N),       box (64, 128, 1)
  CUtensorMap b_map;   // X^T : dims (H*W, Cin, 3N or N),  box (64, BN, 1)
  int H, W, N;
  int chunks_img;      // 64-pixel K chunks per image = ceil(H * W / 64); the ragged tail is TMA zero fill
  int taps;            // 9 or 1
  int cout, cin;
  int co_tiles, ci_tiles, ksplit;
  int chunks_total;    // N * H * W / 64
  int chunks_per_split;
  float* partial;      // [ksplit][taps][cout][cin]
  uint32_t idesc;
};

template <int BN>
__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_kernel(const __grid_constant__ WgradParams p) {
  constexpr int A_BYTES = 128 * 128;
  constexpr int B_BYTES = BN * 128;
  constexpr int STAGE = A_BYTES + B_BYTES;
  constexpr int TMEM_COLS = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WG_STAGES * STAGE);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + WG_STAGES;
  uint64_t* done_bar = bars + 2 * WG_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * WG_STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // work item decode
  int item = blockIdx.x;
  const int ks = item % p.ksplit; item /= p.ksplit;
  const int cit = item % p.ci_tiles; item /= p.ci_tiles;
  const int cot = item % p.co_tiles; item /= p.co_tiles;
  const int tap = item;
  const int dy = p.taps == 9 ? tap / 3 - 1 : 0;
  const int dx = p.taps == 9 ? tap % 3 - 1 : -1;  // 1x1: a single unshifted copy (image index (dx + 1) * N + n = n)
  const int chunk0 = ks * p.chunks_per_split;
  int chunk1 = chunk0 + p.chunks_per_split;
  if (chunk1 > p.chunks_total) chunk1 = p.chunks_total;
  const int nchunks = chunk1 > chunk0 ? chunk1 - chunk0 : 0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.a_map);
    prefetch_tmap(&p.b_map);
    for (int i = 0; i < WG_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int c = 0; c < nchunks; ++c) {
        const int chunk = chunk0 + c;
        const int p0 = (chunk % p.chunks_img) * 64;
        const int n = chunk / p.chunks_img;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], A_BYTES + B_BYTES);
        tma_load_3d(&p.a_map, &full_bar[stage], smem + stage * STAGE, p0, cot * 128, n);
        tma_load_3d(&p.b_map, &full_bar[stage], smem + stage * STAGE + A_BYTES, p0 + dy * p.W, cit * BN, (dx + 1) * p.N + n);
        if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int c = 0; c < nchunks; ++c) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint64_t da = make_smem_desc<128>(smem_u32(smem + stage * STAGE));
        const uint64_t db = make_smem_desc<128>(smem_u32(smem + stage * STAGE + A_BYTES));
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma_f16(tmem_base, da + static_cast<uint64_t>(k * 2), db + static_cast<uint64_t>(k * 2), p.idesc, (c | k) != 0 ? 1u : 0u);
        tc_commit(&empty_bar[stage]);
        if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
      }
      tc_commit(done_bar);
    }
  } else {
    const int sub = warp & 3;
    const int co = cot * 128 + sub * 32 + lane;
    mbar_wait_long(done_bar, 0);
    tc_fence_after();
    float* dst = p.partial + ((static_cast<long long>(ks) * p.taps + tap) * p.cout + co) * p.cin + cit * BN;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(sub * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < BN; c += 16) {
      uint32_t raw[16];
      tc_ld16(taddr + c, raw);
      tc_wait_ld();
      if (co < p.cout && nchunks > 0) {
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          if (cit * BN + c + j < p.cin)
            *reinterpret_cast<float4*>(dst + c + j) = make_float4(__uint_as_float(raw[j]), __uint_as_float(raw[j + 1]),
                                                                  __uint_as_float(raw[j + 2]), __uint_as_float(raw[j + 3]));
      } else if (co < p.cout) {
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          if (cit * BN + c + j < p.cin) *reinterpret_cast<float4*>(dst + c + j) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---- NHWC variant: both operands are read straight from the pixel-major activations / gradients as MN-MAJOR UMMA
// operands (the contracted index - the pixel - is the slow one in memory, exactly what MN-major means), so no transposed
// copies exist at all.  A K chunk = a (bw x bh) = 64-pixel box of 64 channels = one 4-D TMA box, 128-byte swizzled:
// shared memory holds [64 pixels][128 bytes], the canonical MN-major SW128 atom stack (8 pixel rows = 1024 bytes per
// K group -> stride byte offset 1024; the next 64 channels are the next box -> leading byte offset 8192).  The tap shift
// is a shifted box in (x, y), zero-filled outside the image by the TMA unit like in the forward conv.
struct WgradNhwcParams {
  CUtensorMap a_map;   // dY: dims (Cout, W, H, N), box (64, bw, bh, 1)
  CUtensorMap b_map;   // X : dims (Cin,  W, H, N), box (64, bw, bh, 1)
  int H, W, N;
  int bw, bh;
  int taps;
  int cout, cin;
  int co_tiles, ci_tiles, ksplit;
  int chunks_total, chunks_per_split;
  float* partial;
  uint32_t idesc;
  signed char tap_dx[16], tap_dy[16];   // pixel shift of X for every tap (3x3: -1..1; 1x1: 0; sub-pixel upsample: see host)
};

__device__ __forceinline__ uint64_t make_smem_desc_mn128(uint32_t smem_addr) {
  uint64_t d = static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);   // start address
  d |= static_cast<uint64_t>(8192 >> 4) << 16;                       // leading byte offset: next 64-channel atom column
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                       // stride byte offset: next group of 8 pixels (K)
  d |= static_cast<uint64_t>(1) << 46;                               // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;                               // SWIZZLE_128B
  return d;
}

// TPI = taps per work item: narrow-Cin layers put TPI taps side by side in the MMA's N dimension (N = TPI * BN <= 256), so
// the dY tile is loaded once per TPI taps and every instruction runs at the N = 256 operand-traffic ratio.
template <int BN, int TPI>
__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_nhwc_kernel(const __grid_constant__ WgradNhwcParams p) {
  constexpr int A_BYTES = 128 * 128;
  constexpr int NT = BN * TPI;
  constexpr int B_BYTES = NT * 128;
  constexpr int STAGE = A_BYTES + B_BYTES;
  constexpr int TMEM_COLS = NT <= 64 ? 64 : (NT <= 128 ? 128 : 256);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WG_STAGES * STAGE);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + WG_STAGES;
  uint64_t* done_bar = bars + 2 * WG_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * WG_STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // item order: taps fastest, K split slowest - the CTAs resident at the same time read the SAME pixel range (all taps
  // and channel tiles of it), so each operand byte comes from DRAM once and from L2 otherwise
  int item = blockIdx.x;
  const int groups = (p.taps + TPI - 1) / TPI;
  const int tap0 = (item % groups) * TPI; item /= groups;
  const int cit = item % p.ci_tiles; item /= p.ci_tiles;
  const int cot = item % p.co_tiles; item /= p.co_tiles;
  const int ks = item;
  const int ntaps = p.taps - tap0 < TPI ? p.taps - tap0 : TPI;
  const int chunk0 = ks * p.chunks_per_split;
  int chunk1 = chunk0 + p.chunks_per_split;
  if (chunk1 > p.chunks_total) chunk1 = p.chunks_total;
  const int nchunks = chunk1 > chunk0 ? chunk1 - chunk0 : 0;
  const int chunks_w = p.W / p.bw, chunks_h = p.H / p.bh;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.a_map);
    prefetch_tmap(&p.b_map);
    for (int i = 0; i < WG_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // chunk -> (x box, y box, image): decoded once, then advanced by increments (three runtime divisions per chunk in
      // this single thread were a third of its per-chunk latency budget)
      int bx = chunk0 % chunks_w, by = (chunk0 / chunks_w) % chunks_h, n = chunk0 / (chunks_w * chunks_h);
      for (int c = 0; c < nchunks; ++c) {
        const int x0 = bx * p.bw;
        const int y0 = by * p.bh;
        const int n_cur = n;
        if (++bx == chunks_w) { bx = 0; if (++by == chunks_h) { by = 0; ++n; } }
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], A_BYTES + ntaps * BN * 128);
        uint8_t* sa = smem + stage * STAGE;
#pragma unroll
        for (int a = 0; a < 2; ++a) tma_load_4d(&p.a_map, &full_bar[stage], sa + a * 8192, cot * 128 + a * 64, x0, y0, n_cur);
        for (int j = 0; j < ntaps; ++j) {
          const int tap = tap0 + j;
          const int dy = p.tap_dy[tap], dx = p.tap_dx[tap];
#pragma unroll
          for (int b = 0; b < BN / 64; ++b)
            tma_load_4d(&p.b_map, &full_bar[stage], sa + A_BYTES + (j * (BN / 64) + b) * 8192, cit * BN + b * 64, x0 + dx, y0 + dy, n_cur);
        }
        if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int c = 0; c < nchunks; ++c) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * STAGE);
#pragma unroll
        for (int k = 0; k < 4; ++k)  // 16 pixels (two 8-pixel K groups = 2048 bytes) per MMA
          tc_mma_f16(tmem_base, make_smem_desc_mn128(sa + k * 2048), make_smem_desc_mn128(sa + A_BYTES + k * 2048), p.idesc,
                     (c | k) != 0 ? 1u : 0u);
        tc_commit(&empty_bar[stage]);
        if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
      }
      tc_commit(done_bar);
    }
  } else {
    const int sub = warp & 3;
    const int co = cot * 128 + sub * 32 + lane;
    mbar_wait_long(done_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(sub * 32) << 16);
#pragma unroll 1
    for (int cc = 0; cc < NT; cc += 64) {  // BN is a multiple of 64: a 64-column step stays inside one tap slot
      const int jt = cc / BN, c = cc % BN;  // tap slot, channel inside the Cin tile
      if (jt >= ntaps) break;
      uint32_t raw[64];
#pragma unroll
      for (int q = 0; q < 4; ++q) tc_ld16(taddr + cc + 16 * q, raw + 16 * q);  // four TMEM loads in flight per wait
      tc_wait_ld();
      if (co >= p.cout) continue;
      float* dst = p.partial + ((static_cast<long long>(ks) * p.taps + tap0 + jt) * p.cout + co) * p.cin + cit * BN;
#pragma unroll
      for (int j = 0; j < 64; j += 4)
        if (cit * BN + c + j < p.cin)
          *reinterpret_cast<float4*>(dst + c + j) =
              nchunks > 0 ? make_float4(__uint_as_float(raw[j]), __uint_as_float(raw[j + 1]), __uint_as_float(raw[j + 2]),
                                        __uint_as_float(raw[j + 3]))
                          : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---- CTA-pair variant for the wide layers (Cout % 256 == 0, Cin % 256 == 0): one tcgen05.mma.cta_group::2 instruction
// covers 256 Cout rows x 256 Cin columns; CTA r of the pair stages ITS 128 dY channels and HALF of the X tile (128 of the
// 256 Cin columns), so an instruction reads 8 KB of shared memory per CTA instead of 12 KB (the 1-CTA 128x256x16 MMA is
// operand-bandwidth bound at ~187 cycles; the pair runs at the ~128-cycle tensor rate).  Barrier protocol = the forward
// igemm's: all TMA bytes of both CTAs complete on the LEADER's full barrier, the leader issues the MMAs and multicasts
// the commit to both CTAs' empty barriers.
constexpr int WG2_STAGES = 3;
constexpr int WG2_KC = 2;      // 64-pixel K chunks per pipeline stage (the single producer / issuer threads pay ~300 cycles per stage)
__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_nhwc2_kernel(const __grid_constant__ WgradNhwcParams p) {
  constexpr int A_BYTES = 128 * 128;
  constexpr int B_BYTES = 128 * 128;   // this CTA's half of the 256 Cin columns
  constexpr int CHUNK = A_BYTES + B_BYTES;
  constexpr int STAGE = WG2_KC * CHUNK;
  constexpr int TMEM_COLS = 256;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WG2_STAGES * STAGE);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + WG2_STAGES;
  uint64_t* done_bar = bars + 2 * WG2_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * WG2_STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  int item = blockIdx.x >> 1;
  const int tap = item % p.taps; item /= p.taps;
  const int cit = item % p.ci_tiles; item /= p.ci_tiles;
  const int cot2 = item % p.co_tiles; item /= p.co_tiles;   // here co_tiles counts 256-row tiles
  const int ks = item;
  const int dy = p.tap_dy[tap], dx = p.tap_dx[tap];
  const int chunk0 = ks * p.chunks_per_split;
  int chunk1 = chunk0 + p.chunks_per_split;
  if (chunk1 > p.chunks_total) chunk1 = p.chunks_total;
  const int nchunks = chunk1 > chunk0 ? chunk1 - chunk0 : 0;
  const int chunks_w = p.W / p.bw, chunks_h = p.H / p.bh;
  const int co_base = cot2 * 256 + static_cast<int>(rank) * 128;
  const int ci_base = cit * 256 + static_cast<int>(rank) * 128;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.a_map);
    prefetch_tmap(&p.b_map);
    for (int i = 0; i < WG2_STAGES; ++i) {
      mbar_init(&full_bar[i], 2);    // one producer arrive per CTA of the pair (the leader's copy is the live one)
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // chunk -> (x box, y box, image): decoded once, then advanced by increments (no division per chunk in this thread)
      int bx = chunk0 % chunks_w, by = (chunk0 / chunks_w) % chunks_h, bn = chunk0 / (chunks_w * chunks_h);
      for (int c = 0; c < nchunks; c += WG2_KC) {
        const int nck = nchunks - c < WG2_KC ? nchunks - c : WG2_KC;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (leader) mbar_expect_tx(&full_bar[stage], 2 * nck * CHUNK); else mbar_arrive_leader(&full_bar[stage]);
        for (int q = 0; q < nck; ++q) {
          const int x0 = bx * p.bw;
          const int y0 = by * p.bh;
          const int n = bn;
          if (++bx == chunks_w) { bx = 0; if (++by == chunks_h) { by = 0; ++bn; } }
          uint8_t* sa = smem + stage * STAGE + q * CHUNK;
#pragma unroll
          for (int a = 0; a < 2; ++a) tma2_load_4d(&p.a_map, &full_bar[stage], sa + a * 8192, co_base + a * 64, x0, y0, n);
#pragma unroll
          for (int b = 0; b < 2; ++b)
            tma2_load_4d(&p.b_map, &full_bar[stage], sa + A_BYTES + b * 8192, ci_base + b * 64, x0 + dx, y0 + dy, n);
        }
        if (++stage == WG2_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      int stage = 0;
      uint32_t phase = 0;
      for (int c = 0; c < nchunks; c += WG2_KC) {
        const int nck = nchunks - c < WG2_KC ? nchunks - c : WG2_KC;
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        for (int q = 0; q < nck; ++q) {
          const uint32_t sa = smem_u32(smem + stage * STAGE + q * CHUNK);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc2_mma_f16(tmem_base, make_smem_desc_mn128(sa + k * 2048), make_smem_desc_mn128(sa + A_BYTES + k * 2048), p.idesc,
                        (c | q | k) != 0 ? 1u : 0u);
        }
        tc2_commit_mc(&empty_bar[stage]);
        if (++stage == WG2_STAGES) { stage = 0; phase ^= 1; }
      }
      tc2_commit_mc(done_bar);
    }
  } else {
    const int sub = warp & 3;
    const int co = co_base + sub * 32 + lane;
    if (nchunks > 0) mbar_wait_long(done_bar, 0);
    tc_fence_after();
    float* dst = p.partial + ((static_cast<long long>(ks) * p.taps + tap) * p.cout + co) * p.cin + cit * 256;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(sub * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < 256; c += 64) {  // four TMEM loads in flight per wait (one per wait cost 16 load latencies per item)
      uint32_t raw[64];
      if (nchunks > 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) tc_ld16(taddr + c + 16 * q, raw + 16 * q);
        tc_wait_ld();
      }
#pragma unroll
      for (int j = 0; j < 64; j += 4)
        *reinterpret_cast<float4*>(dst + c + j) =
            nchunks > 0 ? make_float4(__uint_as_float(raw[j]), __uint_as_float(raw[j + 1]), __uint_as_float(raw[j + 2]),
                                      __uint_as_float(raw[j + 3]))
                        : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// dW[co][ci][tap] (OIHW, taps innermost) (+)= sum_ks partial[ks][tap][co][ci]
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int ksplit, int taps, int cout,
                                    int cin, int accumulate, long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;  // over (tap, co, ci), ci fastest
  if (i >= total) return;
  const int ci = static_cast<int>(i % cin);
  long long t = i / cin;
  const int co = static_cast<int>(t % cout);
  const int tap = static_cast<int>(t / cout);
  double acc = 0.0;
#pragma unroll 8
  for (int k = 0; k < ksplit; ++k) acc += partial[static_cast<long long>(k) * total + i];  // unrolled: loads issued in batches
  float* o = dw + (static_cast<long long>(co) * cin + ci) * taps + tap;
  *o = (accumulate ? *o : 0.f) + static_cast<float>(acc);
}

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn encode_fn() {
  static EncodeFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(ptr);
  }
  return fn;
}

int make_map(CUtensorMap* m, int dtype, const void* base, long long pixels, int c, int n, int rows) {
  EncodeFn fn = encode_fn();
  EOVAE_CHECK(fn != nullptr, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(pixels), static_cast<cuuint64_t>(c), static_cast<cuuint64_t>(n)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(pixels) * 2, static_cast<cuuint64_t>(pixels) * c * 2};
  cuuint32_t box[3] = {64, static_cast<cuuint32_t>(rows), 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = fn(m, dtype == EOVAE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3,
                  const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EOVAE_CHECK(r == CUDA_SUCCESS, "wgrad: cuTensorMapEncodeTiled failed (%d) pixels %lld c %d n %d rows %d", (int)r, pixels, c, n, rows);
  return 0;
}

// K split: `base` independent tiles, `chunks` 64-pixel K chunks each, `slots` CTAs (or CTA pairs) resident at once.
// cost(ks) = waves(ks) x (chunks per item + fixed per-item cost): an item pays ~4 us of TMEM allocation, barrier setup,
// pipeline fill and fp32 partial-tile write-out, about the MMA time of 16 chunks; whole waves only (1 CTA per SM).
int pick_ksplit(int base, int chunks, int slots) {
  int best = 1;
  long long best_cost = -1;
  const int max_ks = chunks / 8 > 1 ? chunks / 8 : 1;
  for (int ks = 1; ks <= max_ks && ks <= 4096; ++ks) {
    const long long waves = ceil_div(base * ks, slots);
    const long long cost = waves * (ceil_div(chunks, ks) + 16);
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best = ks;
    }
  }
  return best;
}

int pick_bn(int cin) {
  if (cin >= 256) return 256;
  if (cin >= 128) return 128;
  if (cin >= 64) return 64;
  if (cin >= 32) return 32;
  return 16;
}

void plan(int n, int h, int w, int cin, int cout, int taps, int* bn, int* co_tiles, int* ci_tiles, int* ksplit, int* cps,
          int* chunks_total) {
  *bn = pick_bn(cin);
  *co_tiles = ceil_div(cout, 128);
  *ci_tiles = ceil_div(cin, *bn);
  *chunks_total = n * ceil_div(h * w, 64);
  const int base = taps * *co_tiles * *ci_tiles;
  const int ks = pick_ksplit(base, *chunks_total, eovae_num_sms());
  *cps = ceil_div(*chunks_total, ks);
  *ksplit = ceil_div(*chunks_total, *cps);
}

template <int BN>
int launch(const WgradParams& p, int items, cudaStream_t stream) {
  constexpr int SMEM = WG_STAGES * (128 * 128 + BN * 128) + 1024 + 256;
  auto kern = wgrad_kernel<BN>;
  static bool set = false;
  if (!set) {
    EOVAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    set = true;
  }
  kern<<<items, WG_THREADS, SMEM, stream>>>(p);
  EOVAE_LAUNCH_CHECK();
  return 0;
}


// lattice > 1: the map covers the sub-lattice base[:, ::lattice, ::lattice, :] of an image of (lattice*h) x (lattice*w) pixels
int make_map_nhwc(CUtensorMap* m, int dtype, const void* base, int c, long long pitch, int w, int h, int n, int bw, int bh,
                  int lattice = 1) {
  EncodeFn fn = encode_fn();
  EOVAE_CHECK(fn != nullptr, "cuTensorMapEncodeTiled entry point not available");
  const cuuint64_t L = static_cast<cuuint64_t>(lattice);
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(c), static_cast<cuuint64_t>(w), static_cast<cuuint64_t>(h), static_cast<cuuint64_t>(n)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(pitch) * 2 * L, static_cast<cuuint64_t>(pitch) * 2 * (L * w) * L,
                           static_cast<cuuint64_t>(pitch) * 2 * (L * w) * (L * h)};
  cuuint32_t box[4] = {64, static_cast<cuuint32_t>(bw), static_cast<cuuint32_t>(bh), 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = fn(m, dtype == EOVAE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4,
                  const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EOVAE_CHECK(r == CUDA_SUCCESS, "wgrad: cuTensorMapEncodeTiled (NHWC) failed (%d) c %d pitch %lld w %d h %d n %d box %d %d", (int)r,
              c, pitch, w, h, n, bw, bh);
  return 0;
}

bool g_no_pair = false;  // EOVAE_WGRAD_NO_PAIR=1: keep the 1-CTA kernel (A/B measurements)

int pick_bn_nhwc(int cin) { return cin > 128 ? 256 : (cin > 64 ? 128 : 64); }

int pick_tpi(int bn, int taps) { return taps == 1 ? 1 : 256 / bn; }

void plan_nhwc(int n, int h, int w, int cin, int cout, int taps, int* bn, int* co_tiles, int* ci_tiles, int* ksplit, int* cps,
               int* chunks_total) {
  *bn = pick_bn_nhwc(cin);
  *co_tiles = ceil_div(cout, 128);
  *ci_tiles = ceil_div(cin, *bn);
  *chunks_total = static_cast<int>(static_cast<long long>(n) * h * w / 64);
  const int base = ceil_div(taps, pick_tpi(*bn, taps)) * *co_tiles * *ci_tiles;
  const int ks = pick_ksplit(base, *chunks_total, eovae_num_sms());
  *cps = ceil_div(*chunks_total, ks);
  *ksplit = ceil_div(*chunks_total, *cps);
}

template <int BN, int TPI>
int launch_nhwc(const WgradNhwcParams& p, int items, cudaStream_t stream) {
  constexpr int SMEM = WG_STAGES * (128 * 128 + BN * TPI * 128) + 1024 + 256;
  auto kern = wgrad_nhwc_kernel<BN, TPI>;
  static bool set = false;
  if (!set) {
    EOVAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    set = true;
  }
  kern<<<items, WG_THREADS, SMEM, stream>>>(p);
  EOVAE_LAUNCH_CHECK();
  return 0;
}

}  // namespace

extern "C" {

int eovae_conv2d_wgrad_nhwc_ok(int h, int w) {
  const int bw = w >= 64 ? 64 : w;
  if (bw <= 0 || 64 % bw != 0 || w % bw != 0) return 0;
  return h % (64 / bw) == 0 ? 1 : 0;
}

size_t eovae_conv2d_wgrad_nhwc_workspace_bytes(int n, int h, int w, int cin, int cout, int ksize) {
  int bn, cot, cit, ks, cps, ct;
  plan_nhwc(n, h, w, cin, cout, ksize * ksize, &bn, &cot, &cit, &ks, &cps, &ct);
  if (cout % 256 == 0 && cin % 256 == 0) {  // CTA-pair plan (see eovae_conv2d_wgrad_nhwc)
    const int base = ksize * ksize * (cout / 256) * (cin / 256);
    int ks2 = pick_ksplit(base, ct, eovae_num_sms() / 2);
    const int cps2 = ceil_div(ct, ks2);
    ks2 = ceil_div(ct, cps2);
    if (ks2 > ks) ks = ks2;
  }
  return sizeof(float) * static_cast<size_t>(ks) * ksize * ksize * cout * cin;
}

}  // extern "C"

namespace {

// One weight-gradient launch: partial[ks][tap][cout][cin] then the fixed-order split-K reduction into dw_out laid out
// [cout][cin][taps] (taps innermost = OIHW for a k x k kernel).  ``taps`` / offsets are free (3x3, 1x1, or the 2x2 taps of one
// sub-pixel phase of the upsample conv); ``dy_lattice`` = 2 reads dY on the parity sub-lattice starting at ``dy``.
int wgrad_nhwc_launch(const void* x, long long x_pix_stride, int x_lattice, const void* dy, long long dy_pix_stride, int dy_lattice,
                      int dtype, int dy_dtype, int n, int h, int w, int cin, int cout, int taps, const signed char* tap_dx,
                      const signed char* tap_dy, float* dw_out, int accumulate, void* workspace, size_t workspace_bytes,
                      cudaStream_t stream) {
  static bool env_read = false;
  if (!env_read) {
    const char* e = getenv("EOVAE_WGRAD_NO_PAIR");
    g_no_pair = e != nullptr && e[0] == '1';
    env_read = true;
  }
  WgradNhwcParams p;
  memset(&p, 0, sizeof(p));
  int bn;
  plan_nhwc(n, h, w, cin, cout, taps, &bn, &p.co_tiles, &p.ci_tiles, &p.ksplit, &p.chunks_per_split, &p.chunks_total);
  p.H = h; p.W = w; p.N = n; p.taps = taps; p.cout = cout; p.cin = cin;
  for (int t = 0; t < taps; ++t) {
    p.tap_dx[t] = tap_dx[t];
    p.tap_dy[t] = tap_dy[t];
  }
  p.bw = w >= 64 ? 64 : w;
  p.bh = 64 / p.bw;
  p.partial = static_cast<float*>(workspace);
  const uint32_t fmt_a = dy_dtype == EOVAE_BF16 ? 1u : 0u;  // A operand = dY, B operand = X
  const uint32_t fmt_b = dtype == EOVAE_BF16 ? 1u : 0u;
  if (cout % 256 == 0 && cin % 256 == 0 && !g_no_pair) {
    // CTA-pair kernel: co_tiles counts 256-row tiles; same split-K plan (the partial layout does not depend on the tiling)
    p.co_tiles = cout / 256;
    p.ci_tiles = cin / 256;
    const int base = p.taps * p.co_tiles * p.ci_tiles;
    const int ks = pick_ksplit(base, p.chunks_total, eovae_num_sms() / 2);  // one CTA pair per two SMs
    p.chunks_per_split = ceil_div(p.chunks_total, ks);
    p.ksplit = ceil_div(p.chunks_total, p.chunks_per_split);
    EOVAE_CHECK(workspace_bytes >= sizeof(float) * static_cast<size_t>(p.ksplit) * p.taps * cout * cin, "conv2d_wgrad_nhwc: workspace too small");
    p.idesc = (1u << 4) | (fmt_a << 7) | (fmt_b << 10) | (1u << 15) | (1u << 16) | (static_cast<uint32_t>(256 >> 3) << 17) |
              (static_cast<uint32_t>(256 >> 4) << 24);
    if (make_map_nhwc(&p.a_map, dy_dtype, dy, cout, dy_pix_stride, w, h, n, p.bw, p.bh, dy_lattice)) return -3;
    if (make_map_nhwc(&p.b_map, dtype, x, cin, x_pix_stride, w, h, n, p.bw, p.bh, x_lattice)) return -3;
    constexpr int SMEM2 = WG2_STAGES * WG2_KC * (2 * 128 * 128) + 1024 + 256;
    static bool set2 = false;
    if (!set2) {
      EOVAE_CUDA(cudaFuncSetAttribute(wgrad_nhwc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2));
      set2 = true;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(static_cast<unsigned>(2 * base * p.ksplit));
    cfg.blockDim = dim3(WG_THREADS);
    cfg.dynamicSmemBytes = SMEM2;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    EOVAE_CUDA(cudaLaunchKernelEx(&cfg, wgrad_nhwc2_kernel, p));
    EOVAE_LAUNCH_CHECK();
    const long long total2 = static_cast<long long>(p.taps) * cout * cin;
    wgrad_reduce_kernel<<<static_cast<unsigned>((total2 + 255) / 256), 256, 0, stream>>>(p.partial, dw_out, p.ksplit, p.taps, cout,
                                                                                        cin, accumulate, total2);
    EOVAE_LAUNCH_CHECK();
    return 0;
  }
  const int tpi = pick_tpi(bn, p.taps);
  p.idesc = (1u << 4) | (fmt_a << 7) | (fmt_b << 10) | (1u << 15) | (1u << 16) | (static_cast<uint32_t>((bn * tpi) >> 3) << 17) |
            (static_cast<uint32_t>(128 >> 4) << 24);  // bits 15 / 16: A and B are MN-major
  if (make_map_nhwc(&p.a_map, dy_dtype, dy, cout, dy_pix_stride, w, h, n, p.bw, p.bh, dy_lattice)) return -3;
  if (make_map_nhwc(&p.b_map, dtype, x, cin, x_pix_stride, w, h, n, p.bw, p.bh, x_lattice)) return -3;
  const int items = ceil_div(p.taps, tpi) * p.co_tiles * p.ci_tiles * p.ksplit;
  int rc;
  switch (bn * 8 + tpi) {
    case 256 * 8 + 1: rc = launch_nhwc<256, 1>(p, items, stream); break;
    case 128 * 8 + 2: rc = launch_nhwc<128, 2>(p, items, stream); break;
    case 128 * 8 + 1: rc = launch_nhwc<128, 1>(p, items, stream); break;
    case 64 * 8 + 4: rc = launch_nhwc<64, 4>(p, items, stream); break;
    default: rc = launch_nhwc<64, 1>(p, items, stream); break;
  }
  if (rc) return rc;
  const long long total = static_cast<long long>(p.taps) * cout * cin;
  wgrad_reduce_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(p.partial, dw_out, p.ksplit, p.taps, cout,
                                                                                      cin, accumulate, total);
  EOVAE_LAUNCH_CHECK();
  return 0;
}

// dW (3x3, OIHW) (+)= adjoint of the sub-pixel weight folding: dW[kh][kw] = sum over the phases (py, px) of
// dW'[py,px][a(py,kh)][b(px,kw)], with a(0, kh) = (kh >= 1), a(1, kh) = (kh == 2) (the inverse of up2x_in_set in igemm_host.cu)
__global__ void up2x_unfold_wgrad_kernel(const float* __restrict__ dwp /*[4][cout][cin][4]*/, float* __restrict__ dw, long long pairs,
                                         int accumulate) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;  // (co, ci)
  if (i >= pairs) return;
  float acc[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) acc[k] = 0.f;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 v = *reinterpret_cast<const float4*>(dwp + (q * pairs + i) * 4);
    const float t[4] = {v.x, v.y, v.z, v.w};
    const int py = q >> 1, px = q & 1;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int a = py == 0 ? (kh >= 1) : (kh == 2), b = px == 0 ? (kw >= 1) : (kw == 2);
        acc[kh * 3 + kw] += t[a * 2 + b];
      }
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) dw[i * 9 + k] = (accumulate ? dw[i * 9 + k] : 0.f) + acc[k];
}

// dW[co][ci][kh][kw] (+)= the per-parity-class gradients of the stride-2 conv: class q = (kh & 1) * 2 + (kw & 1) holds its taps
// in the order (kh >> 1) * (number of kw of the class) + (kw >> 1); classes start at offsets 0, 4, 6, 8 (x cout*cin) of dwp
__global__ void s2_gather_wgrad_kernel(const float* __restrict__ dwp, float* __restrict__ dw, long long pairs, int accumulate) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;  // (co, ci)
  if (i >= pairs) return;
  const int class_off[4] = {0, 4, 6, 8}, class_taps[4] = {4, 2, 2, 1}, class_kw[4] = {2, 1, 2, 1};
#pragma unroll
  for (int kh = 0; kh < 3; ++kh)
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const int q = (kh & 1) * 2 + (kw & 1);
      const int t = (kh >> 1) * class_kw[q] + (kw >> 1);
      const float v = dwp[class_off[q] * pairs + i * class_taps[q] + t];
      float* o = dw + i * 9 + kh * 3 + kw;
      *o = (accumulate ? *o : 0.f) + v;
    }
}

}  // namespace

extern "C" {

// Weight gradient of the Downsample conv (pad (0,1,0,1) + 3x3 stride 2): dW[kh][kw] = sum dy[i][j] x[2i + kh][2j + kw].  x is read
// on its four parity sub-lattices (one launch per class of taps with equal parity; the pad row / column is TMA zero fill),
// dy as it is: 9 tap units over the (h/2 x w/2) output pixels instead of 9 over the zero-interleaved (h x w) image.
size_t eovae_conv2d_s2_wgrad_workspace_bytes(int n, int ho, int wo, int cin, int cout) {
  int bn, cot, cit, ks, cps, ct;
  plan_nhwc(n, ho, wo, cin, cout, 4, &bn, &cot, &cit, &ks, &cps, &ct);
  if (cout % 256 == 0 && cin % 256 == 0) {
    for (int taps = 1; taps <= 4; ++taps) {
      const int base = taps * (cout / 256) * (cin / 256);
      int ks2 = pick_ksplit(base, ct, eovae_num_sms() / 2);
      const int cps2 = ceil_div(ct, ks2);
      ks2 = ceil_div(ct, cps2);
      if (ks2 > ks) ks = ks2;
    }
  }
  for (int taps = 1; taps <= 2; ++taps) {
    int ks1;
    plan_nhwc(n, ho, wo, cin, cout, taps, &bn, &cot, &cit, &ks1, &cps, &ct);
    if (ks1 > ks) ks = ks1;
  }
  return sizeof(float) * (static_cast<size_t>(ks) * 4 * cout * cin + 9 * static_cast<size_t>(cout) * cin);
}

int eovae_conv2d_s2_wgrad(const void* x, long long x_pix_stride, const void* dy, long long dy_pix_stride, int dtype, int n, int ho,
                          int wo, int cin, int cout, float* dw_oihw, int accumulate, void* workspace, size_t workspace_bytes,
                          void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  EOVAE_CHECK(dtype == EOVAE_BF16 || dtype == EOVAE_F16, "conv2d_s2_wgrad: 16-bit operands only");
  EOVAE_CHECK(eovae_conv2d_wgrad_nhwc_ok(ho, wo), "conv2d_s2_wgrad: %dx%d outputs do not tile into 64-pixel boxes", ho, wo);
  EOVAE_CHECK(cin % 4 == 0 && x_pix_stride % 8 == 0 && dy_pix_stride % 8 == 0, "conv2d_s2_wgrad: Cin %% 4 and 16-byte pixel pitches required");
  EOVAE_CHECK(workspace_bytes >= eovae_conv2d_s2_wgrad_workspace_bytes(n, ho, wo, cin, cout), "conv2d_s2_wgrad: workspace too small");
  float* class_dw = static_cast<float*>(workspace);                      // 9 * cout * cin floats: classes of 4, 2, 2, 1 taps
  const size_t pairs = static_cast<size_t>(cout) * cin;
  float* partial = class_dw + 9 * pairs;
  const size_t partial_bytes = workspace_bytes - sizeof(float) * 9 * pairs;
  const int class_off[4] = {0, 4, 6, 8};
  for (int q = 0; q < 4; ++q) {
    const int ph = q >> 1, pw = q & 1;
    const int nkh = ph == 0 ? 2 : 1, nkw = pw == 0 ? 2 : 1;
    signed char dx[16], dyv[16];
    for (int a = 0; a < nkh; ++a)
      for (int b = 0; b < nkw; ++b) {
        dyv[a * nkw + b] = static_cast<signed char>(a);   // x offset inside the lattice: kh >> 1
        dx[a * nkw + b] = static_cast<signed char>(b);
      }
    // lattice (ph, pw) of x (extent 2 ho x 2 wo): rows ph, ph + 2, ...; the map covers ho x wo lattice points, offset + 1 at the
    // last index is out of bounds = the zero pad of F.pad(0, 1, 0, 1)
    const uint8_t* xq = static_cast<const uint8_t*>(x) + (static_cast<size_t>(ph) * (2 * wo) + pw) * x_pix_stride * 2;
    int rc = wgrad_nhwc_launch(xq, x_pix_stride, 2, dy, dy_pix_stride, 1, dtype, dtype, n, ho, wo, cin, cout, nkh * nkw, dx, dyv,
                               class_dw + class_off[q] * pairs, 0, partial, partial_bytes, stream);
    if (rc) return rc;
  }
  s2_gather_wgrad_kernel<<<static_cast<unsigned>((pairs + 255) / 256), 256, 0, stream>>>(class_dw, dw_oihw, static_cast<long long>(pairs),
                                                                                        accumulate);
  EOVAE_LAUNCH_CHECK();
  return 0;
}

int eovae_conv2d_wgrad_nhwc(const void* x, long long x_pix_stride, const void* dy, long long dy_pix_stride, int dtype, int dy_dtype, int n, int h,
                            int w, int cin, int cout, int ksize, float* dw_oihw, int accumulate, void* workspace,
                            size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  EOVAE_CHECK(ksize == 3 || ksize == 1, "conv2d_wgrad_nhwc: kernel size must be 3 or 1");
  EOVAE_CHECK((dtype == EOVAE_BF16 || dtype == EOVAE_F16) && (dy_dtype == EOVAE_BF16 || dy_dtype == EOVAE_F16),
              "conv2d_wgrad_nhwc: 16-bit operands only");
  EOVAE_CHECK(dtype == dy_dtype || getenv("EOVAE_ALLOW_MIXED_MMA") != nullptr,
              "conv2d_wgrad_nhwc: x (%d) and dy (%d) formats must be equal (tcgen05 kind::f16 rejects mixed f16/bf16)", dtype, dy_dtype);
  EOVAE_CHECK(eovae_conv2d_wgrad_nhwc_ok(h, w), "conv2d_wgrad_nhwc: %dx%d images do not tile into 64-pixel boxes", h, w);
  EOVAE_CHECK(cin % 4 == 0 && x_pix_stride % 8 == 0 && dy_pix_stride % 8 == 0 && x_pix_stride >= cin && dy_pix_stride >= cout,
              "conv2d_wgrad_nhwc: Cin %% 4 and 16-byte pixel pitches required (Cin %d, pitches %lld / %lld)", cin, x_pix_stride,
              dy_pix_stride);
  EOVAE_CHECK(workspace_bytes >= eovae_conv2d_wgrad_nhwc_workspace_bytes(n, h, w, cin, cout, ksize), "conv2d_wgrad_nhwc: workspace too small");
  signed char dx[16], dyv[16];
  for (int t = 0; t < ksize * ksize; ++t) {
    dx[t] = ksize == 3 ? static_cast<signed char>(t % 3 - 1) : 0;
    dyv[t] = ksize == 3 ? static_cast<signed char>(t / 3 - 1) : 0;
  }
  return wgrad_nhwc_launch(x, x_pix_stride, 1, dy, dy_pix_stride, 1, dtype, dy_dtype, n, h, w, cin, cout, ksize * ksize, dx, dyv, dw_oihw,
                           accumulate, workspace, workspace_bytes, stream);
}

// Weight gradient of the sub-pixel upsample convolution (eovae_conv2d_up2x): x = its LOW-resolution input [n][h][w][cin], dy =
// the gradient of its high-resolution output [n][2h][2w][cout].  Four launches (one per phase, 2x2 taps, dY read on the
// phase's parity sub-lattice) = 16 instead of 36 MAC units, then the 3x3 gradient is unfolded from the four 2x2 ones.
// workspace: 4 * cout * cin * 4 floats (phase gradients) + the split-K partials of one launch.
size_t eovae_conv2d_up2x_wgrad_workspace_bytes(int n, int h, int w, int cin, int cout) {
  int bn, cot, cit, ks, cps, ct;
  plan_nhwc(n, h, w, cin, cout, 4, &bn, &cot, &cit, &ks, &cps, &ct);
  if (cout % 256 == 0 && cin % 256 == 0) {
    const int base = 4 * (cout / 256) * (cin / 256);
    int ks2 = pick_ksplit(base, ct, eovae_num_sms() / 2);
    const int cps2 = ceil_div(ct, ks2);
    ks2 = ceil_div(ct, cps2);
    if (ks2 > ks) ks = ks2;
  }
  return sizeof(float) * (static_cast<size_t>(ks) * 4 * cout * cin + 16 * static_cast<size_t>(cout) * cin);
}

int eovae_conv2d_up2x_wgrad(const void* x, long long x_pix_stride, const void* dy, long long dy_pix_stride, int dtype, int n, int h,
                            int w, int cin, int cout, float* dw_oihw, int accumulate, void* workspace, size_t workspace_bytes,
                            void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  EOVAE_CHECK(dtype == EOVAE_BF16 || dtype == EOVAE_F16, "conv2d_up2x_wgrad: 16-bit operands only");
  EOVAE_CHECK(eovae_conv2d_wgrad_nhwc_ok(h, w), "conv2d_up2x_wgrad: %dx%d images do not tile into 64-pixel boxes", h, w);
  EOVAE_CHECK(cin % 4 == 0 && x_pix_stride % 8 == 0 && dy_pix_stride % 8 == 0, "conv2d_up2x_wgrad: Cin %% 4 and 16-byte pixel pitches required");
  EOVAE_CHECK(workspace_bytes >= eovae_conv2d_up2x_wgrad_workspace_bytes(n, h, w, cin, cout), "conv2d_up2x_wgrad: workspace too small");
  float* phase_dw = static_cast<float*>(workspace);                       // [4][cout][cin][4]
  const size_t phase_elems = 4 * static_cast<size_t>(cout) * cin;
  float* partial = phase_dw + 4 * phase_elems;
  const size_t partial_bytes = workspace_bytes - sizeof(float) * 4 * phase_elems;
  for (int q = 0; q < 4; ++q) {
    const int py = q >> 1, px = q & 1;
    signed char dx[16], dyv[16];
    for (int t = 0; t < 4; ++t) {
      dyv[t] = static_cast<signed char>((t >> 1) - 1 + py);
      dx[t] = static_cast<signed char>((t & 1) - 1 + px);
    }
    const uint8_t* dyq = static_cast<const uint8_t*>(dy) + (static_cast<size_t>(py) * (2 * w) + px) * dy_pix_stride * 2;
    int rc = wgrad_nhwc_launch(x, x_pix_stride, 1, dyq, dy_pix_stride, 2, dtype, dtype, n, h, w, cin, cout, 4, dx, dyv,
                               phase_dw + q * phase_elems, 0, partial, partial_bytes, stream);
    if (rc) return rc;
  }
  const long long pairs = static_cast<long long>(cout) * cin;
  up2x_unfold_wgrad_kernel<<<static_cast<unsigned>((pairs + 255) / 256), 256, 0, stream>>>(phase_dw, dw_oihw, pairs, accumulate);
  EOVAE_LAUNCH_CHECK();
  return 0;
}

size_t eovae_conv2d_wgrad_workspace_bytes(int n, int h, int w, int cin, int cout, int ksize) {
  int bn, cot, cit, ks, cps, ct;
  plan(n, h, w, cin, cout, ksize * ksize, &bn, &cot, &cit, &ks, &cps, &ct);
  return sizeof(float) * static_cast<size_t>(ks) * ksize * ksize * cout * cin;
}

int eovae_conv2d_wgrad(const void* x_t, const void* dy_t, int dtype, int dy_dtype, int n, int h, int w, int cin, int cout, int ksize,
                       float* dw_oihw, int accumulate, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  EOVAE_CHECK(ksize == 3 || ksize == 1, "conv2d_wgrad: kernel size must be 3 or 1");
  EOVAE_CHECK((dtype == EOVAE_BF16 || dtype == EOVAE_F16) && (dy_dtype == EOVAE_BF16 || dy_dtype == EOVAE_F16),
              "conv2d_wgrad: 16-bit operands only");
  EOVAE_CHECK(dtype == dy_dtype || getenv("EOVAE_ALLOW_MIXED_MMA") != nullptr,
              "conv2d_wgrad: x (%d) and dy (%d) formats must be equal (tcgen05 kind::f16 rejects mixed f16/bf16)", dtype, dy_dtype);
  EOVAE_CHECK(w % 8 == 0 && cin % 16 == 0, "conv2d_wgrad: (padded) W %% 8 and Cin %% 16 required (W %d, Cin %d)", w, cin);
  EOVAE_CHECK(workspace_bytes >= eovae_conv2d_wgrad_workspace_bytes(n, h, w, cin, cout, ksize), "conv2d_wgrad: workspace too small");
  WgradParams p;
  memset(&p, 0, sizeof(p));
  int bn;
  plan(n, h, w, cin, cout, ksize * ksize, &bn, &p.co_tiles, &p.ci_tiles, &p.ksplit, &p.chunks_per_split, &p.chunks_total);
  p.H = h; p.W = w; p.N = n; p.chunks_img = ceil_div(h * w, 64); p.taps = ksize * ksize; p.cout = cout; p.cin = cin;
  p.partial = static_cast<float*>(workspace);
  const uint32_t fmt_a = dy_dtype == EOVAE_BF16 ? 1u : 0u;  // A operand = dY, B operand = X: kind::f16 takes the two
  const uint32_t fmt_b = dtype == EOVAE_BF16 ? 1u : 0u;     // operand formats independently (f16 x bf16 allowed)
  p.idesc = (1u << 4) | (fmt_a << 7) | (fmt_b << 10) | (static_cast<uint32_t>(bn >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
  if (make_map(&p.a_map, dy_dtype, dy_t, static_cast<long long>(h) * w, cout, n, 128)) return -3;
  if (make_map(&p.b_map, dtype, x_t, static_cast<long long>(h) * w, cin, ksize == 3 ? 3 * n : n, bn)) return -3;
  const int items = p.taps * p.co_tiles * p.ci_tiles * p.ksplit;
  int rc;
  switch (bn) {
    case 256: rc = launch<256>(p, items, stream); break;
    case 128: rc = launch<128>(p, items, stream); break;
    case 64: rc = launch<64>(p, items, stream); break;
    case 32: rc = launch<32>(p, items, stream); break;
    default: rc = launch<16>(p, items, stream); break;
  }
  if (rc) return rc;
  const long long total = static_cast<long long>(p.taps) * cout * cin;
  wgrad_reduce_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(p.partial, dw_oihw, p.ksplit, p.taps, cout,
                                                                                      cin, accumulate, total);
  EOVAE_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
