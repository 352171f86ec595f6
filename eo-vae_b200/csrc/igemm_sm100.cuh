// tcgen05 / TMEM / TMA implicit-GEMM kernel for sm_100a.
//
//   D[m, n] = sum_{tap, c} A_tap[m, c] * Wp[n, tap*Kc + c]      (+ bias[n]) (+ residual[m, n]) (* scale)
//
// m runs over output pixels (image n, row y, col x) of an NHWC activation tensor, A_tap is the same tensor
// shifted by the filter tap; every A tile is ONE 4-D TMA box (channels, x, y, image) whose out-of-bounds
// part (conv zero padding, ragged right/bottom edges, channel padding) is zero-filled by the TMA unit, so
// nothing is ever im2col'ed or padded in HBM.  B tiles are 3-D TMA boxes (k, cout, batch) of the packed
// K-major weight matrix (batch = 0 for convolutions, = image for the attention batched GEMMs).
//
// Warp roles (192 threads, 1 CTA / SM, persistent over tiles):
//   warp 0      : TMA producer (one elected lane), STAGES-deep smem ring, mbarrier full/empty
//   warp 1      : TMEM allocator + tcgen05.mma issuer (one lane); accumulators double-buffered in TMEM
//   warps 2..5  : epilogue - tcgen05.ld (32 lanes x 32 bit), bias / residual / scale, 16-byte stores
#pragma once
#include "common.cuh"

namespace igemm {

constexpr int BLOCK_M = 128;
constexpr int NUM_THREADS = 192;
constexpr int MAX_TAPS = 9;

struct Params {
  CUtensorMap a_map[4];
  CUtensorMap b_map;
  int box_w, box_h, box_n;        // pixels per A box (box_w*box_h*box_n <= 128)
  int tiles_w, tiles_h, tiles_n;  // m-tile grid
  int n_tiles;                    // tiles along Cout
  int num_taps, chunks_per_tap;   // K loop length = num_taps * chunks_per_tap
  int k_per_tap;                  // padded channels per tap in the packed weight matrix
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
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (launch error) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) __trap();
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
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
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
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  constexpr uint64_t layout = CHUNK_BYTES == 128 ? 2 : (CHUNK_BYTES == 64 ? 4 : 6);
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);            // start address      bits [0,14)
  d |= static_cast<uint64_t>(1) << 16;                               // leading byte off.  bits [16,30) (unused for swizzled K-major)
  d |= static_cast<uint64_t>((CHUNK_BYTES * 8) >> 4) << 32;          // stride byte offset bits [32,46)
  d |= static_cast<uint64_t>(1) << 46;                               // descriptor version bits [46,48)
  d |= layout << 61;                                                 // swizzle mode       bits [61,64)
  return d;
}

template <int BLOCK_N, int CHUNK_BYTES>
struct Config {
  static constexpr int A_BYTES = BLOCK_M * CHUNK_BYTES;
  static constexpr int B_BYTES_RAW = BLOCK_N * CHUNK_BYTES;
  static constexpr int B_BYTES = (B_BYTES_RAW + 1023) / 1024 * 1024;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BUDGET = 200 * 1024;
  static constexpr int STAGES_RAW = BUDGET / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64 ? 64 : (2 * BLOCK_N <= 128 ? 128 : (2 * BLOCK_N <= 256 ? 256 : 512)));
};

// Epilogue for 16 accumulator columns of one output pixel.
__device__ __forceinline__ void epilogue16(const Params& p, const uint32_t (&acc)[16], long long pix, int n0,
                                           bool valid_pix) {
  if (!valid_pix || n0 >= p.Cout) return;
  float v[16];
  const bool full = (n0 + 16 <= p.Cout);
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(acc[j]) * p.out_scale;
  if (p.bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (full || n0 + j < p.Cout) v[j] += __ldg(p.bias + n0 + j);
  }
  if (p.res != nullptr) {
    if (p.res_dtype == EOVAE_F32) {
      const float* r = reinterpret_cast<const float*>(p.res) + pix * p.res_pix_stride + n0;
      if (full) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          float4 t = *reinterpret_cast<const float4*>(r + j);
          v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
        }
      } else {
        for (int j = 0; j < 16 && n0 + j < p.Cout; ++j) v[j] += r[j];
      }
    } else {
      const uint16_t* r = reinterpret_cast<const uint16_t*>(p.res) + pix * p.res_pix_stride + n0;
      if (full) {
#pragma unroll
        for (int j = 0; j < 16; j += 8) {
          uint4 t = *reinterpret_cast<const uint4*>(r + j);
          float2 a = unpack16(t.x, p.res_dtype), b = unpack16(t.y, p.res_dtype);
          float2 c = unpack16(t.z, p.res_dtype), d = unpack16(t.w, p.res_dtype);
          v[j] += a.x; v[j + 1] += a.y; v[j + 2] += b.x; v[j + 3] += b.y;
          v[j + 4] += c.x; v[j + 5] += c.y; v[j + 6] += d.x; v[j + 7] += d.y;
        }
      } else {
        for (int j = 0; j < 16 && n0 + j < p.Cout; ++j) {
          uint32_t u = r[j];
          v[j] += unpack16(u, p.res_dtype).x;
        }
      }
    }
  }
  if (p.out_dtype == EOVAE_F32) {
    float* o = reinterpret_cast<float*>(p.out) + pix * p.out_pix_stride + n0;
    if (full) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
      for (int j = 0; j < 16 && n0 + j < p.Cout; ++j) o[j] = v[j];
    }
  } else {
    uint16_t* o = reinterpret_cast<uint16_t*>(p.out) + pix * p.out_pix_stride + n0;
    if (full) {
#pragma unroll
      for (int j = 0; j < 16; j += 8) {
        uint4 t;
        t.x = pack16(v[j], v[j + 1], p.out_dtype);
        t.y = pack16(v[j + 2], v[j + 3], p.out_dtype);
        t.z = pack16(v[j + 4], v[j + 5], p.out_dtype);
        t.w = pack16(v[j + 6], v[j + 7], p.out_dtype);
        *reinterpret_cast<uint4*>(o + j) = t;
      }
    } else {
      for (int j = 0; j < 16 && n0 + j < p.Cout; ++j) o[j] = static_cast<uint16_t>(pack16(v[j], 0.f, p.out_dtype) & 0xFFFF);
    }
  }
}

template <int BLOCK_N, int CHUNK_BYTES>
__global__ void __launch_bounds__(NUM_THREADS, 1) igemm_kernel(const __grid_constant__ Params p) {
  using Cfg = Config<BLOCK_N, CHUNK_BYTES>;
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

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int total_tiles = m_tiles * p.n_tiles;
  const int num_kb = p.num_taps * p.chunks_per_tap;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) prefetch_tmap(&p.a_map[i]);
    prefetch_tmap(&p.b_map);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);  // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(Cfg::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const uint32_t a_box_bytes = static_cast<uint32_t>(p.box_w * p.box_h * p.box_n) * CHUNK_BYTES;
      const uint32_t tx_bytes = a_box_bytes + Cfg::B_BYTES_RAW;
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int nt = tile % p.n_tiles;
        const int mt = tile / p.n_tiles;
        const int tw = mt % p.tiles_w;
        const int th = (mt / p.tiles_w) % p.tiles_h;
        const int tn = mt / (p.tiles_w * p.tiles_h);
        const int x0 = tw * p.box_w, y0 = th * p.box_h, img0 = tn * p.box_n;
        const int bb = p.b_batched ? img0 : 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          const int tap = kb / p.chunks_per_tap;
          const int cc = kb - tap * p.chunks_per_tap;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], tx_bytes);
          tma_load_4d(&p.a_map[p.tap_map[tap]], &full_bar[stage], smem_a + stage * Cfg::A_BYTES, cc * CH_ELEMS,
                      x0 + p.tap_dx[tap], y0 + p.tap_dy[tap], img0);
          tma_load_3d(&p.b_map, &full_bar[stage], smem_b + stage * Cfg::B_BYTES, tap * p.k_per_tap + cc * CH_ELEMS,
                      nt * BLOCK_N, bb);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = make_smem_desc<CHUNK_BYTES>(smem_u32(smem_a + stage * Cfg::A_BYTES));
          const uint64_t db = make_smem_desc<CHUNK_BYTES>(smem_u32(smem_b + stage * Cfg::B_BYTES));
#pragma unroll
          for (int k = 0; k < K_STEPS; ++k) {
            // advance the 14-bit start-address field by k*32 bytes (>>4)
            tc_mma_f16(d_tmem, da + static_cast<uint64_t>(k * 2), db + static_cast<uint64_t>(k * 2), p.idesc,
                       (kb | k) != 0 ? 1u : 0u);
          }
          tc_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tmem_full[acc]);  // accumulator complete
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps 2..5
    const int sub = warp & 3;  // TMEM sub-partition this warp may access: lanes [32*sub, 32*sub+32)
    const int row = sub * 32 + lane;
    const int box_pix = p.box_w * p.box_h * p.box_n;
    const int wi = row % p.box_w;
    const int hi = (row / p.box_w) % p.box_h;
    const int ni = row / (p.box_w * p.box_h);
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int nt = tile % p.n_tiles;
      const int mt = tile / p.n_tiles;
      const int tw = mt % p.tiles_w;
      const int th = (mt / p.tiles_w) % p.tiles_h;
      const int tn = mt / (p.tiles_w * p.tiles_h);
      const int ox = tw * p.box_w + wi, oy = th * p.box_h + hi, on = tn * p.box_n + ni;
      const bool valid = row < box_pix && ox < p.Wo && oy < p.Ho && on < p.Nimg;
      const long long pix = (static_cast<long long>(on) * p.Ho + oy) * p.Wo + ox;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(sub * 32) << 16) + acc * BLOCK_N;
      if constexpr (BLOCK_N >= 32) {
#pragma unroll 1
        for (int c = 0; c < BLOCK_N; c += 32) {
          uint32_t v0[16], v1[16];
          tc_ld16(taddr + c, v0);
          tc_ld16(taddr + c + 16, v1);
          tc_wait_ld();
          epilogue16(p, v0, pix, nt * BLOCK_N + c, valid);
          epilogue16(p, v1, pix, nt * BLOCK_N + c + 16, valid);
        }
      } else {
        uint32_t v0[16];
        tc_ld16(taddr, v0);
        tc_wait_ld();
        epilogue16(p, v0, pix, nt * BLOCK_N, valid);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::TMEM_COLS) : "memory");
  }
}

}  // namespace igemm
