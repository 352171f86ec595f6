// cp.async.bulk (TMA, 1-D) shared-memory ring for streaming elementwise kernels: the bytes in flight per SM are set by the
// ring, not by how many loads the register allocator lets a thread keep outstanding.  A block owns one contiguous byte
// range of each of NT tensors and walks it in slots of kRingThreads * VEC 16-byte vectors; thread t owns vectors
// t, t + 256, ... of a slot.  One elected thread issues the copies, every thread waits on the slot's mbarrier.
#pragma once
#include <cstdint>

namespace eovae {

constexpr int kRingThreads = 256;

__device__ __forceinline__ uint32_t s_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void bar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count));
}
__device__ __forceinline__ void bar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(s_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
    if (++spins > (1u << 21)) __trap();  // a protocol bug traps instead of hanging the device
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s_u32(dst)),
               "l"(src), "r"(bytes), "r"(s_u32(bar))
               : "memory");
}

// ring of STAGES slots, each NT tensors x (kRingThreads * VEC) uint4
template <int NT, int VEC, int STAGES>
struct BulkRing {
  static constexpr int kSlotVecs = kRingThreads * VEC;
  static constexpr uint32_t kSlotBytes = kSlotVecs * 16;
  static constexpr size_t kSmemBytes = static_cast<size_t>(STAGES) * NT * kSlotBytes + 128;
  uint4* slots;
  uint64_t* full;
  const char* src[NT];
  long long total_bytes;  // of this block's range, per tensor
  int nchunks;
  __device__ __forceinline__ void init(unsigned char* smem, long long total) {
    slots = reinterpret_cast<uint4*>(smem);
    full = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(STAGES) * NT * kSlotBytes);
    total_bytes = total;
    nchunks = static_cast<int>((total + kSlotBytes - 1) / kSlotBytes);
    if (threadIdx.x == 0) {
      for (int i = 0; i < STAGES; ++i) bar_init(&full[i], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0)
      for (int i = 0; i < STAGES && i < nchunks; ++i) issue(i);
  }
  __device__ __forceinline__ void issue(int chunk) {  // one thread
    const int s = chunk % STAGES;
    const long long off = static_cast<long long>(chunk) * kSlotBytes;
    const long long left = total_bytes - off;
    const uint32_t bytes = left < static_cast<long long>(kSlotBytes) ? static_cast<uint32_t>(left) : kSlotBytes;
    bar_expect_tx(&full[s], bytes * NT);
#pragma unroll
    for (int t = 0; t < NT; ++t) bulk_g2s(slot(s, t), src[t] + off, bytes, &full[s]);
  }
  __device__ __forceinline__ uint4* slot(int s, int t) { return slots + (static_cast<size_t>(s) * NT + t) * kSlotVecs; }
  __device__ __forceinline__ void wait(int chunk) { bar_wait(&full[chunk % STAGES], (chunk / STAGES) & 1); }
  // every thread has copied its vectors of `chunk` to registers: hand the slot back to the copy engine
  __device__ __forceinline__ void release(int chunk) {
    __syncthreads();
    if (threadIdx.x == 0 && chunk + STAGES < nchunks) issue(chunk + STAGES);
  }
  // vectors of the chunk that hold data (the last chunk of a block may be short)
  __device__ __forceinline__ int valid_vecs(int chunk) const {
    const long long left = total_bytes - static_cast<long long>(chunk) * kSlotBytes;
    return left >= static_cast<long long>(kSlotBytes) ? kSlotVecs : static_cast<int>(left >> 4);
  }
};


}  // namespace eovae
