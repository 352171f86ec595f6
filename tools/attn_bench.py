"""Mid-block attention: fused flash-style kernel vs the GEMM -> softmax -> GEMM path (batch 64, L 1024, C 512)."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "eo-vae_b200"))
import torch
from eo_vae import ops
dev = torch.device("cuda:0")
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for n, l, c in ((64, 1024, 512), (32, 4096, 512), (16, 1024, 512)):
    qkv = torch.randn((n, l, 3 * c), device=dev).bfloat16()
    def unfused():
        q, k, v = qkv[:, :, :c], qkv[:, :, c:2 * c], qkv[:, :, 2 * c:]
        s = ops.gemm_tn_batched(q, k, torch.float32, scale=1.0 / math.sqrt(c))
        p = ops.softmax_rows(s, torch.bfloat16)
        return ops.gemm_tn_batched(p, ops.transpose16(v), torch.bfloat16)
    t_f = timeit(lambda: ops.attention_fused(qkv, c))
    t_u = timeit(unfused)
    fl = 4.0 * n * l * l * c
    print(f"n {n} L {l} C {c}: fused {t_f:.3f} ms ({fl / t_f / 1e9:.0f} TFLOP/s algorithmic), unfused {t_u:.3f} ms ({fl / t_u / 1e9:.0f} TFLOP/s); "
          f"scores+probs not written: {n * l * l * 6 / 2**20:.0f} MiB")
