"""GroupNorm-apply launch shapes on the encoder's tensors (batch 64, fp16): mode 0 = cp.async.bulk ring (default), 4 = the register-load kernels
(128 threads x 8 loads from 32 Mi elements up, else 256 x 4), 2 = 128 x 8 always."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "eo-vae_b200"))
import torch  # noqa: E402

from eo_vae import ops  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, n=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for name, n, c, h, w in (("L0 128ch@256", 64, 128, 256, 256), ("L1 256ch@128", 64, 256, 128, 128), ("L1 128ch@128", 64, 128, 128, 128),
                         ("L2 512ch@64", 64, 512, 64, 64), ("L2 256ch@64", 64, 256, 64, 64), ("L3 512ch@32", 64, 512, 32, 32),
                         ("train L0 128ch@256 b16", 16, 128, 256, 256)):
    x = torch.randn((n, h, w, c), device=dev).half().permute(0, 3, 1, 2)
    gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    stats = ops.gn_stats(x)
    out = []
    for mode in (0, 4, 2, 0, 4):
        ops.set_tuning(ops.TUNE_GN_APPLY_CORESIDENT, mode)
        ms = timeit(lambda: ops.gn_apply(x, stats, gamma, beta, True))
        out.append(f"mode {mode}: {ms:.3f} ms {2 * x.numel() * 2 / ms / 1e6:6.0f} GB/s")
    print(f"{name:26s} " + " | ".join(out), flush=True)

# block size of the ring kernel
x = torch.randn((64, 256, 256, 128), device=dev).half().permute(0, 3, 1, 2)
x2 = torch.randn((64, 32, 32, 512), device=dev).half().permute(0, 3, 1, 2)
ops.set_tuning(ops.TUNE_GN_APPLY_CORESIDENT, 0)
for xx, name in ((x, "L0 128ch@256"), (x2, "L3 512ch@32")):
    c = xx.shape[1]
    gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    stats = ops.gn_stats(xx)
    out = []
    for elems in (0, 8192, 16384, 32768, 65536, 131072, 262144):
        ops.set_tuning(ops.TUNE_GN_APPLY_BLOCK_ELEMS, elems)
        ms = timeit(lambda: ops.gn_apply(xx, stats, gamma, beta, True))
        out.append(f"{elems}: {ms:.3f} ms")
    print(f"{name:26s} block elems " + " | ".join(out), flush=True)
ops.set_tuning(ops.TUNE_GN_APPLY_BLOCK_ELEMS, 0)
