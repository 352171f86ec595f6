"""Weight re-packing kernels (forward operand, data-gradient operand) on the model's weight shapes."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "eo-vae_b200"))
import torch  # noqa: E402

from eo_vae import ops  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for cout, cin, k in ((512, 512, 3), (256, 256, 3), (128, 128, 3), (512, 256, 1), (1536, 512, 1)):
    w = torch.randn((cout, cin, k, k), device=dev)
    a = timeit(lambda: ops.pack_conv_weight(w, torch.bfloat16))
    b = timeit(lambda: ops.pack_conv_weight_dgrad(w, torch.bfloat16))
    mb = w.numel() * 6 / 1e6
    print(f"{cout}x{cin}x{k}x{k}: forward pack {a:.1f} us, dgrad pack {b:.1f} us ({mb:.1f} MB moved each)", flush=True)
