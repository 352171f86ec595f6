"""Where does the epilogue-bound dynamic input conv (16 -> 128 channels @256^2, batch 64: 1 GB written, ~0.1 TFLOP) spend its
time?  Variants: GroupNorm statistics on / off, TMA-store epilogue vs per-thread stores, bias on / off, fp32 output."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "eo-vae_b200"))
import torch  # noqa: E402

from eo_vae import ops  # noqa: E402

dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dt = torch.float16


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for cin, cout, label, hw in ((16, 128, "dyn conv-in 16->128", 256), (128, 128, "L0 128->128", 256), (256, 256, "L1 256->256", 128),
                             (512, 512, "L2 512->512", 64)):
    x = torch.randn((n, hw, hw, cin), device=dev, dtype=dt).permute(0, 3, 1, 2)
    w = torch.randn((cout, cin, 3, 3), device=dev) * 0.05
    wp = ops.pack_conv_weight(w, dt)
    bias = torch.zeros((cout,), device=dev)
    res = torch.randn((n, hw, hw, cout), device=dev, dtype=dt).permute(0, 3, 1, 2)
    for name, mode, kw in (("stats + bias", 0, dict(gn_groups=32)), ("no stats", 0, dict()),
                           ("loads+MMA off, no stats", 6, dict()), ("loads+MMA off, no stats, no store issue", 6 | 32, dict()),
                           ("loads+MMA off, no stats, no tcgen05.ld", 6 | 64, dict()),
                           ("loads+MMA off, no stats, no store, no tcgen05.ld", 6 | 32 | 64, dict()),
                           ("loads+MMA off, stats", 6, dict(gn_groups=32)), ("loads+MMA off, stats, no store issue", 6 | 32, dict(gn_groups=32)),
                           ("stats + residual (TMA)", 0, dict(gn_groups=32, residual=res)),
                           ("stats + residual (registers)", 1 << 15, dict(gn_groups=32, residual=res)),
                           ("all actors off", 7, dict())):
        ops._C.lib().eovae_set_debug_mode(mode)
        nobias = kw.pop("nobias", False)
        ms = timeit(lambda: ops.conv2d(x, wp, None if nobias else bias, cout, ops.CONV_3X3, **kw))
        print(f"{label:22s} {name:32s} {ms:.3f} ms  ({n * hw * hw * cout * 2 / ms / 1e6:.0f} GB/s written)", flush=True)
    ops._C.lib().eovae_set_debug_mode(0)
