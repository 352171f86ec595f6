"""Does the HBM-bound GroupNorm-apply pass run UNDER a resident implicit-GEMM convolution launched on another stream?
Times conv alone, gn_apply alone and both (two streams) for the level-0 shape; prints the overlap achieved."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "eo-vae_b200"))
import torch  # noqa: E402

from eo_vae import ops  # noqa: E402

dev = torch.device("cuda:0")
n, c, hw = int(sys.argv[1]) if len(sys.argv) > 1 else 32, 128, 256
dt = torch.float16
x = torch.randn((n, hw, hw, c), device=dev, dtype=dt).permute(0, 3, 1, 2)
y = torch.randn((n, hw, hw, c), device=dev, dtype=dt).permute(0, 3, 1, 2)
w = torch.randn((c, c, 3, 3), device=dev) * 0.05
wp = ops.pack_conv_weight(w, dt)
bias = torch.zeros((c,), device=dev)
gamma, beta = torch.ones((c,), device=dev), torch.zeros((c,), device=dev)
stats = ops.gn_stats(y)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def conv():
    return ops.conv2d(x, wp, bias, c, ops.CONV_3X3, gn_groups=32)


def apply():
    return ops.gn_apply(y, stats, gamma, beta, True)


def timed(fa, fb, reps=10):
    for _ in range(3):
        if fa: fa()
        if fb: fb()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_stream(torch.cuda.current_stream())
    s2.wait_stream(torch.cuda.current_stream())
    for _ in range(reps):
        if fa:
            with torch.cuda.stream(s1):
                fa()
        if fb:
            with torch.cuda.stream(s2):
                fb()
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


if len(sys.argv) > 2:
    ops._C.lib().eovae_set_debug_mode(int(sys.argv[2]))   # e.g. 256 = force single-CTA (no cluster pairs), 4096 = no halo
for co in (0, 1):
    ops.set_tuning(ops.TUNE_GN_APPLY_CORESIDENT, co)
    tc, ta, tb = timed(conv, None), timed(None, apply), timed(conv, apply)
    print(f"batch {n} 128ch@256^2, gn_apply shape {'co-resident (128 thr x 8 loads, 80 regs)' if co else 'default (256 thr x 4 loads)'}: "
          f"conv {tc:.3f} ms, gn_apply {ta:.3f} ms ({2 * x.numel() * 2 / ta / 1e6:.0f} GB/s), both {tb:.3f} ms -> "
          f"hidden {tc + ta - tb:.3f} ms of {ta:.3f}", flush=True)

# timeline of one pair: event timestamps relative to a common origin
ops.set_tuning(ops.TUNE_GN_APPLY_CORESIDENT, 1)
for trial in range(3):
    torch.cuda.synchronize()
    base = torch.cuda.Event(enable_timing=True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    base.record()
    s1.wait_stream(torch.cuda.current_stream())
    s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s1):
        ev[0].record()
        conv()
        ev[1].record()
    with torch.cuda.stream(s2):
        ev[2].record()
        apply()
        ev[3].record()
    torch.cuda.synchronize()
    t = [base.elapsed_time(e) for e in ev]
    print(f"timeline: conv [{t[0]:.3f}, {t[1]:.3f}] ms, gn_apply [{t[2]:.3f}, {t[3]:.3f}] ms", flush=True)

# how does the conv's duration respond to the SIZE of the co-resident pass?
for frac in (1, 2, 4, 8):
    ys = y[: max(1, n // frac)]
    st = ops.gn_stats(ys)
    torch.cuda.synchronize()
    base = torch.cuda.Event(enable_timing=True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    base.record()
    s1.wait_stream(torch.cuda.current_stream())
    s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s1):
        ev[0].record()
        conv()
        ev[1].record()
    with torch.cuda.stream(s2):
        ev[2].record()
        ops.gn_apply(ys, st, gamma, beta, True)
        ev[3].record()
    torch.cuda.synchronize()
    t = [base.elapsed_time(e) for e in ev]
    print(f"gn_apply on {ys.shape[0]} images: conv [{t[0]:.3f}, {t[1]:.3f}] ms, gn_apply [{t[2]:.3f}, {t[3]:.3f}] ms", flush=True)
