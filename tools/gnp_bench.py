"""GroupNorm + SiLU in the conv prologue (transform warps inside the halo mainloop) vs gn_apply + conv, batch 64, fp16."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "eo-vae_b200"))
import torch  # noqa: E402

from eo_vae import ops  # noqa: E402

dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dt = torch.float16 if (len(sys.argv) <= 2 or sys.argv[2] == "fp16") else torch.bfloat16


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


for cin, cout, hw in ((128, 128, 256), (128, 256, 128), (256, 256, 128)):
    x = (torch.randn((n, hw, hw, cin), device=dev) * 1.5 + 0.4).to(dt).permute(0, 3, 1, 2)
    res = torch.randn((n, hw, hw, cout), device=dev).to(dt).permute(0, 3, 1, 2)
    w = torch.randn((cout, cin, 3, 3), device=dev) / (9 * cin) ** 0.5
    wp = ops.pack_conv_weight(w, dt)
    bias = torch.zeros((cout,), device=dev)
    gamma, beta = torch.ones((cin,), device=dev), torch.zeros((cin,), device=dev)
    stats = ops.gn_stats(x)
    assert ops.gn_prologue_ok(x, cout, ops.CONV_3X3)
    for with_res in (False, True):
        kw = dict(residual=res) if with_res else {}
        t_apply = timeit(lambda: ops.gn_apply(x, stats, gamma, beta, True))
        a = ops.gn_apply(x, stats, gamma, beta, True)
        t_conv = timeit(lambda: ops.conv2d(a, wp, bias, cout, ops.CONV_3X3, gn_groups=32, **kw))
        t_fused = timeit(lambda: ops.conv2d(x, wp, bias, cout, ops.CONV_3X3, gn_groups=32, in_gn=(stats, gamma, beta, 32), **kw))
        y0 = ops.conv2d(a, wp, bias, cout, ops.CONV_3X3, out_dtype=torch.float32, **kw)
        y1 = ops.conv2d(x, wp, bias, cout, ops.CONV_3X3, out_dtype=torch.float32, in_gn=(stats, gamma, beta, 32), **kw)
        err = float((y1 - y0).norm() / y0.norm())
        print(f"{cin:3d}->{cout:3d} @{hw}^2 batch {n} {'+res' if with_res else '    '}: gn_apply {t_apply:.3f} + conv {t_conv:.3f} = {t_apply + t_conv:.3f} ms | "
              f"fused prologue {t_fused:.3f} ms ({(t_apply + t_conv) / t_fused:.2f}x) | fused vs unfused rel {err:.2e}", flush=True)
