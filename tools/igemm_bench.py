"""Micro-benchmark of the implicit-GEMM kernel on the encoder's dominant conv shapes (batch 64), with the kernel's
debug modes that switch off one pipeline actor at a time (epilogue / MMA / TMA) to find the binding one."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "eo-vae_b200"))
import torch  # noqa: E402

from eo_vae import _C, ops  # noqa: E402

dev = torch.device("cuda:0")
SHAPES = [  # name, n, h, w, cin, cout, residual, gn
    ("L0 128->128 @256 conv1", 64, 256, 256, 128, 128, False, True),
    ("L0 128->128 @256 conv2+res", 64, 256, 256, 128, 128, True, True),
    ("L1 256->256 @128 conv1", 64, 128, 128, 256, 256, False, True),
    ("L1 256->256 @128 conv2+res", 64, 128, 128, 256, 256, True, True),
    ("L2 512->512 @64 conv1", 64, 64, 64, 512, 512, False, True),
    ("L3 512->512 @32 conv1", 64, 32, 32, 512, 512, False, True),
    ("dyn 16->128 @256", 64, 256, 256, 16, 128, False, True),
]
modes = [int(m) for m in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["0", "1", "2", "3"])]
ctas = int(sys.argv[2]) if len(sys.argv) > 2 else 0  # bits: 1/2 = forced CTA group size, +4 = 64-channel stages, +8 = no TMA store, +16 = no halo
only = sys.argv[3].split(",") if len(sys.argv) > 3 else None
for name, n, h, w, cin, cout, res, gn in SHAPES:
    if only and not any(name.startswith(o) for o in only):
        continue
    x = torch.randn((n, h, w, cin), device=dev).bfloat16().permute(0, 3, 1, 2)
    wgt = (torch.randn(cout, cin, 3, 3, device=dev) / math.sqrt(9 * cin))
    wp = ops.pack_conv_weight(wgt, torch.bfloat16)
    bias = torch.randn(cout, device=dev)
    r = torch.randn((n, h, w, cout), device=dev).bfloat16().permute(0, 3, 1, 2) if res else None
    flops = 2.0 * n * h * w * cout * cin * 9
    line = f"{name:30s}"
    if os.environ.get("GNP") and ops.gn_prologue_ok(x, cout, ops.CONV_3X3):
        st = ops.gn_stats(x)
        gam, bet = torch.ones(cin, device=dev), torch.zeros(cin, device=dev)
        dm = int(os.environ.get("GNP_DEBUG", "0"))
        _C.lib().eovae_set_debug_mode(dm)
        for label, fn in ((f"fused GN prologue dbg{dm}", lambda: ops.conv2d(x, wp, bias, cout, ops.CONV_3X3, residual=r, gn_groups=32, in_gn=(st, gam, bet, 32))),
                          ("gn_apply + conv  ", lambda: ops.conv2d(ops.gn_apply(x, st, gam, bet, True), wp, bias, cout, ops.CONV_3X3, residual=r, gn_groups=32))):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            line += f" | {label} {ms:7.3f} ms {flops / ms / 1e9:7.1f} TF"
        _C.lib().eovae_set_debug_mode(0)
        print(line, flush=True)
        continue
    for mode in modes:
        _C.lib().eovae_set_debug_mode(mode | (ctas << 8))
        for gflag in ((True, False) if (mode == 0 and gn) else (False,)):
            f = lambda: ops.conv2d(x, wp, bias, cout, ops.CONV_3X3, residual=r, gn_groups=32 if gflag else 0)
            for _ in range(2):
                f()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                f()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            line += f" | m{mode}{'g' if gflag else ' '} {ms:7.3f} ms {flops / ms / 1e9:7.1f} TF"
    _C.lib().eovae_set_debug_mode(0)
    print(line, flush=True)
