"""GB/s of the HBM-bound kernels on the encoder's largest tensors (batch 64), against MEASURED_PEAKS hbm_gbs."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "eo-vae_b200"))
import torch  # noqa: E402

from eo_vae import ops  # noqa: E402

dev = torch.device("cuda:0")
peak = 6549.8
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = json.load(open(pk))["hbm_gbs"]


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


rows = []
for name, n, c, h, w in (("L0 128ch@256", 64, 128, 256, 256), ("L1 256ch@128", 64, 256, 128, 128),
                         ("L2 512ch@64", 64, 512, 64, 64), ("L3 512ch@32", 64, 512, 32, 32)):
    x = torch.randn((n, h, w, c), device=dev).bfloat16().permute(0, 3, 1, 2)
    gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    nbytes = x.numel() * 2
    stats = ops.gn_stats(x)
    ms = timeit(lambda: ops.gn_stats(x))
    rows.append((f"gn_stats {name}", nbytes, ms))
    ms = timeit(lambda: ops.gn_apply(x, stats, gamma, beta, True))
    rows.append((f"gn_apply+silu {name}", 2 * nbytes, ms))
x = torch.randn((64, 12, 256, 256), device=dev)
ms = timeit(lambda: ops.nchw_to_act(x, 16, torch.bfloat16))
rows.append(("nchw_to_nhwc16 12->16ch@256", x.numel() * 4 + 64 * 65536 * 16 * 2, ms))
s = torch.randn((64, 1024, 1024), device=dev)
ms = timeit(lambda: ops.softmax_rows(s, torch.bfloat16))
rows.append(("softmax 64x1024x1024 f32->bf16", s.numel() * 6, ms))
a, b = torch.randn((64, 12, 256, 256), device=dev), torch.randn((64, 12, 256, 256), device=dev)
ms = timeit(lambda: ops.l1_charbonnier(a, b))
rows.append(("l1+charbonnier 64x12x256x256", a.numel() * 8, ms))
# ---- training path (batch 16): algorithmic bytes = the minimum traffic of the op (every operand read once, result written once)
for name, n, c, h, w in (("L0 128ch@256", 16, 128, 256, 256), ("L1 256ch@128", 16, 256, 128, 128), ("L2 512ch@64", 16, 512, 64, 64)):
    x = torch.randn((n, h, w, c), device=dev).bfloat16().permute(0, 3, 1, 2)
    g = torch.randn((n, h, w, c), device=dev).bfloat16().permute(0, 3, 1, 2)
    gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    nbytes = x.numel() * 2
    stats = ops.gn_stats(x)
    ms = timeit(lambda: ops.gn_backward(x, g, stats, gamma, beta, True))
    rows.append((f"gn_backward+silu (+bias sums) {name}", 3 * nbytes, ms))
    ms = timeit(lambda: ops.gn_backward(x, g, stats, gamma, beta, True, grad_add=g))
    rows.append((f"gn_backward+silu+grad_add {name}", 4 * nbytes, ms))
    g._colsum = None
    ms = timeit(lambda: ops.bias_grad(g))
    rows.append((f"bias_grad {name}", nbytes, ms))
x = torch.randn((16, 128, 128, 128), device=dev).bfloat16().permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
ms = timeit(lambda: ops.pool2x2_sum(x))
rows.append(("pool2x2_sum 128ch 128->64", x.numel() * 2 * 1.25, ms))
ms = timeit(lambda: ops.upsample2x(x))
rows.append(("upsample2x 128ch 128->256", x.numel() * 2 * 5, ms))
a, b = torch.randn((16, 12, 256, 256), device=dev), torch.randn((16, 12, 256, 256), device=dev)
one = torch.ones((), device=dev)
ms = timeit(lambda: ops.pixel_loss_backward(a, b, 1e-3, 1, one))
rows.append(("charbonnier backward 16x12x256x256", a.numel() * 12, ms))
ms = timeit(lambda: ops.msssim(a, b, 6.0))
rows.append(("ms-ssim forward 16x12x256x256 (10/3 N s)", a.numel() * 4 * 10 / 3, ms))
ms = timeit(lambda: ops.msssim_backward(a, b, 6.0, one))
rows.append(("ms-ssim backward (fwd + adjoint, 6 N s)", a.numel() * 4 * 6, ms))
# optional EOConsistencyLoss branches: forward = read both tensors, backward = read both + write the gradient
ms = timeit(lambda: ops.sam_loss(a, b))
rows.append(("sam forward 16x12x256x256", a.numel() * 8, ms))
ms = timeit(lambda: ops.sam_loss_backward(a, b, 1e-8, one))
rows.append(("sam backward", a.numel() * 12, ms))
ms = timeit(lambda: ops.grad_diff_loss(a, b))
rows.append(("gradient-difference forward", a.numel() * 8, ms))
ms = timeit(lambda: ops.grad_diff_loss_backward(a, b, one))
rows.append(("gradient-difference backward", a.numel() * 12, ms))
ms = timeit(lambda: ops.focal_freq_loss(a, b, 2, 1.0))
ffl_flop = 6 * 2 * 128 * a.numel()  # six real 128-wide DFT-matrix products over every element
print(f"focal-frequency forward 16x12x256x256 (pf 2): {ms:.3f} ms = {ffl_flop / ms / 1e9:.1f} TFLOP/s fp32 SIMT (compute bound, not in the GB/s table)")
_, ffl_ws = ops.focal_freq_loss(a, b, 2, 1.0, keep=True)
ms = timeit(lambda: ops.focal_freq_loss_backward(tuple(a.shape), 2, one, ffl_ws))
print(f"focal-frequency backward: {ms:.3f} ms = {ffl_flop / ms / 1e9:.1f} TFLOP/s fp32 SIMT")
p = torch.softmax(torch.randn((16, 1024, 1024), device=dev), -1).bfloat16()
dp = torch.randn((16, 1024, 1024), device=dev)
ms = timeit(lambda: ops.softmax_backward(p, dp, 1024, 0.044))
rows.append(("softmax backward 16x1024x1024", p.numel() * 8, ms))
for name, nbytes, ms in rows:
    gbs = nbytes / ms / 1e6
    print(f"{name:36s} {ms:8.3f} ms  {gbs:8.1f} GB/s  {gbs / peak:5.2f} of measured HBM peak ({peak:.0f})")
