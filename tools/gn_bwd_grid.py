"""Sweep of the GroupNorm-backward block size (eovae_set_tuning(EOVAE_TUNE_GN_BWD_BLOCK_ELEMS)) on the training step's
shapes (batch 16) and on the whole training step."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "eo-vae_b200"))
import torch  # noqa: E402

import __graft_entry__ as g  # noqa: E402
from eo_vae import ops  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n


SHAPES = [(16, 256, 256, 128), (16, 128, 128, 256), (16, 64, 64, 512), (16, 32, 32, 512)]
for bulk, elems in ((0, 65536), (1, 32768), (1, 65536), (1, 131072), (1, 262144), (1, 524288)):
    ops.set_tuning(ops.TUNE_GN_BWD_BLOCK_ELEMS, elems)
    ops.set_tuning(ops.TUNE_GN_BWD_BULK, bulk)
    line = [f"bulk {bulk} block elems {elems:6d}:"]
    for (n, h, w, c) in SHAPES:
        x = torch.randn((n, h, w, c), device=dev).to(torch.bfloat16).permute(0, 3, 1, 2)
        gr = torch.randn((n, h, w, c), device=dev).to(torch.bfloat16).permute(0, 3, 1, 2)
        gamma = torch.rand(c, device=dev) + 0.5
        beta = torch.randn(c, device=dev) * 0.1
        stats = ops.gn_stats(x)
        ms = timeit(lambda: ops.gn_backward(x, gr, stats, gamma, beta, True))
        ms2 = timeit(lambda: ops.gn_backward(x, gr, stats, gamma, beta, True, grad_add=gr))
        gbs = 3 * x.numel() * 2 / ms / 1e6
        line.append(f"{h}x{w}x{c} {ms:.3f} ms ({gbs:.0f} GB/s) +add {ms2:.3f}")
        del x, gr
    print("  ".join(line), flush=True)

if os.environ.get("STEP", "1") == "1":
    from eo_vae.models.modules.consistency_loss import EOConsistencyLoss  # noqa: E402
    from oracle.weights import FULL_CONFIG, WAVELENGTHS, make_state_dict  # noqa: E402
    model = g._model(FULL_CONFIG, make_state_dict(FULL_CONFIG, 0), dev)
    wvs = torch.tensor(WAVELENGTHS["S2L2A"], device=dev)
    model.train()
    model.loss_fn = EOConsistencyLoss(pixel_weight=1.0, rec_loss_type="char", msssim_weight=1.0, msssim_start_step=0).to(dev)
    model.clip_grad = 1.0
    x = torch.randn((16, 12, 256, 256), device=dev).clamp_(-2, 6)
    batch = {model.image_key: x, "wvs": wvs}
    for bulk, elems in ((0, 65536), (1, 65536), (1, 131072), (1, 262144), (0, 65536), (1, 65536), (1, 131072), (1, 262144)):
        ops.set_tuning(ops.TUNE_GN_BWD_BLOCK_ELEMS, elems)
        ops.set_tuning(ops.TUNE_GN_BWD_BULK, bulk)
        for _ in range(3):
            model.training_step(batch, 0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8):
            model.training_step(batch, 0)
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 8
        print(f"training step, bulk {bulk} block elems {elems}: {t:.2f} ms = {16 / t * 1e3:.1f} patches/s", flush=True)
