"""Experiment: does running the level-0 ResnetBlocks in L2-sized sub-batches (conv output -> GroupNorm-apply -> conv with the
intermediate resident in the 126 MB L2) beat the whole-batch launch order?  Per-sample time of (a) gn_apply alone and
(b) one level-0 ResnetBlock, at batch 64 in one go vs chunks of 1 / 2 / 4 samples."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "eo-vae_b200"))
import torch  # noqa: E402

import __graft_entry__ as g  # noqa: E402
from eo_vae import ops  # noqa: E402
from oracle.weights import FULL_CONFIG, make_state_dict  # noqa: E402

dev = torch.device("cuda:0")
model = g._model(FULL_CONFIG, make_state_dict(FULL_CONFIG, 3), dev)


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


B = 64
with torch.no_grad():
    for name, blk, c, hw in (("L0 block (128ch @256)", model.encoder.down[0].block[0], 128, 256),
                             ("L1 block (256ch @128)", model.encoder.down[1].block[1], 256, 128)):
        x = torch.randn((B, hw, hw, c), device=dev).bfloat16().permute(0, 3, 1, 2)
        gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
        stats = ops.gn_stats(x)
        full = timeit(lambda: ops.gn_apply(x, stats, gamma, beta, True))
        print(f"{name}: gn_apply batch {B}: {full / B * 1e3:.2f} us/sample")
        for chunk in (1, 2, 4, 8):
            xs, ss = x[:chunk], stats[:chunk]
            t = timeit(lambda: ops.gn_apply(xs, ss, gamma, beta, True), 50)
            print(f"    gn_apply chunk {chunk} (L2 resident): {t / chunk * 1e3:.2f} us/sample")
        full = timeit(lambda: blk(x))
        print(f"{name}: ResnetBlock batch {B}: {full / B * 1e3:.2f} us/sample ({full:.3f} ms)")
        for chunk in (2, 4, 8, 16):
            def run():
                for i in range(0, B, chunk):
                    blk(x[i:i + chunk])
            t = timeit(run, 5)
            print(f"    ResnetBlock in chunks of {chunk}: {t / B * 1e3:.2f} us/sample ({t:.3f} ms)")
