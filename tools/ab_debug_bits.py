"""A/B of library debug bits (eovae_set_debug_mode) on the training step and the encode step, alternating in one process.
usage: python tools/ab_debug_bits.py <bits> [<bits> ...]   (0 is always the baseline)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "eo-vae_b200"))
import torch  # noqa: E402

import __graft_entry__ as g  # noqa: E402
from eo_vae import ops  # noqa: E402
from eo_vae.models.modules.consistency_loss import EOConsistencyLoss  # noqa: E402
from oracle.weights import FULL_CONFIG, WAVELENGTHS, make_state_dict  # noqa: E402

# arguments: debug-bit masks (0x4000 ...) or tuning settings "t<key>=<value>" (eovae_set_tuning), e.g. t1=4
modes = [0] + [a if a.startswith("t") else int(a, 0) for a in sys.argv[1:]]
TUNE_DEFAULTS = {}


def apply_mode(m):
    from eo_vae import ops as _ops
    for k, v in TUNE_DEFAULTS.items():
        _ops.set_tuning(k, v)
    if isinstance(m, str):
        k, v = m[1:].split("=")
        TUNE_DEFAULTS.setdefault(int(k), 0)
        _ops.set_tuning(int(k), int(v))
        _ops._C.lib().eovae_set_debug_mode(0)
    else:
        _ops._C.lib().eovae_set_debug_mode(m)


def label(m):
    return m if isinstance(m, str) else f"{m:#x}"


dev = torch.device("cuda:0")
model = g._model(FULL_CONFIG, make_state_dict(FULL_CONFIG, 0), dev)
wvs = torch.tensor(WAVELENGTHS["S2L2A"], device=dev)


def timeit(fn, n):
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


x64 = torch.randn((64, 12, 256, 256), device=dev).clamp_(-2, 6)
model.eval()
for rep in range(2):
    for m in modes:
        apply_mode(m)
        with torch.no_grad():
            t = timeit(lambda: model.encode_spatial_normalized(x64, wvs), 20)
        print(f"encode batch 64, setting {label(m)}: {t:.3f} ms = {64 / t * 1e3:.1f} patches/s", flush=True)
del x64
model.train()
model.loss_fn = EOConsistencyLoss(pixel_weight=1.0, rec_loss_type="char", msssim_weight=1.0, msssim_start_step=0).to(dev)
model.clip_grad = 1.0
x = torch.randn((16, 12, 256, 256), device=dev).clamp_(-2, 6)
batch = {model.image_key: x, "wvs": wvs}
for rep in range(2):
    for m in modes:
        apply_mode(m)
        t = timeit(lambda: model.training_step(batch, 0), 8)
        print(f"training step, setting {label(m)}: {t:.2f} ms = {16 / t * 1e3:.1f} patches/s", flush=True)
apply_mode(0)
