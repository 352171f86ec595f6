"""Where the non-GEMM time of the training step goes: optimiser (clip + Adam) and the two hypernetworks, timed alone."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "eo-vae_b200"))
import torch  # noqa: E402

import __graft_entry__ as g  # noqa: E402
from oracle.weights import FULL_CONFIG, WAVELENGTHS, make_state_dict  # noqa: E402

dev = torch.device("cuda:0")
model = g._model(FULL_CONFIG, make_state_dict(FULL_CONFIG, 0), dev)
model.train()
wvs = torch.tensor(WAVELENGTHS["S2L2A"], device=dev)


def timeit(fn, n=20):
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


opt = model.configure_optimizers()
opt = opt[0] if isinstance(opt, (list, tuple)) else opt
params = opt.param_groups[0]["params"]
for p in params:
    p.grad = torch.randn_like(p) * 1e-3
print(f"parameters: {sum(p.numel() for p in params) / 1e6:.1f} M in {len(params)} tensors")
print(f"torch clip_grad_norm_: {timeit(lambda: torch.nn.utils.clip_grad_norm_(params, 1.0)):.3f} ms")
ref = torch.optim.Adam([p.detach().clone().requires_grad_(True) for p in params], lr=1e-4, fused=True)
for q, p in zip(ref.param_groups[0]["params"], params):
    q.grad = p.grad.clone()
print(f"torch fused Adam step: {timeit(lambda: ref.step()):.3f} ms")
del ref
print(f"{type(opt).__name__}.step(): {timeit(lambda: opt.step()):.3f} ms")
print(f"{type(opt).__name__}.step(clip_norm=1.0) (norm + clipped update): {timeit(lambda: opt.step(clip_norm=1.0)):.3f} ms")
for name, mod in (("encoder.conv_in", model.encoder.conv_in), ("decoder.conv_out", model.decoder.conv_out)):
    t = timeit(lambda: mod._generate_taped(wvs))
    wk, b_raw, tp = mod._generate_taped(wvs)
    c = wvs.numel()
    dw = torch.randn((c, 128, 3, 3) if mod._decoder else (128, 16, 3, 3), device=dev)
    db = torch.randn((c if mod._decoder else 128,), device=dev)
    tb = timeit(lambda: mod._hyper_backward(wvs, dw, db, 0.1, tp))
    print(f"hypernet {name}: taped forward {t:.3f} ms, backward {tb:.3f} ms")
