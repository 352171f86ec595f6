"""Host-side (Python / ctypes issue) cost of one eager training step: wall time of issuing a step without waiting for the GPU,
and the cProfile top list."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "eo-vae_b200"))
import torch  # noqa: E402

import __graft_entry__ as g  # noqa: E402
from eo_vae.models.modules.consistency_loss import EOConsistencyLoss  # noqa: E402
from oracle.weights import FULL_CONFIG, WAVELENGTHS, make_state_dict  # noqa: E402

dev = torch.device("cuda:0")
model = g._model(FULL_CONFIG, make_state_dict(FULL_CONFIG, 0), dev)
model.train()
model.loss_fn = EOConsistencyLoss(pixel_weight=1.0, rec_loss_type="char", msssim_weight=1.0, msssim_start_step=0).to(dev)
model.clip_grad = 1.0
wvs = torch.tensor(WAVELENGTHS["S2L2A"], device=dev)
x = torch.randn((16, 12, 256, 256), device=dev).clamp_(-2, 6)
batch = {model.image_key: x, "wvs": wvs}
for i in range(3):
    model.training_step(batch, i)
torch.cuda.synchronize()
# tiny batch: the GPU is never the bottleneck, so the wall time per step IS the host issue time
xs = torch.randn((1, 12, 192, 192), device=dev)
small = {model.image_key: xs, "wvs": wvs}
for i in range(3):
    model.training_step(small, i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(10):
    model.training_step(small, i)
torch.cuda.synchronize()
print(f"host issue time per eager step (1 x 12 x 192 x 192, host-bound): {(time.perf_counter() - t0) / 10 * 1e3:.1f} ms")
pr = cProfile.Profile()
pr.enable()
for i in range(5):
    model.training_step(small, i)
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(22)
