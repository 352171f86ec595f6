"""BASELINE configs[4]: encode_spatial_normalized on S2L1C 13-band 512x512 patches, batch 32 (device-resident)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "eo-vae_b200"))
import torch  # noqa: E402

import __graft_entry__ as g  # noqa: E402
from oracle.weights import FULL_CONFIG, WAVELENGTHS, make_state_dict  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0")
model = g._model(FULL_CONFIG, make_state_dict(FULL_CONFIG, 0), dev)
wvs = torch.tensor(WAVELENGTHS["S2L1C"], device=dev)
x = torch.randn((batch, 13, 512, 512), device=dev).clamp_(-2, 6)
with torch.no_grad():
    for _ in range(3):
        model.encode_spatial_normalized(x, wvs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        z = model.encode_spatial_normalized(x, wvs)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"config5: batch {batch} 13x512x512 -> {tuple(z.shape)}: {ms:.2f} ms/step, {batch / ms * 1e3:.1f} patches/s, "
      f"{batch / ms * 1124.9:.0f} TFLOP/s algorithmic, peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
