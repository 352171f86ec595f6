"""One warm-up + one measured encode step of the headline workload (batch 64, S2L2A 256x256, bf16): the command that
is run plain and then under ncu for profiles/ (launch list + full capture of the implicit-GEMM kernel)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "eo-vae_b200"))
import torch  # noqa: E402

import __graft_entry__ as g  # noqa: E402
from oracle.weights import FULL_CONFIG, WAVELENGTHS, make_state_dict  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
what = sys.argv[3] if len(sys.argv) > 3 else "encode"  # encode | reconstruct
dev = torch.device("cuda:0")
model = g._model(FULL_CONFIG, make_state_dict(FULL_CONFIG, 0), dev)
wvs = torch.tensor(WAVELENGTHS["S2L2A"], device=dev)
x = torch.randn((batch, 12, 256, 256), device=dev).clamp_(-2, 6)
with torch.no_grad():
    for i in range(steps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        z = model.encode_spatial_normalized(x, wvs) if what == "encode" else model.reconstruct(x, wvs)
        e1.record()
        torch.cuda.synchronize()
        print(f"step {i}: {e0.elapsed_time(e1):.3f} ms, {what} out {tuple(z.shape)} finite={bool(torch.isfinite(z).all())}")
