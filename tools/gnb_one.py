"""One GroupNorm(+SiLU) backward on the level-0 training tensor (16 x 256 x 256 x 128, bf16): the command profiled for
profiles/r2_gn_bwd_bulk.md (plain, then under ncu --set full -k regex:gn_bwd)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "eo-vae_b200"))
import torch  # noqa: E402

from eo_vae import ops  # noqa: E402

dev = torch.device("cuda:0")
n, h, w, c = 16, 256, 256, 128
x = torch.randn((n, h, w, c), device=dev).bfloat16().permute(0, 3, 1, 2)
g = torch.randn((n, h, w, c), device=dev).bfloat16().permute(0, 3, 1, 2)
gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
stats = ops.gn_stats(x)
if len(sys.argv) > 1:
    ops.set_tuning(ops.TUNE_GN_BWD_BULK, int(sys.argv[1]))
for _ in range(3):
    ops.gn_backward(x, g, stats, gamma, beta, True)
torch.cuda.synchronize()
print("ok")
