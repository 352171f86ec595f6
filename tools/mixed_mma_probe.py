"""Does tcgen05 kind::f16 accept an f16 A operand with a bf16 B operand?  (EOVAE_ALLOW_MIXED_MMA=1 lifts the host-side check.)
Result on B200 (round 2): see profiles/r2_mixed_mma_probe.txt."""
import os
import sys

os.environ["EOVAE_ALLOW_MIXED_MMA"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "eo-vae_b200"))
import torch  # noqa: E402

from eo_vae import ops  # noqa: E402

dev = torch.device("cuda:0")
a = torch.randn(2, 256, 128, device=dev)
b = torch.randn(2, 128, 128, device=dev)
for da, db in ((torch.float16, torch.float16), (torch.bfloat16, torch.bfloat16), (torch.float16, torch.bfloat16),
               (torch.bfloat16, torch.float16)):
    try:
        c = ops.gemm_tn_batched(a.to(da), b.to(db), torch.float32)
        torch.cuda.synchronize()
        ref = a.to(da).float() @ b.to(db).float().transpose(1, 2)
        print(f"A {da} x B {db}: ok, rel err {float((c - ref).norm() / ref.norm()):.2e}", flush=True)
    except Exception as exc:  # noqa: BLE001
        print(f"A {da} x B {db}: FAILED {type(exc).__name__}: {str(exc)[:160]}", flush=True)
        break
