"""BASELINE configs[3]: mixed-modality training - every step (and every rank) draws its modality from {S2L2A (12 bands), S1RTC (2),
S2RGB (3)} with ``random.Random(seed + rank)``, so the DynamicConv layers see per-batch band counts and, under torchrun, ranks
exchange gradients of steps run on different modalities (parameter shapes do not depend on the band count).

usage: [torchrun ...] python tools/mixed_bench.py [batch] [steps]"""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "eo-vae_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import __graft_entry__ as g  # noqa: E402
from oracle.weights import FULL_CONFIG, WAVELENGTHS, make_state_dict  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/eovae_nccl.%h.%p.log")
    dist.init_process_group("nccl", device_id=dev)

from eo_vae.models.modules.consistency_loss import EOConsistencyLoss  # noqa: E402

model = g._model(FULL_CONFIG, make_state_dict(FULL_CONFIG, 0), dev)
model.train()
model.loss_fn = EOConsistencyLoss(pixel_weight=1.0, rec_loss_type="char", msssim_weight=1.0, msssim_start_step=0).to(dev)
model.clip_grad = 1.0
if world > 1:
    model.enable_ddp()
rng = random.Random(1234 + rank)
gen = torch.Generator(device=dev).manual_seed(1234 + rank)
mods = ["S2L2A", "S1RTC", "S2RGB"]
batches = {m: {model.image_key: torch.randn((batch, len(WAVELENGTHS[m]), 256, 256), device=dev, generator=gen).clamp_(-2, 6),
               "wvs": torch.tensor(WAVELENGTHS[m], device=dev)} for m in mods}
for m in mods:  # warm-up: every modality once (weight-operand caches, allocator)
    model.training_step(batches[m], 0)
seq = [rng.choice(mods) for _ in range(steps)]
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i, m in enumerate(seq):
    loss = model.training_step(batches[m], i)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"mixed-modality train_step: {world} GPU x batch {batch}: {float(ms):.1f} ms/step, {world * batch / float(ms) * 1e3:.1f} patches/s "
          f"(rank-0 modality sequence {''.join(m[1:3] + ' ' for m in seq)}; loss {float(loss.detach()):.4f})")
if world > 1:
    dist.destroy_process_group()
