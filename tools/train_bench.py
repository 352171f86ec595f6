"""BASELINE configs[2]: EOFluxVAE.training_step on S2L2A 12x256x256, batch 16 per GPU (forward, loss, backward, clip,
Adam).  Single process = 1 GPU; under torchrun every rank wraps the model in DistributedDataParallel (NCCL gradient
all-reduce overlapped with backward) and rank 0 prints the whole-job number (max over ranks).

usage: python tools/train_bench.py [batch] [steps] [msssim_start_step] [profile|graph]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "eo-vae_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import __graft_entry__ as g  # noqa: E402
from oracle.weights import FULL_CONFIG, WAVELENGTHS, make_state_dict  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ms_start = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
profile = len(sys.argv) > 4 and sys.argv[4] == "profile"
use_graph = len(sys.argv) > 4 and sys.argv[4] == "graph"
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=dev)

from eo_vae import ops  # noqa: E402
from eo_vae.models.modules.consistency_loss import EOConsistencyLoss  # noqa: E402

model = g._model(FULL_CONFIG, make_state_dict(FULL_CONFIG, 0), dev)
model.train()
model.loss_fn = EOConsistencyLoss(pixel_weight=1.0, rec_loss_type="char", msssim_weight=1.0, msssim_start_step=ms_start).to(dev)
model.clip_grad = 1.0
if world > 1:
    model.enable_ddp(int(os.environ.get("EOVAE_DDP_BUCKET_MB", "64")) << 20)
wvs = torch.tensor(WAVELENGTHS["S2L2A"], device=dev)
gen = torch.Generator(device=dev).manual_seed(1234 + rank)
x = torch.randn((batch, 12, 256, 256), device=dev, generator=gen).clamp_(-2, 6)
batch_d = {model.image_key: x, "wvs": wvs}


graphed = None
if use_graph:
    from eo_vae.graphs import GraphedTrainStep
    graphed = GraphedTrainStep(model, batch_d)


def step(i):
    return graphed(batch_d) if graphed is not None else model.training_step(batch_d, i)


for i in range(3):
    loss = step(i)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
l0 = ops.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.time()
e0.record()
for i in range(steps):
    loss = step(3 + i)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
ms = float(ms)
if rank == 0:
    pps = world * batch / ms * 1e3
    print(f"train_step{' (CUDA graph)' if use_graph else ''}: {world} GPU x batch {batch}: {ms:.1f} ms/step, {pps:.1f} patches/s, {pps * 2.695:.0f} TFLOP/s algorithmic "
          f"(2695 GF/patch), loss {float(loss):.4f}, {(ops.launch_count() - l0) // steps} kernel launches/step, peak mem "
          f"{torch.cuda.max_memory_allocated() / 2**30:.1f} GiB, wall {1e3 * (time.time() - t0) / steps:.1f} ms/step")
if profile and rank == 0:
    from torch.profiler import ProfilerActivity, profile as tprofile
    with tprofile(activities=[ProfilerActivity.CUDA]) as prof:
        step(100)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))
if world > 1:
    dist.destroy_process_group()
