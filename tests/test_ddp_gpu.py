"""On-hardware check of the gradient exchange: after ``manual_backward`` on N >= 2 GPUs every rank holds the MEAN
of the per-rank gradients, i.e. what one process computes by running each rank's batch in turn and averaging (the latent
BatchNorm uses per-rank batch statistics - the reference does not use SyncBN, new_autoencoder.py:125 - so the comparison
is with the averaged per-shard gradients, not with one pass over the concatenated batch).  Skipped with fewer than 2 GPUs.

Run on a multi-GPU box:  python -m pytest tests/test_ddp_gpu.py -m gpu   (spawns one process per GPU, NCCL)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(dev):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "eo-vae_b200"))
    import __graft_entry__ as g
    from eo_vae.models.modules.consistency_loss import EOConsistencyLoss
    from oracle.weights import TINY_CONFIG, make_state_dict
    model = g._model(TINY_CONFIG, make_state_dict(TINY_CONFIG, 3), dev)
    model.train()
    model.loss_fn = EOConsistencyLoss(pixel_weight=1.0, rec_loss_type="char").to(dev)
    return model, TINY_CONFIG


def _grads(model, x, wvs, seed):
    for p in model.parameters():
        p.grad = None
    torch.manual_seed(seed)            # the posterior noise is drawn on the CPU generator (distributions.py:44)
    recon, _ = model(x, wvs)
    loss, _ = model.loss_fn(inputs=x, wvs=wvs, reconstructions=recon, global_step=0)
    model.manual_backward(loss)
    return [p.grad.detach().clone().float() for p in model.parameters()]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from oracle.weights import WAVELENGTHS, synthetic_patches
    model, cfg = _build(dev)
    wvs = torch.tensor(WAVELENGTHS["S2L2A"], device=dev)
    shards = [synthetic_patches(2, 12, cfg["resolution"], seed=50 + r).to(dev) for r in range(world)]
    # single-process reference on this rank: every shard in turn (same per-shard noise seeds), averaged
    bn0 = {k: v.clone() for k, v in model.bn.state_dict().items()}
    want = None
    for r in range(world):
        model.bn.load_state_dict(bn0)
        gs = _grads(model, shards[r], wvs, 1000 + r)
        want = gs if want is None else [a + b for a, b in zip(want, gs)]
    want = [w / world for w in want]
    model.bn.load_state_dict(bn0)
    model.enable_ddp(bucket_bytes=1 << 16, tail_bytes=1 << 14)     # several buckets on the tiny model
    got = _grads(model, shards[rank], wvs, 1000 + rank)
    num = torch.sqrt(sum(((a - b) ** 2).sum() for a, b in zip(got, want)))
    den = torch.sqrt(sum((b ** 2).sum() for b in want))
    # every rank must also hold the SAME bits
    flat = torch.cat([t.flatten() for t in got])
    ref = flat.clone()
    dist.broadcast(ref, 0)
    torch.save({"rel": float(num / den), "same": bool(torch.equal(flat, ref)), "buckets": len(model._grad_sync.buckets)},
               os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_exchanged_gradients_equal_mean_of_shard_gradients(tmp_path):
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        res = torch.load(os.path.join(tmp_path, f"r{r}.pt"))
        print(f"rank {r}: exchanged vs mean-of-shards gradient rel-L2 {res['rel']:.3e}, {res['buckets']} buckets")
        assert res["buckets"] >= 2
        assert res["same"], "ranks hold different averaged gradients"
        # same kernels, same inputs: only the fp32 summation order of the average differs
        assert res["rel"] < 1e-5, res["rel"]
