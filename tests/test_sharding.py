"""CPU suite, part 4: the multi-GPU path of the encode workload is a pure partition of independent patches
(no data-path collective).  Two gloo ranks shard a patch list, encode their shards with the oracle and the gathered
result equals the single-process result; the max-over-ranks timing reduction used by bench.py is exercised too."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import eovae_oracle as O
from oracle.weights import TINY_CONFIG, WAVELENGTHS, make_state_dict, synthetic_patches


def shard_indices(n_items: int, rank: int, world: int):
    """Dataset index i -> rank i mod world."""
    return list(range(rank, n_items, world))


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    sd = make_state_dict(TINY_CONFIG, 4)
    wvs = torch.tensor(WAVELENGTHS["S2L2A"])
    x = synthetic_patches(6, 12, 32, seed=9)
    idx = shard_indices(6, rank, world)
    with torch.no_grad():
        z = O.encode_spatial_normalized(sd, x[idx], wvs, TINY_CONFIG["hyper_heads"])
    gathered = [torch.zeros_like(z) for _ in range(world)]
    dist.all_gather(gathered, z)              # result collection only (not part of the timed data path)
    ms = torch.tensor([10.0 + rank])          # bench.py: step time = MAX over ranks
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        full = torch.zeros(6, *z.shape[1:])
        for r in range(world):
            full[shard_indices(6, r, world)] = gathered[r]
        torch.save({"z": full, "ms": ms}, os.path.join(out_dir, "out.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_encode_equals_single_process(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    out = torch.load(os.path.join(tmp_path, "out.pt"))
    sd = make_state_dict(TINY_CONFIG, 4)
    wvs = torch.tensor(WAVELENGTHS["S2L2A"])
    x = synthetic_patches(6, 12, 32, seed=9)
    with torch.no_grad():
        ref = O.encode_spatial_normalized(sd, x, wvs, TINY_CONFIG["hyper_heads"])
    assert torch.allclose(out["z"], ref, atol=1e-6)
    assert float(out["ms"]) == 11.0


def test_shard_indices_partition():
    for world in (1, 2, 4, 8):
        seen = sorted(i for r in range(world) for i in shard_indices(37, r, world))
        assert seen == list(range(37))


def _ddp_worker(rank, world, port, out_dir):
    """Two replicas, different data: after GradSync the gradients equal the mean of the per-rank gradients, parameters
    start from rank 0's values, and a parameter that gets no gradient (unused branch) does not stall the exchange."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "eo-vae_b200"))
    from eo_vae.ddp import GradSync
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)                      # replicas start DIFFERENT; the constructor broadcast fixes that
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Tanh(), torch.nn.Linear(16, 4))
    unused = torch.nn.Parameter(torch.ones(3))         # no gradient on any rank
    partial = torch.nn.Parameter(torch.ones(5))        # gradient on rank 1 only (a branch only that rank took)
    sync = GradSync(list(net.parameters()) + [unused, partial], bucket_bytes=256, tail_bytes=300)   # several small buckets
    assert len(sync.buckets) >= 3
    x = torch.randn(5, 8, generator=torch.Generator().manual_seed(rank))
    for _ in range(2):                                 # two steps: the bucket bookkeeping resets in finish()
        for p in list(net.parameters()) + [unused, partial]:
            p.grad = None
        loss = net(x).pow(2).sum()
        if rank == 1:
            loss = loss + (partial * 2.0).sum()
        loss.backward()
        sync.finish()
    assert unused.grad is None                         # unused everywhere: stays None, like DDP
    assert torch.allclose(partial.grad, torch.full((5,), 1.0))   # (0 + 2) / 2 on BOTH ranks: replicas step identically
    assert sync.launches == 2 * len(sync.buckets)
    torch.save({"grads": [p.grad.clone() for p in net.parameters()], "params": [p.detach().clone() for p in net.parameters()],
                "x": x}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_sync(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_ddp_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (torch.load(os.path.join(tmp_path, f"r{r}.pt")) for r in range(2))
    for a, b in zip(r0["params"], r1["params"]):
        assert torch.equal(a, b)                       # broadcast from rank 0
    for a, b in zip(r0["grads"], r1["grads"]):
        assert torch.equal(a, b)                       # both ranks hold the same averaged gradient
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Tanh(), torch.nn.Linear(16, 4))
    with torch.no_grad():
        for p, v in zip(net.parameters(), r0["params"]):
            p.copy_(v)
    want = None
    for r in (r0, r1):
        for p in net.parameters():
            p.grad = None
        net(r["x"]).pow(2).sum().backward()
        gs = [p.grad.clone() for p in net.parameters()]
        want = gs if want is None else [w + g for w, g in zip(want, gs)]
    for got, w in zip(r0["grads"], want):
        assert torch.allclose(got, w / 2, atol=1e-6)


def _stats_worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "eo-vae_b200"))
    from eo_vae.encode_latents import RunningStatsButFast
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x = torch.randn(6, 4, 8, 8, generator=torch.Generator().manual_seed(3)) * 2 + 1
    st = O.running_stats_init(4)
    for i in range(rank, 6, world):                    # this rank's shard, one item per update
        st = O.running_stats_update(st, x[i:i + 1])
    rs = RunningStatsButFast((4,), [0, 2, 3])
    for k in ("mean", "var", "std", "count", "min", "max"):
        getattr(rs, k).copy_(st[k])
    rs.merge_ranks()
    if rank == 0:
        torch.save(rs.get_stats_dict(), os.path.join(out_dir, "stats.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_latent_statistics_merge(tmp_path):
    """per-rank running statistics of a sharded encode merge to the single-process statistics."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_stats_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = torch.load(os.path.join(tmp_path, "stats.pt"))
    x = torch.randn(6, 4, 8, 8, generator=torch.Generator().manual_seed(3)) * 2 + 1
    st = O.running_stats_init(4)
    for i in range(6):
        st = O.running_stats_update(st, x[i:i + 1])
    for k in ("mean", "var", "std", "min", "max", "count"):
        assert torch.allclose(got[k], st[k], rtol=1e-5, atol=1e-6), k
