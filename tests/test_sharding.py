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
