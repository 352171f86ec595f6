"""Latent-encoding pipeline around the encoder: mirror of the reference ``encode_latents.py``
(``RunningStatsButFast`` :36-109, ``encode_raw`` :139-156, ``encode_spatial_norm`` :159-168, ``decode_raw`` :171-186,
``decode_spatial_norm`` :189-196, ``encode_split`` :305-352), re-designed for a multi-GPU box:

* the running per-channel statistics live on the device and are merged by one kernel per batch
  (``eovae_running_stats_update``), so the encode loop never synchronises with the host;
* ranks encode disjoint shards (item ``i`` -> rank ``i mod world``) and merge their statistics once at the end
  (``RunningStatsButFast.merge_ranks``: count / mean / M2 / min / max, parallel-variance formula, rank order -> deterministic);
* ``.npz`` files are written by a background thread from pinned host buffers filled by asynchronous device->host copies
  on a side stream, so ``np.savez_compressed`` (CPU zlib) is off the GPU's critical path (the reference serialises it per
  sample inside the loop).
"""
from __future__ import annotations

import os
import queue
import threading

import numpy as np
import torch
import torch.distributed as dist

from . import _C, ops


class RunningStatsButFast(torch.nn.Module):
    """Drop-in for the reference class: same buffers (``mean``, ``var``, ``std``, ``count``, ``min``, ``max``), same update
    formulas (batch variance is torch's unbiased one), ``forward(x) -> x``, ``get_stats_dict()``.  Built for latents
    ``[B, C, H, W]`` reduced over ``dims = [0, 2, 3]``."""

    def __init__(self, shape, dims):
        super().__init__()
        shape = tuple(shape) if not isinstance(shape, int) else (shape,)
        if len(shape) != 1 or list(dims) != [0, 2, 3]:
            raise NotImplementedError('RunningStatsButFast is built for per-channel statistics of [B, C, H, W] (dims [0, 2, 3])')
        self.register_buffer('mean', torch.zeros(shape))
        self.register_buffer('var', torch.ones(shape))
        self.register_buffer('std', torch.ones(shape))
        self.register_buffer('count', torch.zeros(1))
        self.register_buffer('min', torch.full(shape, float('inf')))
        self.register_buffer('max', torch.full(shape, float('-inf')))
        self.dims = list(dims)

    @torch.no_grad()
    def update(self, x: torch.Tensor) -> None:
        if not x.is_cuda:
            raise RuntimeError('RunningStatsButFast.update: CUDA tensor required (no CPU path)')
        if self.mean.device != x.device:
            self.to(x.device)
        x = x.to(torch.float32).contiguous()
        b, c, h, w = x.shape
        ws = torch.empty((4 * c,), dtype=torch.float32, device=x.device)
        rc = _C.lib().eovae_running_stats_update(x.data_ptr(), b, c, h * w, self.mean.data_ptr(), self.var.data_ptr(),
                                                 self.std.data_ptr(), self.count.data_ptr(), self.min.data_ptr(),
                                                 self.max.data_ptr(), ws.data_ptr(), torch.cuda.current_stream().cuda_stream)
        _C.check(rc, 'eovae_running_stats_update')

    def forward(self, x):
        self.update(x)
        return x

    def get_stats_dict(self):
        return {k: getattr(self, k).cpu() for k in ('mean', 'std', 'var', 'min', 'max', 'count')}

    # ---- multi-GPU: one tiny exchange at the end of the run ------------------------------------------------------
    @staticmethod
    def merge_states(states: list) -> dict:
        """Parallel-variance merge of per-rank states (dicts with mean / var / count / min / max), in list order."""
        out = {k: states[0][k].clone().double() for k in ('mean', 'var', 'count', 'min', 'max')}
        for st in states[1:]:
            na, nb = out['count'], st['count'].double()
            if float(nb) == 0.0:
                continue
            n = na + nb
            delta = st['mean'].double() - out['mean']
            m2 = out['var'] * na + st['var'].double() * nb + delta ** 2 * na * nb / n
            out['mean'] = (out['mean'] * na + st['mean'].double() * nb) / n
            out['var'] = m2 / n
            out['count'] = n
            out['min'] = torch.minimum(out['min'], st['min'].double())
            out['max'] = torch.maximum(out['max'], st['max'].double())
        out = {k: v.float() for k, v in out.items()}
        out['std'] = torch.sqrt(out['var'] + 1e-8)
        return out

    @torch.no_grad()
    def merge_ranks(self, group=None) -> None:
        """All ranks end up with the statistics of the whole (sharded) dataset."""
        world = dist.get_world_size(group)
        c = self.mean.numel()
        mine = torch.cat([self.mean, self.var, self.min, self.max, self.count]).float()
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine, group=group)
        states = [{'mean': g[:c].cpu(), 'var': g[c:2 * c].cpu(), 'min': g[2 * c:3 * c].cpu(), 'max': g[3 * c:4 * c].cpu(),
                   'count': g[4 * c:].cpu()} for g in gathered]
        merged = self.merge_states(states)
        for k in ('mean', 'var', 'std', 'min', 'max', 'count'):
            getattr(self, k).copy_(merged[k].to(self.mean.device))


# ------------------------------------------------------------------------------------------------------ encode / decode
@torch.no_grad()
def encode_raw(model, img, wvs):
    """RAW latent = posterior mean [B, z, H/8, W/8] (no shuffle, no BatchNorm)."""
    if not hasattr(model, 'encoder'):
        raise ValueError(f'Unknown model type: {type(model)}')
    moments = model.encoder.moments_nhwc(img, wvs)
    return ops.act_to_nchw_f32(moments, moments.shape[1] // 2)


@torch.no_grad()
def encode_spatial_norm(model, img, wvs):
    if hasattr(model, 'encode_spatial_normalized'):
        return model.encode_spatial_normalized(img, wvs)
    raise ValueError('Model does not support encode_spatial_normalized method')


@torch.no_grad()
def decode_raw(model, z, wvs):
    if not hasattr(model, 'decoder'):
        raise ValueError(f'Unknown model type: {type(model)}')
    return model.decoder(z, wvs)


@torch.no_grad()
def decode_spatial_norm(model, z, wvs):
    if hasattr(model, 'decode_spatial_normalized'):
        return model.decode_spatial_normalized(z, wvs)
    raise ValueError('Model does not support decode_spatial_normalized method')


# ------------------------------------------------------------------------------------------------------ async writer
class LatentWriter:
    """``submit(path, **device_tensors)``: device->pinned-host copies are queued on a side stream, a worker thread waits
    for the copy event and runs ``np.savez_compressed``.  ``close()`` drains the queue."""

    def __init__(self, device, max_pending: int = 64):
        self.stream = torch.cuda.Stream(device)
        self.q: queue.Queue = queue.Queue(maxsize=max_pending)
        self.error = None
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        while True:
            item = self.q.get()
            if item is None:
                return
            path, host, event = item
            try:
                event.synchronize()
                np.savez_compressed(path, **{k: v.numpy() for k, v in host.items()})
            except Exception as exc:  # noqa: BLE001 - surfaced by close()
                self.error = exc

    def submit(self, path: str, **tensors) -> None:
        self.stream.wait_stream(torch.cuda.current_stream())
        host = {}
        with torch.cuda.stream(self.stream):
            for k, t in tensors.items():
                t = t.detach()
                buf = torch.empty(t.shape, dtype=t.dtype).pin_memory()
                buf.copy_(t, non_blocking=True)
                t.record_stream(self.stream)
                host[k] = buf
            event = torch.cuda.Event()
            event.record(self.stream)
        self.q.put((path, host, event))

    def close(self) -> None:
        self.q.put(None)
        self.thread.join()
        if self.error is not None:
            raise self.error


def shard_batches(batches, rank: int, world: int):
    """Batch i -> rank i mod world (no data-path collective)."""
    for i, batch in enumerate(batches):
        if i % world == rank:
            yield batch


def encode_split(model, dataloader, output_dir, device, wvs_lr, wvs_hr, stats_lr, stats_hr, split_name,
                 encode_fn=encode_raw, writer: LatentWriter | None = None, rank: int = 0, world: int = 1):
    """Reference ``encode_split`` (:305-352) with the same arguments and file format; ``rank`` / ``world`` shard the
    loader's batches, ``writer`` (created here if absent) takes the file output off the critical path."""
    os.makedirs(output_dir, exist_ok=True)
    own = writer is None
    writer = writer or LatentWriter(device)
    for batch in shard_batches(dataloader, rank, world):
        lr_img = batch['image_lr'].to(device, non_blocking=True)
        hr_img = batch['image_hr'].to(device, non_blocking=True)
        z_lr = encode_fn(model, lr_img, wvs_lr)
        z_hr = encode_fn(model, hr_img, wvs_hr)
        stats_lr(z_lr)
        stats_hr(z_hr)
        for i, aoi_id in enumerate(batch['aoi']):
            writer.submit(os.path.join(output_dir, f'{aoi_id}.npz'), lr_latent=z_lr[i], hr_latent=z_hr[i],
                          lr_image=lr_img[i], hr_image=hr_img[i])
    if own:
        writer.close()
