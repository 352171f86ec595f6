"""Data-parallel gradient exchange for the training step: replicated parameters, one all-reduce
(average) of the gradients per step over ``torch.distributed`` (NCCL over NVLink on GPUs; gloo in the CPU tests).

Layout: the trainable parameters are cut into buckets in REVERSE registration order (~ the order backward produces
their gradients).  Each bucket owns ONE flat fp32 buffer allocated at construction; every parameter has a fixed,
16-byte aligned slot in it, whether or not it receives a gradient in a given step (a missing gradient is exchanged as
zeros), so the byte layout of every collective is identical on all ranks by construction.

Schedule: the convolution / weight-gradient kernels of the backward pass are persistent 148-CTA grids; an NCCL kernel
launched next to them either waits for a kernel boundary or delays a few of their CTAs by its whole duration, which
stretches the tail of every compute kernel it meets (round 1: 3.7 ms of a 60 ms step).  What the step offers instead is
a ~1.5 ms phase of tiny latency-bound kernels at the very END of backward - the encoder's input hypernetwork.  So the
buckets are packed as their gradients arrive (one multi-tensor copy each) but LAUNCHED together when only the last
bucket (``tail_bytes``: the parameters registered first, i.e. ``encoder.conv_in``) is outstanding; that last small
bucket goes out in ``finish()``.  ``finish()`` (called by ``manual_backward``) waits and hands the averaged buckets back
as gradient VIEWS (no copy back; the optimiser kernels read them in place).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class _Bucket:
    __slots__ = ("params", "offsets", "flat", "views", "used", "pending", "packed", "work")

    def __init__(self, params, device):
        self.params = params
        self.offsets, off = [], 0
        for p in params:
            self.offsets.append(off)
            off += (p.numel() + 3) // 4 * 4   # every slot starts on a 16-byte boundary
        # trailing len(params) floats: 1.0 where this rank produced a gradient - after the average, > 0 means "some rank
        # did", and then EVERY rank applies the averaged gradient (replicas must take identical optimiser steps)
        self.flat = torch.zeros((off + len(params),), dtype=torch.float32, device=device)
        self.views = [self.flat[o:o + p.numel()].view_as(p) for p, o in zip(params, self.offsets)]
        self.used = self.flat[off:]
        self.pending = len(params)
        self.packed = False
        self.work = None


class GradSync:
    def __init__(self, params, bucket_bytes: int = 64 << 20, group=None, tail_bytes: int = 24 << 20) -> None:
        self.group = group
        self.world = dist.get_world_size(group)
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise RuntimeError("GradSync: no trainable parameters")
        device = self.params[0].device
        # the LAST bucket to become ready = the parameters registered first, up to tail_bytes
        tail, size = [], 0
        for p in self.params:
            if tail and size + p.numel() * 4 > tail_bytes:
                break
            tail.append(p)
            size += p.numel() * 4
        groups, cur, size = [], [], 0
        for p in reversed(self.params[len(tail):]):
            cur.append(p)
            size += p.numel() * 4
            if size >= bucket_bytes:
                groups.append(cur)
                cur, size = [], 0
        if cur:
            groups.append(cur)
        groups.append(list(reversed(tail)))
        self.buckets = [_Bucket(g, device) for g in groups if g]
        self._bucket_of = {id(p): i for i, b in enumerate(self.buckets) for p in b.params}
        self._avg = dist.get_backend(self.group) == "nccl"   # gloo (CPU tests) has no AVG: sum, then divide in finish()
        self.overlap = True   # False: no hooks fire (graph replay) and finish() packs and launches every bucket itself
        self.launches = 0     # collectives issued so far (tests / bench bookkeeping)
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        self.broadcast_parameters()

    def broadcast_parameters(self, src: int = 0) -> None:
        """Replicas start from rank 0's parameters (DDP's constructor broadcast)."""
        with torch.no_grad():
            for p in self.params:
                dist.broadcast(p.data, src, group=self.group)
        # written through .data: advance the version counters that key the modules' derived 16-bit weight operands
        torch.autograd.graph.increment_version(self.params)

    # ------------------------------------------------------------------------------------------------------------
    def _pack(self, b: _Bucket) -> None:
        """Gradients -> their fixed slots (one multi-tensor copy); slots of parameters without a gradient are zeroed."""
        have = [(v, p.grad) for v, p in zip(b.views, b.params) if p.grad is not None and p.grad.data_ptr() != v.data_ptr()]
        with torch.no_grad():
            if have:
                torch._foreach_copy_([v for v, _ in have], [g.reshape(v.shape) for v, g in have])
            missing = [v for v, p in zip(b.views, b.params) if p.grad is None]
            if missing:
                torch._foreach_zero_(missing)
                b.used.copy_(torch.tensor([0.0 if p.grad is None else 1.0 for p in b.params]), non_blocking=True)
            else:
                b.used.fill_(1.0)
        b.packed = True

    def _launch(self, b: _Bucket) -> None:
        op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
        b.work = dist.all_reduce(b.flat, op=op, group=self.group, async_op=True)
        self.launches += 1

    def _on_grad(self, p) -> None:
        if not self.overlap:
            return
        i = self._bucket_of[id(p)]
        b = self.buckets[i]
        b.pending -= 1
        if b.pending != 0:
            return
        self._pack(b)
        # launch everything packed so far once only the tail bucket is outstanding (see the module docstring)
        if all(x.packed for x in self.buckets[:-1]):
            for x in self.buckets:
                if x.packed and x.work is None:
                    self._launch(x)

    def finish(self) -> None:
        """Pack / launch what is still outstanding, wait, average, hand the buckets back as gradient views."""
        for b in self.buckets:
            if not b.packed:
                self._pack(b)
        for b in self.buckets:
            if b.work is None:
                self._launch(b)
        for b in self.buckets:
            b.work.wait()
            if not self._avg:
                b.flat.div_(self.world)
            if all(p.grad is not None for p in b.params):   # the usual step: no host read of the usage flags
                for p, v in zip(b.params, b.views):
                    p.grad = v
            else:
                used = b.used.tolist()   # rare path (a parameter unused on this rank): one small device read
                for p, v, u in zip(b.params, b.views, used):
                    if p.grad is not None or u > 0.0:   # unused on EVERY rank: grad stays None, like DDP
                        p.grad = v
            b.work, b.packed, b.pending = None, False, len(b.params)

    def remove(self) -> None:
        for h in self._hooks:
            h.remove()
