"""Data-parallel gradient exchange for the training step: replicated parameters, one all-reduce
(sum, then / world) of the gradients per step, bucketed in reverse registration order (~ the order backward produces
them) and launched asynchronously from post-accumulate hooks so NCCL traffic over NVLink overlaps the remaining
backward kernels.  ``finish()`` (called by ``manual_backward``) waits and hands the averaged buckets back as gradient views (no copy).

Plumbing only: ``torch.distributed`` does the transport (NCCL on GPUs; gloo in the CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


class GradSync:
    def __init__(self, params, bucket_bytes: int = 64 << 20, group=None) -> None:
        self.group = group
        self.world = dist.get_world_size(group)
        self.params = [p for p in params if p.requires_grad]
        self.buckets: list[list[torch.nn.Parameter]] = []
        cur, size = [], 0
        for p in reversed(self.params):
            cur.append(p)
            size += p.numel() * 4
            if size >= bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
        if cur:
            self.buckets.append(cur)
        self._bucket_of = {id(p): i for i, b in enumerate(self.buckets) for p in b}
        self._pending = [len(b) for b in self.buckets]
        self._inflight: list = []
        self._launched = [False] * len(self.buckets)
        self.overlap = True   # False: no hooks fire (graph replay) and finish() launches every bucket itself
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        self.broadcast_parameters()

    def broadcast_parameters(self, src: int = 0) -> None:
        """Replicas start from rank 0's parameters (DDP's constructor broadcast)."""
        with torch.no_grad():
            for p in self.params:
                dist.broadcast(p.data, src, group=self.group)

    def _launch(self, i: int) -> None:
        ps = [p for p in self.buckets[i] if p.grad is not None]
        self._launched[i] = True
        if not ps:
            return
        # one flat fp32 buffer per bucket; every gradient starts on a 16-byte boundary (zero pads in between) so the views
        # handed back in finish() keep the vectorised path of the optimiser kernels
        pieces, offs, off = [], [], 0
        for p in ps:
            n = p.numel()
            pieces.append(p.grad.reshape(-1).to(torch.float32))
            offs.append(off)
            off += n
            pad = (-n) % 4
            if pad:
                pieces.append(torch.zeros((pad,), dtype=torch.float32, device=p.grad.device))
                off += pad
        flat = torch.cat(pieces)
        # NCCL averages inside the collective; gloo (CPU tests) has no AVG: sum, then divide in finish()
        avg = dist.get_backend(self.group) == "nccl"
        work = dist.all_reduce(flat, op=dist.ReduceOp.AVG if avg else dist.ReduceOp.SUM, group=self.group, async_op=True)
        self._inflight.append((work, flat, ps, offs, avg))

    def _on_grad(self, p) -> None:
        if not self.overlap:
            return
        i = self._bucket_of[id(p)]
        self._pending[i] -= 1
        if self._pending[i] == 0:
            self._launch(i)

    def finish(self) -> None:
        """Flush buckets whose parameters did not all receive a gradient, wait, average, write back."""
        for i in range(len(self.buckets)):
            if not self._launched[i]:
                self._launch(i)
        for work, flat, ps, offs, avg in self._inflight:
            work.wait()
            if not avg:
                flat.div_(self.world)
            for p, off in zip(ps, offs):   # the averaged gradients are VIEWS of the bucket: no copy back
                p.grad = flat[off:off + p.numel()].view_as(p)
        self._inflight.clear()
        self._pending = [len(b) for b in self.buckets]
        self._launched = [False] * len(self.buckets)

    def remove(self) -> None:
        for h in self._hooks:
            h.remove()
