"""``torch.optim.Adam`` whose ``step`` runs on the multi-tensor kernels of ``csrc/optimizer.cu`` with gradient-norm clipping
folded in (reference: ``torch.optim.Adam(params, lr)``, new_autoencoder.py:549-557, and
``clip_grad_norm_(params, clip_grad)`` + ``step()``, :650-657).

Same constructor, ``param_groups`` and ``state_dict`` layout as ``torch.optim.Adam`` (``step`` / ``exp_avg`` / ``exp_avg_sq``
per parameter), so LR schedulers and checkpoints are interchangeable.  ``step(clip_norm=c)`` is equivalent to
``clip_grad_norm_(params, c); step()`` except that the clip factor is applied on the fly: ``p.grad`` is left unscaled.
"""
from __future__ import annotations

import ctypes

import torch

from . import _C

_CHUNK = 1 << 16


class FusedClipAdam(torch.optim.Adam):
    fused_clip = True   # step(clip_norm=...) replaces clip_grad_norm_ + step()

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, **kw):
        if weight_decay != 0 or amsgrad or kw.get('maximize', False):
            raise NotImplementedError('FusedClipAdam implements plain Adam (no weight decay, amsgrad or maximize)')
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False, foreach=False, fused=False)
        self._tables = {}
        self._shared_step = {}
        self.last_grad_norm = None   # device scalar of the most recent clipped step (what clip_grad_norm_ returns)

    # ------------------------------------------------------------------------------------------------------------
    def _chunk_table(self, sizes, device):
        key = (tuple(sizes), str(device))
        tab = self._tables.get(key)
        if tab is None:
            ct, co = [], []
            for t, n in enumerate(sizes):
                for off in range(0, n, _CHUNK):
                    ct.append(t)
                    co.append(off)
            tab = (torch.tensor(ct, dtype=torch.int32, device=device), torch.tensor(co, dtype=torch.int64, device=device),
                   torch.tensor(sizes, dtype=torch.int64, device=device), len(ct))
            self._tables = {key: tab}
        return tab

    @torch.no_grad()
    def step(self, closure=None, clip_norm=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _C.lib()
        stream = None
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group['params'] if p.grad is not None]
            if not ps:
                continue
            beta1, beta2 = group['betas']
            gs, ms, vs, step_tensors = [], [], [], {}
            for p in ps:
                st = self.state[p]
                if len(st) == 0:
                    if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                        raise RuntimeError('FusedClipAdam: contiguous fp32 CUDA parameters only (no CPU path)')
                    # one shared CPU step counter per group: a single increment per step instead of one per parameter
                    st['step'] = self._shared_step.setdefault(gi, torch.tensor(0.0, dtype=torch.float32))
                    st['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                step_tensors[id(st['step'])] = st['step']
                g = p.grad
                if g.dtype != torch.float32 or g.is_sparse or not g.is_cuda:
                    raise RuntimeError('FusedClipAdam: dense fp32 CUDA gradients only')
                gs.append(g if g.is_contiguous() else g.contiguous())
                ms.append(st['exp_avg'])
                vs.append(st['exp_avg_sq'])
            for t in step_tensors.values():   # distinct tensors only after load_state_dict of a torch.optim.Adam checkpoint
                t += 1
            steps = {int(t) for t in step_tensors.values()}
            if len(steps) != 1:
                raise RuntimeError('FusedClipAdam: parameters of one group must share the step count')
            dev = ps[0].device
            if stream is None:
                stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            ct, co, sizes, nchunks = self._chunk_table([p.numel() for p in ps], dev)
            ptrs = torch.tensor([[t.data_ptr() for t in lst] for lst in (ps, gs, ms, vs)], dtype=torch.int64).to(dev, non_blocking=True)
            norm = None
            if clip_norm:
                partial = torch.empty((nchunks,), dtype=torch.float32, device=dev)
                norm = torch.empty((1,), dtype=torch.float32, device=dev)
                _C.check(lib.eovae_grad_norm(ptrs[1].data_ptr(), sizes.data_ptr(), ct.data_ptr(), co.data_ptr(), nchunks, _CHUNK,
                                             partial.data_ptr(), norm.data_ptr(), stream), 'eovae_grad_norm')
                self.last_grad_norm = norm
            _C.check(lib.eovae_adam_step(ptrs[0].data_ptr(), ptrs[1].data_ptr(), ptrs[2].data_ptr(), ptrs[3].data_ptr(),
                                         sizes.data_ptr(), ct.data_ptr(), co.data_ptr(), nchunks, _CHUNK, float(group['lr']),
                                         float(beta1), float(beta2), float(group['eps']), steps.pop(),
                                         None if norm is None else norm.data_ptr(), float(clip_norm or 0.0), stream),
                     'eovae_adam_step')
            self._keepalive = (ptrs, gs)   # the tables / gradient copies must outlive the asynchronous launches
        return loss
