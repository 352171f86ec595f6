"""``torch.optim.Adam`` whose ``step`` runs on the multi-tensor kernels of ``csrc/optimizer.cu`` with gradient-norm clipping
folded in (reference: ``torch.optim.Adam(params, lr)``, new_autoencoder.py:549-557, and
``clip_grad_norm_(params, clip_grad)`` + ``step()``, :650-657).

Same constructor, ``param_groups`` and ``state_dict`` layout as ``torch.optim.Adam`` (``step`` / ``exp_avg`` / ``exp_avg_sq``
per parameter), so LR schedulers and checkpoints are interchangeable.  ``step(clip_norm=c)`` is equivalent to
``clip_grad_norm_(params, c); step()`` except that the clip factor is applied on the fly: ``p.grad`` is left unscaled.

Scope notes: the clipping norm is taken over every parameter of every group that has a gradient (the reference step clips
``param_groups[0]['params']`` of a single-group optimiser - identical there); all parameters of a group share one step
counter, so a parameter that receives its first gradient at step k uses bias correction k rather than 1 (the reference
model has no such parameter: frozen ones never get a gradient, trainable ones always do).  After every update the
parameters' autograd version counters are advanced, because the derived 16-bit weight-operand caches of the modules are
keyed on them.
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
            self._tables[key] = tab
        return tab

    def _group_tensors(self, gi, group):
        """-> (params with a gradient, gradients, exp_avg, exp_avg_sq, step number of this update) for one group."""
        ps = [p for p in group['params'] if p.grad is not None]
        gs, ms, vs, step_tensors = [], [], [], {}
        for p in ps:
            st = self.state[p]
            if len(st) == 0:
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError('FusedClipAdam: contiguous fp32 CUDA parameters only (no CPU path)')
                # one shared CPU step counter per group: a single increment per step instead of one per parameter.  After
                # load_state_dict the loaded per-parameter counters are adopted (they all hold the same step number).
                shared = self._shared_step.get(gi)
                if shared is None:
                    loaded = [self.state[q]['step'] for q in group['params'] if 'step' in self.state.get(q, {})]
                    shared = loaded[0] if loaded else torch.tensor(0.0, dtype=torch.float32)
                    self._shared_step[gi] = shared
                st['step'] = shared
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
        if len(steps) > 1:
            raise RuntimeError('FusedClipAdam: parameters of one group must share the step count')
        return ps, gs, ms, vs, (steps.pop() if steps else 0)

    @torch.no_grad()
    def step(self, closure=None, clip_norm=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _C.lib()
        groups = [(group,) + self._group_tensors(gi, group) for gi, group in enumerate(self.param_groups)]
        groups = [g for g in groups if g[1]]
        if not groups:
            return loss
        dev = groups[0][1][0].device
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        keep = []
        norm = None
        if clip_norm:
            # clip_grad_norm_ semantics: ONE norm over every parameter handed to step(), whatever group it sits in
            all_gs = [g for grp in groups for g in grp[2]]
            ct, co, sizes, nchunks = self._chunk_table([g.numel() for g in all_gs], dev)
            gptrs = torch.tensor([g.data_ptr() for g in all_gs], dtype=torch.int64).to(dev, non_blocking=True)
            partial = torch.empty((nchunks,), dtype=torch.float32, device=dev)
            norm = torch.empty((1,), dtype=torch.float32, device=dev)
            _C.check(lib.eovae_grad_norm(gptrs.data_ptr(), sizes.data_ptr(), ct.data_ptr(), co.data_ptr(), nchunks, _CHUNK,
                                         partial.data_ptr(), norm.data_ptr(), stream), 'eovae_grad_norm')
            self.last_grad_norm = norm
            keep.append((gptrs, partial))
        for group, ps, gs, ms, vs, step_no in groups:
            beta1, beta2 = group['betas']
            ct, co, sizes, nchunks = self._chunk_table([p.numel() for p in ps], dev)
            ptrs = torch.tensor([[t.data_ptr() for t in lst] for lst in (ps, gs, ms, vs)], dtype=torch.int64).to(dev, non_blocking=True)
            _C.check(lib.eovae_adam_step(ptrs[0].data_ptr(), ptrs[1].data_ptr(), ptrs[2].data_ptr(), ptrs[3].data_ptr(),
                                         sizes.data_ptr(), ct.data_ptr(), co.data_ptr(), nchunks, _CHUNK, float(group['lr']),
                                         float(beta1), float(beta2), float(group['eps']), step_no,
                                         None if norm is None else norm.data_ptr(), float(clip_norm or 0.0), stream),
                     'eovae_adam_step')
            # the kernels wrote through raw pointers: tell autograd's version counters, which key the derived 16-bit weight
            # operand caches (Conv2dSM100.packed_weight, ResnetBlock._conv2_with_shortcut, AttnBlock._qkv_operands)
            torch.autograd.graph.increment_version(ps)
            keep.append((ptrs, gs))
        self._keepalive = keep   # the tables / gradient copies must outlive the asynchronous launches
        return loss
