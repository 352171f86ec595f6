"""Flux2-AE building blocks on the sm_100a kernels.

Mirrors the reference interface of ``eo_vae/models/modules/layers.py`` (Normalize :14, swish :21, Downsample :25,
Upsample :40, ResnetBlock :53, AttnBlock :117): same constructor arguments, parameter names and shapes (OIHW fp32
master weights), same forward signatures.  Forward math runs in: GroupNorm statistics (warp-shuffle reduction) ->
normalise+affine+SiLU -> tcgen05 implicit-GEMM conv with bias / residual fused in the epilogue.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
from torch import Tensor

from ... import autograd as tape
from ... import ops
from ...settings import compute_dtype


def Normalize(in_channels: int, num_groups: int = 32) -> nn.Module:
    return GroupNormSM100(num_groups=num_groups, num_channels=in_channels, eps=1e-6, affine=True)


def swish(x: Tensor) -> Tensor:
    """x * sigmoid(x).  Inside the blocks SiLU is fused into the GroupNorm-apply kernel; this free function only
    exists for API parity and works on any tensor."""
    return x * torch.sigmoid(x)


class GroupNormSM100(nn.GroupNorm):
    """nn.GroupNorm parameter container whose forward runs eovae_gn_stats + eovae_gn_apply."""

    def forward(self, x: Tensor, silu: bool = False) -> Tensor:  # noqa: D102
        x = ops.to_act(x, compute_dtype())
        if tape.grad_mode():
            return tape.GroupNormFn.apply(x, self.weight, self.bias, silu, self.num_groups, self.eps)
        return ops.group_norm(x, self.weight, self.bias, silu, self.num_groups, self.eps)


class _ModulatedNorm:
    """GroupNormSM100 stand-in whose affine is the AdaIN-modulated (gamma', beta') of one forward call."""

    def __init__(self, norm: GroupNormSM100, weight: Tensor, bias: Tensor):
        self.weight, self.bias, self.num_groups, self.eps = weight, bias, norm.num_groups, norm.eps

    def __call__(self, x: Tensor, silu: bool = False) -> Tensor:
        return ops.group_norm(ops.to_act(x, compute_dtype()), self.weight, self.bias, silu, self.num_groups, self.eps)


class Conv2dSM100(nn.Conv2d):
    """nn.Conv2d parameter container (same init, same state_dict keys) executed by eovae_conv2d.

    The K-major 16-bit operand is a derived cache keyed on the parameter version, so optimiser steps and
    ``load_state_dict`` invalidate it automatically."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0):
        super().__init__(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding)
        k = self.kernel_size[0]
        if (k, self.stride[0], self.padding[0]) == (3, 1, 1):
            self._mode = ops.CONV_3X3
        elif (k, self.stride[0], self.padding[0]) == (1, 1, 0):
            self._mode = ops.CONV_1X1
        elif (k, self.stride[0], self.padding[0]) == (3, 2, 0):
            self._mode = ops.CONV_3X3_S2  # caller semantics: F.pad(0,1,0,1) first (Downsample)
        else:
            raise ValueError("Conv2dSM100 supports 3x3/s1/p1, 1x1 and the Downsample 3x3/s2/p0 convolution only")
        self._packed = None
        self._packed_key = None

    def packed_weight(self, dtype) -> Tensor:
        key = (self.weight._version, self.weight.data_ptr(), dtype)
        if self._packed is None or self._packed_key != key:
            self._packed = ops.pack_conv_weight(self.weight, dtype)
            self._packed_key = key
        return self._packed

    def bias_f32(self):
        return None if self.bias is None else self.bias.detach()

    def forward(self, x: Tensor, residual: Tensor | None = None, out_dtype=None, gn_next: bool = False,
                pre_norm: 'GroupNormSM100 | None' = None) -> Tensor:  # noqa: D102
        """gn_next: the output feeds a GroupNorm(32, eps 1e-6) next, so let the epilogue produce its statistics.
        pre_norm: compute conv(silu(pre_norm(x))); the normalisation runs inside the conv's mainloop when the shape
        allows it (no normalised tensor in HBM), otherwise as the separate GroupNorm-apply kernel."""
        x = ops.to_act(x, compute_dtype())
        if tape.grad_mode():  # training path: tape entries with hand-written backward kernels (eo_vae/autograd.py)
            if pre_norm is not None:
                x = pre_norm(x, silu=True)
            return tape.ConvFn.apply(x, self.weight, self.bias, residual, self, out_dtype, gn_next)
        in_gn = None
        if pre_norm is not None:
            if ops.USE_GN_PROLOGUE and ops.gn_prologue_ok(x, self.out_channels, self._mode, pre_norm.num_groups):
                fused = getattr(x, "_gn_stats", None)
                if fused is not None and fused[1] == pre_norm.num_groups and fused[2] == float(pre_norm.eps):
                    stats = fused[0]
                else:
                    stats = ops.gn_stats(x, pre_norm.num_groups, pre_norm.eps)
                in_gn = (stats, pre_norm.weight.detach(), pre_norm.bias.detach(), pre_norm.num_groups)
            else:
                x = pre_norm(x, silu=True)
        return ops.conv2d(x, self.packed_weight(x.dtype), self.bias_f32(), self.out_channels, self._mode,
                          residual=residual, out_dtype=out_dtype, gn_groups=32 if gn_next else 0, gn_eps=1e-6,
                          in_gn=in_gn)


class Downsample(nn.Module):
    """layers.py:25-37: zero pad right/bottom by one, 3x3 stride-2 conv.  The pad is never materialised: the four
    stride-2 sub-lattices are separate TMA maps and the missing row/column is TMA out-of-bounds zero fill."""

    def __init__(self, in_channels: int):
        super().__init__()
        self.conv = Conv2dSM100(in_channels, in_channels, kernel_size=3, stride=2, padding=0)

    def forward(self, x: Tensor) -> Tensor:
        return self.conv(x, gn_next=True)


class Upsample(nn.Module):
    """layers.py:40-50: nearest x2 then 3x3 conv - executed in sub-pixel form: four 2x2 convolutions on the low-resolution
    input with the 3x3 taps that land on the same source pixel summed (ops.conv2d_up2x), 16 instead of 36 MAC units and no
    materialised 4x tensor; the materialising path remains for shapes / the fp32 mode the phase kernel does not cover."""

    def __init__(self, in_channels: int):
        super().__init__()
        self.conv = Conv2dSM100(in_channels, in_channels, kernel_size=3, stride=1, padding=1)
        self._up = None
        self._up_key = None

    def packed_weight_up2x(self, dtype) -> Tensor:
        w = self.conv.weight
        key = (w._version, w.data_ptr(), dtype)
        if self._up is None or self._up_key != key:
            self._up, self._up_key = ops.pack_conv_weight_up2x(w, dtype), key
        return self._up

    def forward(self, x: Tensor) -> Tensor:
        x = ops.to_act(x, compute_dtype())
        if tape.grad_mode():
            return tape.UpsampleFn.apply(x, self.conv.weight, self.conv.bias, self)
        if ops.up2x_ok(x, self.conv.out_channels):
            return ops.conv2d_up2x(x, self.packed_weight_up2x(x.dtype), self.conv.bias_f32(), self.conv.out_channels, gn_groups=32)
        return self.conv(ops.upsample2x(x), gn_next=True)


class ResnetBlock(nn.Module):
    """layers.py:53-114: GN -> SiLU -> conv3x3 -> GN -> SiLU -> conv3x3 (+ 1x1 shortcut) + x."""

    def __init__(self, in_channels: int, out_channels: int, cond_dim: int = None):
        super().__init__()
        self.in_channels = in_channels
        out_channels = in_channels if out_channels is None else out_channels
        self.out_channels = out_channels
        self.cond_dim = cond_dim
        self.norm1 = GroupNormSM100(num_groups=32, num_channels=in_channels, eps=1e-6, affine=True)
        self.conv1 = Conv2dSM100(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
        if self.cond_dim is not None:  # AdaIN projection, initialised to identity (layers.py:68-76)
            self.emb_proj = nn.Linear(cond_dim, out_channels * 2)
            nn.init.zeros_(self.emb_proj.bias)
            self.emb_proj.weight.data.zero_()
            self.emb_proj.bias.data[:out_channels] = 1.0
        self.norm2 = GroupNormSM100(num_groups=32, num_channels=out_channels, eps=1e-6, affine=True)
        self.conv2 = Conv2dSM100(out_channels, out_channels, kernel_size=3, stride=1, padding=1)
        if self.in_channels != self.out_channels:
            self.nin_shortcut = Conv2dSM100(in_channels, out_channels, kernel_size=1, stride=1, padding=0)
        self._fused = None
        self._fused_key = None

    def _conv2_with_shortcut(self, dtype):
        """conv2 and nin_shortcut as ONE implicit GEMM: [W2 | Wnin] along K, b2 + bnin (derived cache)."""
        c2, sc = self.conv2, self.nin_shortcut
        key = (c2.weight._version, c2.weight.data_ptr(), sc.weight._version, sc.weight.data_ptr(), c2.bias._version,
               sc.bias._version, dtype)
        if self._fused is None or self._fused_key != key:
            w2, w1 = c2.packed_weight(dtype), sc.packed_weight(dtype)
            w = torch.cat([w2.reshape(w2.shape[0], -1), w1.reshape(w1.shape[0], -1)], dim=1).contiguous()
            b = (c2.bias.detach() + sc.bias.detach()).contiguous()
            self._fused, self._fused_key = (w, b), key
        return self._fused

    def _norm2_affine(self, emb: Tensor | None):
        """norm2's affine, modulated by the AdaIN style when given (layers.py:96-107).  ``emb`` is the conditioner's
        [B, cond_dim] output; its rows are identical (the style is a function of the wavelength vector), row 0 is used."""
        if self.cond_dim is None or emb is None:
            return self.norm2.weight, self.norm2.bias
        if emb.dim() != 2 or emb.shape[1] != self.cond_dim:
            raise RuntimeError(f"ResnetBlock: emb must be [B, {self.cond_dim}], got {tuple(emb.shape)}")
        style = emb[:1]
        if tape.grad_mode():
            return tape.AdaINAffineFn.apply(style, self.emb_proj.weight, self.emb_proj.bias, self.norm2.weight,
                                            self.norm2.bias)
        return ops.adain_affine_forward(style.detach().contiguous(), self.emb_proj.weight.detach(),
                                        self.emb_proj.bias.detach(), self.norm2.weight.detach(),
                                        self.norm2.bias.detach())[:2]

    def forward(self, x: Tensor, emb: Tensor | None = None) -> Tensor:
        x = ops.to_act(x, compute_dtype())
        g2, b2 = self._norm2_affine(emb)
        if tape.grad_mode():
            sc = getattr(self, 'nin_shortcut', None)
            return tape.ResnetBlockFn.apply(x, self.norm1.weight, self.norm1.bias, self.conv1.weight, self.conv1.bias,
                                            g2, b2, self.conv2.weight, self.conv2.bias,
                                            None if sc is None else sc.weight, None if sc is None else sc.bias, self)
        h = self.conv1(x, gn_next=True, pre_norm=self.norm1)  # GN1 + SiLU inside the conv where the shape allows
        norm2 = self.norm2 if g2 is self.norm2.weight else _ModulatedNorm(self.norm2, g2, b2)
        if self.in_channels == self.out_channels:
            # GN2 + SiLU in the prologue, residual add + next GN's statistics in the epilogue
            return self.conv2(h, residual=x, gn_next=True, pre_norm=norm2)
        h = norm2(h, silu=True)
        if self.in_channels % 64 == 0 and self.out_channels % 64 == 0 and x.dtype != torch.float32:
            # 1x1 shortcut folded into conv2's K loop: no shortcut tensor is written or re-read
            w, b = self._conv2_with_shortcut(x.dtype)
            return ops.conv2d(h, w, b, self.out_channels, ops.CONV_3X3, gn_groups=32, gn_eps=1e-6, x2=x)
        return self.conv2(h, residual=self.nin_shortcut(x), gn_next=True)


class AttnBlock(nn.Module):
    """layers.py:117-142: GN -> q,k,v 1x1 -> softmax(q k^T / sqrt(C)) v (single head, d = C) -> proj 1x1 -> + x.

    q, k and v come from ONE fused 1x1 implicit GEMM (Cout = 3C) and are consumed in place as channel slices; the
    reference's three ``rearrange(...).contiguous()`` copies do not exist here (NHWC is already [L, C])."""

    def __init__(self, in_channels: int):
        super().__init__()
        self.in_channels = in_channels
        self.norm = GroupNormSM100(num_groups=32, num_channels=in_channels, eps=1e-6, affine=True)
        self.q = Conv2dSM100(in_channels, in_channels, kernel_size=1)
        self.k = Conv2dSM100(in_channels, in_channels, kernel_size=1)
        self.v = Conv2dSM100(in_channels, in_channels, kernel_size=1)
        self.proj_out = Conv2dSM100(in_channels, in_channels, kernel_size=1)
        self._qkv = None
        self._qkv_key = None

    def _qkv_operands(self, dtype):
        key = tuple((m.weight._version, m.weight.data_ptr(), m.bias._version) for m in (self.q, self.k, self.v)) + (dtype,)
        if self._qkv is None or self._qkv_key != key:
            w = torch.cat([m.packed_weight(dtype) for m in (self.q, self.k, self.v)], dim=0)
            b = torch.cat([m.bias.detach() for m in (self.q, self.k, self.v)], dim=0).contiguous()
            self._qkv, self._qkv_key = (w, b), key
        return self._qkv

    def forward(self, x: Tensor) -> Tensor:
        x = ops.to_act(x, compute_dtype())
        n, c, hh, ww = x.shape
        if c % 16 != 0:
            raise RuntimeError("AttnBlock: channel count must be a multiple of 16 on the sm_100a path")
        L = hh * ww
        if tape.grad_mode():
            return tape.AttnBlockFn.apply(x, self.norm.weight, self.norm.bias, self.q.weight, self.q.bias, self.k.weight,
                                          self.k.bias, self.v.weight, self.v.bias, self.proj_out.weight,
                                          self.proj_out.bias, self)
        h = self.norm(x, silu=False)
        wqkv, bqkv = self._qkv_operands(x.dtype)
        qkv = ops.conv2d(h, wqkv, bqkv, 3 * c, ops.CONV_1X1)  # NHWC [n, L, 3c]
        flat = qkv.permute(0, 2, 3, 1).reshape(n, L, 3 * c)     # view: pixel-major rows, pitch 3c
        if x.dtype == torch.float32:
            # fp32 validation path: fp32 scores, exact-exp softmax, P V with V read in place (SIMT fp32 GEMMs)
            if L % 16 != 0:
                raise RuntimeError("AttnBlock (fp32 validation path): H*W must be a multiple of 16")
            q, k, v = flat[:, :, :c], flat[:, :, c:2 * c], flat[:, :, 2 * c:]
            scores = ops.gemm_tn_batched(q, k, torch.float32, scale=1.0 / math.sqrt(c))
            o = ops.gemm_pv_f32(ops.softmax_rows(scores, torch.float32), v).view(n, hh, ww, c).permute(0, 3, 1, 2)
            return self.proj_out(o, residual=x)
        if ops.attention_fused_ok(L, c, n):
            # flash-style: scores / probabilities live in TMEM and shared memory only
            o = ops.attention_fused(flat, c).view(n, hh, ww, c).permute(0, 3, 1, 2)
            return self.proj_out(o, residual=x, gn_next=True)
        q, k, v = flat[:, :, :c], flat[:, :, c:2 * c], flat[:, :, 2 * c:]
        lp = (L + 15) // 16 * 16  # key extent padded to the MMA K step (zero probabilities / zero values)
        scores = ops.gemm_tn_batched(q, k, torch.float32, scale=1.0 / math.sqrt(c))  # [n, L, L] fp32
        probs = ops.softmax_rows(scores, x.dtype, cols=L, out_cols=lp)               # [n, L, lp]
        vt = ops.transpose16(v, out_rows=lp)                      # [n, c, lp]
        o = ops.gemm_tn_batched(probs, vt, x.dtype)               # [n, L, c]
        o = o.view(n, hh, ww, c).permute(0, 3, 1, 2)
        return self.proj_out(o, residual=x, gn_next=True)
