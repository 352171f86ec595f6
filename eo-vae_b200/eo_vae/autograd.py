"""Tape entries of the training path: ``torch.autograd.Function`` objects whose forward AND backward run libeovae_sm100
kernels.  torch's autograd engine is only the tape (ordering, parameter ``.grad`` accumulation); no ATen compute kernel
sits on the gradient path of the encoder / decoder bodies.

One Function per reference module (layers.py: ResnetBlock :53-114, AttnBlock :117-142, Upsample :40-50; plain convs and
GroupNorm+SiLU for the model edges) so that the gradient fan-in of every residual branch is fused into a kernel epilogue
(``grad_add`` of the GroupNorm backward / the implicit-GEMM data gradient) instead of a separate elementwise add.

Gradient kernels used:
  data gradient     eovae_conv2d on eovae_pack_conv_weight_dgrad operands (flipped taps, swapped channels)
  weight gradient   eovae_conv2d_wgrad (tcgen05, pixels contracted, deterministic split-K)
  bias gradient     eovae_bias_grad
  GroupNorm(+SiLU)  eovae_gn_backward
  attention         eovae_gemm_tn_batched x4 + eovae_softmax_backward
  stride-2 / x2     eovae_scatter_stride2 / eovae_pool2x2_sum
"""
from __future__ import annotations

import math

import torch
from torch.autograd import Function

from . import ops
from .settings import compute_dtype, grad_dtype


def _gd(x: torch.Tensor):
    """Storage type of the gradients between layers = the type of the saved forward activation they meet in the
    weight-gradient GEMMs (tcgen05 kind::f16 needs equal A / B formats): settings.train_dtype on every taped call."""
    if x.dtype == torch.float32:
        raise RuntimeError("eo_vae: the fp32 validation path is forward-only (set_compute_dtype(torch.float16 | torch.bfloat16) "
                           "for training)")
    return x.dtype


def grad_mode() -> bool:
    """The modules take the tape-recording path whenever torch would record a graph (the fp32 validation path is
    forward-only: it never records)."""
    return torch.is_grad_enabled() and compute_dtype() != torch.float32


def _grad_act(g: torch.Tensor, dtype) -> torch.Tensor:
    """Incoming gradient -> 16-bit NHWC-stored tensor whose pixel pitch keeps TMA's 16-byte alignment."""
    n, c, h, w = g.shape
    if g.dtype == dtype:
        try:
            if ops.pix_stride(g) % 8 == 0:
                return g
        except RuntimeError:
            pass
    cp = (c + 15) // 16 * 16
    if g.dtype == torch.float32 and g.is_contiguous():
        return ops.nchw_to_act(g, cp, dtype)[:, :c]
    out = ops.nhwc_empty(n, cp, h, w, dtype, g.device)[:, :c]
    out.copy_(g)  # edge only (a gradient produced by torch glue on the latent)
    return out


def _grad_act_pad8(g: torch.Tensor, dtype) -> torch.Tensor:
    """Like _grad_act, but the channel count itself is padded to a multiple of 8 with ZERO lanes (the data-gradient
    GEMM contracts over them): used where the forward output had a ragged channel count (a 12-band reconstruction)."""
    n, c, h, w = g.shape
    if c % 8 == 0:
        return _grad_act(g, dtype)
    c8 = (c + 7) // 8 * 8
    a = _grad_act(g, dtype)
    ps = ops.pix_stride(a)
    if ps >= c8:
        wide = torch.as_strided(a, (n, c8, h, w), a.stride(), a.storage_offset())
        wide[:, c:].zero_()
        return wide
    out = torch.zeros((n, h, w, (c + 15) // 16 * 16), dtype=dtype, device=g.device).permute(0, 3, 1, 2)
    out[:, :c].copy_(a)
    return out[:, :c8]


def _dense(g: torch.Tensor) -> torch.Tensor:
    if ops.pix_stride(g) == g.shape[1]:
        return g
    return g.contiguous(memory_format=torch.channels_last)


def _wgrad(x: torch.Tensor, g: torch.Tensor, mode: int) -> torch.Tensor:
    if mode == ops.CONV_3X3_S2:
        if ops.s2_wgrad_ok(x, g):
            return ops.conv2d_s2_wgrad(x, g)   # x on its parity sub-lattices: 9 tap units over the OUTPUT pixels
        return ops.conv2d_wgrad(x, ops.scatter_stride2(_dense(g), x.shape[2], x.shape[3]), 3)
    return ops.conv2d_wgrad(x, g, 1 if mode == ops.CONV_1X1 else 3)


def _stats_of(x: torch.Tensor, groups: int = 32, eps: float = 1e-6) -> torch.Tensor:
    """GroupNorm statistics of x: the ones the producing conv's epilogue left on the tensor, else one stats pass."""
    fused = getattr(x, "_gn_stats", None)
    if fused is not None and fused[1] == groups and fused[2] == float(eps):
        return fused[0]
    return ops.gn_stats(x, groups, eps)


def _bias(mod):
    return None if mod.bias is None else mod.bias.detach()


class ConvFn(Function):
    """out = conv(x) + bias (+ residual) for the three built convolution modes."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, mod, out_dtype, gn_next=False):
        out = ops.conv2d(x, mod.packed_weight(x.dtype), _bias(mod), mod.out_channels, mod._mode, residual=residual,
                         out_dtype=out_dtype, gn_groups=32 if gn_next else 0)
        ctx.save_for_backward(x, weight)
        ctx.mode = mod._mode
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        g = _grad_act(dy, _gd(x))
        need = ctx.needs_input_grad
        dx = ops.conv2d_dgrad(_dense(g) if ctx.mode == ops.CONV_3X3_S2 else g, weight, ctx.mode, in_hw=x.shape[2:]) \
            if need[0] else None
        dw = _wgrad(x, g, ctx.mode) if need[1] else None
        if dw is not None and dw.shape[1] != weight.shape[1]:  # input channels were zero-padded for the kernels
            dw = dw[:, :weight.shape[1]].contiguous()
        db = ops.bias_grad(g) if ctx.has_bias and need[2] else None
        dres = g if need[3] else None
        return dx, dw, db, dres, None, None, None


class GroupNormFn(Function):
    """y = [silu](GroupNorm(x)) (model edges: norm_out)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, silu, groups, eps):
        stats = _stats_of(x, groups, eps)
        ctx.save_for_backward(x, stats, gamma, beta)
        ctx.cfg = (silu, groups)
        return ops.gn_apply(x, stats, gamma.detach(), beta.detach(), silu, groups)

    @staticmethod
    def backward(ctx, dy):
        x, stats, gamma, beta = ctx.saved_tensors
        silu, groups = ctx.cfg
        gx, dg, db = ops.gn_backward(x, _dense(_grad_act(dy, _gd(x))), stats, gamma, beta, silu, groups)
        return gx, dg, db, None, None, None


class ResnetBlockFn(Function):
    """x -> conv2(silu(gn2(conv1(silu(gn1(x)))))) + shortcut(x)   (layers.py:96-114)."""

    @staticmethod
    def forward(ctx, x, g1, b1, w1, c1b, g2, b2, w2, c2b, wn, nb, mod):
        dt = x.dtype
        st1 = _stats_of(x)
        a1 = ops.gn_apply(x, st1, g1.detach(), b1.detach(), True, 32)
        h = ops.conv2d(a1, mod.conv1.packed_weight(dt), _bias(mod.conv1), mod.out_channels, ops.CONV_3X3, gn_groups=32)
        fused = getattr(h, "_gn_stats", None)
        st2 = fused[0] if fused is not None else ops.gn_stats(h, 32, 1e-6)
        a2 = ops.gn_apply(h, st2, g2.detach(), b2.detach(), True, 32)
        if wn is None:
            out = ops.conv2d(a2, mod.conv2.packed_weight(dt), _bias(mod.conv2), mod.out_channels, ops.CONV_3X3, residual=x,
                             gn_groups=32)
        elif mod.in_channels % 64 == 0 and mod.out_channels % 64 == 0:
            w, b = mod._conv2_with_shortcut(dt)
            out = ops.conv2d(a2, w, b, mod.out_channels, ops.CONV_3X3, x2=x, gn_groups=32)
        else:
            sc = ops.conv2d(x, mod.nin_shortcut.packed_weight(dt), _bias(mod.nin_shortcut), mod.out_channels, ops.CONV_1X1)
            out = ops.conv2d(a2, mod.conv2.packed_weight(dt), _bias(mod.conv2), mod.out_channels, ops.CONV_3X3, residual=sc,
                             gn_groups=32)
        ctx.save_for_backward(x, st1, a1, h, st2, a2, g1, b1, w1, g2, b2, w2, wn)
        return out

    @staticmethod
    def backward(ctx, dy):
        x, st1, a1, h, st2, a2, g1, b1, w1, g2, b2, w2, wn = ctx.saved_tensors
        g = _dense(_grad_act(dy, _gd(x)))
        # conv2 (+ shortcut)
        da2 = ops.conv2d_dgrad(g, w2, ops.CONV_3X3)
        dw2 = ops.conv2d_wgrad(a2, g, 3)
        db2 = ops.bias_grad(g)
        if wn is None:
            gsc, dwn, dbn = g, None, None
        else:
            gsc = ops.conv2d_dgrad(g, wn, ops.CONV_1X1)
            dwn = ops.conv2d_wgrad(x, g, 1)
            dbn = db2
        dh, dg2, dbe2 = ops.gn_backward(h, da2, st2, g2, b2, True, 32)
        del da2
        # conv1
        da1 = ops.conv2d_dgrad(dh, w1, ops.CONV_3X3)
        dw1 = ops.conv2d_wgrad(a1, dh, 3)
        db1 = ops.bias_grad(dh)
        # gn1, with the shortcut gradient added in the same pass
        dx, dg1, dbe1 = ops.gn_backward(x, da1, st1, g1, b1, True, 32, grad_add=gsc)
        return dx, dg1, dbe1, dw1, db1, dg2, dbe2, dw2, db2, dwn, dbn, None


class AttnBlockFn(Function):
    """x + proj(softmax(q k^T / sqrt(c)) v) with q, k, v = 1x1 convs of GroupNorm(x)   (layers.py:128-142)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, wq, bq, wk, bk, wv, bv, wp, bp, mod):
        n, c, hh, ww = x.shape
        L = hh * ww
        lp = (L + 15) // 16 * 16
        stn = _stats_of(x)
        hn = ops.gn_apply(x, stn, gamma.detach(), beta.detach(), False, 32)
        wqkv, bqkv = mod._qkv_operands(x.dtype)
        qkv = ops.conv2d(hn, wqkv, bqkv, 3 * c, ops.CONV_1X1)
        flat = qkv.permute(0, 2, 3, 1).reshape(n, L, 3 * c)
        q, k, v = flat[:, :, :c], flat[:, :, c:2 * c], flat[:, :, 2 * c:]
        scores = ops.gemm_tn_batched(q, k, torch.float32, scale=1.0 / math.sqrt(c))
        probs = ops.softmax_rows(scores, x.dtype, cols=L, out_cols=lp)
        del scores
        o = ops.gemm_tn_batched(probs, ops.transpose16(v, out_rows=lp), x.dtype)
        o = o.view(n, hh, ww, c).permute(0, 3, 1, 2)
        out = ops.conv2d(o, mod.proj_out.packed_weight(x.dtype), _bias(mod.proj_out), c, ops.CONV_1X1, residual=x, gn_groups=32)
        ctx.save_for_backward(x, stn, hn, qkv, probs, o, gamma, beta, wq, wk, wv, wp)
        return out

    @staticmethod
    def backward(ctx, dy):
        x, stn, hn, qkv, probs, o, gamma, beta, wq, wk, wv, wp = ctx.saved_tensors
        n, c, hh, ww = x.shape
        L = hh * ww
        lp = probs.shape[-1]
        dt = _gd(x)   # gradients (bf16) meet forward activations (x.dtype) in mixed-format tcgen05 GEMMs below
        scale = 1.0 / math.sqrt(c)
        g = _dense(_grad_act(dy, dt))
        # proj_out
        do = ops.conv2d_dgrad(g, wp, ops.CONV_1X1)
        dwp = ops.conv2d_wgrad(o, g, 1)
        dbp = ops.bias_grad(g)
        dof = do.permute(0, 2, 3, 1).reshape(n, L, c)
        flat = qkv.permute(0, 2, 3, 1).reshape(n, L, 3 * c)
        q, k, v = flat[:, :, :c], flat[:, :, c:2 * c], flat[:, :, 2 * c:]
        dqkv = torch.empty((n, L, 3 * c), dtype=dt, device=x.device)
        same = lp == L
        # dV = P^T dO
        pt = ops.transpose16(probs, out_rows=lp)          # [n, lp, lp]  (rows >= L of P^T are zero)
        dot = ops.transpose16(dof, out_rows=lp)           # [n, c, lp]
        if same:
            ops.gemm_tn_batched(pt, dot, dt, out=dqkv[:, :, 2 * c:])
        else:
            dqkv[:, :, 2 * c:] = ops.gemm_tn_batched(pt, dot, dt)[:, :L]
        del pt, dot
        # dP = dO V^T ; dS = scale * P o (dP - rowsum(dP o P))
        dp = ops.gemm_tn_batched(dof, v, torch.float32)   # [n, L, L]
        ds = ops.softmax_backward(probs, dp, L, scale, out_dtype=dt)    # [n, L, lp]
        del dp
        # dQ = dS K ; dK = dS^T Q
        ops.gemm_tn_batched(ds, ops.transpose16(k, out_rows=lp), dt, out=dqkv[:, :, :c])
        dst = ops.transpose16(ds, out_rows=lp)            # [n, lp, lp]
        qt = ops.transpose16(q, out_rows=lp)              # [n, c, lp]
        if same:
            ops.gemm_tn_batched(dst, qt, dt, out=dqkv[:, :, c:2 * c])
        else:
            dqkv[:, :, c:2 * c] = ops.gemm_tn_batched(dst, qt, dt)[:, :L]
        del ds, dst, qt
        # fused q / k / v 1x1 convs
        dqkv4 = dqkv.view(n, hh, ww, 3 * c).permute(0, 3, 1, 2)
        wqkv = torch.cat([wq, wk, wv], dim=0)
        dhn = ops.conv2d_dgrad(dqkv4, wqkv, ops.CONV_1X1)
        dwqkv = ops.conv2d_wgrad(hn, dqkv4, 1)
        dbqkv = ops.bias_grad(dqkv4)
        dx, dgam, dbet = ops.gn_backward(x, dhn, stn, gamma, beta, False, 32, grad_add=g)
        dwq, dwk, dwv = dwqkv[:c], dwqkv[c:2 * c], dwqkv[2 * c:]
        dbq, dbk, dbv = dbqkv[:c], dbqkv[c:2 * c], dbqkv[2 * c:]
        return dx, dgam, dbet, dwq, dbq, dwk, dbk, dwv, dbv, dwp, dbp, None


class UpsampleFn(Function):
    """conv3x3(nearest_x2(x))   (layers.py:47-50); the upsampled tensor is rebuilt in backward, not kept."""

    @staticmethod
    def forward(ctx, x, weight, bias, mod):
        ctx.save_for_backward(x, weight)
        ctx.up2x = ops.up2x_ok(x, mod.conv.out_channels)
        if ctx.up2x:   # sub-pixel form: four 2x2 convs on the low-resolution input
            return ops.conv2d_up2x(x, mod.packed_weight_up2x(x.dtype), _bias(mod.conv), mod.conv.out_channels, gn_groups=32)
        return ops.conv2d(ops.upsample2x(x), mod.conv.packed_weight(x.dtype), _bias(mod.conv), mod.conv.out_channels,
                          ops.CONV_3X3, gn_groups=32)

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        g = _grad_act(dy, _gd(x))
        if ctx.up2x and ops.pix_stride(g) % 8 == 0:
            dx = ops.conv2d_up2x_dgrad(g, weight)           # one 16-tap launch over the parity sub-lattices of g
        else:
            dx = ops.pool2x2_sum(ops.conv2d_dgrad(g, weight, ops.CONV_3X3))
        if ctx.up2x and ops.up2x_wgrad_ok(x, g):
            dw = ops.conv2d_up2x_wgrad(x, g)                # four 2x2-tap launches, unfolded to the 3x3 gradient
        else:
            dw = ops.conv2d_wgrad(ops.upsample2x(x), g, 3)
        return dx, dw, ops.bias_grad(g), None


class ActToNchwFn(Function):
    """internal activation -> fp32 NCHW (model edge); the gradient comes back through eovae_nchw_to_nhwc16."""

    @staticmethod
    def forward(ctx, x):
        return ops.act_to_nchw_f32(x)

    @staticmethod
    def backward(ctx, g):
        n, c, h, w = g.shape
        return ops.nchw_to_act(g, (c + 15) // 16 * 16, grad_dtype())[:, :c]


def act_to_nchw_f32(x: torch.Tensor) -> torch.Tensor:
    if grad_mode() and x.requires_grad:
        return ActToNchwFn.apply(x)
    return ops.act_to_nchw_f32(x)


class DynConvInFn(Function):
    """Encoder input layer (dynamic_conv.py:511-527): kernel [E, C, 3, 3] and bias [E] generated from the wavelengths.
    Backward: weight / bias gradient of the conv, then the hypernetwork's own backward (eovae_hypernet_backward)."""

    @staticmethod
    def forward(ctx, x, mod, wvs, *hparams):
        c = wvs.size(0)
        wk, b_raw, tape = mod._generate_taped(wvs)
        packed, bias, _ = ops.pack_dyn_weight(wk, b_raw, c, mod.embed_dim, False, mod.scaler, mod.scaler, x.dtype, False)
        ctx.save_for_backward(x, wvs, tape)
        ctx.mod = mod
        return ops.conv2d(x, packed, bias, mod.embed_dim, ops.CONV_3X3, algo_cin=c, gn_groups=32)

    @staticmethod
    def backward(ctx, dy):
        x, wvs, tape = ctx.saved_tensors
        mod = ctx.mod
        g = _dense(_grad_act(dy, _gd(x)))
        dw = ops.conv2d_wgrad(x, g, 3)  # [E, C padded to 16, 3, 3]
        grads = mod._hyper_backward(wvs, dw, ops.bias_grad(g), mod.scaler, tape)
        return (None, None, None) + tuple(grads)


class DynConvOutFn(Function):
    """Decoder output layer (dynamic_conv.py:684-710): band kernels [C, E, 3, 3] and per-band bias generated from the
    wavelengths, then a 3x3 conv.  Backward: data gradient through the generated kernel, its weight / bias gradient,
    and the hypernetwork's backward."""

    @staticmethod
    def forward(ctx, x, mod, waves, *hparams):
        c = waves.size(0)
        wk, b_raw, tape = mod._generate_taped(waves)
        packed, bias, oihw = ops.pack_dyn_weight(wk, b_raw, c, mod.embed_dim, True, mod.scaler, mod.scaler * mod.scaler,
                                                 x.dtype, True)
        mod._last = (wk, b_raw, c)
        ctx.save_for_backward(x, oihw, waves, tape)
        ctx.mod = mod
        return ops.conv2d(x, packed, bias, c, ops.CONV_3X3, out_dtype=torch.float32)

    @staticmethod
    def backward(ctx, dy):
        x, oihw, waves, tape = ctx.saved_tensors
        mod = ctx.mod
        c = waves.size(0)
        g = _grad_act_pad8(dy, _gd(x))
        dx = ops.conv2d_dgrad(g, oihw, ops.CONV_3X3) if ctx.needs_input_grad[0] else None
        dw = ops.conv2d_wgrad(x, g[:, :c], 3)  # [C, E, 3, 3]
        grads = mod._hyper_backward(waves, dw, ops.bias_grad(g[:, :c]), mod.scaler * mod.scaler, tape)
        return (dx, None, None) + tuple(grads)


class SampleFn(Function):
    """z = mean + exp(0.5 * clamp(logvar, -30, 20)) * eps   (distributions.py:28-46) on the fused kernel."""

    @staticmethod
    def forward(ctx, moments, eps, zc):
        eps = eps.to(device=moments.device, dtype=torch.float32).contiguous()
        z, _ = ops.kl_reparam(moments, eps, zc, want_z=True)
        ctx.save_for_backward(moments, eps)
        ctx.zc = zc
        return z

    @staticmethod
    def backward(ctx, dz):
        moments, eps = ctx.saved_tensors
        return ops.reparam_backward(moments, eps, dz, ctx.zc), None, None


class PixelLossFn(Function):
    """mean |a - b| (kind 0) or mean sqrt((a - b)^2 + eps^2) (kind 1)   (consistency_loss.py:12-21,418)."""

    @staticmethod
    def forward(ctx, pred, target, eps, kind):
        a = pred.to(torch.float32).contiguous()
        b = target.to(torch.float32).contiguous()
        ctx.save_for_backward(a, b)
        ctx.cfg = (eps, kind)
        return ops.l1_charbonnier(a, b, eps)[kind]

    @staticmethod
    def backward(ctx, gl):
        a, b = ctx.saved_tensors
        eps, kind = ctx.cfg
        return ops.pixel_loss_backward(a, b, eps, kind, gl), None, None, None


class MsssimFn(Function):
    """batch-mean MS-SSIM (consistency_loss.py:24-37); backward = eovae_msssim_backward."""

    @staticmethod
    def forward(ctx, pred, target, data_range):
        a = pred.to(torch.float32).contiguous()
        b = target.to(torch.float32).contiguous()
        ctx.save_for_backward(a, b)
        ctx.data_range = data_range
        return ops.msssim(a, b, data_range)[0]

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        return ops.msssim_backward(a, b, ctx.data_range, g), None, None


class LatentTrainFn(Function):
    """TRAIN-mode latent glue in one kernel each way: pixel-unshuffle -> BatchNorm2d (batch statistics, running buffers
    updated in place) -> inverse normalisation with the updated buffers -> pixel-shuffle -> decoder input activation
    (new_autoencoder.py:466-469,533-543)."""

    @staticmethod
    def forward(ctx, z, bn, eps_inv, dtype):
        z = z.to(torch.float32).contiguous()
        n, zc, h, w = z.shape
        out = ops.nhwc_empty(n, zc, h, w, dtype, z.device)
        save = torch.empty((4 * zc, 3), dtype=torch.float32, device=z.device)
        momentum = 0.1 if bn.momentum is None else float(bn.momentum)
        _C = ops._C
        _C.check(_C.lib().eovae_latent_bn_train_forward(z.data_ptr(), n, zc, h, w, bn.running_mean.data_ptr(),
                                                        bn.running_var.data_ptr(), momentum, float(bn.eps), float(eps_inv),
                                                        out.data_ptr(), ops.DT[dtype], ops.pix_stride(out), save.data_ptr(),
                                                        ops._stream()), "eovae_latent_bn_train_forward")
        bn.num_batches_tracked += 1
        ctx.save_for_backward(z, save)
        return out

    @staticmethod
    def backward(ctx, dout):
        z, save = ctx.saved_tensors
        n, zc, h, w = z.shape
        g = _grad_act(dout, dout.dtype if dout.dtype != torch.float32 else torch.bfloat16)
        dz = torch.empty_like(z)
        _C = ops._C
        _C.check(_C.lib().eovae_latent_bn_train_backward(g.data_ptr(), ops.DT[g.dtype], ops.pix_stride(g), z.data_ptr(), n, zc, h, w,
                                                         save.data_ptr(), dz.data_ptr(), ops._stream()),
                 "eovae_latent_bn_train_backward")
        return dz, None, None, None


class WavelengthStyleFn(Function):
    """WavelengthConditioner (model.py:35-64): wavelengths -> style row [1, d] on eovae_wavelength_style_forward/backward."""

    @staticmethod
    def forward(ctx, wvs, omega, *mlp):
        params = [omega] + [p.detach() for p in mlp]
        style, tape_ws = ops.wavelength_style_forward(wvs, params, omega.numel() * 2)
        ctx.save_for_backward(tape_ws, omega, *mlp)
        return style

    @staticmethod
    def backward(ctx, dstyle):
        tape_ws, omega, *mlp = ctx.saved_tensors
        grads = ops.wavelength_style_backward([omega] + [p.detach() for p in mlp], omega.numel() * 2, dstyle, tape_ws)
        return (None, None) + tuple(grads)


class AdaINAffineFn(Function):
    """(style, emb_proj, norm2 affine) -> modulated GroupNorm affine gamma*scale, beta*scale + shift (layers.py:96-104)."""

    @staticmethod
    def forward(ctx, style, wproj, bproj, gamma, beta):
        g_out, b_out, style2 = ops.adain_affine_forward(style.detach().contiguous(), wproj.detach(), bproj.detach(),
                                                        gamma.detach(), beta.detach())
        ctx.save_for_backward(style, wproj, gamma, beta, style2)
        return g_out, b_out

    @staticmethod
    def backward(ctx, dg_out, db_out):
        style, wproj, gamma, beta, style2 = ctx.saved_tensors
        dgamma, dbeta, dw, dbp, dstyle = ops.adain_affine_backward(style.detach().contiguous(), wproj.detach(), gamma.detach(),
                                                                   beta.detach(), style2, dg_out, db_out)
        return dstyle.view_as(style), dw, dbp, dgamma, dbeta


class SamLossFn(Function):
    """mean(1 - cos(pred, target)) along the band axis (SAMLoss, consistency_loss.py:186-210)."""

    @staticmethod
    def forward(ctx, pred, target, eps):
        a = pred.to(torch.float32).contiguous()
        b = target.to(torch.float32).contiguous()
        ctx.save_for_backward(a, b)
        ctx.eps = eps
        return ops.sam_loss(a, b, eps)

    @staticmethod
    def backward(ctx, gl):
        a, b = ctx.saved_tensors
        return ops.sam_loss_backward(a, b, ctx.eps, gl), None, None


class GradDiffLossFn(Function):
    """GradientDifferenceLoss with alpha = 1 (consistency_loss.py:241-269)."""

    @staticmethod
    def forward(ctx, pred, target):
        a = pred.to(torch.float32).contiguous()
        b = target.to(torch.float32).contiguous()
        ctx.save_for_backward(a, b)
        return ops.grad_diff_loss(a, b)

    @staticmethod
    def backward(ctx, gl):
        a, b = ctx.saved_tensors
        return ops.grad_diff_loss_backward(a, b, gl), None


class DistillWeightFn(Function):
    """Differentiable ``get_distillation_weight`` (dynamic_conv.py:471-497, 638-664): wavelengths -> generated OIHW kernel
    and bias, both scaled by 0.1, for the stage-1 weight distillation loop (weight_distill_train.py:190-264: MSE against
    the Flux conv_in / conv_out weights).  Backward = the hypernetwork's backward kernels on (dW, dbias)."""

    @staticmethod
    def forward(ctx, mod, wvs, *hparams):
        c = wvs.numel()
        wk, b_raw, tape_ws = mod._generate_taped(wvs)
        _, bias, oihw = ops.pack_dyn_weight(wk, b_raw, c, mod.embed_dim, mod._decoder, mod.scaler, mod.scaler,
                                            torch.bfloat16, True)
        ctx.save_for_backward(wvs, tape_ws)
        ctx.mod = mod
        return oihw, bias

    @staticmethod
    def backward(ctx, dw, db):
        wvs, tape_ws = ctx.saved_tensors
        mod = ctx.mod
        if dw is None:
            dw = torch.zeros((wvs.numel(), mod.embed_dim, 3, 3) if mod._decoder else (mod.embed_dim, wvs.numel(), 3, 3),
                             dtype=torch.float32, device=wvs.device)
        if db is None:
            db = torch.zeros((wvs.numel() if mod._decoder else mod.embed_dim,), dtype=torch.float32, device=wvs.device)
        grads = mod._hyper_backward(wvs, dw.contiguous(), db.contiguous(), mod.scaler, tape_ws)
        return (None, None) + tuple(grads)


class FocalFreqLossFn(Function):
    """FocalFrequencyLoss (ffl.py:17-104) with the detached batch-normalised log weight matrix."""

    @staticmethod
    def forward(ctx, pred, target, patch_factor, alpha):
        out, ws = ops.focal_freq_loss(pred, target, patch_factor, alpha, keep=True)
        ctx.save_for_backward(ws)
        ctx.cfg = (tuple(pred.shape), patch_factor)
        return out

    @staticmethod
    def backward(ctx, gl):
        (ws,) = ctx.saved_tensors
        shape, pf = ctx.cfg
        return ops.focal_freq_loss_backward(shape, pf, gl, ws), None, None, None


class LatentResizeRotFn(Function):
    """EQ-VAE latent transform: bilinear rescale then rot90 (new_autoencoder.py:460-464, 519-531) on one gather kernel;
    backward = its adjoint (scatter of the four taps)."""

    @staticmethod
    def forward(ctx, z, size, rot_k):
        ctx.cfg = (tuple(z.shape), size, rot_k)
        return ops.latent_resize_rot(z, size, rot_k)

    @staticmethod
    def backward(ctx, g):
        shape, size, rot_k = ctx.cfg
        return ops.latent_resize_rot_backward(g, shape, size, rot_k), None, None
