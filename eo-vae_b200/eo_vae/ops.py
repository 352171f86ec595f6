"""Torch-tensor front end of the C ABI (include/eovae.h): allocation, stream and argument marshalling only.

Activations are ordinary ``torch.Tensor`` objects of logical shape [N, C, H, W] stored channels-last (NHWC in
memory) in a 16-bit dtype, so every module boundary stays a plain tensor while the kernels see pixel-major data.
Nothing here computes: each function forwards to one hand-written sm_100a kernel and raises on failure.
"""
from __future__ import annotations

import ctypes

import torch

from . import _C
from ._C import BF16, CONV_1X1, CONV_3X3, CONV_3X3_S2, F16, F32  # noqa: F401

DT = {torch.bfloat16: BF16, torch.float16: F16, torch.float32: F32}


# bench.py sets this to a list to collect (family, algorithmic flops, start event, end event) per tensor-core launch
PROFILE = None


def _timed(family: str, flops: float, call):
    if PROFILE is None:
        return call()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    rc = call()
    end.record()
    PROFILE.append((family, flops, start, end))
    return rc


def set_tuning(key: int, value: int) -> None:
    """Launch-shape knobs of the library (include/eovae.h EOVAE_TUNE_*); results never depend on them."""
    _C.lib().eovae_set_tuning(int(key), int(value))


TUNE_GN_APPLY_CORESIDENT = 1
TUNE_GN_BWD_BLOCK_ELEMS = 3
TUNE_GN_BWD_BULK = 4
TUNE_GN_APPLY_BLOCK_ELEMS = 5


def launch_count() -> int:
    """Kernels launched so far by libeovae_sm100.so in this process."""
    return int(_C.lib().eovae_launch_count())


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream() -> int:
    """cudaStream_t of torch's current stream on the current device (raw handle: ~10x cheaper than building a
    torch.cuda.Stream object for each of the ~550 launches of a training step)."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


def _need_cuda(*ts) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("eo_vae: tensors must live on a CUDA device - this path has no CPU implementation")


def nhwc_empty(n: int, c: int, h: int, w: int, dtype, device) -> torch.Tensor:
    """Uninitialised logical-NCHW tensor with NHWC storage."""
    return torch.empty((n, h, w, c), dtype=dtype, device=device).permute(0, 3, 1, 2)


def pix_stride(t: torch.Tensor) -> int:
    """Elements between consecutive pixels of an NHWC-stored [N, C, H, W] view (channel slices allowed)."""
    n, c, h, w = t.shape
    ps = t.stride(3)
    ok = (c == 1 or t.stride(1) == 1) and (h == 1 or t.stride(2) == w * ps) and (n == 1 or t.stride(0) == h * w * ps)
    if not ok or ps < c:
        raise RuntimeError(f"eo_vae: expected channels-last storage, got shape {tuple(t.shape)} strides {t.stride()}")
    return ps


def to_act(x: torch.Tensor, dtype) -> torch.Tensor:
    """Bring an arbitrary [N, C, H, W] tensor to the internal activation form (edge use only)."""
    _need_cuda(x)
    if x.dtype == dtype and x.dim() == 4:
        try:
            pix_stride(x)
            return x
        except RuntimeError:
            pass
    return x.to(dtype=dtype, memory_format=torch.channels_last)


# ------------------------------------------------------------------------------------------------ conv / gemm
def conv_k_per_tap(cin: int) -> int:
    return _C.lib().eovae_conv_k_per_tap(cin)


def pack_conv_weight(w: torch.Tensor, dtype) -> torch.Tensor:
    """OIHW fp32 -> K-major 16-bit [round_up(cout,16)][taps][k_per_tap] (derived cache, never persisted)."""
    _need_cuda(w)
    cout, cin, kh, kw = w.shape
    wf = w.detach().to(torch.float32).contiguous()
    rows = (cout + 15) // 16 * 16
    out = torch.empty((rows, kh * kw, conv_k_per_tap(cin)), dtype=dtype, device=w.device)
    _C.check(_C.lib().eovae_pack_conv_weight(_ptr(wf), _ptr(out), cout, cin, kh, kw, DT[dtype], _stream()),
             "eovae_pack_conv_weight")
    return out


def conv2d(x: torch.Tensor, w_packed: torch.Tensor, bias, cout: int, mode: int, residual=None, out_dtype=None,
           scale: float = 1.0, algo_cin: int | None = None, gn_groups: int = 0, gn_eps: float = 1e-6,
           x2: torch.Tensor | None = None, in_gn=None) -> torch.Tensor:
    """out = scale * conv(x) + bias + residual.  With gn_groups > 0 the epilogue also produces the GroupNorm statistics
    of the output; they ride along as ``out._gn_stats = (stats, groups, eps)`` for the next GroupNormSM100.  ``x2``: a second
    activation whose 1x1 convolution is accumulated in the same mainloop (weights appended along K in ``w_packed``).
    ``in_gn = (stats, gamma, beta, groups)``: x is a RAW tensor and GroupNorm + SiLU is applied to the operand tiles
    inside the mainloop (only where ``gn_prologue_ok`` says so)."""
    _need_cuda(x, w_packed, bias, residual)
    n, cin, h, w = x.shape
    ho, wo = (h, w) if mode != CONV_3X3_S2 else ((h - 2) // 2 + 1, (w - 2) // 2 + 1)
    out_dtype = out_dtype or x.dtype
    # pixel pitch padded to 16 channels so every row stays 16-byte aligned (e.g. a 12-band reconstruction)
    out = nhwc_empty(n, (cout + 15) // 16 * 16, ho, wo, out_dtype, x.device)[:, :cout]
    res_dt, res_ps = 0, 0
    if residual is not None:
        if tuple(residual.shape) != (n, cout, ho, wo):
            raise RuntimeError("eo_vae.conv2d: residual shape mismatch")
        res_dt, res_ps = DT[residual.dtype], pix_stride(residual)
    if bias is not None and (bias.dtype != torch.float32 or bias.numel() != cout):
        raise RuntimeError("eo_vae.conv2d: bias must be fp32 [cout]")
    stats = ws = None
    ws_bytes = 0
    if gn_groups > 0 and x.dtype != torch.float32:   # (fp32 validation path: statistics come from eovae_gn_stats)
        ws_bytes = _C.lib().eovae_conv2d_gn_workspace_bytes(n, h, w, mode, cout, gn_groups)
        if ws_bytes > 0:
            stats = torch.empty((n, gn_groups, 2), dtype=torch.float32, device=x.device)
            ws = torch.empty((ws_bytes // 4,), dtype=torch.float32, device=x.device)
    flops = 2.0 * n * ho * wo * cout * (algo_cin or cin) * (1 if mode == CONV_1X1 else 9)
    cin2, x2_ps = 0, 0
    if x2 is not None:
        _need_cuda(x2)
        if x2.dtype != x.dtype or tuple(x2.shape[0:1] + x2.shape[2:]) != (n, h, w):
            raise RuntimeError("eo_vae.conv2d: fused 1x1 operand must match the main input's batch / size / dtype")
        cin2, x2_ps = x2.shape[1], pix_stride(x2)
        flops += 2.0 * n * ho * wo * cout * cin2
    g_stats = g_gamma = g_beta = g_ws = None
    g_groups, g_ws_bytes = 0, 0
    if in_gn is not None:
        g_stats, g_gamma, g_beta, g_groups = in_gn
        g_ws = torch.empty((n * cin * 2,), dtype=torch.float32, device=x.device)
        g_ws_bytes = g_ws.numel() * 4
    rc = _timed("conv", flops, lambda: _C.lib().eovae_conv2d(
        _ptr(x), n, h, w, cin, pix_stride(x), mode, _ptr(w_packed), cout, _ptr(bias), _ptr(residual), res_dt, res_ps,
        _ptr(out), DT[out_dtype], pix_stride(out), DT[x.dtype], float(scale), _ptr(stats), gn_groups, float(gn_eps),
        _ptr(ws), ws_bytes, _ptr(x2), cin2, x2_ps, _ptr(g_stats), _ptr(g_gamma), _ptr(g_beta), g_groups, _ptr(g_ws),
        g_ws_bytes, _stream()))
    _C.check(rc, "eovae_conv2d")
    if stats is not None:
        out._gn_stats = (stats, gn_groups, float(gn_eps))
    return out


# ------------------------------------------------------------------------------------------------ sub-pixel upsample conv
# Upsample (layers.py:40-50) = nearest x2 + conv3x3.  USE_UP2X: run it as four 2x2 convolutions on the low-resolution input
# (16 instead of 36 MAC units, no materialised 4x tensor) in forward, data gradient and weight gradient.
USE_UP2X = True


def up2x_ok(x: torch.Tensor, cout: int) -> bool:
    n, cin, h, w = x.shape
    return (USE_UP2X and x.dtype != torch.float32 and pix_stride(x) % 8 == 0
            and bool(_C.lib().eovae_conv2d_up2x_ok(n, h, w, cin, cout)))


def pack_conv_weight_up2x(w: torch.Tensor, dtype, dgrad: bool = False) -> torch.Tensor:
    """OIHW fp32 3x3 -> folded 2x2 operands: forward [4 phases][rows][4 taps][k] or data-gradient [cin rows][16][k]."""
    _need_cuda(w)
    cout, cin, kh, kw = w.shape
    if (kh, kw) != (3, 3):
        raise RuntimeError("eo_vae.pack_conv_weight_up2x: 3x3 kernels only")
    wf = w.detach().to(torch.float32).contiguous()
    if dgrad:
        out = torch.empty(((cin + 15) // 16 * 16, 16, conv_k_per_tap((cout + 7) // 8 * 8)), dtype=dtype, device=w.device)
    else:
        out = torch.empty((4, (cout + 15) // 16 * 16, 4, conv_k_per_tap(cin)), dtype=dtype, device=w.device)
    _C.check(_C.lib().eovae_pack_conv_weight_up2x(_ptr(wf), _ptr(out), cout, cin, DT[dtype], 1 if dgrad else 0, _stream()),
             "eovae_pack_conv_weight_up2x")
    return out


def conv2d_up2x(x: torch.Tensor, w_packed: torch.Tensor, bias, cout: int, gn_groups: int = 0, gn_eps: float = 1e-6) -> torch.Tensor:
    """conv3x3(nearest_x2(x)) + bias from the LOW-resolution x; optional GroupNorm statistics of the output."""
    _need_cuda(x, w_packed, bias)
    n, cin, h, w = x.shape
    out = nhwc_empty(n, cout, 2 * h, 2 * w, x.dtype, x.device)
    stats = ws = None
    ws_bytes = 0
    lib = _C.lib()
    if gn_groups > 0:
        ws_bytes = lib.eovae_conv2d_up2x_gn_workspace_bytes(n, h, w, cout, gn_groups)
        if ws_bytes > 0:
            stats = torch.empty((n, gn_groups, 2), dtype=torch.float32, device=x.device)
            ws = torch.empty((ws_bytes // 4,), dtype=torch.float32, device=x.device)
    flops = 2.0 * n * (2 * h) * (2 * w) * cout * cin * 4   # MACs actually executed: 4 taps per output pixel
    rc = _timed("conv", flops, lambda: lib.eovae_conv2d_up2x(
        _ptr(x), n, h, w, cin, pix_stride(x), _ptr(w_packed), cout, _ptr(bias), _ptr(out), DT[out.dtype], pix_stride(out),
        DT[x.dtype], _ptr(stats), gn_groups if stats is not None else 0, float(gn_eps), _ptr(ws), ws_bytes, _stream()))
    _C.check(rc, "eovae_conv2d_up2x")
    if stats is not None:
        out._gn_stats = (stats, gn_groups, float(gn_eps))
    return out


def conv2d_up2x_dgrad(dy: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """Data gradient of conv2d_up2x wrt its low-resolution input (w: OIHW fp32 master weight): one 16-tap launch."""
    _need_cuda(dy, w)
    n, cout, h2, w2 = dy.shape
    cin = w.shape[1]
    wp = pack_conv_weight_up2x(w, dy.dtype, dgrad=True)
    dx = nhwc_empty(n, cin, h2 // 2, w2 // 2, dy.dtype, dy.device)
    flops = 2.0 * n * (h2 // 2) * (w2 // 2) * cin * cout * 16
    rc = _timed("conv", flops, lambda: _C.lib().eovae_conv2d_up2x_dgrad(
        _ptr(dy), n, h2, w2, cout, pix_stride(dy), _ptr(wp), cin, _ptr(dx), DT[dx.dtype], pix_stride(dx), DT[dy.dtype], _stream()))
    _C.check(rc, "eovae_conv2d_up2x_dgrad")
    return dx


def conv2d_up2x_wgrad(x: torch.Tensor, dy: torch.Tensor) -> torch.Tensor:
    """Weight gradient [cout, cin, 3, 3] fp32 of conv2d_up2x (x: low-resolution input, dy: high-resolution output gradient)."""
    _need_cuda(x, dy)
    n, cin, h, w = x.shape
    cout = dy.shape[1]
    lib = _C.lib()
    ws_bytes = lib.eovae_conv2d_up2x_wgrad_workspace_bytes(n, h, w, cin, cout)
    ws = torch.empty((ws_bytes // 4 + 1,), dtype=torch.float32, device=x.device)
    dw = torch.empty((cout, cin, 3, 3), dtype=torch.float32, device=x.device)
    flops = 2.0 * n * h * w * cout * cin * 16
    rc = _timed("wgrad", flops, lambda: lib.eovae_conv2d_up2x_wgrad(
        _ptr(x), pix_stride(x), _ptr(dy), pix_stride(dy), DT[x.dtype], n, h, w, cin, cout, _ptr(dw), 0, _ptr(ws), ws_bytes, _stream()))
    _C.check(rc, "eovae_conv2d_up2x_wgrad")
    return dw


def conv2d_s2_wgrad(x: torch.Tensor, dy: torch.Tensor) -> torch.Tensor:
    """Weight gradient [cout, cin, 3, 3] fp32 of the Downsample conv: x [n, cin, 2ho, 2wo] (read on its parity sub-lattices),
    dy [n, cout, ho, wo] - no zero-interleaved gradient image."""
    _need_cuda(x, dy)
    n, cin, h, w = x.shape
    cout, ho, wo = dy.shape[1], dy.shape[2], dy.shape[3]
    lib = _C.lib()
    ws_bytes = lib.eovae_conv2d_s2_wgrad_workspace_bytes(n, ho, wo, cin, cout)
    ws = torch.empty((ws_bytes // 4 + 1,), dtype=torch.float32, device=x.device)
    dw = torch.empty((cout, cin, 3, 3), dtype=torch.float32, device=x.device)
    flops = 2.0 * n * ho * wo * cout * cin * 9
    rc = _timed("wgrad", flops, lambda: lib.eovae_conv2d_s2_wgrad(
        _ptr(x), pix_stride(x), _ptr(dy), pix_stride(dy), DT[x.dtype], n, ho, wo, cin, cout, _ptr(dw), 0, _ptr(ws), ws_bytes, _stream()))
    _C.check(rc, "eovae_conv2d_s2_wgrad")
    return dw


def s2_wgrad_ok(x: torch.Tensor, dy: torch.Tensor) -> bool:
    n, cin, h, w = x.shape
    ho, wo = dy.shape[2], dy.shape[3]
    return (USE_UP2X and x.dtype == dy.dtype and (h, w) == (2 * ho, 2 * wo) and cin % 4 == 0 and pix_stride(x) % 8 == 0
            and pix_stride(dy) % 8 == 0 and bool(_C.lib().eovae_conv2d_wgrad_nhwc_ok(ho, wo)))


def up2x_wgrad_ok(x: torch.Tensor, dy: torch.Tensor) -> bool:
    n, cin, h, w = x.shape
    return (USE_UP2X and x.dtype == dy.dtype and cin % 4 == 0 and pix_stride(x) % 8 == 0 and pix_stride(dy) % 8 == 0
            and bool(_C.lib().eovae_conv2d_wgrad_nhwc_ok(h, w)))


# The in-mainloop GroupNorm prologue is correct but slower than gn_apply + conv on B200 (see igemm_sm100.cuh): opt-in.
USE_GN_PROLOGUE = False


def gn_prologue_ok(x: torch.Tensor, cout: int, mode: int, groups: int = 32) -> bool:
    """Can eovae_conv2d apply GroupNorm + SiLU to this input inside its mainloop?"""
    n, cin, h, w = x.shape
    return pix_stride(x) == cin and bool(_C.lib().eovae_conv2d_gn_prologue_ok(n, h, w, cin, cout, mode, groups))


def gemm_tn_batched(a: torch.Tensor, b: torch.Tensor, out_dtype, scale: float = 1.0, out: torch.Tensor | None = None) -> torch.Tensor:
    """c[i] = scale * a[i] @ b[i].T ; a [B, M, K], b [B, N, K]; row pitches may exceed K (channel slices).  ``out``: an
    existing [B, M, N] view (e.g. a channel slice of a wider buffer) whose batch stride equals M row pitches."""
    _need_cuda(a, b)
    bsz, m, k = a.shape
    n = b.shape[1]
    # a / b may mix f16 and bf16 (kind::f16 takes the two operand formats independently); fp32 = the validation path
    if (b.shape[0] != bsz or b.shape[2] != k or a.stride(2) != 1 or b.stride(2) != 1 or a.element_size() != b.element_size()):
        raise RuntimeError("eo_vae.gemm_tn_batched: bad operand layout")
    if a.stride(0) != m * a.stride(1):
        raise RuntimeError("eo_vae.gemm_tn_batched: A batches must be contiguous")
    if out is None:
        c = torch.empty((bsz, m, n), dtype=out_dtype, device=a.device)
    else:
        c = out
        if tuple(c.shape) != (bsz, m, n) or c.dtype != out_dtype or c.stride(2) != 1 or c.stride(0) != m * c.stride(1):
            raise RuntimeError("eo_vae.gemm_tn_batched: bad output view")
    rc = _timed("attn_gemm", 2.0 * bsz * m * n * k, lambda: _C.lib().eovae_gemm_tn_batched(
        _ptr(a), a.stride(1), a.stride(0), _ptr(b), b.stride(1), b.stride(0), _ptr(c), DT[out_dtype], c.stride(1), bsz, m, n, k,
        DT[a.dtype], DT[b.dtype], float(scale), _stream()))
    _C.check(rc, "eovae_gemm_tn_batched")
    return c


# Fused flash-style attention (eovae_attention_fused).  Measured against the q k^T GEMM -> softmax -> p v
# GEMM path (tools/attn_bench.py, C 512): 0.124 vs 0.132 ms at 16 x 1024, 0.466 vs 0.437 ms at 64 x 1024, 3.30 vs 1.92 ms at
# 32 x 4096 - on par at the 256-pixel shapes (and 384 MiB of scores never written), slower for very long sequences, where it
# is used only when the score tensor would be unreasonable to materialise.
USE_FUSED_ATTENTION = True
FUSED_ATTENTION_MAX_L = 2048            # above this the GEMM path is faster ...
FUSED_ATTENTION_SCORE_BYTES = 8 << 30   # ... unless N * L^2 * 6 bytes of scores + probabilities exceed this


def gemm_pv_f32(p: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """fp32 validation path: o[i] = p[i] @ v[i] with p [B, L, Lk] dense and v [B, Lk, C] read IN PLACE (a channel slice of the
    qkv tensor: keys along the row pitch), Lk a multiple of 16."""
    _need_cuda(p, v)
    bsz, l, lk = p.shape
    c = v.shape[2]
    if p.dtype != torch.float32 or v.dtype != torch.float32 or not p.is_contiguous() or v.stride(2) != 1 or v.shape[1] != lk:
        raise RuntimeError("eo_vae.gemm_pv_f32: bad operand layout")
    out = torch.empty((bsz, l, c), dtype=torch.float32, device=p.device)
    rc = _C.lib().eovae_gemm_strided_f32(_ptr(p), lk, l * lk, _ptr(v), 1, v.stride(1), v.stride(0), _ptr(out), c, bsz, l, c, lk,
                                        1.0, _stream())
    _C.check(rc, "eovae_gemm_strided_f32")
    return out


def attention_fused_ok(l: int, c: int, n: int = 1) -> bool:
    if not USE_FUSED_ATTENTION or not _C.lib().eovae_attention_fused_ok(l, c):
        return False
    return l <= FUSED_ATTENTION_MAX_L or 6 * n * l * l > FUSED_ATTENTION_SCORE_BYTES


def attention_fused(qkv: torch.Tensor, c: int) -> torch.Tensor:
    """qkv: [N, L, 3c] 16-bit rows (pitch may exceed 3c) -> softmax(q k^T / sqrt(c)) v as [N, L, c]; one fused kernel."""
    _need_cuda(qkv)
    n, l, _ = qkv.shape
    if qkv.stride(2) != 1 or qkv.stride(0) != l * qkv.stride(1):
        raise RuntimeError("eo_vae.attention_fused: bad qkv layout")
    out = torch.empty((n, l, c), dtype=qkv.dtype, device=qkv.device)
    flops = 4.0 * n * l * l * c
    rc = _timed("attn_fused", flops, lambda: _C.lib().eovae_attention_fused(_ptr(qkv), qkv.stride(1), n, l, c, _ptr(out), c,
                                                                            DT[qkv.dtype], _stream()))
    _C.check(rc, "eovae_attention_fused")
    return out


# ------------------------------------------------------------------------------------------------ group norm
def gn_stats(x: torch.Tensor, groups: int = 32, eps: float = 1e-6) -> torch.Tensor:
    _need_cuda(x)
    n, c, h, w = x.shape
    stats = torch.empty((n, groups, 2), dtype=torch.float32, device=x.device)
    ws_bytes = _C.lib().eovae_gn_stats_workspace_bytes(n, h * w, c, groups)
    ws = torch.empty((max(ws_bytes // 8, 1),), dtype=torch.float64, device=x.device)
    rc = _C.lib().eovae_gn_stats(_ptr(x), DT[x.dtype], n, h * w, c, pix_stride(x), groups, float(eps), _ptr(stats),
                                _ptr(ws), ws.numel() * 8, _stream())
    _C.check(rc, "eovae_gn_stats")
    return stats


def gn_apply(x: torch.Tensor, stats: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, silu: bool,
             groups: int = 32, out_dtype=None) -> torch.Tensor:
    _need_cuda(x, stats, gamma, beta)
    n, c, h, w = x.shape
    y = nhwc_empty(n, c, h, w, out_dtype or x.dtype, x.device)
    rc = _C.lib().eovae_gn_apply(_ptr(x), DT[x.dtype], pix_stride(x), _ptr(stats), _ptr(gamma), _ptr(beta), _ptr(y),
                                DT[y.dtype], pix_stride(y), n, h * w, c, groups, 1 if silu else 0, _stream())
    _C.check(rc, "eovae_gn_apply")
    return y


def group_norm(x, gamma, beta, silu: bool, groups: int = 32, eps: float = 1e-6):
    """GroupNorm (+SiLU).  Statistics come from the producing conv's epilogue when it left them on the tensor."""
    fused = getattr(x, "_gn_stats", None)
    if fused is not None and fused[1] == groups and fused[2] == float(eps):
        stats = fused[0]
    else:
        stats = gn_stats(x, groups, eps)
    return gn_apply(x, stats, gamma, beta, silu, groups)


# ------------------------------------------------------------------------------------------------ edges
def nchw_to_act(x: torch.Tensor, c_pad: int, dtype) -> torch.Tensor:
    """fp32 NCHW image batch -> internal activation with channels zero-padded to c_pad."""
    _need_cuda(x)
    x = x.to(torch.float32).contiguous()
    n, c, h, w = x.shape
    out = nhwc_empty(n, c_pad, h, w, dtype, x.device)
    _C.check(_C.lib().eovae_nchw_to_nhwc16(_ptr(x), _ptr(out), n, c, h, w, c_pad, DT[dtype], _stream()),
             "eovae_nchw_to_nhwc16")
    return out


def act_to_nchw_f32(x: torch.Tensor, channels: int | None = None) -> torch.Tensor:
    _need_cuda(x)
    n, c, h, w = x.shape
    c = channels or c
    out = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
    _C.check(_C.lib().eovae_nhwc_to_nchw_f32(_ptr(x), DT[x.dtype], pix_stride(x), _ptr(out), n, c, h, w, _stream()),
             "eovae_nhwc_to_nchw_f32")
    return out


def upsample2x(x: torch.Tensor) -> torch.Tensor:
    _need_cuda(x)
    n, c, h, w = x.shape
    if pix_stride(x) != c:
        raise RuntimeError("eo_vae.upsample2x: dense channels-last input required")
    out = nhwc_empty(n, c, 2 * h, 2 * w, x.dtype, x.device)
    # the kernel copies 16-byte vectors of 2-byte lanes: an fp32 pixel is twice as many lanes
    _C.check(_C.lib().eovae_upsample2x(_ptr(x), _ptr(out), n, h, w, c * x.element_size() // 2, _stream()), "eovae_upsample2x")
    return out


def softmax_rows(s: torch.Tensor, out_dtype, cols: int | None = None, out_cols: int | None = None) -> torch.Tensor:
    """softmax over the first ``cols`` entries of every row of s [..., ld]; output [..., out_cols] with zeros beyond
    ``cols`` (K padding for the following GEMM)."""
    _need_cuda(s)
    if not s.is_contiguous():
        raise RuntimeError("eo_vae.softmax_rows: contiguous input required")
    ld = s.shape[-1]
    cols = cols or ld
    out_cols = out_cols or cols
    rows = s.numel() // ld
    p = torch.empty(s.shape[:-1] + (out_cols,), dtype=out_dtype, device=s.device)
    _C.check(_C.lib().eovae_softmax_rows(_ptr(s), DT[s.dtype], ld, _ptr(p), DT[out_dtype], out_cols, rows, cols,
                                         _stream()), "eovae_softmax_rows")
    return p


def transpose16(x: torch.Tensor, out_rows: int | None = None) -> torch.Tensor:
    """[B, R, C] (row pitch may exceed C) 16-bit -> dense [B, C, out_rows >= R], zero padded."""
    _need_cuda(x)
    b, r, c = x.shape
    out_rows = out_rows or r
    if x.stride(2) != 1 or x.stride(0) != r * x.stride(1) or x.element_size() != 2:
        raise RuntimeError("eo_vae.transpose16: bad layout")
    out = torch.empty((b, c, out_rows), dtype=x.dtype, device=x.device)
    _C.check(_C.lib().eovae_transpose16(_ptr(x), x.stride(1), _ptr(out), out_rows, b, r, c, _stream()),
             "eovae_transpose16")
    return out


# ------------------------------------------------------------------------------------------------ latent glue
def _strides4(t: torch.Tensor):
    return (ctypes.c_longlong * 4)(*t.stride())


def latent_norm(moments: torch.Tensor, running_mean, running_var, eps: float, zc: int, out: torch.Tensor | None = None) -> torch.Tensor:
    """moments: fp32 logical [N, 2zc, H, W] (any strides) -> normalised spatial latent NCHW fp32 [N, zc, H, W] (``out``: a
    contiguous destination, e.g. a batch slice of a larger latent tensor)."""
    _need_cuda(moments, running_mean, running_var)
    n, c2, h, w = moments.shape
    if moments.dtype != torch.float32 or c2 != 2 * zc:
        raise RuntimeError("eo_vae.latent_norm: bad moments tensor")
    z = out if out is not None else torch.empty((n, zc, h, w), dtype=torch.float32, device=moments.device)
    if tuple(z.shape) != (n, zc, h, w) or z.dtype != torch.float32 or not z.is_contiguous():
        raise RuntimeError("eo_vae.latent_norm: bad output tensor")
    rc = _C.lib().eovae_latent_norm(_ptr(moments), _strides4(moments), _ptr(running_mean), _ptr(running_var), float(eps),
                                   _ptr(z), n, h, w, zc, _stream())
    _C.check(rc, "eovae_latent_norm")
    return z


def latent_denorm(z: torch.Tensor, running_mean, running_var, eps: float, dtype) -> torch.Tensor:
    """normalised spatial latent NCHW fp32 -> inverse BN -> decoder input activation (NHWC 16-bit)."""
    _need_cuda(z, running_mean, running_var)
    z = z.to(torch.float32).contiguous()
    n, zc, h, w = z.shape
    out = nhwc_empty(n, zc, h, w, dtype, z.device)
    rc = _C.lib().eovae_latent_denorm(_ptr(z), _ptr(running_mean), _ptr(running_var), float(eps), _ptr(out), DT[dtype], n,
                                     h, w, zc, _stream())
    _C.check(rc, "eovae_latent_denorm")
    return out


def kl_reparam(moments: torch.Tensor, eps, zc: int, want_z: bool = True):
    """moments fp32 logical [N, 2zc, H, W] (any strides) -> (z NCHW fp32 or None, kl [N])."""
    _need_cuda(moments, eps)
    n, c2, h, w = moments.shape
    if moments.dtype != torch.float32 or c2 != 2 * zc:
        raise RuntimeError("eo_vae.kl_reparam: bad moments tensor")
    z = torch.empty((n, zc, h, w), dtype=torch.float32, device=moments.device) if want_z else None
    kl = torch.empty((n,), dtype=torch.float32, device=moments.device)
    if eps is not None:
        eps = eps.to(device=moments.device, dtype=torch.float32).contiguous()
        if eps.numel() != n * zc * h * w:
            raise RuntimeError(f"eo_vae.kl_reparam: noise has {eps.numel()} elements, the latent {n * zc * h * w}")
    rc = _C.lib().eovae_kl_reparam(_ptr(moments), _strides4(moments), _ptr(eps), _ptr(z), _ptr(kl), n, h, w, zc, _stream())
    _C.check(rc, "eovae_kl_reparam")
    return z, kl


def l1_charbonnier(a: torch.Tensor, b: torch.Tensor, eps: float = 1e-3):
    """-> fp32 tensor [2] = (mean |a-b|, mean sqrt((a-b)^2 + eps^2))."""
    _need_cuda(a, b)
    a = a.to(torch.float32).contiguous()
    b = b.to(torch.float32).contiguous()
    out = torch.empty((2,), dtype=torch.float32, device=a.device)
    ws = torch.empty((2,), dtype=torch.float64, device=a.device)
    _C.check(_C.lib().eovae_l1_charbonnier(_ptr(a), _ptr(b), a.numel(), float(eps), _ptr(out), _ptr(ws), 16, _stream()),
             "eovae_l1_charbonnier")
    return out


def _nchw_f32_pair(pred: torch.Tensor, target: torch.Tensor):
    _need_cuda(pred, target)
    if pred.dim() != 4 or pred.shape != target.shape:
        raise RuntimeError(f"loss inputs must be two [B, C, H, W] tensors of one shape, got {tuple(pred.shape)} / {tuple(target.shape)}")
    return pred.to(torch.float32).contiguous(), target.to(torch.float32).contiguous()


def sam_loss(pred: torch.Tensor, target: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """SAMLoss (consistency_loss.py:186-210) -> fp32 scalar tensor."""
    a, b = _nchw_f32_pair(pred, target)
    n, c, h, w = a.shape
    out = torch.empty((1,), dtype=torch.float32, device=a.device)
    ws = torch.empty((2,), dtype=torch.float64, device=a.device)
    _C.check(_C.lib().eovae_sam_loss(_ptr(a), _ptr(b), n, c, h * w, float(eps), _ptr(out), _ptr(ws), 16, _stream()), "eovae_sam_loss")
    return out[0]


def sam_loss_backward(a: torch.Tensor, b: torch.Tensor, eps: float, grad_scale: torch.Tensor) -> torch.Tensor:
    _need_cuda(a, b, grad_scale)
    n, c, h, w = a.shape
    ga = torch.empty_like(a)
    gs = grad_scale.to(torch.float32).reshape(1)
    _C.check(_C.lib().eovae_sam_loss_backward(_ptr(a), _ptr(b), n, c, h * w, float(eps), _ptr(gs), _ptr(ga), _stream()),
             "eovae_sam_loss_backward")
    return ga


def grad_diff_loss(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """GradientDifferenceLoss, alpha = 1 (consistency_loss.py:241-269) -> fp32 scalar tensor."""
    a, b = _nchw_f32_pair(pred, target)
    n, c, h, w = a.shape
    out = torch.empty((1,), dtype=torch.float32, device=a.device)
    ws = torch.empty((2,), dtype=torch.float64, device=a.device)
    _C.check(_C.lib().eovae_grad_diff_loss(_ptr(a), _ptr(b), n * c, h, w, _ptr(out), _ptr(ws), 16, _stream()), "eovae_grad_diff_loss")
    return out[0]


def grad_diff_loss_backward(a: torch.Tensor, b: torch.Tensor, grad_scale: torch.Tensor) -> torch.Tensor:
    _need_cuda(a, b, grad_scale)
    n, c, h, w = a.shape
    ga = torch.empty_like(a)
    gs = grad_scale.to(torch.float32).reshape(1)
    _C.check(_C.lib().eovae_grad_diff_loss_backward(_ptr(a), _ptr(b), n * c, h, w, _ptr(gs), _ptr(ga), _stream()),
             "eovae_grad_diff_loss_backward")
    return ga


def focal_freq_loss(pred: torch.Tensor, target: torch.Tensor, patch_factor: int = 1, alpha: float = 1.0, keep: bool = False):
    """FocalFrequencyLoss (ffl.py:17-104; batch_matrix, log_matrix) -> (fp32 scalar tensor, workspace | None).  ``keep``:
    return the workspace (weight * spectrum) for focal_freq_loss_backward."""
    a, b = _nchw_f32_pair(pred, target)
    n, c, h, w = a.shape
    lib = _C.lib()
    ws_bytes = lib.eovae_focal_freq_loss_workspace_bytes(n, c, h, w, patch_factor)
    ws = torch.empty((ws_bytes // 4 + 1,), dtype=torch.float32, device=a.device)
    out = torch.empty((1,), dtype=torch.float32, device=a.device)
    _C.check(lib.eovae_focal_freq_loss(_ptr(a), _ptr(b), n, c, h, w, int(patch_factor), float(alpha), 1 if keep else 0, _ptr(out),
                                       _ptr(ws), ws_bytes, _stream()), "eovae_focal_freq_loss")
    return out[0], (ws if keep else None)


def focal_freq_loss_backward(shape, patch_factor: int, grad_scale: torch.Tensor, ws: torch.Tensor) -> torch.Tensor:
    _need_cuda(grad_scale, ws)
    n, c, h, w = shape
    ga = torch.empty(shape, dtype=torch.float32, device=ws.device)
    gs = grad_scale.to(torch.float32).reshape(1)
    _C.check(_C.lib().eovae_focal_freq_loss_backward(n, c, h, w, int(patch_factor), _ptr(gs), _ptr(ga), _ptr(ws),
                                                     (ws.numel() - 1) * 4, _stream()), "eovae_focal_freq_loss_backward")
    return ga


def _rot_shape(nh: int, nw: int, k: int):
    return (nw, nh) if k & 1 else (nh, nw)


def latent_resize_rot(z: torch.Tensor, size, rot_k: int) -> torch.Tensor:
    """rot90(bilinear resize(z, size), k, dims=[-1, -2]) (new_autoencoder.py:460-464, 519-531); size None = rotation only."""
    _need_cuda(z)
    z = z.to(torch.float32).contiguous()
    n, c, h, w = z.shape
    nh, nw = (h, w) if size is None else size
    oh, ow = _rot_shape(nh, nw, rot_k)
    out = torch.empty((n, c, oh, ow), dtype=torch.float32, device=z.device)
    _C.check(_C.lib().eovae_latent_resize_rot(_ptr(z), n * c, h, w, nh, nw, int(rot_k) & 3, _ptr(out), _stream()),
             "eovae_latent_resize_rot")
    return out


def latent_resize_rot_backward(grad_out: torch.Tensor, in_shape, size, rot_k: int) -> torch.Tensor:
    _need_cuda(grad_out)
    g = grad_out.to(torch.float32).contiguous()
    n, c, h, w = in_shape
    nh, nw = (h, w) if size is None else size
    gz = torch.empty(in_shape, dtype=torch.float32, device=g.device)
    _C.check(_C.lib().eovae_latent_resize_rot_backward(_ptr(g), n * c, h, w, nh, nw, int(rot_k) & 3, _ptr(gz), _stream()),
             "eovae_latent_resize_rot_backward")
    return gz


def area_resize_rot(x: torch.Tensor, size, rot_k: int) -> torch.Tensor:
    """rot90(F.interpolate(x, size, mode='area'), k, dims=[-1, -2]): the EQ-VAE reconstruction target (:611-636)."""
    _need_cuda(x)
    x = x.to(torch.float32).contiguous()
    n, c, h, w = x.shape
    nh, nw = size
    oh, ow = _rot_shape(nh, nw, rot_k)
    out = torch.empty((n, c, oh, ow), dtype=torch.float32, device=x.device)
    _C.check(_C.lib().eovae_area_resize_rot(_ptr(x), n * c, h, w, nh, nw, int(rot_k) & 3, _ptr(out), _stream()),
             "eovae_area_resize_rot")
    return out


def msssim(pred: torch.Tensor, target: torch.Tensor, data_range: float = 6.0):
    """-> (mean MS-SSIM over the batch [1], per-sample MS-SSIM [B]); fp32 NCHW inputs."""
    _need_cuda(pred, target)
    pred = pred.to(torch.float32).contiguous()
    target = target.to(torch.float32).contiguous()
    b, c, h, w = pred.shape
    lib = _C.lib()
    ws_bytes = lib.eovae_msssim_workspace_bytes(b, c, h, w)
    ws = torch.empty((ws_bytes // 4 + 1,), dtype=torch.float32, device=pred.device)
    out = torch.empty((1,), dtype=torch.float32, device=pred.device)
    per = torch.empty((b,), dtype=torch.float32, device=pred.device)
    _C.check(lib.eovae_msssim(_ptr(pred), _ptr(target), b, c, h, w, float(data_range), _ptr(out), _ptr(per), _ptr(ws),
                              ws_bytes, _stream()), "eovae_msssim")
    return out, per


def msssim_backward(pred: torch.Tensor, target: torch.Tensor, data_range: float, grad_scale: torch.Tensor) -> torch.Tensor:
    """grad_scale * d(mean MS-SSIM)/d(pred), fp32 NCHW."""
    _need_cuda(pred, target, grad_scale)
    pred = pred.to(torch.float32).contiguous()
    target = target.to(torch.float32).contiguous()
    b, c, h, w = pred.shape
    lib = _C.lib()
    ws_bytes = lib.eovae_msssim_backward_workspace_bytes(b, c, h, w)
    ws = torch.empty((ws_bytes // 4 + 1,), dtype=torch.float32, device=pred.device)
    g = torch.empty_like(pred)
    gs = grad_scale.to(torch.float32).reshape(1)
    _C.check(lib.eovae_msssim_backward(_ptr(pred), _ptr(target), b, c, h, w, float(data_range), _ptr(gs), _ptr(g), _ptr(ws),
                                       ws_bytes, _stream()), "eovae_msssim_backward")
    return g


# ------------------------------------------------------------------------------------------------ backward pieces
def pack_conv_weight_dgrad(w: torch.Tensor, dtype) -> torch.Tensor:
    """OIHW fp32 -> operand of the data-gradient conv (flipped taps, in/out channels swapped)."""
    _need_cuda(w)
    cout, cin, kh, kw = w.shape
    wf = w.detach().to(torch.float32).contiguous()
    out = torch.empty(((cin + 15) // 16 * 16, kh * kw, conv_k_per_tap((cout + 7) // 8 * 8)), dtype=dtype, device=w.device)
    _C.check(_C.lib().eovae_pack_conv_weight_dgrad(_ptr(wf), _ptr(out), cout, cin, kh, kw, DT[dtype], _stream()),
             "eovae_pack_conv_weight_dgrad")
    return out


def scatter_stride2(dy: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """dy [n, c, ho, wo] -> z [n, c, h, w], zero except z[2i+1][2j+1] = dy[i][j] (adjoint of the Downsample gather)."""
    _need_cuda(dy)
    n, c, ho, wo = dy.shape
    if pix_stride(dy) != c:
        raise RuntimeError("eo_vae.scatter_stride2: dense channels-last gradient required")
    z = nhwc_empty(n, c, h, w, dy.dtype, dy.device)
    _C.check(_C.lib().eovae_scatter_stride2(_ptr(dy), _ptr(z), n, ho, wo, h, w, c, _stream()), "eovae_scatter_stride2")
    return z


def conv2d_s2_dgrad(dy: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """Data gradient of the Downsample conv in sub-pixel form: four 2x2 convolutions over dy, one per input parity, each
    stored on its parity sub-lattice of dx [n, cin, 2 ho, 2 wo] - no zero-interleaved gradient image, 16 / 36 of the MACs."""
    _need_cuda(dy, w)
    n, cout, ho, wo = dy.shape
    cin = w.shape[1]
    wf = w.detach().to(torch.float32).contiguous()
    wp = torch.empty((4, (cin + 15) // 16 * 16, 4, conv_k_per_tap((cout + 7) // 8 * 8)), dtype=dy.dtype, device=dy.device)
    _C.check(_C.lib().eovae_pack_conv_weight_up2x(_ptr(wf), _ptr(wp), cout, cin, DT[dy.dtype], 2, _stream()),
             "eovae_pack_conv_weight_up2x")
    dx = nhwc_empty(n, cin, 2 * ho, 2 * wo, dy.dtype, dy.device)
    flops = 2.0 * n * ho * wo * cin * cout * 16
    rc = _timed("conv", flops, lambda: _C.lib().eovae_conv2d_s2_dgrad(
        _ptr(dy), n, ho, wo, cout, pix_stride(dy), _ptr(wp), cin, _ptr(dx), DT[dx.dtype], pix_stride(dx), DT[dy.dtype], _stream()))
    _C.check(rc, "eovae_conv2d_s2_dgrad")
    return dx


def s2_dgrad_ok(dy: torch.Tensor, cin: int, in_hw) -> bool:
    n, cout, ho, wo = dy.shape
    return (USE_UP2X and in_hw is not None and tuple(in_hw) == (2 * ho, 2 * wo) and cout % 64 == 0 and pix_stride(dy) % 8 == 0
            and bool(_C.lib().eovae_conv2d_up2x_ok(n, ho, wo, cout, cin)))


def conv2d_dgrad(dy: torch.Tensor, w: torch.Tensor, mode: int, in_hw=None, grad_add=None) -> torch.Tensor:
    """Data gradient of eovae_conv2d (w: OIHW fp32 master weight) as another implicit GEMM; ``grad_add`` (same shape
    as the result) is accumulated in the epilogue (gradient fan-in of a residual branch)."""
    cout, cin = w.shape[0], w.shape[1]
    if mode == CONV_3X3_S2 and grad_add is None and s2_dgrad_ok(dy, cin, in_hw):
        return conv2d_s2_dgrad(dy, w)
    wp = pack_conv_weight_dgrad(w, dy.dtype)
    if mode == CONV_3X3_S2:
        return conv2d(scatter_stride2(dy, in_hw[0], in_hw[1]), wp, None, cin, CONV_3X3, residual=grad_add)
    return conv2d(dy, wp, None, cin, mode, residual=grad_add)


USE_WGRAD_NHWC = True  # False: always take the transposed-copy kernel (kept for image sizes that do not tile into 64-pixel boxes)


def _pixel_rows(t: torch.Tensor) -> torch.Tensor:
    """[N, C, H, W] NHWC-stored (pixel pitch may exceed C) -> [N, H*W, C] view with the pitch as row stride."""
    n, c, h, w = t.shape
    ps = pix_stride(t)
    return torch.as_strided(t, (n, h * w, c), (h * w * ps, ps, 1), t.storage_offset())


def conv2d_wgrad(x: torch.Tensor, dy: torch.Tensor, ksize: int, dw: torch.Tensor | None = None) -> torch.Tensor:
    """Weight gradient [cout, cin, k, k] fp32 of a stride-1 conv (x: its NHWC 16-bit input, dy: output gradient)."""
    _need_cuda(x, dy)
    n, cin, h, w = x.shape
    cout = dy.shape[1]
    lib = _C.lib()
    if (USE_WGRAD_NHWC and cin % 4 == 0 and pix_stride(x) % 8 == 0 and pix_stride(dy) % 8 == 0
            and lib.eovae_conv2d_wgrad_nhwc_ok(h, w)):
        # operands read in place (MN-major UMMA operands): no transposed copies
        ws_bytes = lib.eovae_conv2d_wgrad_nhwc_workspace_bytes(n, h, w, cin, cout, ksize)
        ws = torch.empty((ws_bytes // 4 + 1,), dtype=torch.float32, device=x.device)
        acc = dw is not None
        if dw is None:
            dw = torch.empty((cout, cin, ksize, ksize), dtype=torch.float32, device=x.device)
        flops = 2.0 * n * h * w * cout * cin * ksize * ksize
        rc = _timed("wgrad", flops, lambda: lib.eovae_conv2d_wgrad_nhwc(
            _ptr(x), pix_stride(x), _ptr(dy), pix_stride(dy), DT[x.dtype], DT[dy.dtype], n, h, w, cin, cout, ksize, _ptr(dw), 1 if acc else 0,
            _ptr(ws), ws_bytes, _stream()))
        _C.check(rc, "eovae_conv2d_wgrad_nhwc")
        return dw
    if cin % 16 != 0:  # narrow edge layers (e.g. an 8-channel latent): zero-pad the channels, slice the result
        xp = torch.zeros((n, h, w, (cin + 15) // 16 * 16), dtype=x.dtype, device=x.device).permute(0, 3, 1, 2)
        xp[:, :cin].copy_(x)
        fresh = conv2d_wgrad(xp, dy, ksize)[:, :cin].contiguous()
        return fresh if dw is None else dw.add_(fresh)
    # transposed-operand kernel: channel-major copies, image rows padded to a multiple of 8 pixels
    wp = (w + 7) // 8 * 8
    xt = torch.empty((3 if ksize == 3 else 1, n, cin, h * wp), dtype=x.dtype, device=x.device)
    _C.check(lib.eovae_transpose16_xshift(_ptr(x), pix_stride(x), _ptr(xt), n, h, w, wp, cin, -1 if ksize == 3 else 0,
                                          3 if ksize == 3 else 1, _stream()), "eovae_transpose16_xshift")
    dyt = torch.empty((n, cout, h * wp), dtype=dy.dtype, device=dy.device)
    _C.check(lib.eovae_transpose16_xshift(_ptr(dy), pix_stride(dy), _ptr(dyt), n, h, w, wp, cout, 0, 1, _stream()),
             "eovae_transpose16_xshift")
    ws_bytes = lib.eovae_conv2d_wgrad_workspace_bytes(n, h, wp, cin, cout, ksize)
    ws = torch.empty((ws_bytes // 4 + 1,), dtype=torch.float32, device=x.device)
    acc = dw is not None
    if dw is None:
        dw = torch.empty((cout, cin, ksize, ksize), dtype=torch.float32, device=x.device)
    _C.check(lib.eovae_conv2d_wgrad(_ptr(xt), _ptr(dyt), DT[x.dtype], DT[dy.dtype], n, h, wp, cin, cout, ksize, _ptr(dw), 1 if acc else 0,
                                    _ptr(ws), ws_bytes, _stream()), "eovae_conv2d_wgrad")
    return dw


def bias_grad(dy: torch.Tensor) -> torch.Tensor:
    _need_cuda(dy)
    n, c, h, w = dy.shape
    fused = getattr(dy, "_colsum", None)  # left by eovae_gn_backward when dy is its grad_x output
    if fused is not None and fused.numel() == c:
        return fused
    cp = pix_stride(dy)  # a padded pixel pitch is summed as extra columns and dropped
    lib = _C.lib()
    ws_bytes = lib.eovae_bias_grad_workspace_bytes(n * h * w, cp)
    ws = torch.empty((ws_bytes // 4 + 1,), dtype=torch.float32, device=dy.device)
    out = torch.empty((cp,), dtype=torch.float32, device=dy.device)
    _C.check(lib.eovae_bias_grad(_ptr(dy), DT[dy.dtype], n * h * w, cp, _ptr(out), 0, _ptr(ws), ws_bytes, _stream()),
             "eovae_bias_grad")
    return out[:c]


def gn_backward(x: torch.Tensor, grad_out: torch.Tensor, stats, gamma, beta, silu: bool, groups: int = 32,
                grad_add=None):
    """-> (grad_x, dgamma, dbeta) of y = [silu](GroupNorm(x)).  x is the forward activation (its dtype), grad_out / grad_add
    / grad_x share the gradient dtype (the training default mixes f16 activations with bf16 gradients)."""
    _need_cuda(x, grad_out, stats, gamma, beta, grad_add)
    n, c, h, w = x.shape
    if pix_stride(x) != c or pix_stride(grad_out) != c or (grad_add is not None and pix_stride(grad_add) != c):
        raise RuntimeError("eo_vae.gn_backward: dense channels-last tensors required")
    if grad_add is not None and grad_add.dtype != grad_out.dtype:
        raise RuntimeError("eo_vae.gn_backward: grad_add must have the gradient dtype")
    lib = _C.lib()
    ws_bytes = lib.eovae_gn_backward_workspace_bytes(n, h * w, c, groups)
    ws = torch.empty((ws_bytes // 4 + 1,), dtype=torch.float32, device=x.device)
    gx = nhwc_empty(n, c, h, w, grad_out.dtype, x.device)
    dg = torch.empty((c,), dtype=torch.float32, device=x.device)
    db = torch.empty((c,), dtype=torch.float32, device=x.device)
    cs = torch.empty((c,), dtype=torch.float32, device=x.device)
    _C.check(lib.eovae_gn_backward(_ptr(x), _ptr(grad_out), DT[x.dtype], DT[grad_out.dtype], _ptr(stats), _ptr(gamma), _ptr(beta), n, h * w, c,
                                   groups, 1 if silu else 0, _ptr(grad_add), _ptr(gx), _ptr(dg), _ptr(db), 0, _ptr(cs), _ptr(ws),
                                   ws_bytes, _stream()), "eovae_gn_backward")
    gx._colsum = cs  # per-channel sum of gx (bias gradient of the conv that produced x): rides along like _gn_stats
    return gx, dg, db


def softmax_backward(p: torch.Tensor, dp: torch.Tensor, cols: int, scale: float, out_dtype=None) -> torch.Tensor:
    """ds = scale * p o (dp - rowsum(dp o p)); p 16-bit [..., lp], dp fp32 [..., >= cols] -> ds 16-bit [..., lp] (out_dtype)."""
    _need_cuda(p, dp)
    if not p.is_contiguous() or not dp.is_contiguous() or dp.dtype != torch.float32:
        raise RuntimeError("eo_vae.softmax_backward: contiguous p (16-bit) and dp (fp32) required")
    lp = p.shape[-1]
    rows = p.numel() // lp
    ds = torch.empty_like(p, dtype=out_dtype or p.dtype)
    _C.check(_C.lib().eovae_softmax_backward(_ptr(p), lp, _ptr(dp), dp.shape[-1], _ptr(ds), lp, DT[p.dtype], DT[ds.dtype], rows, cols, lp,
                                             float(scale), _stream()), "eovae_softmax_backward")
    return ds


def reparam_backward(moments: torch.Tensor, eps: torch.Tensor, dz: torch.Tensor, zc: int) -> torch.Tensor:
    _need_cuda(moments, eps, dz)
    n, c2, h, w = moments.shape
    dz = dz.to(torch.float32).contiguous()
    dm = torch.empty((n, c2, h, w), dtype=torch.float32, device=moments.device)
    _C.check(_C.lib().eovae_reparam_backward(_ptr(moments), _strides4(moments), _ptr(eps), _ptr(dz), _ptr(dm), n, h, w, zc,
                                             _stream()), "eovae_reparam_backward")
    return dm


def pixel_loss_backward(a: torch.Tensor, b: torch.Tensor, eps: float, kind: int, grad_scale: torch.Tensor) -> torch.Tensor:
    """gradient wrt a (fp32, a's shape) of mean|a-b| (kind 0) / Charbonnier (kind 1) times the device scalar grad_scale."""
    _need_cuda(a, b, grad_scale)
    ga = torch.empty_like(a)
    gs = grad_scale.to(torch.float32).reshape(1)
    _C.check(_C.lib().eovae_pixel_loss_backward(_ptr(a), _ptr(b), a.numel(), float(eps), kind, _ptr(gs), _ptr(ga), _stream()),
             "eovae_pixel_loss_backward")
    return ga


def pool2x2_sum(g: torch.Tensor) -> torch.Tensor:
    _need_cuda(g)
    n, c, h2, w2 = g.shape
    out = nhwc_empty(n, c, h2 // 2, w2 // 2, g.dtype, g.device)
    _C.check(_C.lib().eovae_pool2x2_sum(_ptr(g), _ptr(out), DT[g.dtype], n, h2 // 2, w2 // 2, c, _stream()),
             "eovae_pool2x2_sum")
    return out


# ------------------------------------------------------------------------------------------------ hypernetwork
def hypernet_forward(wvs: torch.Tensor, params: list, num_layers: int, d: int, heads: int, ff: int, embed: int,
                     decoder: bool):
    """-> (wk [C, 9*embed], bias_raw [embed] | [C]) fp32, unscaled (see eovae_hypernet_forward)."""
    _need_cuda(wvs, *params)
    wvs = wvs.to(torch.float32).contiguous()
    c = wvs.numel()
    lib = _C.lib()
    ws_bytes = lib.eovae_hypernet_workspace_bytes(c, d, ff, embed)
    ws = torch.empty((ws_bytes // 4 + 1,), dtype=torch.float32, device=wvs.device)
    wk = torch.empty((c, 9 * embed), dtype=torch.float32, device=wvs.device)
    bias = torch.empty((c if decoder else embed,), dtype=torch.float32, device=wvs.device)
    arr = (ctypes.c_void_p * len(params))(*[p.data_ptr() for p in params])
    rc = lib.eovae_hypernet_forward(_ptr(wvs), c, arr, num_layers, d, heads, ff, embed, 1 if decoder else 0, _ptr(wk),
                                    _ptr(bias), _ptr(ws), ws_bytes, _stream())
    _C.check(rc, "eovae_hypernet_forward")
    return wk, bias


def hypernet_forward_taped(wvs: torch.Tensor, params: list, num_layers: int, d: int, heads: int, ff: int, embed: int,
                           decoder: bool):
    """hypernet_forward that also returns the activation tape (a workspace tensor) for hypernet_backward(tape=...)."""
    _need_cuda(wvs, *params)
    wvs = wvs.to(torch.float32).contiguous()
    c = wvs.numel()
    lib = _C.lib()
    ws_bytes = lib.eovae_hypernet_backward_workspace_bytes(c, d, ff, embed, num_layers)
    tape = torch.empty((ws_bytes // 4 + 1,), dtype=torch.float32, device=wvs.device)
    wk = torch.empty((c, 9 * embed), dtype=torch.float32, device=wvs.device)
    bias = torch.empty((c if decoder else embed,), dtype=torch.float32, device=wvs.device)
    arr = (ctypes.c_void_p * len(params))(*[p.data_ptr() for p in params])
    rc = lib.eovae_hypernet_forward_taped(_ptr(wvs), c, arr, num_layers, d, heads, ff, embed, 1 if decoder else 0, _ptr(wk),
                                          _ptr(bias), _ptr(tape), ws_bytes, _stream())
    _C.check(rc, "eovae_hypernet_forward_taped")
    return wk, bias, tape


def hypernet_backward(wvs: torch.Tensor, params: list, num_layers: int, d: int, heads: int, ff: int, embed: int, decoder: bool,
                      dw_oihw: torch.Tensor, w_scale: float, dbias: torch.Tensor, bias_scale: float, tape=None) -> list:
    """Gradients of params[1:] (params[0] is the sincos table) given the gradient of the generated kernel / bias.
    ``tape``: the workspace hypernet_forward_taped returned for the same inputs (skips the forward re-run)."""
    _need_cuda(wvs, dw_oihw, dbias, *params)
    wvs = wvs.to(torch.float32).contiguous()
    c = wvs.numel()
    dw_oihw = dw_oihw.to(torch.float32).contiguous()
    dbias = dbias.to(torch.float32).contiguous()
    lib = _C.lib()
    ws_bytes = lib.eovae_hypernet_backward_workspace_bytes(c, d, ff, embed, num_layers)
    ws = tape if tape is not None else torch.empty((ws_bytes // 4 + 1,), dtype=torch.float32, device=wvs.device)
    grads = [None] + [torch.empty_like(p) for p in params[1:]]
    parr = (ctypes.c_void_p * len(params))(*[p.data_ptr() for p in params])
    garr = (ctypes.c_void_p * len(params))(*[None if g is None else g.data_ptr() for g in grads])
    rc = lib.eovae_hypernet_backward(_ptr(wvs), c, parr, num_layers, d, heads, ff, embed, 1 if decoder else 0, _ptr(dw_oihw),
                                     dw_oihw.shape[1], float(w_scale), _ptr(dbias), float(bias_scale), garr,
                                     1 if tape is not None else 0, _ptr(ws), ws_bytes, _stream())
    _C.check(rc, "eovae_hypernet_backward")
    return grads[1:]


def hypernet_factorized_forward(wvs: torch.Tensor, params: list, num_layers: int, d: int, heads: int, ff: int, embed: int,
                                rank: int, decoder: bool):
    """FactorizedWeightGenerator(_decoder): -> (wk [C, 9*embed], bias_raw, tape); the tape (activations) feeds
    hypernet_factorized_backward.  Parameter order: see eovae_hypernet_factorized_forward (include/eovae.h)."""
    _need_cuda(wvs, *params)
    wvs = wvs.to(torch.float32).contiguous()
    c = wvs.numel()
    lib = _C.lib()
    ws_bytes = lib.eovae_hypernet_factorized_workspace_bytes(c, d, ff, embed, rank, num_layers)
    tape = torch.empty((ws_bytes // 4 + 1,), dtype=torch.float32, device=wvs.device)
    wk = torch.empty((c, 9 * embed), dtype=torch.float32, device=wvs.device)
    bias = torch.empty((c if decoder else embed,), dtype=torch.float32, device=wvs.device)
    arr = (ctypes.c_void_p * len(params))(*[p.data_ptr() for p in params])
    rc = lib.eovae_hypernet_factorized_forward(_ptr(wvs), c, arr, num_layers, d, heads, ff, embed, rank, 1 if decoder else 0,
                                               _ptr(wk), _ptr(bias), _ptr(tape), ws_bytes, _stream())
    _C.check(rc, "eovae_hypernet_factorized_forward")
    return wk, bias, tape


def hypernet_factorized_backward(wvs: torch.Tensor, params: list, num_layers: int, d: int, heads: int, ff: int, embed: int,
                                 rank: int, decoder: bool, dw_oihw: torch.Tensor, w_scale: float, dbias: torch.Tensor,
                                 bias_scale: float, tape: torch.Tensor) -> list:
    """Gradients of params[1:] given the gradient of the generated kernel / bias; ``tape`` from the forward."""
    _need_cuda(wvs, dw_oihw, dbias, tape, *params)
    wvs = wvs.to(torch.float32).contiguous()
    c = wvs.numel()
    dw_oihw = dw_oihw.to(torch.float32).contiguous()
    dbias = dbias.to(torch.float32).contiguous()
    lib = _C.lib()
    ws_bytes = lib.eovae_hypernet_factorized_workspace_bytes(c, d, ff, embed, rank, num_layers)
    if tape.numel() * 4 < ws_bytes:
        raise RuntimeError("hypernet_factorized_backward: tape does not belong to this configuration")
    grads = [None] + [torch.empty_like(p) for p in params[1:]]
    parr = (ctypes.c_void_p * len(params))(*[p.data_ptr() for p in params])
    garr = (ctypes.c_void_p * len(params))(*[None if g is None else g.data_ptr() for g in grads])
    rc = lib.eovae_hypernet_factorized_backward(_ptr(wvs), c, parr, num_layers, d, heads, ff, embed, rank,
                                                1 if decoder else 0, _ptr(dw_oihw), dw_oihw.shape[1], float(w_scale),
                                                _ptr(dbias), float(bias_scale), garr, _ptr(tape), ws_bytes, _stream())
    _C.check(rc, "eovae_hypernet_factorized_backward")
    return grads[1:]


def wavelength_style_forward(wvs: torch.Tensor, params: list, d: int):
    """WavelengthConditioner (model.py:35-64): -> (style [1, d] fp32, tape).  params: omega, mlp.0/2/4 (w, b)."""
    _need_cuda(wvs, *params)
    wvs = wvs.to(torch.float32).contiguous()
    lib = _C.lib()
    ws_bytes = lib.eovae_wavelength_style_workspace_bytes(d)
    tape = torch.empty((ws_bytes // 4,), dtype=torch.float32, device=wvs.device)
    style = torch.empty((1, d), dtype=torch.float32, device=wvs.device)
    arr = (ctypes.c_void_p * len(params))(*[p.data_ptr() for p in params])
    _C.check(lib.eovae_wavelength_style_forward(_ptr(wvs), wvs.numel(), arr, d, _ptr(style), _ptr(tape), ws_bytes, _stream()),
             "eovae_wavelength_style_forward")
    return style, tape


def wavelength_style_backward(params: list, d: int, dstyle: torch.Tensor, tape: torch.Tensor) -> list:
    """-> gradients of params[1:] (mlp.0/2/4 weights and biases)."""
    _need_cuda(dstyle, tape, *params)
    dstyle = dstyle.to(torch.float32).contiguous()
    lib = _C.lib()
    grads = [None] + [torch.empty_like(p) for p in params[1:]]
    parr = (ctypes.c_void_p * len(params))(*[p.data_ptr() for p in params])
    garr = (ctypes.c_void_p * len(params))(*[None if g is None else g.data_ptr() for g in grads])
    _C.check(lib.eovae_wavelength_style_backward(parr, d, _ptr(dstyle), garr, _ptr(tape), tape.numel() * 4, _stream()),
             "eovae_wavelength_style_backward")
    return grads[1:]


def adain_affine_forward(style: torch.Tensor, wproj: torch.Tensor, bproj: torch.Tensor, gamma: torch.Tensor,
                         beta: torch.Tensor):
    """[scale | shift] = emb_proj(style); -> (gamma * scale, beta * scale + shift, style2) (layers.py:96-104 folded)."""
    _need_cuda(style, wproj, bproj, gamma, beta)
    cout, d = gamma.numel(), style.numel()
    if tuple(wproj.shape) != (2 * cout, d):
        raise RuntimeError(f"adain_affine: emb_proj weight {tuple(wproj.shape)} does not match ({2 * cout}, {d})")
    g_out, b_out = torch.empty_like(gamma), torch.empty_like(beta)
    style2 = torch.empty((2 * cout,), dtype=torch.float32, device=style.device)
    _C.check(_C.lib().eovae_adain_affine_forward(_ptr(style), d, _ptr(wproj), _ptr(bproj), _ptr(gamma), _ptr(beta), cout,
                                                 _ptr(g_out), _ptr(b_out), _ptr(style2), _stream()),
             "eovae_adain_affine_forward")
    return g_out, b_out, style2


def adain_affine_backward(style, wproj, gamma, beta, style2, dg_out, db_out):
    """-> (dgamma, dbeta, dwproj, dbproj, dstyle [1, d])."""
    _need_cuda(style, wproj, gamma, beta, style2, dg_out, db_out)
    cout, d = gamma.numel(), style.numel()
    dg_out = dg_out.to(torch.float32).contiguous()
    db_out = db_out.to(torch.float32).contiguous()
    dgamma, dbeta = torch.empty_like(gamma), torch.empty_like(beta)
    dw = torch.empty_like(wproj)
    dbp = torch.empty((2 * cout,), dtype=torch.float32, device=style.device)
    dstyle = torch.empty((1, d), dtype=torch.float32, device=style.device)
    dstyle2 = torch.empty((2 * cout,), dtype=torch.float32, device=style.device)
    _C.check(_C.lib().eovae_adain_affine_backward(_ptr(style), d, _ptr(wproj), _ptr(gamma), _ptr(beta), _ptr(style2), cout,
                                                  _ptr(dg_out), _ptr(db_out), _ptr(dgamma), _ptr(dbeta), _ptr(dw), _ptr(dbp),
                                                  _ptr(dstyle), _ptr(dstyle2), _stream()),
             "eovae_adain_affine_backward")
    return dgamma, dbeta, dw, dbp, dstyle


def pack_dyn_weight(wk: torch.Tensor, bias_raw: torch.Tensor, c: int, embed: int, decoder: bool, scale: float,
                    bias_scale: float, dtype, want_oihw: bool):
    """generated kernel -> (igemm B operand, scaled bias, optional fp32 OIHW weight)."""
    _need_cuda(wk, bias_raw)
    rows, cin = (c, embed) if decoder else (embed, c)
    rows_pad = (rows + 15) // 16 * 16
    kpt = conv_k_per_tap(dyn_cin_pad(cin))
    packed = torch.empty((rows_pad, 9, kpt), dtype=dtype, device=wk.device)
    oihw = torch.empty((rows, cin, 3, 3), dtype=torch.float32, device=wk.device) if want_oihw else None
    bias = torch.empty_like(bias_raw)
    rc = _C.lib().eovae_pack_dyn_weight(_ptr(wk), c, embed, 1 if decoder else 0, float(scale), _ptr(packed), DT[dtype],
                                       kpt, rows_pad, _ptr(oihw), _ptr(bias_raw), float(bias_scale), _ptr(bias),
                                       bias.numel(), _stream())
    _C.check(rc, "eovae_pack_dyn_weight")
    return packed, bias, oihw


def dyn_cin_pad(cin: int) -> int:
    """channel padding of the NHWC operand of a dynamic conv (16 = one 32-byte TMA/UMMA K-chunk)."""
    return (cin + 15) // 16 * 16
