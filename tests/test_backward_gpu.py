"""Backward kernels (first pieces of the training path) against torch autograd in fp32 on the same 16-bit-rounded
operands: conv data gradient (implicit GEMM on flipped weights, incl. the stride-2 and upsample adjoints), conv weight
gradient (tcgen05, contraction over pixels), bias gradient, GroupNorm(+SiLU) backward."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-30))


def _act(n, c, h, w, dev, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (scale * torch.randn((n, c, h, w), generator=g)).to(dev).to(dtype=torch.bfloat16, memory_format=torch.channels_last)


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 32, 32, 64, 128), (1, 64, 64, 128, 128), (2, 16, 16, 256, 512),
                                            (1, 8, 128, 64, 128), (2, 16, 16, 32, 64)])
@pytest.mark.parametrize("mode", ["3x3", "1x1", "s2"])
def test_conv_dgrad(cuda, n, h, w, cin, cout, mode):
    from eo_vae import ops
    k = 1 if mode == "1x1" else 3
    wgt = (torch.randn(cout, cin, k, k) / math.sqrt(cin * k * k)).to(cuda)
    x = _act(n, cin, h, w, cuda, seed=1).float().requires_grad_(True)
    w16 = wgt.bfloat16().float()
    if mode == "3x3":
        y = F.conv2d(x, w16, padding=1)
    elif mode == "1x1":
        y = F.conv2d(x, w16)
    else:
        y = F.conv2d(F.pad(x, (0, 1, 0, 1)), w16, stride=2)
    dy = _act(*y.shape, cuda, seed=2)
    y.backward(dy.float())
    m = {"3x3": ops.CONV_3X3, "1x1": ops.CONV_1X1, "s2": ops.CONV_3X3_S2}[mode]
    dx = ops.conv2d_dgrad(dy, wgt, m, in_hw=(h, w))
    assert dx.shape == x.shape
    assert _rel(dx, x.grad) < 4e-3
    add = _act(n, cin, h, w, cuda, seed=3)
    dx2 = ops.conv2d_dgrad(dy, wgt, m, in_hw=(h, w), grad_add=add)
    assert _rel(dx2, x.grad + add.float()) < 4e-3


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 32, 32, 64, 128), (1, 64, 64, 128, 128), (2, 16, 16, 256, 512),
                                            (4, 8, 128, 64, 256), (2, 16, 16, 32, 64), (3, 32, 32, 512, 64),
                                            (2, 64, 64, 16, 128), (1, 128, 128, 128, 24), (2, 8, 8, 8, 32),
                                            (2, 14, 14, 64, 128), (1, 28, 28, 128, 64), (2, 12, 20, 32, 64), (1, 7, 9, 16, 16)])
@pytest.mark.parametrize("k", [3, 1])
@pytest.mark.parametrize("nhwc", [True, False])
def test_conv_wgrad_and_bias_grad(cuda, monkeypatch, n, h, w, cin, cout, k, nhwc):
    """nhwc=True: operands read in place as MN-major tcgen05 operands; False: the transposed-copy kernel."""
    from eo_vae import ops
    monkeypatch.setattr(ops, "USE_WGRAD_NHWC", nhwc)
    x = _act(n, cin, h, w, cuda, seed=5)
    wgt = (torch.randn(cout, cin, k, k) / math.sqrt(cin * k * k)).to(cuda).requires_grad_(True)
    bias = torch.zeros(cout, device=cuda, requires_grad=True)
    y = F.conv2d(x.float(), wgt, bias, padding=k // 2)
    dy = _act(*y.shape, cuda, seed=6)
    y.backward(dy.float())
    dw = ops.conv2d_wgrad(x, dy, k)
    assert dw.shape == wgt.shape
    assert _rel(dw, wgt.grad) < 2e-5, _rel(dw, wgt.grad)
    dw_acc = ops.conv2d_wgrad(x, dy, k, dw=dw.clone())
    assert _rel(dw_acc, 2 * wgt.grad) < 2e-5
    assert torch.equal(ops.conv2d_wgrad(x, dy, k), dw), "wgrad must be deterministic"
    assert _rel(ops.bias_grad(dy), bias.grad) < 1e-5


@pytest.mark.parametrize("n,c,h,w", [(2, 128, 32, 32), (3, 512, 8, 8), (1, 256, 64, 64), (2, 32, 16, 16), (2, 64, 12, 20)])
@pytest.mark.parametrize("silu", [True, False])
def test_group_norm_backward(cuda, n, c, h, w, silu):
    from eo_vae import ops
    x = (_act(n, c, h, w, cuda, seed=7, scale=2.0).float() + 0.5).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    gamma = (1 + 0.2 * torch.randn(c)).to(cuda).requires_grad_(True)
    beta = (0.2 * torch.randn(c)).to(cuda).requires_grad_(True)
    xf = x.float().requires_grad_(True)
    y = F.group_norm(xf, 32, gamma, beta, eps=1e-6)
    if silu:
        y = y * torch.sigmoid(y)
    g = _act(n, c, h, w, cuda, seed=8)
    y.backward(g.float())
    stats = ops.gn_stats(x)
    gx, dg, db = ops.gn_backward(x, g, stats, gamma.detach(), beta.detach(), silu)
    assert _rel(gx, xf.grad) < 6e-3
    assert _rel(dg, gamma.grad) < 2e-3 and _rel(db, beta.grad) < 2e-3
    add = _act(n, c, h, w, cuda, seed=9)
    gx2, _, _ = ops.gn_backward(x, g, stats, gamma.detach(), beta.detach(), silu, grad_add=add)
    assert _rel(gx2, xf.grad + add.float()) < 6e-3


def test_upsample_adjoint(cuda):
    from eo_vae import ops
    x = _act(2, 64, 6, 10, cuda, seed=11).float().requires_grad_(True)
    up = F.interpolate(x, scale_factor=2.0, mode="nearest")
    g = _act(2, 64, 12, 20, cuda, seed=12)
    up.backward(g.float())
    assert _rel(ops.pool2x2_sum(g), x.grad) < 4e-3


@pytest.mark.parametrize("silu", [True, False])
@pytest.mark.parametrize("n,c,h,w", [(3, 128, 96, 80), (2, 256, 64, 72), (5, 512, 32, 40), (16, 128, 64, 64)])
def test_gn_backward_fused_equals_two_pass(cuda, silu, n, c, h, w):
    """The fused single-kernel GroupNorm backward (x and g from HBM once, pass B out of L2, per-image block rendezvous) against the
    two-pass kernels and against torch autograd: dx, dgamma, dbeta, the bias column sums, with a residual-branch gradient added."""
    from eo_vae import ops
    import torch.nn.functional as F
    assert n * c * h * w >= 1 << 21
    g0 = torch.Generator().manual_seed(c + h)
    x = (torch.randn((n, h, w, c), generator=g0) * 1.5 + 0.3).to(cuda).bfloat16().permute(0, 3, 1, 2)
    gy = torch.randn((n, h, w, c), generator=g0).to(cuda).bfloat16().permute(0, 3, 1, 2)
    ga = torch.randn((n, h, w, c), generator=g0).to(cuda).bfloat16().permute(0, 3, 1, 2)
    gamma = (1.0 + 0.1 * torch.randn((c,), generator=g0)).to(cuda)
    beta = (0.1 * torch.randn((c,), generator=g0)).to(cuda)
    stats = ops.gn_stats(x)
    outs = []
    for fused in (1, 0):
        ops.set_tuning(ops.TUNE_GN_BWD_FUSED, fused)
        try:
            dx, dg, db = ops.gn_backward(x, gy, stats, gamma, beta, silu, 32, grad_add=ga)
            outs.append((dx.float().clone(), dg.clone(), db.clone(), dx._colsum.clone()))
        finally:
            ops.set_tuning(ops.TUNE_GN_BWD_FUSED, 1)
    for a, b in zip(*outs):
        assert _rel(a, b) < 2e-3, _rel(a, b)     # same math, different (fixed) summation orders; dx rounds to bf16
    xr = x.float().detach().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y = F.group_norm(xr, 32, gr, br, eps=1e-6)
    y = y * torch.sigmoid(y) if silu else y
    y.backward(gy.float())
    dx, dg, db, cs = outs[0]
    assert _rel(dx, xr.grad + ga.float()) < 8e-3
    assert _rel(dg, gr.grad) < 5e-3 and _rel(db, br.grad) < 5e-3
    assert _rel(cs, dx.sum(dim=(0, 2, 3))) < 5e-3
    # determinism: the rendezvous order must not leak into the bits
    dx2 = ops.gn_backward(x, gy, stats, gamma, beta, silu, 32, grad_add=ga)[0]
    assert torch.equal(dx2.float(), dx)
