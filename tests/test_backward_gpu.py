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


@pytest.mark.parametrize("n,c,h,w", [(2, 128, 64, 64), (1, 64, 37, 29), (3, 512, 24, 24), (2, 256, 128, 128), (1, 8, 5, 7)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_group_norm_backward_staged_equals_register_path(cuda, n, c, h, w, dtype):
    """The cp.async.bulk staged kernels (default) and the register-load kernels walk every thread's pixels in the same
    order: identical bits, over block sizes that give short last stages, one-stage blocks and ragged last blocks."""
    from eo_vae import ops
    x = (_act(n, c, h, w, cuda, seed=11, scale=2.0).float() + 0.25).to(dtype).contiguous(memory_format=torch.channels_last)
    g = _act(n, c, h, w, cuda, seed=12).to(dtype)
    add = _act(n, c, h, w, cuda, seed=13).to(dtype)
    groups = min(32, c // 8) if c < 32 else 32
    gamma, beta = (1 + 0.2 * torch.randn(c)).to(cuda), (0.2 * torch.randn(c)).to(cuda)
    stats = ops.gn_stats(x, groups)
    try:
        for elems in (0, 4096, 40960, 131072):
            ops.set_tuning(ops.TUNE_GN_BWD_BLOCK_ELEMS, elems)
            outs = []
            for bulk in (0, 1):
                ops.set_tuning(ops.TUNE_GN_BWD_BULK, bulk)
                for silu in (True, False):
                    outs.append((bulk, ops.gn_backward(x, g, stats, gamma, beta, silu, groups=groups),
                                 ops.gn_backward(x, g, stats, gamma, beta, silu, groups=groups, grad_add=add)))
            half = len(outs) // 2
            for (_, a0, a1), (_, b0, b1) in zip(outs[:half], outs[half:]):
                for t0, t1 in zip(a0 + a1, b0 + b1):
                    assert torch.isfinite(t1.float()).all()
                    assert torch.equal(t0, t1)
    finally:
        ops.set_tuning(ops.TUNE_GN_BWD_BLOCK_ELEMS, 0)
        ops.set_tuning(ops.TUNE_GN_BWD_BULK, 1)


def test_upsample_adjoint(cuda):
    from eo_vae import ops
    x = _act(2, 64, 6, 10, cuda, seed=11).float().requires_grad_(True)
    up = F.interpolate(x, scale_factor=2.0, mode="nearest")
    g = _act(2, 64, 12, 20, cuda, seed=12)
    up.backward(g.float())
    assert _rel(ops.pool2x2_sum(g), x.grad) < 4e-3
