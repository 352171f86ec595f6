"""Per-kernel parity on the GPU: every C-ABI entry point against a plain PyTorch fp32 statement of the same op
on the same (already 16-bit-rounded) operands, so the tolerances only cover accumulation order and the output
rounding.  Everything goes through eo_vae.ops -> ctypes -> libeovae_sm100.so."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-30))


def _act(n, c, h, w, dev, dtype=torch.bfloat16, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn((n, c, h, w), generator=g).to(dev)
    return x.to(dtype=dtype, memory_format=torch.channels_last)


CONV_CASES = [
    # n, h, w, cin, cout
    (2, 32, 32, 64, 64),
    (1, 64, 64, 128, 128),
    (2, 16, 16, 512, 512),
    (1, 256, 256, 128, 128),
    (3, 24, 24, 32, 64),     # 64-byte K chunks
    (2, 8, 8, 16, 32),       # 32-byte K chunks, several images per tile
    (1, 48, 40, 64, 256),    # ragged tiles
    (2, 32, 32, 256, 512),
    (5, 4, 4, 64, 16),       # tiny maps, Cout 16
    (2, 16, 16, 8, 32),      # Cin below one K chunk (TMA channel OOB fill)
    (1, 6, 128, 64, 128),    # m-tile = one image row: halo-reuse mainloop (three taps from one shared-memory tile)
    (2, 4, 256, 128, 256),
    (1, 3, 200, 64, 128),    # halo mainloop with a ragged right edge
]


@pytest.fixture(params=[1, 2, 5, 6, 10, 18], ids=["1cta", "2cta", "1cta-k64", "2cta-k64", "2cta-stg", "2cta-nohalo"])
def cta_group(request, cuda):
    """Run the implicit GEMM as single CTAs and as CTA pairs (tcgen05 cta_group::2); shapes that cannot pair fall
    back to single CTAs inside the library."""
    from eo_vae import _C
    _C.lib().eovae_set_debug_mode(request.param << 8)
    yield request.param
    _C.lib().eovae_set_debug_mode(0)


@pytest.mark.parametrize("n,h,w,cin,cout", CONV_CASES)
@pytest.mark.parametrize("mode", ["3x3", "1x1", "s2"])
def test_conv2d(cuda, cta_group, n, h, w, cin, cout, mode):
    from eo_vae import ops
    torch.manual_seed(1)
    x = _act(n, cin, h, w, cuda, seed=n * 7 + h)
    k = 1 if mode == "1x1" else 3
    wgt = (torch.randn(cout, cin, k, k) / math.sqrt(cin * k * k)).to(cuda)
    bias = torch.randn(cout).to(cuda)
    wp = ops.pack_conv_weight(wgt, torch.bfloat16)
    m = {"3x3": ops.CONV_3X3, "1x1": ops.CONV_1X1, "s2": ops.CONV_3X3_S2}[mode]
    w16 = wgt.bfloat16().float()
    xf = x.float()
    if mode == "3x3":
        ref = F.conv2d(xf, w16, bias, padding=1)
    elif mode == "1x1":
        ref = F.conv2d(xf, w16, bias)
    else:
        ref = F.conv2d(F.pad(xf, (0, 1, 0, 1)), w16, bias, stride=2)
    out = ops.conv2d(x, wp, bias, cout, m, out_dtype=torch.float32)
    assert out.shape == ref.shape
    assert _rel(out, ref) < 2e-5, f"fp32-out conv mismatch {_rel(out, ref)}"
    # fused residual + 16-bit output
    res = _act(*ref.shape[:2], *ref.shape[2:], cuda, seed=99).contiguous(memory_format=torch.channels_last)
    out2 = ops.conv2d(x, wp, bias, cout, m, residual=res)
    assert out2.dtype == torch.bfloat16
    assert _rel(out2, ref + res.float()) < 4e-3


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 32, 32, 64, 128), (1, 64, 64, 128, 256), (3, 16, 16, 128, 512),
                                            (2, 24, 40, 32, 64), (2, 16, 16, 16, 32), (1, 256, 256, 16, 128)])
@pytest.mark.parametrize("mode", ["3x3", "s2"])
def test_conv2d_fused_gn_stats(cuda, cta_group, n, h, w, cin, cout, mode):
    """The epilogue's GroupNorm statistics of the conv OUTPUT (incl. bias + residual) match a two-pass fp32 statement
    and are bit-reproducible."""
    from eo_vae import ops
    x = _act(n, cin, h, w, cuda, seed=3)
    wgt = (torch.randn(cout, cin, 3, 3) / math.sqrt(cin * 9)).to(cuda)
    bias = torch.randn(cout).to(cuda)
    wp = ops.pack_conv_weight(wgt, torch.bfloat16)
    m = ops.CONV_3X3 if mode == "3x3" else ops.CONV_3X3_S2
    xf, w16 = x.float(), wgt.bfloat16().float()
    ref = F.conv2d(xf, w16, bias, padding=1) if mode == "3x3" else F.conv2d(F.pad(xf, (0, 1, 0, 1)), w16, bias, stride=2)
    res = _act(*ref.shape, cuda, seed=11)
    ref = ref + res.float()
    out = ops.conv2d(x, wp, bias, cout, m, residual=res, gn_groups=32)
    from eo_vae import _C
    if _C.lib().eovae_conv2d_gn_workspace_bytes(n, h, w, m, cout, 32) == 0:
        assert not hasattr(out, "_gn_stats")  # several images per tile: the caller falls back to eovae_gn_stats
        return
    assert hasattr(out, "_gn_stats"), "fused statistics were not produced for a supported shape"
    stats = out._gn_stats[0]
    rf = ref.reshape(ref.shape[0], 32, -1)
    assert torch.allclose(stats[..., 0], rf.mean(-1), atol=2e-4)
    assert torch.allclose(stats[..., 1], 1 / torch.sqrt(rf.var(-1, unbiased=False) + 1e-6), rtol=2e-4)
    out2 = ops.conv2d(x, wp, bias, cout, m, residual=res, gn_groups=32)
    assert torch.equal(out2._gn_stats[0], stats) and torch.equal(out2, out)
    # group_norm() consumes them: same result as recomputing from the stored tensor up to the 16-bit rounding
    gamma, beta = torch.ones(cout, device=cuda), torch.zeros(cout, device=cuda)
    y = ops.group_norm(out, gamma, beta, True)
    y_ref = F.group_norm(ref, 32, gamma, beta, eps=1e-6)
    assert _rel(y, y_ref * torch.sigmoid(y_ref)) < 6e-3


@pytest.mark.parametrize("n,h,w,cin2,c", [(2, 32, 32, 128, 256), (1, 64, 64, 256, 512), (3, 16, 16, 64, 128), (2, 8, 8, 64, 64)])
def test_conv2d_fused_shortcut(cuda, cta_group, n, h, w, cin2, c):
    """ResnetBlock tail as one implicit GEMM: conv3x3(h) + conv1x1(x) + biases (+ GroupNorm statistics)."""
    from eo_vae import ops
    hh = _act(n, c, h, w, cuda, seed=21)
    x = _act(n, cin2, h, w, cuda, seed=22)
    w3 = (torch.randn(c, c, 3, 3) / math.sqrt(9 * c)).to(cuda)
    w1 = (torch.randn(c, cin2, 1, 1) / math.sqrt(cin2)).to(cuda)
    b3, b1 = torch.randn(c).to(cuda), torch.randn(c).to(cuda)
    p3, p1 = ops.pack_conv_weight(w3, torch.bfloat16), ops.pack_conv_weight(w1, torch.bfloat16)
    wf = torch.cat([p3.reshape(p3.shape[0], -1), p1.reshape(p1.shape[0], -1)], dim=1).contiguous()
    out = ops.conv2d(hh, wf, (b3 + b1).contiguous(), c, ops.CONV_3X3, out_dtype=torch.float32, x2=x, gn_groups=32)
    ref = F.conv2d(hh.float(), w3.bfloat16().float(), b3, padding=1) + F.conv2d(x.float(), w1.bfloat16().float(), b1)
    assert _rel(out, ref) < 2e-5
    if hasattr(out, "_gn_stats"):
        rf = ref.reshape(n, 32, -1)
        assert torch.allclose(out._gn_stats[0][..., 0], rf.mean(-1), atol=2e-4)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("n,h,w,cin,cout", [(1, 6, 128, 64, 128), (2, 5, 256, 128, 128), (2, 4, 128, 256, 256),
                                            (1, 3, 200, 128, 256), (3, 2, 384, 64, 128)])
def test_conv2d_gn_prologue(cuda, n, h, w, cin, cout, dtype):
    """conv3x3(silu(GroupNorm(x))) with the normalisation applied to the operand tiles inside the mainloop (transform
    warps) vs the fp32 statement and vs the unfused kernels; also with residual + next-GN statistics in the epilogue."""
    from eo_vae import ops
    x = (_act(n, cin, h, w, cuda, dtype=dtype, seed=31) * 2.5 + 0.7).contiguous(memory_format=torch.channels_last)
    gamma = (1 + 0.2 * torch.randn(cin)).to(cuda)
    beta = (0.3 * torch.randn(cin)).to(cuda)
    wgt = (torch.randn(cout, cin, 3, 3) / math.sqrt(9 * cin)).to(cuda)
    bias = torch.randn(cout).to(cuda)
    wp = ops.pack_conv_weight(wgt, dtype)
    assert ops.gn_prologue_ok(x, cout, ops.CONV_3X3)
    stats = ops.gn_stats(x)
    res = _act(n, cout, h, w, cuda, dtype=dtype, seed=32)
    fused = ops.conv2d(x, wp, bias, cout, ops.CONV_3X3, residual=res, out_dtype=torch.float32, gn_groups=32,
                       in_gn=(stats, gamma, beta, 32))
    y = F.group_norm(x.float(), 32, gamma, beta, eps=1e-6)
    y = (y * torch.sigmoid(y)).to(dtype).float()
    ref = F.conv2d(y, wgt.to(dtype).float(), bias, padding=1) + res.float()
    unfused = ops.conv2d(ops.gn_apply(x, stats, gamma, beta, True), wp, bias, cout, ops.CONV_3X3, residual=res,
                         out_dtype=torch.float32)
    # bf16: fp32 transform, operand may differ by one ulp; fp16: packed half2 transform (3 fp16 roundings + tanh.approx.f16x2)
    tol = 3e-3 if dtype == torch.bfloat16 else 2e-3
    assert _rel(fused, ref) < tol, _rel(fused, ref)
    assert _rel(fused, unfused) < tol
    rf = ref.reshape(n, 32, -1)
    assert torch.allclose(fused._gn_stats[0][..., 0], rf.mean(-1), atol=3e-3)


def test_conv2d_fp16_operands(cuda):
    from eo_vae import ops
    x = _act(2, 64, 16, 16, cuda, dtype=torch.float16)
    wgt = (torch.randn(64, 64, 3, 3) / 24).to(cuda)
    wp = ops.pack_conv_weight(wgt, torch.float16)
    out = ops.conv2d(x, wp, None, 64, ops.CONV_3X3, out_dtype=torch.float32)
    ref = F.conv2d(x.float(), wgt.half().float(), None, padding=1)
    assert _rel(out, ref) < 2e-5


def test_conv2d_channel_slices(cuda):
    """A conv may read a channel slice of a wider NHWC tensor and write into one."""
    from eo_vae import ops
    big = _act(2, 192, 16, 16, cuda)
    x = big[:, 64:128]
    wgt = (torch.randn(32, 64, 1, 1) / 8).to(cuda)
    wp = ops.pack_conv_weight(wgt, torch.bfloat16)
    out = ops.conv2d(x, wp, None, 32, ops.CONV_1X1, out_dtype=torch.float32)
    ref = F.conv2d(x.float(), wgt.bfloat16().float())
    assert _rel(out, ref) < 2e-5


@pytest.mark.parametrize("b,m,n,k", [(2, 256, 256, 64), (3, 1024, 1024, 512), (1, 100, 72, 32), (4, 64, 512, 1024)])
def test_gemm_tn_batched(cuda, cta_group, b, m, n, k):
    from eo_vae import ops
    torch.manual_seed(2)
    a = torch.randn(b, m, k, device=cuda).bfloat16()
    bb = torch.randn(b, n, k, device=cuda).bfloat16()
    c = ops.gemm_tn_batched(a, bb, torch.float32, scale=0.125)
    ref = 0.125 * torch.einsum("bmk,bnk->bmn", a.float(), bb.float())
    assert _rel(c, ref) < 2e-5


@pytest.mark.parametrize("n,c,h,w", [(2, 128, 32, 32), (3, 512, 8, 8), (1, 256, 64, 64), (2, 32, 16, 16), (2, 64, 12, 20)])
def test_group_norm(cuda, n, c, h, w):
    from eo_vae import ops
    x = _act(n, c, h, w, cuda, seed=5) * 3 + 1.5
    gamma = (1 + 0.1 * torch.randn(c)).to(cuda)
    beta = (0.1 * torch.randn(c)).to(cuda)
    stats = ops.gn_stats(x)
    xf = x.float().reshape(n, 32, -1)
    assert torch.allclose(stats[..., 0], xf.mean(-1), atol=1e-4)
    assert torch.allclose(stats[..., 1], 1 / torch.sqrt(xf.var(-1, unbiased=False) + 1e-6), rtol=1e-4)
    for silu in (False, True):
        y = ops.gn_apply(x, stats, gamma, beta, silu)
        ref = F.group_norm(x.float(), 32, gamma, beta, eps=1e-6)
        if silu:
            ref = ref * torch.sigmoid(ref)
        assert _rel(y, ref) < 4e-3


@pytest.mark.parametrize("n,c,h,w", [(2, 128, 64, 64), (1, 64, 37, 29), (3, 512, 24, 24), (1, 8, 5, 7), (2, 256, 16, 16)])
@pytest.mark.parametrize("dt_in,dt_out", [(torch.bfloat16, torch.bfloat16), (torch.float16, torch.float16),
                                          (torch.float16, torch.bfloat16)])
def test_gn_apply_ring_equals_register_path(cuda, n, c, h, w, dt_in, dt_out):
    """The cp.async.bulk ring kernel (default) and the register-load kernels compute every element with the same
    instructions: identical bits, over block sizes that give one-slot blocks, short last slots and ragged last blocks; also
    with the output written as a channel slice of a wider tensor (through the C ABI)."""
    from eo_vae import _C, ops
    from eo_vae.ops import DT, _ptr, _stream
    groups = 32 if c >= 32 else c // 8
    x = (_act(n, c, h, w, cuda, seed=21).float() * 2 + 0.5).to(dt_in).contiguous(memory_format=torch.channels_last)
    gamma, beta = (1 + 0.1 * torch.randn(c)).to(cuda), (0.1 * torch.randn(c)).to(cuda)
    stats = ops.gn_stats(x, groups)
    try:
        ops.set_tuning(ops.TUNE_GN_APPLY_CORESIDENT, 4)
        ref = {silu: ops.gn_apply(x, stats, gamma, beta, silu, groups, out_dtype=dt_out) for silu in (False, True)}
        ops.set_tuning(ops.TUNE_GN_APPLY_CORESIDENT, 0)
        for elems in (0, 2048, 40960):
            ops.set_tuning(ops.TUNE_GN_APPLY_BLOCK_ELEMS, elems)
            for silu in (False, True):
                y = ops.gn_apply(x, stats, gamma, beta, silu, groups, out_dtype=dt_out)
                assert torch.isfinite(y.float()).all() and torch.equal(y, ref[silu])
        # output = channels [c, 2c) of a (n, h, w, 3c) tensor
        wide = torch.zeros((n, h, w, 3 * c), dtype=dt_out, device=cuda)
        ysl = wide[..., c:2 * c]
        rc = _C.lib().eovae_gn_apply(_ptr(x), DT[x.dtype], c, _ptr(stats), _ptr(gamma), _ptr(beta), _ptr(ysl), DT[dt_out], 3 * c,
                                    n, h * w, c, groups, 1, _stream())
        _C.check(rc, "eovae_gn_apply")
        assert torch.equal(ysl.permute(0, 3, 1, 2), ref[True])
        assert float(wide[..., :c].abs().max()) == 0.0 and float(wide[..., 2 * c:].abs().max()) == 0.0
    finally:
        ops.set_tuning(ops.TUNE_GN_APPLY_CORESIDENT, 0)
        ops.set_tuning(ops.TUNE_GN_APPLY_BLOCK_ELEMS, 0)


def test_softmax_transpose_upsample(cuda):
    from eo_vae import ops
    s = torch.randn(3, 200, 1024, device=cuda) * 4
    p = ops.softmax_rows(s, torch.bfloat16)
    assert _rel(p, torch.softmax(s, -1)) < 4e-3
    s2 = torch.randn(2, 8, 5000, device=cuda)
    assert _rel(ops.softmax_rows(s2, torch.bfloat16), torch.softmax(s2, -1)) < 4e-3
    s3 = torch.randn(2, 50, 324, device=cuda)
    p3 = ops.softmax_rows(s3, torch.bfloat16, cols=324, out_cols=336)
    assert p3.shape == (2, 50, 336) and float(p3[..., 324:].abs().max()) == 0.0
    assert _rel(p3[..., :324], torch.softmax(s3, -1)) < 4e-3
    vt = ops.transpose16(torch.randn(2, 324, 64, device=cuda).bfloat16(), out_rows=336)
    assert vt.shape == (2, 64, 336) and float(vt[..., 324:].abs().max()) == 0.0
    t = torch.randn(3, 70, 96 * 3, device=cuda).bfloat16()
    v = t[:, :, 96:192]
    assert torch.equal(ops.transpose16(v), v.transpose(1, 2).contiguous())
    x = _act(2, 64, 6, 10, cuda)
    up = ops.upsample2x(x)
    assert torch.equal(up, F.interpolate(x.float(), scale_factor=2.0, mode="nearest").to(x.dtype))


def test_layout_edges(cuda):
    from eo_vae import ops
    x = torch.randn(3, 12, 20, 28, device=cuda)
    a = ops.nchw_to_act(x, 16, torch.bfloat16)
    assert a.shape == (3, 16, 20, 28)
    assert torch.equal(a[:, :12].float(), x.bfloat16().float())
    assert float(a[:, 12:].abs().max()) == 0.0
    back = ops.act_to_nchw_f32(a, 12)
    assert back.is_contiguous() and torch.equal(back, x.bfloat16().float())


def test_latent_glue(cuda):
    from eo_vae import ops
    from oracle import eovae_oracle as O
    n, zc, h, w = 3, 8, 6, 10
    moments = torch.randn(n, 2 * zc, h, w, device=cuda)
    moments_cl = moments.contiguous(memory_format=torch.channels_last)
    sd = {"bn.running_mean": torch.randn(4 * zc), "bn.running_var": torch.rand(4 * zc) + 0.5}
    rm, rv = sd["bn.running_mean"].to(cuda), sd["bn.running_var"].to(cuda)
    ref = O.pixel_shuffle2(O.bn_eval(sd, O.pixel_unshuffle2(moments.cpu()[:, :zc])))
    for mom in (moments, moments_cl):
        z = ops.latent_norm(mom, rm, rv, 1e-5, zc)
        assert torch.allclose(z.cpu(), ref, atol=1e-5, rtol=1e-5)
    eps = torch.randn(n, zc, h, w)
    zs, kl = ops.kl_reparam(moments_cl, eps.to(cuda), zc)
    assert torch.allclose(zs.cpu(), O.posterior_sample(moments.cpu(), eps), atol=1e-5, rtol=1e-5)
    assert torch.allclose(kl.cpu(), O.posterior_kl(moments.cpu()), rtol=1e-5)
    zn = torch.randn(n, zc, h, w, device=cuda)
    d = ops.latent_denorm(zn, rm, rv, 1e-4, torch.bfloat16)
    ref_d = O.pixel_shuffle2(O.bn_inverse(sd, O.pixel_unshuffle2(zn.cpu())))
    assert _rel(d.cpu(), ref_d) < 4e-3


def test_pixel_losses(cuda):
    from eo_vae import ops
    a = torch.randn(2, 12, 33, 47, device=cuda)
    b = torch.randn(2, 12, 33, 47, device=cuda)
    out = ops.l1_charbonnier(a, b, 1e-3).cpu()
    assert abs(float(out[0]) - float((a - b).abs().mean())) < 1e-5
    assert abs(float(out[1]) - float(torch.sqrt((a - b) ** 2 + 1e-6).mean())) < 1e-5


@pytest.mark.parametrize("b,c,h,w", [(2, 12, 256, 256), (3, 2, 192, 176), (1, 3, 512, 256)])
def test_msssim_and_consistency_loss(cuda, b, c, h, w):
    """MS-SSIM kernels and the EOConsistencyLoss forward against the CPU oracle (1e-3 relative, BASELINE tolerance)."""
    from eo_vae import ops
    from eo_vae.models.modules.consistency_loss import EOConsistencyLoss
    from oracle import eovae_oracle as O
    g = torch.Generator().manual_seed(b * h)
    x = torch.randn((b, c, h, w), generator=g).clamp_(-2, 6)
    r = x + 0.4 * torch.randn((b, c, h, w), generator=g)
    ref = float(O.ms_ssim(r, x))
    out, per = ops.msssim(r.to(cuda), x.to(cuda))
    assert abs(float(out) - ref) / ref < 1e-4
    assert per.shape == (b,) and abs(float(per.mean()) - ref) / ref < 1e-4
    assert abs(float(ops.msssim(x.to(cuda), x.to(cuda))[0]) - 1.0) < 1e-5
    for kind in ("l1", "char"):
        loss = EOConsistencyLoss(pixel_weight=1.0, rec_loss_type=kind, msssim_weight=1.0, msssim_start_step=0)
        total, logs = loss(x.to(cuda), None, r.to(cuda), global_step=3, split="train")
        t_ref, rec_ref, ms_ref = O.consistency_loss(x, r, kind, 1.0, 1.0, 3, 0)
        assert abs(float(total) - float(t_ref)) / float(t_ref) < 1e-3
        assert abs(float(logs["train/loss_rec"]) - float(rec_ref)) / float(rec_ref) < 1e-3
        assert abs(float(logs["train/loss_msssim"]) - float(ms_ref)) / float(ms_ref) < 1e-3
        assert set(logs) == {"train/loss_rec", "train/loss_msssim", "train/loss_total"}
    gated = EOConsistencyLoss(pixel_weight=1.0, msssim_weight=1.0, msssim_start_step=2000)
    _, logs = gated(x.to(cuda), None, r.to(cuda), global_step=10)
    assert "train/loss_msssim" not in logs


def test_msssim_large_batch_value_and_gradient(cuda):
    """More samples than one pass of the finalize / coefficient kernels takes (64): per-sample values, their mean and the
    gradient against autograd over the CPU oracle."""
    from eo_vae import ops
    from oracle import eovae_oracle as O
    b, c, h, w = 70, 1, 176, 176
    g = torch.Generator().manual_seed(5)
    x = torch.randn((b, c, h, w), generator=g).clamp_(-2, 6)
    r = (x + (0.1 + 0.4 * torch.rand((b, 1, 1, 1), generator=g)) * torch.randn((b, c, h, w), generator=g)).requires_grad_(True)
    ref = O.ms_ssim(r, x)
    ref.backward()
    out, per = ops.msssim(r.detach().to(cuda), x.to(cuda))
    assert abs(float(out) - float(ref)) / float(ref) < 1e-4
    per_ref = torch.stack([O.ms_ssim(r.detach()[i:i + 1], x[i:i + 1]) for i in range(b)])
    assert torch.allclose(per.cpu(), per_ref.float(), rtol=1e-4, atol=1e-6)
    gr = ops.msssim_backward(r.detach().to(cuda), x.to(cuda), 6.0, torch.ones(1, device=cuda))
    assert _rel(gr.cpu(), r.grad) < 1e-3


@pytest.mark.parametrize("modality", ["S2RGB", "S1RTC", "S2L2A", "S2L1C"])
@pytest.mark.parametrize("cfg_name", ["tiny", "full"])
def test_hypernet(cuda, modality, cfg_name):
    """Generated conv kernels vs the CPU oracle (fp32 both sides)."""
    import __graft_entry__ as g
    from oracle import eovae_oracle as O
    from oracle.weights import FULL_CONFIG, TINY_CONFIG, WAVELENGTHS, make_state_dict
    cfg = TINY_CONFIG if cfg_name == "tiny" else FULL_CONFIG
    sd = make_state_dict(cfg, 1)
    model = g._model(cfg, sd, cuda)
    wvs = torch.tensor(WAVELENGTHS[modality])
    w_ref, b_ref = O.hypernet(sd, "encoder.conv_in", wvs, False, cfg["hyper_heads"])
    w, b = model.encoder.conv_in.get_distillation_weight(wvs.to(cuda))
    assert _rel(w.cpu(), w_ref) < 2e-4 and _rel(b.cpu(), b_ref) < 2e-4
    w_ref, b_ref = O.hypernet(sd, "decoder.conv_out", wvs, True, cfg["hyper_heads"])
    w, b = model.decoder.conv_out.get_distillation_weight(wvs.to(cuda))
    assert _rel(w.cpu(), w_ref) < 2e-4
    assert _rel(b.cpu(), b_ref * 10.0) < 2e-4  # get_distillation_weight scales the bias once (x0.1), forward twice


@pytest.mark.parametrize("modality", ["S2RGB", "S1RTC", "S2L2A", "S2L1C"])
@pytest.mark.parametrize("cfg_name", ["tiny", "full"])
def test_hypernet_factorized(cuda, modality, cfg_name):
    """generator_type='factorized' (FactorizedWeightGenerator(_decoder), dynamic_conv.py:186-302: pre-norm layers,
    low-rank head): generated kernels / biases vs the CPU oracle (fp32 both sides; oracle pinned against the reference
    in tests/test_oracle_vs_reference.py)."""
    import __graft_entry__ as g
    from oracle import eovae_oracle as O
    from oracle.weights import FULL_CONFIG, TINY_FACTORIZED_CONFIG, WAVELENGTHS, make_state_dict
    cfg = TINY_FACTORIZED_CONFIG if cfg_name == "tiny" else dict(FULL_CONFIG, generator_type="factorized", rank_ratio=2)
    sd = make_state_dict(cfg, 1)
    model = g._model(cfg, sd, cuda)
    assert type(model.encoder.conv_in.weight_generator).__name__ == "FactorizedWeightGenerator"
    wvs = torch.tensor(WAVELENGTHS[modality])
    w_ref, b_ref = O.hypernet(sd, "encoder.conv_in", wvs, False, cfg["hyper_heads"])
    w, b = model.encoder.conv_in.get_distillation_weight(wvs.to(cuda))
    assert _rel(w.cpu(), w_ref) < 2e-4 and _rel(b.cpu(), b_ref) < 2e-4
    w_ref, b_ref = O.hypernet(sd, "decoder.conv_out", wvs, True, cfg["hyper_heads"])
    w, b = model.decoder.conv_out.get_distillation_weight(wvs.to(cuda))
    assert _rel(w.cpu(), w_ref) < 2e-4
    assert _rel(b.cpu(), b_ref * 10.0) < 2e-4


def test_factorized_model_forward_and_gradients(cuda):
    """Whole tiny model with factorized generators: latents / reconstruction (bf16 path) and, in train mode, every
    hypernetwork parameter gradient (fp32 kernels end to end) vs autograd over the oracle."""
    import __graft_entry__ as g
    from oracle import eovae_oracle as O
    from oracle.weights import TINY_FACTORIZED_CONFIG, WAVELENGTHS, make_state_dict, synthetic_patches
    cfg = TINY_FACTORIZED_CONFIG
    sd = make_state_dict(cfg, 4)
    model = g._model(cfg, sd, cuda)
    wvs = torch.tensor(WAVELENGTHS["S2L2A"])
    x = synthetic_patches(2, 12, cfg["resolution"], seed=21)
    with torch.no_grad():
        z = model.encode_spatial_normalized(x.to(cuda), wvs.to(cuda))
        r = model.reconstruct(x.to(cuda), wvs.to(cuda))
    z_ref = O.encode_spatial_normalized(sd, x, wvs, cfg["hyper_heads"])
    r_ref = O.reconstruct(sd, x, wvs, cfg["hyper_heads"])
    print(f"factorized tiny: latent {_rel(z.cpu(), z_ref):.3e} recon {_rel(r.cpu(), r_ref):.3e}")
    assert _rel(z.cpu(), z_ref) < 1.5e-2 and _rel(r.cpu(), r_ref) < 4e-2
    # gradients: deterministic forward (mode of the posterior), Charbonnier loss
    model.train()
    recon, _ = model(x.to(cuda), wvs.to(cuda), sample_posterior=False)
    torch.sqrt((recon - x.to(cuda)) ** 2 + 1e-6).mean().backward()
    osd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v.clone()) for k, v in sd.items()}
    rr, _ = O.forward(osd, x, wvs, None, True, cfg["hyper_heads"])
    O.charbonnier_loss(rr, x).backward()
    got, want = [], []
    for name, p in model.named_parameters():
        if "weight_generator" in name or "fclayer" in name:
            assert p.grad is not None, name
            got.append(p.grad.flatten().float().cpu())
            want.append(osd[name].grad.flatten())
    assert len(got) > 60
    for side in ("encoder.conv_in", "decoder.conv_out"):
        a = torch.cat([p.grad.flatten().float().cpu() for n, p in model.named_parameters() if n.startswith(side)])
        b = torch.cat([osd[n].grad.flatten() for n, p in model.named_parameters() if n.startswith(side)])
        print(f"factorized hypernet gradients {side}: rel-L2 {_rel(a, b):.3e}")
        assert _rel(a, b) < 6e-2, (side, _rel(a, b))


@pytest.mark.parametrize("modality", ["S2RGB", "S1RTC", "S2L2A"])
def test_wavelength_conditioner_and_adain_affine(cuda, modality):
    """eovae_wavelength_style_forward/backward and eovae_adain_affine_forward/backward (fp32) vs autograd over the oracle:
    style vector, modulated GroupNorm affine, and the gradients of the conditioner MLP / emb_proj / norm2 affine."""
    from eo_vae import autograd as tape
    from eo_vae.models.model import WavelengthConditioner
    from eo_vae.models.modules.layers import ResnetBlock
    from oracle import eovae_oracle as O
    from oracle.weights import WAVELENGTHS
    torch.manual_seed(11)
    cond = WavelengthConditioner(512).to(cuda)
    blk = ResnetBlock(64, 128, cond_dim=512).to(cuda)
    with torch.no_grad():
        blk.emb_proj.weight.normal_(0, 0.05)
        blk.emb_proj.bias.add_(0.05 * torch.randn_like(blk.emb_proj.bias))
        blk.norm2.weight.add_(0.1 * torch.randn_like(blk.norm2.weight))
        blk.norm2.bias.add_(0.1 * torch.randn_like(blk.norm2.bias))
    wvs = torch.tensor(WAVELENGTHS[modality], dtype=torch.float32)
    gsel = torch.randn(128, generator=torch.Generator().manual_seed(1))
    bsel = torch.randn(128, generator=torch.Generator().manual_seed(2))
    with torch.enable_grad():
        emb = cond(wvs.to(cuda), 3)
        assert emb.shape == (3, 512)
        g2, b2 = blk._norm2_affine(emb)
        ((g2 * gsel.to(cuda)).sum() + (b2 * bsel.to(cuda)).sum()).backward()
    sd = {"c." + k: v.detach().cpu().clone().requires_grad_(True) for k, v in cond.state_dict().items()}
    sd.update({"b." + k: v.detach().cpu().clone().requires_grad_(True) for k, v in blk.state_dict().items()})
    style = O.wavelength_style(sd, "c", wvs)
    assert _rel(emb[0].detach().cpu(), style[0].detach()) < 1e-4
    scale, shift = (style @ sd["b.emb_proj.weight"].t() + sd["b.emb_proj.bias"]).reshape(-1).chunk(2)
    g_ref = sd["b.norm2.weight"] * scale
    b_ref = sd["b.norm2.bias"] * scale + shift
    assert _rel(g2.detach().cpu(), g_ref.detach()) < 1e-4 and _rel(b2.detach().cpu(), b_ref.detach()) < 1e-4
    ((g_ref * gsel).sum() + (b_ref * bsel).sum()).backward()
    for name, p in list(cond.named_parameters()) + [(n, q) for n, q in blk.named_parameters() if "emb_proj" in n or "norm2" in n]:
        key = ("c." if name.startswith("mlp") else "b.") + name
        assert p.grad is not None, name
        assert _rel(p.grad.cpu(), sd[key].grad) < 1e-3, (name, _rel(p.grad.cpu(), sd[key].grad))


def test_adain_model_forward_and_gradients(cuda):
    """Tiny model with use_adain=True: latents / reconstruction (bf16 path) vs the oracle, and the whole-model parameter
    gradient (conditioner, emb_proj included) in train mode."""
    import __graft_entry__ as g
    from oracle import eovae_oracle as O
    from oracle.weights import TINY_ADAIN_CONFIG, WAVELENGTHS, make_state_dict, synthetic_patches
    cfg = TINY_ADAIN_CONFIG
    sd = make_state_dict(cfg, 4)
    model = g._model(cfg, sd, cuda)
    assert model.encoder.use_adain and model.decoder.use_adain
    wvs = torch.tensor(WAVELENGTHS["S2L2A"])
    x = synthetic_patches(2, 12, cfg["resolution"], seed=22)
    with torch.no_grad():
        z = model.encode_spatial_normalized(x.to(cuda), wvs.to(cuda))
        r = model.reconstruct(x.to(cuda), wvs.to(cuda))
    z_ref = O.encode_spatial_normalized(sd, x, wvs, cfg["hyper_heads"])
    r_ref = O.reconstruct(sd, x, wvs, cfg["hyper_heads"])
    # the same model WITHOUT the modulation must differ visibly (the AdaIN terms are exercised, not identity)
    plain = {k: v for k, v in sd.items() if "conditioner" not in k and "emb_proj" not in k}
    assert _rel(O.reconstruct(plain, x, wvs, cfg["hyper_heads"]), r_ref) > 0.1
    print(f"adain tiny: latent {_rel(z.cpu(), z_ref):.3e} recon {_rel(r.cpu(), r_ref):.3e}")
    assert _rel(z.cpu(), z_ref) < 1.5e-2 and _rel(r.cpu(), r_ref) < 4e-2
    model.train()
    recon, _ = model(x.to(cuda), wvs.to(cuda), sample_posterior=False)
    torch.sqrt((recon - x.to(cuda)) ** 2 + 1e-6).mean().backward()
    osd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v.clone()) for k, v in sd.items()}
    rr, _ = O.forward(osd, x, wvs, None, True, cfg["hyper_heads"])
    O.charbonnier_loss(rr, x).backward()
    for sel in ("", "conditioner", "emb_proj"):
        a = torch.cat([p.grad.flatten().float().cpu() for n, p in model.named_parameters() if sel in n])
        b = torch.cat([osd[n].grad.flatten() for n, p in model.named_parameters() if sel in n])
        print(f"adain gradients [{sel or 'all'}]: rel-L2 {_rel(a, b):.3e}")
        assert _rel(a, b) < 6e-2, (sel, _rel(a, b))


@pytest.mark.parametrize("shape", [(2, 12, 64, 64), (3, 2, 37, 51), (16, 12, 256, 256), (1, 13, 16, 16)])
def test_spectral_and_spatial_losses(cuda, shape):
    """eovae_sam_loss / eovae_grad_diff_loss forward + backward and the EOConsistencyLoss branches that use them
    (consistency_loss.py:186-210, 241-269, 426-440) vs the oracle (pinned against the reference classes on CPU)."""
    from eo_vae.models.modules.consistency_loss import EOConsistencyLoss, GradientDifferenceLoss, SAMLoss
    from oracle import eovae_oracle as O
    g = torch.Generator().manual_seed(9)
    x = torch.randn(shape, generator=g)
    r0 = x + 0.3 * torch.randn(shape, generator=g)
    r0[0, :, 0, 0] = 0.0
    r0[0, 0, 1, 1:4] = x[0, 0, 1, 1:4]        # exact ties: sign(0) = 0 in the gradient-difference adjoint
    for mod, fn in ((SAMLoss(), O.sam_loss), (GradientDifferenceLoss(), O.grad_diff_loss)):
        a = r0.clone().to(cuda).requires_grad_(True)
        b = r0.clone().requires_grad_(True)
        la = mod(a, x.to(cuda))
        lb = fn(b, x)
        (3.0 * la).backward(); (3.0 * lb).backward()
        assert abs(float(la) - float(lb)) < 1e-5 * abs(float(lb)) + 1e-7, (type(mod).__name__, float(la), float(lb))
        assert _rel(a.grad.cpu(), b.grad) < 1e-4, (type(mod).__name__, _rel(a.grad.cpu(), b.grad))
        with torch.no_grad():
            assert abs(float(mod(r0.to(cuda), x.to(cuda))) - float(lb)) < 1e-5 * abs(float(lb)) + 1e-7
    loss = EOConsistencyLoss(pixel_weight=1.0, rec_loss_type="char", spectral_weight=0.5, spatial_weight=2.0,
                             spatial_start_step=10).to(cuda)
    for step in (0, 10):
        a = r0.clone().to(cuda).requires_grad_(True)
        b = r0.clone().requires_grad_(True)
        total, logs = loss(inputs=x.to(cuda), wvs=None, reconstructions=a, global_step=step)
        want, _, _ = O.consistency_loss(x, b, "char", 1.0, 0.0, step, 0, spectral_weight=0.5, spatial_weight=2.0,
                                        spatial_start_step=10)
        total.backward(); want.backward()
        assert abs(float(total) - float(want)) < 1e-5 * abs(float(want))
        assert _rel(a.grad.cpu(), b.grad) < 1e-4
        assert ("train/loss_spatial" in logs) == (step >= 10) and "train/loss_spectral" in logs


@pytest.mark.parametrize("shape,pf,alpha", [((2, 3, 64, 64), 2, 1.0), ((1, 12, 24, 36), 1, 1.0), ((2, 2, 28, 44), 2, 0.5),
                                            ((4, 12, 256, 256), 2, 1.0)])
def test_focal_frequency_loss(cuda, shape, pf, alpha):
    """eovae_focal_freq_loss forward + backward (DFT-matrix products, power-of-two and other patch sizes) and the freq
    branch of EOConsistencyLoss with its warm-up vs the oracle (pinned against the reference's ffl.py on CPU)."""
    from eo_vae.models.modules.consistency_loss import EOConsistencyLoss, FocalFrequencyLoss
    from oracle import eovae_oracle as O
    g = torch.Generator().manual_seed(10)
    x = torch.randn(shape, generator=g)
    r0 = x + 0.3 * torch.randn(shape, generator=g)
    mod = FocalFrequencyLoss(alpha=alpha, patch_factor=pf, batch_matrix=True, log_matrix=True)
    a = r0.clone().to(cuda).requires_grad_(True)
    b = r0.clone().requires_grad_(True)
    la, lb = mod(a, x.to(cuda)), O.focal_freq_loss(b, x, pf, alpha)
    (2.0 * la).backward(); (2.0 * lb).backward()
    print(f"ffl {shape} pf {pf}: value {float(la.detach()):.6e} vs {float(lb.detach()):.6e}, grad rel {_rel(a.grad.cpu(), b.grad):.2e}")
    assert abs(float(la.detach()) - float(lb.detach())) < 2e-4 * abs(float(lb.detach()))
    assert _rel(a.grad.cpu(), b.grad) < 1e-3
    loss = EOConsistencyLoss(pixel_weight=1.0, rec_loss_type="l1", freq_weight=3.0, freq_start_step=100, patch_factor=pf,
                             ffl_alpha=alpha).to(cuda)
    for step in (0, 600):
        with torch.no_grad():
            total, logs = loss(inputs=x.to(cuda), wvs=None, reconstructions=r0.to(cuda), global_step=step)
        want, _, _ = O.consistency_loss(x, r0, "l1", 1.0, 0.0, step, 0, freq_weight=3.0, freq_start_step=100, patch_factor=pf,
                                        ffl_alpha=alpha)
        assert abs(float(total) - float(want)) < 2e-4 * abs(float(want))
        assert ("train/loss_freq_raw" in logs) == (step >= 100)


@pytest.mark.parametrize("shape,size,k", [((2, 8, 16, 16), (8, 8), 1), ((2, 8, 16, 16), (12, 12), 3), ((3, 4, 16, 24), (6, 18), 2),
                                          ((1, 32, 32, 32), (16, 24), 1), ((2, 8, 16, 16), None, 1), ((2, 8, 16, 16), (6, 6), 0)])
def test_latent_resize_rot_and_area_target(cuda, shape, size, k):
    """eovae_latent_resize_rot (+ backward) and eovae_area_resize_rot vs the torch fp32 ops the reference calls
    (F.interpolate bilinear / area, torch.rot90 with dims=[-1, -2]; new_autoencoder.py:460-464, 519-531, 614-624)."""
    import torch.nn.functional as F
    from eo_vae import autograd as tape
    from eo_vae import ops
    g = torch.Generator().manual_seed(4)
    z = torch.randn(shape, generator=g)
    a = z.clone().to(cuda).requires_grad_(True)
    b = z.clone().requires_grad_(True)
    out = tape.LatentResizeRotFn.apply(a, size, k)
    ref = b if size is None else F.interpolate(b, size=size, mode="bilinear", align_corners=False)
    ref = torch.rot90(ref, k=k, dims=[-1, -2])
    assert out.shape == ref.shape
    assert float((out.detach().cpu() - ref.detach()).abs().max()) < 1e-5
    sel = torch.randn(ref.shape, generator=g)
    (out * sel.to(cuda)).sum().backward(); (ref * sel).sum().backward()
    assert float((a.grad.cpu() - b.grad).abs().max()) < 1e-4
    if size is not None:
        x = torch.randn((shape[0], 3, shape[2] * 8, shape[3] * 8), generator=g)
        big = (size[0] * 8, size[1] * 8)
        t = ops.area_resize_rot(x.to(cuda), big, k)
        t_ref = torch.rot90(F.interpolate(x, size=big, mode="area"), k=k, dims=[-1, -2])
        assert t.shape == t_ref.shape and float((t.cpu() - t_ref).abs().max()) < 1e-5


@pytest.mark.parametrize("scale,angle", [(0.5, 1), (0.75, 3), (0.5, None)])
def test_eq_vae_forward_and_training_step(cuda, scale, angle):
    """EQ-VAE modes end to end on the tiny model: train-mode forward with scale / angle vs the oracle (same noise), and a
    training_step with p_prior = 1 (latent equivariance mode) runs on the kernels and moves the parameters."""
    import __graft_entry__ as g
    from eo_vae.models.modules.consistency_loss import EOConsistencyLoss
    from oracle import eovae_oracle as O
    from oracle.weights import TINY_CONFIG, WAVELENGTHS, make_state_dict, synthetic_patches
    cfg = TINY_CONFIG
    sd = make_state_dict(cfg, 6)
    model = g._model(cfg, sd, cuda)
    model.train()
    wvs = torch.tensor(WAVELENGTHS["S2RGB"])
    x = synthetic_patches(2, 3, cfg["resolution"], seed=83)
    hl = cfg["resolution"] // 2 ** (len(cfg["ch_mult"]) - 1)
    torch.manual_seed(7)
    eps = torch.randn((2, cfg["z_channels"], hl, hl))
    torch.manual_seed(7)
    with torch.no_grad():
        recon, _ = model(x.to(cuda), wvs.to(cuda), scale=scale, angle=angle)
    recon_ref, _ = O.forward(sd, x, wvs, eps=eps, train=True, heads=cfg["hyper_heads"], scale=scale, angle=angle)
    assert recon.shape == recon_ref.shape
    err = _rel(recon.cpu(), recon_ref)
    print(f"eq-vae forward scale {scale} angle {angle}: recon rel-L2 {err:.3e}")
    assert err < 4e-2
    if angle is not None:
        import random
        from eo_vae import _lightning  # noqa: F401
        model2 = g._model(cfg, make_state_dict(cfg, 6), cuda)
        model2.train()
        model2.p_prior = 1.0
        model2.loss_fn = EOConsistencyLoss(pixel_weight=1.0, rec_loss_type="char").to(cuda)
        before = {k: v.detach().clone() for k, v in model2.named_parameters()}
        random.seed(3)
        loss = model2.training_step({model2.image_key: x.to(cuda), "wvs": wvs.to(cuda)}, 0)
        assert float(loss) == float(loss)
        moved = sum(not torch.equal(v.detach(), before[k]) for k, v in model2.named_parameters())
        assert moved > 100


@pytest.mark.parametrize("cout,cin,k", [(128, 128, 3), (512, 256, 3), (64, 512, 3), (12, 128, 3), (128, 12, 3), (24, 40, 3),
                                        (512, 512, 1), (64, 64, 1), (32, 8, 1), (256, 1024, 3)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_weight_pack_layouts(cuda, cout, cin, k, dtype):
    """eovae_pack_conv_weight / eovae_pack_conv_weight_dgrad (coalesced shared-memory kernels): exact bits of the K-major
    forward operand [round_up(cout,16)][taps][k_per_tap] and of the data-gradient operand (taps flipped, channels swapped),
    zero padding included."""
    from eo_vae import ops
    g = torch.Generator().manual_seed(cout * 131 + cin)
    w = torch.randn((cout, cin, k, k), generator=g)
    taps = k * k
    fwd = ops.pack_conv_weight(w.to(cuda), dtype).cpu()
    rows, kpt = fwd.shape[0], fwd.shape[2]
    want = torch.zeros((rows, taps, kpt), dtype=dtype)
    want[:cout, :, :cin] = w.reshape(cout, cin, taps).permute(0, 2, 1).to(dtype)
    assert fwd.shape == want.shape and torch.equal(fwd, want)
    dg = ops.pack_conv_weight_dgrad(w.to(cuda), dtype).cpu()
    rows, kpt = dg.shape[0], dg.shape[2]
    want = torch.zeros((rows, taps, kpt), dtype=dtype)
    want[:cin, :, :cout] = w.reshape(cout, cin, taps).flip(2).permute(1, 2, 0).to(dtype)
    assert dg.shape == want.shape and torch.equal(dg, want)


# ---------------------------------------------------------------------------------------------------------------------
# Upsample in sub-pixel form (layers.py:40-50): forward, data gradient and weight gradient against torch on the SAME rounded
# operands (nearest x2 + 3x3 conv, fp32), plus GroupNorm statistics of the output from the phase epilogue.
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("n,c,cout,h,w", [(2, 64, 64, 32, 32), (1, 128, 64, 64, 64), (2, 256, 256, 16, 32), (3, 64, 96, 8, 16)])
def test_upsample_subpixel_forward_and_gradients(cuda, dtype, n, c, cout, h, w):
    from eo_vae import ops
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(n * 1000 + c + h)
    x = torch.randn((n, h, w, c), generator=g).to(cuda).to(dtype).permute(0, 3, 1, 2)
    wgt = (torch.randn((cout, c, 3, 3), generator=g) * 0.05).to(cuda)
    bias = torch.randn((cout,), generator=g).to(cuda)
    assert ops.up2x_ok(x, cout)
    out = ops.conv2d_up2x(x, ops.pack_conv_weight_up2x(wgt, dtype), bias, cout, gn_groups=32)
    assert out.shape == (n, cout, 2 * h, 2 * w)
    xr = x.float().detach().requires_grad_(True)
    wr = wgt.clone().requires_grad_(True)
    ref = F.conv2d(F.interpolate(xr, scale_factor=2.0, mode="nearest"), wr, bias, padding=1)
    # the kernel rounds the FOLDED taps (sums of up to four fp32 weights) once; the reference rounds nothing: fp16/bf16 operand error
    tol = 2e-3 if dtype == torch.float16 else 1.2e-2
    assert _rel(out.float(), ref) < tol, _rel(out.float(), ref)
    if 32 % (cout // 32) == 0:   # group widths the statistics epilogue supports (the network's: 4, 8, 16 channels)
        st = out._gn_stats[0]
        grp = out.float().permute(0, 2, 3, 1).reshape(n, -1, 32, cout // 32).permute(0, 2, 1, 3).reshape(n, 32, -1)
        assert torch.allclose(st[..., 0], grp.mean(-1), atol=2e-3)
        assert torch.allclose(st[..., 1], 1.0 / torch.sqrt(grp.var(-1, unbiased=False) + 1e-6), rtol=2e-3)
    else:
        assert not hasattr(out, "_gn_stats")
    dy = torch.randn((n, 2 * h, 2 * w, cout), generator=g).to(cuda).to(dtype).permute(0, 3, 1, 2)
    ref.backward(dy.float())
    dx = ops.conv2d_up2x_dgrad(dy, wgt)
    assert dx.shape == x.shape and _rel(dx.float(), xr.grad) < tol * 1.5, _rel(dx.float(), xr.grad)
    assert ops.up2x_wgrad_ok(x, dy)
    dw = ops.conv2d_up2x_wgrad(x, dy)
    assert _rel(dw, wr.grad) < 2e-3, _rel(dw, wr.grad)       # fp32 accumulate over exactly representable operands


def test_upsample_module_subpixel_equals_materialised(cuda):
    """The Upsample module through both paths (sub-pixel vs nearest-upsample + 3x3) and its taped gradients."""
    from eo_vae import ops
    from eo_vae.models.modules.layers import Upsample
    torch.manual_seed(5)
    up = Upsample(64).to(cuda)
    x = torch.randn((2, 32, 32, 64), device=cuda).permute(0, 3, 1, 2)
    with torch.no_grad():
        a = up(x).float()
        ops.USE_UP2X = False
        try:
            b = up(x).float()
        finally:
            ops.USE_UP2X = True
    assert _rel(a, b) < 2e-3
    grads = []
    for flag in (True, False):
        ops.USE_UP2X = flag
        try:
            xr = x.clone().requires_grad_(True)
            up.zero_grad()
            y = up(xr)
            y.float().square().mean().backward()
            grads.append((xr.grad.float().clone(), up.conv.weight.grad.clone(), up.conv.bias.grad.clone()))
        finally:
            ops.USE_UP2X = True
    for u, v in zip(*grads):
        assert _rel(u, v) < 1.5e-2, _rel(u, v)       # bf16 training operands: folded vs unfolded rounding


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("n,c,cout,h,w", [(2, 64, 64, 64, 64), (1, 128, 128, 32, 64), (2, 64, 128, 16, 32)])
def test_downsample_subpixel_data_gradient(cuda, dtype, n, c, cout, h, w):
    """Data gradient of the Downsample conv (pad (0,1,0,1) + 3x3 stride 2) as four 2x2 phase convolutions vs torch autograd."""
    from eo_vae import ops
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(n * 100 + c + h)
    wgt = (torch.randn((cout, c, 3, 3), generator=g) * 0.05).to(cuda)
    xr = torch.randn((n, c, h, w), generator=g).to(cuda).requires_grad_(True)
    y = F.conv2d(F.pad(xr, (0, 1, 0, 1)), wgt, None, stride=2)
    dy = torch.randn((n, h // 2, w // 2, cout), generator=g).to(cuda).to(dtype).permute(0, 3, 1, 2)
    y.backward(dy.float())
    assert ops.s2_dgrad_ok(dy, c, (h, w))
    dx = ops.conv2d_dgrad(dy, wgt, ops.CONV_3X3_S2, in_hw=(h, w))
    assert dx.shape == (n, c, h, w)
    tol = 2e-3 if dtype == torch.float16 else 1.2e-2
    assert _rel(dx.float(), xr.grad) < tol, _rel(dx.float(), xr.grad)
    ops.USE_UP2X = False
    try:
        dx_old = ops.conv2d_dgrad(dy, wgt, ops.CONV_3X3_S2, in_hw=(h, w))
    finally:
        ops.USE_UP2X = True
    assert _rel(dx.float(), dx_old.float()) < (2e-2 if dtype == torch.bfloat16 else 3e-3)
    # weight gradient: x read on its parity sub-lattices
    x16 = xr.detach().to(dtype).contiguous(memory_format=torch.channels_last)
    xr2 = x16.float().requires_grad_(True)
    wr = wgt.clone().requires_grad_(True)
    F.conv2d(F.pad(xr2, (0, 1, 0, 1)), wr, None, stride=2).backward(dy.float())
    assert ops.s2_wgrad_ok(x16, dy)
    dw = ops.conv2d_s2_wgrad(x16, dy)
    assert _rel(dw, wr.grad) < 2e-3, _rel(dw, wr.grad)
