"""Training path: the tape entries of eo_vae/autograd.py (forward + hand-written backward kernels) against torch autograd
over the fp32 oracle, first per reference module, then for the whole EOFluxVAE forward + Charbonnier loss at the tiny
configuration, then one ``training_step`` (manual optimisation, reference new_autoencoder.py:587-690)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

# bf16 operands and bf16 inter-layer gradients against an fp32 reference: relative L2 per tensor
TOL_BLOCK = 3e-2


def _rel(a, b):
    return float((a.detach().float() - b.detach().float()).norm() / (b.detach().float().norm() + 1e-30))


def _randomise(mod, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in mod.named_parameters():
            if name.endswith("norm.weight") or "norm1.weight" in name or "norm2.weight" in name:
                p.copy_(1.0 + 0.2 * torch.randn(p.shape, generator=g))
            elif p.dim() == 1:
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
            else:
                fan_in = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g) / fan_in ** 0.5)


def _check_module(mod, oracle_fn, x16, cuda, tol=TOL_BLOCK):
    """grads of sum(out * r) wrt the input and every parameter, kernels vs autograd over the oracle in fp32."""
    g = torch.Generator().manual_seed(99)
    x = x16.clone().requires_grad_(True)
    out = mod(x)
    r = torch.randn(out.shape, generator=g).to(cuda)
    (out.float() * r).sum().backward()
    sd = {"m." + k: v.detach().clone().float().requires_grad_(True) for k, v in mod.state_dict().items()}
    xr = x16.float().contiguous().requires_grad_(True)
    ref = oracle_fn(sd, "m", xr)
    assert _rel(out, ref) < 2e-2
    (ref * r).sum().backward()
    pairs = {"x": (x.grad, xr.grad)}
    for name, p in mod.named_parameters():
        assert p.grad is not None, name
        pairs[name] = (p.grad, sd["m." + name].grad)
    # a gradient that is analytically zero (the key bias: softmax is shift invariant) is compared on the scale of the
    # largest parameter gradient instead of its own (rounding-noise) norm
    bad = {}
    for name, (a, b) in pairs.items():
        err = float((a.float() - b.float()).norm())
        scale = float(b.float().norm())
        if name == "k.bias":
            scale = float(pairs["k.weight"][1].float().norm())
        if not err < tol * scale:
            bad[name] = (err, scale)
    assert not bad, f"gradient mismatch (abs err, ref norm): {bad}"


def _act(n, c, h, w, dev, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn((n, c, h, w), generator=g).to(dev).to(dtype=torch.bfloat16, memory_format=torch.channels_last)


@pytest.mark.parametrize("cin,cout,hw", [(64, 64, 16), (64, 128, 16), (32, 64, 32), (128, 128, 8)])
def test_resnet_block_gradients(cuda, cin, cout, hw):
    from eo_vae.models.modules.layers import ResnetBlock
    from oracle import eovae_oracle as O
    mod = ResnetBlock(cin, cout).to(cuda)
    _randomise(mod, 1)
    _check_module(mod, O.resnet_block, _act(2, cin, hw, hw, cuda, 3), cuda)


@pytest.mark.parametrize("c,hw", [(64, 16), (128, 8)])
def test_attn_block_gradients(cuda, c, hw):
    from eo_vae.models.modules.layers import AttnBlock
    from oracle import eovae_oracle as O
    mod = AttnBlock(c).to(cuda)
    _randomise(mod, 2)
    _check_module(mod, O.attn_block, _act(2, c, hw, hw, cuda, 4), cuda)


@pytest.mark.parametrize("kind", ["up", "down"])
def test_resample_gradients(cuda, kind):
    from eo_vae.models.modules.layers import Downsample, Upsample
    from oracle import eovae_oracle as O
    mod = (Upsample(64) if kind == "up" else Downsample(64)).to(cuda)
    _randomise(mod, 5)
    _check_module(mod, O.upsample if kind == "up" else O.downsample, _act(2, 64, 16, 16, cuda, 6), cuda)


@pytest.mark.parametrize("generator", ["transformer", "factorized"])
@pytest.mark.parametrize("decoder", [False, True])
@pytest.mark.parametrize("modality", ["S2L2A", "S1RTC"])
def test_hypernet_backward(cuda, decoder, modality, generator):
    """eovae_hypernet_backward / eovae_hypernet_factorized_backward (full-size generator: d 256, 4 layers, 4 heads, ff 2048
    / 1024 + rank-576 head) vs autograd over the oracle, fp32 on both sides."""
    from eo_vae.models.modules.dynamic_conv import DynamicConv, DynamicConv_decoder
    from oracle import eovae_oracle as O
    from oracle.weights import WAVELENGTHS
    torch.manual_seed(7)
    cls = DynamicConv_decoder if decoder else DynamicConv
    mod = cls(wv_planes=256, inter_dim=128, kernel_size=3, stride=1, padding=1, embed_dim=128, num_layers=4, num_heads=4,
              generator_type=generator, rank_ratio=2).to(cuda).eval()
    wvs = torch.tensor(WAVELENGTHS[modality], dtype=torch.float32, device=cuda)
    c, e = wvs.numel(), 128
    g = torch.Generator().manual_seed(3)
    shape = (c, e, 3, 3) if decoder else (e, c, 3, 3)
    dw = torch.randn(shape, generator=g).to(cuda)
    db = torch.randn((c if decoder else e,), generator=g).to(cuda)
    bias_scale = 0.01 if decoder else 0.1
    if decoder:
        dw_in = dw
    else:  # the encoder layer's weight gradient comes with the band dimension padded to 16
        dw_in = torch.zeros((e, 16, 3, 3), device=cuda)
        dw_in[:, :c] = dw
    grads = mod._hyper_backward(wvs, dw_in, db, bias_scale)
    sd = {"m." + k: v.detach().cpu().clone().requires_grad_(True) for k, v in mod.state_dict().items()}
    w_ref, b_ref = O.hypernet(sd, "m", wvs.cpu(), decoder, heads=4)
    ((w_ref * dw.cpu()).sum() + (b_ref * db.cpu()).sum()).backward()
    names = [n for n, _ in mod.named_parameters()]
    by_ptr = {p.data_ptr(): n for n, p in mod.named_parameters()}
    plist = mod.weight_generator.parameter_list(mod.fclayer)
    assert len(plist) == len(names)
    bad = {}
    for p, got in zip(plist, grads):
        name = by_ptr[p.data_ptr()]
        want = sd["m." + name].grad
        err = float((got.cpu() - want).norm())
        if not err < 2e-3 * float(want.norm()) + 1e-6:
            bad[name] = (err, float(want.norm()))
    assert not bad, bad


@pytest.mark.parametrize("shape", [(2, 3, 176, 192), (1, 12, 256, 256)])
@pytest.mark.parametrize("rec", ["char", "l1"])
def test_loss_gradients(cuda, shape, rec):
    """EOConsistencyLoss (pixel + MS-SSIM, the shipped training configuration) value and d/d(reconstruction)."""
    from eo_vae.models.modules.consistency_loss import EOConsistencyLoss
    from oracle import eovae_oracle as O
    g = torch.Generator().manual_seed(21)
    target = torch.randn(shape, generator=g).clamp_(-2, 6)
    pred0 = (target + 0.3 * torch.randn(shape, generator=g))
    loss_fn = EOConsistencyLoss(pixel_weight=1.0, rec_loss_type=rec, msssim_weight=1.0, msssim_start_step=0).to(cuda)
    pred = pred0.to(cuda).requires_grad_(True)
    loss, logs = loss_fn(inputs=target.to(cuda), wvs=None, reconstructions=pred, global_step=10)
    loss.backward()
    pr = pred0.clone().requires_grad_(True)
    ref = O.consistency_loss(target, pr, rec_loss_type=rec, pixel_weight=1.0, msssim_weight=1.0)
    ref = ref[0] if isinstance(ref, tuple) else ref
    ref.backward()
    assert abs(float(loss) - float(ref)) < 1e-3 * abs(float(ref))
    assert _rel(pred.grad.cpu(), pr.grad) < 1e-3


def _tiny(cuda, seed=3):
    import __graft_entry__ as ge
    from oracle.weights import TINY_CONFIG, make_state_dict
    sd = make_state_dict(TINY_CONFIG, seed=seed)
    model = ge._model(TINY_CONFIG, sd, cuda)
    return model, sd, TINY_CONFIG


def test_tiny_model_gradients(cuda):
    """d Charbonnier(recon, x) / d(body parameters): sampled posterior, BatchNorm in train mode."""
    from eo_vae.models.modules.consistency_loss import EOConsistencyLoss
    from oracle import eovae_oracle as O
    from oracle.weights import WAVELENGTHS, synthetic_patches
    model, sd, cfg = _tiny(cuda)
    model.train()
    wvs = torch.tensor(WAVELENGTHS["S2L2A"], dtype=torch.float32)
    x = synthetic_patches(2, 12, cfg["resolution"], seed=11)
    loss_fn = EOConsistencyLoss(pixel_weight=1.0, rec_loss_type="char").to(cuda)
    torch.manual_seed(1234)
    recon, post = model(x.to(cuda), wvs.to(cuda))
    loss, _ = loss_fn(inputs=x.to(cuda), wvs=wvs.to(cuda), reconstructions=recon, global_step=0)
    loss.backward()

    ref_sd = {k: (v.clone().float().requires_grad_(True) if v.is_floating_point() else v.clone()) for k, v in sd.items()}
    torch.manual_seed(1234)
    zc = cfg["z_channels"]
    hl = cfg["resolution"] // 2 ** (len(cfg["ch_mult"]) - 1)
    eps = torch.randn((2, zc, hl, hl))  # the draw DiagonalGaussianDistribution.sample makes on the CPU generator
    recon_ref, moments_ref = O.forward(ref_sd, x, wvs, eps, train=True, heads=cfg["hyper_heads"])
    loss_ref = O.charbonnier_loss(recon_ref, x)
    loss_ref.backward()
    assert abs(float(loss) - float(loss_ref)) < 1e-2 * abs(float(loss_ref)) + 1e-3
    # the fused train-mode latent kernel also updated the BatchNorm running statistics like nn.BatchNorm2d does
    with torch.no_grad():
        _, new = O.bn_train(ref_sd, O.pixel_unshuffle2(O.posterior_sample(moments_ref, eps)))
    assert torch.allclose(model.bn.running_mean.cpu(), new["bn.running_mean"], atol=2e-3)
    assert torch.allclose(model.bn.running_var.cpu(), new["bn.running_var"], rtol=2e-2, atol=1e-3)
    assert int(model.bn.num_batches_tracked) == int(sd["bn.num_batches_tracked"]) + 1

    got, want, worst = [], [], {}
    for name, p in model.named_parameters():
        assert p.grad is not None, name
        gr = ref_sd[name].grad
        got.append(p.grad.flatten().cpu())
        want.append(gr.flatten())
        worst[name] = _rel(p.grad.cpu(), gr)
    total = _rel(torch.cat(got), torch.cat(want))
    wn = float(torch.cat(want).norm())
    contrib = sorted(((round(float((p.grad.float().cpu() - ref_sd[n].grad).norm()) / wn, 5), round(worst[n], 4), n)
                      for n, p in model.named_parameters()), reverse=True)[:5]
    print("largest contributions (share of total error, own rel err, name):", contrib)
    top = sorted(worst.items(), key=lambda kv: -kv[1])[:5]
    print(f"tiny model gradient: total rel-L2 {total:.3e}; worst tensors {top}")
    assert total < 5e-2, (total, top)


@pytest.mark.parametrize("modality", ["S1RTC", "S2RGB", "S2L1C"])
def test_mixed_modality_training_steps(cuda, modality):
    """BASELINE configs[3]: the band count changes from batch to batch (2 / 3 / 13 bands), the parameters do not; every
    trainable tensor - the wavelength hypernetworks included - receives a finite gradient."""
    from eo_vae.models.modules.consistency_loss import EOConsistencyLoss
    from oracle.weights import WAVELENGTHS, synthetic_patches
    model, sd, cfg = _tiny(cuda)
    model.train()
    model.loss_fn = EOConsistencyLoss(pixel_weight=1.0, rec_loss_type="l1").to(cuda)
    model.clip_grad = 1.0
    losses = []
    for step, mod in enumerate(["S2L2A", modality, "S2L2A", modality]):
        wvs = torch.tensor(WAVELENGTHS[mod], dtype=torch.float32).to(cuda)
        x = synthetic_patches(2, wvs.numel(), cfg["resolution"], seed=30 + step).to(cuda)
        losses.append(float(model.training_step({model.image_key: x, "wvs": wvs}, step)))
    assert all(l == l and abs(l) < 1e6 for l in losses), losses
    for name, p in model.named_parameters():
        assert p.grad is not None and bool(torch.isfinite(p.grad).all()), name


def test_training_step_odd_image_size(cuda):
    """56x56 patches (activations 56 / 28 / 14 pixels wide: none tiles into 64-pixel TMA boxes) - the weight gradients
    take the row-padded transposed-operand kernel; gradients still match autograd over the oracle."""
    from eo_vae.models.modules.consistency_loss import EOConsistencyLoss
    from oracle import eovae_oracle as O
    from oracle.weights import WAVELENGTHS, synthetic_patches
    model, sd, cfg = _tiny(cuda)
    model.train()
    wvs = torch.tensor(WAVELENGTHS["S2RGB"], dtype=torch.float32)
    x = synthetic_patches(2, 3, 56, seed=13)
    loss_fn = EOConsistencyLoss(pixel_weight=1.0, rec_loss_type="l1").to(cuda)
    torch.manual_seed(99)
    recon, _ = model(x.to(cuda), wvs.to(cuda))
    loss, _ = loss_fn(inputs=x.to(cuda), wvs=wvs.to(cuda), reconstructions=recon, global_step=0)
    loss.backward()
    ref_sd = {k: (v.clone().float().requires_grad_(True) if v.is_floating_point() else v.clone()) for k, v in sd.items()}
    torch.manual_seed(99)
    eps = torch.randn((2, cfg["z_channels"], 14, 14))
    recon_ref, _ = O.forward(ref_sd, x, wvs, eps, train=True, heads=cfg["hyper_heads"])
    O.l1_loss(recon_ref, x).backward()
    got = torch.cat([p.grad.flatten().cpu() for n, p in model.named_parameters()])
    want = torch.cat([ref_sd[n].grad.flatten() for n, p in model.named_parameters()])
    assert _rel(got, want) < 6e-2, _rel(got, want)


def test_graphed_train_step_matches_eager(cuda):
    """GraphedTrainStep (forward + loss + backward as one CUDA graph) produces the eager step's gradients."""
    from eo_vae.graphs import GraphedTrainStep
    from eo_vae.models.modules.consistency_loss import EOConsistencyLoss
    from oracle.weights import WAVELENGTHS, synthetic_patches
    wvs = torch.tensor(WAVELENGTHS["S2L2A"], dtype=torch.float32).to(cuda)
    grads = []
    for graphed in (False, True):
        model, sd, cfg = _tiny(cuda)
        model.train()
        model.loss_fn = EOConsistencyLoss(pixel_weight=1.0, rec_loss_type="char").to(cuda)
        model.base_lr = 0.0   # keep the parameters fixed so both runs differentiate the same function
        batch = {model.image_key: synthetic_patches(2, 12, cfg["resolution"], seed=5).to(cuda), "wvs": wvs}
        torch.manual_seed(77)
        if graphed:
            step = GraphedTrainStep(model, batch, warmup=2)
            with torch.no_grad():  # the warm-up / capture passes moved the BatchNorm running statistics: restore them
                model.bn.running_mean.copy_(sd["bn.running_mean"])
                model.bn.running_var.copy_(sd["bn.running_var"])
            torch.manual_seed(77)
            loss = step(batch)
        else:
            loss = model.training_step(batch, 0)
        grads.append((float(loss), torch.cat([p.grad.flatten().float() for p in model.parameters()]).clone()))
    assert abs(grads[0][0] - grads[1][0]) < 1e-6 * abs(grads[0][0]) + 1e-7
    assert torch.equal(grads[0][1], grads[1][1])


def test_graph_training_flag_mixed_modalities(cuda):
    """model.graph_training = True: training_step keeps one CUDA graph per (shape, band count) signature; a mixed
    S2L2A / S1RTC sequence reuses them and the loss on a fixed batch goes down."""
    from eo_vae.models.modules.consistency_loss import EOConsistencyLoss
    from oracle.weights import WAVELENGTHS, synthetic_patches
    model, sd, cfg = _tiny(cuda)
    model.train()
    model.loss_fn = EOConsistencyLoss(pixel_weight=1.0, rec_loss_type="char").to(cuda)
    model.base_lr = 2e-4
    model.clip_grad = 1.0
    model.graph_training = True
    batches = {}
    for mod in ("S2L2A", "S1RTC"):
        wvs = torch.tensor(WAVELENGTHS[mod], dtype=torch.float32).to(cuda)
        batches[mod] = {model.image_key: synthetic_patches(2, wvs.numel(), cfg["resolution"], seed=60).to(cuda), "wvs": wvs}
    torch.manual_seed(0)
    losses = {"S2L2A": [], "S1RTC": []}
    for step in range(8):
        mod = ("S2L2A", "S1RTC")[step % 2]
        losses[mod].append(float(model.training_step(batches[mod], step)))
    assert len(model._train_graphs) == 2
    for mod, ls in losses.items():
        assert all(l == l for l in ls) and ls[-1] < ls[0], (mod, ls)


def test_freeze_body_trains_only_the_dynamic_layers(cuda):
    """freeze_body=True (the reference default, new_autoencoder.py:284-293): only the wavelength hypernetworks receive
    gradients / move; the frozen Flux body still back-propagates the data gradient to the dynamic input layer."""
    from eo_vae.models.modules.consistency_loss import EOConsistencyLoss
    from oracle.weights import WAVELENGTHS, synthetic_patches
    model, sd, cfg = _tiny(cuda)
    model._freeze_body()
    model.train()
    model.loss_fn = EOConsistencyLoss(pixel_weight=1.0, rec_loss_type="l1").to(cuda)
    wvs = torch.tensor(WAVELENGTHS["S2L2A"], dtype=torch.float32).to(cuda)
    batch = {model.image_key: synthetic_patches(2, 12, cfg["resolution"], seed=5).to(cuda), "wvs": wvs}
    before = {k: v.detach().clone() for k, v in model.named_parameters()}
    loss = model.training_step(batch, 0)
    assert float(loss) == float(loss)
    for k, v in model.named_parameters():
        dyn = "encoder.conv_in" in k or "decoder.conv_out" in k
        assert v.requires_grad == dyn, k
        if dyn:
            assert v.grad is not None and bool(torch.isfinite(v.grad).all()), k
        else:
            assert torch.equal(v.detach(), before[k]), k
    moved = [k for k, v in model.named_parameters() if v.requires_grad and not torch.equal(v.detach(), before[k])]
    assert len(moved) > 10


def test_training_step_fp16_operands(cuda):
    """The training path with the reference trainer's operand type (precision 16-mixed -> fp16): gradients vs the oracle."""
    import eo_vae
    from eo_vae.models.modules.consistency_loss import EOConsistencyLoss
    from oracle import eovae_oracle as O
    from oracle.weights import WAVELENGTHS, synthetic_patches
    eo_vae.set_train_dtype(torch.float16)
    try:
        model, sd, cfg = _tiny(cuda)
        model.train()
        wvs = torch.tensor(WAVELENGTHS["S2L2A"], dtype=torch.float32)
        x = synthetic_patches(2, 12, cfg["resolution"], seed=11)
        loss_fn = EOConsistencyLoss(pixel_weight=1.0, rec_loss_type="char").to(cuda)
        torch.manual_seed(1234)
        recon, _ = model(x.to(cuda), wvs.to(cuda))
        loss, _ = loss_fn(inputs=x.to(cuda), wvs=wvs.to(cuda), reconstructions=recon, global_step=0)
        # fp16 gradients need the usual loss scaling (Lightning's 16-mixed plugin wraps manual_backward in a GradScaler):
        # unscaled, the ~1e-6 inter-layer gradients fall into fp16's subnormal range (measured 1.2e-2 instead of 3e-3)
        scale = 4096.0
        (loss * scale).backward()
        for p in model.parameters():
            p.grad.div_(scale)
    finally:
        eo_vae.set_train_dtype(torch.bfloat16)
    ref_sd = {k: (v.clone().float().requires_grad_(True) if v.is_floating_point() else v.clone()) for k, v in sd.items()}
    torch.manual_seed(1234)
    hl = cfg["resolution"] // 2 ** (len(cfg["ch_mult"]) - 1)
    eps = torch.randn((2, cfg["z_channels"], hl, hl))
    recon_ref, _ = O.forward(ref_sd, x, wvs, eps, train=True, heads=cfg["hyper_heads"])
    O.charbonnier_loss(recon_ref, x).backward()
    got = torch.cat([p.grad.flatten().float().cpu() for _, p in model.named_parameters()])
    want = torch.cat([ref_sd[n].grad.flatten() for n, _ in model.named_parameters()])
    err = _rel(got, want)
    contrib = sorted(((round(float((p.grad.float().cpu() - ref_sd[n].grad).norm()) / float(want.norm()), 5),
                       round(_rel(p.grad.cpu(), ref_sd[n].grad), 4), round(float(ref_sd[n].grad.norm()) / float(want.norm()), 3), n)
                      for n, p in model.named_parameters()), reverse=True)[:6]
    print(f"fp16-operand training gradient rel-L2 {err:.3e}; largest contributions {contrib}")
    assert err < 8e-3, err  # measured 5.6e-3


def test_training_step_reduces_loss(cuda):
    """A few manual-optimisation steps (Adam, clip 1.0) on one batch: finite, parameters move, loss goes down."""
    from eo_vae.models.modules.consistency_loss import EOConsistencyLoss
    from oracle.weights import WAVELENGTHS, synthetic_patches
    model, sd, cfg = _tiny(cuda)
    model.train()
    model.loss_fn = EOConsistencyLoss(pixel_weight=1.0, rec_loss_type="char", msssim_weight=1.0,
                                      msssim_start_step=2000).to(cuda)
    model.base_lr = 2e-4
    wvs = torch.tensor(WAVELENGTHS["S2L2A"], dtype=torch.float32).to(cuda)
    batch = {model.image_key: synthetic_patches(2, 12, cfg["resolution"], seed=5).to(cuda), "wvs": wvs}
    before = {k: v.detach().clone() for k, v in model.named_parameters() if "down.0.block.0.conv1.weight" in k}
    losses = []
    torch.manual_seed(0)
    for step in range(6):
        losses.append(float(model.training_step(batch, step)))
    assert all(l == l and abs(l) < 1e6 for l in losses), losses
    assert losses[-1] < losses[0], losses
    for k, v in model.named_parameters():
        if k in before:
            assert not torch.equal(v.detach(), before[k])


@pytest.mark.parametrize("generator", ["transformer", "factorized"])
def test_weight_distillation_loop(cuda, generator):
    """Stage-1 weight distillation (weight_distill_train.py:190-264): MSE between ``get_distillation_weight(rgb_wvs)`` of
    both dynamic layers and fixed teacher conv weights, Adam on the hypernetworks only.  First-step gradients vs autograd
    over the oracle, then the loss must fall."""
    import torch.nn.functional as F
    from eo_vae.models.modules.dynamic_conv import DynamicConv, DynamicConv_decoder
    from oracle import eovae_oracle as O
    torch.manual_seed(3)
    kw = dict(wv_planes=128, inter_dim=128, kernel_size=3, stride=1, padding=1, embed_dim=128, num_layers=2, num_heads=4,
              generator_type=generator, rank_ratio=4)
    enc, dec = DynamicConv(**kw).to(cuda).eval(), DynamicConv_decoder(**kw).to(cuda).eval()
    wvs = torch.tensor([0.665, 0.56, 0.49], device=cuda)
    g = torch.Generator().manual_seed(1)
    t_enc_w, t_enc_b = 0.1 * torch.randn((128, 3, 3, 3), generator=g), 0.1 * torch.randn((128,), generator=g)
    t_dec_w, t_dec_b = 0.1 * torch.randn((3, 128, 3, 3), generator=g), 0.1 * torch.randn((3,), generator=g)

    def loss_of(ew, eb, dw, db, dev):
        return (F.mse_loss(ew, t_enc_w.to(dev)) + F.mse_loss(eb, t_enc_b.to(dev)) + F.mse_loss(dw, t_dec_w.to(dev))
                + F.mse_loss(db.reshape(-1), t_dec_b.to(dev)))

    params = list(enc.parameters()) + list(dec.parameters())
    opt = torch.optim.Adam(params, lr=1e-3)
    losses = []
    for step in range(12):
        opt.zero_grad(set_to_none=True)
        loss = loss_of(*enc.get_distillation_weight(wvs), *dec.get_distillation_weight(wvs), cuda)
        loss.backward()
        if step == 0:
            sd = {"e." + k: v.detach().cpu().clone().requires_grad_(True) for k, v in enc.state_dict().items()}
            sd.update({"d." + k: v.detach().cpu().clone().requires_grad_(True) for k, v in dec.state_dict().items()})
            ew, eb = O.hypernet(sd, "e", wvs.cpu(), False, heads=4)
            dw, db = O.hypernet(sd, "d", wvs.cpu(), True, heads=4)
            want = loss_of(ew, eb, dw, db * 10.0, "cpu")   # distillation bias: scaled once (dynamic_conv.py:660)
            want.backward()
            assert abs(float(loss) - float(want)) < 1e-4 * abs(float(want))
            for pre, mod in (("e.", enc), ("d.", dec)):
                a = torch.cat([p.grad.flatten().cpu() for _, p in mod.named_parameters()])
                b = torch.cat([sd[pre + n].grad.flatten() for n, _ in mod.named_parameters()])
                assert _rel(a, b) < 2e-3, (pre, _rel(a, b))
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < 0.7 * losses[0], losses


@pytest.mark.parametrize("clip", [None, 1.0, 1e-3])
def test_fused_clip_adam_matches_torch(cuda, clip):
    """eovae_grad_norm + eovae_adam_step (FusedClipAdam) vs clip_grad_norm_ + torch.optim.Adam over several steps on tensors
    of awkward sizes (odd lengths, a scalar, > one chunk), including the state_dict layout."""
    from eo_vae.optim import FusedClipAdam
    g = torch.Generator().manual_seed(2)
    shapes = [(3,), (), (257, 129), (70001,), (64, 3, 3, 3), (5, 7)]
    pa = [torch.randn(s, generator=g).to(cuda).requires_grad_(True) for s in shapes]
    pb = [p.detach().clone().requires_grad_(True) for p in pa]
    oa, ob = FusedClipAdam(pa, lr=3e-3), torch.optim.Adam(pb, lr=3e-3)
    for it in range(5):
        for a, b in zip(pa, pb):
            gr = torch.randn(a.shape, generator=g).to(cuda) * (10.0 if it % 2 else 0.1)
            a.grad, b.grad = gr.clone(), gr.clone()
        if clip:
            want_norm = torch.nn.utils.clip_grad_norm_(pb, clip)
        oa.step(clip_norm=clip)
        ob.step()
        if clip:
            assert abs(float(oa.last_grad_norm) - float(want_norm)) < 1e-5 * float(want_norm)
        for a, b in zip(pa, pb):
            assert float((a - b).abs().max()) < 2e-6 + 1e-5 * float(b.abs().max()), (it, tuple(a.shape))
    sa, sb = oa.state_dict(), ob.state_dict()
    assert sa["param_groups"][0]["lr"] == sb["param_groups"][0]["lr"] and set(sa["state"].keys()) == set(sb["state"].keys())
    for k in sa["state"]:
        assert set(sa["state"][k].keys()) == set(sb["state"][k].keys()) == {"step", "exp_avg", "exp_avg_sq"}
        assert float(sa["state"][k]["step"]) == float(sb["state"][k]["step"]) == 5.0
        for key in ("exp_avg", "exp_avg_sq"):   # fma vs mul + addcmul rounding: a few ulp
            assert float((sa["state"][k][key] - sb["state"][k][key]).abs().max()) < 1e-6 + 1e-4 * float(sb["state"][k][key].abs().max())
    ob.load_state_dict(sa)      # interchangeable checkpoints
    oa.load_state_dict(sb)


def test_eager_training_forward_uses_the_updated_weights(cuda):
    """Round-1 advisor finding: FusedClipAdam writes parameters through raw pointers, and the derived 16-bit weight operands
    (Conv2dSM100.packed_weight, the fused conv2 + shortcut operand, the q/k/v operand, the folded Upsample taps) are keyed on
    the parameters' autograd versions.  After eager steps with freeze_body=False, a forward of the trained model must equal
    a forward of a FRESH model loaded from its state_dict (whose operands are packed from scratch), in eval and in train mode."""
    import __graft_entry__ as g
    from eo_vae.models.modules.consistency_loss import EOConsistencyLoss
    from eo_vae.optim import FusedClipAdam
    from oracle.weights import TINY_CONFIG, WAVELENGTHS, synthetic_patches
    model, sd, cfg = _tiny(cuda)
    model.train()
    model.loss_fn = EOConsistencyLoss(pixel_weight=1.0, rec_loss_type="char").to(cuda)
    model.clip_grad = 1.0
    model.base_lr = 1e-3
    wvs = torch.tensor(WAVELENGTHS["S2L2A"], dtype=torch.float32).to(cuda)
    x = synthetic_patches(2, 12, cfg["resolution"], seed=5).to(cuda)
    batch = {model.image_key: x, "wvs": wvs}
    with torch.no_grad():
        before = model.reconstruct(x, wvs).clone()      # packs every operand cache in the inference dtype
    torch.manual_seed(0)
    for step in range(3):
        model.training_step(batch, step)
    opt = model.optimizers()
    assert isinstance(getattr(opt, "optimizer", opt), FusedClipAdam)
    assert model.global_step == 3                        # the stand-in advances global_step like Lightning does
    fresh = g._model(TINY_CONFIG, {k: v.detach().clone() for k, v in model.state_dict().items()}, cuda)
    model.eval()
    with torch.no_grad():
        after = model.reconstruct(x, wvs)
        want = fresh.reconstruct(x, wvs)
    assert not torch.equal(after, before), "parameters did not move"
    assert torch.equal(after, want), f"stale weight operands in eval forward: rel {_rel(after, want):.3e}"
    # and a taped (train-dtype) forward on both: same posterior noise, same bits
    model.train()
    fresh.train()
    torch.manual_seed(7)
    r1, _ = model(x, wvs)
    torch.manual_seed(7)
    r2, _ = fresh(x, wvs)
    assert torch.equal(r1.detach(), r2.detach()), f"stale weight operands in train forward: rel {_rel(r1, r2):.3e}"


def test_training_step_runs_the_discriminator_half(cuda):
    """The generator / discriminator control flow of the reference step (new_autoencoder.py:638-684) with a user-supplied GAN
    loss module: the generator half receives get_last_layer(), the discriminator half starts at disc_start, trains only the
    discriminator, and its logs are merged."""
    from oracle.weights import WAVELENGTHS, synthetic_patches

    class ToyGanLoss(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.discriminator = torch.nn.Sequential(torch.nn.Conv2d(12, 4, 3, padding=1), torch.nn.LeakyReLU(0.2),
                                                     torch.nn.Conv2d(4, 1, 3, padding=1))
            self.disc_start, self.disc_weight = 2, 0.5
            self.seen = []

        def forward(self, inputs, wvs, reconstructions, optimizer_idx, global_step, last_layer=None, split="train"):
            self.seen.append((optimizer_idx, global_step, None if last_layer is None else tuple(last_layer.shape),
                              self.discriminator.training))
            if optimizer_idx == 0:
                rec = (reconstructions - inputs).abs().mean()
                gen = -self.discriminator(reconstructions).mean() if global_step >= self.disc_start else rec * 0.0
                return rec + self.disc_weight * gen, {f"{split}/rec": rec.detach()}
            d = torch.relu(1.0 - self.discriminator(inputs)).mean() + torch.relu(1.0 + self.discriminator(reconstructions)).mean()
            return d, {f"{split}/disc": d.detach()}

    model, sd, cfg = _tiny(cuda)
    model.train()
    model.loss_fn = ToyGanLoss().to(cuda)
    wvs = torch.tensor(WAVELENGTHS["S2L2A"], dtype=torch.float32).to(cuda)
    batch = {model.image_key: synthetic_patches(2, 12, cfg["resolution"], seed=9).to(cuda), "wvs": wvs}
    d0 = [p.detach().clone() for p in model.loss_fn.discriminator.parameters()]
    e0 = model.encoder.down[0].block[0].conv1.weight.detach().clone()
    torch.manual_seed(0)
    for step in range(3):
        model.training_step(batch, step)
    seen = model.loss_fn.seen
    # two optimisers step per iteration once the discriminator trains: global_step counts optimiser steps (Lightning semantics)
    gen_calls = [s for s in seen if s[0] == 0]
    disc_calls = [s for s in seen if s[0] == 1]
    assert len(gen_calls) == 3 and len(disc_calls) >= 1
    assert all(s[2] == (12, cfg["ch"], 3, 3) for s in gen_calls)         # last_layer = generated decoder kernel [C, E, 3, 3]
    assert all(s[3] is False for s in gen_calls) and all(s[3] is True for s in disc_calls)   # eval() / train() toggling
    assert all(s[1] >= 2 for s in disc_calls)
    assert any(not torch.equal(a, b.detach()) for a, b in zip(d0, model.loss_fn.discriminator.parameters()))
    assert not torch.equal(e0, model.encoder.down[0].block[0].conv1.weight.detach())
    assert "train/disc" in model.logged and "train/rec" in model.logged
