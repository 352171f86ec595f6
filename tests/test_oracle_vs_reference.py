"""CPU suite, part 2 (build container only): the oracle and the drop-in package's checkpoint layout against the
unmodified reference executed from /root/reference.  Skipped where the reference tree does not exist (GPU box)."""
import pytest
import torch

from oracle import eovae_oracle as O
from oracle import ref_shim
from oracle.weights import FULL_CONFIG, TINY_ADAIN_CONFIG, TINY_CONFIG, TINY_FACTORIZED_CONFIG, WAVELENGTHS, make_state_dict, state_dict_spec, synthetic_patches

pytestmark = pytest.mark.skipif(ref_shim.reference_root() is None, reason="reference tree not available")


FULL_FACTORIZED_CONFIG = dict(FULL_CONFIG, generator_type="factorized", rank_ratio=2)  # finetune_consistency_factor.yaml


@pytest.mark.parametrize("cfg", [TINY_CONFIG, FULL_CONFIG, TINY_FACTORIZED_CONFIG, FULL_FACTORIZED_CONFIG,
                                 TINY_ADAIN_CONFIG, dict(FULL_CONFIG, use_adain=True)],
                         ids=["tiny", "full", "tiny-factorized", "full-factorized", "tiny-adain", "full-adain"])
def test_state_dict_layout(cfg):
    model = ref_shim.build_reference_model(cfg)
    ref_sd = model.state_dict()
    spec = state_dict_spec(cfg)
    assert list(ref_sd.keys()) == list(spec.keys())
    for k, shape in spec.items():
        assert tuple(ref_sd[k].shape) == tuple(shape), k


@pytest.mark.parametrize("modality", ["S2RGB", "S1RTC", "S2L2A"])
def test_oracle_equals_reference_live(modality):
    cfg = TINY_CONFIG
    sd = make_state_dict(cfg, 5)
    model = ref_shim.build_reference_model(cfg, sd, train=False)
    wvs = torch.tensor(WAVELENGTHS[modality])
    x = synthetic_patches(2, len(wvs), 32, seed=77)
    with torch.no_grad():
        z_ref = model.encode_spatial_normalized(x, wvs)
        r_ref = model.reconstruct(x, wvs)
        post = model.encode(x, wvs)
        z = O.encode_spatial_normalized(sd, x, wvs, cfg["hyper_heads"])
        r = O.reconstruct(sd, x, wvs, cfg["hyper_heads"])
        m = O.encoder_forward(sd, x, wvs, cfg["hyper_heads"])
    assert torch.allclose(z, z_ref, atol=2e-5) and torch.allclose(r, r_ref, atol=5e-5)
    assert torch.allclose(O.posterior_kl(m), post.kl(), rtol=1e-5)
    assert torch.allclose(O.posterior(m)[1], post.logvar, atol=2e-5)


@pytest.mark.parametrize("modality", ["S2RGB", "S1RTC", "S2L2A"])
def test_factorized_generator_oracle_equals_reference(modality):
    """generator_type='factorized' (dynamic_conv.py:186-302): generated kernels / biases of both dynamic layers, latents
    and reconstructions, and (eval mode, dropout off) every hypernetwork parameter gradient vs reference autograd."""
    cfg = TINY_FACTORIZED_CONFIG
    sd = make_state_dict(cfg, 8)
    model = ref_shim.build_reference_model(cfg, sd, train=False)
    wvs = torch.tensor(WAVELENGTHS[modality])
    x = synthetic_patches(2, len(wvs), 32, seed=80)
    with torch.no_grad():
        w_ref, b_ref = model.encoder.conv_in.get_distillation_weight(wvs)
        w, b = O.hypernet(sd, "encoder.conv_in", wvs, decoder=False, heads=cfg["hyper_heads"])
        assert torch.allclose(w, w_ref, atol=1e-5) and torch.allclose(b, b_ref, atol=1e-5)
        w_ref, b_ref = model.decoder.conv_out.get_distillation_weight(wvs)
        w, b = O.hypernet(sd, "decoder.conv_out", wvs, decoder=True, heads=cfg["hyper_heads"])
        # get_distillation_weight scales the decoder bias once (:660), the forward twice (:692-697)
        assert torch.allclose(w, w_ref, atol=1e-5) and torch.allclose(b * 10.0, b_ref.reshape(-1), atol=1e-5)
        z_ref = model.encode_spatial_normalized(x, wvs)
        r_ref = model.reconstruct(x, wvs)
        z = O.encode_spatial_normalized(sd, x, wvs, cfg["hyper_heads"])
        r = O.reconstruct(sd, x, wvs, cfg["hyper_heads"])
    assert torch.allclose(z, z_ref, atol=2e-5) and torch.allclose(r, r_ref, atol=5e-5)
    # gradients (eval mode: deterministic)
    r_ref = model(x, wvs, sample_posterior=False)[0]       # reconstruct() itself runs under no_grad
    torch.sqrt((r_ref - x) ** 2 + 1e-6).mean().backward()
    osd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v.clone()) for k, v in sd.items()}
    O.charbonnier_loss(O.forward(osd, x, wvs, None, False, cfg["hyper_heads"])[0], x).backward()
    floor = 1e-6 * max(float(p.grad.abs().max()) for p in model.parameters() if p.grad is not None)
    checked = 0
    for name, p in model.named_parameters():
        if "weight_generator" not in name and "fclayer" not in name:
            continue
        got = osd[name].grad
        assert got is not None and p.grad is not None, name
        assert float((got - p.grad).abs().max()) < 2e-4 * float(p.grad.abs().max()) + floor, name
        checked += 1
    assert checked > 60


@pytest.mark.parametrize("modality", ["S2RGB", "S2L2A"])
def test_adain_oracle_equals_reference(modality):
    """use_adain=True (model.py:35-64, 96-100, 173-191, 331-343; layers.py:68-76, 96-104): latents, reconstruction and
    every parameter gradient (conditioner MLPs and emb_proj included) vs the unmodified reference."""
    cfg = TINY_ADAIN_CONFIG
    sd = make_state_dict(cfg, 9)
    model = ref_shim.build_reference_model(cfg, sd, train=False)
    wvs = torch.tensor(WAVELENGTHS[modality])
    x = synthetic_patches(2, len(wvs), 32, seed=81)
    with torch.no_grad():
        z_ref = model.encode_spatial_normalized(x, wvs)
        r_ref = model.reconstruct(x, wvs)
        z = O.encode_spatial_normalized(sd, x, wvs, cfg["hyper_heads"])
        r = O.reconstruct(sd, x, wvs, cfg["hyper_heads"])
    assert torch.allclose(z, z_ref, atol=2e-5) and torch.allclose(r, r_ref, atol=5e-5)
    r_ref = model(x, wvs, sample_posterior=False)[0]
    torch.sqrt((r_ref - x) ** 2 + 1e-6).mean().backward()
    osd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v.clone()) for k, v in sd.items()}
    O.charbonnier_loss(O.forward(osd, x, wvs, None, False, cfg["hyper_heads"])[0], x).backward()
    floor = 1e-6 * max(float(p.grad.abs().max()) for p in model.parameters() if p.grad is not None)
    checked = 0
    for name, p in model.named_parameters():
        got = osd[name].grad
        assert got is not None and p.grad is not None, name
        assert float((got - p.grad).abs().max()) < 2e-4 * float(p.grad.abs().max()) + floor, name
        checked += "conditioner" in name or "emb_proj" in name
    assert checked >= 12 + 2 * 13  # two conditioners (6 tensors each) + emb_proj (w, b) of 13 ResnetBlocks


def test_train_mode_forward_matches_reference():
    """Sampled, train-mode forward (batch-statistics BN, inverse BN on running stats) with the noise passed in."""
    cfg = TINY_CONFIG
    sd = make_state_dict(cfg, 6)
    model = ref_shim.build_reference_model(cfg, sd, train=True)
    wvs = torch.tensor(WAVELENGTHS["S2L2A"])
    x = synthetic_patches(3, 12, 32, seed=78)
    torch.manual_seed(123)
    with torch.no_grad():
        recon_ref, post = model(x, wvs)          # draws eps with the CPU generator (distributions.py:44)
        torch.manual_seed(123)
        eps = torch.randn(post.mean.shape)
        recon, _ = O.forward(sd, x, wvs, eps=eps, train=True, heads=cfg["hyper_heads"])
    assert torch.allclose(recon, recon_ref, atol=1e-4)


def test_train_mode_gradients_match_reference():
    """Pins the oracle's GRADIENT semantics: train-mode forward + Charbonnier loss, every parameter gradient of the
    unmodified reference (torch autograd) equals autograd over the oracle - in particular the BatchNorm running-statistics
    update is outside the graph, although the inverse normalisation of the same forward reads the updated buffers."""
    cfg = TINY_CONFIG
    sd = make_state_dict(cfg, 6)
    model = ref_shim.build_reference_model(cfg, sd, train=True)
    wvs = torch.tensor(WAVELENGTHS["S2L2A"])
    x = synthetic_patches(2, 12, 32, seed=79)
    torch.manual_seed(321)
    recon_ref, post = model(x, wvs)
    torch.sqrt((recon_ref - x) ** 2 + 1e-3 ** 2).mean().backward()        # CharbonnierLoss, consistency_loss.py:12-21
    torch.manual_seed(321)
    eps = torch.randn(post.mean.shape)
    osd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v.clone()) for k, v in sd.items()}
    recon, _ = O.forward(osd, x, wvs, eps=eps, train=True, heads=cfg["hyper_heads"])
    O.charbonnier_loss(recon, x).backward()
    checked = 0
    # conv biases in front of a GroupNorm with one channel per group have a mathematically zero gradient (rounding noise
    # ~1e-9 on both sides), hence the absolute floor relative to the largest gradient of the model
    floor = 1e-6 * max(float(p.grad.abs().max()) for p in model.parameters())
    for name, p in model.named_parameters():
        assert p.grad is not None, name
        got = osd[name].grad
        assert got is not None, name
        scale = float(p.grad.abs().max())
        assert float((got - p.grad).abs().max()) < 2e-4 * scale + floor, name
        checked += 1
    assert checked > 150


def test_dropin_package_has_reference_checkpoint_layout():
    """eo_vae (this repo) registers exactly the reference's parameters/buffers: a reference checkpoint loads strictly."""
    import __graft_entry__ as g
    import sys
    if g.PKG not in sys.path:
        sys.path.insert(0, g.PKG)
    from eo_vae.models import Decoder, Encoder, EOFluxVAE
    cfg = TINY_CONFIG
    dyn = dict(num_layers=cfg["hyper_layers"], wv_planes=cfg["wv_planes"], num_heads=cfg["hyper_heads"])
    enc = Encoder(resolution=64, in_channels=3, ch=cfg["ch"], ch_mult=list(cfg["ch_mult"]), num_res_blocks=1,
                  z_channels=cfg["z_channels"], use_dynamic_ops=True, dynamic_conv_kwargs=dict(dyn))
    dec = Decoder(ch=cfg["ch"], out_ch=3, ch_mult=list(cfg["ch_mult"]), num_res_blocks=1, resolution=64,
                  z_channels=cfg["z_channels"], use_dynamic_ops=True, dynamic_conv_kwargs=dict(dyn))
    ours = EOFluxVAE(enc, dec, torch.nn.Identity(), freeze_body=False)
    ref = ref_shim.build_reference_model(cfg, make_state_dict(cfg, 2))
    assert list(ours.state_dict().keys()) == list(ref.state_dict().keys())
    ours.load_state_dict(ref.state_dict(), strict=True)
    # same trainable set after freezing the body (new_autoencoder.py:274-293)
    frozen = EOFluxVAE(enc, dec, torch.nn.Identity(), freeze_body=True)
    names = {n for n, p in frozen.named_parameters() if p.requires_grad}
    assert names and all(n.startswith(("encoder.conv_in.", "decoder.conv_out.")) for n in names)


@pytest.mark.parametrize("shape", [(2, 12, 24, 20), (3, 2, 17, 33)])
def test_spectral_and_spatial_losses_match_reference(shape):
    """SAMLoss / GradientDifferenceLoss / EOConsistencyLoss with spectral + spatial branches (consistency_loss.py:186-210,
    241-269, 426-440): oracle restatement vs the unmodified reference classes - values and d/d(reconstruction)."""
    ref = ref_shim.load_reference_losses()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(shape, generator=g)
    r0 = x + 0.3 * torch.randn(shape, generator=g)
    r0[0, :, 0, 0] = 0.0                      # a zero spectrum: norm subgradient 0, eps keeps the quotient finite
    for fn_ref, fn in ((ref.SAMLoss(), O.sam_loss), (ref.GradientDifferenceLoss(), O.grad_diff_loss)):
        a, b = r0.clone().requires_grad_(True), r0.clone().requires_grad_(True)
        la, lb = fn_ref(a, x), fn(b, x)
        la.backward(); lb.backward()
        assert torch.allclose(la, lb, rtol=1e-6, atol=1e-7)
        assert torch.allclose(a.grad, b.grad, rtol=1e-5, atol=1e-9)
    loss = ref.EOConsistencyLoss(pixel_weight=1.0, rec_loss_type="char", spectral_weight=0.5, spatial_weight=2.0,
                                 spatial_start_step=10)
    for step in (0, 10):
        total, logs = loss(inputs=x, wvs=None, reconstructions=r0, global_step=step)
        want, _, _ = O.consistency_loss(x, r0, "char", 1.0, 0.0, step, 0, spectral_weight=0.5, spatial_weight=2.0,
                                        spatial_start_step=10)
        assert torch.allclose(total, want, rtol=1e-6)
        assert ("train/loss_spatial" in logs) == (step >= 10) and "train/loss_spectral" in logs


@pytest.mark.parametrize("shape,pf,alpha", [((2, 3, 32, 32), 2, 1.0), ((1, 12, 24, 36), 1, 1.0), ((2, 2, 28, 28), 2, 0.5)])
def test_focal_frequency_loss_matches_reference(shape, pf, alpha):
    """FocalFrequencyLoss (ffl.py:17-104) in EOConsistencyLoss's configuration, and the freq branch with its weight warm-up
    (consistency_loss.py:442-463): oracle vs the unmodified reference, value and d/d(reconstruction)."""
    ref = ref_shim.load_reference_losses()
    g = torch.Generator().manual_seed(6)
    x = torch.randn(shape, generator=g)
    r0 = x + 0.3 * torch.randn(shape, generator=g)
    fn_ref = ref.FFL(loss_weight=1.0, alpha=alpha, patch_factor=pf, ave_spectrum=False, batch_matrix=True, log_matrix=True)
    a, b = r0.clone().requires_grad_(True), r0.clone().requires_grad_(True)
    la, lb = fn_ref(a, x), O.focal_freq_loss(b, x, pf, alpha)
    la.backward(); lb.backward()
    assert torch.allclose(la, lb, rtol=1e-5) and torch.allclose(a.grad, b.grad, rtol=1e-4, atol=1e-8)
    loss = ref.EOConsistencyLoss(pixel_weight=1.0, rec_loss_type="l1", freq_weight=3.0, freq_start_step=100, patch_factor=pf,
                                 ffl_alpha=alpha)
    for step in (0, 100, 600, 5000):
        total, logs = loss(inputs=x, wvs=None, reconstructions=r0, global_step=step)
        want, _, _ = O.consistency_loss(x, r0, "l1", 1.0, 0.0, step, 0, freq_weight=3.0, freq_start_step=100, patch_factor=pf,
                                        ffl_alpha=alpha)
        assert torch.allclose(total, want, rtol=1e-5), step
        assert ("train/loss_freq_raw" in logs) == (step >= 100)


@pytest.mark.parametrize("scale,angle", [(0.5, 1), (0.75, 3), ((0.375, 0.75), 2), (0.5, None), (None, 1)])
def test_eq_vae_transforms_match_reference(scale, angle):
    """EQ-VAE modes (new_autoencoder.py:460-464, 519-531, 605-627): train-mode forward with a rescaled / rotated latent and
    the area-averaged, rotated target - oracle vs the unmodified reference (same posterior noise)."""
    cfg = TINY_CONFIG
    sd = make_state_dict(cfg, 6)
    model = ref_shim.build_reference_model(cfg, sd, train=True)
    wvs = torch.tensor(WAVELENGTHS["S2RGB"])
    x = synthetic_patches(2, 3, 64, seed=83)
    torch.manual_seed(7)
    with torch.no_grad():
        recon_ref, post = model(x, wvs, scale=scale, angle=angle)
        torch.manual_seed(7)
        eps = torch.randn(post.mean.shape)
        recon, _ = O.forward(sd, x, wvs, eps=eps, train=True, heads=cfg["hyper_heads"], scale=scale, angle=angle)
    assert recon.shape == recon_ref.shape and torch.allclose(recon, recon_ref, atol=1e-4)
    import torch.nn.functional as F
    t_ref = F.interpolate(x, size=recon_ref.shape[-2:], mode="area")
    if angle is not None:
        t_ref = torch.rot90(t_ref, k=angle, dims=[-1, -2])
    assert torch.equal(O.eq_target(x, recon_ref.shape[-2:], angle), t_ref)


# ----------------------------------------------------------------------------------------------------------------------
# Pins of the "next" rows' oracles: the reference files cannot be imported whole (hydra, lightning,
# matplotlib ... at module level), so their definitions are executed from the parsed, otherwise unmodified source
# (ref_shim.load_reference_definitions).

def _ref_encode_latents():
    return ref_shim.load_reference_definitions(
        "encode_latents.py", drop_imports=("lightning", "omegaconf", "hydra", "matplotlib", "eo_vae", "tqdm", "einops"))


def _ref_datamodule():
    return ref_shim.load_reference_definitions("eo_vae/datasets/terramesh_datamodule.py",
                                               drop_imports=("lightning", "omegaconf"), drop_defs=("TerraMeshDataModule",))


@pytest.mark.parametrize("shape", [(3, 4, 8, 8), (1, 32, 16, 16), (5, 2, 7, 9)])
def test_running_stats_oracle_equals_reference(shape):
    """O.running_stats_update vs RunningStatsButFast.update (encode_latents.py:36-109): every buffer after every update of
    a stream of batches with drifting mean / scale (the parallel-variance merge is order dependent), bit for bit."""
    ref = _ref_encode_latents().RunningStatsButFast((shape[1],), [0, 2, 3])
    st = O.running_stats_init(shape[1])
    g = torch.Generator().manual_seed(17)
    for i in range(6):
        x = torch.randn(shape, generator=g) * (1.0 + 0.5 * i) + 0.3 * i
        ref(x)
        st = O.running_stats_update(st, x)
        for k in ("mean", "var", "std", "count", "min", "max"):
            assert torch.equal(getattr(ref, k), st[k]), (i, k)
    # the module in the datamodule file is the same class body: pin that copy too
    ref2 = _ref_datamodule().RunningStatsButFast((shape[1],), [0, 2, 3])
    ref2(x)
    assert torch.equal(ref2.mean, O.running_stats_update(O.running_stats_init(shape[1]), x)["mean"])


@pytest.mark.parametrize("modality,scheme", [("S2L2A", "custom"), ("S2L1C", "custom"), ("S2L2A", "legacy"),
                                             ("S1RTC", "legacy"), ("S2RGB", "legacy")])
@pytest.mark.parametrize("target", [None, (24, 24), (40, 56)])
def test_preprocess_oracle_equals_reference_collate(modality, scheme, target):
    """O.preprocess vs the reference's own collate closure (terramesh_datamodule.py:418-503: normaliser -> bilinear resize
    -> apply_batch_augmentations), train mode, for eight seeds (all three D4 draws vary).  The oracle receives the flags the
    reference drew - Python's ``random`` replayed with the same seed, in the reference's draw order."""
    import random
    dm = _ref_datamodule()
    bands = len(dm.WAVELENGTHS[modality])
    collate = dm.single_modality_collate_fn([modality], normalize=True, norm_scheme=scheme, target_size=target, mode="train")
    norm = dm.NormalizerFactory.create(modality, scheme)
    custom = not isinstance(norm, dm.LegacyZScoreNorm)
    mean, std = norm.mean.reshape(-1).float(), norm.std.reshape(-1).float()
    g = torch.Generator().manual_seed(5)
    # raw digital numbers incl. values outside the clip range of the custom scheme
    x = torch.rand((2, bands, 32, 48), generator=g) * 14000.0 - 2000.0
    for seed in range(8):
        random.seed(seed)
        got = collate({"image": x.clone()})
        random.seed(seed)
        fh, fv, k = random.random() > 0.5, random.random() > 0.5, random.randint(0, 3)
        want = O.preprocess(x, mean, std, custom, target, fh, fv, k)
        assert got["image"].shape == want.shape, (seed, got["image"].shape, want.shape)
        assert torch.equal(got["image"], want), (seed, float((got["image"] - want).abs().max()))
        assert torch.equal(got["wvs"], torch.tensor(WAVELENGTHS[modality])) and got["modality"] == modality
    # eval mode: no augmentation
    ev = dm.single_modality_collate_fn([modality], normalize=True, norm_scheme=scheme, target_size=target, mode="eval")
    assert torch.equal(ev({"image": x.clone()})["image"], O.preprocess(x, mean, std, custom, target))


def test_product_preprocess_constants_equal_reference():
    """The statistics baked into eo_vae.preprocess are the reference's (terramesh_datamodule.py:141-182)."""
    from eo_vae import preprocess as P
    n = _ref_datamodule().Sentinel2L2ANorm()
    assert torch.equal(n.mean.reshape(-1), torch.tensor(P.S2L2A_CUSTOM_MEAN))
    assert torch.equal(n.std.reshape(-1), torch.tensor(P.S2L2A_CUSTOM_STD))
