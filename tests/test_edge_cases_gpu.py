"""Edge cases of the drop-in surface: ragged (non-square) patches, batch 1, every modality's band count, and the error
behaviour the boundary promises (no CPU path, wavelength / band mismatch, missing wavelengths - model.py:170,353)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-30))


@pytest.mark.parametrize("modality,batch,h,w", [("S2L2A", 1, 64, 96), ("S1RTC", 3, 48, 80), ("S2RGB", 2, 112, 64),
                                                ("S2L1C", 1, 32, 32), ("S2L2A", 5, 16, 48)])
def test_ragged_patches_forward_and_backward(cuda, modality, batch, h, w):
    """H != W, H and W any multiple of 16 (three stride-2 stages + 2 x 2 latent packing), odd batch sizes: latents,
    reconstruction and the all-parameter gradient vs the oracle."""
    import __graft_entry__ as g
    from oracle import eovae_oracle as O
    from oracle.weights import TINY_CONFIG, WAVELENGTHS, _rng, make_state_dict
    cfg = TINY_CONFIG
    sd = make_state_dict(cfg, 2)
    model = g._model(cfg, sd, cuda)
    wvs = torch.tensor(WAVELENGTHS[modality])
    x = torch.from_numpy(_rng(5, f"edge{batch}x{h}x{w}").standard_normal((batch, len(wvs), h, w)).astype("float32")).clamp(-2, 6)
    with torch.no_grad():
        z = model.encode_spatial_normalized(x.to(cuda), wvs.to(cuda))
        r = model.reconstruct(x.to(cuda), wvs.to(cuda))
    z_ref = O.encode_spatial_normalized(sd, x, wvs, cfg["hyper_heads"])
    r_ref = O.reconstruct(sd, x, wvs, cfg["hyper_heads"])
    assert z.shape == z_ref.shape == (batch, cfg["z_channels"], h // 4, w // 4) and r.shape == x.shape
    print(f"ragged {modality} {batch}x{h}x{w}: latent {_rel(z.cpu(), z_ref):.3e} recon {_rel(r.cpu(), r_ref):.3e}")
    assert _rel(z.cpu(), z_ref) < 2e-2 and _rel(r.cpu(), r_ref) < 5e-2
    model.train()
    recon, _ = model(x.to(cuda), wvs.to(cuda), sample_posterior=False)
    torch.sqrt((recon - x.to(cuda)) ** 2 + 1e-6).mean().backward()
    osd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v.clone()) for k, v in sd.items()}
    rr, _ = O.forward(osd, x, wvs, None, True, cfg["hyper_heads"])
    O.charbonnier_loss(rr, x).backward()
    a = torch.cat([p.grad.flatten().float().cpu() for _, p in model.named_parameters()])
    b = torch.cat([osd[n].grad.flatten() for n, _ in model.named_parameters()])
    print(f"    all-parameter gradient rel-L2 {_rel(a, b):.3e}")
    assert _rel(a, b) < 8e-2


def test_error_behaviour(cuda):
    import __graft_entry__ as g
    from oracle.weights import TINY_CONFIG, WAVELENGTHS, make_state_dict
    cfg = TINY_CONFIG
    model = g._model(cfg, make_state_dict(cfg, 2), cuda)
    wvs = torch.tensor(WAVELENGTHS["S2RGB"], device=cuda)
    x = torch.zeros((1, 3, 32, 32), device=cuda)
    with torch.no_grad():
        with pytest.raises(AssertionError):                       # model.py:170: wvs is required by the dynamic encoder
            model.encoder(x, None)
        with pytest.raises(AssertionError):                       # model.py:353
            model.decoder(torch.zeros((1, cfg["z_channels"], 8, 8), device=cuda), None)
        with pytest.raises(RuntimeError, match="wavelengths"):    # 3 bands, 2 wavelengths
            model.encode(x, wvs[:2])
        with pytest.raises(RuntimeError, match="CUDA"):           # no CPU implementation behind the API
            model.encode(x.cpu(), wvs)
        # latent height 36 / 4 = 9 is odd: the 2 x 2 latent packing is impossible (the reference's rearrange raises too)
        with pytest.raises(RuntimeError, match="even"):
            model.encode_spatial_normalized(torch.zeros((1, 3, 36, 32), device=cuda), wvs)
