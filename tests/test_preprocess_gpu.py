"""the data-side prologue kernel (clip, z-score, bilinear resize, D4 augmentation) against the oracle's
restatement of the reference collate function, for every element of the D4 group and both normalisation schemes."""
import itertools

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("scheme", ["custom", "legacy"])
@pytest.mark.parametrize("in_dtype", [torch.float32, torch.int16, torch.uint16])
def test_preprocess_matches_collate(cuda, scheme, in_dtype):
    from eo_vae.preprocess import S2L2A_CUSTOM_MEAN, S2L2A_CUSTOM_STD, BatchPreprocessor
    from oracle import eovae_oracle as O
    g = torch.Generator().manual_seed(8)
    raw = torch.randint(-200, 12000, (3, 12, 40, 56), generator=g)
    if in_dtype == torch.uint16:
        raw = raw.clamp_min(0)
    raw_t = raw.to(in_dtype) if in_dtype != torch.float32 else raw.float() + 0.25
    mean, std = torch.tensor(S2L2A_CUSTOM_MEAN), torch.tensor(S2L2A_CUSTOM_STD)
    for target in (None, (64, 48), (40, 56)):
        pre = BatchPreprocessor(mean, std, scheme=scheme, target_size=target).to(cuda)
        for fh, fv, k in itertools.product((False, True), (False, True), range(4)):
            got = pre(raw_t.to(cuda), augment=(fh, fv, k)).cpu()
            ref = O.preprocess(raw_t.float(), mean, std, scheme == "custom", target, fh, fv, k)
            assert got.shape == ref.shape, (target, fh, fv, k)
            assert torch.allclose(got, ref, atol=2e-4, rtol=1e-5), (target, fh, fv, k, float((got - ref).abs().max()))


def test_preprocess_feeds_the_model(cuda):
    """raw uint16 DNs -> prologue kernel -> encode_spatial_normalized equals the oracle on the oracle-preprocessed batch."""
    import __graft_entry__ as ge
    from eo_vae.preprocess import S2L2A_CUSTOM_MEAN, S2L2A_CUSTOM_STD, BatchPreprocessor
    from oracle import eovae_oracle as O
    from oracle.weights import TINY_CONFIG, WAVELENGTHS, make_state_dict
    sd = make_state_dict(TINY_CONFIG, seed=3)
    model = ge._model(TINY_CONFIG, sd, cuda)
    raw = torch.randint(0, 9000, (2, 12, 48, 48), generator=torch.Generator().manual_seed(2)).to(torch.uint16)
    mean, std = torch.tensor(S2L2A_CUSTOM_MEAN), torch.tensor(S2L2A_CUSTOM_STD)
    pre = BatchPreprocessor(mean, std, scheme="custom", target_size=(64, 64)).to(cuda)
    wvs = torch.tensor(WAVELENGTHS["S2L2A"], dtype=torch.float32)
    with torch.no_grad():
        z = model.encode_spatial_normalized(pre(raw.to(cuda), augment=(True, False, 1)), wvs.to(cuda)).cpu()
        z_ref = O.encode_spatial_normalized(sd, O.preprocess(raw.float(), mean, std, True, (64, 64), True, False, 1), wvs,
                                            TINY_CONFIG["hyper_heads"])
    assert float((z - z_ref).norm() / z_ref.norm()) < 3e-2
