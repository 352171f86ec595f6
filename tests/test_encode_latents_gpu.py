"""the encode_latents pipeline around the encoder - device-side running statistics (one kernel per batch),
raw-latent encode, asynchronous .npz writer - against the oracle's restatement of the reference formulas."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_running_stats_match_reference_formulas(cuda):
    from eo_vae.encode_latents import RunningStatsButFast
    from oracle import eovae_oracle as O
    rs = RunningStatsButFast((32,), [0, 2, 3]).to(cuda)
    st = O.running_stats_init(32)
    g = torch.Generator().manual_seed(5)
    for b in (4, 1, 7):
        x = torch.randn((b, 32, 16, 16), generator=g) * 3 + 0.5
        assert rs(x.to(cuda)) is not None
        st = O.running_stats_update(st, x)
    got = rs.get_stats_dict()
    for k in ("mean", "var", "std", "min", "max", "count"):
        assert torch.allclose(got[k], st[k], rtol=2e-5, atol=2e-6), (k, got[k][:4], st[k][:4])


def test_encode_split_writes_latents_and_statistics(cuda, tmp_path):
    import __graft_entry__ as ge
    from eo_vae.encode_latents import RunningStatsButFast, encode_raw, encode_spatial_norm, encode_split
    from oracle import eovae_oracle as O
    from oracle.weights import TINY_CONFIG, WAVELENGTHS, make_state_dict, synthetic_patches
    sd = make_state_dict(TINY_CONFIG, seed=3)
    model = ge._model(TINY_CONFIG, sd, cuda)
    wv_lr = torch.tensor(WAVELENGTHS["S2RGB"], dtype=torch.float32, device=cuda)
    wv_hr = torch.tensor(WAVELENGTHS["S2L2A"], dtype=torch.float32, device=cuda)
    batches = [{"image_lr": synthetic_patches(2, 3, 64, seed=40 + i), "image_hr": synthetic_patches(2, 12, 64, seed=50 + i),
                "aoi": [f"aoi_{i}_{j}" for j in range(2)]} for i in range(3)]
    s_lr, s_hr = RunningStatsButFast((8,), [0, 2, 3]), RunningStatsButFast((8,), [0, 2, 3])
    encode_split(model, batches, str(tmp_path), cuda, wv_lr, wv_hr, s_lr, s_hr, "train", encode_fn=encode_raw)
    files = sorted(os.listdir(tmp_path))
    assert files == sorted(f"aoi_{i}_{j}.npz" for i in range(3) for j in range(2))
    # file contents = the raw latents (posterior means) of the oracle, statistics = the reference formulas over them
    st = O.running_stats_init(8)
    for b in batches:
        with torch.no_grad():
            mean = O.encoder_forward(sd, b["image_hr"], wv_hr.cpu(), TINY_CONFIG["hyper_heads"])[:, :8]
        st = O.running_stats_update(st, mean)
        for j, aoi in enumerate(b["aoi"]):
            f = np.load(os.path.join(tmp_path, f"{aoi}.npz"))
            assert f["hr_latent"].shape == (8, 16, 16) and f["lr_image"].shape == (3, 64, 64)
            ref = mean[j].numpy()
            assert np.linalg.norm(f["hr_latent"] - ref) / np.linalg.norm(ref) < 3e-2
            assert np.array_equal(f["hr_image"], b["image_hr"][j].numpy())
    got = s_hr.get_stats_dict()
    assert float(got["count"]) == float(st["count"])
    assert torch.allclose(got["mean"], st["mean"], atol=3e-2) and torch.allclose(got["std"], st["std"], rtol=5e-2)
    z = encode_spatial_norm(model, batches[0]["image_hr"].to(cuda), wv_hr)
    assert tuple(z.shape) == (2, 8, 16, 16)
