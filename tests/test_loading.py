"""CPU suite: the construction / checkpoint entry points of the drop-in ``EOFluxVAE`` (from_config, from_pretrained, the three
checkpoint formats of ``_load_checkpoint``: new_autoencoder.py:143-263, 295-417) - no kernel is launched.  Where the
reference tree is available (build container) the same files are loaded by the unmodified reference and the resulting
``state_dict`` must be identical tensor by tensor."""
import os
import sys

import pytest
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "eo-vae_b200"))

from oracle import ref_shim  # noqa: E402
from oracle.weights import TINY_ADAIN_CONFIG, TINY_CONFIG, TINY_FACTORIZED_CONFIG, make_state_dict  # noqa: E402


def _yaml_config(cfg: dict, full_train_layout: bool) -> dict:
    """The two layouts from_config accepts: the training config (``model:`` section with hydra ``_target_`` keys,
    configs/eo-vae.yaml:12-57) and the minimal ``model_config.yaml`` shipped next to a checkpoint."""
    dyn = dict(num_layers=cfg["hyper_layers"], wv_planes=cfg["wv_planes"], num_heads=cfg["hyper_heads"])
    if cfg.get("generator_type", "transformer") != "transformer":
        dyn.update(generator_type=cfg["generator_type"], rank_ratio=cfg["rank_ratio"])
    if cfg.get("use_adain"):
        dyn.update(use_adain=True)
    common = dict(ch=cfg["ch"], ch_mult=list(cfg["ch_mult"]), num_res_blocks=cfg["num_res_blocks"],
                  resolution=cfg["resolution"], z_channels=cfg["z_channels"], use_dynamic_ops=True,
                  dynamic_conv_kwargs=dyn)
    model = dict(encoder=dict(_target_="eo_vae.models.Encoder", in_channels=3, **common),
                 decoder=dict(_target_="eo_vae.models.Decoder", out_ch=3, **common),
                 freeze_body=False, base_lr=2e-4, clip_grad=1.0, image_key="image")
    if full_train_layout:
        model["_target_"] = "eo_vae.models.new_autoencoder.EOFluxVAE"
        return dict(model=model, trainer=dict(max_epochs=3))
    return model


def _write(tmp_path, cfg, full_train_layout=False):
    path = os.path.join(tmp_path, "model_config.yaml")
    with open(path, "w") as f:
        yaml.safe_dump(_yaml_config(cfg, full_train_layout), f)
    return path


def _same(sd_a, sd_b):
    assert list(sd_a.keys()) == list(sd_b.keys())
    for k in sd_a:
        assert torch.equal(sd_a[k], sd_b[k]), k


@pytest.mark.parametrize("cfg", [TINY_CONFIG, TINY_FACTORIZED_CONFIG, TINY_ADAIN_CONFIG], ids=["transformer", "factorized", "adain"])
@pytest.mark.parametrize("layout", ["minimal", "train"])
def test_from_config_full_checkpoint(tmp_path, cfg, layout):
    from eo_vae.models.new_autoencoder import EOFluxVAE
    sd = make_state_dict(cfg, 12)
    cfg_path = _write(str(tmp_path), cfg, layout == "train")
    ckpt = os.path.join(str(tmp_path), "eo-vae.ckpt")
    torch.save({"state_dict": sd, "epoch": 3}, ckpt)               # Lightning layout (new_autoencoder.py:329)
    model = EOFluxVAE.from_config(cfg_path, ckpt)
    assert not model.training and model.base_lr == 2e-4 and model.clip_grad == 1.0 and not model.freeze_body
    _same(model.state_dict(), sd)
    assert all(p.requires_grad for p in model.parameters())
    frozen = EOFluxVAE.from_config(cfg_path, ckpt, freeze_body=True, eval_mode=False)
    assert frozen.training
    for n, p in frozen.named_parameters():
        assert p.requires_grad == ("encoder.conv_in" in n or "decoder.conv_out" in n), n
    if ref_shim.reference_root() is not None:
        ref = ref_shim.load_reference().new_autoencoder.EOFluxVAE.from_config(cfg_path, ckpt)
        _same(ref.state_dict(), model.state_dict())
        assert ref.base_lr == model.base_lr and ref.clip_grad == model.clip_grad and ref.image_key == model.image_key


def test_flux_body_safetensors_and_distilled_pt(tmp_path):
    """Stage-0/1 initialisation: a Flux AE ``.safetensors`` (static conv_in / conv_out, no hypernetworks, no bn) loads the
    body only; a distilled ``.pt`` loads the two dynamic layers only (new_autoencoder.py:311-316, 347-372)."""
    from safetensors.torch import save_file

    from eo_vae.models.new_autoencoder import EOFluxVAE
    cfg = TINY_CONFIG
    sd = make_state_dict(cfg, 13)
    cfg_path = _write(str(tmp_path), cfg)
    body = {k: v.contiguous() for k, v in sd.items()
            if "weight_generator" not in k and "fclayer" not in k and not k.startswith("bn.")}
    body["encoder.conv_in.weight"] = torch.randn(cfg["ch"], 3, 3, 3)      # Flux's static edge layers: must be skipped
    body["encoder.conv_in.bias"] = torch.randn(cfg["ch"])
    body["decoder.conv_out.weight"] = torch.randn(3, cfg["ch"], 3, 3)
    body["decoder.conv_out.bias"] = torch.randn(3)
    st = os.path.join(str(tmp_path), "ae.safetensors")
    save_file(body, st)
    with pytest.raises(RuntimeError):          # the latent BatchNorm buffers are not in a Flux AE file (reference: same)
        EOFluxVAE.from_config(cfg_path, st)
    torch.manual_seed(0)
    model = EOFluxVAE.from_config(cfg_path, st, ignore_keys=["bn"])
    got = model.state_dict()
    for k, v in sd.items():
        if "weight_generator" in k or "fclayer" in k or k.startswith("bn."):
            continue
        assert torch.equal(got[k], v), k
    assert not torch.equal(got["encoder.conv_in.weight_generator.weight_tokens"], sd["encoder.conv_in.weight_generator.weight_tokens"])
    # distilled dynamic layers on top
    pt = os.path.join(str(tmp_path), "distilled.pt")
    torch.save({"encoder_conv_in_state_dict": {k[len("encoder.conv_in."):]: v for k, v in sd.items() if k.startswith("encoder.conv_in.")},
                "decoder_conv_out_state_dict": {k[len("decoder.conv_out."):]: v for k, v in sd.items() if k.startswith("decoder.conv_out.")}}, pt)
    model._load_checkpoint(pt, [])
    got = model.state_dict()
    for k, v in sd.items():
        if not k.startswith("bn."):
            assert torch.equal(got[k], v), k
    if ref_shim.reference_root() is not None:
        torch.manual_seed(0)
        ref = ref_shim.load_reference().new_autoencoder.EOFluxVAE.from_config(cfg_path, st, ignore_keys=["bn"])
        ref._load_checkpoint(pt, [])
        for k, v in ref.state_dict().items():
            if not k.startswith("bn."):
                assert torch.equal(got[k], v), k


def test_checkpoint_errors(tmp_path, capsys):
    from eo_vae.models.new_autoencoder import EOFluxVAE
    cfg = TINY_CONFIG
    sd = make_state_dict(cfg, 14)
    cfg_path = _write(str(tmp_path), cfg)
    with pytest.raises(FileNotFoundError):
        EOFluxVAE.from_config(os.path.join(str(tmp_path), "nope.yaml"))
    bad = os.path.join(str(tmp_path), "bad.yaml")
    with open(bad, "w") as f:
        yaml.safe_dump(dict(model=dict(encoder=dict())), f)
    with pytest.raises(ValueError):
        EOFluxVAE.from_config(bad)
    # a missing checkpoint file is reported and skipped (new_autoencoder.py:303-305)
    model = EOFluxVAE.from_config(cfg_path, os.path.join(str(tmp_path), "missing.ckpt"))
    assert "Checkpoint not found" in capsys.readouterr().out and model is not None
    # body weights missing -> RuntimeError; the same keys under ignore_keys -> accepted
    partial = {k: v for k, v in sd.items() if not k.startswith("decoder.mid")}
    ck = os.path.join(str(tmp_path), "partial.ckpt")
    torch.save(partial, ck)
    with pytest.raises(RuntimeError):
        EOFluxVAE.from_config(cfg_path, ck)
    EOFluxVAE.from_config(cfg_path, ck, ignore_keys=["decoder.mid"])


def test_from_pretrained_uses_hub_files(tmp_path, monkeypatch):
    """from_pretrained = hf_hub_download(config) + hf_hub_download(checkpoint) + from_config (new_autoencoder.py:224-263);
    the hub call is replaced by a local lookup (no network in the test environment)."""
    import huggingface_hub

    from eo_vae.models.new_autoencoder import EOFluxVAE
    cfg = TINY_CONFIG
    sd = make_state_dict(cfg, 15)
    _write(str(tmp_path), cfg)
    torch.save({"state_dict": sd}, os.path.join(str(tmp_path), "eo-vae.ckpt"))
    calls = []

    def fake_download(repo_id, filename, revision=None, cache_dir=None, local_files_only=False, **kw):
        calls.append((repo_id, filename, revision, local_files_only))
        return os.path.join(str(tmp_path), filename)

    monkeypatch.setattr(huggingface_hub, "hf_hub_download", fake_download)
    model = EOFluxVAE.from_pretrained("nilsleh/eo-vae", revision="main", local_files_only=True)
    assert calls == [("nilsleh/eo-vae", "model_config.yaml", "main", True), ("nilsleh/eo-vae", "eo-vae.ckpt", "main", True)]
    _same(model.state_dict(), sd)
    assert not model.training


def test_optimizer_selection_and_no_cpu_path():
    """configure_optimizers (new_autoencoder.py:549-585): plain torch Adam while the model lives on the CPU (construction /
    checkpoint handling only); the kernel-backed FusedClipAdam refuses CPU tensors instead of falling back."""
    from eo_vae.models.new_autoencoder import EOFluxVAE  # noqa: F401
    from eo_vae.optim import FusedClipAdam
    import __graft_entry__ as g
    cfg = TINY_CONFIG
    model = g._model(cfg, make_state_dict(cfg, 16), torch.device("cpu"))
    opt = model.configure_optimizers()
    opt = opt[0] if isinstance(opt, (list, tuple)) else opt
    assert type(opt) is torch.optim.Adam and opt.param_groups[0]["lr"] == model.base_lr
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError, match="CUDA"):
        FusedClipAdam([p], lr=1e-3).step(clip_norm=1.0)
    with pytest.raises(NotImplementedError):
        FusedClipAdam([p], lr=1e-3, weight_decay=0.1)
