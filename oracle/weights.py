"""TEST INFRASTRUCTURE ONLY - deterministic, torch-RNG-independent parameter sets for parity tests.

Parity is checked with the reference, the oracle and the CUDA path all loaded from the *same*
``state_dict``.  95.5 M parameters cannot be committed as fixtures, so every tensor is regenerated from
(seed, key name) with numpy's Philox generator: identical bits in this container and on the GPU box
(same image, same numpy).  The key list / shapes restate the reference checkpoint layout
(``eo_vae/models/model.py:67-165, 223-322``, ``dynamic_conv.py:62-108, 372-436``,
``new_autoencoder.py:125``) and are verified against the real reference ``state_dict`` in
``tests/test_oracle_vs_reference.py``.
"""
from __future__ import annotations

import zlib
from collections import OrderedDict

import numpy as np
import torch

# configs/eo-vae.yaml:33-57 of the reference
FULL_CONFIG = dict(resolution=256, ch=128, ch_mult=(1, 2, 4, 4), num_res_blocks=2, z_channels=32,
                   hyper_layers=4, wv_planes=256, hyper_heads=4)
# Same topology at toy width: every op of the path, seconds on CPU.
TINY_CONFIG = dict(resolution=64, ch=32, ch_mult=(1, 2, 2), num_res_blocks=1, z_channels=8,
                   hyper_layers=1, wv_planes=64, hyper_heads=4)

# generator_type='factorized' (configs/finetune_consistency_factor.yaml:50-73: num_layers 4, rank_ratio 2) at toy width
TINY_FACTORIZED_CONFIG = dict(TINY_CONFIG, hyper_layers=2, generator_type="factorized", rank_ratio=2)

# use_adain=True in dynamic_conv_kwargs (model.py:96-100): WavelengthConditioner + emb_proj in every ResnetBlock
TINY_ADAIN_CONFIG = dict(TINY_CONFIG, use_adain=True)

# eo_vae/datasets/terramesh_datamodule.py:18-50 (micrometres)
WAVELENGTHS = {
    "S2RGB": [0.665, 0.56, 0.49],
    "S1RTC": [5.4, 5.6],
    "S2L2A": [0.443, 0.490, 0.560, 0.665, 0.705, 0.740, 0.783, 0.842, 0.865, 1.610, 2.190, 0.945],
    "S2L1C": [0.443, 0.490, 0.560, 0.665, 0.705, 0.740, 0.783, 0.842, 0.865, 0.945, 1.375, 1.610, 2.190],
}


_ADAIN = [False]  # set by state_dict_spec while it walks a use_adain config (layers.py:68-76: emb_proj after conv1)


def _conditioner(spec, p, d=512):
    """WavelengthConditioner.mlp (model.py:42-48)."""
    for i, (o, k) in ((0, (2 * d, d)), (2, (d, 2 * d)), (4, (d, d))):
        spec[f"{p}.mlp.{i}.weight"] = (o, k)
        spec[f"{p}.mlp.{i}.bias"] = (o,)


def _resblock(spec, p, cin, cout):
    spec[p + ".norm1.weight"] = (cin,)
    spec[p + ".norm1.bias"] = (cin,)
    spec[p + ".conv1.weight"] = (cout, cin, 3, 3)
    spec[p + ".conv1.bias"] = (cout,)
    if _ADAIN[0]:
        spec[p + ".emb_proj.weight"] = (2 * cout, 512)
        spec[p + ".emb_proj.bias"] = (2 * cout,)
    spec[p + ".norm2.weight"] = (cout,)
    spec[p + ".norm2.bias"] = (cout,)
    spec[p + ".conv2.weight"] = (cout, cout, 3, 3)
    spec[p + ".conv2.bias"] = (cout,)
    if cin != cout:
        spec[p + ".nin_shortcut.weight"] = (cout, cin, 1, 1)
        spec[p + ".nin_shortcut.bias"] = (cout,)


def _attn(spec, p, c):
    spec[p + ".norm.weight"] = (c,)
    spec[p + ".norm.bias"] = (c,)
    for n in ("q", "k", "v", "proj_out"):
        spec[f"{p}.{n}.weight"] = (c, c, 1, 1)
        spec[f"{p}.{n}.bias"] = (c,)


def _hypernet_factorized(spec, p, d, embed, layers, decoder, rank_ratio):
    """FactorizedWeightGenerator(_decoder) registration order (dynamic_conv.py:186-236, 267-285): transformer layers
    (ff = 4 d), fc_weight.{0,2}, fc_bias; then fclayer."""
    g = p + ".weight_generator"
    rank = max(32, 9 * embed // rank_ratio)
    spec[g + ".weight_tokens"] = (128, d)   # direct parameters precede submodules in a state_dict
    spec[g + ".bias_token"] = (1, d)
    for i in range(layers):
        l = f"{g}.transformer_encoder.layers.{i}"
        spec[l + ".self_attn.in_proj_weight"] = (3 * d, d)
        spec[l + ".self_attn.in_proj_bias"] = (3 * d,)
        spec[l + ".self_attn.out_proj.weight"] = (d, d)
        spec[l + ".self_attn.out_proj.bias"] = (d,)
        spec[l + ".linear1.weight"] = (4 * d, d)
        spec[l + ".linear1.bias"] = (4 * d,)
        spec[l + ".linear2.weight"] = (d, 4 * d)
        spec[l + ".linear2.bias"] = (d,)
        spec[l + ".norm1.weight"] = (d,)
        spec[l + ".norm1.bias"] = (d,)
        spec[l + ".norm2.weight"] = (d,)
        spec[l + ".norm2.bias"] = (d,)
    spec[g + ".fc_weight.0.weight"] = (rank, d)
    spec[g + ".fc_weight.0.bias"] = (rank,)
    spec[g + ".fc_weight.2.weight"] = (9 * embed, rank)
    spec[g + ".fc_weight.2.bias"] = (9 * embed,)
    spec[g + ".fc_bias.weight"] = (1 if decoder else embed, d)
    spec[g + ".fc_bias.bias"] = (1 if decoder else embed,)
    for w in ("w1", "w2"):
        spec[f"{p}.fclayer.{w}.weight"] = (d, d)
        spec[f"{p}.fclayer.{w}.bias"] = (d,)


def _hypernet(spec, p, d, embed, layers, decoder, cfg=None):
    if cfg is not None and cfg.get("generator_type", "transformer") == "factorized":
        return _hypernet_factorized(spec, p, d, embed, layers, decoder, cfg.get("rank_ratio", 4))
    g = p + ".weight_generator"
    spec[g + ".weight_tokens"] = (128, d)
    spec[g + ".bias_token"] = (1, d)
    for i in range(layers):
        l = f"{g}.transformer_encoder.layers.{i}"
        spec[l + ".self_attn.in_proj_weight"] = (3 * d, d)
        spec[l + ".self_attn.in_proj_bias"] = (3 * d,)
        spec[l + ".self_attn.out_proj.weight"] = (d, d)
        spec[l + ".self_attn.out_proj.bias"] = (d,)
        spec[l + ".linear1.weight"] = (2048, d)
        spec[l + ".linear1.bias"] = (2048,)
        spec[l + ".linear2.weight"] = (d, 2048)
        spec[l + ".linear2.bias"] = (d,)
        spec[l + ".norm1.weight"] = (d,)
        spec[l + ".norm1.bias"] = (d,)
        spec[l + ".norm2.weight"] = (d,)
        spec[l + ".norm2.bias"] = (d,)
    spec[g + ".fc_weight.weight"] = (9 * embed, d)
    spec[g + ".fc_weight.bias"] = (9 * embed,)
    spec[g + ".fc_bias.weight"] = (1 if decoder else embed, d)
    spec[g + ".fc_bias.bias"] = (1 if decoder else embed,)
    for w in ("w1", "w2"):
        spec[f"{p}.fclayer.{w}.weight"] = (d, d)
        spec[f"{p}.fclayer.{w}.bias"] = (d,)


def state_dict_spec(cfg: dict) -> "OrderedDict[str, tuple]":
    """(key -> shape) in the reference's registration order."""
    ch, mult, nrb, zc = cfg["ch"], tuple(cfg["ch_mult"]), cfg["num_res_blocks"], cfg["z_channels"]
    d, hl = cfg["wv_planes"], cfg["hyper_layers"]
    nres = len(mult)
    spec: OrderedDict[str, tuple] = OrderedDict()
    _ADAIN[0] = bool(cfg.get("use_adain", False))
    # ---- encoder (model.py:67-165)
    if _ADAIN[0]:
        _conditioner(spec, "encoder.conditioner")
    _hypernet(spec, "encoder.conv_in", d, ch, hl, decoder=False, cfg=cfg)
    in_mult = (1,) + mult
    block_in = ch
    for lvl in range(nres):
        block_in = ch * in_mult[lvl]
        block_out = ch * mult[lvl]
        for b in range(nrb):
            _resblock(spec, f"encoder.down.{lvl}.block.{b}", block_in, block_out)
            block_in = block_out
        if lvl != nres - 1:
            spec[f"encoder.down.{lvl}.downsample.conv.weight"] = (block_in, block_in, 3, 3)
            spec[f"encoder.down.{lvl}.downsample.conv.bias"] = (block_in,)
    _resblock(spec, "encoder.mid.block_1", block_in, block_in)
    _attn(spec, "encoder.mid.attn_1", block_in)
    _resblock(spec, "encoder.mid.block_2", block_in, block_in)
    spec["encoder.norm_out.weight"] = (block_in,)
    spec["encoder.norm_out.bias"] = (block_in,)
    spec["encoder.conv_out.weight"] = (2 * zc, block_in, 3, 3)
    spec["encoder.conv_out.bias"] = (2 * zc,)
    spec["encoder.quant_conv.weight"] = (2 * zc, 2 * zc, 1, 1)
    spec["encoder.quant_conv.bias"] = (2 * zc,)
    # ---- decoder (model.py:223-322)
    spec["decoder.post_quant_conv.weight"] = (zc, zc, 1, 1)
    spec["decoder.post_quant_conv.bias"] = (zc,)
    if _ADAIN[0]:
        _conditioner(spec, "decoder.conditioner")
    block_in = ch * mult[-1]
    spec["decoder.conv_in.weight"] = (block_in, zc, 3, 3)
    spec["decoder.conv_in.bias"] = (block_in,)
    _resblock(spec, "decoder.mid.block_1", block_in, block_in)
    _attn(spec, "decoder.mid.attn_1", block_in)
    _resblock(spec, "decoder.mid.block_2", block_in, block_in)
    up_specs = {}
    for lvl in reversed(range(nres)):
        s: OrderedDict[str, tuple] = OrderedDict()
        block_out = ch * mult[lvl]
        for b in range(nrb + 1):
            _resblock(s, f"decoder.up.{lvl}.block.{b}", block_in, block_out)
            block_in = block_out
        if lvl != 0:
            s[f"decoder.up.{lvl}.upsample.conv.weight"] = (block_in, block_in, 3, 3)
            s[f"decoder.up.{lvl}.upsample.conv.bias"] = (block_in,)
        up_specs[lvl] = s
    for lvl in range(nres):  # ModuleList.insert(0, ...) => ascending index order in state_dict
        spec.update(up_specs[lvl])
    spec["decoder.norm_out.weight"] = (block_in,)
    spec["decoder.norm_out.bias"] = (block_in,)
    _hypernet(spec, "decoder.conv_out", d, block_in, hl, decoder=True, cfg=cfg)
    # ---- latent BatchNorm (new_autoencoder.py:125)
    spec["bn.running_mean"] = (4 * zc,)
    spec["bn.running_var"] = (4 * zc,)
    spec["bn.num_batches_tracked"] = ()
    return spec


def _rng(seed: int, key: str) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=[seed, zlib.crc32(key.encode())]))


def make_state_dict(cfg: dict, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """Deterministic 'looks-trained' parameters: fan-in scaled weights, non-trivial affine terms and
    non-trivial BN running statistics (so that every term of the path is exercised)."""
    out: OrderedDict[str, torch.Tensor] = OrderedDict()
    for key, shape in state_dict_spec(cfg).items():
        g = _rng(seed, key)
        if key == "bn.num_batches_tracked":
            out[key] = torch.tensor(7, dtype=torch.long)
            continue
        if key == "bn.running_mean":
            a = 0.3 * g.standard_normal(shape)
        elif key == "bn.running_var":
            a = 0.5 + g.random(shape)
        elif key.endswith(("weight_tokens", "bias_token")):
            a = 0.02 * g.standard_normal(shape)
        elif ".norm" in key and key.endswith(".weight") and len(shape) == 1:
            a = 1.0 + 0.1 * g.standard_normal(shape)
        elif key.endswith("emb_proj.bias"):          # [scale | shift]: scale around the identity init of layers.py:72-76
            a = 0.05 * g.standard_normal(shape)
            a[: shape[0] // 2] += 1.0
        elif key.endswith((".bias", "in_proj_bias")):
            a = 0.05 * g.standard_normal(shape)
        else:
            fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else shape[0]
            a = g.standard_normal(shape) / np.sqrt(fan_in)
        out[key] = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    return out


def synthetic_patches(batch: int, bands: int, size: int, seed: int = 1234) -> torch.Tensor:
    """z-scored synthetic patches: N(0,1) clipped to the range the reference documents
    (eo_vae/datasets/findings.md:229-242: min ~ -2, max ~ +5.7)."""
    g = _rng(seed, f"x{batch}x{bands}x{size}")
    a = g.standard_normal((batch, bands, size, size), dtype=np.float32)
    return torch.from_numpy(np.clip(a, -2.0, 6.0))
