"""Flux2-style encoder / decoder bodies on the sm_100a kernels.

Interface mirror of the reference ``eo_vae/models/model.py`` (Encoder :67-197, Decoder :223-358): same constructor
arguments, submodule names (hence ``state_dict`` keys), attributes read by callers (``use_dynamic_ops``,
``z_channels``, ``conv_in``, ``conv_out``) and forward signatures.  Activations stay NHWC / 16-bit between layers;
the NCHW fp32 <-> NHWC 16-bit conversions happen once, at the model edges.
"""
from __future__ import annotations

import torch
import torch.nn as nn
from torch import Tensor

from .. import autograd as tape
from .. import ops
from ..settings import compute_dtype
from .modules.dynamic_conv import DynamicConv, DynamicConv_decoder, _omega_table
from .modules.layers import AttnBlock, Conv2dSM100, Downsample, GroupNormSM100, ResnetBlock, Upsample


def swish(x: Tensor) -> Tensor:
    return x * torch.sigmoid(x)


def get_1d_sincos_pos_embed(embed_dim: int, pos: Tensor) -> Tensor:
    """[N] or [B, N] positions -> [B, N, D] (sin | cos) embedding (model.py:17-32).  Host/torch utility kept for API
    parity; the CUDA path evaluates the same formula inside eovae_wavelength_style_forward."""
    if pos.dim() == 1:
        pos = pos.unsqueeze(0)
    out = torch.einsum('bn,d->bnd', pos.float(), _omega_table(embed_dim).to(pos.device))
    return torch.cat([torch.sin(out), torch.cos(out)], dim=2)


class WavelengthConditioner(nn.Module):
    """Wavelength set -> global AdaIN style vector (model.py:35-64): mean over bands of the sincos embedding (wavelengths
    in micrometres, unscaled), then Linear(d, 2d) -> SiLU -> Linear(2d, d) -> SiLU -> Linear(d, d), computed by
    eovae_wavelength_style_forward/backward.  The style is a function of the wavelength vector only, so one row is
    computed and returned as an expanded [batch_size, d] view (the reference materialises the repeat)."""

    def __init__(self, embed_dim: int = 512):
        super().__init__()
        self.embed_dim = embed_dim
        self.mlp = nn.Sequential(nn.Linear(embed_dim, embed_dim * 2), nn.SiLU(), nn.Linear(embed_dim * 2, embed_dim),
                                 nn.SiLU(), nn.Linear(embed_dim, embed_dim))
        self._omega = None

    def forward(self, wvs: Tensor, batch_size: int) -> Tensor:
        if wvs.dim() != 1:
            raise RuntimeError('WavelengthConditioner: one wavelength vector per batch ([N]) on the sm_100a path')
        dev = self.mlp[0].weight.device
        if self._omega is None or self._omega.device != dev:
            self._omega = _omega_table(self.embed_dim).to(dev)
        ps = [self.mlp[0].weight, self.mlp[0].bias, self.mlp[2].weight, self.mlp[2].bias, self.mlp[4].weight,
              self.mlp[4].bias]
        wvs = wvs.to(device=dev, dtype=torch.float32)
        if tape.grad_mode():
            style = tape.WavelengthStyleFn.apply(wvs, self._omega, *ps)
        else:
            style = ops.wavelength_style_forward(wvs, [self._omega] + [p.detach() for p in ps], self.embed_dim)[0]
        return style.expand(batch_size, -1)


def _split_dynamic_kwargs(dynamic_conv_kwargs):
    """-> (use_adain, wv_planes, inter_dim, kwargs passed on to the dynamic layer)   (model.py:94-104, 248-251, 299-308)"""
    kw = dict(dynamic_conv_kwargs) if dynamic_conv_kwargs else {}
    use_adain = bool(kw.pop('use_adain', False))
    kw.pop('mode', 'conv')
    return use_adain, kw.pop('wv_planes', 128), kw.pop('inter_dim', 128), kw


class Encoder(nn.Module):
    def __init__(self, resolution: int, in_channels: int, ch: int, ch_mult: list[int], num_res_blocks: int,
                 z_channels: int, use_dynamic_ops: bool = False, dynamic_conv_kwargs: dict = None):
        super().__init__()
        self.ch = ch
        self.num_resolutions = len(ch_mult)
        self.num_res_blocks = num_res_blocks
        self.resolution = resolution
        self.in_channels = in_channels
        self.z_channels = z_channels
        self.use_dynamic_ops = use_dynamic_ops
        self.use_adain = False
        self.cond_dim = None
        if use_dynamic_ops:
            self.use_adain, wv_planes, inter_dim, rest = _split_dynamic_kwargs(dynamic_conv_kwargs)
            if self.use_adain:
                self.cond_dim = 512
                self.conditioner = WavelengthConditioner(embed_dim=self.cond_dim)
            self.conv_in = DynamicConv(wv_planes=wv_planes, inter_dim=inter_dim, kernel_size=3, stride=1, padding=1,
                                       embed_dim=ch, **rest)
        else:
            self.conv_in = Conv2dSM100(in_channels, ch, kernel_size=3, stride=1, padding=1)
        in_ch_mult = (1,) + tuple(ch_mult)
        self.in_ch_mult = in_ch_mult
        self.down = nn.ModuleList()
        block_in = ch
        for lvl in range(self.num_resolutions):
            stage = nn.Module()
            stage.block = nn.ModuleList()
            stage.attn = nn.ModuleList()
            block_in = ch * in_ch_mult[lvl]
            block_out = ch * ch_mult[lvl]
            for _ in range(num_res_blocks):
                stage.block.append(ResnetBlock(in_channels=block_in, out_channels=block_out, cond_dim=self.cond_dim))
                block_in = block_out
            if lvl != self.num_resolutions - 1:
                stage.downsample = Downsample(block_in)
            self.down.append(stage)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in, cond_dim=self.cond_dim)
        self.mid.attn_1 = AttnBlock(block_in)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in, cond_dim=self.cond_dim)
        self.norm_out = GroupNormSM100(num_groups=32, num_channels=block_in, eps=1e-6, affine=True)
        self.conv_out = Conv2dSM100(block_in, 2 * z_channels, kernel_size=3, stride=1, padding=1)
        self.quant_conv = Conv2dSM100(2 * z_channels, 2 * z_channels, 1)

    def moments_nhwc(self, x: Tensor, wvs: Tensor = None) -> Tensor:
        """fp32 moments, logical [B, 2z, H/8, W/8] with NHWC storage (what the fused latent kernels consume)."""
        emb = None
        if self.use_dynamic_ops:
            assert wvs is not None, 'wvs must be provided for Dynamic Encoder'
            if self.use_adain:
                emb = self.conditioner(wvs, x.shape[0])
            h = self.conv_in(x, wvs)
        else:
            h = self.conv_in(ops.nchw_to_act(x, (x.shape[1] + 15) // 16 * 16, compute_dtype()))
        for lvl, stage in enumerate(self.down):
            for block in stage.block:
                h = block(h, emb)
            if lvl != self.num_resolutions - 1:
                h = stage.downsample(h)
        h = self.mid.block_1(h, emb)
        h = self.mid.attn_1(h)
        h = self.mid.block_2(h, emb)
        h = self.conv_out(self.norm_out(h, silu=True))
        return self.quant_conv(h, out_dtype=torch.float32)

    def forward(self, x: Tensor, wvs: Tensor = None) -> Tensor:
        return tape.act_to_nchw_f32(self.moments_nhwc(x, wvs))

    def load_flux_weights(self, state_dict, strict=True):
        own = self.state_dict()
        skip = (['conv_in'] if self.use_dynamic_ops else []) + (['conditioner', 'emb_proj'] if self.use_adain else [])
        for name, param in state_dict.items():
            if any(s in name for s in skip):
                continue
            if name not in own:
                if strict:
                    raise KeyError(f'Unexpected key {name} in state_dict')
                continue
            own[name].copy_(param)


class Decoder(nn.Module):
    def __init__(self, ch: int, out_ch: int, ch_mult: list[int], num_res_blocks: int, resolution: int, z_channels: int,
                 use_dynamic_ops: bool = False, dynamic_conv_kwargs: dict = None):
        super().__init__()
        self.post_quant_conv = Conv2dSM100(z_channels, z_channels, 1)
        self.ch = ch
        self.num_resolutions = len(ch_mult)
        self.num_res_blocks = num_res_blocks
        self.z_channels = z_channels
        self.resolution = resolution
        self.use_dynamic_ops = use_dynamic_ops
        self.use_adain = False
        self.cond_dim = None
        if use_dynamic_ops:
            self.use_adain, wv_planes, inter_dim, rest = _split_dynamic_kwargs(dynamic_conv_kwargs)
            if self.use_adain:
                self.cond_dim = 512
                self.conditioner = WavelengthConditioner(embed_dim=self.cond_dim)
        block_in = ch * ch_mult[self.num_resolutions - 1]
        curr_res = resolution // 2 ** (self.num_resolutions - 1)
        self.z_shape = (1, z_channels, curr_res, curr_res)
        self.conv_in = Conv2dSM100(z_channels, block_in, kernel_size=3, stride=1, padding=1)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in, cond_dim=self.cond_dim)
        self.mid.attn_1 = AttnBlock(block_in)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in, cond_dim=self.cond_dim)
        self.up = nn.ModuleList()
        for lvl in reversed(range(self.num_resolutions)):
            stage = nn.Module()
            stage.block = nn.ModuleList()
            stage.attn = nn.ModuleList()
            block_out = ch * ch_mult[lvl]
            for _ in range(num_res_blocks + 1):
                stage.block.append(ResnetBlock(in_channels=block_in, out_channels=block_out, cond_dim=self.cond_dim))
                block_in = block_out
            if lvl != 0:
                stage.upsample = Upsample(block_in)
            self.up.insert(0, stage)
        self.norm_out = GroupNormSM100(num_groups=32, num_channels=block_in, eps=1e-6, affine=True)
        if use_dynamic_ops:
            self.conv_out = DynamicConv_decoder(wv_planes=wv_planes, inter_dim=inter_dim, kernel_size=3, stride=1,
                                                padding=1, embed_dim=block_in, **rest)
        else:
            self.conv_out = Conv2dSM100(block_in, out_ch, kernel_size=3, stride=1, padding=1)

    def forward_act(self, z: Tensor, wvs: Tensor = None) -> Tensor:
        """z: activation (NHWC 16-bit) or any [B, z, h, w] tensor -> fp32 NHWC-stored reconstruction."""
        h = self.conv_in(self.post_quant_conv(z), gn_next=True)
        emb = None
        if self.use_dynamic_ops and self.use_adain:
            assert wvs is not None
            emb = self.conditioner(wvs, z.shape[0])
        h = self.mid.block_1(h, emb)
        h = self.mid.attn_1(h)
        h = self.mid.block_2(h, emb)
        for lvl in reversed(range(self.num_resolutions)):
            for block in self.up[lvl].block:
                h = block(h, emb)
            if lvl != 0:
                h = self.up[lvl].upsample(h)
        h = self.norm_out(h, silu=True)
        if self.use_dynamic_ops:
            assert wvs is not None, 'wvs must be provided for Dynamic Decoder'
            return self.conv_out(h, wvs)
        return self.conv_out(h, out_dtype=torch.float32)

    def forward(self, z: Tensor, wvs: Tensor = None) -> Tensor:
        return tape.act_to_nchw_f32(self.forward_act(z, wvs))

    def load_flux_weights(self, state_dict, strict=True):
        own = self.state_dict()
        skip = (['conv_out'] if self.use_dynamic_ops else []) + (['conditioner', 'emb_proj'] if self.use_adain else [])
        for name, param in state_dict.items():
            if any(s in name for s in skip) or name not in own:
                continue
            own[name].copy_(param)
