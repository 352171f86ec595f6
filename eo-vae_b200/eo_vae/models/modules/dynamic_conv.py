"""Wavelength-conditioned dynamic input / output convolutions on the sm_100a kernels.

Interface mirror of the reference ``eo_vae/models/modules/dynamic_conv.py`` (TransformerWeightGenerator :62,
TransformerWeightGenerator_decoder :133, FactorizedWeightGenerator :186, FactorizedWeightGenerator_decoder :267,
FCResLayer :336, DynamicConv :369, DynamicConv_decoder :538): identical constructor arguments, parameter names / shapes / initialisation, ``forward`` and ``get_distillation_weight``.
The hypernetwork (sincos -> FCRes -> post-norm transformer over 128 + C + 1 tokens -> linear heads) runs as fp32
CUDA kernels (``eovae_hypernet_forward``) and its output is packed straight into the K-major tensor-core operand
of the band-mixing 3x3 implicit GEMM (``eovae_pack_dyn_weight`` -> ``eovae_conv2d``); no OIHW weight tensor is
materialised on the forward path.

Deliberate difference: the reference re-seeds torch's global RNG at import time (``torch.manual_seed(1234)``,
dynamic_conv.py:7-8).  Importing this module has no such side effect.
"""
from __future__ import annotations

import torch
import torch.nn as nn
from torch import Tensor

from ... import autograd as tape
from ... import ops
from ...settings import compute_dtype


def get_1d_sincos_pos_embed_from_grid_torch(embed_dim: int, pos: Tensor) -> Tensor:
    """[M] positions -> [M, D] (sin | cos) embedding (dynamic_conv.py:37-59).  Host/torch utility kept for API
    parity; the CUDA path evaluates the same formula inside eovae_hypernet_forward."""
    assert embed_dim % 2 == 0
    return_dev = pos.device
    omega = _omega_table(embed_dim).to(return_dev)
    ang = pos.reshape(-1).float()[:, None] * omega[None, :]
    return torch.cat([torch.sin(ang), torch.cos(ang)], dim=1)


def _omega_table(embed_dim: int) -> Tensor:
    """1 / 10000^(i / (D/2)), evaluated exactly like the reference (fp32, CPU) so the frequencies are bit-equal."""
    omega = torch.arange(embed_dim // 2, dtype=torch.float32)
    omega /= embed_dim / 2.0
    return 1.0 / 10000**omega


class FCResLayer(nn.Module):
    """x + relu(W2 relu(W1 x + b1) + b2) (dynamic_conv.py:336-366); parameter container on the CUDA path."""

    def __init__(self, linear_size: int = 128) -> None:
        super().__init__()
        self.l_size = linear_size
        self.nonlin1 = nn.ReLU(inplace=True)
        self.nonlin2 = nn.ReLU(inplace=True)
        self.w1 = nn.Linear(self.l_size, self.l_size)
        self.w2 = nn.Linear(self.l_size, self.l_size)


class TransformerWeightGenerator(nn.Module):
    """Parameter container with the reference's registration order and init (dynamic_conv.py:62-108)."""

    _decoder_head = False

    def __init__(self, input_dim: int, output_dim: int, embed_dim: int, num_heads: int = 4, num_layers: int = 1) -> None:
        super().__init__()
        layer = nn.TransformerEncoderLayer(d_model=input_dim, nhead=num_heads, activation='gelu', norm_first=False,
                                           batch_first=False, dropout=False)
        self.transformer_encoder = nn.TransformerEncoder(layer, num_layers=num_layers, enable_nested_tensor=False)
        self.fc_weight = nn.Linear(input_dim, output_dim)
        self.fc_bias = nn.Linear(input_dim, embed_dim)
        self.wt_num = 128
        self.weight_tokens = nn.Parameter(torch.empty([self.wt_num, input_dim]))
        self.bias_token = nn.Parameter(torch.empty([1, input_dim]))
        torch.nn.init.normal_(self.weight_tokens, std=0.02)
        torch.nn.init.normal_(self.bias_token, std=0.02)
        self.input_dim, self.embed_dim, self.num_heads, self.num_layers = input_dim, embed_dim, num_heads, num_layers

    def parameter_list(self, fclayer: FCResLayer) -> list:
        """The trainable tensors in kernel order (entry 0 of the kernel list, the sincos table, is not a parameter)."""
        ps = [self.weight_tokens, self.bias_token, fclayer.w1.weight, fclayer.w1.bias, fclayer.w2.weight,
              fclayer.w2.bias, self.fc_weight.weight, self.fc_weight.bias, self.fc_bias.weight, self.fc_bias.bias]
        for l in self.transformer_encoder.layers:
            ps += [l.self_attn.in_proj_weight, l.self_attn.in_proj_bias, l.self_attn.out_proj.weight,
                   l.self_attn.out_proj.bias, l.linear1.weight, l.linear1.bias, l.linear2.weight, l.linear2.bias,
                   l.norm1.weight, l.norm1.bias, l.norm2.weight, l.norm2.bias]
        return ps

    def kernel_params(self, fclayer: FCResLayer, omega: Tensor) -> list:
        """Device pointers in the order eovae_hypernet_forward expects (include/eovae.h)."""
        ps = [omega] + self.parameter_list(fclayer)
        for p in ps:
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("hypernetwork parameters must be contiguous fp32 tensors")
        return [p.detach() for p in ps]

    @property
    def ff_dim(self) -> int:
        return self.transformer_encoder.layers[0].linear1.out_features


class TransformerWeightGenerator_decoder(TransformerWeightGenerator):
    """Per-band scalar bias head (dynamic_conv.py:133-183)."""

    _decoder_head = True

    def __init__(self, input_dim: int, output_dim: int, embed_dim: int, num_heads: int = 4, num_layers: int = 1) -> None:
        super().__init__(input_dim, output_dim, embed_dim, num_heads=num_heads, num_layers=num_layers)
        self.fc_bias = nn.Linear(input_dim, 1)


class FactorizedWeightGenerator(nn.Module):
    """Low-rank head + pre-norm transformer (dynamic_conv.py:186-264): parameter container with the reference's
    registration order (transformer_encoder, fc_weight.{0,2}, fc_bias, weight_tokens, bias_token) and init.
    Runs on eovae_hypernet_factorized_forward/backward.  The reference's train-mode dropout (p = 0.1 inside the
    transformer layers) is NOT applied by the kernels: the generated kernel is deterministic in both modes."""

    _decoder_head = False
    factorized = True

    def __init__(self, input_dim: int, output_dim: int, embed_dim: int, num_heads: int = 4, num_layers: int = 2,
                 rank_ratio: int = 4) -> None:
        super().__init__()
        layer = nn.TransformerEncoderLayer(d_model=input_dim, nhead=num_heads, dim_feedforward=input_dim * 4,
                                           activation='gelu', norm_first=True, batch_first=False, dropout=0.1)
        self.transformer_encoder = nn.TransformerEncoder(layer, num_layers=num_layers, enable_nested_tensor=False)
        rank = max(32, output_dim // rank_ratio)
        self.fc_weight = nn.Sequential(nn.Linear(input_dim, rank), nn.GELU(), nn.Linear(rank, output_dim))
        self.fc_bias = nn.Linear(input_dim, embed_dim)
        self.wt_num = 128
        self.weight_tokens = nn.Parameter(torch.empty([self.wt_num, input_dim]))
        self.bias_token = nn.Parameter(torch.empty([1, input_dim]))
        torch.nn.init.normal_(self.weight_tokens, std=0.02)
        torch.nn.init.normal_(self.bias_token, std=0.02)
        self._init_head()
        self.input_dim, self.embed_dim, self.num_heads, self.num_layers = input_dim, embed_dim, num_heads, num_layers
        self.rank = rank

    def _init_head(self) -> None:
        nn.init.xavier_uniform_(self.fc_weight[0].weight)
        nn.init.zeros_(self.fc_weight[-1].weight)
        nn.init.zeros_(self.fc_weight[-1].bias)

    def parameter_list(self, fclayer: FCResLayer) -> list:
        """Trainable tensors in the order of eovae_hypernet_factorized_forward (entry 0, the sincos table, excluded)."""
        ps = [self.weight_tokens, self.bias_token, fclayer.w1.weight, fclayer.w1.bias, fclayer.w2.weight,
              fclayer.w2.bias, self.fc_weight[0].weight, self.fc_weight[0].bias, self.fc_weight[2].weight,
              self.fc_weight[2].bias, self.fc_bias.weight, self.fc_bias.bias]
        for l in self.transformer_encoder.layers:
            ps += [l.self_attn.in_proj_weight, l.self_attn.in_proj_bias, l.self_attn.out_proj.weight,
                   l.self_attn.out_proj.bias, l.linear1.weight, l.linear1.bias, l.linear2.weight, l.linear2.bias,
                   l.norm1.weight, l.norm1.bias, l.norm2.weight, l.norm2.bias]
        return ps

    kernel_params = TransformerWeightGenerator.kernel_params
    ff_dim = TransformerWeightGenerator.ff_dim


class FactorizedWeightGenerator_decoder(FactorizedWeightGenerator):
    """Per-band scalar bias head on (features + bias_token) (dynamic_conv.py:267-302)."""

    _decoder_head = True

    def __init__(self, input_dim: int, output_dim: int, embed_dim: int, num_heads: int = 4, num_layers: int = 2,
                 rank_ratio: int = 4) -> None:
        super().__init__(input_dim, output_dim, embed_dim, num_heads, num_layers, rank_ratio)
        self.fc_bias = nn.Linear(input_dim, 1)


def _xavier_linear(m: nn.Module) -> None:
    if isinstance(m, nn.Linear):
        nn.init.xavier_uniform_(m.weight)
        if m.bias is not None:
            m.bias.data.fill_(0.01)


class _DynamicBase(nn.Module):
    _decoder = False

    def __init__(self, wv_planes: int, inter_dim: int = 128, kernel_size: int = 3, stride: int = 1, padding: int = 1,
                 embed_dim: int = 128, num_layers: int = 1, num_heads: int = 4, generator_type: str = 'transformer',
                 rank_ratio: int = 4) -> None:
        super().__init__()
        if (kernel_size, stride, padding) != (3, 1, 1):
            raise NotImplementedError('dynamic conv kernels are built for kernel 3, stride 1, padding 1')
        self.kernel_size = kernel_size
        self.wv_planes = wv_planes
        self.embed_dim = embed_dim
        self._num_kernel = kernel_size * kernel_size * embed_dim
        self.inter_dim = inter_dim
        self.patch_size = (kernel_size, kernel_size)
        self.num_patches = -1
        self.stride = stride
        self.padding = padding
        self.generator_type = generator_type
        if generator_type == 'factorized':  # dynamic_conv.py:412-421, 581-590
            gen = FactorizedWeightGenerator_decoder if self._decoder else FactorizedWeightGenerator
            self.weight_generator = gen(wv_planes, self._num_kernel, embed_dim, num_heads=num_heads,
                                        num_layers=num_layers, rank_ratio=rank_ratio)
        else:
            gen = TransformerWeightGenerator_decoder if self._decoder else TransformerWeightGenerator
            self.weight_generator = gen(wv_planes, self._num_kernel, embed_dim, num_heads=num_heads,
                                        num_layers=num_layers)
        self.use_weight_standardization = False
        self.scaler = 0.1
        self.fclayer = FCResLayer(wv_planes)
        self._init_weights()
        self._omega = None

    def weight_init(self, m: nn.Module) -> None:
        _xavier_linear(m)

    def _init_weights(self) -> None:
        self.weight_generator.apply(self.weight_init)
        self.fclayer.apply(self.weight_init)

    # -- kernel plumbing ---------------------------------------------------------------------------------------
    def _omega_dev(self, device) -> Tensor:
        if self._omega is None or self._omega.device != device:
            self._omega = _omega_table(self.wv_planes).to(device)
        return self._omega

    def _generate(self, wvs: Tensor):
        """-> raw fc_weight output [C, 9E] and raw bias head output, both fp32 on device."""
        g = self.weight_generator
        dev = g.weight_tokens.device
        wvs = wvs.to(device=dev, dtype=torch.float32)
        params = g.kernel_params(self.fclayer, self._omega_dev(dev))
        if getattr(g, 'factorized', False):
            return ops.hypernet_factorized_forward(wvs, params, g.num_layers, g.input_dim, g.num_heads, g.ff_dim,
                                                   self.embed_dim, g.rank, self._decoder)[:2]
        return ops.hypernet_forward(wvs, params, g.num_layers, g.input_dim, g.num_heads, g.ff_dim, self.embed_dim,
                                    self._decoder)

    def _generate_taped(self, wvs: Tensor):
        """_generate that also returns the activation tape for _hyper_backward (training forward)."""
        g = self.weight_generator
        dev = g.weight_tokens.device
        params = g.kernel_params(self.fclayer, self._omega_dev(dev))
        if getattr(g, 'factorized', False):
            return ops.hypernet_factorized_forward(wvs.to(device=dev, dtype=torch.float32), params, g.num_layers,
                                                   g.input_dim, g.num_heads, g.ff_dim, self.embed_dim, g.rank,
                                                   self._decoder)
        return ops.hypernet_forward_taped(wvs.to(device=dev, dtype=torch.float32), params, g.num_layers, g.input_dim,
                                          g.num_heads, g.ff_dim, self.embed_dim, self._decoder)

    def _hyper_backward(self, wvs: Tensor, dw_oihw: Tensor, dbias: Tensor, bias_scale: float, tape=None) -> list:
        """Gradients of ``weight_generator.parameter_list(fclayer)`` given the generated kernel's / bias' gradient."""
        g = self.weight_generator
        dev = g.weight_tokens.device
        params = g.kernel_params(self.fclayer, self._omega_dev(dev))
        if getattr(g, 'factorized', False):
            if tape is None:
                tape = self._generate_taped(wvs)[2]
            return ops.hypernet_factorized_backward(wvs.to(device=dev, dtype=torch.float32), params, g.num_layers,
                                                    g.input_dim, g.num_heads, g.ff_dim, self.embed_dim, g.rank,
                                                    self._decoder, dw_oihw, self.scaler, dbias, bias_scale, tape)
        return ops.hypernet_backward(wvs.to(device=dev, dtype=torch.float32), params, g.num_layers, g.input_dim, g.num_heads,
                                     g.ff_dim, self.embed_dim, self._decoder, dw_oihw, self.scaler, dbias, bias_scale, tape)

    # -- eval-mode operand cache ---------------------------------------------------------------------------------
    # The generated kernel is a function of (wavelengths, hypernetwork parameters) only - not of the batch (SURVEY App. B
    # item 8) - so outside training the packed conv operand is kept per wavelength vector and rebuilt when any parameter
    # of the hypernetwork changes (autograd version counters; optimiser steps and load_state_dict advance them).
    CACHE_EVAL_OPERANDS = True

    def _wvs_key(self, wvs: Tensor):
        if wvs.is_cuda:   # no device read: identity of the tensor (pointer + version + length)
            return ('cuda', wvs.data_ptr(), wvs._version, wvs.numel(), str(wvs.dtype))
        return ('cpu',) + tuple(float(v) for v in wvs.reshape(-1).tolist())

    def _eval_operands(self, wvs: Tensor, dtype, bias_scale: float):
        """-> (packed igemm B operand, scaled bias, (wk, bias_raw)) for the non-taped forward."""
        c = wvs.numel()
        if not self.CACHE_EVAL_OPERANDS:
            wk, b_raw = self._generate(wvs)
            packed, bias, _ = ops.pack_dyn_weight(wk, b_raw, c, self.embed_dim, self._decoder, self.scaler, bias_scale, dtype, False)
            return packed, bias, (wk, b_raw)
        params = list(self.weight_generator.parameters()) + list(self.fclayer.parameters())
        key = (self._wvs_key(wvs), dtype, float(bias_scale), float(self.scaler),
               tuple(p._version for p in params), tuple(p.data_ptr() for p in params))
        cache = self.__dict__.setdefault('_operand_cache', {})
        hit = cache.get(key[0])
        if hit is not None and hit[0] == key:
            return hit[1]
        wk, b_raw = self._generate(wvs)
        packed, bias, _ = ops.pack_dyn_weight(wk, b_raw, c, self.embed_dim, self._decoder, self.scaler, bias_scale, dtype, False)
        if len(cache) >= 16:
            cache.clear()
        # a CUDA wavelength tensor is identified by pointer: keep it alive so the address cannot be recycled
        cache[key[0]] = (key, (packed, bias, (wk, b_raw)), wvs)
        return packed, bias, (wk, b_raw)

    def _get_weights(self, waves: Tensor):
        raise NotImplementedError('the CUDA path generates weights from wavelengths directly; use _generate(wvs)')


class DynamicConv(_DynamicBase):
    """Input layer: [B, C, H, W] image + [C] wavelengths (um) -> [B, embed_dim, H, W] (dynamic_conv.py:369-535)."""

    _decoder = False

    def get_distillation_weight(self, wvs_microns: Tensor):
        if tape.grad_mode():  # stage-1 weight distillation (weight_distill_train.py:190-264) trains the hypernetwork
            dev = self.weight_generator.weight_tokens.device
            return tape.DistillWeightFn.apply(self, wvs_microns.to(device=dev, dtype=torch.float32),
                                              *self.weight_generator.parameter_list(self.fclayer))
        wk, b_raw = self._generate(wvs_microns)
        _, bias, oihw = ops.pack_dyn_weight(wk, b_raw, wvs_microns.numel(), self.embed_dim, False, self.scaler,
                                            self.scaler, compute_dtype(), want_oihw=True)
        return oihw, bias

    def forward(self, img_feat: Tensor, wvs: Tensor) -> Tensor:
        c = wvs.size(0)
        if img_feat.shape[1] != c:
            raise RuntimeError(f'DynamicConv: {img_feat.shape[1]} image bands but {c} wavelengths')
        dt = compute_dtype()
        if tape.grad_mode():
            x = ops.nchw_to_act(img_feat, ops.dyn_cin_pad(c), dt)
            return tape.DynConvInFn.apply(x, self, wvs, *self.weight_generator.parameter_list(self.fclayer))
        packed, bias, _ = self._eval_operands(wvs, dt, self.scaler)
        x = ops.nchw_to_act(img_feat, ops.dyn_cin_pad(c), dt)
        return ops.conv2d(x, packed, bias, self.embed_dim, ops.CONV_3X3, algo_cin=c, gn_groups=32, gn_eps=1e-6)


class DynamicConv_decoder(_DynamicBase):
    """Output layer: [B, embed_dim, H, W] features -> [B, C, H, W] bands (dynamic_conv.py:538-710).  As in the
    reference forward the generated bias is scaled twice (0.1 * 0.1, :692-697) while ``get_distillation_weight``
    scales it once (:660)."""

    _decoder = True

    def __init__(self, wv_planes: int, inter_dim: int = 128, kernel_size: int = 3, stride: int = 1, padding: int = 1,
                 embed_dim: int = 128, num_layers: int = 2, num_heads: int = 4, generator_type: str = 'transformer',
                 rank_ratio: int = 4) -> None:
        super().__init__(wv_planes, inter_dim, kernel_size, stride, padding, embed_dim, num_layers, num_heads,
                         generator_type, rank_ratio)
        self._last = None

    def get_distillation_weight(self, wvs_microns: Tensor):
        if tape.grad_mode():
            dev = self.weight_generator.weight_tokens.device
            return tape.DistillWeightFn.apply(self, wvs_microns.to(device=dev, dtype=torch.float32),
                                              *self.weight_generator.parameter_list(self.fclayer))
        wk, b_raw = self._generate(wvs_microns)
        _, bias, oihw = ops.pack_dyn_weight(wk, b_raw, wvs_microns.numel(), self.embed_dim, True, self.scaler,
                                            self.scaler, compute_dtype(), want_oihw=True)
        return oihw, bias

    @property
    def weight(self) -> Tensor:
        """Generated [C, embed, 3, 3] kernel of the last forward (the reference stashes it at :708; read by
        ``EOFluxVAE.get_last_layer``).  Unpacked lazily so the forward path never materialises it."""
        if self._last is None:
            raise AttributeError('DynamicConv_decoder.weight is only defined after a forward pass')
        wk, b_raw, c = self._last
        return ops.pack_dyn_weight(wk, b_raw, c, self.embed_dim, True, self.scaler, 0.01, compute_dtype(), True)[2]

    def forward(self, img_feat: Tensor, waves: Tensor) -> Tensor:
        c = waves.size(0)
        self.scaler = 0.1
        x = ops.to_act(img_feat, compute_dtype())
        if tape.grad_mode():
            return tape.DynConvOutFn.apply(x, self, waves, *self.weight_generator.parameter_list(self.fclayer))
        packed, bias, (wk, b_raw) = self._eval_operands(waves, x.dtype, self.scaler * self.scaler)
        self._last = (wk, b_raw, c)
        return ops.conv2d(x, packed, bias, c, ops.CONV_3X3, out_dtype=torch.float32)
