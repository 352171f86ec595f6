"""EOFluxVAE on the sm_100a kernels: drop-in for ``eo_vae.models.new_autoencoder.EOFluxVAE``.

Interface mirror of the reference ``eo_vae/models/new_autoencoder.py:64-738``: constructor signature (:76-100),
``encode`` :418, ``decode`` :423, ``decode_raw`` :431, ``forward`` :447, ``encode_spatial_normalized`` :481,
``decode_spatial_normalized`` :505, ``reconstruct`` :725, ``encode_to_latent`` :731, ``from_config`` :188,
``from_pretrained`` :224, the three checkpoint formats of ``_load_checkpoint`` :295, ``configure_optimizers`` :549,
``training_step`` :587, ``validation_step`` :692, ``get_last_layer`` :718 and the ``state_dict`` layout
(``encoder.*``, ``decoder.*``, ``bn.running_mean|running_var|num_batches_tracked``).

The latent glue is fused: pixel-unshuffle -> BatchNorm2d(eval) -> pixel-shuffle collapses to one per-(channel, row
parity, column parity) affine applied while the posterior mean is read (``eovae_latent_norm``); the inverse is fused
with the NCHW->NHWC / fp32->16-bit conversion of the decoder input (``eovae_latent_denorm``).
"""
from __future__ import annotations

import math
import os
import random
from typing import Any

import torch
import torch.nn.functional as F
from torch import Tensor
from torch.optim import Optimizer
from torch.optim.lr_scheduler import LambdaLR

from .. import autograd as tape
from .. import ops
from .._lightning import LightningModule
from ..settings import compute_dtype
from .model import Decoder, Encoder
from .modules.distributions import DiagonalGaussianDistribution


def get_cosine_schedule_with_warmup(optimizer: Optimizer, num_warmup_steps: int, num_training_steps: int, base_lr: float,
                                    final_lr: float, num_cycles: float = 0.5) -> LambdaLR:
    """Linear warm-up then cosine decay from base_lr to final_lr (reference :36-56)."""

    def scale(step: int) -> float:
        if step < num_warmup_steps:
            return step / max(1, num_warmup_steps)
        t = (step - num_warmup_steps) / max(1, num_training_steps - num_warmup_steps)
        cos = 0.5 * (1.0 + math.cos(math.pi * num_cycles * 2.0 * t))
        return ((base_lr - final_lr) * cos + final_lr) / base_lr

    return LambdaLR(optimizer, scale)


def _clip_and_step(opt, clip_grad) -> None:
    """clip_grad_norm_(params, clip_grad) + opt.step() (:650-657); one fused pass when the optimiser supports it
    (eo_vae.optim.FusedClipAdam, possibly behind a LightningOptimizer wrapper)."""
    core = getattr(opt, 'optimizer', opt)
    if getattr(core, 'fused_clip', False):
        opt.step(clip_norm=clip_grad or None)
        return
    if clip_grad:
        torch.nn.utils.clip_grad_norm_(opt.param_groups[0]['params'], clip_grad)
    opt.step()


def _unshuffle2(z: Tensor) -> Tensor:
    """'c (i pi) (j pj) -> (c pi pj) i j', pi = pj = 2 (index permutation only)."""
    *lead, c, h, w = z.shape
    z = z.reshape(*lead, c, h // 2, 2, w // 2, 2)
    nd = z.dim()
    return z.permute(*range(nd - 5), nd - 5, nd - 3, nd - 1, nd - 4, nd - 2).reshape(*lead, c * 4, h // 2, w // 2)


def _shuffle2(z: Tensor) -> Tensor:
    """'(c pi pj) i j -> c (i pi) (j pj)'."""
    *lead, c4, h, w = z.shape
    c = c4 // 4
    z = z.reshape(*lead, c, 2, 2, h, w)
    nd = z.dim()
    return z.permute(*range(nd - 5), nd - 5, nd - 2, nd - 4, nd - 1, nd - 3).reshape(*lead, c, h * 2, w * 2)


def _load_yaml_config(path: str) -> dict:
    try:
        from omegaconf import OmegaConf  # type: ignore
        return OmegaConf.to_container(OmegaConf.load(path), resolve=True)
    except ImportError:
        import json
        import re

        import yaml
        with open(path) as f:
            data = json.load(f) if path.endswith('.json') else yaml.safe_load(f)

        def lookup(root, dotted):
            cur = root
            for part in dotted.split('.'):
                cur = cur[part]
            return cur

        def resolve(node, root):
            if isinstance(node, dict):
                return {k: resolve(v, root) for k, v in node.items()}
            if isinstance(node, list):
                return [resolve(v, root) for v in node]
            if isinstance(node, str):
                m = re.fullmatch(r'\$\{([\w.]+)\}', node)
                if m:
                    return resolve(lookup(root, m.group(1)), root)
            return node

        return resolve(data, data)


class EOFluxVAE(LightningModule):
    def __init__(self, encoder: torch.nn.Module, decoder: torch.nn.Module, loss_fn: torch.nn.Module,
                 ckpt_path: str | None = None, ignore_keys: list[str] | None = None, freeze_body: bool = True,
                 base_lr: float = 1e-4, final_lr: float | None = None, warmup_epochs: int | None = None,
                 decay_end_epoch: int | None = None, clip_grad: float | None = None, p_prior: float = 0.0,
                 p_prior_s: float = 0.0, anisotropic: bool = False, latent_noise_p: float = 0.0, noise_tau: float = 0.8,
                 image_key: str = 'image') -> None:
        super().__init__()
        self.encoder = encoder
        self.decoder = decoder
        self.loss_fn = loss_fn
        self.image_key = image_key
        self.base_lr = base_lr
        self.final_lr = final_lr
        self.warmup_epochs = warmup_epochs
        self.decay_end_epoch = decay_end_epoch
        self.clip_grad = clip_grad
        self.p_prior = p_prior
        self.p_prior_s = p_prior_s
        self.anisotropic = anisotropic
        self.latent_noise_p = latent_noise_p
        self.noise_tau = noise_tau
        self.ps = [2, 2]
        self.bn_eps = 1e-4
        self.bn = torch.nn.BatchNorm2d(math.prod(self.ps) * encoder.z_channels, affine=False, track_running_stats=True)
        self.automatic_optimization = False
        self.freeze_body = freeze_body
        if self.freeze_body:
            self._freeze_body()
        if ckpt_path:
            self._load_checkpoint(ckpt_path, ignore_keys or [])

    # ------------------------------------------------------------------------------------------- construction
    @staticmethod
    def _read_config_file(config_path: str) -> dict[str, Any]:
        if not os.path.exists(config_path):
            raise FileNotFoundError(f'Config file not found: {config_path}')
        data = _load_yaml_config(config_path)
        if not isinstance(data, dict):
            raise ValueError('Model config must deserialize to a dictionary')
        return data

    @staticmethod
    def _extract_model_sections(config: dict[str, Any]):
        model_cfg = config.get('model', config)
        if not isinstance(model_cfg, dict):
            raise ValueError('Invalid config: `model` section must be a dictionary')
        if 'encoder' not in model_cfg or 'decoder' not in model_cfg:
            raise ValueError('Invalid config: expected `encoder` and `decoder` sections')
        enc = {k: v for k, v in model_cfg['encoder'].items() if k != '_target_'}
        dec = {k: v for k, v in model_cfg['decoder'].items() if k != '_target_'}
        keys = ('freeze_body', 'base_lr', 'final_lr', 'warmup_epochs', 'decay_end_epoch', 'clip_grad', 'p_prior',
                'p_prior_s', 'anisotropic', 'latent_noise_p', 'noise_tau', 'image_key')
        return enc, dec, {k: model_cfg[k] for k in keys if k in model_cfg}

    @classmethod
    def from_config(cls, config_path: str, ckpt_path: str | None = None, *, loss_fn: torch.nn.Module | None = None,
                    freeze_body: bool | None = None, ignore_keys: list[str] | None = None,
                    device: str | torch.device | None = None, eval_mode: bool = True) -> 'EOFluxVAE':
        enc_cfg, dec_cfg, vae_kwargs = cls._extract_model_sections(cls._read_config_file(config_path))
        if freeze_body is not None:
            vae_kwargs['freeze_body'] = freeze_body
        model = cls(encoder=Encoder(**enc_cfg), decoder=Decoder(**dec_cfg),
                    loss_fn=loss_fn if loss_fn is not None else torch.nn.Identity(),
                    freeze_body=vae_kwargs.pop('freeze_body', False), **vae_kwargs)
        if ckpt_path:
            model._load_checkpoint(ckpt_path, ignore_keys or [])
        if device is not None:
            model = model.to(device)
        if eval_mode:
            model.eval()
        return model

    @classmethod
    def from_pretrained(cls, repo_id: str, *, ckpt_filename: str = 'eo-vae.ckpt',
                        config_filename: str = 'model_config.yaml', revision: str | None = None,
                        cache_dir: str | None = None, local_files_only: bool = False,
                        loss_fn: torch.nn.Module | None = None, freeze_body: bool | None = None,
                        ignore_keys: list[str] | None = None, device: str | torch.device | None = None,
                        eval_mode: bool = True) -> 'EOFluxVAE':
        try:
            from huggingface_hub import hf_hub_download
        except ImportError as exc:
            raise ImportError('huggingface_hub is required for from_pretrained') from exc
        common = dict(repo_id=repo_id, revision=revision, cache_dir=cache_dir, local_files_only=local_files_only)
        config_path = hf_hub_download(filename=config_filename, **common)
        ckpt_path = hf_hub_download(filename=ckpt_filename, **common)
        return cls.from_config(config_path=config_path, ckpt_path=ckpt_path, loss_fn=loss_fn, freeze_body=freeze_body,
                               ignore_keys=ignore_keys, device=device, eval_mode=eval_mode)

    def _freeze_body(self) -> None:
        for p in list(self.encoder.parameters()) + list(self.decoder.parameters()):
            p.requires_grad = False
        if self.encoder.use_dynamic_ops:
            for p in self.encoder.conv_in.parameters():
                p.requires_grad = True
        if self.decoder.use_dynamic_ops:
            for p in self.decoder.conv_out.parameters():
                p.requires_grad = True

    def _load_checkpoint(self, path: str, ignore_keys: list[str]) -> None:
        """Flux AE ``.safetensors`` (body only), distilled ``.pt`` (dynamic layers only) or full ``.ckpt``."""
        if not os.path.exists(path):
            print(f'Checkpoint not found: {path}')
            return
        if path.endswith('.pt'):
            ckpt = torch.load(path, map_location='cpu')
            if 'encoder_conv_in_state_dict' in ckpt or 'decoder_conv_out_state_dict' in ckpt:
                self._load_distilled_checkpoint(ckpt)
                return
        if path.endswith('.safetensors'):
            from safetensors import safe_open
            sd = {}
            with safe_open(path, framework='pt', device='cpu') as f:
                for k in f.keys():
                    sd[k] = f.get_tensor(k)
        else:
            sd = torch.load(path, map_location='cpu')
            sd = sd.get('state_dict', sd)

        def is_static_edge(k: str) -> bool:
            dyn = (self.encoder.use_dynamic_ops and 'encoder.conv_in' in k) or \
                  (self.decoder.use_dynamic_ops and 'decoder.conv_out' in k)
            return dyn and 'weight_generator' not in k and 'fclayer' not in k

        sd = {k: v for k, v in sd.items()
              if not is_static_edge(k) and not any(k.startswith(ik) for ik in ignore_keys)}
        missing, unexpected = self.load_state_dict(sd, strict=False)
        self._verify_loading(missing, unexpected, ignore_keys)

    def _load_distilled_checkpoint(self, ckpt: dict) -> None:
        if self.encoder.use_dynamic_ops and ckpt.get('encoder_conv_in_state_dict'):
            self.encoder.conv_in.load_state_dict(ckpt['encoder_conv_in_state_dict'])
        if self.decoder.use_dynamic_ops and ckpt.get('decoder_conv_out_state_dict'):
            self.decoder.conv_out.load_state_dict(ckpt['decoder_conv_out_state_dict'])

    def _verify_loading(self, missing_keys: list[str], unexpected_keys: list[str], ignore_keys: list[str]) -> None:
        allowed = list(ignore_keys)
        if self.encoder.use_dynamic_ops:
            allowed.append('encoder.conv_in')
        if self.decoder.use_dynamic_ops:
            allowed.append('decoder.conv_out')
        critical = [k for k in missing_keys if not any(k.startswith(p) for p in allowed)]
        if critical:
            raise RuntimeError(f'Critical weights missing from checkpoint:\n{critical[:20]}...\n'
                               f'Total: {len(critical)} missing keys')

    # ------------------------------------------------------------------------------------------- forward path
    def _moments(self, x: Tensor, wvs: Tensor) -> Tensor:
        return self.encoder.moments_nhwc(x, wvs)

    def encode(self, x: Tensor, wvs: Tensor) -> DiagonalGaussianDistribution:
        return DiagonalGaussianDistribution(self.encoder(x, wvs))

    def _decoder_input(self, z_packed: Tensor) -> Tensor:
        """packed normalised latent [B, 4z, h, w] -> inverse BN (running stats, eps 1e-4) -> unshuffle -> NHWC act."""
        if tape.grad_mode() and z_packed.requires_grad:
            # training: per-channel affine + index permutation on the (tiny) latent stay torch tensor glue so the tape
            # sees them; everything from post_quant_conv on is kernels again
            return ops.to_act(_shuffle2(self._inv_normalize_latent(z_packed)), compute_dtype())
        z_spatial = _shuffle2(z_packed)  # index permutation; the affine is applied per (c, parity) in the kernel
        return ops.latent_denorm(z_spatial, self.bn.running_mean, self.bn.running_var, self.bn_eps, compute_dtype())

    def decode(self, z: Tensor, wvs: Tensor) -> Tensor:
        self.bn.eval()
        return tape.act_to_nchw_f32(self.decoder.forward_act(self._decoder_input(z), wvs))

    def decode_raw(self, z: Tensor, wvs: Tensor) -> Tensor:
        return self.decoder(z, wvs)

    def noising(self, x: Tensor) -> Tensor:
        sigma = self.noise_tau * torch.rand((x.size(0),) + (1,) * (len(x.shape) - 1), device=x.device)
        return x + sigma * torch.randn_like(x)

    def forward(self, x: Tensor, wvs: Tensor, sample_posterior: bool = True, scale=None, angle: int | None = None):
        moments = self._moments(x, wvs)
        posterior = DiagonalGaussianDistribution(moments)
        plain = scale is None and angle is None and not self.training
        if plain and not sample_posterior:
            # eval fast path: mode -> shuffle -> BN(eval) -> inverse BN -> unshuffle, all per-(c, parity) affines
            z_norm = ops.latent_norm(moments, self.bn.running_mean, self.bn.running_var, self.bn.eps,
                                     self.encoder.z_channels)
            h = ops.latent_denorm(z_norm, self.bn.running_mean, self.bn.running_var, self.bn_eps, compute_dtype())
            return tape.act_to_nchw_f32(self.decoder.forward_act(h, wvs)), posterior
        # _static_eps: a device buffer the caller refills every step (eo_vae.graphs.GraphedTrainStep); default = CPU draw
        z = posterior.sample(getattr(self, '_static_eps', None)) if sample_posterior else posterior.mode()
        if scale is not None or angle is not None:
            # EQ-VAE transforms (:460-464): bilinear rescale + rot90 as one gather kernel (adjoint scatter in backward)
            size = self._scaled_size(z.shape[-2:], scale) if scale is not None else None
            k = 0 if angle is None else int(angle) % 4
            z = tape.LatentResizeRotFn.apply(z, size, k) if tape.grad_mode() and z.requires_grad else \
                ops.latent_resize_rot(z, size, k)
        if (self.training and self.latent_noise_p == 0 and self.bn.track_running_stats
                and self.bn.momentum is not None and z.is_cuda):
            # train-mode glue fused: unshuffle -> BN(batch stats, buffers updated) -> inverse BN -> shuffle -> activation
            h = tape.LatentTrainFn.apply(z, self.bn, self.bn_eps, compute_dtype())
            return tape.act_to_nchw_f32(self.decoder.forward_act(h, wvs)), posterior
        z_normalized = self._normalize_latent(_unshuffle2(z))
        if self.training and random.random() < self.latent_noise_p:
            z_normalized = self.noising(z_normalized)
        return self.decode(z_normalized, wvs), posterior

    # Dual-stream encode (OPT-IN, off by default): patches are independent (GroupNorm, attention and the eval-mode latent
    # BatchNorm are per sample), so a batch can be encoded as two halves on two CUDA streams, the idea being that the
    # HBM-bound passes of one half (GroupNorm apply) run UNDER the tensor-bound implicit GEMMs of the other half.  Built,
    # bit-identical, and measured to gain nothing on B200 (3033 vs 3059 patches/s): a GroupNorm-apply CTA does co-reside
    # with a resident implicit-GEMM CTA (128 threads x 80 registers, max-shared carve-out), but the pair takes exactly the
    # SUM of the two kernels' times - the N = 128 convolutions already move ~7 TB/s from L2 into shared memory, and the
    # streaming pass competes for that same fabric (tools/overlap_probe.py, profiles/r2_dual_stream_experiment.md).
    DUAL_STREAM_MIN_BATCH = int(os.environ.get('EOVAE_DUAL_STREAM_MIN_BATCH', '0'))   # 0 = off; e.g. 8 to enable

    def _side_stream(self, dev) -> 'torch.cuda.Stream':
        streams = self.__dict__.setdefault('_side_streams', {})
        if dev not in streams:
            streams[dev] = torch.cuda.Stream(dev)
        return streams[dev]

    @torch.no_grad()
    def encode_spatial_normalized(self, x: Tensor, wvs: Tensor) -> Tensor:
        """[B, C, H, W] -> spatial normalised latent [B, z, H/8, W/8] (reference :480-502) in one fused tail."""
        self.bn.eval()
        b = x.shape[0]
        if (not x.is_cuda or b < self.DUAL_STREAM_MIN_BATCH or not self.DUAL_STREAM_MIN_BATCH
                or torch.cuda.is_current_stream_capturing()):
            return ops.latent_norm(self._moments(x, wvs), self.bn.running_mean, self.bn.running_var, self.bn.eps,
                                   self.encoder.z_channels)
        dev = x.device
        f = 2 ** (self.encoder.num_resolutions - 1)
        z = torch.empty((b, self.encoder.z_channels, x.shape[2] // f, x.shape[3] // f), dtype=torch.float32, device=dev)
        half = b // 2
        main, side = torch.cuda.current_stream(dev), self._side_stream(dev)
        ops.set_tuning(ops.TUNE_GN_APPLY_CORESIDENT, 1)
        try:
            side.wait_stream(main)                      # x, wvs, the parameters and z's allocation are ready
            with torch.cuda.stream(side):
                ops.latent_norm(self._moments(x[half:], wvs), self.bn.running_mean, self.bn.running_var, self.bn.eps,
                                self.encoder.z_channels, out=z[half:])
            ops.latent_norm(self._moments(x[:half], wvs), self.bn.running_mean, self.bn.running_var, self.bn.eps,
                            self.encoder.z_channels, out=z[:half])
            main.wait_stream(side)
        finally:
            ops.set_tuning(ops.TUNE_GN_APPLY_CORESIDENT, 0)
        return z

    @torch.no_grad()
    def decode_spatial_normalized(self, z: Tensor, wvs: Tensor) -> Tensor:
        self.bn.eval()
        h = ops.latent_denorm(z, self.bn.running_mean, self.bn.running_var, self.bn_eps, compute_dtype())
        return tape.act_to_nchw_f32(self.decoder.forward_act(h, wvs))

    def _scaled_size(self, hw, scale):
        """(new_h, new_w) of the rescaled latent, multiples of the 2 x 2 packing (:519-527)."""
        h, w = hw
        sh, sw = scale if isinstance(scale, (tuple, list)) else (scale, scale)
        return round(h * sh / self.ps[0]) * self.ps[0], round(w * sw / self.ps[1]) * self.ps[1]

    def _apply_scale(self, z: Tensor, scale) -> Tensor:
        size = self._scaled_size(z.shape[-2:], scale)
        if tape.grad_mode() and z.requires_grad:
            return tape.LatentResizeRotFn.apply(z, size, 0)
        return ops.latent_resize_rot(z, size, 0)

    def _normalize_latent(self, z: Tensor) -> Tensor:
        self.bn.train() if self.training else self.bn.eval()
        return self.bn(z)

    def _inv_normalize_latent(self, z: Tensor) -> Tensor:
        self.bn.eval()
        s = torch.sqrt(self.bn.running_var.view(1, -1, 1, 1) + self.bn_eps)
        return z * s + self.bn.running_mean.view(1, -1, 1, 1)

    @torch.no_grad()
    def reconstruct(self, x: Tensor, wvs: Tensor) -> Tensor:
        return self.forward(x, wvs, sample_posterior=False)[0]

    @torch.no_grad()
    def encode_to_latent(self, x: Tensor, wvs: Tensor) -> Tensor:
        """packed normalised latent [B, 4z, H/16, W/16] (reference :730-738)."""
        return _unshuffle2(self.encode_spatial_normalized(x, wvs))

    # ------------------------------------------------------------------------------------------- training
    def configure_optimizers(self):
        params = [p for p in self.encoder.parameters() if p.requires_grad] + \
                 [p for p in self.decoder.parameters() if p.requires_grad]
        # same optimiser as the reference (:556).  With everything on the GPU: the multi-tensor Adam of csrc/optimizer.cu,
        # a torch.optim.Adam subclass (same state_dict) whose step(clip_norm=...) folds clip_grad_norm_ into the update
        on_gpu = bool(params) and all(p.is_cuda and p.dtype == torch.float32 for p in params)
        if on_gpu:
            from ..optim import FusedClipAdam
            optimizers = [FusedClipAdam(params, lr=self.base_lr)]
        else:
            optimizers = [torch.optim.Adam(params, lr=self.base_lr)]
        if hasattr(self.loss_fn, 'discriminator'):
            optimizers.append(torch.optim.Adam(self.loss_fn.discriminator.parameters(), lr=self.base_lr))
        schedulers = []
        if all([self.final_lr, self.warmup_epochs, self.decay_end_epoch]):
            steps_per_epoch = 2000
            for opt in optimizers:
                sch = get_cosine_schedule_with_warmup(opt, self.warmup_epochs * steps_per_epoch,
                                                      self.decay_end_epoch * steps_per_epoch, self.base_lr, self.final_lr)
                schedulers.append({'scheduler': sch, 'interval': 'step'})
        return (optimizers, schedulers) if schedulers else optimizers

    def training_step(self, batch, batch_idx):
        """Mirror of the reference step (:587-690): forward (sampled posterior) -> loss -> manual backward -> clip ->
        Adam -> scheduler -> log.  With gradients enabled every module records one tape entry of eo_vae/autograd.py, so
        ``manual_backward`` runs the hand-written backward kernels (conv dgrad / wgrad, GroupNorm, attention,
        hypernetwork, reparameterisation, Charbonnier / L1 and MS-SSIM adjoints); under ``enable_ddp()`` the gradients
        are averaged over the ranks while backward is still running (eo_vae/ddp.py)."""
        has_disc = hasattr(self.loss_fn, 'discriminator')
        if (getattr(self, 'graph_training', False) and not has_disc
                and not (self.p_prior or self.p_prior_s or self.latent_noise_p)):
            return self._graphed_training_step(batch)
        opts, schs = self.optimizers(), self.lr_schedulers()
        opts = opts if isinstance(opts, list) else [opts]
        schs = schs if isinstance(schs, list) else ([schs] if schs else [])
        opt_gen, opt_disc = opts[0], (opts[1] if len(opts) > 1 else None)
        sch_gen, sch_disc = (schs[0] if schs else None), (schs[1] if len(schs) > 1 else None)
        images, wvs = batch[self.image_key], batch['wvs']
        bins = [0.375, 0.5, 0.75]
        target = images
        if random.random() < self.p_prior:
            angle = random.choice([1, 2, 3])
            scale = (random.choice(bins), random.choice(bins)) if self.anisotropic else random.choice(bins)
            recon, _ = self.forward(images, wvs, scale=scale, angle=angle)
            with torch.no_grad():  # area-average to the reconstruction size, then the same rotation (:618-624)
                target = ops.area_resize_rot(images, tuple(recon.shape[-2:]), angle)
        elif random.random() < self.p_prior_s:
            recon, _ = self.forward(images, wvs, scale=random.choice(bins))
            with torch.no_grad():
                target = ops.area_resize_rot(images, tuple(recon.shape[-2:]), 0)
        else:
            recon, _ = self.forward(images, wvs)
        # generator half (:638-657)
        opt_gen.zero_grad()
        if opt_disc is not None and has_disc:
            self.loss_fn.discriminator.eval()
        gen_loss, logs = self.loss_fn(inputs=target, wvs=wvs, reconstructions=recon, optimizer_idx=0,
                                      global_step=self.global_step, last_layer=self.get_last_layer(), split='train')
        self.manual_backward(gen_loss)
        _clip_and_step(opt_gen, self.clip_grad)
        if sch_gen:
            sch_gen.step()
        # discriminator half (:659-684): only for a user-supplied GAN loss exposing discriminator / disc_start / disc_weight
        # (such losses are torch modules outside the built path; the control flow around them is the reference's)
        if (opt_disc is not None and self.global_step >= self.loss_fn.disc_start and self.loss_fn.disc_weight > 0.0):
            if has_disc:
                self.loss_fn.discriminator.train()
            opt_disc.zero_grad()
            disc_loss, disc_logs = self.loss_fn(inputs=target, wvs=wvs, reconstructions=recon.detach(), optimizer_idx=1,
                                                global_step=self.global_step, last_layer=None, split='train')
            self.manual_backward(disc_loss)
            opt_disc.step()
            if sch_disc:
                sch_disc.step()
            logs.update(disc_logs)
        logs['train/lr'] = opt_gen.param_groups[0]['lr']
        self.log_dict(logs, prog_bar=True, logger=True, on_step=True, on_epoch=False)
        return gen_loss

    def _graphed_training_step(self, batch):
        """Opt-in (``model.graph_training = True``): the same step with forward + loss + backward replayed as one CUDA
        graph per (batch shape, band count, active loss terms) signature - e.g. three graphs for a mixed S2L2A / S1RTC /
        S2RGB collate.  Semantics of the eager step are kept (CPU-drawn posterior noise, clip, Adam, scheduler, logs)."""
        from ..graphs import GraphedTrainStep
        images, wvs = batch[self.image_key], batch['wvs']
        starts = getattr(self.loss_fn, 'starts', {})
        active = tuple(sorted(k for k, v in starts.items() if self.global_step >= v))
        key = (tuple(images.shape), int(wvs.numel()), active)
        cache = self.__dict__.setdefault('_train_graphs', {})
        if key not in cache:
            cache[key] = GraphedTrainStep(self, batch)
        step = cache[key]
        loss = step(batch)
        logs = dict(step.logs)
        opts = self.optimizers()
        opt_gen = opts[0] if isinstance(opts, list) else opts
        logs['train/lr'] = opt_gen.param_groups[0]['lr']
        self.log_dict(logs, prog_bar=True, logger=True, on_step=True, on_epoch=False)
        return loss

    def validation_step(self, batch, batch_idx):
        images, wvs = batch[self.image_key], batch['wvs']
        recon, _ = self.forward(images, wvs)
        val_loss, logs = self.loss_fn(inputs=images, wvs=wvs, reconstructions=recon, optimizer_idx=0,
                                      global_step=self.global_step, last_layer=None, split='val')
        self.log_dict(logs, prog_bar=True, logger=True, on_step=False, on_epoch=True)
        return val_loss

    def get_last_layer(self) -> Tensor:
        if hasattr(self.decoder, 'output_conv_weight'):
            return self.decoder.output_conv_weight
        return self.decoder.conv_out.weight
