"""EO consistency loss (pixel + MS-SSIM branches) on the sm_100a reduction kernels.

Interface mirror of the reference ``eo_vae/models/modules/consistency_loss.py`` (CharbonnierLoss :12-21, SSIMLoss
:24-37, EOConsistencyLoss :329-483): same constructor arguments, ``forward(inputs, wvs, reconstructions,
global_step, split, **kwargs) -> (total, logs)`` and log keys.  Pixel (L1 / Charbonnier), MS-SSIM, spectral-angle,
gradient-difference and focal-frequency branches all run on the reduction kernels (values and gradients); only the
DOFA semantic-feature branch (``feature_weight > 0``: needs an external pretrained network) raises.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ... import autograd as tape
from ... import ops


class CharbonnierLoss(nn.Module):
    def __init__(self, eps=1e-3):
        super().__init__()
        self.eps = eps

    def forward(self, pred, target):
        return ops.l1_charbonnier(pred, target, self.eps)[1]


class SSIMLoss(nn.Module):
    """1 - MS-SSIM (data_range 6, 5 scales) on the fused per-scale kernels (reference :24-37 via torchmetrics)."""

    def __init__(self, channels=12):
        super().__init__()
        self.data_range = 6.0

    def forward(self, pred, target):
        if tape.grad_mode() and pred.requires_grad:
            return 1.0 - tape.MsssimFn.apply(pred, target, self.data_range)[0]
        return 1.0 - ops.msssim(pred, target, self.data_range)[0][0]


class SAMLoss(nn.Module):
    """1 - cosine similarity along the band axis, averaged (consistency_loss.py:186-210) on eovae_sam_loss."""

    def __init__(self, eps=1e-8):
        super().__init__()
        self.eps = eps

    def forward(self, x_rec, x_true):
        if tape.grad_mode() and x_rec.requires_grad:
            return tape.SamLossFn.apply(x_rec, x_true, self.eps)
        return ops.sam_loss(x_rec, x_true, self.eps)


class GradientDifferenceLoss(nn.Module):
    """| |dx pred| - |dx target| | + the same along y, averaged (consistency_loss.py:241-269) on eovae_grad_diff_loss.
    The kernels implement alpha = 1, the value EOConsistencyLoss constructs it with (:386)."""

    def __init__(self, alpha=1.0):
        super().__init__()
        if float(alpha) != 1.0:
            raise NotImplementedError('GradientDifferenceLoss: only alpha = 1 is built (EOConsistencyLoss default)')
        self.alpha = alpha

    def forward(self, pred, target):
        if tape.grad_mode() and pred.requires_grad:
            return tape.GradDiffLossFn.apply(pred, target)
        return ops.grad_diff_loss(pred, target)


class FocalFrequencyLoss(nn.Module):
    """ffl.py:17-104 on eovae_focal_freq_loss: the configuration EOConsistencyLoss uses (ave_spectrum False, batch_matrix
    and log_matrix True, no external weight matrix); other option combinations raise."""

    def __init__(self, loss_weight=1.0, alpha=1.0, patch_factor=1, ave_spectrum=False, log_matrix=False, batch_matrix=False):
        super().__init__()
        if ave_spectrum or not (log_matrix and batch_matrix):
            raise NotImplementedError('FocalFrequencyLoss: built for ave_spectrum=False, log_matrix=True, batch_matrix=True')
        self.loss_weight, self.alpha, self.patch_factor = loss_weight, alpha, patch_factor
        self.ave_spectrum, self.log_matrix, self.batch_matrix = ave_spectrum, log_matrix, batch_matrix

    def forward(self, pred, target, matrix=None, **kwargs):
        if matrix is not None:
            raise NotImplementedError('FocalFrequencyLoss: external weight matrix is not supported')
        if tape.grad_mode() and pred.requires_grad:
            return tape.FocalFreqLossFn.apply(pred, target, self.patch_factor, self.alpha) * self.loss_weight
        return ops.focal_freq_loss(pred, target, self.patch_factor, self.alpha)[0] * self.loss_weight


class EOConsistencyLoss(nn.Module):
    def __init__(self, pixel_weight: float = 1.0, rec_loss_type: str = 'l1', spectral_weight: float = 0.0,
                 spatial_weight: float = 0.0, freq_weight: float = 0.0, feature_weight: float = 0.0,
                 msssim_weight: float = 0.0, spectral_start_step: int = 0, spatial_start_step: int = 0,
                 freq_start_step: int = 0, feature_start_step: int = 0, msssim_start_step: int = 0,
                 patch_factor: int = 2, ffl_alpha: float = 1.0, dofa_net: nn.Module = None):
        super().__init__()
        if feature_weight > 0 or dofa_net is not None:  # DOFA feature branch: needs an external pretrained network
            raise NotImplementedError('feature_weight > 0: the DOFA semantic-feature branch is outside the built hot path')
        if rec_loss_type not in ('l1', 'char'):
            raise ValueError("rec_loss_type must be 'l1' or 'char'")
        self.rec_loss_type = rec_loss_type
        self.starts = {'spectral': spectral_start_step, 'spatial': spatial_start_step, 'freq': freq_start_step,
                       'feature': feature_start_step, 'msssim': msssim_start_step}
        self.weights = {'pixel': pixel_weight, 'spectral': spectral_weight, 'spatial': spatial_weight,
                        'freq': freq_weight, 'feature': feature_weight, 'msssim': msssim_weight}
        self.sam_loss = SAMLoss()
        self.grad_loss = GradientDifferenceLoss()
        self.fft_loss = FocalFrequencyLoss(loss_weight=1.0, alpha=ffl_alpha, patch_factor=patch_factor, ave_spectrum=False,
                                           batch_matrix=True, log_matrix=True)
        self.char_loss = CharbonnierLoss()
        self.msssim_loss = SSIMLoss()
        self._graph_ffl = None   # device scalar read by a CAPTURED step (see refresh_graph_scalars)

    def ffl_weight_at(self, global_step: int) -> float:
        """freq_weight x linear warm-up over 1000 steps after freq_start_step (consistency_loss.py:443-452)."""
        warmup = min(1.0, max(0.0, (global_step - self.starts['freq']) / 1000))
        return self.weights['freq'] * warmup

    def refresh_graph_scalars(self, global_step: int, device) -> None:
        """Every ``global_step``-dependent scalar of the loss, as device memory a CUDA graph reads at replay time.  A
        Python float would be frozen into the captured kernels (the warm-up would stay at its capture-time value
        forever); ``eo_vae.graphs.GraphedTrainStep`` calls this before the capture and before every replay."""
        if self._graph_ffl is None or self._graph_ffl.device != torch.device(device):
            self._graph_ffl = torch.zeros((), dtype=torch.float32, device=device)
        self._graph_ffl.fill_(self.ffl_weight_at(global_step))

    def forward(self, inputs: torch.Tensor, wvs: torch.Tensor, reconstructions: torch.Tensor, global_step: int = 0,
                split: str = 'train', **kwargs):
        logs = {}
        total = torch.zeros((), device=inputs.device)
        if self.weights['pixel'] > 0:
            kind = 0 if self.rec_loss_type == 'l1' else 1
            if tape.grad_mode() and reconstructions.requires_grad:
                l_rec = tape.PixelLossFn.apply(reconstructions, inputs, self.char_loss.eps, kind)
            else:
                l_rec = ops.l1_charbonnier(reconstructions, inputs, self.char_loss.eps)[kind]
            total = total + self.weights['pixel'] * l_rec
            logs[f'{split}/loss_rec'] = l_rec.detach()
        if self.weights['spectral'] > 0 and global_step >= self.starts['spectral']:
            l_sam = self.sam_loss(reconstructions, inputs)
            total = total + self.weights['spectral'] * l_sam
            logs[f'{split}/loss_spectral'] = l_sam.detach()
        if self.weights['spatial'] > 0 and global_step >= self.starts['spatial']:
            l_spat = self.grad_loss(reconstructions, inputs)
            total = total + self.weights['spatial'] * l_spat
            logs[f'{split}/loss_spatial'] = l_spat.detach()
        if self.weights['freq'] > 0 and global_step >= self.starts['freq']:
            raw = self.fft_loss(reconstructions, inputs)
            if inputs.is_cuda and torch.cuda.is_current_stream_capturing():
                if self._graph_ffl is None:
                    raise RuntimeError('EOConsistencyLoss: call refresh_graph_scalars(global_step, device) before capturing '
                                       'a step with freq_weight > 0 (the warm-up must not be baked into the graph)')
                current = self._graph_ffl          # device scalar, refilled by the caller before each replay
                logs[f'{split}/ffl_weight'] = current
            else:
                current = self.ffl_weight_at(global_step)
                logs[f'{split}/ffl_weight'] = torch.tensor(current)
            total = total + raw * current
            logs[f'{split}/loss_freq_raw'] = raw.detach()
        if self.weights['msssim'] > 0 and global_step >= self.starts['msssim']:
            l_msssim = self.msssim_loss(reconstructions, inputs)
            total = total + self.weights['msssim'] * l_msssim
            logs[f'{split}/loss_msssim'] = l_msssim.detach()
        logs[f'{split}/loss_total'] = total.detach()
        return total, logs
