"""Data-side prologue of the hot path: what the reference's collate function does on the CPU with five
tensor passes - clip, z-score, bilinear resize to ``target_size``, D4 augmentation
(``eo_vae/datasets/terramesh_datamodule.py:130-339`` normalisers, ``:347-369`` ``apply_batch_augmentations``,
``:476-482`` resize + augment in ``single_modality_collate_fn``) - as ONE gather kernel on the GPU
(``eovae_preprocess``).  Raw 16-bit digital numbers can be uploaded as they are (half the host->device bytes of fp32).
"""
from __future__ import annotations

import random

import torch

from . import _C

_IN_DT = {torch.float32: 2, torch.int16: 3, torch.uint16: 4}

# 'custom' scheme statistics of Sentinel-2 L2A (terramesh_datamodule.py:141-182); other modalities: pass mean / std
S2L2A_CUSTOM_MEAN = (1718.9949, 1825.5669, 2043.5834, 2175.4543, 2522.9522, 3114.2216, 3323.3469, 3417.3660, 3470.9655,
                     3489.4869, 2725.9735, 2152.0551)
S2L2A_CUSTOM_STD = (2126.3409, 2140.1035, 2044.6618, 2125.3351, 2065.3251, 1874.4652, 1808.0426, 1839.0210, 1737.9521,
                    1738.5136, 1456.5919, 1365.1743)


def draw_d4(rng=random):
    """The reference's draw order (terramesh_datamodule.py:357-367): h-flip, v-flip, k in [0, 3]."""
    return rng.random() > 0.5, rng.random() > 0.5, rng.randint(0, 3)


class BatchPreprocessor(torch.nn.Module):
    """``scheme='custom'``: clip to [0, 10000] then (x - mean) / std (Sentinel2L2ANorm / Sentinel2L1CNorm);
    ``scheme='legacy'``: (x - mean) / (std + 1e-8) (LegacyZScoreNorm).  ``forward(images, augment=None)`` returns the
    normalised [B, C, H', W'] fp32 batch; ``augment = (flip_h, flip_v, k)`` or ``True`` to draw like the reference."""

    def __init__(self, mean, std, scheme: str = 'legacy', target_size=(224, 224)):
        super().__init__()
        if scheme not in ('legacy', 'custom'):
            raise ValueError("scheme must be 'legacy' or 'custom'")
        self.register_buffer('mean', torch.as_tensor(mean, dtype=torch.float32).reshape(-1))
        self.register_buffer('std', torch.as_tensor(std, dtype=torch.float32).reshape(-1))
        self.scheme = scheme
        self.target_size = None if target_size is None else tuple(target_size)

    @torch.no_grad()
    def forward(self, images: torch.Tensor, augment=None) -> torch.Tensor:
        if not images.is_cuda:
            raise RuntimeError('BatchPreprocessor: CUDA tensor required (upload the raw batch, no CPU path)')
        if images.dtype not in _IN_DT:
            images = images.float()
        images = images.contiguous()
        n, c, hi, wi = images.shape
        if c != self.mean.numel():
            raise RuntimeError(f'BatchPreprocessor: {c} bands but statistics for {self.mean.numel()}')
        hr, wr = self.target_size if self.target_size is not None else (hi, wi)
        if augment is True:
            augment = draw_d4()
        fh, fv, k = augment if augment else (False, False, 0)
        ho, wo = (wr, hr) if k % 2 else (hr, wr)
        out = torch.empty((n, c, ho, wo), dtype=torch.float32, device=images.device)
        custom = self.scheme == 'custom'
        rc = _C.lib().eovae_preprocess(images.data_ptr(), _IN_DT[images.dtype], n, c, hi, wi, hr, wr, self.mean.data_ptr(),
                                       self.std.data_ptr(), 0.0 if custom else 1e-8, 1 if custom else 0, 0.0, 10000.0,
                                       int(bool(fh)), int(bool(fv)), int(k), out.data_ptr(),
                                       torch.cuda.current_stream().cuda_stream)
        _C.check(rc, 'eovae_preprocess')
        return out
