"""Diagonal Gaussian posterior on the fused KL / reparameterisation kernel.

Interface mirror of the reference ``eo_vae/models/modules/distributions.py:19-102``.  ``sample`` and ``kl`` each
run ONE kernel over the moments (clamp, exp, reparameterise, per-sample reduction) instead of ~6 elementwise passes.
"""
from __future__ import annotations

import torch

from ... import autograd as tape
from ... import ops


class DiagonalGaussianDistribution:
    def __init__(self, parameters: torch.Tensor, deterministic: bool = False) -> None:
        if parameters.dim() != 4 or parameters.shape[1] % 2 != 0:
            raise ValueError('moments must be [B, 2*z, H, W]')
        self.parameters = parameters if parameters.dtype == torch.float32 else parameters.float()
        self.deterministic = deterministic
        self._zc = parameters.shape[1] // 2

    # tensors of the reference object, materialised lazily (views / tiny elementwise ops, off the hot path)
    @property
    def mean(self) -> torch.Tensor:
        return self.parameters[:, :self._zc]

    @property
    def logvar(self) -> torch.Tensor:
        return torch.clamp(self.parameters[:, self._zc:], -30.0, 20.0)

    @property
    def std(self) -> torch.Tensor:
        return torch.zeros_like(self.mean) if self.deterministic else torch.exp(0.5 * self.logvar)

    @property
    def var(self) -> torch.Tensor:
        return torch.zeros_like(self.mean) if self.deterministic else torch.exp(self.logvar)

    def sample(self, eps: torch.Tensor | None = None) -> torch.Tensor:
        """mean + std * eps.  Like the reference (:44-46) the noise is drawn with the CPU generator and copied to
        the device unless the caller passes ``eps``."""
        if self.deterministic:
            return self.mean.contiguous()
        if eps is None:
            eps = torch.randn(self.mean.shape).to(device=self.parameters.device)
        if tape.grad_mode() and self.parameters.requires_grad:
            return tape.SampleFn.apply(self.parameters, eps, self._zc)
        z, _ = ops.kl_reparam(self.parameters, eps, self._zc, want_z=True)
        return z

    def kl(self, other: 'DiagonalGaussianDistribution | None' = None) -> torch.Tensor:
        if self.deterministic:
            return torch.Tensor([0.0])
        if other is not None:
            return 0.5 * torch.sum(torch.pow(self.mean - other.mean, 2) / other.var + self.var / other.var - 1.0
                                   - self.logvar + other.logvar, dim=[1, 2, 3])
        return ops.kl_reparam(self.parameters, None, self._zc, want_z=False)[1]

    def nll(self, sample: torch.Tensor, dims: list[int] = [1, 2, 3]) -> torch.Tensor:
        if self.deterministic:
            return torch.Tensor([0.0])
        logtwopi = 1.8378770664093453
        return 0.5 * torch.sum(logtwopi + self.logvar + torch.pow(sample - self.mean, 2) / self.var, dim=dims)

    def mode(self) -> torch.Tensor:
        return self.mean
