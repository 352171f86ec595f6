"""Process-wide numeric settings of the CUDA path.

Three modes, selected with ``set_compute_dtype``:

* ``torch.float16`` (DEFAULT): activations and tensor-core operands in IEEE half, fp32 accumulation, fp32 GroupNorm
  statistics / softmax / hypernetwork / reductions.  This is the reference trainer's own 16-bit type
  (``configs/eo-vae.yaml:79`` ``precision: 16-mixed``: every conv input and output is fp16 under autocast), runs the same
  kernels at the same tensor-core rate as bf16 (one field of the UMMA instruction descriptor), and is the only 16-bit
  format that meets the stated parity (latents and reconstructions within 1e-2 of the fp32 reference, losses within 1e-3):
  measured 1.0-1.4e-3 / 2.4-3.3e-3 against 0.8-1.1e-2 / 1.9-2.8e-2 for bf16, whose 8-bit significand puts the *ideal*
  bf16 implementation of this network at 0.7e-2 / 1.6e-2 (``tools/precision_roles.py``: bf16 storage of the residual
  stream alone costs 0.7e-2 / 1.4e-2, bf16 rounding of the un-normalised conv operands another 0.6e-2 / 1.2e-2).
  Gradients flowing between layers in the training step are bf16 (``grad_dtype``): their dynamic range is what needs the
  wide exponent (1 / (B*C*H*W) ~ 8e-8 at the shipped batch), exactly why the reference pairs fp16 with a GradScaler.
* ``torch.bfloat16``: everything 16-bit in bf16 (fp16's range is 65504; a checkpoint whose residual stream exceeds it
  needs this mode and accepts the ~1e-2 / 2e-2 deviation).
* ``torch.float32``: the validation path - fp32 activations and fp32 SIMT kernels end to end (``csrc/fp32_path.cu``), eval
  only, within 1e-4 of the fp32 reference.  Not a fast path.
"""
import torch

_COMPUTE_DTYPE = torch.float16
_GRAD_DTYPE = torch.bfloat16


def compute_dtype() -> torch.dtype:
    """Storage / operand type of the activations."""
    return _COMPUTE_DTYPE


def grad_dtype() -> torch.dtype:
    """Storage type of the gradients that flow between layers in the training step (parameter gradients are fp32)."""
    return _GRAD_DTYPE if _COMPUTE_DTYPE != torch.float32 else torch.float32


def set_compute_dtype(dtype: torch.dtype) -> None:
    global _COMPUTE_DTYPE
    if dtype not in (torch.bfloat16, torch.float16, torch.float32):
        raise ValueError("compute dtype must be torch.float16 (default), torch.bfloat16 or torch.float32 (validation path)")
    _COMPUTE_DTYPE = dtype


def default_compute_dtype() -> torch.dtype:
    return torch.float16


def numerics_description() -> str:
    """One line for logs / bench output."""
    if _COMPUTE_DTYPE == torch.float32:
        return "fp32 validation path: fp32 activations, fp32 SIMT kernels (no tensor cores)"
    name = "fp16" if _COMPUTE_DTYPE == torch.float16 else "bf16"
    return (f"{name} activations and tcgen05 operands (kind::f16), fp32 accumulate; fp32 GroupNorm statistics, softmax, "
            f"hypernetwork and reductions; bf16 inter-layer gradients in training")
