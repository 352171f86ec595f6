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
  The TRAINING step (any call made with autograd enabled) runs in ``train_dtype`` instead, bf16 by default: the gradients
  flowing between layers need the wide exponent (1 / (B*C*H*W) ~ 8e-8 at the shipped batch - why the reference pairs fp16
  with a GradScaler), and tcgen05 ``kind::f16`` rejects mixed f16 x bf16 operands (illegal instruction on sm_100a, measured),
  so the activations that meet those gradients in the weight-gradient GEMMs must be bf16 as well.  The losses of a bf16
  forward stay within the stated 1e-3.  ``set_train_dtype(torch.float16)`` trains in half precision throughout for callers
  that bring their own loss scaling (Lightning ``precision: 16-mixed``).
* ``torch.bfloat16``: everything 16-bit in bf16 (fp16's range is 65504; a checkpoint whose residual stream exceeds it
  needs this mode and accepts the ~1e-2 / 2e-2 deviation).
* ``torch.float32``: the validation path - fp32 activations and fp32 SIMT kernels end to end (``csrc/fp32_path.cu``), eval
  only, within 1e-4 of the fp32 reference.  Not a fast path.
"""
import torch

_COMPUTE_DTYPE = torch.float16
_TRAIN_DTYPE = torch.bfloat16


def compute_dtype() -> torch.dtype:
    """Storage / operand type of the activations of the call being made: the inference mode, or ``train_dtype`` when the
    call records an autograd tape."""
    if _COMPUTE_DTYPE != torch.float32 and torch.is_grad_enabled():
        return _TRAIN_DTYPE
    return _COMPUTE_DTYPE


def inference_dtype() -> torch.dtype:
    return _COMPUTE_DTYPE


def grad_dtype() -> torch.dtype:
    """Storage type of the activations AND inter-layer gradients of the training step (parameter gradients are fp32)."""
    return _TRAIN_DTYPE


def set_compute_dtype(dtype: torch.dtype) -> None:
    global _COMPUTE_DTYPE
    if dtype not in (torch.bfloat16, torch.float16, torch.float32):
        raise ValueError("compute dtype must be torch.float16 (default), torch.bfloat16 or torch.float32 (validation path)")
    _COMPUTE_DTYPE = dtype


def set_train_dtype(dtype: torch.dtype) -> None:
    global _TRAIN_DTYPE
    if dtype not in (torch.bfloat16, torch.float16):
        raise ValueError("train dtype must be torch.bfloat16 (default) or torch.float16 (caller provides loss scaling)")
    _TRAIN_DTYPE = dtype


def default_compute_dtype() -> torch.dtype:
    return torch.float16


def default_train_dtype() -> torch.dtype:
    return torch.bfloat16


def numerics_description() -> str:
    """One line for logs / bench output."""
    if _COMPUTE_DTYPE == torch.float32:
        return "fp32 validation path: fp32 activations, fp32 SIMT kernels (no tensor cores)"
    name = "fp16" if _COMPUTE_DTYPE == torch.float16 else "bf16"
    tname = "fp16" if _TRAIN_DTYPE == torch.float16 else "bf16"
    return (f"inference: {name} activations and tcgen05 operands (kind::f16), fp32 accumulate; training step: {tname} "
            f"activations, operands and inter-layer gradients, fp32 parameter gradients; fp32 GroupNorm statistics, softmax, "
            f"hypernetwork and reductions in both")
