"""Process-wide numeric settings of the CUDA path."""
import torch

_COMPUTE_DTYPE = torch.bfloat16


def compute_dtype() -> torch.dtype:
    """16-bit storage / tensor-core operand type of the activations (fp32 accumulate everywhere)."""
    return _COMPUTE_DTYPE


def set_compute_dtype(dtype: torch.dtype) -> None:
    """bf16 (default, what BASELINE.json names) or fp16 (the reference trainer's ``precision: 16-mixed``)."""
    global _COMPUTE_DTYPE
    if dtype not in (torch.bfloat16, torch.float16):
        raise ValueError("compute dtype must be torch.bfloat16 or torch.float16")
    _COMPUTE_DTYPE = dtype
