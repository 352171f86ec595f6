"""B200-native drop-in for the EOFluxVAE hot path of nilsleh/eo-vae.

Same module paths, class names, constructor arguments and ``state_dict`` keys as the reference package
(``eo_vae.models.new_autoencoder.EOFluxVAE``, ``eo_vae.models.Encoder`` / ``Decoder``,
``eo_vae.models.modules.consistency_loss.EOConsistencyLoss``), with every tensor op of the path executed by the
hand-written sm_100a kernels in ``libeovae_sm100.so`` (C ABI: ``include/eovae.h``).
"""
from .settings import (compute_dtype, grad_dtype, inference_dtype, numerics_description, set_compute_dtype,  # noqa: F401
                       set_train_dtype)

__all__ = ["compute_dtype", "grad_dtype", "inference_dtype", "numerics_description", "set_compute_dtype", "set_train_dtype"]
