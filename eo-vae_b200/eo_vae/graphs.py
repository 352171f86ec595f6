"""CUDA-graph replay of the encode path.

One ``encode_spatial_normalized`` call is ~120 kernel launches; for a fixed (batch, bands, size) the whole sequence -
hypernetwork, every implicit-GEMM conv with its TMA descriptors, GroupNorm, attention, latent tail - is captured once
into a CUDA graph and replayed with a single launch.  The graph owns static input / wavelength / output buffers; a call
copies the new batch in (device-to-device or host-to-device, asynchronous) and replays.
"""
from __future__ import annotations

import torch


class GraphedEncoder:
    """``GraphedEncoder(model, example_x, wvs)(x) -> latents`` with the semantics of
    ``model.encode_spatial_normalized(x, wvs)`` for inputs of the example's shape.  The returned tensor is the graph's
    static output buffer: consume or copy it before the next call."""

    def __init__(self, model, example_x: torch.Tensor, wvs: torch.Tensor, warmup: int = 2):
        dev = wvs.device
        self.model = model
        self.x = torch.empty(example_x.shape, dtype=torch.float32, device=dev)
        self.x.copy_(example_x)
        self.wvs = wvs.detach().clone()
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):  # builds every weight-operand cache and kernel attribute outside the capture
                model.encode_spatial_normalized(self.x, self.wvs)
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.z = model.encode_spatial_normalized(self.x, self.wvs)

    @torch.no_grad()
    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        if tuple(x.shape) != tuple(self.x.shape):
            raise RuntimeError(f"GraphedEncoder was captured for shape {tuple(self.x.shape)}, got {tuple(x.shape)}")
        if x.data_ptr() != self.x.data_ptr():
            self.x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.z


def _invalidate_operand_caches(model) -> None:
    """The derived 16-bit weight operands are keyed on ``weight._version``, which a graph replay does not advance."""
    for m in model.modules():
        if hasattr(m, "_packed_key"):
            m._packed_key = None
        if hasattr(m, "_fused_key"):
            m._fused_key = None
        if hasattr(m, "_qkv_key"):
            m._qkv_key = None
        if hasattr(m, "_up_key"):
            m._up_key = None


class GraphedTrainStep:
    """One EOFluxVAE optimisation step with forward + loss + backward replayed as ONE CUDA graph (~1.4k kernel launches
    whose Python/ctypes issue cost otherwise leaves the GPU idle ~20 % of the step), followed by the eager gradient
    exchange / clip / Adam / scheduler of ``training_step`` (new_autoencoder.py:587-690).

    Fixed at capture: batch shape, wavelengths (band count), the python-side branches (``p_prior = p_prior_s = 0``,
    MS-SSIM active or not).  The reparameterisation noise is drawn per step on the CPU generator like the reference
    (distributions.py:44) and copied into the graph's static buffer."""

    def __init__(self, model, batch: dict, warmup: int = 3):
        if model.p_prior or model.p_prior_s or model.latent_noise_p:
            raise RuntimeError("GraphedTrainStep: the EQ-VAE scale/rotation and latent-noise branches are python-side random "
                               "choices and cannot be frozen into a graph")
        self.model = model
        dev = batch["wvs"].device
        self.x = batch[model.image_key].detach().to(torch.float32).clone()
        self.wvs = batch["wvs"].detach().clone()
        b, _, h, w = self.x.shape
        f = 2 ** (model.encoder.num_resolutions - 1)
        self.eps_shape = (b, model.encoder.z_channels, h // f, w // f)
        self.eps = torch.zeros(self.eps_shape, dtype=torch.float32, device=dev)
        self.params = [p for p in model.parameters() if p.requires_grad]
        model._static_eps = self.eps
        if getattr(model, "_grad_sync", None) is not None:
            model._grad_sync.overlap = False  # autograd hooks do not run on replay: exchange after the graph instead
        self._refresh = getattr(model.loss_fn, 'refresh_graph_scalars', None)
        if self._refresh is not None:
            self._refresh(model.global_step, dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._zero()
                self.eps.copy_(torch.randn(self.eps_shape))
                self._forward_backward()
        torch.cuda.current_stream(dev).wait_stream(side)
        self._zero()
        # the 16-bit weight operands must be REBUILT INSIDE the graph (the optimiser changes the master weights between
        # replays): drop the caches so the pack kernels are captured and their outputs live in the graph's pool
        _invalidate_operand_caches(model)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._forward_backward()
        self.grads = [p.grad for p in self.params]
        self.step_index = 0
        model._static_eps = None  # only the capture reads it; eager steps draw their own noise again

    def _zero(self) -> None:
        for p in self.params:
            p.grad = None

    def _forward_backward(self) -> torch.Tensor:
        m = self.model
        recon, _ = m(self.x, self.wvs)
        loss, self.logs = m.loss_fn(inputs=self.x, wvs=self.wvs, reconstructions=recon, optimizer_idx=0,
                                    global_step=m.global_step, last_layer=m.get_last_layer(), split='train')
        loss.backward()
        return loss.detach()

    def __call__(self, batch: dict) -> torch.Tensor:
        m = self.model
        x = batch[m.image_key]
        if tuple(x.shape) != tuple(self.x.shape) or batch["wvs"].numel() != self.wvs.numel():
            raise RuntimeError("GraphedTrainStep was captured for another batch shape / band count")
        if x.data_ptr() != self.x.data_ptr():
            self.x.copy_(x, non_blocking=True)
        self.wvs.copy_(batch["wvs"], non_blocking=True)
        self.eps.copy_(torch.randn(self.eps_shape), non_blocking=False)
        if self._refresh is not None:   # global_step-dependent loss scalars live in device memory the graph reads
            self._refresh(m.global_step, self.eps.device)
        self.graph.replay()
        for p, g in zip(self.params, self.grads):
            p.grad = g
        if getattr(m, "_grad_sync", None) is not None:
            m._grad_sync.finish()
        opts = m.optimizers()
        opt = opts[0] if isinstance(opts, list) else opts
        from .models.new_autoencoder import _clip_and_step
        _clip_and_step(opt, m.clip_grad)
        schs = m.lr_schedulers()
        sch = schs[0] if isinstance(schs, list) and schs else schs
        if sch:
            sch.step()
        self.step_index += 1
        return self.loss
