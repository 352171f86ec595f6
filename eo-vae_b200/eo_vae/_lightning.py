"""``lightning.LightningModule`` when Lightning is installed, otherwise the minimal stand-in the EOFluxVAE step
needs (manual optimisation: ``optimizers()``, ``lr_schedulers()``, ``manual_backward``, ``log_dict``,
``global_step``).  The stand-in lets ``training_step`` run under a plain loop (tests, bench.py) unchanged."""
from __future__ import annotations

import torch

try:  # pragma: no cover - depends on the environment
    from lightning import LightningModule  # type: ignore
    HAVE_LIGHTNING = True
except Exception:  # noqa: BLE001
    HAVE_LIGHTNING = False

    class _CountingOptimizer:
        """What ``LightningOptimizer`` does for manual optimisation: forwards to the wrapped optimiser (``.optimizer``)
        and advances the module's ``global_step`` once per ``step()`` call, so ``*_start_step`` loss schedules and
        warm-ups move under a plain loop exactly as they do under ``Trainer.fit``."""

        def __init__(self, optimizer, module) -> None:
            self.__dict__['optimizer'] = optimizer
            self.__dict__['_module'] = module

        def step(self, *args, **kwargs):
            out = self.optimizer.step(*args, **kwargs)
            self._module.global_step += 1
            return out

        def __getattr__(self, name):
            return getattr(self.__dict__['optimizer'], name)

        def __setattr__(self, name, value):
            setattr(self.__dict__['optimizer'], name, value)

    class LightningModule(torch.nn.Module):  # type: ignore[no-redef]
        def __init__(self) -> None:
            super().__init__()
            self.automatic_optimization = True
            self.global_step = 0
            self._optimizers = None
            self._schedulers = None
            self.logged: dict = {}
            self._grad_sync = None

        def enable_ddp(self, bucket_bytes: int = 64 << 20, tail_bytes: int = 24 << 20) -> None:
            """Stand-in for Trainer(strategy='ddp'): average the gradients over the default process group inside
            ``manual_backward`` (bucketed NCCL all-reduce overlapped with backward, eo_vae/ddp.py)."""
            from .ddp import GradSync
            self._grad_sync = GradSync(self.parameters(), bucket_bytes, tail_bytes=tail_bytes)

        def attach_optimizers(self) -> None:
            """Stand-in for Trainer wiring: calls ``configure_optimizers`` once and stores the result."""
            cfg = self.configure_optimizers()
            if isinstance(cfg, tuple):
                opts, schs = cfg
                self._schedulers = [s['scheduler'] if isinstance(s, dict) else s for s in schs]
            else:
                opts, self._schedulers = cfg, []
            opts = list(opts) if isinstance(opts, (list, tuple)) else [opts]
            self._optimizers = [_CountingOptimizer(o, self) for o in opts]

        def optimizers(self):
            if self._optimizers is None:
                self.attach_optimizers()
            return self._optimizers if len(self._optimizers) > 1 else self._optimizers[0]

        def lr_schedulers(self):
            if self._optimizers is None:
                self.attach_optimizers()
            if not self._schedulers:
                return None
            return self._schedulers if len(self._schedulers) > 1 else self._schedulers[0]

        def manual_backward(self, loss: torch.Tensor) -> None:
            loss.backward()
            if self._grad_sync is not None:
                self._grad_sync.finish()

        def log_dict(self, d: dict, **_: object) -> None:
            self.logged.update({k: (v.detach() if isinstance(v, torch.Tensor) else v) for k, v in d.items()})
