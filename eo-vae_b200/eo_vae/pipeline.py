"""Host-fed encode loop with copy/compute overlap.

The reference's latent-encoding driver (``encode_latents.py:305-352``) copies a batch to the GPU, encodes it and copies
the latents back, serially.  ``encode_stream`` keeps the same per-batch call (``model.encode_spatial_normalized``) but
runs the host->device copy of batch i+1 and the device->host copy of latents i-1 on side streams while batch i is in
the kernels (two device input buffers, CUDA events for the hand-offs; pinned host memory required for true overlap).
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch


@torch.no_grad()
def encode_stream(model, host_batches: Iterable[torch.Tensor], wvs: torch.Tensor,
                  host_out: Optional[List[torch.Tensor]] = None) -> List[torch.Tensor]:
    """Encode an iterable of host (ideally pinned) [B, C, H, W] batches; returns the host latent tensors, complete
    (all streams synchronised) on return.  ``host_out``: optional preallocated pinned output tensors, one per batch."""
    dev = wvs.device
    main = torch.cuda.current_stream(dev)
    h2d, d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    bufs: list = [None, None]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [None, None]
    outs: List[torch.Tensor] = []
    for i, hb in enumerate(host_batches):
        b = i & 1
        if bufs[b] is None or bufs[b].shape != hb.shape:
            bufs[b] = torch.empty(hb.shape, dtype=torch.float32, device=dev)
        with torch.cuda.stream(h2d):
            if consumed[b] is not None:
                h2d.wait_event(consumed[b])      # the kernels that read this buffer two batches ago are done
            bufs[b].copy_(hb, non_blocking=True)
            ready[b].record(h2d)
        main.wait_event(ready[b])
        z = model.encode_spatial_normalized(bufs[b], wvs)
        consumed[b] = torch.cuda.Event()
        consumed[b].record(main)
        out = host_out[i] if host_out is not None else torch.empty(z.shape, dtype=z.dtype).pin_memory()
        with torch.cuda.stream(d2h):
            d2h.wait_event(consumed[b])
            out.copy_(z, non_blocking=True)
        z.record_stream(d2h)
        outs.append(out)
    d2h.synchronize()
    main.synchronize()
    return outs
