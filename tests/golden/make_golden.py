"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU fp32.

Run in the build container only (the reference does not exist on the GPU box):
    python tests/golden/make_golden.py
Weights come from oracle.weights.make_state_dict (deterministic, numpy Philox), inputs from
oracle.weights.synthetic_patches, so the fixtures hold outputs only (plus a few spot inputs).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402
from oracle.weights import FULL_CONFIG, TINY_CONFIG, WAVELENGTHS, make_state_dict, synthetic_patches  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def run(tag, cfg, modality, size, batch, seed, recon_stride=1):
    """recon_stride > 1: the reconstruction is stored on the pixel sub-lattice [::s, ::s] (a 12 x 256 x 256 fp32 batch
    does not compress; the sub-lattice keeps the fixture small and the rel-L2 statistic unbiased)."""
    torch.manual_seed(0)
    sd = make_state_dict(cfg, seed)
    model = ref_shim.build_reference_model(cfg, sd, train=False)
    wvs = torch.tensor(WAVELENGTHS[modality], dtype=torch.float32)
    x = synthetic_patches(batch, len(wvs), size, seed=1234 + seed)
    with torch.no_grad():
        moments = model.encoder(x, wvs)
        post = model.encode(x, wvs)
        kl = post.kl()
        z_norm = model.encode_spatial_normalized(x, wvs)
        recon = model.reconstruct(x, wvs)
        w_in, b_in = model.encoder.conv_in.get_distillation_weight(wvs)
        w_out, b_out = model.decoder.conv_out.get_distillation_weight(wvs)
        eps = torch.from_numpy(np.random.Generator(np.random.Philox(key=[seed, 99])).standard_normal(
            post.mean.shape, dtype=np.float32))
        z_s = post.mean + post.std * eps
        l1 = torch.nn.functional.l1_loss(recon, x)
        char = torch.sqrt((recon - x) ** 2 + 1e-6).mean()
    np.savez_compressed(
        os.path.join(OUT, f"{tag}.npz"),
        modality=modality, size=size, batch=batch, seed=seed,
        moments=moments.numpy().astype(np.float32), kl=kl.numpy(), z_norm=z_norm.numpy(),
        recon=recon[..., ::recon_stride, ::recon_stride].contiguous().numpy(), recon_stride=recon_stride,
        recon_sum=recon.double().sum().numpy(), recon_sumsq=(recon.double() ** 2).sum().numpy(), w_in=w_in.numpy(), b_in=b_in.numpy(), w_out=w_out.numpy(), b_out=b_out.numpy(),
        z_sample=z_s.numpy(), l1=l1.numpy(), char=char.numpy())
    print(tag, "moments", tuple(moments.shape), "recon", tuple(recon.shape), "kl", kl.tolist(), "l1", float(l1))


if __name__ == "__main__":
    run("tiny_s2l2a", TINY_CONFIG, "S2L2A", 64, 2, 0)
    run("tiny_s1rtc", TINY_CONFIG, "S1RTC", 32, 3, 1)
    run("tiny_s2l1c", TINY_CONFIG, "S2L1C", 48, 1, 2)
    run("full_s2rgb", FULL_CONFIG, "S2RGB", 256, 1, 0)     # BASELINE.json configs[0]
    run("full_s2l2a_64", FULL_CONFIG, "S2L2A", 64, 2, 3)
    run("full_s2l2a_256", FULL_CONFIG, "S2L2A", 256, 2, 5, recon_stride=4)   # BASELINE.json configs[1] patch shape
