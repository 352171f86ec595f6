"""Whole-path parity on the GPU: the drop-in EOFluxVAE (CUDA kernels via the C ABI) against the committed golden
vectors produced by the unmodified reference, and against the CPU oracle on fresh seeded inputs.

Tolerances (BASELINE.json north_star): latents / reconstructions within 1e-2 relative (rel-L2 against the fp32
reference) for 16-bit tensor-core operands; KL and pixel losses within 1e-3 relative."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rel(a, b):
    a, b = torch.as_tensor(a).float(), torch.as_tensor(b).float()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _setup(tag, cuda):
    import __graft_entry__ as g
    from oracle.weights import FULL_CONFIG, TINY_CONFIG, WAVELENGTHS, make_state_dict, synthetic_patches
    gold = np.load(os.path.join(GOLDEN, tag + ".npz"))
    cfg = TINY_CONFIG if tag.startswith("tiny") else FULL_CONFIG
    seed, size, batch, modality = int(gold["seed"]), int(gold["size"]), int(gold["batch"]), str(gold["modality"])
    sd = make_state_dict(cfg, seed)
    model = g._model(cfg, sd, cuda)
    wvs = torch.tensor(WAVELENGTHS[modality], dtype=torch.float32)
    x = synthetic_patches(batch, len(wvs), size, seed=1234 + seed)
    return gold, model, x.to(cuda), wvs.to(cuda)


# rel-L2 bounds (latent, reconstruction, KL, L1) per numeric mode, against the fp32 reference's golden vectors.
#  * torch.float16 = the package DEFAULT and the dtype of every bench line: the north-star values, unwidened (latents and
#    reconstructions within 1e-2, losses within 1e-3; measured 1.0-1.4e-3 / 2.4-3.3e-3).
#  * torch.float32 = the fp32 validation path: 1e-4.
#  * torch.bfloat16 = the opt-in wide-range mode.  Its 8-bit significand puts an IDEAL bf16 implementation of this network
#    (only MMA operands rounded) at 0.7e-2 / 1.6e-2 (tools/precision_roles.py), so it is held to documented looser bounds
#    and is not what the package ships as default.
BOUNDS = {torch.float16: (1e-2, 1e-2, 1e-3, 1e-3), torch.float32: (1e-4, 1e-4, 1e-4, 1e-4),
          torch.bfloat16: (1.5e-2, 3.5e-2, 1e-3, 2e-3)}
DEFAULT = torch.float16


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32, torch.bfloat16])
@pytest.mark.parametrize("tag", ["tiny_s2l2a", "tiny_s1rtc", "tiny_s2l1c", "full_s2l2a_64", "full_s2rgb", "full_s2l2a_256"])
def test_against_reference_golden(cuda, tag, dtype):
    import eo_vae
    gold, model, x, wvs = _setup(tag, cuda)
    assert eo_vae.inference_dtype() == DEFAULT        # the default mode is the one held to the north-star bounds
    eo_vae.set_compute_dtype(dtype)
    with torch.no_grad():
        moments = model.encoder(x, wvs)
        z = model.encode_spatial_normalized(x, wvs)
        recon = model.reconstruct(x, wvs)
        post = model.encode(x, wvs)
        kl = post.kl()
    st = int(gold["recon_stride"]) if "recon_stride" in gold.files else 1   # large fixtures keep a pixel sub-lattice
    assert moments.shape == gold["moments"].shape and moments.is_contiguous()
    assert z.shape == gold["z_norm"].shape and recon.shape == x.shape
    e_m, e_z = _rel(moments.cpu(), gold["moments"]), _rel(z.cpu(), gold["z_norm"])
    e_r = _rel(recon[..., ::st, ::st].cpu(), gold["recon"])
    e_kl = _rel(kl.cpu(), gold["kl"])
    l1 = float((recon.cpu() - x.cpu()).abs().mean())
    e_l1 = abs(l1 - float(gold["l1"])) / float(gold["l1"])
    print(f"PARITY {tag} {str(dtype).split('.')[-1]}: moments {e_m:.3e} latent {e_z:.3e} recon {e_r:.3e} "
          f"kl {e_kl:.3e} l1 {e_l1:.3e}")
    bz, br, bkl, bl1 = BOUNDS[dtype]
    assert e_z < bz, f"latent rel-L2 {e_z}"
    assert e_r < br, f"reconstruction rel-L2 {e_r}"
    assert e_kl < bkl, f"KL {e_kl}"
    assert e_l1 < bl1, f"L1 {e_l1}"


def test_sample_and_decode_api(cuda):
    from oracle import eovae_oracle as O
    from oracle.weights import TINY_CONFIG, make_state_dict
    gold, model, x, wvs = _setup("tiny_s2l2a", cuda)
    sd = make_state_dict(TINY_CONFIG, int(gold["seed"]))
    with torch.no_grad():
        post = model.encode(x, wvs)
        eps = torch.from_numpy(np.random.Generator(np.random.Philox(key=[int(gold["seed"]), 99])).standard_normal(
            tuple(post.mean.shape), dtype=np.float32))
        zs = post.sample(eps.to(cuda))
        assert _rel(zs.cpu(), gold["z_sample"]) < BOUNDS[DEFAULT][0]
        # decode_spatial_normalized(encode_spatial_normalized(x)) == reconstruct(x)
        z = model.encode_spatial_normalized(x, wvs)
        r1 = model.decode_spatial_normalized(z, wvs)
        r2 = model.reconstruct(x, wvs)
        assert _rel(r1, r2) < 2e-3
        # packed-latent API
        zp = model.encode_to_latent(x, wvs)
        assert zp.shape == (x.shape[0], 4 * TINY_CONFIG["z_channels"], x.shape[2] // 8, x.shape[3] // 8)
        r3 = model.decode(zp, wvs)
        assert _rel(r3, r2) < 2e-3
        ref = O.decode(sd, zp.cpu(), wvs.cpu(), TINY_CONFIG["hyper_heads"])
        assert _rel(r3.cpu(), ref) < BOUNDS[DEFAULT][1]


def test_full_size_properties(cuda):
    """BASELINE configs[1] shape (12 x 256 x 256) at reduced batch: batch-independence (the path shards by patch)
    and determinism - the size-independent properties the domain offers."""
    import __graft_entry__ as g
    from oracle.weights import FULL_CONFIG, WAVELENGTHS, make_state_dict, synthetic_patches
    sd = make_state_dict(FULL_CONFIG, 0)
    model = g._model(FULL_CONFIG, sd, cuda)
    wvs = torch.tensor(WAVELENGTHS["S2L2A"]).to(cuda)
    x = synthetic_patches(4, 12, 256, seed=7).to(cuda)
    with torch.no_grad():
        z_all = model.encode_spatial_normalized(x, wvs)
        z_again = model.encode_spatial_normalized(x, wvs)
        z_one = model.encode_spatial_normalized(x[2:3], wvs)
    assert z_all.shape == (4, 32, 32, 32)
    assert torch.equal(z_all, z_again), "encode is not deterministic"
    assert torch.equal(z_all[2:3], z_one), "a patch's latent depends on its batch neighbours"
    assert torch.isfinite(z_all).all()


def test_config5_512px_13band(cuda):
    """BASELINE configs[4] shape: S2L1C 13-band 512x512 (attention over L = 4096 tokens), parity against the oracle."""
    import __graft_entry__ as g
    from oracle import eovae_oracle as O
    from oracle.weights import FULL_CONFIG, WAVELENGTHS, make_state_dict, synthetic_patches
    sd = make_state_dict(FULL_CONFIG, 1)
    model = g._model(FULL_CONFIG, sd, cuda)
    wvs = torch.tensor(WAVELENGTHS["S2L1C"])
    x = synthetic_patches(2, 13, 512, seed=21)
    with torch.no_grad():
        z = model.encode_spatial_normalized(x.to(cuda), wvs.to(cuda))
        torch.set_num_threads(max(1, os.cpu_count() or 1))
        z_ref = O.encode_spatial_normalized(sd, x[:1], wvs, FULL_CONFIG["hyper_heads"])
    assert z.shape == (2, 32, 64, 64)
    e = _rel(z[:1].cpu(), z_ref)
    print(f"PARITY config5 512px default numerics: latent {e:.3e}")
    assert e < BOUNDS[DEFAULT][0]


def test_cuda_graph_replay_is_bit_identical(cuda):
    """eo_vae.graphs.GraphedEncoder: the ~120 launches of an encode captured once, replayed with new inputs."""
    import __graft_entry__ as g
    from eo_vae.graphs import GraphedEncoder
    from oracle.weights import FULL_CONFIG, WAVELENGTHS, make_state_dict, synthetic_patches
    model = g._model(FULL_CONFIG, make_state_dict(FULL_CONFIG, 0), cuda)
    wvs = torch.tensor(WAVELENGTHS["S2L2A"]).to(cuda)
    x1 = synthetic_patches(2, 12, 256, seed=31).to(cuda)
    x2 = synthetic_patches(2, 12, 256, seed=32).to(cuda)
    with torch.no_grad():
        ref1 = model.encode_spatial_normalized(x1, wvs).clone()
        ref2 = model.encode_spatial_normalized(x2, wvs).clone()
        ge = GraphedEncoder(model, x1, wvs)
        assert torch.equal(ge(x1), ref1)
        assert torch.equal(ge(x2), ref2)


@pytest.mark.parametrize("modality,size,batch", [("S2RGB", 224, 1), ("S1RTC", 96, 3), ("S2L2A", 144, 2)])
def test_forward_like_reference_test(cuda, modality, size, batch):
    """The reference's own test (tests/test_model.py:19-27: x = randn(1, 3, 224, 224), recon.shape == x.shape) plus
    numerical parity with the oracle, on sizes that are NOT multiples of the 128-pixel tile (ragged tiles)."""
    import __graft_entry__ as g
    from oracle import eovae_oracle as O
    from oracle.weights import FULL_CONFIG, WAVELENGTHS, make_state_dict, synthetic_patches
    sd = make_state_dict(FULL_CONFIG, 4)
    model = g._model(FULL_CONFIG, sd, cuda)
    wvs = torch.tensor(WAVELENGTHS[modality])
    x = synthetic_patches(batch, len(wvs), size, seed=41)
    with torch.no_grad():
        recon, posterior = model(x.to(cuda), wvs.to(cuda), sample_posterior=False)
        assert isinstance(recon, torch.Tensor) and recon.shape == x.shape
        assert posterior.mean.shape == (batch, 32, size // 8, size // 8)
        torch.set_num_threads(max(1, os.cpu_count() or 1))
        ref = O.reconstruct(sd, x, wvs, FULL_CONFIG["hyper_heads"])
    e = _rel(recon.cpu(), ref)
    print(f"PARITY forward {modality} {size}px default numerics: recon {e:.3e}")
    assert e < BOUNDS[DEFAULT][1]


def test_eval_operand_cache_follows_parameters_and_wavelengths(cuda):
    """The generated dynamic-conv operands are cached per wavelength vector in eval (they do not depend on the batch); the
    cache must follow in-place parameter updates (optimiser steps), new wavelength vectors and the numeric mode."""
    import eo_vae
    gold, model, x, wvs = _setup("tiny_s2l2a", cuda)
    with torch.no_grad():
        z0 = model.encode_spatial_normalized(x, wvs).clone()
        assert torch.equal(model.encode_spatial_normalized(x, wvs), z0)            # served from the cache: identical bits
        cache = model.encoder.conv_in._operand_cache
        assert len(cache) == 1
        p = model.encoder.conv_in.weight_generator.fc_weight.weight
        p.mul_(1.5)                                                                 # what an optimiser step does: in place
        z1 = model.encode_spatial_normalized(x, wvs).clone()
        assert not torch.equal(z1, z0)
        model.encoder.conv_in.CACHE_EVAL_OPERANDS = False
        assert torch.equal(model.encode_spatial_normalized(x, wvs), z1)            # == a fresh generation
        model.encoder.conv_in.CACHE_EVAL_OPERANDS = True
        wvs2 = (wvs * 1.01).clone()                                                 # another wavelength vector
        z2 = model.encode_spatial_normalized(x, wvs2)
        assert not torch.equal(z2, z1) and len(cache) == 2
        wvs2.copy_(wvs)                                                             # same tensor, new contents
        assert torch.equal(model.encode_spatial_normalized(x, wvs2), z1)
        r_cpu_wvs = model.reconstruct(x, wvs.cpu().to(cuda))                        # a fresh tensor with equal values
        assert torch.equal(r_cpu_wvs, model.reconstruct(x, wvs))
        eo_vae.set_compute_dtype(torch.bfloat16)
        zb = model.encode_spatial_normalized(x, wvs)
        assert not torch.equal(zb, z1)


def test_dual_stream_encode_is_bit_identical(cuda):
    """Opt-in dual-stream encode (two halves of the batch on two streams): the latents must not depend on it."""
    import __graft_entry__ as g
    from oracle.weights import FULL_CONFIG, WAVELENGTHS, make_state_dict, synthetic_patches
    model = g._model(FULL_CONFIG, make_state_dict(FULL_CONFIG, 0), cuda)
    wvs = torch.tensor(WAVELENGTHS["S2L2A"]).to(cuda)
    x = synthetic_patches(9, 12, 128, seed=17).to(cuda)      # odd batch: halves of 4 and 5
    with torch.no_grad():
        model.DUAL_STREAM_MIN_BATCH = 8
        z2 = model.encode_spatial_normalized(x, wvs)
        z2b = model.encode_spatial_normalized(x, wvs)
        model.DUAL_STREAM_MIN_BATCH = 0
        z1 = model.encode_spatial_normalized(x, wvs)
    assert torch.equal(z2, z1) and torch.equal(z2b, z1)
