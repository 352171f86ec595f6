"""Fused flash-style attention kernel (eovae_attention_fused) against softmax(q k^T / sqrt(c)) v in fp32 on the same
16-bit operands, including ragged sequence lengths (keys masked, query rows clipped) and both head-dimension layouts
(one CTA per query tile for C <= 256, two for C = 512)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,l,c", [(2, 1024, 512), (3, 256, 64), (2, 324, 128), (1, 4096, 512), (2, 196, 256), (1, 64, 384),
                                   (2, 100, 64)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_attention_fused_matches_reference(cuda, n, l, c, dtype):
    from eo_vae import ops
    g = torch.Generator().manual_seed(l + c)
    qkv = (torch.randn((n, l, 3 * c), generator=g) * 1.5).to(cuda).to(dtype)
    assert ops._C.lib().eovae_attention_fused_ok(l, c)
    out = ops.attention_fused(qkv, c)
    q, k, v = (t.float() for t in qkv.split(c, dim=2))
    ref = torch.softmax(q @ k.transpose(1, 2) / math.sqrt(c), dim=-1) @ v
    err = float((out.float() - ref).norm() / ref.norm())
    assert err < (1.2e-2 if dtype == torch.bfloat16 else 2e-3), err
    # the padded channel pitch of the fused q/k/v conv output is honoured
    wide = torch.zeros((n, l, 3 * c + 16), device=cuda, dtype=dtype)
    wide[:, :, :3 * c] = qkv
    assert torch.equal(ops.attention_fused(wide[:, :, :3 * c], c), out)


def test_attn_block_fused_equals_unfused(cuda, monkeypatch):
    from eo_vae import ops
    from eo_vae.models.modules.layers import AttnBlock
    torch.manual_seed(0)
    blk = AttnBlock(128).to(cuda)
    x = torch.randn((2, 128, 16, 16), device=cuda).to(dtype=torch.bfloat16, memory_format=torch.channels_last)
    with torch.no_grad():
        blk(x)  # builds the weight-operand caches
        monkeypatch.setattr(ops, "USE_FUSED_ATTENTION", True)
        before = ops.launch_count()
        fused = blk(x)
        fused_launches = ops.launch_count() - before
        monkeypatch.setattr(ops, "USE_FUSED_ATTENTION", False)
        before = ops.launch_count()
        unfused = blk(x)
        assert ops.launch_count() - before > fused_launches   # the fused path really replaced GEMM + softmax + GEMM
    assert float((fused.float() - unfused.float()).norm() / unfused.float().norm()) < 5e-3
