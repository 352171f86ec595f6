"""CPU emulation of per-ROLE rounding schemes on the oracle network (design study for the mixed default mode, not product).

Roles: an  = GroupNorm-normalised conv operands (bounded by construction)      wn = their weights
       ar  = raw conv operands (Down/Upsample, nin_shortcut, image, latent...)  wr = their weights
       sh  = storage of the residual stream (conv2 / proj_out / resample outputs)
       st  = storage of block-internal raw tensors (conv1 output, q/k/v, attention output)
"""
import sys, math
sys.path.insert(0, '/root/repo')
import torch, torch.nn.functional as F
from oracle import eovae_oracle as O
from oracle.weights import *


def make(an, wn, ar, wr, sh, st):
    def conv(sd, p, x, stride=1, padding=1, res=None, norm=True, store=None):
        rA, rW = (an, wn) if norm else (ar, wr)
        y = F.conv2d(rA(x), rW(sd[p + '.weight']), sd[p + '.bias'], stride=stride, padding=padding)
        if res is not None: y = y + res
        return (store or sh)(y)
    def gn(sd, p, x, silu=True):
        y = F.group_norm(x, 32, sd[p + '.weight'], sd[p + '.bias'], eps=1e-6)
        return y * torch.sigmoid(y) if silu else y
    def res(sd, p, x):
        h = conv(sd, p + '.conv1', gn(sd, p + '.norm1', x), store=st)
        h = gn(sd, p + '.norm2', h)
        if p + '.nin_shortcut.weight' in sd:
            # folded into conv2's K loop: x is a raw operand of the same accumulation
            sc = F.conv2d(ar(x), wr(sd[p + '.nin_shortcut.weight']), sd[p + '.nin_shortcut.bias'])
        else:
            sc = x
        return conv(sd, p + '.conv2', h, res=sc)
    def attn(sd, p, x):
        b, c, hh, ww = x.shape
        h = gn(sd, p + '.norm', x, silu=False)
        q = conv(sd, p + '.q', h, padding=0, store=st).reshape(b, c, -1).transpose(1, 2)
        k = conv(sd, p + '.k', h, padding=0, store=st).reshape(b, c, -1).transpose(1, 2)
        v = conv(sd, p + '.v', h, padding=0, store=st).reshape(b, c, -1).transpose(1, 2)
        att = torch.softmax(q @ k.transpose(1, 2) / math.sqrt(c), -1)
        o = st((st(att) @ v)).transpose(1, 2).reshape(b, c, hh, ww)
        return conv(sd, p + '.proj_out', o, padding=0, res=x, norm=False)
    def enc(sd, x, wvs, heads):
        w, bb = O.hypernet(sd, 'encoder.conv_in', wvs, False, heads)
        h = sh(F.conv2d(ar(x), wr(w), bb, padding=1))
        nlev = O._levels(sd, 'encoder.down')
        for l in range(nlev):
            for b in range(O._blocks(sd, f'encoder.down.{l}')): h = res(sd, f'encoder.down.{l}.block.{b}', h)
            if l != nlev - 1: h = conv(sd, f'encoder.down.{l}.downsample.conv', F.pad(h, (0, 1, 0, 1)), stride=2, padding=0, norm=False)
        h = res(sd, 'encoder.mid.block_1', h); h = attn(sd, 'encoder.mid.attn_1', h); h = res(sd, 'encoder.mid.block_2', h)
        h = conv(sd, 'encoder.conv_out', gn(sd, 'encoder.norm_out', h), store=lambda t: t)
        return F.conv2d(h, sd['encoder.quant_conv.weight'], sd['encoder.quant_conv.bias'])
    def dec(sd, z, wvs, heads):
        h = F.conv2d(z, sd['decoder.post_quant_conv.weight'], sd['decoder.post_quant_conv.bias'])
        h = conv(sd, 'decoder.conv_in', h, norm=False)
        h = res(sd, 'decoder.mid.block_1', h); h = attn(sd, 'decoder.mid.attn_1', h); h = res(sd, 'decoder.mid.block_2', h)
        nlev = O._levels(sd, 'decoder.up')
        for l in reversed(range(nlev)):
            for b in range(O._blocks(sd, f'decoder.up.{l}')): h = res(sd, f'decoder.up.{l}.block.{b}', h)
            if l != 0: h = conv(sd, f'decoder.up.{l}.upsample.conv', F.interpolate(h, scale_factor=2.0, mode='nearest'), norm=False)
        h = gn(sd, 'decoder.norm_out', h)
        w, bb = O.hypernet(sd, 'decoder.conv_out', wvs, True, heads)
        return F.conv2d(an(h), wn(w), bb, padding=1)
    return enc, dec


bf = lambda t: t.bfloat16().float()
hf = lambda t: t.half().float()
idt = lambda t: t
def split2(t):  # bf16 hi + bf16 lo: 16 significand bits, bf16 range
    hi = bf(t)
    return hi + bf(t - hi)
def rel(a, b): return float((a - b).norm() / b.norm())

SCHEMES = {
    #                                    an  wn  ar     wr     sh   st
    'all bf16':                         (bf, bf, bf,    bf,    bf,  bf),
    'all fp16':                         (hf, hf, hf,    hf,    hf,  hf),
    'norm fp16, raw bf16, h bf16, t f16': (hf, hf, bf,   bf,    bf,  hf),
    'norm fp16, raw bf16, h f32, t f16': (hf, hf, bf,    bf,    idt, hf),
    'norm fp16, raw split, h f32, t f16': (hf, hf, split2, split2, idt, hf),
    'norm fp16, raw fp16, h f32, t f16': (hf, hf, hf,    hf,    idt, hf),
    'norm fp16, raw fp16, h bf16, t f16': (hf, hf, hf,   hf,    bf,  hf),
    'norm fp16, raw split, h bf16, t f16': (hf, hf, split2, split2, bf, hf),
}

if __name__ == '__main__':
    for cfgname, cfg, size in (('tiny', TINY_CONFIG, 64), ('full', FULL_CONFIG, 64)):
        sd = make_state_dict(cfg, 3)
        wvs = torch.tensor(WAVELENGTHS['S2L2A']); x = synthetic_patches(2, 12, size, seed=5)
        with torch.no_grad():
            zr = O.encode_spatial_normalized(sd, x, wvs, cfg['hyper_heads']); rr = O.reconstruct(sd, x, wvs, cfg['hyper_heads'])
            for name, roles in SCHEMES.items():
                enc, dec = make(*roles)
                m = enc(sd, x, wvs, cfg['hyper_heads'])
                z = O.pixel_shuffle2(O.bn_eval(sd, O.pixel_unshuffle2(O.posterior(m)[0])))
                r = dec(sd, O.pixel_shuffle2(O.bn_inverse(sd, O.pixel_unshuffle2(z))), wvs, cfg['hyper_heads'])
                print(f'{cfgname:5s} {name:38s} latent {rel(z, zr):.2e} recon {rel(r, rr):.2e}', flush=True)
