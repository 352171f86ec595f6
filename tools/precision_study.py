"""CPU emulation of storage / operand rounding schemes on the oracle network (design study, not product)."""
import sys, math
sys.path.insert(0, '/root/repo')
import torch, torch.nn.functional as F
from oracle import eovae_oracle as O
from oracle.weights import *

def make(rA, rS, rW):
    """rA: operand rounding of conv inputs, rS: storage rounding of conv outputs, rW: weight rounding."""
    def conv(sd, p, x, stride=1, padding=1, res=None):
        y = F.conv2d(rA(x), rW(sd[p + '.weight']), sd[p + '.bias'], stride=stride, padding=padding)
        if res is not None: y = y + res
        return rS(y)
    def gn(sd, p, x, silu=True):
        y = F.group_norm(x, 32, sd[p + '.weight'], sd[p + '.bias'], eps=1e-6)
        return y * torch.sigmoid(y) if silu else y
    def res(sd, p, x):
        h = conv(sd, p + '.conv1', gn(sd, p + '.norm1', x))
        h = gn(sd, p + '.norm2', h)
        sc = conv(sd, p + '.nin_shortcut', x, padding=0) if p + '.nin_shortcut.weight' in sd else x
        return conv(sd, p + '.conv2', h, res=sc)
    def attn(sd, p, x):
        b, c, hh, ww = x.shape
        h = gn(sd, p + '.norm', x, silu=False)
        q = conv(sd, p + '.q', h, padding=0).reshape(b, c, -1).transpose(1, 2)
        k = conv(sd, p + '.k', h, padding=0).reshape(b, c, -1).transpose(1, 2)
        v = conv(sd, p + '.v', h, padding=0).reshape(b, c, -1).transpose(1, 2)
        att = torch.softmax(rA(q) @ rA(k).transpose(1, 2) / math.sqrt(c), -1)
        o = rS((rA(att) @ rA(v))).transpose(1, 2).reshape(b, c, hh, ww)
        return conv(sd, p + '.proj_out', o, padding=0, res=x)
    def enc(sd, x, wvs, heads):
        w, bb = O.hypernet(sd, 'encoder.conv_in', wvs, False, heads)
        h = rS(F.conv2d(rA(x), rW(w), bb, padding=1))
        nlev = O._levels(sd, 'encoder.down')
        for l in range(nlev):
            for b in range(O._blocks(sd, f'encoder.down.{l}')): h = res(sd, f'encoder.down.{l}.block.{b}', h)
            if l != nlev - 1: h = conv(sd, f'encoder.down.{l}.downsample.conv', F.pad(h, (0, 1, 0, 1)), stride=2, padding=0)
        h = res(sd, 'encoder.mid.block_1', h); h = attn(sd, 'encoder.mid.attn_1', h); h = res(sd, 'encoder.mid.block_2', h)
        h = conv(sd, 'encoder.conv_out', gn(sd, 'encoder.norm_out', h))
        return F.conv2d(rA(h), rW(sd['encoder.quant_conv.weight']), sd['encoder.quant_conv.bias'])
    def dec(sd, z, wvs, heads):
        h = conv(sd, 'decoder.post_quant_conv', rS(z), padding=0)
        h = conv(sd, 'decoder.conv_in', h)
        h = res(sd, 'decoder.mid.block_1', h); h = attn(sd, 'decoder.mid.attn_1', h); h = res(sd, 'decoder.mid.block_2', h)
        nlev = O._levels(sd, 'decoder.up')
        for l in reversed(range(nlev)):
            for b in range(O._blocks(sd, f'decoder.up.{l}')): h = res(sd, f'decoder.up.{l}.block.{b}', h)
            if l != 0: h = conv(sd, f'decoder.up.{l}.upsample.conv', F.interpolate(h, scale_factor=2.0, mode='nearest'))
        h = gn(sd, 'decoder.norm_out', h)
        w, bb = O.hypernet(sd, 'decoder.conv_out', wvs, True, heads)
        return F.conv2d(rA(h), rW(w), bb, padding=1)
    return enc, dec

bf = lambda t: t.bfloat16().float()
hf = lambda t: t.half().float()
idt = lambda t: t
def rel(a, b): return float((a - b).norm() / b.norm())

for cfgname, cfg, size in (('tiny', TINY_CONFIG, 64), ('full', FULL_CONFIG, 64)):
    sd = make_state_dict(cfg, 3)
    wvs = torch.tensor(WAVELENGTHS['S2L2A']); x = synthetic_patches(2, 12, size, seed=5)
    with torch.no_grad():
        m_ref = O.encoder_forward(sd, x, wvs, cfg['hyper_heads'])
        zr = O.encode_spatial_normalized(sd, x, wvs, cfg['hyper_heads']); rr = O.reconstruct(sd, x, wvs, cfg['hyper_heads'])
        for name, (rA, rS, rW) in {'bf16 operands + bf16 storage': (bf, bf, bf), 'bf16 operands + fp32 storage': (bf, idt, bf),
                                   'fp16 operands + bf16 storage': (hf, bf, hf), 'fp16 operands + fp16 storage': (hf, hf, hf),
                                   'fp16 operands + fp32 storage': (hf, idt, hf)}.items():
            enc, dec = make(rA, rS, rW)
            m = enc(sd, x, wvs, cfg['hyper_heads'])
            z = O.pixel_shuffle2(O.bn_eval(sd, O.pixel_unshuffle2(O.posterior(m)[0])))
            r = dec(sd, O.pixel_shuffle2(O.bn_inverse(sd, O.pixel_unshuffle2(z))), wvs, cfg['hyper_heads'])
            print(f'{cfgname:5s} {name:32s} latent {rel(z, zr):.2e} recon {rel(r, rr):.2e}')
