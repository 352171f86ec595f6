import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "eo-vae_b200"))
import torch
import eo_vae
import __graft_entry__ as ge
from oracle import eovae_oracle as O
from oracle.weights import TINY_CONFIG, WAVELENGTHS, make_state_dict, synthetic_patches
from eo_vae.models.modules.consistency_loss import EOConsistencyLoss
cuda = torch.device("cuda:0")
eo_vae.set_compute_dtype(torch.float16)
cfg = TINY_CONFIG
sd = make_state_dict(cfg, 3)
model = ge._model(cfg, sd, cuda); model.train()
wvs = torch.tensor(WAVELENGTHS["S2L2A"], dtype=torch.float32)
x = synthetic_patches(2, 12, cfg["resolution"], seed=11)
loss_fn = EOConsistencyLoss(pixel_weight=1.0, rec_loss_type="char").to(cuda)
torch.manual_seed(1234)
recon, post = model(x.to(cuda), wvs.to(cuda))
loss, _ = loss_fn(inputs=x.to(cuda), wvs=wvs.to(cuda), reconstructions=recon, global_step=0)
(loss * 4096).backward()
ref = {k: (v.clone().float().requires_grad_(True) if v.is_floating_point() else v.clone()) for k, v in sd.items()}
torch.manual_seed(1234)
eps = torch.randn((2, cfg["z_channels"], 16, 16))
recon_ref, mom_ref = O.forward(ref, x, wvs, eps, train=True, heads=cfg["hyper_heads"])
mom_ref.retain_grad()
O.charbonnier_loss(recon_ref, x).backward()
def rel(a, b): return float((a - b).norm() / (b.norm() + 1e-30))
for name in ("encoder.conv_out.weight", "encoder.quant_conv.weight", "encoder.conv_out.bias", "encoder.quant_conv.bias",
             "encoder.norm_out.weight", "encoder.mid.block_2.conv2.weight", "decoder.post_quant_conv.weight", "decoder.conv_in.weight"):
    g = dict(model.named_parameters())[name].grad.float().cpu() / 4096
    r = ref[name].grad
    print(name, "rel", round(rel(g, r), 4), "norm", float(r.norm()))
    if g.dim() == 4:
        per_out = [(round(rel(g[o], r[o]), 3), round(float(r[o].norm()), 6)) for o in range(min(g.shape[0], 16))]
        print("   per out channel (rel, norm):", per_out)
print("moments (fwd) rel:", rel(post.parameters.float().cpu(), mom_ref.detach()))
