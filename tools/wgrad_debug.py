import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "eo-vae_b200"))
import torch, torch.nn.functional as F
from eo_vae import ops
dev = torch.device("cuda:0")
k = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n, h, w, cin, cout = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else '2,32,32,64,128').split(',')]
g = torch.Generator().manual_seed(0)
x = torch.randn((n, cin, h, w), generator=g).to(dev).to(dtype=torch.bfloat16, memory_format=torch.channels_last)
dy = torch.randn((n, cout, h, w), generator=g).to(dev).to(dtype=torch.bfloat16, memory_format=torch.channels_last)
import time
t0 = time.time()
try:
    dw = ops.conv2d_wgrad(x, dy, k)
    torch.cuda.synchronize()
finally:
    print('elapsed', round(time.time() - t0, 2))
wgt = torch.zeros(cout, cin, k, k, device=dev, requires_grad=True)
y = F.conv2d(x.float(), wgt, padding=k // 2); y.backward(dy.float())
print("rel err", float((dw - wgt.grad).norm() / wgt.grad.norm()))
