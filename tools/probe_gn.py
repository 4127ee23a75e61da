import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from htd_b200 import ops
torch.backends.cudnn.enabled = False
g = torch.Generator().manual_seed(0)
N, C, G = 24, 576, 36
x = torch.randn(N, C, 7, 7, generator=g)
w = 1.0 + 0.05 * torch.randn(C, generator=g); b = 0.05 * torch.randn(C, generator=g)
for name, dy in (('random', torch.randn(N, C, 7, 7, generator=g)),
                 ('spatially-constant', torch.randn(N, C, 1, 1, generator=g).expand(N, C, 7, 7).contiguous()),
                 ('constant*1e-6', 1e-6 * torch.randn(N, C, 1, 1, generator=g).expand(N, C, 7, 7).contiguous())):
    xo, wo, bo = (t.double().requires_grad_(True) for t in (x, w, b))
    go = torch.autograd.grad((torch.relu(F.group_norm(xo, G, wo, bo)) * dy.double()).sum(), [xo, wo, bo])
    def rel(a, want):
        return float((a.double().cpu() - want).abs().max() / want.abs().max())
    xa, wa, ba = (t.cuda().requires_grad_(True) for t in (x, w, b))
    ga = torch.autograd.grad((torch.relu(F.group_norm(xa.contiguous(memory_format=torch.channels_last), G, wa, ba)) * dy.cuda()).sum(), [xa, wa, ba])
    xm, wm, bm = (t.cuda().requires_grad_(True) for t in (x, w, b))
    gm = torch.autograd.grad((ops.group_norm_relu(xm.contiguous(memory_format=torch.channels_last), wm, bm, G) * dy.cuda()).sum(), [xm, wm, bm])
    print(name, 'aten', ['%.1e' % rel(a, o) for a, o in zip(ga, go)], 'mine', ['%.1e' % rel(a, o) for a, o in zip(gm, go)], flush=True)
