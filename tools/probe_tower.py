"""fp32 accuracy of the regression conv tower (conv -> GN -> ReLU x3, conv -> ReLU, mean) on the GPU
against an fp64 CPU run of the same modules: which variant (fused GN, channels-last, cuDNN)
accounts for the deviation of the weight gradients."""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import copy
import torch
import torch.nn as nn
import htd_b200
from htd_b200 import synth
from htd_b200.bbox_heads import ConvModule

torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(0)
head = htd_b200.build_htd_roi_head()
synth.fill_params_(head, 'n005', 0)
tower = head.bbox_head[1].convs
x0 = torch.randn(32, 256, 7, 7)
w = torch.randn(32, 1024)


def run(dev, dt, cl, fused, cudnn):
    torch.backends.cudnn.enabled = cudnn
    ConvModule.fused_gn = fused
    t = copy.deepcopy(tower).to(dt).to(dev)
    if not cl:
        for m in t:
            m.conv.weight.data = m.conv.weight.data.contiguous(memory_format=torch.contiguous_format)
    x = x0.to(dt).to(dev).requires_grad_(True)
    xi = x.contiguous(memory_format=torch.channels_last) if cl else x
    y = t(xi).mean((2, 3))
    (y * w.to(dt).to(dev)).sum().backward()
    out = {'y': y.detach().cpu().double(), 'dx': x.grad.cpu().double()}
    for k, p in t.named_parameters():
        out[k] = p.grad.cpu().double()
    return out


ref = run('cpu', torch.float64, False, False, False)
for cl, fused, cudnn in itertools.product((True, False), (True, False), (False, True)):
    got = run('cuda', torch.float32, cl, fused, cudnn)
    errs = {k: float((got[k] - ref[k]).abs().max() / ref[k].abs().max()) for k in ref}
    print(f'channels_last={cl!s:5} fused_gn={fused!s:5} cudnn={cudnn!s:5} ' +
          ' '.join(f'{k.replace(".weight", ".w").replace("conv.", "c.")}={e:.1e}' for k, e in errs.items()))
