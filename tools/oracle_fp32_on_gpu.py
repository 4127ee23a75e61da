"""How far is the ORACLE itself, run in fp32 on the GPU (PyTorch eager CUDA ops, torchvision CUDA
roi_align), from its fp64 CPU run on the TRAIN_ASSIGNED case - the fp32 noise floor of the library
ops on this hardware, against which the product's fp32 gradients are judged (test_gpu_assign.py)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torchvision.ops import roi_align
from htd_b200 import synth
from oracle import cases, restate

torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.enabled = False
c = cases.TRAIN_ASSIGNED


def run(dtype, device):
    own = restate.HTDRoIHead()
    synth.fill_params_(own, c['scheme'], c['wseed'])
    own = own.to(dtype).to(device)
    return cases.run_train_assigned(lambda h, *a: h.forward_train_assigned(*a), own, dtype, device)


want = run(torch.float64, 'cpu')
orig = restate.RoIAlign.forward
restate.RoIAlign.forward = lambda self, x, rois: roi_align(
    x, rois.to(x.dtype), self.output_size, self.spatial_scale, self.sampling_ratio, self.aligned)
try:
    got = run(torch.float32, 'cuda')
finally:
    restate.RoIAlign.forward = orig
errs = {k: cases.rel_err(got[k], w) for k, w in want.items() if w.is_floating_point()}
for k, e in sorted(errs.items(), key=lambda x: -x[1])[:25]:
    print(f'{k:60s} {e:.2e}')
