"""Every ATen op of ONE eager training step (bench workload) with the repo line that issued it:
a TorchDispatchMode in the calling thread, autograd multithreading off so that backward runs in
the same thread.  Ops that launch no kernel (views, metadata) are filtered by name."""
import collections
import os
import sys
import traceback

import torch
from torch.utils._python_dispatch import TorchDispatchMode

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import htd_b200  # noqa: E402
from htd_b200 import synth  # noqa: E402

VIEWS = {'view', 'reshape', '_unsafe_view', 'permute', 'transpose', 't', 'slice', 'select', 'expand',
         'unsqueeze', 'squeeze', 'detach', 'alias', 'as_strided', 'narrow', 'split', 'unbind',
         'split_with_sizes', 'empty', 'empty_like', 'empty_strided', 'new_empty', 'size', 'stride',
         'is_contiguous', '_local_scalar_dense', 'lift_fresh', 'new_empty_strided', 'unfold',
         'contiguous', 'flatten', 'view_as', 'result_type', 'set_', 'record_stream'}


class Tracer(TorchDispatchMode):
    def __init__(self):
        super().__init__()
        self.log = collections.Counter()

    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        name = func.__name__.split('.')[0]
        if name not in VIEWS:
            fr = [f for f in traceback.extract_stack() if ROOT in f.filename and 'tools/' not in f.filename]
            site = ' <- '.join(f'{os.path.basename(f.filename)}:{f.lineno}' for f in reversed(fr[-3:]))
            shp = next((tuple(a.shape) for a in args if isinstance(a, torch.Tensor)), None)
            self.log[(name, site or '(engine)', str(shp))] += 1
        return func(*args, **(kwargs or {}))


dev = torch.device('cuda')
head = htd_b200.build_htd_roi_head()
synth.fill_params_(head, 'init', 0)
head = head.to(dev).to(torch.bfloat16)
head.compute_dtype = torch.bfloat16
head.train()
IMGS, ROIS, POS = 2, 512, 128
pyr = synth.make_pyramid(IMGS)
props_h = synth.make_proposals(IMGS, ROIS)
gts = [{k: v.to(dev) for k, v in g.items()} for g in synth.make_gt(IMGS, props_h, num_pos=POS)]
x = [t.to(dev).requires_grad_(True) for t in pyr]
props = [p.to(dev) for p in props_h]
shapes = [(800, 1333, 3)] * IMGS
torch.autograd.set_multithreading_enabled(False)


def step():
    for p in head.parameters():
        p.grad = None
    for t in x:
        t.grad = None
    losses = synth.sampled_forward_train(head, x, props, gts, shapes, POS)
    sum(v for k, v in losses.items() if 'loss' in k).backward()


step()
tr = Tracer()
with tr:
    step()
torch.cuda.synchronize()
print('ops:', sum(tr.log.values()))
for (name, site, shp), n in sorted(tr.log.items(), key=lambda kv: (kv[0][1], kv[0][0])):
    print(f'{n:3d} {name:24s} {shp:26s} {site}')
