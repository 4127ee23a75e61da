"""torch.profiler view of ONE eager training step (bench workload): which ATen ops (name, input
shapes) launch the library / elementwise kernels around the own kernels, and where from."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import htd_b200  # noqa: E402
from htd_b200 import synth  # noqa: E402

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


def step():
    for p in head.parameters():
        p.grad = None
    for t in x:
        t.grad = None
    losses = synth.sampled_forward_train(head, x, props, gts, shapes, POS)
    sum(v for k, v in losses.items() if 'loss' in k).backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True,
             with_stack=True) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages(group_by_input_shape=True).table(sort_by='cuda_time_total', row_limit=70,
                                                          max_name_column_width=48,
                                                          max_shapes_column_width=70))
print(prof.key_averages(group_by_stack_n=6).table(sort_by='cuda_time_total', row_limit=60,
                                                  max_name_column_width=40, max_src_column_width=90))
