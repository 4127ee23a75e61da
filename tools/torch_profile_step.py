"""torch.profiler table of one steady-state bench step (device time per ATen op with input shapes):
which of the non-own kernels (ATen elementwise / reduce / cuBLAS / cuDNN) the step still spends on.
usage: python tools/torch_profile_step.py [bf16|fp32]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import htd_b200
from htd_b200 import synth

dt = torch.bfloat16 if (len(sys.argv) < 2 or sys.argv[1] == 'bf16') else torch.float32
IMGS, ROIS, POS = 2, 512, 128
head = htd_b200.build_htd_roi_head()
synth.fill_params_(head, 'init', 0)
head = head.cuda().to(dt)
head.compute_dtype = dt
x = [t.cuda().requires_grad_(True) for t in synth.make_pyramid(IMGS)]
props_h = synth.make_proposals(IMGS, ROIS)
props = [p.cuda() for p in props_h]
gts = [{k: v.cuda() for k, v in g.items()} for g in synth.make_gt(IMGS, props_h, num_pos=POS)]
shapes = [(800, 1333, 3)] * IMGS


def step():
    for p in head.parameters():
        p.grad = None
    for t in x:
        t.grad = None
    losses = synth.sampled_forward_train(head, x, props, gts, shapes, POS)
    sum(v for k, v in losses.items() if 'loss' in k).backward()


for _ in range(5):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages(group_by_input_shape=True).table(sort_by='self_cuda_time_total', row_limit=70,
                                                         max_name_column_width=48, max_shapes_column_width=70))
