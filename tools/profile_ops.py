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
import collections
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if not ev.name.startswith('aten::') or ev.device_time_total <= 0 and ev.self_device_time_total <= 0:
        continue
    if ev.self_device_time_total <= 0:
        continue
    frames = [f for f in (ev.stack or []) if 'htd_b200' in f or 'oracle' in f or 'synth' in f]
    site = ' <- '.join(f.split('/')[-1].strip() for f in frames[:3]) or '(autograd engine / library)'
    k = (ev.name, site)
    agg[k][0] += 1
    agg[k][1] += ev.self_device_time_total
tot = sum(v[1] for v in agg.values())
print(f'ATen ops with device time: {sum(v[0] for v in agg.values())} launches, {tot:.0f} us')
for (name, site), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:120]:
    print(f'{n:4d} {t:8.1f} us  {name:28s} {site[:200]}')
