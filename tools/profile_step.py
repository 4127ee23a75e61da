"""One steady-state bench step between cudaProfilerStart/Stop (for `ncu --profile-from-start off`),
plus wall-clock vs CUDA-event time of a step (launch-bound or not).
usage: python tools/profile_step.py [bf16|fp32] [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import htd_b200
from htd_b200 import synth, _lib

dt = torch.bfloat16 if (len(sys.argv) < 2 or sys.argv[1] == 'bf16') else torch.float32
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
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
t0 = time.perf_counter()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    step()
b.record()
cpu_issue = (time.perf_counter() - t0) / 5
torch.cuda.synchronize()
print('step: cuda-event %.3f ms, cpu issue time %.3f ms, own launches/step %d' %
      (a.elapsed_time(b) / 5, cpu_issue * 1e3, _lib.LAUNCHES['total'] // 10), flush=True)
torch.cuda.profiler.start()
for _ in range(nsteps):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
