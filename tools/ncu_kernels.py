"""One launch of each RoIAlign kernel (bf16, BASELINE sizes) between cudaProfilerStart/Stop, for
  ncu --profile-from-start off --set full --import-source on -o gpurun_out/prof python tools/ncu_kernels.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from htd_b200 import ops, synth

dev = 'cuda'
dtype = torch.bfloat16 if (len(sys.argv) < 2 or sys.argv[1] == 'bf16') else torch.float32
imgs, nroi, npos, C = 2, 512, 128, 256
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
pyr = synth.make_pyramid(imgs)[:4]
props = synth.make_proposals(imgs, nroi)
rois = torch.cat([torch.cat([p.new_full((p.size(0), 1), i), p], 1) for i, p in enumerate(props)]).to(dev)
pos_rois = torch.cat([torch.cat([p.new_full((npos, 1), i), p[:npos]], 1) for i, p in enumerate(props)]).to(dev)
scales = [0.25, 0.125, 0.0625, 0.03125]
x = [ops.to_channels_last(t.to(dev), dtype) for t in pyr]
shapes = [tuple(t.shape) for t in x]
lv = ops.level_assign(rois, 4)
K, P = rois.shape[0], pos_rois.shape[0]
plan_s = ops.RoIPlan(x, scales, rois, lv, 7, 0)
PLAN_IN_RUN = len(sys.argv) > 2 and sys.argv[2] == "plan"
plan_b = ops.RoIPlan(x, scales, pos_rois, None, 7, 0)
out_s = torch.empty(K, 7, 7, C, device=dev, dtype=dtype)
out_b = torch.empty(4, P, 7, 7, C, device=dev, dtype=dtype)
g_s = torch.randn(K, 7, 7, C, device=dev).to(dtype)
g_b = torch.randn(P, 7, 7, C, device=dev).to(dtype)
wts = torch.rand(4, P, device=dev)
dm = torch.randn(4 * P, C, device=dev)

def run_all():
    flush.fill_(1.0)
    if PLAN_IN_RUN:
        ops.RoIPlan(x, scales, rois, lv, 7, 0)
        ops.RoIPlan(x, scales, pos_rois, None, 7, 0)
    ops._fwd_launch('f', x, scales, rois, lv, 7, 0, None, out_s, plan=plan_s)
    flush.fill_(2.0)
    ops._roi_align_bwd(shapes, dtype, scales, rois, plan_s.tensors(), 7, g_s, False)
    flush.fill_(3.0)
    ops._fwd_launch('f', x, scales, pos_rois, None, 7, 0, None, out_b, plan=plan_b)
    flush.fill_(4.0)
    ops._roi_align_bwd(shapes, dtype, scales, pos_rois, plan_b.tensors(), 7, g_b, False, scale=wts,
                       ring_edge=1, addvec=dm)

for _ in range(3):
    run_all()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run_all()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('done')
