"""Cost model of the fused backward gather (htd_roi_align_bwd_multi) at the BASELINE sizes:
the real three-source launch of a training step, the same launch with every RoI moved to a
non-existent image (no hit anywhere: scan + zero-fill only = the fixed per-tile cost), and each
source alone.  CUDA events, L2 flushed between iterations.  One JSON line per case.
Usage: [HTD_BWD_KERNEL=mma3] python tools/bench_bwd_cost.py [--iters 20]
"""
import argparse
import json
import os

os.environ.setdefault('HTD_B200_HOOKS', '1')   # variant switches live in the hooks build only
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from htd_b200 import ops, synth  # noqa: E402
from tools.bench_kernels import timeit  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--iters', type=int, default=20)
    a = ap.parse_args()
    dev, dtype = 'cuda', torch.bfloat16
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    pyr = synth.make_pyramid(2)[:4]
    props = synth.make_proposals(2, 512)
    rois = torch.cat([torch.cat([p.new_full((p.size(0), 1), i), p], 1)
                      for i, p in enumerate(props)]).to(dev)
    pos = torch.cat([torch.cat([p.new_full((128, 1), i), p[:128]], 1)
                     for i, p in enumerate(props)]).to(dev)
    scales = [0.25, 0.125, 0.0625, 0.03125]
    x = [ops.to_channels_last(t.to(dev), dtype) for t in pyr]
    shapes = [tuple(t.shape) for t in x]
    C = 256

    def sources(r, p):
        lv = ops.level_assign(r, 4)
        ps = ops.RoIPlan(x, scales, r, lv, 7, 0)
        pb = ops.RoIPlan(x, scales, p, None, 7, 0)
        g = torch.randn(r.shape[0], 7, 7, C, device=dev).to(dtype)
        gp = torch.randn(p.shape[0], 7, 7, C, device=dev).to(dtype)
        single = dict(rois=r, plan=ps.tensors(), dy=g, dy_per_level=False)
        ba = dict(rois=p, plan=pb.tensors(), dy=gp, dy_per_level=False,
                  scale=torch.rand(4, p.shape[0], device=dev), ring_edge=1,
                  addvec=torch.randn(4 * p.shape[0], C, device=dev))
        return single, ba, (ps, pb)

    single, ba, keep = sources(rois, pos)
    far_r, far_p = rois.clone(), pos.clone()
    far_r[:, 0] = 9
    far_p[:, 0] = 9
    esingle, eba, keep2 = sources(far_r, far_p)
    cases = {
        'fused(step: single + single + BA)': [single, dict(single), ba],
        'fused, no hits (fixed cost)': [esingle, dict(esingle), eba],
        'single only': [single],
        'single only, no hits': [esingle],
        'BA only': [ba],
        'BA only, no hits': [eba],
    }
    for name, src in cases.items():
        med, best = timeit(lambda: ops._bwd_multi(shapes, dtype, False, scales, src, 7), a.iters,
                           flush)
        print(json.dumps(dict(case=name, variant=os.environ.get('HTD_BWD_KERNEL', 'default'),
                              ms=med, ms_best=best)))


if __name__ == '__main__':
    main()
