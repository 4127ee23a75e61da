"""Per-CTA trace of the fused bf16 backward gather at the BASELINE sizes (htd_debug_set_bwd_trace):
start / end globaltimer, hits, K-step blocks, SM and level of every tile.  Prints a least-squares
cost model  t_tile = a + b * hits + c * blocks,  the per-level means, the heaviest tiles and the
busy time of the busiest SM against the kernel span.  Usage: python tools/trace_bwd.py
"""
import json
import os

os.environ.setdefault('HTD_B200_HOOKS', '1')   # variant switches live in the hooks build only
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from htd_b200 import _lib, ops, synth  # noqa: E402


def main():
    dev, dtype = 'cuda', torch.bfloat16
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
    lv = ops.level_assign(rois, 4)
    ps = ops.RoIPlan(x, scales, rois, lv, 7, 0)
    pb = ops.RoIPlan(x, scales, pos, None, 7, 0)
    g = torch.randn(rois.shape[0], 7, 7, C, device=dev).to(dtype)
    gp = torch.randn(pos.shape[0], 7, 7, C, device=dev).to(dtype)
    single = dict(rois=rois, plan=ps.tensors(), dy=g, dy_per_level=False)
    ba = dict(rois=pos, plan=pb.tensors(), dy=gp, dy_per_level=False,
              scale=torch.rand(4, pos.shape[0], device=dev), ring_edge=1,
              addvec=torch.randn(4 * pos.shape[0], C, device=dev))
    src = [single, dict(single), ba]
    ntiles = sum(s[0] * ((s[2] + 7) // 8) * ((s[3] + 7) // 8) for s in shapes)
    rec = torch.zeros(ntiles, 6, dtype=torch.int64, device=dev)
    for _ in range(3):
        ops._bwd_multi(shapes, dtype, False, scales, src, 7)
    torch.cuda.synchronize()
    _lib.lib().htd_debug_set_bwd_trace(rec.data_ptr())
    ops._bwd_multi(shapes, dtype, False, scales, src, 7)
    torch.cuda.synchronize()
    _lib.lib().htd_debug_set_bwd_trace(None)
    r = rec.cpu().numpy().astype(np.float64)
    t0, t1, hits, blocks, sm, lvl = r.T
    dur = (t1 - t0) / 1e3                                   # us
    span = (t1.max() - t0.min()) / 1e3
    A = np.stack([np.ones_like(hits), hits, blocks], 1)
    coef, *_ = np.linalg.lstsq(A, dur, rcond=None)
    out = dict(variant=os.environ.get('HTD_BWD_KERNEL', 'default'), tiles=int(ntiles),
               span_us=span, hits=int(hits.sum()), blocks=int(blocks.sum()),
               model_us=dict(fixed=coef[0], per_hit=coef[1], per_block=coef[2]),
               tile_us=dict(mean=dur.mean(), p50=float(np.median(dur)), p99=float(np.percentile(dur, 99)),
                            max=dur.max()))
    out['per_level'] = {int(k): dict(tiles=int((lvl == k).sum()), mean_us=dur[lvl == k].mean(),
                                     max_us=dur[lvl == k].max(), mean_hits=hits[lvl == k].mean(),
                                     max_hits=hits[lvl == k].max(), mean_blocks=blocks[lvl == k].mean(),
                                     max_blocks=blocks[lvl == k].max(),
                                     first_start_us=(t0[lvl == k].min() - t0.min()) / 1e3,
                                     last_end_us=(t1[lvl == k].max() - t0.min()) / 1e3)
                        for k in np.unique(lvl)}
    busy = np.zeros(int(sm.max()) + 1)
    np.add.at(busy, sm.astype(int), dur)
    out['sm_busy_us'] = dict(mean=busy.mean(), max=busy.max(), min=busy.min())
    top = np.argsort(-dur)[:8]
    out['heaviest'] = [dict(us=dur[i], hits=int(hits[i]), blocks=int(blocks[i]), level=int(lvl[i]),
                            start_us=(t0[i] - t0.min()) / 1e3) for i in top]
    print(json.dumps(out))


if __name__ == '__main__':
    main()
