"""Per-kernel timing of the RoIAlign / BA kernels at the BASELINE sizes (CUDA events, L2 flushed
between iterations) with the algorithmic-byte accounting of SURVEY 8(d).  Prints one JSON line
per kernel.  Usage: python tools/bench_kernels.py [--imgs 2] [--rois 512] [--pos 128] [--iters 20]
"""
import argparse
import json
import os

os.environ.setdefault('HTD_B200_HOOKS', '1')   # variant switches live in the hooks build only
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from htd_b200 import _lib, ops, synth  # noqa: E402

PEAK = 6543.7
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                       'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:
    pass


def timeit(fn, iters, flush):
    ts = []
    for i in range(iters + 3):
        flush.fill_(float(i))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--imgs', type=int, default=2)
    ap.add_argument('--rois', type=int, default=512)
    ap.add_argument('--pos', type=int, default=128)
    ap.add_argument('--iters', type=int, default=20)
    a = ap.parse_args()
    dev = 'cuda'
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    pyr = synth.make_pyramid(a.imgs)[:4]
    props = synth.make_proposals(a.imgs, a.rois)
    rois = torch.cat([torch.cat([p.new_full((p.size(0), 1), i), p], 1)
                      for i, p in enumerate(props)]).to(dev)
    pos_rois = torch.cat([torch.cat([p.new_full((a.pos, 1), i), p[:a.pos]], 1)
                          for i, p in enumerate(props)]).to(dev)
    scales = [0.25, 0.125, 0.0625, 0.03125]
    C, PP = 256, 49
    for dtype in (torch.bfloat16, torch.float32):
        bs = 2 if dtype == torch.bfloat16 else 4
        x = [ops.to_channels_last(t.to(dev), dtype) for t in pyr]
        lv = ops.level_assign(rois, 4)
        K, P = rois.shape[0], pos_rois.shape[0]
        # ---- single level
        plan = ops.RoIPlan(x, scales, rois, lv, 7, 0)
        px = plan.pixels()
        med, best = timeit(lambda: ops.RoIPlan(x, scales, rois, lv, 7, 0), a.iters, flush)
        print(json.dumps(dict(kernel='roi_plan(single: footprints+scan+tables)', K=K, ms=med,
                              ms_best=best)))
        by_f = px * C * bs + K * PP * C * bs + 20 * K
        by_b = K * PP * C * bs + px * C * 4
        out = torch.empty(K, 7, 7, C, device=dev, dtype=dtype)
        g = torch.randn(K, 7, 7, C, device=dev).to(dtype)
        shapes = [tuple(t.shape) for t in x]
        med, best = timeit(lambda: ops._fwd_launch('f', x, scales, rois, lv, 7, 0, None, out, plan=plan),
                           a.iters, flush)
        print(json.dumps(dict(kernel='roi_align_fwd(single)', dtype=str(dtype), K=K, ms=med,
                              ms_best=best, alg_MB=by_f / 1e6, GBs=by_f / med / 1e6,
                              frac=by_f / med / 1e6 / PEAK, px_per_roi=px / K)))
        med, best = timeit(lambda: ops._roi_align_bwd(shapes, dtype, scales, rois, plan.tensors(), 7,
                                                      g, False), a.iters, flush)
        dxb = sum(s[0] * s[2] * s[3] for s in shapes) * C * bs
        print(json.dumps(dict(kernel='roi_align_bwd(single)', dtype=str(dtype), K=K, ms=med,
                              ms_best=best, alg_MB=by_b / 1e6, GBs=by_b / med / 1e6,
                              frac=by_b / med / 1e6 / PEAK, dx_write_MB=dxb / 1e6,
                              dx_write_GBs=dxb / med / 1e6)))
        # ---- BA (all levels)
        plan = ops.RoIPlan(x, scales, pos_rois, None, 7, 0)
        px = plan.pixels()
        bx = plan.boxes.long()
        cnt = ((bx[..., 1] - bx[..., 0] + 1).clamp(min=0) * (bx[..., 3] - bx[..., 2] + 1).clamp(min=0)).sum(1)
        med, best = timeit(lambda: ops.RoIPlan(x, scales, pos_rois, None, 7, 0), a.iters, flush)
        print(json.dumps(dict(kernel='roi_plan(BA: footprints+scan+tables)', K=P, ms=med,
                              ms_best=best)))
        by_f = px * C * bs + 4 * P * PP * C * bs + 20 * P
        outb = torch.empty(4, P, 7, 7, C, device=dev, dtype=dtype)
        med, best = timeit(lambda: ops._fwd_launch('f', x, scales, pos_rois, None, 7, 0, None, outb,
                                                   plan=plan), a.iters, flush)
        print(json.dumps(dict(kernel='roi_align_fwd(BA all levels)', dtype=str(dtype), K=P, ms=med,
                              ms_best=best, alg_MB=by_f / 1e6, GBs=by_f / med / 1e6,
                              frac=by_f / med / 1e6 / PEAK, px_per_roi=px / P,
                              px_levels=cnt.tolist())))
        gp = torch.randn(P, 7, 7, C, device=dev).to(dtype)
        wts = torch.rand(4, P, device=dev)
        dm = torch.randn(4 * P, C, device=dev)
        by_b = P * PP * C * bs + px * C * 4
        med, best = timeit(lambda: ops._roi_align_bwd(shapes, dtype, scales, pos_rois, plan.tensors(),
                                                      7, gp, False, scale=wts, ring_edge=1, addvec=dm),
                           a.iters, flush)
        print(json.dumps(dict(kernel='roi_align_bwd(BA all levels)', dtype=str(dtype), K=P, ms=med,
                              ms_best=best, alg_MB=by_b / 1e6, GBs=by_b / med / 1e6,
                              frac=by_b / med / 1e6 / PEAK, dx_write_MB=dxb / 1e6,
                              dx_write_GBs=dxb / med / 1e6)))
        # ---- layout conversion of the pyramid (NCHW fp32 -> channels-last dtype)
        src = [t.to(dev) for t in pyr]
        nb = sum(t.numel() for t in src) * (4 + bs)
        med, best = timeit(lambda: [ops.to_channels_last(t, dtype) for t in src], a.iters, flush)
        print(json.dumps(dict(kernel='layout_convert(pyramid)', dtype=str(dtype), ms=med,
                              ms_best=best, alg_MB=nb / 1e6, GBs=nb / med / 1e6,
                              frac=nb / med / 1e6 / PEAK)))

        # ---- flatten-order change of the pooled maps ([P,49,C] <-> [P,C,49])
        for (R, S) in ((49, C), (C, 49)):
            a_ = torch.randn(1024, R, S, device=dev).to(dtype)
            b_ = torch.empty(1024, S, R, device=dev, dtype=dtype)
            nb = 2 * a_.numel() * bs
            med, best = timeit(lambda: ops._convert(a_, b_, 1024, R, S), a.iters, flush)
            print(json.dumps(dict(kernel=f'layout_convert(pooled {R}x{S})', dtype=str(dtype), ms=med,
                                  ms_best=best, alg_MB=nb / 1e6, GBs=nb / med / 1e6,
                                  frac=nb / med / 1e6 / PEAK)))


if __name__ == '__main__':
    main()
