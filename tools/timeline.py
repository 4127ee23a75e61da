"""Kernel timeline of ONE replay of the captured step (the bench workload), from CUPTI through
torch.profiler: every kernel with its stream, start and duration, so that the idle gaps, the
branch overlap and the critical path of the graph can be read.  A profiler run is not a
bench number; the per-kernel durations are warm and concurrent, unlike the serialised ncu list.

    python tools/timeline.py --out gpurun_out/timeline.json [--imgs-per-gpu 2]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default='gpurun_out/timeline.json')
    ap.add_argument('--imgs-per-gpu', type=int, default=2)
    ap.add_argument('--dtype', default='bf16')
    a = ap.parse_args()
    import torch
    import htd_b200
    import bench
    from htd_b200 import synth
    from htd_b200.graphed import GraphedTrainStep
    dev = torch.device('cuda', 0)
    dtype = torch.bfloat16 if a.dtype == 'bf16' else torch.float32
    head = htd_b200.build_htd_roi_head()
    synth.fill_params_(head, 'init', 0)
    head = head.to(dev).to(dtype)
    head.compute_dtype = dtype
    head.train()
    imgs = a.imgs_per_gpu
    pyr = [t.to(dev).requires_grad_(True) for t in synth.make_pyramid(imgs, bench.IMG_H, bench.IMG_W, seed=1000)]
    props_h = synth.make_proposals(imgs, bench.ROIS, bench.IMG_H, bench.IMG_W, seed=1234)
    gts = synth.make_gt(imgs, props_h, num_pos=bench.POS, seed=4321)
    gts = [{k: v.to(dev) for k, v in g.items()} for g in gts]
    props = [p.to(dev) for p in props_h]
    shapes = [(bench.IMG_H, bench.IMG_W, 3)] * imgs
    for _ in range(3):
        for p in list(head.parameters()) + pyr:
            p.grad = None
        losses = synth.sampled_forward_train(head, pyr, props, gts, shapes, bench.POS)
        sum(v for k, v in losses.items() if 'loss' in k).backward()
    torch.cuda.synchronize()
    del losses                                  # drops the eager graph and its AccumulateGrad nodes
    for p in list(head.parameters()) + pyr:     # (bound to the default stream: they break a capture)
        p.grad = None
    gstep = GraphedTrainStep(head, pyr, props, gts, shapes, bench.POS)
    for _ in range(5):
        gstep()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            gstep()
        torch.cuda.synchronize()
    raw = a.out + '.chrome.json'
    prof.export_chrome_trace(raw)
    tr = json.load(open(raw))
    os.remove(raw)
    recs = [dict(name=e['name'], start_us=e['ts'], dur_us=e.get('dur', 0),
                 stream=e.get('args', {}).get('stream', 0))
            for e in tr['traceEvents'] if e.get('cat') in ('kernel', 'gpu_memcpy', 'gpu_memset')]
    recs.sort(key=lambda r: r['start_us'])
    n = len(recs) // 3                              # three identical replays
    last = recs[2 * n:]
    t0 = last[0]['start_us']
    for r in last:
        r['start_us'] = round(r['start_us'] - t0, 3)
    os.makedirs(os.path.dirname(a.out) or '.', exist_ok=True)
    json.dump(last, open(a.out, 'w'))
    end = max(r['start_us'] + r['dur_us'] for r in last)
    print(json.dumps(dict(kernels=len(last), span_us=end, busy_sum_us=sum(r['dur_us'] for r in last),
                          streams=sorted({r['stream'] for r in last}))))


if __name__ == '__main__':
    main()
