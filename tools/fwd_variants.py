"""Times the level-assigned forward extraction (plan + gather, CUDA-graph replay, cold L2) for every
library given on the command line (tuning builds of tools/build_fwd_variants.sh).  One subprocess
per library (HTD_B200_LIB).  usage: python tools/fwd_variants.py lib1.so lib2.so ..."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def worker():
    import torch
    sys.path.insert(0, ROOT)
    from htd_b200 import ops, synth
    dev = 'cuda'
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    res = {}
    scales = [0.25, 0.125, 0.0625, 0.03125]
    for rois_per_img in (512, 2048):
        pyr = synth.make_pyramid(2)[:4]
        props = synth.make_proposals(2, rois_per_img)
        rois = torch.cat([torch.cat([p.new_full((p.size(0), 1), i), p], 1)
                          for i, p in enumerate(props)]).to(dev)
        x = [ops.to_channels_last(t.to(dev), torch.bfloat16) for t in pyr]
        lv = ops.level_assign(rois, 4)
        K = rois.shape[0]
        out = torch.empty(K, 7, 7, 256, device=dev, dtype=torch.bfloat16)

        def run():
            ops._fwd_launch('f', x, scales, rois, lv, 7, 0, None, out)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            run()
        ts = []
        for i in range(25):
            flush.fill_(float(i))
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            g.replay()
            b.record()
            torch.cuda.synchronize()
            if i >= 5:
                ts.append(a.elapsed_time(b))
        ts.sort()
        res[K] = round(ts[len(ts) // 2], 4)
    print(json.dumps(res))


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == '--worker':
        worker()
    else:
        for lib in sys.argv[1:]:
            env = dict(os.environ, HTD_B200_LIB=os.path.abspath(lib))
            r = subprocess.run([sys.executable, os.path.abspath(__file__), '--worker'], env=env,
                               capture_output=True, text=True)
            print(os.path.basename(lib), r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-400:],
                  flush=True)
