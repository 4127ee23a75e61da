"""PGraph stress (BASELINE config 4): all N RoIs of one image forced onto one FPN level, so the
graph is ONE dense N x N group; times the forward aggregation chain (A_local X, sam sam^T,
softmax, A_g Xm, graph Linear) and forward+backward on the tcgen05 path, and reports TFLOP/s
against the measured bf16 peak.  Prints one JSON line per N."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from htd_b200 import _lib, ops, pgraph

def peak():
    try:
        p = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))
        return p['bf16_tflops'], p.get('bf16_tflops_sustained', p['bf16_tflops'])
    except Exception:
        return 1590.0, 1400.0

def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--sizes', default='256,512,1000,2048,4096')
    ap.add_argument('--dtype', default='bf16')
    a = ap.parse_args()
    dt = torch.bfloat16 if a.dtype == 'bf16' else torch.float32
    burst, sust = peak()
    g = torch.Generator().manual_seed(0)
    d, ds = 1024, 1025
    for N in [int(s) for s in a.sizes.split(',')]:
        s = 112 + 100 * torch.rand(N, generator=g)          # sqrt(w*h) in [112, 224): level 1
        cx, cy = 1333 * torch.rand(N, generator=g), 800 * torch.rand(N, generator=g)
        rois = torch.stack([torch.zeros(N), cx - s / 2, cy - s / 2, cx + s / 2, cy + s / 2], 1).cuda()
        lv = ops.level_assign(rois, 4)
        assert int((lv == 1).sum()) == N
        x = torch.randn(N, d, generator=g).cuda().to(dt).requires_grad_(True)
        sam = (0.3 * torch.randn(N, ds, generator=g)).cuda().to(dt).requires_grad_(True)
        W = [(0.02 * torch.randn(d, d, generator=g)).cuda().to(dt).requires_grad_(True) for _ in range(4)]
        b = [torch.zeros(d, device='cuda', dtype=dt, requires_grad=True) for _ in range(4)]
        dy = torch.randn(N, d, generator=g).cuda().to(dt)
        plan = pgraph.GraphPlan(rois, lv, 1, 4, dt, max_group=N)
        fwd_flops = plan.flops()
        def fwd():
            with torch.no_grad():
                return pgraph.pgraph_refine(x, sam, W, b, plan)
        def fwdbwd():
            out = pgraph.pgraph_refine(x, sam, W, b, plan)
            torch.autograd.grad((out * dy).sum(), [x, sam] + W + b)
        t_f = timeit(fwd)
        t_fb = timeit(fwdbwd)
        # the same forward as one CUDA graph: device time without host launch gaps
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fwd()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            fwd()
        t_g = timeit(gr.replay)
        # only the contractions, timed per launch
        _lib.TIMER = _lib.KernelTimer()
        fwd(); torch.cuda.synchronize()
        n, tg = _lib.TIMER.summary().get('pgraph_gemm', (0, 0.0))
        _lib.TIMER = None
        print(json.dumps(dict(N=N, dtype=a.dtype, fwd_ms=t_f, fwd_graph_ms=t_g, fwd_graph_tflops=fwd_flops / t_g / 1e9,
                              fwd_graph_frac_of_burst_peak=fwd_flops / t_g / 1e9 / burst, fwdbwd_ms=t_fb, fwd_gflop=fwd_flops / 1e9,
                              fwd_tflops=fwd_flops / t_f / 1e9, gemm_launches=n, gemm_ms=tg,
                              gemm_tflops=fwd_flops / tg / 1e9 if tg else None,
                              frac_of_burst_peak=(fwd_flops / tg / 1e9 / burst) if tg else None,
                              fwdbwd_tflops=3 * fwd_flops / t_fb / 1e9, peak_tflops=burst)), flush=True)

if __name__ == '__main__':
    main()
