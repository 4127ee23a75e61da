"""Per-shape timing of the own dense tcgen05 kernels (csrc/dense_gemm.cu) at the step's shapes
(bench workload: K = 1024 RoIs, P = 256 positives) next to the library call for the same product
(cuBLAS via torch.matmul / cuDNN via F.conv2d, bf16).  CUDA events, L2 flushed before every launch.
One JSON line per shape.  Usage: python tools/bench_dense.py [--iters 10]"""
import argparse
import json
import os

os.environ.setdefault('HTD_B200_HOOKS', '1')   # variant switches live in the hooks build only
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from htd_b200 import _lib, dense  # noqa: E402

BF16 = torch.bfloat16


def timeit(fn, iters, flush):
    """Median device time of one launch of `fn`: captured once into a CUDA graph and replayed, so
    that the host side of the call (descriptor setup, tensor-map encoding, ctypes) is not in the
    interval - an event pair around an eager call of a 50 us kernel measures mostly that."""
    fn()
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        fn()
    torch.cuda.current_stream().wait_stream(st)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        fn()
    ts = []
    for i in range(iters + 3):
        flush.fill_(float(i))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        gr.replay()
        b.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--iters', type=int, default=10)
    ap.add_argument('--splits', type=int, default=0)
    ap.add_argument('--only', default='', help='substring filter on the shape name')
    ap.add_argument('--no-lib', action='store_true')
    a = ap.parse_args()
    dev = 'cuda'
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    g = torch.Generator().manual_seed(0)

    def rnd(*s):
        return torch.randn(*s, generator=g).to(BF16).to(dev)

    def report(name, flops, own, libf):
        if a.only and a.only not in name:
            return
        if a.no_lib:
            libf = None
        t_own = timeit(own, a.iters, flush)
        t_lib = timeit(libf, a.iters, flush) if libf is not None else None
        print(json.dumps(dict(shape=name, gflop=round(flops / 1e9, 2), own_ms=round(t_own, 4),
                              own_tflops=round(flops / t_own / 1e9, 1),
                              lib_ms=None if t_lib is None else round(t_lib, 4),
                              lib_tflops=None if t_lib is None else round(flops / t_lib / 1e9, 1))),
              flush=True)

    for (M, N, K) in ((1024, 1024, 12544), (2048, 1024, 1024), (1024, 1024, 1024), (1024, 88, 1024)):
        A, B = rnd(M, K), rnd(N, K)
        D = torch.empty(M, N, dtype=BF16, device=dev)
        bias = torch.randn(N, device=dev)
        report(f'fc_fwd NT {M}x{N}x{K}', 2.0 * M * N * K,
               lambda: dense.gemm(_lib.DENSE_NT, A, B, D, M=M, N=N, K=K, lda=K, ldb=K, ldd=N, bias=bias,
                                  relu=True, splits=a.splits),
               lambda: torch.relu_(F.linear(A, B, bias.to(BF16))))
    for (M, N, K) in ((1024, 12544, 1024), (2048, 1024, 1024), (1024, 1024, 88)):
        A, B = rnd(M, K), rnd(K, N)
        D = torch.empty(M, N, dtype=BF16, device=dev)
        report(f'fc_dgrad NN {M}x{N}x{K}', 2.0 * M * N * K,
               lambda: dense.gemm(_lib.DENSE_NN, A, B, D, M=M, N=N, K=K, lda=K, ldb=N, ldd=N,
                                  splits=a.splits),
               lambda: torch.matmul(A, B))
    for (M, N, K) in ((1024, 12544, 1024), (1024, 1024, 2048), (88, 1024, 1024)):
        A, B = rnd(K, M), rnd(K, N)
        D = torch.empty(M, N, dtype=BF16, device=dev)
        report(f'fc_wgrad TN {M}x{N}x{K}', 2.0 * M * N * K,
               lambda: dense.gemm(_lib.DENSE_TN, A, B, D, M=M, N=N, K=K, lda=M, ldb=N, ldd=N,
                                  splits=a.splits),
               lambda: torch.matmul(A.t(), B))
    P = 256
    for (Cin, Cout) in ((256, 576), (576, 576), (576, 1024)):
        x = rnd(P, Cin, 7, 7).contiguous(memory_format=torch.channels_last)
        w = (0.03 * rnd(Cout, Cin, 3, 3)).contiguous(memory_format=torch.channels_last)
        dy = rnd(P, Cout, 7, 7).contiguous(memory_format=torch.channels_last)
        y = torch.empty(P, 7, 7, Cout, dtype=BF16, device=dev)
        dx = torch.empty(P, 7, 7, Cin, dtype=BF16, device=dev)
        dw = torch.empty(Cout, 3, 3, Cin, dtype=BF16, device=dev)
        fl = 2.0 * P * 49 * 9 * Cin * Cout
        report(f'conv_fprop {Cin}->{Cout} P={P}', fl,
               lambda: dense.gemm(_lib.DENSE_CONV_FPROP, w, x, y, P=P, Cin=Cin, Cout=Cout, ldd=Cout),
               lambda: F.conv2d(x, w, padding=1))
        report(f'conv_dgrad {Cin}->{Cout} P={P}', fl,
               lambda: dense.gemm(_lib.DENSE_CONV_DGRAD, w, dy, dx, P=P, Cin=Cin, Cout=Cout, ldd=Cin),
               lambda: torch.ops.aten.convolution_backward(dy, x, w, None, (1, 1), (1, 1), (1, 1), False,
                                                           (0, 0), 1, (True, False, False)))
        report(f'conv_wgrad {Cin}->{Cout} P={P}', fl,
               lambda: dense.gemm(_lib.DENSE_CONV_WGRAD, dy, x, dw, P=P, Cin=Cin, Cout=Cout, ldd=9 * Cin,
                                  splits=a.splits),
               lambda: torch.ops.aten.convolution_backward(dy, x, w, None, (1, 1), (1, 1), (1, 1), False,
                                                           (0, 0), 1, (False, True, False)))


if __name__ == '__main__':
    main()
