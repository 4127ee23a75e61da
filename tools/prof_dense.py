"""A few launches of the own dense kernel at the step's largest shapes, for `ncu --set full`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from htd_b200 import _lib, dense  # noqa: E402

BF16 = torch.bfloat16
dev = 'cuda'
g = torch.Generator().manual_seed(0)


def rnd(*s):
    return torch.randn(*s, generator=g).to(BF16).to(dev)


which = sys.argv[1] if len(sys.argv) > 1 else 'all'
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
if which in ('all', 'nt'):
    M, N, K = 1024, 1024, 12544
    A, B = rnd(M, K), rnd(N, K)
    D = torch.empty(M, N, dtype=BF16, device=dev)
    for _ in range(2):
        flush.fill_(1.0)
        dense.gemm(_lib.DENSE_NT, A, B, D, M=M, N=N, K=K, lda=K, ldb=K, ldd=N)
if which in ('all', 'nn'):
    M, N, K = 1024, 12544, 1024
    A, B = rnd(M, K), rnd(K, N)
    D = torch.empty(M, N, dtype=BF16, device=dev)
    for _ in range(2):
        flush.fill_(1.0)
        dense.gemm(_lib.DENSE_NN, A, B, D, M=M, N=N, K=K, lda=K, ldb=N, ldd=N)
if which in ('all', 'conv'):
    P, Cin, Cout = 256, 576, 576
    x = rnd(P, 7, 7, Cin)
    w = 0.03 * rnd(Cout, 3, 3, Cin)
    dy = rnd(P, 7, 7, Cout)
    y = torch.empty(P, 7, 7, Cout, dtype=BF16, device=dev)
    dw = torch.empty(Cout, 3, 3, Cin, dtype=BF16, device=dev)
    for _ in range(2):
        flush.fill_(1.0)
        dense.gemm(_lib.DENSE_CONV_FPROP, w, x, y, P=P, Cin=Cin, Cout=Cout, ldd=Cout)
        flush.fill_(1.0)
        dense.gemm(_lib.DENSE_CONV_WGRAD, dy, x, dw, P=P, Cin=Cin, Cout=Cout, ldd=9 * Cin)
torch.cuda.synchronize()
