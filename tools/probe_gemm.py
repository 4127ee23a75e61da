"""Throwaway GPU probe: tcgen05 grouped GEMM on the shapes the head uses (small ld, unaligned k0, bf16 D)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['CUDA_LAUNCH_BLOCKING'] = '1'
import torch
from htd_b200 import pgraph

def run(tag, a_ld, Np, spans, d_bf16, with_dt, unaligned=True):
    g = torch.Generator().manual_seed(1)
    d = 1024
    A = torch.zeros(Np, a_ld)
    XT = torch.randn(d, Np, generator=g)
    groups = []
    for off, n in spans:
        A[off:off + n, :n] = torch.randn(n, n, generator=g)
        groups.append(dict(M=n, N=d, K=n, a_row=off, b_k0=off, d_row=off, dt_col=off))
    A, XT = A.cuda().bfloat16(), XT.cuda().bfloat16()
    D = torch.zeros(Np, d, device='cuda', dtype=torch.bfloat16 if d_bf16 else torch.float32)
    DT = torch.zeros(d, Np, device='cuda', dtype=torch.bfloat16) if with_dt else None
    pgraph._gemm(A, XT, groups, D=D, ldd=d, DT=DT, ldt=Np)
    torch.cuda.synchronize()
    err = 0.0
    for off, n in spans:
        want = A[off:off + n, :n].double() @ XT[:, off:off + n].double().t()
        err = max(err, ((D[off:off + n].double() - want).abs().max() / want.abs().max()).item())
    print(tag, 'ok err=%.2e' % err, flush=True)

run('aligned ld1088 f32', 1088, 256, [(0, 64), (64, 100)], False, False)
run('ld64 aligned f32', 64, 256, [(0, 64), (64, 64)], False, False)
run('ld64 unaligned-k0 f32', 64, 256, [(0, 12), (16, 50), (72, 37)], False, False)
run('ld64 unaligned-k0 bf16D', 64, 256, [(0, 12), (16, 50), (72, 37)], True, False)
run('ld64 unaligned-k0 bf16D+DT', 64, 256, [(0, 12), (16, 50), (72, 37)], True, True)
run('Np64 rows<box', 64, 64, [(0, 12), (16, 48)], True, True)
