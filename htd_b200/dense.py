"""Dense bf16 layers of the head on the package's own tcgen05 kernels (csrc/dense_gemm.cu,
SURVEY.md §8 row f1): the FC stacks (``convfc_bbox_head.py:141-148``, ``htd_bbox_head.py:114-121,
191-192,227-228``) and the 3x3 regression conv tower on 7x7 RoI maps (``htd_bbox_head.py:75-113,
186``) - forward, data gradient and weight gradient - which the reference runs through cuBLAS /
cuDNN.  bf16 CUDA tensors only: the fp32 parity configuration keeps the library calls (exact fp32
FC / conv arithmetic is not what the tensor cores provide), and there is no CPU path.

    gemm(kind, ...)                 one contraction, the C-ABI call
    linear(x, w, b, relu)           y = act(x w^T + b) with own forward / dgrad / wgrad / bias-grad
    conv3x3(x, w)                   channels-last [P,C,7,7] -> [P,Cout,7,7], stride 1, padding 1
"""
import ctypes

import torch

from . import _lib
from ._lib import check, lib, stream

BF16 = torch.bfloat16


def usable(*tensors):
    """The own dense kernels apply: CUDA, bf16, nothing else (callers fall back to the library
    call for the fp32 parity configuration)."""
    return all(t is not None and t.is_cuda and t.dtype == BF16 for t in tensors)


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


_WS = {}


def _workspace(nbytes, device):
    """Split-K partial sums: one growing scratch buffer per device and stream (the kernels of one
    stream run in order, so consecutive GEMMs can share it; CUDA-graph capture keeps it alive)."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _WS.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = _WS[key] = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
    return buf


def gemm(kind, A, B, D, M=0, N=0, K=0, lda=0, ldb=0, ldd=0, bias=None, relu=False, gate=None, ldg=0,
         D2=None, row_bias=None, row_class=None, P=0, Cin=0, Cout=0, splits=0, name='dense_gemm'):
    """One ``htd_dense_gemm`` call (see include/htd_b200.h for the six kinds)."""
    _lib.require_cuda(A, B, D)
    g = _lib.HtdDenseGemm()
    g.kind, g.M, g.N, g.K = int(kind), int(M), int(N), int(K)
    g.P, g.Cin, g.Cout, g.pooled = int(P), int(Cin), int(Cout), 7
    g.d_dtype, g.relu, g.splits = _lib.dt(D), int(bool(relu)), int(splits)
    g.bias_dtype = _lib.dt(bias) if bias is not None else 0
    g.A, g.B, g.D = A.data_ptr(), B.data_ptr(), D.data_ptr()
    g.D2 = D2.data_ptr() if D2 is not None else None
    g.bias = bias.data_ptr() if bias is not None else None
    g.row_bias = row_bias.data_ptr() if row_bias is not None else None
    g.row_class = row_class.data_ptr() if row_class is not None else None
    g.gate = gate.data_ptr() if gate is not None else None
    g.lda, g.ldb, g.ldd, g.ldg = int(lda), int(ldb), int(ldd), int(ldg)
    g.ld_row_bias = int(row_bias.stride(0)) if row_bias is not None else 0
    need = int(lib().htd_dense_gemm_workspace_bytes(ctypes.byref(g)))
    ws = _workspace(need, D.device)
    if _lib.ACCOUNT is not None:
        flops = 2.0 * M * N * K if kind <= _lib.DENSE_TN else 2.0 * P * 49 * 9 * Cin * Cout
        _lib.ACCOUNT.append((name + ':flops', flops))
    with _lib.timed(name):
        check(lib().htd_dense_gemm(ctypes.byref(g), _p(ws), ws.numel(), stream()), 'htd_dense_gemm')
    return D


def _rows2d(t):
    """[rows, cols] bf16 view with a row pitch that is a multiple of 8 elements and a 16-byte
    aligned base (copies only when the caller's tensor does not qualify)."""
    assert t.dim() == 2
    if t.stride(1) != 1 or t.stride(0) % 8 or t.data_ptr() % 16 or t.stride(0) < t.size(1):
        u = torch.empty((t.size(0), (t.size(1) + 7) // 8 * 8), dtype=t.dtype, device=t.device)
        u = u[:, :t.size(1)]
        u.copy_(t)
        return u
    return t


def gate_colsum(dy, y=None, want_dz=True, out_dtype=torch.float32):
    """(dz, colsum): dz = dy * [y > 0] (dy itself when y is None) and the column sums of dz."""
    dy = _rows2d(dy)
    rows, N = dy.shape
    dz = None
    if y is not None and want_dz:
        dz = torch.empty((rows, (N + 7) // 8 * 8), dtype=BF16, device=dy.device)[:, :N]
    part = torch.empty(((rows + 63) // 64 or 1, N), dtype=torch.float32, device=dy.device)
    out = torch.empty(N, dtype=out_dtype, device=dy.device)
    check(lib().htd_gate_colsum(_p(dy), dy.stride(0), _p(y), y.stride(0) if y is not None else 0,
                                rows, N, _p(dz), dz.stride(0) if dz is not None else 0, _p(part),
                                _p(out), _lib.dt(out), stream()), 'htd_gate_colsum')
    return (dz if dz is not None else dy), out


class _Linear(torch.autograd.Function):
    """y = act(x w^T + b) on the tcgen05 kernels.  forward: NT with bias / ReLU in the epilogue;
    backward: dz = dy * [y > 0] fused with the bias gradient (one pass), dx = dz w (NN),
    dw = dz^T x (TN)."""

    @staticmethod
    def forward(ctx, x, w, b, relu):
        x2 = _rows2d(x.detach())
        w2 = _rows2d(w.detach())
        M, K = x2.shape
        N = w2.shape[0]
        ldy = (N + 7) // 8 * 8
        y = torch.empty((M, ldy), dtype=BF16, device=x.device)[:, :N]
        bf = b.detach().contiguous() if b is not None else None      # bf16 or fp32: read as it is
        gemm(_lib.DENSE_NT, x2, w2, y, M=M, N=N, K=K, lda=x2.stride(0), ldb=w2.stride(0), ldd=ldy,
             bias=bf, relu=relu, name='fc_fwd')
        ctx.relu = relu
        ctx.has_bias = b is not None
        ctx.bdtype = b.dtype if b is not None else None
        ctx.save_for_backward(x2, w2, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x2, w2, y = ctx.saved_tensors
        M, K = x2.shape
        N = w2.shape[0]
        dy = dy if dy.dtype == BF16 else dy.to(BF16)
        need_db = ctx.has_bias and ctx.needs_input_grad[2]
        db = None
        if ctx.relu or need_db:
            dz, colsum = gate_colsum(dy, y if ctx.relu else None,
                                     out_dtype=ctx.bdtype if ctx.bdtype in _lib._DT else torch.float32)
            if need_db:
                db = colsum if colsum.dtype == ctx.bdtype else colsum.to(ctx.bdtype)
        else:
            dz = _rows2d(dy)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty((M, K), dtype=BF16, device=dy.device)
            gemm(_lib.DENSE_NN, dz, w2, dx, M=M, N=K, K=N, lda=dz.stride(0), ldb=w2.stride(0), ldd=K,
                 name='fc_dgrad')
        if ctx.needs_input_grad[1]:
            dw = torch.empty((N, K), dtype=BF16, device=dy.device)
            gemm(_lib.DENSE_TN, dz, x2, dw, M=N, N=K, K=M, lda=dz.stride(0), ldb=x2.stride(0), ldd=K,
                 name='fc_wgrad')
        return dx, dw, db, None


class _LinearDual(torch.autograd.Function):
    """H [2M, N]: H[:M] = relu(x w^T + b), H[M:] = relu(x w^T + b + corr[cls[m]]) - the two FC inputs
    of HTDBBoxHead (`fcs(x_cls)` and `fcs(x_cls + global_feat)`, htd_bbox_head.py:161-164,191-192)
    from one product: the second output is written by the same epilogue (D2 / row_bias)."""

    @staticmethod
    def forward(ctx, x, w, b, corr, cls):
        x2, w2 = _rows2d(x.detach()), _rows2d(w.detach())
        M, K = x2.shape
        N = w2.shape[0]
        assert N % 8 == 0
        H = torch.empty((2 * M, N), dtype=BF16, device=x.device)
        rb = corr.detach().float().contiguous()
        rc = cls.detach().to(torch.int32).contiguous()
        gemm(_lib.DENSE_NT, x2, w2, H[:M], M=M, N=N, K=K, lda=x2.stride(0), ldb=w2.stride(0), ldd=N,
             bias=b.detach().contiguous(), relu=True, D2=H[M:], row_bias=rb, row_class=rc,
             name='fc_fwd')
        ctx.save_for_backward(x2, w2, H, rc)
        ctx.meta = (b.dtype, corr.dtype, corr.shape[0])
        return H

    @staticmethod
    def backward(ctx, dH):
        x2, w2, H, rc = ctx.saved_tensors
        bdt, cdt, R = ctx.meta
        M, K = x2.shape
        N = w2.shape[0]
        dH = dH if dH.dtype == BF16 else dH.to(BF16)
        dH = dH.contiguous()
        # one pass: gradient of the shared pre-activation, bias gradient, gradient of corr per class
        dz = torch.empty((M, N), dtype=BF16, device=dH.device)
        part = torch.empty(((M + 63) // 64 or 1, 1 + R, N), dtype=torch.float32, device=dH.device)
        sums = torch.empty((1 + R, N), dtype=torch.float32, device=dH.device)
        check(lib().htd_dual_gate(_p(dH), _p(H), _p(rc), M, N, R, _p(dz), _p(part), _p(sums),
                                  _lib.dt(sums), stream()), 'htd_dual_gate')
        db, dcorr = sums[0], sums[1:]
        dx = torch.empty((M, K), dtype=BF16, device=dH.device)
        gemm(_lib.DENSE_NN, dz, w2, dx, M=M, N=K, K=N, lda=N, ldb=w2.stride(0), ldd=K, name='fc_dgrad')
        dw = torch.empty((N, K), dtype=BF16, device=dH.device)
        gemm(_lib.DENSE_TN, dz, x2, dw, M=N, N=K, K=M, lda=N, ldb=x2.stride(0), ldd=K, name='fc_wgrad')
        return dx, dw, db.to(bdt), dcorr.to(cdt), None


def linear_dual(x, w, b, corr, cls):
    return _LinearDual.apply(x, w, b, corr, cls)


class _MatMul(torch.autograd.Function):
    """c = a @ b for [M,K] x [K,N] bf16 (b read MN-major)."""

    @staticmethod
    def forward(ctx, a, b):
        a2, b2 = _rows2d(a.detach()), _rows2d(b.detach())
        M, K = a2.shape
        N = b2.shape[1]
        ldc = (N + 7) // 8 * 8
        c = torch.empty((M, ldc), dtype=BF16, device=a.device)[:, :N]
        gemm(_lib.DENSE_NN, a2, b2, c, M=M, N=N, K=K, lda=a2.stride(0), ldb=b2.stride(0), ldd=ldc,
             name='fc_small')
        ctx.save_for_backward(a2, b2)
        return c

    @staticmethod
    def backward(ctx, dc):
        a2, b2 = ctx.saved_tensors
        M, K = a2.shape
        N = b2.shape[1]
        dc = _rows2d(dc if dc.dtype == BF16 else dc.to(BF16))
        da = db = None
        if ctx.needs_input_grad[0]:          # da = dc b^T : [M,N] x [K,N]^T
            lda_ = (K + 7) // 8 * 8
            da = torch.empty((M, lda_), dtype=BF16, device=dc.device)[:, :K]
            gemm(_lib.DENSE_NT, dc, b2, da, M=M, N=K, K=N, lda=dc.stride(0), ldb=b2.stride(0), ldd=lda_,
                 name='fc_small')
        if ctx.needs_input_grad[1]:          # db = a^T dc
            ldb_ = (N + 7) // 8 * 8
            db = torch.empty((K, ldb_), dtype=BF16, device=dc.device)[:, :N]
            gemm(_lib.DENSE_TN, a2, dc, db, M=K, N=N, K=M, lda=a2.stride(0), ldb=dc.stride(0), ldd=ldb_,
                 name='fc_small')
        return da, db


def mm(a, b):
    return _MatMul.apply(a, b)


class _Add3(torch.autograd.Function):
    """a + alpha * b + g[image of the RoI] on channels-last RoI maps in one pass (the regression
    branch input of HTDBBoxHead, htd_bbox_head.py:163,184); backward: the same gradient for a, alpha
    times it for b, and its per-image sum (htd_bias_grad) for g."""

    @staticmethod
    def forward(ctx, a, b, g, rois, alpha):
        P, C, S, S2 = a.shape
        ac = a.detach() if a.is_contiguous(memory_format=torch.channels_last) else \
            a.detach().contiguous(memory_format=torch.channels_last)
        bc = b.detach() if b.is_contiguous(memory_format=torch.channels_last) else \
            b.detach().contiguous(memory_format=torch.channels_last)
        out = torch.empty((P, S, S2, C), dtype=BF16, device=a.device)
        gc = g.detach().reshape(g.shape[0], C).contiguous() if g is not None else None
        r = rois.detach().float().contiguous()
        check(lib().htd_add3(_p(ac), _p(bc), float(alpha), _p(gc), _p(r), P, S * S2, C,
                             gc.shape[0] if gc is not None else 0, _p(out), stream()), 'htd_add3')
        ctx.alpha = float(alpha)
        ctx.gshape = None if g is None else tuple(g.shape)
        ctx.save_for_backward(r)
        return out.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dy):
        from . import ops
        r, = ctx.saved_tensors
        if not dy.is_contiguous(memory_format=torch.channels_last):
            dy = dy.contiguous(memory_format=torch.channels_last)
        dg = None
        if ctx.gshape is not None and ctx.needs_input_grad[2]:
            dg = ops._bias_grad(dy.permute(0, 2, 3, 1), r, ctx.gshape[0]).reshape(ctx.gshape).to(dy.dtype)
        db = dy if ctx.alpha == 1.0 else dy * ctx.alpha
        return dy, db, dg, None, None


def add3(a, b, g, rois, alpha=1.0):
    return _Add3.apply(a, b, g, rois, alpha)


def linear(x, w, b=None, relu=False):
    """``act(F.linear(x, w, b))`` for bf16 CUDA tensors on the own kernels ([M,K] x [N,K])."""
    return _Linear.apply(x, w, b, bool(relu))


class _Conv3x3(torch.autograd.Function):
    """3x3 / stride 1 / padding 1 convolution without bias on channels-last 7x7 RoI maps as an
    implicit GEMM whose halo is the TMA unit's out-of-bounds zero fill."""

    @staticmethod
    def forward(ctx, x, w):
        P, Cin, S, _ = x.shape
        Cout = w.shape[0]
        assert S == 7 and x.shape[3] == 7 and tuple(w.shape[1:]) == (Cin, 3, 3)
        xc = x.detach()
        if not xc.is_contiguous(memory_format=torch.channels_last):
            xc = xc.contiguous(memory_format=torch.channels_last)
        wc = w.detach()
        if not wc.is_contiguous(memory_format=torch.channels_last):
            wc = wc.contiguous(memory_format=torch.channels_last)
        y = torch.empty((P, 7, 7, Cout), dtype=BF16, device=x.device)
        if P:
            gemm(_lib.DENSE_CONV_FPROP, wc, xc, y, P=P, Cin=Cin, Cout=Cout, ldd=Cout, name='conv_fprop')
        ctx.save_for_backward(xc, wc)
        return y.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dy):
        xc, wc = ctx.saved_tensors
        P, Cin = xc.shape[0], xc.shape[1]
        Cout = wc.shape[0]
        dy = dy if dy.dtype == BF16 else dy.to(BF16)
        if not dy.is_contiguous(memory_format=torch.channels_last):
            dy = dy.contiguous(memory_format=torch.channels_last)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty((P, 7, 7, Cin), dtype=BF16, device=dy.device)
            if P:
                gemm(_lib.DENSE_CONV_DGRAD, wc, dy, dx, P=P, Cin=Cin, Cout=Cout, ldd=Cin,
                     name='conv_dgrad')
            dx = dx.permute(0, 3, 1, 2)
        if ctx.needs_input_grad[1]:
            dw = torch.zeros((Cout, 3, 3, Cin), dtype=BF16, device=dy.device) if not P else \
                torch.empty((Cout, 3, 3, Cin), dtype=BF16, device=dy.device)
            if P:
                gemm(_lib.DENSE_CONV_WGRAD, dy, xc, dw, P=P, Cin=Cin, Cout=Cout, ldd=9 * Cin,
                     name='conv_wgrad')
            dw = dw.permute(0, 3, 1, 2)
        return dx, dw


def conv3x3(x, w):
    """``F.conv2d(x, w, padding=1)`` for bf16 channels-last [P,Cin,7,7] RoI maps."""
    return _Conv3x3.apply(x, w)
