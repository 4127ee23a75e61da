"""Progressive Graph (PGraph) aggregation of the HTD head on the C-ABI kernels.

Replaces the per-(image, level) Python loop of ``HTDBBoxHead.forward``
(``mmdet/models/roi_heads/bbox_heads/htd_bbox_head.py:195-219``): for every group the reference
builds the IoU mask, both adjacency matrices and runs three small matmuls plus a Linear, with a
host sync per group.  Here all groups of the batch share every launch:

    plan (stable sort by level, image) -> IoU mask bits / degrees / A_local      [csrc/pgraph.cu]
    Xm = A_local X          S = sam sam^T        A_g = softmax((1-M) * S)
    Z  = A_g Xm             refined = relu(Z W_lvl^T + b_lvl)   (scattered back to RoI order)

The contractions are ``htd_pgraph_gemm`` launches (csrc/pgraph_gemm.cu): bf16 operands run on
tcgen05 tensor cores (TMA-fed, TMEM accumulators), fp32 operands on exact-fp32 FFMA tiles.
Backward follows SURVEY.md Appendix D.  The only host read is the (L*B+L)*2+1 int plan table.
"""
import ctypes

import torch

from . import _lib
from ._lib import HtdGemmGroup, check, dt, lib, ptr, stream

ALIGN = 64          # level blocks start on multiples of the GEMM K tile
GROUP_ALIGN = 8     # groups start on 16-byte (8 x bf16) boundaries: TMA coordinate alignment


def _round_up(a, b):
    return (a + b - 1) // b * b


class GraphPlan:
    """Sorted-space layout + local graph of one batch of RoIs (no gradients involved)."""

    def __init__(self, rois, levels, num_imgs, num_levels, op_dtype):
        _lib.require_cuda(rois, levels)
        dev = rois.device
        rois = rois.detach().float().contiguous()
        levels = levels.to(torch.int32).contiguous()
        K, B, L = rois.shape[0], int(num_imgs), int(num_levels)
        if L * B > _lib.MAX_GROUPS:
            raise ValueError(f'PGraph supports at most {_lib.MAX_GROUPS} (image, level) groups per '
                             f'call, got {L}x{B}')
        ncap = K + L * ALIGN + L * B * GROUP_ALIGN
        self.K, self.B, self.L, self.op_dtype = K, B, L, op_dtype
        self.perm = torch.empty(ncap, dtype=torch.int32, device=dev)
        self.pos = torch.empty(K, dtype=torch.int32, device=dev)
        self.rowspan = torch.empty((ncap, 2), dtype=torch.int32, device=dev)
        self.boxes = torch.empty((ncap, 4), dtype=torch.float32, device=dev)
        nt = 2 * L * B + 2 * L + 1
        self.table_dev = torch.empty(nt, dtype=torch.int32, device=dev)
        check(lib().htd_pgraph_plan(ptr(rois), ptr(levels), K, B, L, ALIGN, ncap, ptr(self.perm),
                                    ptr(self.pos), ptr(self.rowspan), ptr(self.boxes),
                                    ptr(self.table_dev), stream()), 'htd_pgraph_plan')
        tab = self.table_dev.cpu().tolist()          # the one host read of the PGraph path
        self.groups = [(g // B, g % B, tab[2 * g], tab[2 * g + 1]) for g in range(L * B)
                       if tab[2 * g + 1] > 0]        # (level, image, off, n)
        self.level_blocks = [(l, tab[2 * L * B + 2 * l], tab[2 * L * B + 2 * l + 1])
                             for l in range(L) if tab[2 * L * B + 2 * l + 1] > 0]
        self.level_seg = self.table_dev[2 * L * B:2 * L * B + 2 * L]
        self.Npad = tab[-1]
        nmax = max([g[3] for g in self.groups], default=0)
        self.ldn = max(_round_up(nmax, ALIGN), ALIGN)
        self.ldb = self.ldn // 32
        Np = self.Npad
        self.bits = torch.empty((max(Np, 1), self.ldb), dtype=torch.int32, device=dev)
        self.deg = torch.empty(max(Np, 1), dtype=torch.int32, device=dev)
        self.adj = torch.empty((max(Np, 1), self.ldn), dtype=op_dtype, device=dev)
        check(lib().htd_iou_graph_build(ptr(self.boxes), ptr(self.rowspan), Np, ptr(self.bits),
                                        self.ldb, ptr(self.deg), ptr(self.adj), dt(op_dtype),
                                        self.ldn, stream()), 'htd_iou_graph_build')

    # ---- inspection helpers (tests / parity of the "neighbour indices") ----------------------
    def group_mask(self, level, image):
        """(original RoI indices, dense 0/1 mask [n,n], degrees) of one group."""
        for l, b, off, n in self.groups:
            if l == level and b == image:
                idx = self.perm[off:off + n].long()
                words = self.bits[off:off + n].to(torch.int64) & 0xFFFFFFFF
                j = torch.arange(n, device=words.device)
                m = (words[:, j // 32] >> (j % 32)) & 1
                return idx, m.to(torch.float32), self.deg[off:off + n]
        return None

    def flops(self, d=1024, ds=1025):
        """Algorithmic forward flops: 4 n^2 d + 2 n^2 ds + 2 n d^2 per group."""
        return sum(4 * n * n * d + 2 * n * n * ds + 2 * n * d * d for _, _, _, n in self.groups)


def _groups_array(items):
    arr = (HtdGemmGroup * max(len(items), 1))()
    for i, kw in enumerate(items):
        g = arr[i]
        for f, _ in HtdGemmGroup._fields_:
            setattr(g, f, int(kw.get(f, 0)))
    return arr


def _gemm(A, B, groups, D=None, ldd=0, rowmap=None, DT=None, ldt=0, bias=None, relu=False):
    if not groups:
        return
    assert A.dtype == B.dtype and A.dim() == 2 and B.dim() == 2
    arr = _groups_array(groups)
    check(lib().htd_pgraph_gemm(
        ptr(A), A.shape[0], A.stride(0), ptr(B), B.shape[0], B.stride(0), dt(A), arr, len(groups),
        ptr(D), dt(D) if D is not None else 0, int(ldd), ptr(rowmap),
        ptr(DT), dt(DT) if DT is not None else 0, int(ldt), ptr(bias), int(bool(relu)), stream()),
        'htd_pgraph_gemm')


def _pack(src, perm, Npad, ldd=None, want_rows=True, want_t=False, gate=None, out_dtype=None):
    D = src.shape[1]
    out_dtype = out_dtype or src.dtype
    dev = src.device
    dst = torch.empty((Npad, ldd or D), dtype=out_dtype, device=dev) if want_rows else None
    dstT = torch.empty((D, Npad), dtype=out_dtype, device=dev) if want_t else None
    check(lib().htd_pgraph_pack(ptr(src), dt(src), src.stride(0), ptr(gate),
                                dt(gate) if gate is not None else 0,
                                gate.stride(0) if gate is not None else 0, ptr(perm), Npad, D,
                                ptr(dst), dst.stride(0) if dst is not None else 0, ptr(dstT),
                                dstT.stride(0) if dstT is not None else 0, dt(out_dtype), stream()),
          'htd_pgraph_pack')
    return dst, dstT


class _PGraphFunction(torch.autograd.Function):
    """refined[k] = relu(graph_lvl(A_global (A_local X)))[k]  for every RoI k (0 for RoIs in no
    group).  x [K,d], sam [K,ds], W [L,d,d] (stacked graph_lvl{i}_cls.weight), b [L,d]."""

    @staticmethod
    def forward(ctx, x, sam, W, b, plan):
        _lib.require_cuda(x, sam, W, b)
        op = plan.op_dtype
        dev = x.device
        K, d = x.shape
        ds = sam.shape[1]
        Np, ldn = plan.Npad, plan.ldn
        L = plan.L
        need_grad = any(ctx.needs_input_grad[:4])
        refined = torch.zeros((K, d), dtype=x.dtype, device=dev)
        if Np == 0:
            ctx.empty = True
            ctx.shapes = (x.shape, sam.shape, W.shape, b.shape, x.dtype, sam.dtype, W.dtype, b.dtype)
            return refined
        ctx.empty = False
        x = x.detach().contiguous()
        sam = sam.detach().contiguous()
        Wc = W.detach().to(op).contiguous().reshape(L * d, d)
        bc = b.detach().float().contiguous().reshape(L * d)
        lds = _round_up(ds, ALIGN)
        # ---- sorted-space operands
        _, XT = _pack(x, plan.perm, Np, want_rows=False, want_t=True, out_dtype=op)
        sam_s, samT = _pack(sam, plan.perm, Np, ldd=lds, want_rows=True, want_t=need_grad,
                            out_dtype=op)
        G = plan.groups
        # ---- Xm = A_local X
        Xm = torch.zeros((Np, d), dtype=op, device=dev)
        XmT = torch.zeros((d, Np), dtype=op, device=dev)
        _gemm(plan.adj, XT, [dict(M=n, N=d, K=n, a_row=off, b_k0=off, d_row=off, dt_col=off)
                             for _, _, off, n in G], D=Xm, ldd=d, DT=XmT, ldt=Np)
        # ---- S = sam sam^T, A_g = softmax((1 - M) * S)
        S = torch.empty((Np, ldn), dtype=torch.float32, device=dev)
        _gemm(sam_s, sam_s, [dict(M=n, N=n, K=ds, a_row=off, b_row=off, d_row=off)
                             for _, _, off, n in G], D=S, ldd=ldn)
        Ag = torch.empty((Np, ldn), dtype=op, device=dev)
        check(lib().htd_pgraph_masked_softmax(ptr(S), ldn, ptr(plan.bits), plan.ldb,
                                              ptr(plan.rowspan), Np, ptr(Ag), dt(op), ldn,
                                              stream()), 'htd_pgraph_masked_softmax')
        # ---- Z = A_g Xm
        Z = torch.zeros((Np, d), dtype=op, device=dev)
        ZT = torch.zeros((d, Np), dtype=op, device=dev) if need_grad else None
        _gemm(Ag, XmT, [dict(M=n, N=d, K=n, a_row=off, b_k0=off, d_row=off, dt_col=off)
                        for _, _, off, n in G], D=Z, ldd=d, DT=ZT, ldt=Np)
        # ---- refined = relu(Z W_l^T + b_l), one problem per level block, scattered to RoI order
        _gemm(Z, Wc, [dict(M=n, N=d, K=d, a_row=off, b_row=l * d, d_row=off, bias_off=l * d)
                      for l, off, n in plan.level_blocks], D=refined, ldd=d, rowmap=plan.perm,
              bias=bc, relu=True)
        if need_grad:
            ctx.plan = plan
            ctx.save_for_backward(refined, Xm, Ag, ZT, sam_s, samT, Wc)
            ctx.meta = (K, d, ds, x.dtype, sam.dtype, W.dtype, b.dtype, W.shape, b.shape)
        return refined

    @staticmethod
    def backward(ctx, dY):
        if ctx.empty:
            xs, ss, ws, bs, xd, sd, wd, bd = ctx.shapes
            dev = dY.device
            return (torch.zeros(xs, dtype=xd, device=dev), torch.zeros(ss, dtype=sd, device=dev),
                    torch.zeros(ws, dtype=wd, device=dev), torch.zeros(bs, dtype=bd, device=dev),
                    None)
        plan = ctx.plan
        refined, Xm, Ag, ZT, sam_s, samT, Wc = ctx.saved_tensors
        K, d, ds, xdt, sdt, wdt, bdt, wshape, bshape = ctx.meta
        op = plan.op_dtype
        dev = dY.device
        Np, ldn, L = plan.Npad, plan.ldn, plan.L
        G = plan.groups
        dY = dY.contiguous()
        if dY.dtype != refined.dtype:
            dY = dY.to(refined.dtype)
        # dU = dY * [Y > 0] in sorted space (+ transposed copy for the weight gradient)
        dU, dUT = _pack(dY, plan.perm, Np, want_rows=True, want_t=True, gate=refined, out_dtype=op)
        db = torch.empty((L, d), dtype=torch.float32, device=dev)
        check(lib().htd_pgraph_segment_colsum(ptr(dU), dt(dU), d, ptr(plan.level_seg), L, d,
                                              ptr(db), stream()), 'htd_pgraph_segment_colsum')
        # dW_l = dU_l^T Z_l   (K = RoIs of the level block; both operands zero in the pad rows)
        dW = torch.zeros((L * d, d), dtype=torch.float32, device=dev)
        _gemm(dUT, ZT, [dict(M=d, N=d, K=n, a_k0=off, b_k0=off, d_row=l * d)
                        for l, off, n in plan.level_blocks], D=dW, ldd=d)
        # dZ = dU W_l
        WT = torch.empty((L * d, d), dtype=op, device=dev)
        check(lib().htd_layout_convert(ptr(Wc), dt(Wc), ptr(WT), dt(WT), L, d, d, stream()),
              'htd_layout_convert(W^T)')
        dZ = torch.zeros((Np, d), dtype=op, device=dev)
        dZT = torch.zeros((d, Np), dtype=op, device=dev)
        _gemm(dU, WT, [dict(M=n, N=d, K=d, a_row=off, b_row=l * d, d_row=off, dt_col=off)
                       for l, off, n in plan.level_blocks], D=dZ, ldd=d, DT=dZT, ldt=Np)
        # dXm = A_g^T dZ  (only its transpose is needed, as the operand of dX)
        AgT = torch.empty((Np, ldn), dtype=op, device=dev)
        check(lib().htd_pgraph_group_transpose(ptr(Ag), dt(Ag), ldn, ptr(plan.rowspan), Np, 0.0, 1.0,
                                               ptr(AgT), dt(AgT), ldn, stream()),
              'htd_pgraph_group_transpose(A_g)')
        dXmT = torch.zeros((d, Np), dtype=op, device=dev)
        _gemm(AgT, dZT, [dict(M=n, N=d, K=n, a_row=off, b_k0=off, dt_col=off)
                         for _, _, off, n in G], DT=dXmT, ldt=Np)
        # dA_g = dZ Xm^T ; softmax backward ; dS symmetrised
        dAg = torch.empty((Np, ldn), dtype=torch.float32, device=dev)
        _gemm(dZ, Xm, [dict(M=n, N=n, K=d, a_row=off, b_row=off, d_row=off)
                       for _, _, off, n in G], D=dAg, ldd=ldn)
        dS = torch.empty((Np, ldn), dtype=torch.float32, device=dev)
        check(lib().htd_pgraph_softmax_bwd(ptr(Ag), dt(Ag), ldn, ptr(dAg), ldn, ptr(plan.bits),
                                           plan.ldb, ptr(plan.rowspan), Np, ptr(dS), ldn, stream()),
              'htd_pgraph_softmax_bwd')
        Ssym = torch.empty((Np, ldn), dtype=op, device=dev)
        check(lib().htd_pgraph_group_transpose(ptr(dS), dt(dS), ldn, ptr(plan.rowspan), Np, 1.0, 1.0,
                                               ptr(Ssym), dt(Ssym), ldn, stream()),
              'htd_pgraph_group_transpose(dS)')
        # dsam = (dS + dS^T) sam      dX = A_local dXm      (both scattered back to RoI order)
        dsam = torch.zeros((K, ds), dtype=sdt, device=dev)
        _gemm(Ssym, samT, [dict(M=n, N=ds, K=n, a_row=off, b_k0=off, d_row=off)
                           for _, _, off, n in G], D=dsam, ldd=ds, rowmap=plan.perm)
        dx = torch.zeros((K, d), dtype=xdt, device=dev)
        _gemm(plan.adj, dXmT, [dict(M=n, N=d, K=n, a_row=off, b_k0=off, d_row=off)
                               for _, _, off, n in G], D=dx, ldd=d, rowmap=plan.perm)
        return dx, dsam, dW.reshape(wshape).to(wdt), db.reshape(bshape).to(bdt), None


def pgraph_refine(x_cls, sam, weights, biases, plan):
    """x_cls [K,d], sam [K,ds], per-level Linear weights/biases (lists of L tensors)."""
    W = torch.stack(list(weights), 0)
    b = torch.stack(list(biases), 0)
    return _PGraphFunction.apply(x_cls, sam, W, b, plan)
