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
    """Sorted-space layout + local graph of one batch of RoIs (no gradients involved).

    Everything is sized from static bounds (K RoIs, ``max_group`` RoIs per (image, level) group at
    most) and scheduled on the device, so building and using a plan involves NO host read and
    the whole step can be captured in a CUDA graph; ``groups`` / ``level_blocks`` / ``Npad``
    (inspection, tests) read the table lazily."""

    def __init__(self, rois, levels, num_imgs, num_levels, op_dtype, max_group=None, d=1024,
                 ds=1025):
        _lib.require_cuda(rois, levels)
        dev = rois.device
        rois = rois.detach().float().contiguous()
        levels = levels.to(torch.int32).contiguous()
        K, B, L = rois.shape[0], int(num_imgs), int(num_levels)
        if L * B > _lib.MAX_GROUPS:
            raise ValueError(f'PGraph supports at most {_lib.MAX_GROUPS} (image, level) groups per '
                             f'call, got {L}x{B}')
        self.K, self.B, self.L, self.op_dtype, self.d, self.ds = K, B, L, op_dtype, d, ds
        self.max_group = max(1, min(K, int(max_group) if max_group else K))
        ncap = _round_up(K + L * ALIGN + L * B * GROUP_ALIGN, ALIGN)
        self.Ncap = ncap
        self.ldn = max(_round_up(self.max_group, ALIGN), ALIGN)
        self.ldb = self.ldn // 32
        self.perm = torch.empty(ncap, dtype=torch.int32, device=dev)
        self.pos = torch.empty(K, dtype=torch.int32, device=dev)
        self.rowspan = torch.empty((ncap, 2), dtype=torch.int32, device=dev)
        self.boxes = torch.empty((ncap, 4), dtype=torch.float32, device=dev)
        nt = 2 * L * B + 2 * L + 1
        self.table_dev = torch.empty(nt, dtype=torch.int32, device=dev)
        check(lib().htd_pgraph_plan(ptr(rois), ptr(levels), K, B, L, ALIGN, ncap, ptr(self.perm),
                                    ptr(self.pos), ptr(self.rowspan), ptr(self.boxes),
                                    ptr(self.table_dev), stream()), 'htd_pgraph_plan')
        self.level_seg = self.table_dev[2 * L * B:2 * L * B + 2 * L]
        self.sched = torch.empty(_lib.SCHED_BYTES, dtype=torch.uint8, device=dev)
        check(lib().htd_pgraph_schedule(ptr(self.table_dev), B, L, d, ds, dt(op_dtype),
                                        ptr(self.sched), stream()), 'htd_pgraph_schedule')
        self.max_tiles = [int(lib().htd_pgraph_max_tiles(s, K, B, L, self.max_group, ncap, d, ds,
                                                         dt(op_dtype)))
                          for s in range(_lib.SCHED_SETS)]
        self.bits = torch.empty((ncap, self.ldb), dtype=torch.int32, device=dev)
        self.deg = torch.empty(ncap, dtype=torch.int32, device=dev)
        self.adj = torch.empty((ncap, self.ldn), dtype=op_dtype, device=dev)
        check(lib().htd_iou_graph_build(ptr(self.boxes), ptr(self.rowspan), ncap, ptr(self.bits),
                                        self.ldb, ptr(self.deg), ptr(self.adj), dt(op_dtype),
                                        self.ldn, stream()), 'htd_iou_graph_build')
        self._tab = None

    # ---- inspection helpers: these DO read the table back (tests / accounting only) ------------
    def _table(self):
        if self._tab is None:
            self._tab = self.table_dev.cpu().tolist()
            if max([self._tab[2 * g + 1] for g in range(self.L * self.B)], default=0) > self.max_group:
                raise RuntimeError('a PGraph group is larger than max_group - pass the real bound')
        return self._tab

    @property
    def groups(self):
        """(level, image, off, n) of every non-empty group."""
        t, B = self._table(), self.B
        return [(g // B, g % B, t[2 * g], t[2 * g + 1]) for g in range(self.L * B) if t[2 * g + 1] > 0]

    @property
    def level_blocks(self):
        t, G = self._table(), self.L * self.B
        return [(l, t[2 * G + 2 * l], t[2 * G + 2 * l + 1]) for l in range(self.L)
                if t[2 * G + 2 * l + 1] > 0]

    @property
    def Npad(self):
        return self._table()[-1]

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

    def gemm(self, A, B, sset, D=None, ldd=0, rowmap=None, DT=None, ldt=0, bias=None, relu=False):
        """One scheduled grouped contraction D = A B^T of descriptor set ``sset``."""
        assert A.dtype == B.dtype and A.dim() == 2 and B.dim() == 2
        with _lib.timed('pgraph_gemm'):
            check(lib().htd_pgraph_gemm_scheduled(
                ptr(A), A.shape[0], A.stride(0), ptr(B), B.shape[0], B.stride(0), dt(A),
                ptr(self.sched), int(sset), self.max_tiles[sset], ptr(D),
                dt(D) if D is not None else 0, int(ldd), ptr(rowmap), ptr(DT),
                dt(DT) if DT is not None else 0, int(ldt), ptr(bias), int(bool(relu)), stream()),
                'htd_pgraph_gemm_scheduled')


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
        assert (d, ds) == (plan.d, plan.ds), 'plan was scheduled for other feature widths'
        Np, ldn = plan.Ncap, plan.ldn
        L = plan.L
        need_grad = any(ctx.needs_input_grad[:4])
        refined = torch.zeros((K, d), dtype=x.dtype, device=dev)
        x = x.detach().contiguous()
        sam = sam.detach()
        if sam.stride(1) != 1:
            sam = sam.contiguous()          # the pack kernel takes any row pitch
        Wc = W.detach().to(op).contiguous().reshape(L * d, d)
        bc = b.detach().float().contiguous().reshape(L * d)
        lds = _round_up(ds, ALIGN)
        # ---- sorted-space operands
        _, XT = _pack(x, plan.perm, Np, want_rows=False, want_t=True, out_dtype=op)
        sam_s, samT = _pack(sam, plan.perm, Np, ldd=lds, want_rows=True, want_t=need_grad,
                            out_dtype=op)
        # zero-initialised where a buffer is later read as an operand with pad rows / K tails
        z = torch.zeros((3 if need_grad else 2, Np * d), dtype=op, device=dev)
        Xm, XmT = z[0].view(Np, d), z[1].view(d, Np)
        ZT = z[2].view(d, Np) if need_grad else None
        # ---- Xm = A_local X
        plan.gemm(plan.adj, XT, _lib.SCHED_GROUP_ND, D=Xm, ldd=d, DT=XmT, ldt=Np)
        # ---- S = sam sam^T, A_g = softmax((1 - M) * S)
        S = torch.empty((Np, ldn), dtype=torch.float32, device=dev)
        plan.gemm(sam_s, sam_s, _lib.SCHED_GROUP_NN_S, D=S, ldd=ldn)
        Ag = torch.empty((Np, ldn), dtype=op, device=dev)
        check(lib().htd_pgraph_masked_softmax(ptr(S), ldn, ptr(plan.bits), plan.ldb,
                                              ptr(plan.rowspan), Np, ptr(Ag), dt(op), ldn,
                                              stream()), 'htd_pgraph_masked_softmax')
        # ---- Z = A_g Xm
        Z = torch.zeros((Np, d), dtype=op, device=dev)
        plan.gemm(Ag, XmT, _lib.SCHED_GROUP_ND, D=Z, ldd=d, DT=ZT, ldt=Np)
        # ---- refined = relu(Z W_l^T + b_l), one problem per level block, scattered to RoI order
        plan.gemm(Z, Wc, _lib.SCHED_LEVEL_ND, D=refined, ldd=d, rowmap=plan.perm, bias=bc, relu=True)
        if need_grad:
            ctx.plan = plan
            ctx.save_for_backward(refined, Xm, Ag, ZT, sam_s, samT, Wc)
            ctx.meta = (K, d, ds, x.dtype, sam.dtype, W.dtype, b.dtype, W.shape, b.shape)
        return refined

    @staticmethod
    def backward(ctx, dY):
        plan = ctx.plan
        refined, Xm, Ag, ZT, sam_s, samT, Wc = ctx.saved_tensors
        K, d, ds, xdt, sdt, wdt, bdt, wshape, bshape = ctx.meta
        op = plan.op_dtype
        dev = dY.device
        Np, ldn, L = plan.Ncap, plan.ldn, plan.L
        dY = dY.contiguous()
        if dY.dtype != refined.dtype:
            dY = dY.to(refined.dtype)
        # dU = dY * [Y > 0] in sorted space (+ transposed copy for the weight gradient)
        dU, dUT = _pack(dY, plan.perm, Np, want_rows=True, want_t=True, gate=refined, out_dtype=op)
        db = torch.empty((L, d), dtype=torch.float32, device=dev)
        check(lib().htd_pgraph_segment_colsum(ptr(dU), dt(dU), d, ptr(plan.level_seg), L, d,
                                              ptr(db), stream()), 'htd_pgraph_segment_colsum')
        # dW_l = dU_l^T Z_l   (K = RoIs of the level block; both operands zero in the pad rows)
        # weight gradient straight in the parameter dtype (no fp32 buffer + cast of 4 x d x d)
        dW = torch.zeros((L * d, d), dtype=wdt if wdt in _lib._DT else torch.float32, device=dev)
        plan.gemm(dUT, ZT, _lib.SCHED_LEVEL_DD, D=dW, ldd=d)
        # dZ = dU W_l
        WT = torch.empty((L * d, d), dtype=op, device=dev)
        check(lib().htd_layout_convert(ptr(Wc), dt(Wc), ptr(WT), dt(WT), L, d, d, stream()),
              'htd_layout_convert(W^T)')
        z = torch.zeros((3, Np * d), dtype=op, device=dev)
        dZ, dZT, dXmT = z[0].view(Np, d), z[1].view(d, Np), z[2].view(d, Np)
        plan.gemm(dU, WT, _lib.SCHED_LEVEL_ND, D=dZ, ldd=d, DT=dZT, ldt=Np)
        # dXm = A_g^T dZ  (only its transpose is needed, as the operand of dX)
        AgT = torch.empty((Np, ldn), dtype=op, device=dev)
        check(lib().htd_pgraph_group_transpose(ptr(Ag), dt(Ag), ldn, ptr(plan.rowspan), Np, 0.0, 1.0,
                                               ptr(AgT), dt(AgT), ldn, stream()),
              'htd_pgraph_group_transpose(A_g)')
        plan.gemm(AgT, dZT, _lib.SCHED_GROUP_ND, DT=dXmT, ldt=Np)
        # dA_g = dZ Xm^T ; softmax backward ; dS symmetrised
        dAg = torch.empty((Np, ldn), dtype=torch.float32, device=dev)
        plan.gemm(dZ, Xm, _lib.SCHED_GROUP_NN_D, D=dAg, ldd=ldn)
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
        plan.gemm(Ssym, samT, _lib.SCHED_GROUP_NS, D=dsam, ldd=ds, rowmap=plan.perm)
        dx = torch.zeros((K, d), dtype=xdt, device=dev)
        plan.gemm(plan.adj, dXmT, _lib.SCHED_GROUP_ND, D=dx, ldd=d, rowmap=plan.perm)
        return dx, dsam, dW.reshape(wshape).to(wdt), db.reshape(bshape).to(bdt), None


def pgraph_refine(x_cls, sam, weights, biases, plan):
    """x_cls [K,d], sam [K,ds], per-level Linear weights/biases (lists of L tensors)."""
    W = torch.stack(list(weights), 0)
    b = torch.stack(list(biases), 0)
    return _PGraphFunction.apply(x_cls, sam, W, b, plan)
