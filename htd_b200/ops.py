"""torch.autograd bindings of the C-ABI kernels (RoIAlign / BA part).

PyTorch only provides device memory, the current stream and autograd plumbing here; all
arithmetic of the path happens in ``libhtd_b200.so``.  Feature maps are handled as
channels-last: a ``[B,C,H,W]`` tensor whose memory is ``[B,H,W,C]``; RoI features are returned as
``[K,C,P,P]`` tensors with channels_last strides (values index exactly like the reference's
NCHW tensors).
"""
import ctypes

import os

import torch

from . import _lib
from ._lib import check, dt, lib, ptr, stream


# ------------------------------------------------------------------------------------------
# layout
# ------------------------------------------------------------------------------------------
def _is_cl(t):
    return t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last)


def _convert(src, dst, N, R, S):
    check(lib().htd_layout_convert(ptr(src), dt(src), ptr(dst), dt(dst), N, R, S, stream()),
          'htd_layout_convert')


class _ToChannelsLast(torch.autograd.Function):
    """NCHW-contiguous -> channels-last (+ optional dtype change) with the transpose kernel."""

    @staticmethod
    def forward(ctx, x, dtype):
        B, C, H, W = x.shape
        out = torch.empty((B, H, W, C), dtype=dtype, device=x.device)
        _convert(x, out, B, C, H * W)
        ctx.src_dtype = x.dtype
        return out.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, g):
        B, C, H, W = g.shape
        if not _is_cl(g):
            g = g.contiguous(memory_format=torch.channels_last)
        out = torch.empty((B, C, H, W), dtype=ctx.src_dtype, device=g.device)
        _convert(g, out, B, H * W, C)
        return out, None


def to_channels_last(x, dtype=None):
    """Feature map in the layout the kernels read.  No copy if it already is channels-last in
    the requested dtype."""
    _lib.require_cuda(x)
    dtype = dtype or x.dtype
    if _is_cl(x) and x.dtype == dtype:
        return x
    if x.is_contiguous():
        return _ToChannelsLast.apply(x, dtype)
    return x.to(dtype).contiguous(memory_format=torch.channels_last)


class _FlattenRoIFeats(torch.autograd.Function):
    """[K,C,P,P] RoI features with channels-last strides -> [K, C*P*P] in the reference's
    flatten order (c*PP + bin), which the FC weights index (htd_bbox_head.py:191,
    convfc_bbox_head.py:145).  Transpose kernel in both directions instead of ATen's strided copy
    (0.1 ms per 25 MB in the round-1 profile).  With ``bias`` ([B,C,1,1]) and ``rois`` the SFA
    vector of each RoI's image is added on the way (``_fuse_global``, htd_roi_head.py:133-141)."""

    @staticmethod
    def forward(ctx, x, bias, rois):
        K, C, P, Q = x.shape
        out = torch.empty((K, C * P * Q), dtype=x.dtype, device=x.device)
        ctx.bias_shape = None
        if bias is None:
            _convert(x, out, K, P * Q, C)        # memory [K, PQ, C] -> [K, C, PQ]
        else:
            B = bias.shape[0]
            bias_c = bias.detach().reshape(B, C).float().contiguous()
            rois_c = rois.detach().float().contiguous()
            check(lib().htd_roi_flatten(ptr(x), dt(x), ptr(out), dt(out), K, P * Q, C, ptr(bias_c),
                                        ptr(rois_c), B, stream()), 'htd_roi_flatten')
            ctx.bias_shape, ctx.bias_dtype = tuple(bias.shape), bias.dtype
            ctx.save_for_backward(rois_c)
        ctx.shape = (K, C, P, Q)
        return out

    @staticmethod
    def backward(ctx, g):
        K, C, P, Q = ctx.shape
        g = g.contiguous()
        out = torch.empty((K, P, Q, C), dtype=g.dtype, device=g.device)
        _convert(g, out, K, C, P * Q)            # [K, C, PQ] -> [K, PQ, C]
        dbias = None
        if ctx.bias_shape is not None and ctx.needs_input_grad[1]:
            rois_c, = ctx.saved_tensors
            dbias = _bias_grad(out, rois_c, ctx.bias_shape[0]).reshape(ctx.bias_shape).to(ctx.bias_dtype)
        return out.permute(0, 3, 1, 2), dbias, None


class _FlattenWithPrefix(torch.autograd.Function):
    """(flatten_roi_feats(x), cat of the row spans of x) with ONE gradient tensor: the stage-1 head
    reads the same RoI features twice - flattened for the cls branch and, as the positive prefix of
    every image's block, for the reg branch (htd_roi_head.py:157-170).  Plain autograd answers
    with two zero-filled full-size gradients (one per slice) and their additions; here the
    flatten-backward tensor is the gradient and the spans are added into it in place."""

    @staticmethod
    def forward(ctx, x, spans):
        K, C, P, Q = x.shape
        flat = torch.empty((K, C * P * Q), dtype=x.dtype, device=x.device)
        _convert(x, flat, K, P * Q, C)
        pos = torch.cat([x[o:o + n] for o, n in spans], 0) if spans else x[:0]
        ctx.shape, ctx.spans = (K, C, P, Q), tuple(spans)
        return flat, pos

    @staticmethod
    def backward(ctx, g_flat, g_pos):
        K, C, P, Q = ctx.shape
        g_flat = g_flat.contiguous()
        out = torch.empty((K, P, Q, C), dtype=g_flat.dtype, device=g_flat.device)
        _convert(g_flat, out, K, C, P * Q)
        out = out.permute(0, 3, 1, 2)
        off = 0
        for o, n in ctx.spans:
            if n:
                out[o:o + n] += g_pos[off:off + n].to(out.dtype)
            off += n
        return out, None


def flatten_with_prefix(x, spans):
    """x [K,C,P,P] channels-last RoI features, spans = [(first row, rows)] -> (x flattened in the
    reference's c*PP+bin order [K, C*P*P], the span rows [sum rows, C, P, P])."""
    _lib.require_cuda(x)
    if not _is_cl(x):
        x = x.contiguous(memory_format=torch.channels_last)
    return _FlattenWithPrefix.apply(x, spans)


def flatten_fuses_bias(x):
    """True when ``flatten_roi_feats(x, bias, rois)`` adds the bias inside the transpose kernel."""
    return x.is_cuda and x.dim() == 4 and _is_cl(x) and x.dtype in _lib._DT and x.shape[1] > 1 and \
        x.shape[1] * x.shape[2] * x.shape[3] <= 12800 and (x.shape[1] * x.shape[2] * x.shape[3]) % 8 == 0


def flatten_fuses_bias_shape(C, output_size):
    """Same question for a [K, C, P, Q] channels-last CUDA map that does not exist yet."""
    P, Q = (output_size, output_size) if isinstance(output_size, int) else tuple(output_size)
    return C > 1 and C * P * Q <= 12800 and (C * P * Q) % 8 == 0


def flatten_roi_feats(x, bias=None, rois=None):
    """``x.flatten(1)`` for RoI features (+ the per-image ``bias`` [B,C,1,1] of the RoIs' images,
    ``rois`` [K,5]); uses the transpose kernel when ``x`` is a CUDA channels-last tensor, plain
    flatten otherwise."""
    if bias is not None and not flatten_fuses_bias(x):
        x = x + bias[rois[:, 0].long()].to(x.dtype)
        bias = None
    if x.is_cuda and x.dim() == 4 and _is_cl(x) and x.dtype in _lib._DT and x.shape[1] > 1:
        return _FlattenRoIFeats.apply(x, bias, rois)
    return x.flatten(1)


def _bhwc(x):
    """[B,C,H,W] channels-last tensor -> its [B,H,W,C] memory view."""
    return x.permute(0, 2, 3, 1)


# ------------------------------------------------------------------------------------------
# level assignment
# ------------------------------------------------------------------------------------------
def level_assign(rois, num_levels, finest_scale=56.0):
    """int32 level per RoI (-1: NaN scale).  SingleRoIExtractor.map_roi_levels."""
    _lib.require_cuda(rois)
    rois = rois.detach().float().contiguous()
    out = torch.empty(rois.shape[0], dtype=torch.int32, device=rois.device)
    check(lib().htd_level_assign(ptr(rois), rois.shape[0], int(num_levels), float(finest_scale),
                                 ptr(out), stream()), 'htd_level_assign')
    return out


def roi_footprints(feats_cl, scales, rois, roi_level, pooled, sampling_ratio, count=False):
    L, K = len(feats_cl), rois.shape[0]
    B = feats_cl[0].shape[0]
    boxes = torch.empty((L, K, 4), dtype=torch.int32, device=rois.device)
    cnt = torch.zeros(L, dtype=torch.int64, device=rois.device) if count else None
    lv = _lib.make_levels([_bhwc(f) for f in feats_cl], scales)
    check(lib().htd_roi_footprints(lv, L, B, ptr(rois), K, ptr(roi_level), pooled, sampling_ratio,
                                   ptr(boxes), ptr(cnt), stream()), 'htd_roi_footprints')
    return (boxes, cnt) if count else boxes


class RoIPlan:
    """Sampling plan of one extractor call (htd_roi_plan): footprint boxes, table offsets, per-bin
    pixel ranges and the separable axis-weight tables - shared by forward and backward."""

    def __init__(self, feats_cl, scales, rois, roi_level, pooled, sampling_ratio):
        L, K = len(feats_cl), rois.shape[0]
        B = feats_cl[0].shape[0]
        dev = rois.device
        lv = _lib.make_levels([_bhwc(f) for f in feats_cl], scales)
        rows = int(lib().htd_roi_plan_rows_bound(lv, L, K, int(roi_level is not None)))
        self.boxes = torch.empty((L, K, 4), dtype=torch.int32, device=dev)
        self.offsets = torch.empty(L * K + 1, dtype=torch.int32, device=dev)
        self.ranges = torch.empty((L * K, 4 * 8), dtype=torch.int32, device=dev)
        self.weights = torch.empty((rows, 8), dtype=torch.float32, device=dev)
        self.pooled, self.sampling_ratio = pooled, sampling_ratio
        check(lib().htd_roi_plan(lv, L, B, ptr(rois), K, ptr(roi_level), pooled, sampling_ratio,
                                 ptr(self.boxes), ptr(self.offsets), ptr(self.ranges),
                                 ptr(self.weights), rows, None, stream()), 'htd_roi_plan')

    def tensors(self):
        return self.boxes, self.offsets, self.ranges, self.weights

    def pixels(self):
        bx = self.boxes.reshape(-1, 4).long()
        h = (bx[:, 1] - bx[:, 0] + 1).clamp(min=0)
        w = (bx[:, 3] - bx[:, 2] + 1).clamp(min=0)
        return int((h * w).sum().item())


# ------------------------------------------------------------------------------------------
# multi-level RoIAlign
# ------------------------------------------------------------------------------------------
def _roi_align_bwd(feat_shapes, feat_dtype, scales, rois, plan_tensors, pooled, dy,
                   dy_per_level, scale=None, ring_edge=-1, addvec=None):
    """Single-source gather backward; returns per-level dX as [B,C,H,W] channels-last tensors."""
    src = dict(rois=rois, plan=plan_tensors, dy=dy, dy_per_level=dy_per_level, scale=scale,
               ring_edge=ring_edge, addvec=addvec)
    return _bwd_multi(feat_shapes, feat_dtype, False, scales, [src], pooled)


def _bwd_multi(feat_shapes, out_dtype, nchw, scales, sources, pooled):
    """ONE tile-gather launch for several extractor calls (htd_roi_align_bwd_multi).  ``sources``:
    dicts with rois, plan (boxes, offsets, ranges, weights), dy, dy_per_level, scale, ring_edge,
    addvec.  Returns per-level dX: [B,C,H,W] contiguous when ``nchw`` else channels-last views."""
    dev = sources[0]['rois'].device
    L = len(feat_shapes)
    B, C = feat_shapes[0][0], feat_shapes[0][1]
    if nchw:
        bufs = [torch.empty((B, C, s[2], s[3]), dtype=out_dtype, device=dev) for s in feat_shapes]
        lv = _lib.make_levels([b.permute(0, 2, 3, 1) for b in bufs], scales)   # H, W from dims 1, 2
    else:
        bufs = [torch.empty((B, s[2], s[3], C), dtype=out_dtype, device=dev) for s in feat_shapes]
        lv = _lib.make_levels(bufs, scales)
    arr = (_lib.HtdBwdSource * len(sources))()
    name = 'roi_align_bwd(fused)' if len(sources) > 1 else (
        'roi_align_bwd(BA)' if sources[0].get('scale') is not None else 'roi_align_bwd(single)')
    acct = 0
    for i, q in enumerate(sources):
        boxes, offsets, ranges, weights = q['plan']
        a = arr[i]
        a.rois, a.boxes, a.offsets = q['rois'].data_ptr(), boxes.data_ptr(), offsets.data_ptr()
        a.ranges, a.weights, a.dy = ranges.data_ptr(), weights.data_ptr(), q['dy'].data_ptr()
        a.scale = q['scale'].data_ptr() if q.get('scale') is not None else None
        av = q.get('addvec')
        if av is not None and pooled < 8 and \
                lib().htd_roi_align_bwd_uses_tensor_pipe(int(C), int(pooled), dt(q['dy'].dtype)):
            # the tensor-pipe gather takes the add vector as one more bf16 K row of the hit
            av = q['_addvec_bf16'] = av.to(torch.bfloat16)
        a.addvec = av.data_ptr() if av is not None else None
        a.addvec_dtype = dt(av.dtype) if av is not None else 0
        a.K, a.dy_per_level = q['rois'].shape[0], int(bool(q.get('dy_per_level', False)))
        a.ring_edge = int(q.get('ring_edge', -1))
        if _lib.ACCOUNT is not None:  # 49*C*b_dy + fh*fw*C*4 per (RoI, level)
            acct += q['dy'].numel() * q['dy'].element_size() + _count_pixels(boxes) * C * 4
    if _lib.ACCOUNT is not None:
        _lib.ACCOUNT.append((name, acct))
    dyt = sources[0]['dy'].dtype
    assert all(q['dy'].dtype == dyt for q in sources), 'all gradient sources must share a dtype'
    with _lib.timed(name):
        check(lib().htd_roi_align_bwd_multi(lv, L, B, C, dt(out_dtype), int(bool(nchw)), arr,
                                            len(sources), pooled, dt(dyt), stream()),
              'htd_roi_align_bwd_multi')
    return bufs if nchw else [b.permute(0, 3, 1, 2) for b in bufs]


class GradSink:
    """Deferred pyramid gradient: the extractor calls of one step park their backward inputs here
    and the pyramid node gathers all of them in one launch (see ``make_pyramid``)."""

    def __init__(self):
        self.sources = []
        self.scales = None
        self.pooled = None

    def park(self, scales, pooled, source):
        """One gather launch serves every parked source, so they must agree on the spatial scales
        and the pooled size (extractors sharing a Pyramid with different ``output_size`` or
        ``featmap_strides`` would otherwise get a silently wrong dX)."""
        if self.sources and (tuple(self.scales) != tuple(scales) or self.pooled != pooled):
            raise RuntimeError(
                'extractors that share one Pyramid must use the same featmap_strides and '
                f'output_size: parked {self.scales} / {self.pooled}, got {scales} / {pooled}')
        self.scales, self.pooled = scales, pooled
        self.sources.append(source)

    def hand_over(self, stream):
        """The parked tensors travel outside autograd, so its cross-stream bookkeeping does not
        see them: an extractor that ran on a side stream parks tensors from that stream's pool,
        and the gather that reads them runs on the pyramid node's stream.  Tell the allocator,
        or the blocks go back to the side stream's pool (and to its next kernels) as soon as the
        sources are dropped, while the gather is still reading them."""
        def walk(v):
            if torch.is_tensor(v):
                if v.is_cuda:
                    v.record_stream(stream)
            elif isinstance(v, (tuple, list)):
                for u in v:
                    walk(u)
            elif isinstance(v, dict):
                for u in v.values():
                    walk(u)
        walk(self.sources)


_AUX_STREAMS = {}


def _aux_stream(device):
    """One auxiliary stream per device for launches that run next to the current stream's."""
    key = (device.index if device.index is not None else torch.cuda.current_device())
    st = _AUX_STREAMS.get(key)
    if st is None:
        st = _AUX_STREAMS[key] = torch.cuda.Stream(device=device)
    return st


class Pyramid(list):
    """Channels-last feature maps of one step + the token / sink that defer their gradient."""
    token = None
    sink = None


class _PyramidFn(torch.autograd.Function):
    """NCHW maps -> channels-last maps in the compute dtype (one transpose+cast launch per level).
    The outputs are NOT differentiable themselves: extractors that read them hang their autograd
    edge on ``token`` and park (rois, plan, dY ...) in ``sink``; this node's backward then runs
    ONE multi-source tile gather that writes dX directly as [B,C,H,W] in the input dtype - instead
    of one gather + a full dX write per extractor call, autograd's dX additions and a transpose
    back (the reference: 13 RoIAlign backward launches with atomics into zero-filled maps)."""

    @staticmethod
    def forward(ctx, dtype, sink, *xs):
        outs, is_cl = [], []
        # the small levels are converted on a side stream NEXT to the largest one (their launches
        # are latency-bound: one after the other they add ~25 us in front of the first extraction)
        cur = torch.cuda.current_stream(xs[0].device)
        side = _aux_stream(xs[0].device) if len(xs) > 1 else None
        big = max(range(len(xs)), key=lambda i: xs[i].numel())
        if side is not None:
            side.wait_stream(cur)
        for i, x in enumerate(xs):
            B, C, H, W = x.shape
            if _is_cl(x):
                # producer already emits channels-last: no transpose; a cast only
                # if its dtype is not the compute dtype
                outs.append(x.detach() if x.dtype == dtype else x.detach().to(dtype))
                is_cl.append(True)
                continue
            with torch.cuda.stream(cur if (side is None or i == big) else side):
                o = torch.empty((B, H, W, C), dtype=dtype, device=x.device)
                _convert(x, o, B, C, H * W)
            if side is not None and i != big:
                o.record_stream(cur)
            outs.append(o.permute(0, 3, 1, 2))
            is_cl.append(False)
        if side is not None:
            cur.wait_stream(side)
        token = torch.zeros(1, dtype=torch.float32, device=xs[0].device)
        ctx.sink = sink
        ctx.meta = ([tuple(x.shape) for x in xs], [x.dtype for x in xs], is_cl)
        ctx.mark_non_differentiable(*outs)
        # autograd otherwise hands backward a ZERO-FILLED tensor for every non-differentiable output:
        # five fills of the whole bf16 pyramid (92 MB, 38 us) on the critical path right before the
        # backward gather (profiles/r02_launches_step_bf16_final.csv of the previous build)
        ctx.set_materialize_grads(False)
        ctx.device = xs[0].device
        return (token,) + tuple(outs)

    @staticmethod
    def backward(ctx, gtoken, *gouts):
        shapes, xdtypes, is_cl = ctx.meta
        sink = ctx.sink
        if not sink.sources:
            return (None, None) + tuple(torch.zeros(s, dtype=dt_, device=ctx.device)
                                        for s, dt_ in zip(shapes, xdtypes))
        sink.hand_over(torch.cuda.current_stream(ctx.device))
        # channels-last gather (512 B coalesced stores) + one transpose/cast pass back to the
        # reference's NCHW layout; writing NCHW straight from the gather measured 30% slower
        cl = _bwd_multi(shapes, sink.sources[0]['dy'].dtype, False, sink.scales, sink.sources,
                        sink.pooled)
        grads = []
        cur = torch.cuda.current_stream(ctx.device)
        side = _aux_stream(ctx.device) if len(cl) > 1 else None
        big = max(range(len(cl)), key=lambda i: cl[i].numel())
        if side is not None:
            side.wait_stream(cur)
        for i, (g, shp, xdt, keep_cl) in enumerate(zip(cl, shapes, xdtypes, is_cl)):
            B, C, H, W = shp
            if keep_cl:                            # gradient stays channels-last, like the input
                assert g.shape == shp              # [B,C,H,W] view of the [B,H,W,C] gather buffer
                grads.append(g if g.dtype == xdt else g.to(xdt))
                continue
            on_side = side is not None and i != big
            with torch.cuda.stream(side if on_side else cur):
                o = torch.empty(shp, dtype=xdt, device=g.device)
                _convert(g, o, B, H * W, C)
            if on_side:
                g.record_stream(side)              # read there, allocated on `cur`
                o.record_stream(cur)
            grads.append(o)
        if side is not None:
            cur.wait_stream(side)
        sink.sources = []
        return (None, None) + tuple(grads)


def make_pyramid(xs, dtype=None):
    """Feature maps in the layout the kernels read, converted once per step and shared by all
    extractor calls.  Contiguous NCHW or channels-last CUDA inputs get the deferred single-launch
    backward (channels-last inputs are read in place and receive a channels-last gradient: no
    layout pass at all); anything else falls back to per-call ``to_channels_last``."""
    dtype = dtype or xs[0].dtype
    if all(x.is_cuda and x.dim() == 4 and (x.is_contiguous() or _is_cl(x)) for x in xs) and \
            any(x.requires_grad for x in xs) and torch.is_grad_enabled():
        sink = GradSink()
        res = _PyramidFn.apply(dtype, sink, *xs)
        pyr = Pyramid(res[1:])
        pyr.token, pyr.sink = res[0], sink
        return pyr
    return Pyramid(to_channels_last(x, dtype) for x in xs)


def _count_pixels(boxes):
    """Sum of footprint pixels fh*fw over all (level, RoI) of a footprint plan."""
    bx = boxes.reshape(-1, 4).long()
    h = (bx[:, 1] - bx[:, 0] + 1).clamp(min=0)
    w = (bx[:, 3] - bx[:, 2] + 1).clamp(min=0)
    return int((h * w).sum().item())


def _fwd_launch(name, feats, scales, rois, roi_level, pooled, sampling_ratio, bias_c, out,
                plan=None):
    """Builds the sampling plan (unless given) and launches the forward gather; returns the plan."""
    L, K = len(feats), rois.shape[0]
    B, C = feats[0].shape[0], feats[0].shape[1]
    lv = _lib.make_levels([_bhwc(f) for f in feats], scales)
    # the plan (footprints, scan, axis-weight tables) is part of the extraction: it is timed under
    # the same name as the gather it serves
    with _lib.timed(name):
        if plan is None:
            plan = RoIPlan(feats, scales, rois, roi_level, pooled, sampling_ratio)
        check(lib().htd_roi_align_fwd(lv, L, B, C, dt(feats[0]), ptr(rois), K, ptr(roi_level),
                                      pooled, ptr(plan.boxes), ptr(plan.offsets), ptr(plan.ranges),
                                      ptr(plan.weights), ptr(bias_c), ptr(out), dt(out), stream()),
              'htd_roi_align_fwd')
    if _lib.ACCOUNT is not None:      # fh*fw*C*b_in + 49*C*b_out + 20 per (RoI, level)
        tasks = K if roi_level is not None else K * L
        _lib.ACCOUNT.append((name, plan.pixels() * C * feats[0].element_size() +
                             out.numel() * out.element_size() + 20 * tasks))
    return plan


def _bias_grad(g_kppc, rois, B):
    K, PP, C = g_kppc.shape[0], g_kppc.shape[1] * g_kppc.shape[2], g_kppc.shape[3]
    nblk = max((K + 3) // 4, 1)
    partial = torch.empty(max(K, nblk * B, 1) * C, dtype=torch.float32, device=g_kppc.device)
    dbias = torch.empty((B, C), dtype=torch.float32, device=g_kppc.device)
    check(lib().htd_bias_grad(ptr(g_kppc), dt(g_kppc), ptr(rois), K, PP, C, B, ptr(partial),
                              ptr(dbias), stream()), 'htd_bias_grad')
    return dbias


class _RoIAlignLevels(torch.autograd.Function):
    """rois [K,5] + L channels-last maps -> [K,C,P,P] (roi_level given) or [L,K,C,P,P]."""

    @staticmethod
    def forward(ctx, rois, roi_level, bias, scales, pooled, sampling_ratio, out_dtype, token, sink,
                *feats):
        _lib.require_cuda(rois, *feats)
        L, K = len(feats), rois.shape[0]
        B, C = feats[0].shape[0], feats[0].shape[1]
        for f in feats:
            if not _is_cl(f) or f.dtype != feats[0].dtype or f.shape[:2] != feats[0].shape[:2]:
                raise ValueError('feature maps must be channels-last, same dtype, same [B,C]')
        rois = rois.detach().float().contiguous()
        out_dtype = out_dtype or feats[0].dtype
        lead = (K,) if roi_level is not None else (L, K)
        out = torch.empty(lead + (pooled, pooled, C), dtype=out_dtype, device=rois.device)
        bias_shape = None if bias is None else tuple(bias.shape)
        bias_c = None if bias is None else bias.detach().reshape(B, C).float().contiguous()
        plan = _fwd_launch('roi_align_fwd(single)' if roi_level is not None else 'roi_align_fwd(all)',
                           feats, scales, rois, roi_level, pooled, sampling_ratio, bias_c, out)
        ctx.deferred = token is not None and ctx.needs_input_grad[7]
        ctx.sink = sink if ctx.deferred else None
        ctx.with_dx = ctx.deferred or any(ctx.needs_input_grad[9:])
        if ctx.with_dx:                        # the gather backward reuses the plan
            ctx.save_for_backward(rois, roi_level, *plan.tensors())
        else:
            ctx.save_for_backward(rois, roi_level)
        ctx.cfg = (scales, pooled, sampling_ratio, [tuple(f.shape) for f in feats], feats[0].dtype,
                   bias_shape, B)
        return out.permute(0, 3, 1, 2) if roi_level is not None else out.permute(0, 1, 4, 2, 3)

    @staticmethod
    def backward(ctx, g):
        rois, roi_level = ctx.saved_tensors[:2]
        scales, pooled, sr, shapes, fdtype, bias_shape, B = ctx.cfg
        single = roi_level is not None
        g = (g.permute(0, 2, 3, 1) if single else g.permute(0, 1, 3, 4, 2)).contiguous()
        dbias = None
        grads = [None] * len(shapes)
        gtoken = None
        if ctx.deferred:                       # park for the pyramid's single gather launch
            ctx.sink.park(scales, pooled, dict(rois=rois, plan=tuple(ctx.saved_tensors[2:]), dy=g,
                                               dy_per_level=not single))
            gtoken = torch.zeros(1, dtype=torch.float32, device=g.device)
        elif ctx.with_dx:
            grads = _roi_align_bwd(shapes, fdtype, scales, rois, ctx.saved_tensors[2:], pooled, g,
                                   dy_per_level=not single)
        if bias_shape is not None and ctx.needs_input_grad[2]:
            gb = g if single else g.sum(0)
            dbias = _bias_grad(gb, rois, B).reshape(bias_shape)
        return (None, None, dbias, None, None, None, None, gtoken, None) + tuple(grads)


def roi_align_levels(feats_cl, rois, scales, pooled=7, sampling_ratio=0, roi_level=None, bias=None,
                     out_dtype=None):
    out = _RoIAlignLevels.apply(rois, roi_level, bias, tuple(float(s) for s in scales), int(pooled),
                                int(sampling_ratio), out_dtype, getattr(feats_cl, 'token', None),
                                getattr(feats_cl, 'sink', None), *feats_cl)
    return out


# ------------------------------------------------------------------------------------------
# BA: multi-level sample + attention softmax + fuse (+ ring) in one autograd node
# ------------------------------------------------------------------------------------------
class _BAFunction(torch.autograd.Function):
    """AdptRoIExtractor.forward (adaptative_roi_extractor.py:49-91), see csrc/layout_ba.cu.

    Inputs: rois, conv1 weight [128,256,1,1] / bias, conv2 weight [1,128,1,1] / bias, optional
    residual `add` [K,C,P,P] and SFA `bias` [B,C,1,1] (both folded into the fuse pass), then the
    L channels-last maps.  Saves R (the L sampled maps) for the backward dot products.
    """

    @staticmethod
    def forward(ctx, rois, w1, b1, w2, b2, add, bias, scales, pooled, sampling_ratio, edge, token,
                sink, *feats):
        _lib.require_cuda(rois, *feats)
        L, K = len(feats), rois.shape[0]
        B, C = feats[0].shape[0], feats[0].shape[1]
        PP = pooled * pooled
        dev = rois.device
        rois = rois.detach().float().contiguous()
        fdt = feats[0].dtype
        R = torch.empty((L, K, pooled, pooled, C), dtype=fdt, device=dev)
        plan = _fwd_launch('roi_align_fwd(BA)', feats, scales, rois, None, pooled, sampling_ratio,
                           None, R)
        m = torch.empty((L * K, C), dtype=torch.float32, device=dev)
        check(lib().htd_ba_bin_mean(ptr(R), dt(R), L * K, PP, C, ptr(m), stream()),
              'htd_ba_bin_mean')
        # attention MLP on the pooled vectors ([L*K,C] -> 128 -> 1): one fused launch for the
        # configured sizes, plain library GEMMs otherwise
        H1 = w1.shape[0]
        pdt = w1.dtype
        fused_mlp = bool(lib().htd_ba_mlp_supported(C, H1)) and w2.numel() == H1 and \
            all(t.dtype == pdt and t.is_contiguous() for t in (w1, b1, w2, b2)) and \
            pdt in (torch.float32, torch.bfloat16)
        if fused_mlp:
            W1, W2 = w1.detach(), w2.detach()
            h = torch.empty((L * K, H1), dtype=torch.float32, device=dev)
            logits = torch.empty((L, K), dtype=torch.float32, device=dev)
            check(lib().htd_ba_mlp_fwd(ptr(m), L * K, C, H1, ptr(W1), ptr(b1.detach()), ptr(W2),
                                       ptr(b2.detach()), dt(W1), ptr(h), ptr(logits), stream()),
                  'htd_ba_mlp_fwd')
        else:
            W1 = w1.detach().reshape(w1.shape[0], -1).float()
            W2 = w2.detach().reshape(1, -1).float()
            h = torch.tanh(torch.addmm(b1.detach().float(), m, W1.t()))
            logits = torch.addmm(b2.detach().float(), h, W2.t()).reshape(L, K).contiguous()
        wts = torch.empty((L, K), dtype=torch.float32, device=dev)
        out = torch.empty((K, pooled, pooled, C), dtype=fdt, device=dev)
        add_c = None
        if add is not None:
            add_c = add.detach().permute(0, 2, 3, 1).to(fdt).contiguous()
        bias_c = None if bias is None else bias.detach().reshape(B, C).float().contiguous()
        check(lib().htd_ba_fuse_fwd(ptr(R), dt(R), ptr(logits), L, K, pooled, C, int(edge),
                                    ptr(add_c), dt(fdt), ptr(bias_c), ptr(rois), B, ptr(wts),
                                    ptr(out), dt(out), stream()), 'htd_ba_fuse_fwd')
        ctx.deferred = token is not None and ctx.needs_input_grad[11]
        ctx.sink = sink if ctx.deferred else None
        ctx.with_dx = ctx.deferred or any(ctx.needs_input_grad[13:])
        ctx.save_for_backward(rois, R, m, h, wts, W1, W2, *(plan.tensors() if ctx.with_dx else ()))
        ctx.cfg = (scales, pooled, sampling_ratio, edge, [tuple(f.shape) for f in feats], fdt, B,
                   w1.shape, w2.shape, add is not None,
                   None if bias is None else tuple(bias.shape))
        ctx.fused_mlp = fused_mlp
        return out.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, g):
        rois, R, m, h, wts, W1, W2 = ctx.saved_tensors[:7]
        (scales, pooled, sr, edge, shapes, fdt, B, w1_shape, w2_shape, has_add,
         bias_shape) = ctx.cfg
        L, K = wts.shape
        C = R.shape[-1]
        PP = pooled * pooled
        g = g.permute(0, 2, 3, 1).contiguous()
        da = torch.empty((L, K), dtype=torch.float32, device=g.device)
        check(lib().htd_ba_fuse_bwd(ptr(R), dt(R), ptr(g), dt(g), ptr(wts), L, K, PP, C, ptr(da),
                                    stream()), 'htd_ba_fuse_bwd')
        # backward of the tiny attention MLP (Appendix D of SURVEY.md)
        if ctx.fused_mlp:
            H1 = W1.shape[0]
            dm = torch.empty((L * K, C), dtype=torch.float32, device=g.device)
            ws = torch.empty(int(lib().htd_ba_mlp_workspace_floats(L * K, C)), dtype=torch.float32,
                             device=g.device)
            dW1 = torch.empty(w1_shape, dtype=W1.dtype, device=g.device)
            db1 = torch.empty(H1, dtype=W1.dtype, device=g.device)
            dW2 = torch.empty(w2_shape, dtype=W1.dtype, device=g.device)
            db2 = torch.empty(1, dtype=W1.dtype, device=g.device)
            check(lib().htd_ba_mlp_bwd(ptr(da), ptr(h), ptr(m), L * K, C, H1, ptr(W1), ptr(W2),
                                       dt(W1), 1.0 / PP, ptr(dm), ptr(ws), ptr(dW1), ptr(db1),
                                       ptr(dW2), ptr(db2), stream()), 'htd_ba_mlp_bwd')
        else:
            da_f = da.reshape(L * K, 1)
            dW2 = (da_f * h).sum(0).reshape(w2_shape)
            db2 = da_f.sum().reshape(1)
            dpre = (da_f * W2) * (1.0 - h * h)
            dW1 = (dpre.t() @ m).reshape(w1_shape)
            db1 = dpre.sum(0)
            dm = (dpre @ W1) * (1.0 / PP)                  # [L*K, C], gradient of the bin mean
        grads = [None] * L
        gtoken = None
        if ctx.deferred:
            ctx.sink.park(scales, pooled, dict(rois=rois, plan=tuple(ctx.saved_tensors[7:]), dy=g,
                                               dy_per_level=False, scale=wts, ring_edge=edge,
                                               addvec=dm.contiguous()))
            gtoken = torch.zeros(1, dtype=torch.float32, device=g.device)
        elif ctx.with_dx:
            grads = _roi_align_bwd(shapes, fdt, scales, rois, ctx.saved_tensors[7:], pooled, g,
                                   dy_per_level=False, scale=wts, ring_edge=edge,
                                   addvec=dm.contiguous())
        dadd = g.permute(0, 3, 1, 2) if (has_add and ctx.needs_input_grad[5]) else None
        dbias = None
        if bias_shape is not None and ctx.needs_input_grad[6]:
            dbias = _bias_grad(g, rois, B).reshape(bias_shape)
        return (None, dW1, db1, dW2, db2, dadd, dbias, None, None, None, None, gtoken, None) + \
            tuple(grads)


def ba_extract(feats_cl, rois, scales, conv1, conv2, pooled=7, sampling_ratio=0, edge=1, add=None,
               bias=None):
    return _BAFunction.apply(rois, conv1.weight, conv1.bias, conv2.weight, conv2.bias, add, bias,
                             tuple(float(s) for s in scales), int(pooled), int(sampling_ratio),
                             int(edge), getattr(feats_cl, 'token', None),
                             getattr(feats_cl, 'sink', None), *feats_cl)


# ------------------------------------------------------------------------------------------
# fused GroupNorm + ReLU (regression conv tower)
# ------------------------------------------------------------------------------------------
class _GroupNormReLU(torch.autograd.Function):
    """relu(group_norm(x)) on channels-last [N,C,H,W] tensors: one launch forward, two backward
    (csrc/gn_relu.cu) instead of ATen's moments / affine / relu / internal-gradients / backward /
    parameter-reduction kernels."""

    @staticmethod
    def forward(ctx, x, weight, bias, groups, eps):
        _lib.require_cuda(x, weight, bias)
        if not _is_cl(x):
            x = x.contiguous(memory_format=torch.channels_last)
        N, C, H, W = x.shape
        y = torch.empty_like(x)               # preserves channels_last
        sdt = torch.float64 if x.dtype == torch.float32 else torch.float32
        mean = torch.empty(N * groups, dtype=sdt, device=x.device)
        rstd = torch.empty_like(mean)
        # affine parameters in the tensor dtype (no conversion launches in the bf16 configuration)
        gamma = weight.detach().to(x.dtype).contiguous()
        beta = bias.detach().to(x.dtype).contiguous()
        check(lib().htd_gn_relu_fwd(ptr(x), dt(x), N, H * W, C, int(groups), ptr(gamma), ptr(beta),
                                    float(eps), ptr(y), ptr(mean), ptr(rstd), stream()),
              'htd_gn_relu_fwd')
        ctx.save_for_backward(x, mean, rstd, gamma, beta)
        ctx.cfg = (int(groups), weight.dtype, bias.dtype)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd, gamma, beta = ctx.saved_tensors
        groups, wdt, bdt = ctx.cfg
        N, C, H, W = x.shape
        if not _is_cl(dy) or dy.dtype != x.dtype:
            dy = dy.to(x.dtype).contiguous(memory_format=torch.channels_last)
        dx = torch.empty_like(x)
        part = torch.empty((2, max(N, 1), C), dtype=torch.float32, device=x.device)
        dgamma = torch.empty(C, dtype=x.dtype, device=x.device)
        dbeta = torch.empty_like(dgamma)
        check(lib().htd_gn_relu_bwd(ptr(x), ptr(dy), dt(x), ptr(mean), ptr(rstd), ptr(gamma),
                                    ptr(beta), N, H * W, C, groups, ptr(dx), ptr(part), ptr(dgamma),
                                    ptr(dbeta), stream()), 'htd_gn_relu_bwd')
        return dx, dgamma.to(wdt), dbeta.to(bdt), None, None


def group_norm_relu(x, weight, bias, groups, eps=1e-5):
    return _GroupNormReLU.apply(x, weight, bias, groups, eps)


class _ReLUMeanPool(torch.autograd.Function):
    """mean over (H, W) of relu(x) for a channels-last [N,C,H,W] tensor -> [N,C]: one pass forward,
    one pass backward (csrc/gn_relu.cu) instead of relu / mean / expand-div / threshold_backward."""

    @staticmethod
    def forward(ctx, x):
        _lib.require_cuda(x)
        if not _is_cl(x):
            x = x.contiguous(memory_format=torch.channels_last)
        N, C, H, W = x.shape
        y = torch.empty((N, C), dtype=x.dtype, device=x.device)
        check(lib().htd_relu_mean_fwd(ptr(x), dt(x), N, H * W, C, ptr(y), stream()),
              'htd_relu_mean_fwd')
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, g):
        x, = ctx.saved_tensors
        N, C, H, W = x.shape
        g = g.to(x.dtype).contiguous()
        dx = torch.empty_like(x)
        check(lib().htd_relu_mean_bwd(ptr(x), ptr(g), dt(x), N, H * W, C, ptr(dx), stream()),
              'htd_relu_mean_bwd')
        return dx


def relu_mean_pool(x):
    return _ReLUMeanPool.apply(x)


# ------------------------------------------------------------------------------------------
# target / loss / decode glue (csrc/rcnn_glue.cu)
# ------------------------------------------------------------------------------------------
def _f4(v):
    return (ctypes.c_float * 4)(*[float(t) for t in v])


def bbox_targets(boxes, gt_boxes, gt_labels, is_pos, num_classes, pos_weight, means, stds):
    """labels, label_weights, bbox_targets, bbox_weights of all sampled RoIs in one launch."""
    _lib.require_cuda(boxes, gt_boxes, gt_labels, is_pos)
    K = boxes.shape[0]
    dev = boxes.device
    boxes = boxes.detach().float().contiguous()
    gt_boxes = gt_boxes.detach().float().contiguous()
    gt_labels = gt_labels.detach().long().contiguous()
    labels = torch.empty(K, dtype=torch.long, device=dev)
    lw = torch.empty(K, dtype=torch.float32, device=dev)
    bt = torch.empty((K, 4), dtype=torch.float32, device=dev)
    bw = torch.empty((K, 4), dtype=torch.float32, device=dev)
    check(lib().htd_bbox_targets(ptr(boxes), ptr(gt_boxes), ptr(gt_labels), ptr(is_pos), K,
                                 int(num_classes), float(pos_weight), _f4(means), _f4(stds),
                                 ptr(labels), ptr(lw), ptr(bt), ptr(bw), stream()),
          'htd_bbox_targets')
    return labels, lw, bt, bw


def bbox_decode(rois, deltas, means, stds, max_shape=None, wh_ratio_clip=16 / 1000, clip=True):
    """delta2bbox + clip for class-agnostic deltas; rois [K,4] or [K,5] -> same shape."""
    _lib.require_cuda(rois, deltas)
    K, rs = rois.shape
    rois = rois.detach().float().contiguous()
    deltas = deltas.detach()
    if deltas.dtype not in _lib._DT:
        deltas = deltas.float()
    deltas = deltas.contiguous()
    out = torch.empty((K, rs), dtype=torch.float32, device=rois.device)
    do_clip = bool(clip and max_shape is not None)
    mh, mw = (float(max_shape[0]), float(max_shape[1])) if do_clip else (0.0, 0.0)
    check(lib().htd_bbox_decode(ptr(rois), rs, ptr(deltas), dt(deltas), K, _f4(means), _f4(stds),
                                float(wh_ratio_clip), int(do_clip), mh, mw, ptr(out), rs, stream()),
          'htd_bbox_decode')
    return out


class _RCNNLoss(torch.autograd.Function):
    """BBoxHead.loss for class-agnostic regression (bbox_head.py:141-186): weighted softmax CE /
    avg_factor, top-1 accuracy, smooth-L1 on positives / K - forward computes the (unnormalised)
    gradients too, backward only scales them by the incoming loss gradients."""

    @staticmethod
    def forward(ctx, cls_score, bbox_pred, labels, label_weights, bbox_targets, bbox_weights,
                num_classes, beta, w_cls, w_bbox, pad_rows=False):
        _lib.require_cuda(cls_score, bbox_pred, labels)
        K, C1 = cls_score.shape
        dev = cls_score.device
        cs = cls_score.detach().contiguous()
        bp = bbox_pred.detach().to(cs.dtype).contiguous()
        dcls = torch.empty_like(cs)
        dbbox = torch.empty_like(bp)
        nblk = max((K + 7) // 8, 1)
        partial = torch.empty((nblk, 4), dtype=torch.float32, device=dev)
        out4 = torch.empty(4, dtype=torch.float32, device=dev)
        # converted copies are bound to locals: ptr() keeps only the address, and a temporary
        # would be freed (and its block reused by the next conversion) before the launch
        lab_c = labels.contiguous()
        lw_c = label_weights.float().contiguous()
        bt_c = bbox_targets.float().contiguous()
        bw_c = bbox_weights.float().contiguous()
        check(lib().htd_rcnn_loss_fwd(ptr(cs), C1, ptr(bp), dt(cs), ptr(lab_c), ptr(lw_c),
                                      ptr(bt_c), ptr(bw_c), K, int(num_classes),
                                      float(beta), float(w_cls), float(w_bbox), int(bool(pad_rows)),
                                      ptr(dcls), ptr(dbbox), ptr(partial), ptr(out4), stream()),
              'htd_rcnn_loss_fwd')
        ctx.save_for_backward(dcls, dbbox, out4)
        ctx.cfg = (float(w_cls), float(w_bbox), K, cls_score.dtype, bbox_pred.dtype,
                   int(bool(pad_rows)))
        loss_cls, acc, loss_bbox = out4[0], out4[1:2], out4[2]
        ctx.mark_non_differentiable(acc)
        return loss_cls, acc, loss_bbox

    @staticmethod
    def backward(ctx, g_cls, g_acc, g_bbox):
        dcls, dbbox, out4 = ctx.saved_tensors
        w_cls, w_bbox, K, cdt, bdt, pad_rows = ctx.cfg
        g_cls = None if g_cls is None else g_cls.detach().float().contiguous()
        g_bbox = None if g_bbox is None else g_bbox.detach().float().contiguous()
        gc, gb = torch.empty_like(dcls), torch.empty_like(dbbox)   # saved buffers stay untouched
        check(lib().htd_rcnn_loss_bwd(ptr(dcls), dcls.numel(), ptr(dbbox), dbbox.numel(), dt(dcls),
                                      ptr(g_cls), ptr(g_bbox), ptr(out4), w_cls, w_bbox, K, pad_rows,
                                      ptr(gc), ptr(gb), stream()), 'htd_rcnn_loss_bwd')
        return gc.to(cdt), gb.to(bdt), None, None, None, None, None, None, None, None, None


def rcnn_loss(cls_score, bbox_pred, labels, label_weights, bbox_targets, bbox_weights, num_classes,
              beta=1.0, w_cls=1.0, w_bbox=1.0, pad_rows=False):
    """``pad_rows``: rows with label weight 0 are pad rows of ``assign_sample`` (left out of the
    accuracy; accuracy / loss_bbox are averaged over the real rows)."""
    return _RCNNLoss.apply(cls_score, bbox_pred, labels, label_weights, bbox_targets, bbox_weights,
                           num_classes, beta, w_cls, w_bbox, pad_rows)


class StaticSample:
    """Result of ``assign_sample``: ``num`` rows per image (positives, negatives, pad rows), all
    device tensors of static shape.  ``rois`` [B*num,5]; ``kind`` [B*num] uint8 1/0/2 =
    positive/negative/pad; ``gt_boxes`` [B*num,4] / ``gt_labels`` [B*num] / ``gt_index`` [B*num] of
    the positives' matched gt; ``is_gt`` [B*num] row is an appended gt box; ``cand`` [B*num] index
    into the reference's cat([gt_bboxes, bboxes]); ``counts`` [B,4] int32 = sampled positives,
    sampled negatives, positive candidates, negative candidates."""
    __slots__ = ('rois', 'kind', 'gt_boxes', 'gt_labels', 'gt_index', 'is_gt', 'cand', 'counts',
                 'gt_inds', 'max_overlaps', 'num', 'num_pos', 'num_imgs')


def assign_sample(props, gt_boxes, gt_labels, num_gt, keys, valid=None, pos_iou_thr=0.5,
                  neg_iou_thr=0.5, min_pos_iou=0.5, match_low_quality=False,
                  add_gt_as_proposals=True, num=512, pos_fraction=0.25, neg_pos_ub=-1,
                  want_assignment=False):
    """MaxIoUAssigner + RandomSampler of one stage for all images in one launch (csrc/
    assign_sample.cu).  props [B,N,4], gt_boxes [B,G,4], gt_labels [B,G], num_gt [B] (device
    int32), keys [B,G+N] uniform random numbers, valid [B,N] bool/uint8 or None."""
    _lib.require_cuda(props, gt_boxes, gt_labels, num_gt, keys)
    B, N = props.shape[:2]
    G = gt_boxes.shape[1]
    dev = props.device
    assert gt_boxes.shape[0] == B and gt_labels.shape == (B, G) and keys.shape == (B, G + N), \
        (props.shape, gt_boxes.shape, gt_labels.shape, keys.shape)
    props = props.detach().float().contiguous()
    gt_boxes = gt_boxes.detach().float().contiguous()
    gt_labels = gt_labels.detach().long().contiguous()
    num_gt = num_gt.detach().to(torch.int32).contiguous()
    keys = keys.detach().float().contiguous()
    if valid is not None:
        assert valid.shape == (B, N)
        valid = valid.detach().to(torch.uint8).contiguous()
    K = B * num
    s = StaticSample()
    s.num, s.num_pos, s.num_imgs = int(num), int(num * pos_fraction), B
    s.rois = torch.empty((K, 5), dtype=torch.float32, device=dev)
    s.kind = torch.empty(K, dtype=torch.uint8, device=dev)
    s.gt_boxes = torch.empty((K, 4), dtype=torch.float32, device=dev)
    s.gt_labels = torch.empty(K, dtype=torch.long, device=dev)
    s.is_gt = torch.empty(K, dtype=torch.uint8, device=dev)
    s.cand = torch.empty(K, dtype=torch.int32, device=dev)
    s.gt_index = torch.empty(K, dtype=torch.int32, device=dev)
    s.counts = torch.empty((B, 4), dtype=torch.int32, device=dev)
    s.gt_inds = torch.empty((B, G + N), dtype=torch.int32, device=dev) if want_assignment else None
    s.max_overlaps = torch.empty((B, G + N), dtype=torch.float32, device=dev) \
        if want_assignment else None
    check(lib().htd_assign_sample(ptr(props), ptr(valid), B, N, ptr(gt_boxes), ptr(gt_labels),
                                  ptr(num_gt), G, ptr(keys), float(pos_iou_thr), float(neg_iou_thr),
                                  float(min_pos_iou), int(bool(match_low_quality)),
                                  int(bool(add_gt_as_proposals)), int(num), s.num_pos,
                                  float(neg_pos_ub), ptr(s.rois), ptr(s.kind), ptr(s.gt_boxes),
                                  ptr(s.gt_labels), ptr(s.is_gt), ptr(s.cand), ptr(s.gt_index),
                                  ptr(s.counts), ptr(s.gt_inds), ptr(s.max_overlaps), stream()),
          'htd_assign_sample')
    return s


def multiclass_nms(multi_bboxes, multi_scores, score_thr, iou_thr, max_num, soft=None):
    """multiclass_nms (core/post_processing/bbox_nms.py:7-71) on the device, no host sync: returns
    det [max_num,5], labels [max_num] (int64) and count [1] (int32, device); rows >= count are
    unspecified.  multi_bboxes [K,4] or [K,C*4], multi_scores [K,C+1].  ``soft`` = None: hard NMS;
    dict(min_score=..., method='linear'|'naive'): mmcv's soft_nms (htd_multiclass_soft_nms)."""
    _lib.require_cuda(multi_bboxes, multi_scores)
    K, C1 = multi_scores.shape
    C = C1 - 1
    bc = multi_bboxes.shape[1] // 4
    assert bc in (1, C), multi_bboxes.shape
    dev = multi_scores.device
    boxes = multi_bboxes.detach().float().contiguous()
    scores = multi_scores.detach().float().contiguous()
    max_num = int(max_num) if max_num > 0 else max(K * C, 1)
    det = torch.empty((max_num, 5), dtype=torch.float32, device=dev)
    labels = torch.empty(max_num, dtype=torch.long, device=dev)
    count = torch.empty(1, dtype=torch.int32, device=dev)
    if soft is not None:
        method = {'naive': 0, 'linear': 1}.get(soft.get('method', 'linear'))
        if method is None:
            raise NotImplementedError("soft_nms method 'gaussian' is not provided (configs/htd use "
                                      "the mmcv default 'linear')")
        ws = torch.empty(int(lib().htd_multiclass_soft_nms_workspace_bytes(K, C)),
                         dtype=torch.uint8, device=dev)
        check(lib().htd_multiclass_soft_nms(ptr(boxes), bc, ptr(scores), K, C, float(score_thr),
                                            float(iou_thr), float(soft.get('min_score', 1e-3)),
                                            method, max_num, ptr(det), ptr(labels), ptr(count),
                                            ptr(ws), stream()), 'htd_multiclass_soft_nms')
        return det, labels, count
    ws = torch.empty(int(lib().htd_multiclass_nms_workspace_bytes(K, C)), dtype=torch.uint8,
                     device=dev)
    check(lib().htd_multiclass_nms(ptr(boxes), bc, ptr(scores), K, C, float(score_thr),
                                   float(iou_thr), max_num, ptr(det), ptr(labels), ptr(count),
                                   ptr(ws), stream()), 'htd_multiclass_nms')
    return det, labels, count


def topk_sorted(keys, k):
    """The k largest of every row of a [rows, n] fp32 tensor (rows may be strided) in descending
    order, equal keys by ascending position: (values [rows, k], positions [rows, k] int64).
    htd_topk_sorted - the ranking of the RPN proposal path (rpn_head.py:125-134)."""
    _lib.require_cuda(keys)
    assert keys.dim() == 2 and keys.dtype == torch.float32 and keys.stride(1) == 1, \
        (keys.shape, keys.dtype, keys.stride())
    rows, n = keys.shape
    vals = torch.empty((rows, k), dtype=torch.float32, device=keys.device)
    idx = torch.empty((rows, k), dtype=torch.int32, device=keys.device)
    check(lib().htd_topk_sorted(ptr(keys), keys.stride(0) if rows > 1 else n, rows, n, int(k), ptr(vals),
                                ptr(idx), stream()), 'htd_topk_sorted')
    return vals, idx.long()

