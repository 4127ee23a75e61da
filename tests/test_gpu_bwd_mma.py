"""GPU tests of the bf16 tensor-pipe backward gather (roi_align_bwd_mma_kernel) through the C ABI:
every warp layout / ring depth against the exact fp32 scalar gather on identical inputs (the fp32
gather itself is pinned to the fp64 oracle and the golden fixtures in test_gpu_extractors.py),
the rescan path taken when one tile collects more hits than the record table holds, narrow channel
counts, the NCHW fp32 output form, and bit-reproducibility.

Tolerance: the kernel rounds the interpolation weights and dX to bf16 (2^-9 relative each) and
accumulates in fp32: max|a-b| / max|b| <= 1e-2 (inside the 2e-2 bf16 gate of BASELINE.json).
"""
import pytest
import torch

from htd_b200 import _lib, ops

pytestmark = pytest.mark.gpu
TOL = 1e-2
SCALES = [0.25, 0.125, 0.0625, 0.03125]


def _pyramid_shapes(B, C, H0, W0):
    return [(B, C, max(H0 >> l, 1), max(W0 >> l, 1)) for l in range(4)]


def _rand_rois(K, B, H, W, gen, smin=8., smax=None):
    smax = smax or max(H, W)
    s = torch.exp(torch.empty(K).uniform_(float(torch.tensor(smin).log()), float(torch.tensor(smax).log()),
                                           generator=gen))
    r = torch.exp(torch.empty(K).uniform_(-0.69, 0.69, generator=gen))
    w, h = s * r.sqrt(), s / r.sqrt()
    cx = torch.empty(K).uniform_(0, W, generator=gen)
    cy = torch.empty(K).uniform_(0, H, generator=gen)
    b = torch.randint(0, B, (K,), generator=gen).float()
    return torch.stack([b, (cx - w / 2).clamp(0, W), (cy - h / 2).clamp(0, H),
                        (cx + w / 2).clamp(0, W), (cy + h / 2).clamp(0, H)], 1)


def _sources(shapes, rois, pos, C, gen, with_ba=True, pooled=7):
    """Three sources as in a training step: two single-level extractions and the BA extraction."""
    P = pooled
    dev = 'cuda'
    x = [torch.empty(s, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
         for s in shapes]
    rois, pos = rois.to(dev), pos.to(dev)
    lv = ops.level_assign(rois, 4)
    ps = ops.RoIPlan(x, SCALES, rois, lv, P, 0)
    g1 = torch.randn(rois.shape[0], P, P, C, generator=gen).to(dev).to(torch.bfloat16)
    g2 = torch.randn(rois.shape[0], P, P, C, generator=gen).to(dev).to(torch.bfloat16)
    src = [dict(rois=rois, plan=ps.tensors(), dy=g1, dy_per_level=False),
           dict(rois=rois, plan=ps.tensors(), dy=g2, dy_per_level=False)]
    keep = [ps]
    if with_ba:
        pb = ops.RoIPlan(x, SCALES, pos, None, P, 0)
        gp = torch.randn(pos.shape[0], P, P, C, generator=gen).to(dev).to(torch.bfloat16)
        src.append(dict(rois=pos, plan=pb.tensors(), dy=gp, dy_per_level=False,
                        scale=torch.rand(4, pos.shape[0], generator=gen).to(dev), ring_edge=1,
                        addvec=torch.randn(4 * pos.shape[0], C, generator=gen).to(dev)))
        keep.append(pb)
    return src, keep


def _fp32(src):
    out = []
    for q in src:
        q = dict(q)
        q['dy'] = q['dy'].float()
        out.append(q)
    return out


def _run(shapes, dtype, nchw, src, variant, pooled=7):
    """One backward gather with a forced kernel variant: through the -DHTD_DEBUG_HOOKS build of
    the library (the product library has no variant selection)."""
    with _lib.hooks_library() as L:
        L.htd_debug_set_bwd_variant(variant)
        try:
            out = ops._bwd_multi(shapes, dtype, nchw, SCALES, [dict(q) for q in src], pooled)
            torch.cuda.synchronize()
        finally:
            L.htd_debug_set_bwd_variant(-1)
    return [o.float() for o in out]


def _err(a, b):
    return max(float((x - y).abs().max()) for x, y in zip(a, b)) / \
        max(max(float(y.abs().max()) for y in b), 1e-30)


@pytest.mark.parametrize('variant', [1, 2, 3, 4, 5])
def test_mma_variants_match_exact_fp32_gather(variant):
    gen = torch.Generator().manual_seed(7)
    B, C, H, W = 2, 256, 72, 104                     # ragged tiles on every level
    shapes = _pyramid_shapes(B, C, H, W)
    rois = _rand_rois(700, B, H * 4, W * 4, gen)
    pos = _rand_rois(96, B, H * 4, W * 4, gen)
    src, keep = _sources(shapes, rois, pos, C, gen)
    ref = _run(shapes, torch.float32, False, _fp32(src), 0)
    got = _run(shapes, torch.bfloat16, False, src, variant)
    assert _err(got, ref) <= TOL
    scalar = _run(shapes, torch.bfloat16, False, src, 0)    # the bf16 FFMA gather it replaces
    assert _err(got, scalar) <= TOL


def test_mma_more_hits_than_the_record_table_rescans():
    """3000 + 3000 + 600 RoIs around one spot: a single tile collects several times the 1024 hit
    records of a batch, so the scan is truncated and resumed (chunks rescanned) repeatedly."""
    gen = torch.Generator().manual_seed(11)
    B, C, H, W = 1, 256, 48, 64
    shapes = _pyramid_shapes(B, C, H, W)
    K = 3000
    c = torch.tensor([W * 2.0, H * 2.0])
    d = torch.empty(K, 2).uniform_(-6, 6, generator=gen)
    s = torch.empty(K, 2).uniform_(20, 200, generator=gen)
    rois = torch.cat([torch.zeros(K, 1), (c + d - s / 2).clamp(min=0), (c + d + s / 2)], 1)
    rois[:, 3].clamp_(max=W * 4.0)
    rois[:, 4].clamp_(max=H * 4.0)
    src, keep = _sources(shapes, rois, rois[:600].clone(), C, gen)
    ref = _run(shapes, torch.float32, False, _fp32(src), 0)
    got = _run(shapes, torch.bfloat16, False, src, 3)
    assert _err(got, ref) <= TOL
    got4 = _run(shapes, torch.bfloat16, False, src, 4)
    assert _err(got4, ref) <= TOL


@pytest.mark.parametrize('pooled', [2, 3, 8])
def test_mma_other_output_sizes(pooled):
    """pooled == 8 fills all eight K slots of a bin row, so the BA add vector takes the fp32 rank-1
    form instead of K slot 7; small grids leave most slots at weight 0."""
    gen = torch.Generator().manual_seed(20 + pooled)
    B, C, H, W = 2, 256, 40, 56
    shapes = _pyramid_shapes(B, C, H, W)
    rois = _rand_rois(300, B, H * 4, W * 4, gen)
    pos = _rand_rois(48, B, H * 4, W * 4, gen)
    src, keep = _sources(shapes, rois, pos, C, gen, pooled=pooled)
    ref = _run(shapes, torch.float32, False, _fp32(src), 0, pooled)
    for variant in (3, 4):
        assert _err(_run(shapes, torch.bfloat16, False, src, variant, pooled), ref) <= TOL


@pytest.mark.parametrize('C', [64, 128, 192])
def test_mma_narrow_channel_counts(C):
    gen = torch.Generator().manual_seed(3 + C)
    B, H, W = 2, 40, 56
    shapes = _pyramid_shapes(B, C, H, W)
    rois = _rand_rois(300, B, H * 4, W * 4, gen)
    pos = _rand_rois(40, B, H * 4, W * 4, gen)
    src, keep = _sources(shapes, rois, pos, C, gen)
    ref = _run(shapes, torch.float32, False, _fp32(src), 0)
    for variant in (3, 4):
        assert _err(_run(shapes, torch.bfloat16, False, src, variant), ref) <= TOL


def test_mma_nchw_fp32_output_and_determinism():
    gen = torch.Generator().manual_seed(5)
    B, C, H, W = 2, 256, 40, 56
    shapes = _pyramid_shapes(B, C, H, W)
    rois = _rand_rois(400, B, H * 4, W * 4, gen)
    pos = _rand_rois(64, B, H * 4, W * 4, gen)
    src, keep = _sources(shapes, rois, pos, C, gen)
    ref = _run(shapes, torch.float32, False, _fp32(src), 0)
    nchw = _run(shapes, torch.float32, True, src, 3)          # fp32 NCHW dX from bf16 dY
    assert all(o.is_contiguous() for o in nchw)
    assert _err(nchw, ref) <= TOL
    a = _run(shapes, torch.bfloat16, False, src, 3)
    b = _run(shapes, torch.bfloat16, False, src, 3)
    assert all(torch.equal(x, y) for x, y in zip(a, b)), 'the gather must be bit-reproducible'


def test_bf16_addvec_needs_the_tensor_pipe_kernel():
    """The C ABI rejects a bf16 add vector on the scalar path instead of misreading it."""
    gen = torch.Generator().manual_seed(9)
    B, C, H, W = 1, 256, 24, 32
    shapes = _pyramid_shapes(B, C, H, W)
    pos = _rand_rois(16, B, H * 4, W * 4, gen)
    src, keep = _sources(shapes, pos, pos, C, gen)
    with _lib.hooks_library() as L:
        L.htd_debug_set_bwd_variant(0)
        try:
            with pytest.raises(RuntimeError, match='addvec'):
                q = dict(src[2])
                q['addvec'] = q['addvec'].to(torch.bfloat16)
                _force_bf16_addvec(shapes, q)
        finally:
            L.htd_debug_set_bwd_variant(-1)


def _force_bf16_addvec(shapes, q):
    """Build the source array by hand so the add vector really arrives as bf16."""
    B, C = shapes[0][0], shapes[0][1]
    bufs = [torch.empty((B, s[2], s[3], C), dtype=torch.bfloat16, device='cuda') for s in shapes]
    lv = _lib.make_levels(bufs, SCALES)
    arr = (_lib.HtdBwdSource * 1)()
    boxes, offsets, ranges, weights = q['plan']
    a = arr[0]
    a.rois, a.boxes, a.offsets = q['rois'].data_ptr(), boxes.data_ptr(), offsets.data_ptr()
    a.ranges, a.weights, a.dy = ranges.data_ptr(), weights.data_ptr(), q['dy'].data_ptr()
    a.scale, a.addvec, a.addvec_dtype = q['scale'].data_ptr(), q['addvec'].data_ptr(), 1
    a.K, a.dy_per_level, a.ring_edge = q['rois'].shape[0], 0, 1
    _lib.check(_lib.lib().htd_roi_align_bwd_multi(lv, 4, B, C, 1, 0, arr, 1, 7, 1, _lib.stream()),
               'htd_roi_align_bwd_multi')
