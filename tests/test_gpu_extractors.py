"""GPU parity tests (through the C-ABI) of level assignment, multi-level RoIAlign forward /
atomic-free backward and the BA extractor, against the CPU oracle (fp64) on the same seeded
inputs and against the golden fixtures generated from the reference.

Tolerances (BASELINE.json north_star / SURVEY F12): indices bit-exact; fp32 features and
gradients max|a-b|/max|b| <= 1e-5 against the fp64 oracle; bf16 <= 2e-2.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

from htd_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
TOL_F32, TOL_BF16 = 1e-5, 2e-2
ROI_LAYER = dict(type='RoIAlign', output_size=7, sampling_ratio=0)


class _Ext(nn.Module):
    def __init__(self):
        super().__init__()
        from htd_b200.roi_extractors import AdptRoIExtractor, SingleRoIExtractor
        self.bbox_roi_extractor = nn.ModuleList([
            SingleRoIExtractor(dict(ROI_LAYER), 256, [4, 8, 16, 32]),
            AdptRoIExtractor(edge=1, roi_layer=dict(ROI_LAYER), out_channels=256,
                             featmap_strides=[4, 8, 16, 32])])


def _oracle_outs(name):
    from oracle import cases, restate
    c = cases.CASES[name]
    head = restate.HTDRoIHead().double()
    synth.fill_params_(head, c['scheme'], c['seed'])
    return cases.run_extractors(head, name, torch.float64)


_ORACLE_CACHE = {}


def oracle_outs(name):
    if name not in _ORACLE_CACHE:
        _ORACLE_CACHE[name] = _oracle_outs(name)
    return _ORACLE_CACHE[name]


def _product_outs(name, dtype):
    from oracle import cases
    c = cases.CASES[name]
    ext = _Ext()
    synth.fill_params_(ext, c['scheme'], c['seed'])
    ext = ext.cuda()
    return cases.run_extractors(ext, name, dtype, device='cuda')


def test_level_assign_bit_exact_golden():
    from htd_b200 import ops
    z = np.load(os.path.join(GOLD, 'levels.npz'))
    rois = torch.from_numpy(z['rois']).cuda()
    lv = ops.level_assign(rois, 4, 56.0)
    assert np.array_equal(lv.cpu().numpy().astype(np.int8), z['levels'])


@pytest.mark.parametrize('name', ['small', 'mid'])
def test_extractors_fp32_vs_fp64_oracle_and_golden(name):
    from oracle import cases
    got = _product_outs(name, torch.float32)
    want = oracle_outs(name)
    assert torch.equal(got['levels'].cpu(), want['levels'])
    errs = {k: cases.rel_err(got[k], want[k]) for k in want if k != 'levels'}
    bad = {k: v for k, v in errs.items() if not v <= TOL_F32}
    assert not bad, bad
    fix = cases.load_fixture(os.path.join(GOLD, f'{name}_f64.npz'))
    cases.compare_to_fixture(got, fix, TOL_F32, names=set(want))
    print({k: f'{v:.1e}' for k, v in errs.items()})


@pytest.mark.parametrize('name', ['small'])
def test_extractors_bf16(name):
    from oracle import cases
    got = _product_outs(name, torch.bfloat16)
    want = oracle_outs(name)
    errs = {k: cases.rel_err(got[k].float(), want[k]) for k in want if k != 'levels'}
    bad = {k: v for k, v in errs.items() if not v <= TOL_BF16}
    assert not bad, bad


def test_backward_is_deterministic():
    a = _product_outs('small', torch.float32)
    b = _product_outs('small', torch.float32)
    for k in a:
        assert torch.equal(a[k], b[k]), k


def _edge_rois(H, W, stride):
    return torch.tensor([
        [0, 10.3, 7.9, 60.2, 51.1], [0, -20.0, -14.0, 30.0, 25.0],
        [1, W * stride - 30.0, H * stride - 20.0, W * stride + 90.0, H * stride + 70.0],
        [1, 33.0, 21.0, 33.0, 21.0], [0, 0.0, 0.0, W * stride, H * stride],
        [1, 50.0, 40.0, 52.5, 41.0], [0, 70.0, 30.0, 20.0, 10.0],
        [0, -500.0, -500.0, -300.0, -300.0], [1, -1e4, -1e4, 1e4, 1e4],
        [1, 3.0, 5.0, 3.0 + 7 * stride, 5.0 + 14 * stride]])


@pytest.mark.parametrize('C,stride,sr', [(256, 4, 0), (64, 8, 0), (40, 16, 2), (512, 4, 0)])
def test_roialign_module_edge_cases(C, stride, sr):
    """mmcv.ops.RoIAlign drop-in: NCHW in, contiguous NCHW out; empty / degenerate / outside /
    whole-image RoIs; C not a multiple of 256; fixed sampling_ratio."""
    from htd_b200.roi_extractors import RoIAlign
    from oracle import cases, restate
    H, W = 23, 37
    g = torch.Generator().manual_seed(C + stride)
    x = torch.randn(2, C, H, W, generator=g)
    rois = _edge_rois(H, W, stride)
    if sr > 0:
        rois = rois[(rois[:, 3] >= rois[:, 1]) & (rois[:, 4] >= rois[:, 2])]
    xo = x.double().requires_grad_(True)
    yo = restate.RoIAlign(7, 1.0 / stride, sr)(xo, rois.double())
    dy = torch.randn(yo.shape, generator=g)
    gxo, = torch.autograd.grad((yo * dy.double()).sum(), xo)
    xg = x.cuda().requires_grad_(True)
    layer = RoIAlign(7, 1.0 / stride, sr)
    yg = layer(xg, rois.cuda())
    assert yg.is_contiguous() and yg.shape == yo.shape
    gxg, = torch.autograd.grad((yg * dy.cuda()).sum(), xg)
    assert gxg.shape == x.shape
    assert cases.rel_err(yg, yo) <= TOL_F32
    assert cases.rel_err(gxg, gxo) <= TOL_F32
    # empty roi set
    ye = layer(xg, torch.zeros(0, 5, device='cuda'))
    assert ye.shape == (0, C, 7, 7)


def test_sfa_bias_fused_and_bias_grad():
    from htd_b200 import ops
    from oracle import cases
    c, x, props, gts, shapes = cases.case_inputs('small', torch.float32, 'cuda')
    rois = cases._rois(props)
    ext = _Ext().cuda().bbox_roi_extractor[0]
    bias = torch.randn(c['B'], 256, 1, 1, device='cuda', requires_grad=True)
    xs = [t.clone().requires_grad_(True) for t in x[:4]]
    y0 = ext(xs, rois)
    y1 = ext(xs, rois, bias=bias)
    want = y0 + bias[rois[:, 0].long()]
    assert cases.rel_err(y1, want) <= 1e-6
    dy = torch.randn_like(y1)
    gb, = torch.autograd.grad((y1 * dy).sum(), bias)
    wb = torch.zeros_like(bias).index_add_(0, rois[:, 0].long(), dy.sum((2, 3), keepdim=True))
    assert cases.rel_err(gb, wb) <= 1e-5


def test_layout_roundtrip():
    from htd_b200 import ops
    x = torch.randn(3, 72, 19, 45, device='cuda', requires_grad=True)
    y = ops.to_channels_last(x, torch.float32)
    assert y.is_contiguous(memory_format=torch.channels_last) and torch.equal(y, x)
    g = torch.randn_like(x)
    gx, = torch.autograd.grad((y * g).sum(), x)
    assert torch.equal(gx, g) and gx.is_contiguous()
    yb = ops.to_channels_last(x, torch.bfloat16)
    assert torch.equal(yb, x.to(torch.bfloat16))


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_flatten_with_sfa_bias(dtype):
    """ops.flatten_roi_feats(x, bias, rois) = (x + bias[image]).flatten(1) and its gradients
    (htd_roi_flatten: the SFA vector added while the RoI maps change to the FC flatten order)."""
    from htd_b200 import ops
    from oracle import cases
    g = torch.Generator(device='cuda').manual_seed(3)
    K, C, B = 37, 256, 3
    x = torch.randn(K, 7, 7, C, device='cuda', generator=g).to(dtype).permute(0, 3, 1, 2).requires_grad_(True)
    bias = torch.randn(B, C, 1, 1, device='cuda', generator=g).to(dtype).requires_grad_(True)
    rois = torch.rand(K, 5, device='cuda', generator=g) * 100
    rois[:, 0] = torch.randint(0, B, (K,), device='cuda', generator=g).float()
    assert ops.flatten_fuses_bias(x)
    y = ops.flatten_roi_feats(x, bias, rois)
    want = (x.float() + bias.float()[rois[:, 0].long()]).flatten(1)
    tol = 1e-6 if dtype == torch.float32 else 8e-3
    assert y.shape == want.shape and cases.rel_err(y.float(), want) <= tol
    dy = torch.randn_like(y)
    gx, gb = torch.autograd.grad((y.float() * dy.float()).sum(), (x, bias))
    wx, wb = torch.autograd.grad((want * dy.float()).sum(), (x, bias))
    assert cases.rel_err(gx.float(), wx.float()) <= tol and cases.rel_err(gb.float(), wb.float()) <= tol


@pytest.mark.parametrize('rows,C', [(1024, 256), (37, 256), (8, 32), (260, 128)])
@pytest.mark.parametrize('pd', ['f32', 'bf16'])
def test_ba_attention_mlp_fused(rows, C, pd):
    """htd_ba_mlp_fwd / _bwd (tanh MLP of the BA attention on the bin means) against the same
    expressions in torch fp32 on the parameter values the kernel reads."""
    from htd_b200 import _lib
    from htd_b200._lib import check, dt, lib, ptr, stream
    T = dict(f32=torch.float32, bf16=torch.bfloat16)[pd]
    g = torch.Generator(device='cuda').manual_seed(rows + C)
    rnd = lambda *s: torch.randn(*s, device='cuda', generator=g)
    m = rnd(rows, C)
    w1, b1, w2, b2 = (0.08 * rnd(128, C)).to(T), (0.1 * rnd(128)).to(T), rnd(128).to(T), rnd(1).to(T)
    h = torch.empty(rows, 128, device='cuda')
    logits = torch.empty(rows, device='cuda')
    check(lib().htd_ba_mlp_fwd(ptr(m), rows, C, 128, ptr(w1), ptr(b1), ptr(w2), ptr(b2), dt(w1),
                               ptr(h), ptr(logits), stream()), 'htd_ba_mlp_fwd')
    W1, B1, W2, B2 = w1.float(), b1.float(), w2.float(), b2.float()
    saved = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        h_ref = torch.tanh(m @ W1.t() + B1)
        l_ref = h_ref @ W2 + B2
        assert (h - h_ref).abs().max() <= 5e-6 and (logits - l_ref).abs().max() <= 5e-5   # fp32 sums, other order
        da = rnd(rows)
        dm = torch.empty(rows, C, device='cuda')
        ws = torch.empty(int(lib().htd_ba_mlp_workspace_floats(rows, C)), device='cuda')
        dw1, db1 = torch.empty_like(w1), torch.empty_like(b1)
        dw2, db2 = torch.empty_like(w2), torch.empty_like(b2)
        check(lib().htd_ba_mlp_bwd(ptr(da), ptr(h), ptr(m), rows, C, 128, ptr(w1), ptr(w2), dt(w1),
                                   1.0 / 49, ptr(dm), ptr(ws), ptr(dw1), ptr(db1), ptr(dw2), ptr(db2),
                                   stream()), 'htd_ba_mlp_bwd')
        dpre = (da[:, None] * W2[None, :]) * (1 - h_ref * h_ref)
        tol = 1e-5 if pd == 'f32' else 8e-3
        from oracle import cases
        assert cases.rel_err(dm, dpre @ W1 / 49) <= 1e-5
        assert cases.rel_err(dw1.float(), dpre.t() @ m) <= tol
        assert cases.rel_err(db1.float(), dpre.sum(0)) <= tol
        assert cases.rel_err(dw2.float(), (da[:, None] * h_ref).sum(0)) <= tol
        assert cases.rel_err(db2.float(), da.sum().reshape(1)) <= tol
    finally:
        torch.backends.cuda.matmul.allow_tf32 = saved


@pytest.mark.parametrize('N,R,S', [
    (5, 49, 256), (5, 256, 49),            # whole-matrix form, both directions of the flatten order
    (2, 256, 200 * 336), (2, 256, 100 * 167), (2, 256, 25 * 42), (2, 256, 13 * 21),  # tiled, widths 4/4/2/1
    (3, 72, 855), (1, 7, 3), (2, 64, 64), (4, 100, 96),
])
@pytest.mark.parametrize('sd,dd', [('f32', 'bf16'), ('bf16', 'f32'), ('bf16', 'bf16'), ('f32', 'f32')])
def test_layout_convert_forms(N, R, S, sd, dd):
    """htd_layout_convert [N,R,S] -> [N,S,R] in every kernel form the host dispatch picks
    (whole-matrix / tiled, each vector width), bit-exact against a torch permute."""
    from htd_b200 import ops
    T = dict(f32=torch.float32, bf16=torch.bfloat16)
    g = torch.Generator(device='cuda').manual_seed(N * 1000 + R + S)
    src = torch.randn(N, R, S, device='cuda', generator=g).to(T[sd])
    dst = torch.full((N, S, R), float('nan'), device='cuda', dtype=T[dd])
    ops._convert(src, dst, N, R, S)
    want = src.permute(0, 2, 1).to(T[dd])
    assert torch.equal(dst, want)


def test_full_size_forward_properties():
    """BASELINE config sizes (800x1333, 512 RoIs): linearity in the features and agreement of
    the single-level extractor with the all-level sampler on the assigned level; oracle compare
    on a bounded subset."""
    from htd_b200 import ops
    from oracle import cases, restate
    x = [t.cuda() for t in synth.make_pyramid(1)[:4]]
    props = synth.make_proposals(1, 512)
    rois = cases._rois(props).cuda()
    ext = _Ext().cuda()
    e0 = ext.bbox_roi_extractor[0]
    y = e0(x, rois)
    y2 = e0([2.0 * t for t in x], rois)
    assert cases.rel_err(y2, 2.0 * y) <= 1e-6
    xcl = [ops.to_channels_last(t) for t in x]
    allv = ops.roi_align_levels(xcl, rois, [0.25, 0.125, 0.0625, 0.03125])
    lv = e0.map_roi_levels(rois, 4)
    pick = allv[lv, torch.arange(rois.shape[0], device='cuda')]
    assert torch.equal(pick, y)
    sub = torch.arange(0, 512, 16)
    ref = restate.SingleRoIExtractor()([t.cpu().double() for t in x], rois.cpu().double()[sub])
    assert cases.rel_err(y[sub.cuda()], ref) <= TOL_F32


def test_library_has_no_cpu_path():
    from htd_b200 import ops
    with pytest.raises(RuntimeError):
        ops.level_assign(torch.zeros(4, 5), 4)


@pytest.mark.parametrize('dtype,tol', [(torch.float32, 1e-5), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize('C,G,HW', [(576, 36, (7, 7)), (64, 8, (5, 3)), (96, 4, (7, 7))])
def test_fused_group_norm_relu(dtype, tol, C, G, HW):
    """csrc/gn_relu.cu against torch's fp64 GroupNorm + ReLU (forward, dx, dgamma, dbeta)."""
    from htd_b200 import ops
    from oracle import cases
    g = torch.Generator().manual_seed(C + G)
    N = 24
    x = (torch.randn(N, C, *HW, generator=g) * 1.5 + 0.3).to(dtype)
    w = (1.0 + 0.2 * torch.randn(C, generator=g)).to(dtype)
    b = (0.2 * torch.randn(C, generator=g)).to(dtype)
    dy = torch.randn(N, C, *HW, generator=g).to(dtype)
    xo, wo, bo = (t.double().requires_grad_(True) for t in (x, w, b))
    yo = torch.relu(torch.nn.functional.group_norm(xo, G, wo, bo, 1e-5))
    go = torch.autograd.grad((yo * dy.double()).sum(), [xo, wo, bo])
    xg = x.cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    wg, bg = w.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    yg = ops.group_norm_relu(xg, wg, bg, G)
    gg = torch.autograd.grad((yg.float() * dy.cuda().float()).sum(), [xg, wg, bg])
    assert yg.is_contiguous(memory_format=torch.channels_last)
    assert cases.rel_err(yg.float(), yo) <= tol
    # gradients: the ReLU mask is recomputed from x - in bf16 an activation within rounding of
    # zero can flip, so compare dx in relative L2 there (max-norm in fp32)
    for a, want, name in zip(gg, go, ('dx', 'dgamma', 'dbeta')):
        if dtype == torch.float32:
            assert cases.rel_err(a.float(), want) <= tol, name
        else:
            num = (a.double().cpu() - want).norm() / want.norm()
            assert float(num) <= 3e-2, (name, float(num))


@pytest.mark.parametrize('C,hw,out_dtype', [(256, (200, 336), torch.bfloat16),
                                            (256, (200, 336), torch.float32),
                                            (128, (96, 160), torch.float32),
                                            (64, (640, 704), torch.float32)])
def test_single_level_tensor_pipe_forward(C, hw, out_dtype):
    """Level-assigned extraction of bf16 features (roi_align_fwd_mma_kernel: TMA tensor tiles +
    tensor-pipe x reduction) against the fp64 oracle RoIAlign on the SAME bf16-rounded features:
    fp32 output <= 2e-5 (weights enter as bf16 hi + lo pairs), bf16 output <= 2e-2.  Covers 1 / 2
    / 4 channel chunks, the SFA bias, RoIs outside / degenerate / on the border, and footprints
    wider than 64 px (hw[1] / 32 > 64 on the coarsest level: reduced from global memory)."""
    from htd_b200 import ops
    from oracle import restate
    H, W = hw
    B, strides = 2, [4, 8, 16, 32]
    g = torch.Generator().manual_seed(C + H)
    feats = [torch.randn(B, C, -(-H // (s // 4)), -(-W // (s // 4)), generator=g).bfloat16() for s in strides]
    props = synth.make_proposals(B, 150, H * 4, W * 4, seed=7, min_scale=6.0, max_scale=3.0 * W)
    rois = torch.cat([torch.cat([p.new_full((p.size(0), 1), i), p], 1) for i, p in enumerate(props)])
    rois = torch.cat([rois, _edge_rois(H, W, 4), torch.tensor([[0, 0.0, 0.0, 4.0 * W, 4.0 * H],
                                                               [1, 3.0, 2.0, 4.0 * W - 5, 40.0]])])
    bias = torch.randn(B, C, generator=g)
    # oracle: per-level RoIAlign of the assigned level, fp64, on the bf16 values
    lv_cpu = restate.map_roi_levels(rois, 4) if hasattr(restate, 'map_roi_levels') else None
    x_cl = [ops.to_channels_last(f.cuda(), torch.bfloat16) for f in feats]
    r = rois.cuda()
    lv = ops.level_assign(r, 4)
    if lv_cpu is not None:
        assert torch.equal(lv.cpu().long(), lv_cpu.long())
    want = torch.zeros(rois.size(0), C, 7, 7, dtype=torch.float64)
    for l, s in enumerate(strides):
        idx = (lv.cpu() == l).nonzero().squeeze(1)
        if idx.numel():
            want[idx] = restate.RoIAlign(7, 1.0 / s, 0)(feats[l].double(), rois[idx].double())
    want = want + bias.double()[rois[:, 0].long()][:, :, None, None]
    out = torch.empty(rois.size(0), 7, 7, C, device='cuda', dtype=out_dtype)
    ops._fwd_launch('f', x_cl, [1.0 / s for s in strides], r, lv, 7, 0, bias.cuda().contiguous(), out)
    got = out.permute(0, 3, 1, 2).double().cpu()
    den = want.abs().max().item()
    err = (got - want).abs().max().item() / den
    assert err <= (2e-5 if out_dtype == torch.float32 else TOL_BF16), err
    if W // 8 > 64:
        bx = ops.RoIPlan(x_cl, [1.0 / s for s in strides], r, lv, 7, 0).boxes
        widths = (bx[..., 3] - bx[..., 2] + 1).max().item()
        assert widths > 64, widths                     # the wide-footprint path was taken
