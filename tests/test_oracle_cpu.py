"""CPU tests that PIN the oracle: the C RoIAlign against torchvision-CPU, the restatement
against the reference's own modules (when /root/reference is present) and against the golden
fixtures generated from the reference (anywhere), plus the reference's own known-answer vectors
for the adjacent helpers."""
import os

import numpy as np
import pytest
import torch

from htd_b200 import synth
from oracle import cases, refshim, restate

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _edge_rois():
    return torch.tensor([
        [0, 10.3, 7.9, 60.2, 51.1],      # interior
        [0, -20.0, -14.0, 30.0, 25.0],   # negative coordinates
        [1, 100.0, 60.0, 400.0, 300.0],  # beyond the image
        [1, 33.0, 21.0, 33.0, 21.0],     # zero area
        [0, 0.0, 0.0, 159.0, 99.0],      # full image
        [1, 50.0, 40.0, 52.5, 41.0],     # sub-bin
        [0, 70.0, 30.0, 20.0, 10.0],     # inverted (x2 < x1)
    ])


@pytest.mark.parametrize('dtype', [torch.float32, torch.float64])
@pytest.mark.parametrize('scale', [1.0, 0.25])
def test_c_roialign_equals_torchvision(dtype, scale):
    from torchvision.ops import roi_align
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 5, 25, 40, generator=g).to(dtype).requires_grad_(True)
    rois = _edge_rois().to(dtype)
    rois[:, 1:] /= (0.25 / scale) if scale != 1.0 else 4.0
    rois[:, 1:] *= 1.0 if scale == 1.0 else 4.0
    ya = restate.RoIAlign(7, scale, 0)(x, rois)
    yb = roi_align(x, rois, (7, 7), scale, 0, True)
    tol = 1e-12 if dtype == torch.float64 else 1e-6
    assert (ya - yb).abs().max().item() <= tol
    gy = torch.randn(ya.shape, generator=g).to(dtype)
    ga, = torch.autograd.grad((ya * gy).sum(), x)
    gb, = torch.autograd.grad((yb * gy).sum(), x)
    assert (ga - gb).abs().max().item() <= tol * 10


def test_bbox_overlaps_known_answers():
    # doctest of iou2d_calculator.py:65-86 and the aligned IoU of tests/test_iou2d_calculator.py
    b1 = torch.FloatTensor([[0, 0, 10, 10], [10, 10, 20, 20], [32, 32, 38, 42]])
    b2 = torch.FloatTensor([[0, 0, 10, 20], [0, 10, 10, 19], [10, 10, 20, 20]])
    ov = restate.bbox_overlaps(b1, b2)
    assert ov.shape == (3, 3)
    assert torch.allclose(ov.diag(), torch.tensor([0.5, 0.0, 0.0]))
    assert ov[1, 2].item() == 1.0
    empty = torch.empty(0, 4)
    assert tuple(restate.bbox_overlaps(empty, b1).shape) == (0, 3)
    assert tuple(restate.bbox_overlaps(b1, empty).shape) == (3, 0)


def test_delta2bbox_golden():
    # golden tensor of delta_xywh_bbox_coder.py:148-169
    rois = torch.Tensor([[0., 0., 1., 1.], [0., 0., 1., 1.], [0., 0., 1., 1.], [5., 5., 5., 5.]])
    deltas = torch.Tensor([[0., 0., 0., 0.], [1., 1., 1., 1.], [0., 0., 2., -1.],
                           [0.7, -1.9, -0.5, 0.3]])
    out = restate.delta2bbox(rois, deltas, (0., 0., 0., 0.), (1., 1., 1., 1.), max_shape=(32, 32))
    want = torch.tensor([[0.0000, 0.0000, 1.0000, 1.0000], [0.1409, 0.1409, 2.8591, 2.8591],
                         [0.0000, 0.3161, 4.1945, 0.6839], [5.0000, 5.0000, 5.0000, 5.0000]])
    assert torch.allclose(out, want, atol=1e-4)
    back = restate.bbox2delta(rois[:3], out[:3], (0., 0., 0., 0.), (1., 1., 1., 1.))
    assert torch.isfinite(back).all()


def test_levels_match_golden():
    z = np.load(os.path.join(GOLD, 'levels.npz'))
    rois = torch.from_numpy(z['rois'])
    lv = restate.map_roi_levels(rois, 4)
    assert np.array_equal(lv.numpy().astype(np.int8), z['levels'])
    out = np.zeros(rois.shape[0], dtype=np.int64)
    import ctypes
    restate.lib().map_roi_levels_f32(ctypes.c_void_p(rois.data_ptr()),
                                     out.ctypes.data_as(ctypes.c_void_p), rois.shape[0], 4,
                                     ctypes.c_float(56.0))
    assert np.array_equal(out.astype(np.int8), z['levels'])


@pytest.mark.parametrize('name', list(cases.CASES))
def test_masks_match_golden(name):
    z = np.load(os.path.join(GOLD, f'masks_{name}.npz'))
    c, x, pr, gts, shapes = cases.case_inputs(name)
    r = cases._rois(pr)
    masks = cases.graph_masks(restate.bbox_overlaps, restate.map_roi_levels(r, 4), r)
    assert {f'{b}_{i}' for b, i in masks} == {k.split('|')[0] for k in z.files}
    for (b, i), (idx, M, deg) in masks.items():
        assert np.array_equal(idx.numpy(), z[f'{b}_{i}|idx'])
        assert np.array_equal(np.packbits(M.numpy().astype(np.uint8), axis=1), z[f'{b}_{i}|bits'])
        assert np.array_equal(deg.numpy().astype(np.int32), z[f'{b}_{i}|deg'])


def _restate_outs(name, dt, which=('ext', 'head', 'train')):
    c = cases.CASES[name]
    head = restate.HTDRoIHead().to(dt)
    synth.fill_params_(head, c['scheme'], c['seed'])
    outs = {}
    if 'ext' in which:
        outs.update(cases.run_extractors(head, name, dt))
    if 'head' in which:
        outs.update(cases.run_head(head, name, dt))
    if 'train' in which:
        outs.update(cases.run_train(
            head, lambda h, *a: h.forward_train_sampled(*a),
            lambda h, *a: h.simple_test_scores(*a), name, dt))
    return outs


@pytest.mark.parametrize('name,tag,dt,tol', [('small', 'f64', torch.float64, 1e-10),
                                             ('small', 'f32', torch.float32, 2e-5),
                                             ('mid', 'f32', torch.float32, 2e-5),
                                             ('small_s', 'f64', torch.float64, 1e-10),
                                             ('c2', 'f64', torch.float64, 1e-10)])
def test_restatement_matches_golden(name, tag, dt, tol):
    fix = cases.load_fixture(os.path.join(GOLD, f'{name}_{tag}.npz'))
    outs = _restate_outs(name, dt)
    assert set(outs) == set(fix)
    cases.compare_to_fixture(outs, fix, tol)


@pytest.mark.skipif(not refshim.available(), reason='/root/reference not present')
def test_restatement_equals_reference_live():
    """Runs the reference's own unmodified modules next to the restatement (fp32, 'small')."""
    from oracle import ref_driver
    c = cases.CASES['small']
    ref = refshim.build_head()
    synth.fill_params_(ref, c['scheme'], c['seed'])
    assert sum(p.numel() for p in set(ref.parameters())) == 47189116
    a = {}
    a.update(cases.run_extractors(ref, 'small', torch.float32))
    a.update(cases.run_head(ref, 'small', torch.float32))
    b = _restate_outs('small', torch.float32, ('ext', 'head'))
    assert set(ref.state_dict()) == set(restate.HTDRoIHead().state_dict())
    for k in a:
        assert cases.rel_err(b[k], a[k]) <= 1e-6, k


# ---- test-time augmentation (htd_roi_head.py:388-440) ---------------------------------------------
@pytest.mark.parametrize('tag,dt,tol', [('f64', torch.float64, 1e-10), ('f32', torch.float32, 2e-5)])
def test_aug_test_restatement_matches_reference_golden(tag, dt, tol):
    """tests/golden/aug_small_*.npz hold the merged boxes / scores the reference's OWN aug_test
    hands to multiclass_nms (three views: plain, h-flip, v-flip, scale factor 0.5)."""
    c = cases.CASES['small']
    head = restate.HTDRoIHead().to(dt)
    synth.fill_params_(head, c['scheme'], c['seed'])
    outs = cases.run_aug(head, lambda h, *a: h.aug_test_merged(*a), 'small', dt)
    fix = cases.load_fixture(os.path.join(GOLD, f'aug_small_{tag}.npz'))
    assert set(outs) == set(fix)
    cases.compare_to_fixture(outs, fix, tol)


@pytest.mark.skipif(not refshim.available(), reason='/root/reference not present')
def test_aug_test_restatement_equals_reference_live():
    from oracle import ref_driver
    c = cases.CASES['small']
    ref = refshim.build_head(double=True)
    synth.fill_params_(ref, c['scheme'], c['seed'])
    a = cases.run_aug(ref, lambda h, *x: ref_driver.ref_aug_test(h, *x)[:2], 'small', torch.float64)
    head = restate.HTDRoIHead().double()
    synth.fill_params_(head, c['scheme'], c['seed'])
    b = cases.run_aug(head, lambda h, *x: h.aug_test_merged(*x), 'small', torch.float64)
    for k in a:
        assert cases.rel_err(b[k], a[k]) <= 1e-12, k


def test_bbox_mapping_helpers_round_trip():
    """Host logic of the product's aug_test: mapping to an augmented view and back is the identity,
    and flips are involutions (core/bbox/transforms.py:5-56)."""
    from htd_b200.core import bbox_flip, bbox_mapping, bbox_mapping_back, merge_aug_bboxes
    g = torch.Generator().manual_seed(3)
    b = torch.rand(17, 4, generator=g) * 200
    sf = np.array([0.5, 0.75, 0.5, 0.75], dtype=np.float32)
    for d in ('horizontal', 'vertical', 'diagonal'):
        assert torch.allclose(bbox_flip(bbox_flip(b, (300, 400), d), (300, 400), d), b, atol=1e-4)
        m = bbox_mapping(b, (300, 400), sf, True, d)
        assert torch.allclose(bbox_mapping_back(m, (300, 400), sf, True, d), b, atol=1e-4)
    wide = torch.rand(5, 8, generator=g) * 100
    assert bbox_flip(wide, (300, 400), 'horizontal').shape == wide.shape
    metas = [[dict(img_shape=(300, 400, 3), scale_factor=sf, flip=False)],
             [dict(img_shape=(300, 400, 3), scale_factor=sf, flip=True, flip_direction='vertical')]]
    views = [bbox_mapping(b, (300, 400), sf, m[0]['flip'], m[0].get('flip_direction', 'horizontal'))
             for m in metas]
    merged, sc = merge_aug_bboxes(views, [torch.ones(17, 3), 3 * torch.ones(17, 3)], metas)
    assert torch.allclose(merged, b, atol=1e-4) and torch.equal(sc, 2 * torch.ones(17, 3))


# ---- FPN neck -----------------------------------------------------------------------
def test_fpn_restatement_matches_reference_golden():
    """tests/golden/fpn_f64.npz: outputs and input / weight gradients of the reference's own FPN
    (necks/fpn.py) on odd-sized maps (non-integer nearest scales)."""
    fpn = cases.fpn_fill_(restate.FPN().double())
    outs = cases.run_fpn(fpn, torch.float64)
    fix = cases.load_fixture(os.path.join(GOLD, 'fpn_f64.npz'))
    assert set(outs) == set(fix)
    cases.compare_to_fixture(outs, fix, 1e-10)


@pytest.mark.skipif(not refshim.available(), reason='/root/reference not present')
def test_fpn_restatement_equals_reference_live():
    ns = refshim.load()
    ref = cases.fpn_fill_(ns.FPN([256, 512, 1024, 2048], 256, 5).double())
    a = cases.run_fpn(ref, torch.float64)
    b = cases.run_fpn(cases.fpn_fill_(restate.FPN().double()), torch.float64)
    assert set(ref.state_dict()) == set(restate.FPN().state_dict())
    for k in a:
        assert cases.rel_err(b[k], a[k]) <= 1e-12, k


def test_product_fpn_and_rpn_state_dicts_match_the_reference_layout():
    """Checkpoint compatibility (host logic): parameter names and shapes of htd_b200.necks.FPN /
    dense_heads.RPNHead are the reference's (necks/fpn.py:110-131, rpn_head.py:24-31)."""
    from htd_b200.dense_heads import RPNHead
    from htd_b200.necks import FPN
    mine = {k: tuple(v.shape) for k, v in FPN([256, 512, 1024, 2048], 256, 5).state_dict().items()}
    want = {k: tuple(v.shape) for k, v in restate.FPN().state_dict().items()}
    assert mine == want
    rpn = {k: tuple(v.shape) for k, v in RPNHead(256, 256).state_dict().items()}
    assert rpn == {'rpn_conv.weight': (256, 256, 3, 3), 'rpn_conv.bias': (256,),
                   'rpn_cls.weight': (3, 256, 1, 1), 'rpn_cls.bias': (3,),
                   'rpn_reg.weight': (12, 256, 1, 1), 'rpn_reg.bias': (12,)}
    if refshim.available():
        ns = refshim.load()
        ref = {k: tuple(v.shape) for k, v in ns.FPN([256, 512, 1024, 2048], 256, 5).state_dict().items()}
        assert mine == ref


# ---- RPN proposals ------------------------------------------------------------------
@pytest.mark.parametrize('name', list(cases.RPN_CASES))
def test_rpn_restatement_matches_reference_golden(name):
    """tests/golden/rpn_proposals.npz: output of the reference's own RPNHead._get_bboxes_single
    (source of dense_heads/rpn_head.py:77-168 executed unmodified) with its AnchorGenerator."""
    z = np.load(os.path.join(GOLD, 'rpn_proposals.npz'))
    cls, reg, shape, cfg = cases.rpn_inputs(name)
    anchors = restate.anchor_grid([c.shape[-2:] for c in cls])
    det = restate.rpn_proposals_single(cls, reg, anchors, shape, cfg['nms_pre'], cfg['nms_post'],
                                       cfg['nms_thr'], cfg['min_bbox_size'])
    assert np.array_equal(det.numpy(), z[name])


@pytest.mark.skipif(not refshim.available(), reason='/root/reference not present')
def test_rpn_restatement_equals_reference_live():
    from oracle import ref_driver
    cls, reg, shape, cfg = cases.rpn_inputs('minsize')
    want = ref_driver.ref_rpn_proposals(cls, reg, shape, cfg)
    anchors = restate.anchor_grid([c.shape[-2:] for c in cls])
    got = restate.rpn_proposals_single(cls, reg, anchors, shape, cfg['nms_pre'], cfg['nms_post'],
                                       cfg['nms_thr'], cfg['min_bbox_size'])
    assert torch.equal(got, want)


def test_product_anchor_generator_equals_the_restatement():
    """Host logic of htd_b200.dense_heads.AnchorGenerator (anchor_generator.py:142-272)."""
    from htd_b200.dense_heads import AnchorGenerator
    sizes = [(50, 76), (25, 38), (13, 19), (7, 10), (4, 5)]
    ag = AnchorGenerator(strides=[4, 8, 16, 32, 64], ratios=[0.5, 1.0, 2.0], scales=[8])
    assert ag.num_base_anchors == [3] * 5 and ag.num_levels == 5
    got = ag.grid_anchors(sizes, device='cpu')
    want = restate.anchor_grid(sizes)
    for a, b in zip(got, want):
        assert torch.equal(a, b)
    assert ag.grid_anchors(sizes, device='cpu')[0] is got[0]          # cached


# ---- assign + sample --------------------------------------------------------------
def _assign_fixture():
    z = np.load(os.path.join(GOLD, 'assign_sample.npz'))
    return {k: z[k] for k in z.files}


@pytest.mark.parametrize('name', list(cases.ASSIGN_CASES))
def test_assign_sample_restatement_matches_reference_golden(name):
    """oracle/restate.assign_sample_image == the reference's MaxIoUAssigner + RandomSampler
    (tests/golden/assign_sample.npz, made by oracle/gen_golden.py from the reference classes):
    every index, flag and count bit for bit, boxes exactly."""
    fix = _assign_fixture()
    out, res = cases.run_assign_case(name, restate.assign_sample_image)
    for k, v in out.items():
        assert np.array_equal(v.numpy(), fix[f'{name}|{k}']), (name, k)
    for b, r in enumerate(res):
        assert np.array_equal(r.gt_inds.numpy().astype(np.int32), fix[f'{name}|gt_inds{b}'])
        assert np.array_equal(r.max_overlaps.numpy(), fix[f'{name}|max_overlaps{b}'])


def test_assigner_known_answers_of_the_reference_tests():
    """/root/reference/tests/test_assigner.py:14-36 (gt_inds [1,0,2,0]) and :66-83 (no gt ->
    all background), through the restatement and in the committed fixture."""
    fix = _assign_fixture()
    assert fix['kat|gt_inds0'].tolist() == [1, 0, 2, 0]
    assert fix['kat|gt_inds1'].tolist() == [0, 0, 0, 0]
    d = cases.assign_case_inputs('kat')
    gi, mo, lab = restate.max_iou_assign(d['props'][0], d['gt_boxes'][0], d['gt_labels'][0], 0.5, 0.5)
    assert gi.tolist() == [1, 0, 2, 0] and lab.tolist() == [2, -1, 3, -1]


@pytest.mark.skipif(not refshim.available(), reason='/root/reference not present')
def test_assign_sample_restatement_equals_reference_live():
    from oracle import ref_driver
    for name in cases.ASSIGN_CASES:
        a, _ = cases.run_assign_case(name, ref_driver.ref_assign_sample_image)
        b, _ = cases.run_assign_case(name, restate.assign_sample_image)
        for k in a:
            assert torch.equal(a[k], b[k]), (name, k)


def test_key_choice_has_the_distribution_of_randperm_prefix():
    """choose_by_keys with i.i.d. uniform keys picks every member with probability num/n."""
    g = torch.Generator().manual_seed(0)
    gallery = torch.arange(3, 43)
    hits = torch.zeros(50)
    for _ in range(2000):
        keys = torch.rand(50, generator=g)
        hits[restate.choose_by_keys(gallery, 10, keys)] += 1
    assert hits[:3].sum() == 0 and hits[43:].sum() == 0
    assert (hits[3:43] / 2000 - 0.25).abs().max() < 0.05


def _restate_train_assigned(dtype=torch.float32):
    c = cases.TRAIN_ASSIGNED
    own = restate.HTDRoIHead()
    synth.fill_params_(own, c['scheme'], c['wseed'])
    own = own.to(dtype)
    return cases.run_train_assigned(lambda h, *a: h.forward_train_assigned(*a), own, dtype)


def test_restated_forward_train_with_assigner_matches_reference_golden():
    """restate.forward_train_assigned == the reference's unmodified HTDRoIHead.forward_train with
    its own MaxIoUAssigner / RandomSampler (keys instead of randperm): sampled indices exactly,
    losses and gradients to 1e-6 (tests/golden/train_assigned_f32.npz)."""
    fix = cases.load_fixture(os.path.join(GOLD, 'train_assigned_f32.npz'))
    outs = _restate_train_assigned()
    assert set(outs) == set(fix)
    cases.compare_to_fixture(outs, fix, 1e-6)


@pytest.mark.skipif(not refshim.available(), reason='/root/reference not present')
def test_restated_forward_train_with_assigner_equals_reference_live():
    from oracle import ref_driver
    c = cases.TRAIN_ASSIGNED
    ref = refshim.build_head()
    synth.fill_params_(ref, c['scheme'], c['wseed'])
    a = cases.run_train_assigned(ref_driver.ref_forward_train_assigned, ref)
    b = _restate_train_assigned()
    assert set(a) == set(b)
    for k in a:
        if a[k].is_floating_point():
            assert cases.rel_err(b[k], a[k]) <= 1e-6, k
        else:
            assert torch.equal(a[k], b[k]), k


# ---- multi-class NMS ----------------------------------------------------------------
@pytest.mark.parametrize('name', list(cases.NMS_CASES))
def test_nms_restatement_matches_reference_golden(name):
    """restate.multiclass_nms == the reference's multiclass_nms (tests/golden/nms.npz): the same
    detections in the same order, boxes and scores bit for bit."""
    z = np.load(os.path.join(GOLD, 'nms.npz'))
    boxes, scores, c = cases.nms_case_inputs(name)
    dets, labels = restate.multiclass_nms(boxes, scores, c['score_thr'], c['iou_thr'], c['max_num'])
    assert np.array_equal(dets.numpy(), z[f'{name}|dets']) and \
        np.array_equal(labels.numpy(), z[f'{name}|labels'])


@pytest.mark.parametrize('name', list(cases.NMS_CASES))
def test_soft_nms_restatement_matches_reference_golden(name):
    """restate.multiclass_nms(nms_type='soft_nms') == the reference's multiclass_nms called with
    the R-101 configs' nms_cfg (tests/golden/nms_soft.npz; mmcv's op = oracle/soft_nms_ref.c)."""
    z = np.load(os.path.join(GOLD, 'nms_soft.npz'))
    boxes, scores, c = cases.nms_case_inputs(name)
    dets, labels = restate.multiclass_nms(boxes, scores, c['score_thr'], c['iou_thr'], c['max_num'],
                                          nms_type='soft_nms', min_score=c['score_thr'])
    assert np.array_equal(dets.numpy(), z[f'{name}|dets']) and \
        np.array_equal(labels.numpy(), z[f'{name}|labels'])


def test_soft_nms_parallel_formulation_equals_the_literal_loop():
    """The per-pass array formulation the CUDA kernel uses (first-occurrence argmax, decay of all
    later boxes at once, removals as ONE unstable compaction) selects exactly what the literal
    sequential loop selects - tied scores included - and stops early consistently."""
    g = torch.Generator().manual_seed(0)
    for trial in range(24):
        n = int(torch.randint(1, 400, (1,), generator=g))
        ctr = torch.rand(n, 2, generator=g) * 200
        wh = torch.rand(n, 2, generator=g) * 80 + 5
        b = torch.cat([ctr - wh / 2, ctr + wh / 2], 1)
        s = torch.rand(n, generator=g)
        if trial % 2:
            s = torch.floor(s * 16) / 16 + 0.05           # many exact ties
        for method in ('linear', 'naive'):
            d1, k1 = restate.soft_nms(b, s, 0.5, 0.5, 0.1, method)
            d2, k2 = restate.soft_nms_vectorised(b, s, 0.5, 0.5, 0.1, method)
            assert torch.equal(k1, k2) and torch.equal(d1, d2.float()), (trial, method)
            _, k3 = restate.soft_nms_vectorised(b, s, 0.5, 0.5, 0.1, method, max_out=7)
            assert torch.equal(k1[:7], k3)
    # hand-checked: two identical boxes, linear decay 1 - IoU = 0 -> the second is dropped
    b = torch.tensor([[0., 0., 10., 10.], [0., 0., 10., 10.], [20., 20., 30., 30.]])
    d, k = restate.soft_nms(b, torch.tensor([0.9, 0.8, 0.7]), 0.5, 0.5, 0.05)
    assert k.tolist() == [0, 2] and d[:, 4].tolist() == pytest.approx([0.9, 0.7])


def test_nms_restatement_equals_torchvision_nms():
    """The greedy NMS of the restatement against the library implementation of the same published
    algorithm (torchvision.ops.nms, CPU) on clustered boxes."""
    from torchvision.ops import nms
    boxes, scores, c = cases.nms_case_inputs('htd')
    s = scores[:, 3] + torch.arange(scores.size(0)) * 1e-7
    for thr in (0.3, 0.5, 0.7):
        assert torch.equal(restate.nms_greedy(boxes, s, thr), nms(boxes, s, thr))
