"""GPU parity tests (through the C-ABI) of the multi-class NMS kernels (SURVEY §8 f3,
csrc/nms.cu): the detections and their order must equal, bit for bit, the fixture written by the
reference's own multiclass_nms (tests/golden/nms.npz) and the CPU restatement."""
import os

import numpy as np
import pytest
import torch

from htd_b200 import core, ops
from oracle import cases, restate

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.mark.parametrize('name', list(cases.NMS_CASES))
def test_multiclass_nms_matches_reference_golden(name):
    z = np.load(os.path.join(GOLD, 'nms.npz'))
    boxes, scores, c = cases.nms_case_inputs(name)
    dets, labels = core.multiclass_nms(boxes.cuda(), scores.cuda(), c['score_thr'],
                                       dict(type='nms', iou_threshold=c['iou_thr']), c['max_num'])
    assert np.array_equal(dets.cpu().numpy(), z[f'{name}|dets']), name
    assert np.array_equal(labels.cpu().numpy(), z[f'{name}|labels']), name


def test_multiclass_nms_matches_restatement_random_sizes():
    g = torch.Generator().manual_seed(7)
    # (2048, 5), (3000, 3) and (4096, 2): few classes with many candidates each - the bit-matrix path
    for K, C, per_class in ((1, 1, False), (37, 3, True), (1000, 80, False), (2048, 5, False),
                            (3000, 3, True), (4096, 2, False), (513, 8, False)):
        cases.NMS_CASES['_t'] = dict(K=K, C=C, per_class=per_class, score_thr=0.02, iou_thr=0.45,
                                     max_num=150, seed=int(torch.randint(0, 1000, (1,), generator=g)),
                                     dup=0.5, temp=2.0)
        try:
            boxes, scores, c = cases.nms_case_inputs('_t')
        finally:
            del cases.NMS_CASES['_t']
        want_d, want_l = restate.multiclass_nms(boxes, scores, c['score_thr'], c['iou_thr'], c['max_num'])
        det, lab, cnt = ops.multiclass_nms(boxes.cuda(), scores.cuda(), c['score_thr'], c['iou_thr'],
                                           c['max_num'])
        n = int(cnt)
        assert n == want_d.size(0), (K, C, n, want_d.size(0))
        assert torch.equal(det[:n].cpu(), want_d) and torch.equal(lab[:n].cpu(), want_l), (K, C)


@pytest.mark.parametrize('name', list(cases.NMS_CASES))
def test_multiclass_soft_nms_matches_reference_golden(name):
    """nms=dict(type='soft_nms', iou_thr=..., min_score=...) (configs/htd/htd_resnet101_2x.py:298):
    boxes, decayed scores, labels and the ORDER (tied scores included) equal the fixture written by
    the reference's multiclass_nms bit for bit."""
    z = np.load(os.path.join(GOLD, 'nms_soft.npz'))
    boxes, scores, c = cases.nms_case_inputs(name)
    dets, labels = core.multiclass_nms(boxes.cuda(), scores.cuda(), c['score_thr'],
                                       dict(type='soft_nms', iou_thr=c['iou_thr'],
                                            min_score=c['score_thr']), c['max_num'])
    assert np.array_equal(dets.cpu().numpy(), z[f'{name}|dets']), name
    assert np.array_equal(labels.cpu().numpy(), z[f'{name}|labels']), name


def test_multiclass_soft_nms_matches_restatement_random_sizes():
    g = torch.Generator().manual_seed(9)
    for K, C, per_class, max_num, quant in ((1, 1, False, 10, None), (37, 3, True, -1, 8),
                                            (1000, 80, False, 100, None), (2048, 5, False, 300, 16),
                                            (600, 4, False, -1, None)):
        cases.NMS_CASES['_t'] = dict(K=K, C=C, per_class=per_class, score_thr=0.05, iou_thr=0.5,
                                     max_num=max_num, dup=0.6, temp=2.0,
                                     seed=int(torch.randint(0, 1000, (1,), generator=g)))
        if quant:
            cases.NMS_CASES['_t']['quant'] = quant
        try:
            boxes, scores, c = cases.nms_case_inputs('_t')
        finally:
            del cases.NMS_CASES['_t']
        for method in ('linear', 'naive'):
            want_d, want_l = restate.multiclass_nms(boxes, scores, 0.05, 0.5, max_num,
                                                    nms_type='soft_nms', min_score=0.05, method=method)
            d, l = core.multiclass_nms(boxes.cuda(), scores.cuda(), 0.05,
                                       dict(type='soft_nms', iou_threshold=0.5, min_score=0.05,
                                            method=method), max_num)
            assert d.shape == want_d.shape, (K, C, method, d.shape, want_d.shape)
            assert torch.equal(d.cpu(), want_d) and torch.equal(l.cpu(), want_l), (K, C, method)


def test_simple_test_with_the_r101_soft_nms_test_cfg():
    """BASELINE config 4 (R-101-DCN) post-processing: simple_test with the real test_cfg of
    configs/htd/htd_resnet101_2x.py:292-300 (soft_nms) == the restatement on the head's own scores."""
    import htd_b200
    from htd_b200 import synth
    head = htd_b200.build_htd_roi_head().cuda()
    synth.fill_params_(head, 'n005', 2)
    head.eval()
    head.test_cfg = htd_b200.core.as_cfg(dict(score_thr=0.05, max_per_img=100,
                                              nms=dict(type='soft_nms', iou_thr=0.5, min_score=0.05)))
    H, W = 256, 320
    x = [t.cuda() for t in synth.make_pyramid(1, H, W)]
    props = [p.cuda() for p in synth.make_proposals(1, 300, H, W, min_scale=8, max_scale=300)]
    metas = [dict(img_shape=(H, W, 3), scale_factor=1.0)]
    with torch.no_grad():
        res = head.simple_test(x, props, metas)
        rois, cls_score, bbox_pred = head.simple_test_scores(x, props, metas)
        boxes, scores = head.bbox_head[-1].get_bboxes(rois, cls_score, bbox_pred, (H, W, 3), 1.0)
    wd, wl = restate.multiclass_nms(boxes.cpu(), scores.cpu(), 0.05, 0.5, 100, nms_type='soft_nms',
                                    min_score=0.05)
    assert sum(a.shape[0] for a in res[0]) == wd.shape[0] > 0
    for c in range(80):
        assert np.array_equal(res[0][c], wd[wl == c].numpy()), c


def test_simple_test_returns_reference_format_through_the_nms_kernel():
    """HTDRoIHead.simple_test (htd_roi_head.py:319-386): per image a list of num_classes arrays
    [n_c,5], at most max_per_img detections, scores descending within the NMS output."""
    import htd_b200
    from htd_b200 import synth
    head = htd_b200.build_htd_roi_head().cuda()
    head.init_weights()
    head.eval()
    H, W = 256, 320
    x = [t.cuda() for t in synth.make_pyramid(2, H, W)]
    props = [p.cuda() for p in synth.make_proposals(2, 200, H, W, min_scale=8, max_scale=300)]
    metas = [dict(img_shape=(H, W, 3), scale_factor=1.0) for _ in props]
    with torch.no_grad():
        res = head.simple_test(x, props, metas)
    assert len(res) == 2
    for per_img in res:
        assert len(per_img) == 80 and all(a.shape[1] == 5 for a in per_img)
        assert sum(a.shape[0] for a in per_img) <= 100


def test_multiclass_nms_edge_cases():
    dev = 'cuda'
    det, lab, cnt = ops.multiclass_nms(torch.zeros(0, 4, device=dev), torch.zeros(0, 81, device=dev),
                                       0.05, 0.5, 100)
    assert int(cnt) == 0 and det.shape == (100, 5)
    # identical boxes, identical scores: the first (lowest index) survives per class
    boxes = torch.tensor([[10., 10., 50., 50.]] * 4, device=dev)
    scores = torch.tensor([[0.4, 0.3, 0.3]] * 4, device=dev)
    d, l = core.multiclass_nms(boxes, scores, 0.05, dict(type='nms', iou_threshold=0.5), 10)
    assert d.shape == (2, 5) and l.tolist() == [0, 1] and d[:, 4].tolist() == pytest.approx([0.4, 0.3])
    # zero-area boxes: IoU is 0/0 = nan, never above the threshold - nothing is suppressed
    z = torch.tensor([[5., 5., 5., 5.]] * 3, device=dev)
    s = torch.tensor([[0.9, 0.1], [0.8, 0.2], [0.7, 0.3]], device=dev)
    d, l = core.multiclass_nms(z, s, 0.05, dict(type='nms', iou_threshold=0.5), 10)
    wd, wl = restate.multiclass_nms(z.cpu(), s.cpu(), 0.05, 0.5, 10)
    assert torch.equal(d.cpu(), wd) and torch.equal(l.cpu(), wl) and d.shape[0] == 3
