"""GPU parity tests of the RPN proposal path (producer side):
htd_b200.dense_heads.RPNHead.get_bboxes (torch.sort ranking, htd_bbox_decode, htd_multiclass_nms with
the level as the class) against the CPU oracle (oracle/restate.py rpn_proposals_single) and the
fixture written by the reference's own RPNHead._get_bboxes_single."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _head():
    from htd_b200.dense_heads import RPNHead
    return RPNHead(256, 256).cuda()


@pytest.mark.parametrize('name', ['train', 'test', 'minsize'])
def test_rpn_proposals_vs_oracle_and_reference_fixture(name):
    """Same proposals in the same order (see _same_proposals for the comparison)."""
    from oracle import cases, restate
    cls, reg, shape, cfg = cases.rpn_inputs(name)
    anchors = restate.anchor_grid([c.shape[-2:] for c in cls])
    want = restate.rpn_proposals_single(cls, reg, anchors, shape, cfg['nms_pre'], cfg['nms_post'],
                                        cfg['nms_thr'], cfg['min_bbox_size'])
    head = _head()
    got = head.get_bboxes([c[None].cuda() for c in cls], [r[None].cuda() for r in reg],
                          [dict(img_shape=shape, scale_factor=1.0)], cfg)
    assert len(got) == 1
    g = got[0].cpu()
    assert (g[:-1, 4] >= g[1:, 4]).all()
    z = np.load(os.path.join(GOLD, 'rpn_proposals.npz'))
    for ref in (want, torch.from_numpy(z[name])):
        _same_proposals(g, ref)


def _same_proposals(got, want):
    """Scores are (nearly all) distinct by construction, so a proposal is identified by its score
    (host and device sigmoid may differ in the last bit: match within 3e-7).  The decoded boxes
    differ in the last bits too (exp), which can flip a suppression whose IoU sits on the
    threshold: at most 1 % of the proposals may differ; the others agree to 1e-3 px and keep
    their order."""
    gsc, wsc = got[:, 4].double(), want[:, 4].double()
    # nearest got score for every want score (both descending)
    pos = torch.searchsorted(-gsc.contiguous(), -wsc.contiguous()).clamp(max=len(gsc) - 1)
    cand = torch.stack([(pos - 1).clamp(min=0), pos])
    dist = (gsc[cand] - wsc[None]).abs()
    best = cand.gather(0, dist.argmin(0, keepdim=True))[0]
    ok = dist.min(0).values <= 3e-7
    # drop scores shared by several proposals (the levels have a few logits in common)
    uniq_w = torch.ones_like(ok)
    uniq_w[1:] &= (wsc[:-1] - wsc[1:]) > 1e-6
    uniq_w[:-1] &= (wsc[:-1] - wsc[1:]) > 1e-6
    ok &= uniq_w
    slack = max(2, int(0.01 * want.shape[0]))
    assert abs(got.shape[0] - want.shape[0]) <= slack, (got.shape, want.shape)
    assert int(ok.sum()) >= want.shape[0] - slack - int((~uniq_w).sum()), (int(ok.sum()), want.shape[0])
    gi, wi = best[ok], torch.nonzero(ok).squeeze(1)
    assert torch.allclose(got[gi, :4], want[wi, :4], atol=1e-3, rtol=0)
    assert (gi[1:] > gi[:-1]).all()


def test_rpn_head_end_to_end_two_images_from_fpn_outputs():
    """FPN (channels-last bf16) -> RPN head -> proposals for two images -> HTDRoIHead.simple_test."""
    import htd_b200
    from htd_b200 import synth
    from htd_b200.necks import FPN
    from oracle import cases
    fpn = cases.fpn_fill_(FPN([256, 512, 1024, 2048], 256, 5)).cuda().bfloat16()
    rpn = _head().bfloat16().to(memory_format=torch.channels_last)
    head = htd_b200.build_htd_roi_head()
    synth.fill_params_(head, 'n005', 0)
    head = head.cuda().bfloat16()
    head.compute_dtype = torch.bfloat16
    xs = cases.fpn_inputs(torch.float32, 'cuda')
    H, W = cases.FPN_SIZES[0]
    metas = [dict(img_shape=(4 * H, 4 * W, 3), scale_factor=1.0)] * 2
    cfg = dict(nms_across_levels=False, nms_pre=1000, nms_post=300, max_num=300, nms_thr=0.7,
               min_bbox_size=0)
    with torch.no_grad():
        feats = fpn(xs)
        props = rpn.get_bboxes(*rpn(feats), metas, cfg)
        assert len(props) == 2 and all(p.shape[1] == 5 and 0 < p.shape[0] <= 300 for p in props)
        for p in props:
            assert (p[:, 0] >= 0).all() and (p[:, 2] <= 4 * W).all() and (p[:, 3] <= 4 * H).all()
        res = head.simple_test(feats, props, metas)
    assert len(res) == 2 and len(res[0]) == head.bbox_head[-1].num_classes


@pytest.mark.parametrize('n,k', [(201600, 2000), (50400, 1000), (3150, 2000), (819, 819), (5000, 4096),
                                 (7, 3), (1, 1)])
def test_topk_sorted_equals_a_stable_descending_sort(n, k):
    """htd_topk_sorted against torch.sort(stable=True, descending=True): values and positions
    bit-exact, several rows (strided), negative / positive / repeated keys."""
    from htd_b200 import ops
    g = torch.Generator().manual_seed(n + k)
    keys = torch.randn(3, n + 5, generator=g).cuda()[:, :n]            # row stride n + 5
    keys[1] = (keys[1] * 4).round() / 4                                 # many exact ties
    if n > 10:
        keys[2, ::3] = -keys[2, ::3].abs() * 1e-20                      # denormal-range negatives
    vals, idx = ops.topk_sorted(keys, k)
    sv, si = torch.sort(keys, dim=1, descending=True, stable=True)
    assert torch.equal(vals, sv[:, :k])
    assert torch.equal(idx, si[:, :k])


def test_topk_sorted_massive_ties_take_the_index_ordered_path():
    """More threshold-valued keys than the collection buffer holds (saturated logits): the first
    ones in position order are taken."""
    from htd_b200 import ops
    n, k = 60000, 1500
    keys = torch.full((2, n), 3.5, device='cuda')
    keys[0, 100:600] = 9.0                      # 500 above the tie value
    keys[1, ::7] = -1.0
    vals, idx = ops.topk_sorted(keys, k)
    sv, si = torch.sort(keys, dim=1, descending=True, stable=True)
    assert torch.equal(vals, sv[:, :k]) and torch.equal(idx, si[:, :k])


def test_rpn_proposals_without_ranking_and_with_one_level_empty_after_the_size_filter():
    """nms_pre larger than every level (the reference then keeps the anchor order: no ranking) and a
    min_bbox_size that removes most boxes; compared with the oracle like the other cases."""
    from oracle import cases, restate
    cls, reg, shape, _ = cases.rpn_inputs('minsize')
    cfg = dict(nms_across_levels=False, nms_pre=4000, nms_post=50, max_num=50, nms_thr=0.6, min_bbox_size=40)
    anchors = restate.anchor_grid([c.shape[-2:] for c in cls])
    want = restate.rpn_proposals_single(cls, reg, anchors, shape, cfg['nms_pre'], cfg['nms_post'],
                                        cfg['nms_thr'], cfg['min_bbox_size'])
    got = _head().get_bboxes([c[None].cuda() for c in cls], [r[None].cuda() for r in reg],
                             [dict(img_shape=shape, scale_factor=1.0)], cfg)[0].cpu()
    _same_proposals(got, want)
