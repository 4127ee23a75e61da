"""ORACLE (test infrastructure).  Drives the reference's own HTDRoIHead (via refshim) through
the same synthetic-sampling protocol as ``oracle/restate.py``: the reference's
``glbctx_head``, ``_bbox_forward_train`` (htd_roi_head.py:203-215), ``refine_bboxes`` and
``_bbox_forward`` are called unmodified; only the random assign+sample steps
(htd_roi_head.py:254-264, :300-310) are replaced by positives-first synthetic SamplingResults.
Only usable where /root/reference exists.
"""
import torch

from . import refshim


def make_ref_sampling(ns, bboxes, num_pos, gt):
    n = min(num_pos, bboxes.size(0))
    r = ns.SamplingResult.__new__(ns.SamplingResult)
    r.pos_bboxes = bboxes[:n]
    r.neg_bboxes = bboxes[n:]
    r.pos_gt_bboxes = gt['pos_gt_bboxes'][:n].to(bboxes.dtype)
    r.pos_gt_labels = gt['pos_gt_labels'][:n]
    r.pos_is_gt = torch.zeros(n, dtype=torch.uint8)
    r.pos_inds = torch.arange(n)
    r.neg_inds = torch.arange(n, bboxes.size(0))
    r.num_gts = n
    r.pos_assigned_gt_inds = torch.arange(n)
    return r


def ref_forward_train_sampled(head, x, proposals, gts, img_shapes, num_pos=128,
                              return_intermediates=False):
    ns = refshim.load()
    img_metas = [dict(img_shape=s) for s in img_shapes]
    losses = {}
    inter = {}
    mc_pred, g = head.glbctx_head(x)
    losses['loss_global'] = head.glbctx_head.loss(mc_pred, [gt['gt_labels_unique'] for gt in gts])
    samp = [make_ref_sampling(ns, p, num_pos, gt) for p, gt in zip(proposals, gts)]
    res = head._bbox_forward_train(0, x, samp, None, None, head.train_cfg[0], img_metas, g)
    for k, v in res['loss_bbox'].items():
        losses[f's0.{k}'] = v * head.stage_loss_weights[0] if 'loss' in k else v
    inter['s0.cls_score'], inter['s0.bbox_pred'] = res['cls_score'], res['bbox_pred']
    roi_labels = res['bbox_targets'][0]
    with torch.no_grad():
        roi_labels = torch.where(roi_labels == head.bbox_head[0].num_classes,
                                 res['cls_score'][:, :-1].argmax(1), roi_labels)
        refined = head.bbox_head[0].refine_bboxes(res['rois'], roi_labels, res['bbox_pred'],
                                                  [r.pos_is_gt for r in samp], img_metas)
    inter['refined'] = torch.cat(refined)
    samp = [make_ref_sampling(ns, p, num_pos, gt) for p, gt in zip(refined, gts)]
    res = head._bbox_forward_train(1, x, samp, None, None, head.train_cfg[1], img_metas, g)
    for k, v in res['loss_bbox'].items():
        losses[f's1.{k}'] = v * head.stage_loss_weights[1] if 'loss' in k else v
    inter['s1.cls_score'], inter['s1.bbox_pred'] = res['cls_score'], res['bbox_pred']
    return (losses, inter) if return_intermediates else losses


def ref_simple_test_scores(head, x, proposals, img_shapes):
    """htd_roi_head.py:319-366 (body of simple_test before get_bboxes/NMS)."""
    ns = refshim.load()
    img_metas = [dict(img_shape=s) for s in img_shapes]
    rois = ns.bbox2roi(proposals)
    _, g = head.glbctx_head(x)
    r0 = head._bbox_forward(0, x, rois, g)
    n = tuple(len(p) for p in proposals)
    cls = r0['cls_score'].split(n, 0)
    bp = r0['bbox_pred'].split(n, 0)
    rs = rois.split(n, 0)
    label = [s[:, :-1].argmax(dim=1) for s in cls]
    new_rois = torch.cat([head.bbox_head[0].regress_by_class(rs[j], label[j], bp[j], img_metas[j])
                          for j in range(len(proposals))])
    r1 = head._bbox_forward(1, x, new_rois, g)
    return new_rois, (r0['cls_score'] + r1['cls_score']) / 2.0, r1['bbox_pred']


def ref_assign_sample_image(bboxes, gt_bboxes, gt_labels, keys, cfg, valid=None):
    """One image through the REFERENCE's MaxIoUAssigner.assign + RandomSampler.sample (the calls of
    htd_roi_head.py:254-264).  The only substitution: ``random_choice`` (randperm of the global
    RNG) draws by the given keys (restate.choose_by_keys), so that the result is reproducible."""
    from types import SimpleNamespace
    from . import restate
    ns = refshim.load()
    a = {k: v for k, v in cfg['assigner'].items() if k != 'type'}
    s = {k: v for k, v in cfg['sampler'].items() if k != 'type'}
    assigner, sampler = ns.MaxIoUAssigner(**a), ns.RandomSampler(**s)
    keep = torch.arange(bboxes.size(0)) if valid is None else torch.nonzero(valid).squeeze(1)
    boxes = bboxes[keep]
    g = gt_bboxes.size(0)
    added = s.get('add_gt_as_proposals', True) and g > 0
    src = torch.cat([torch.arange(g), keep + g]) if added else keep
    ckeys = keys[src]
    sampler.random_choice = lambda gallery, num: restate.choose_by_keys(gallery, num, ckeys)
    ar = assigner.assign(boxes, gt_bboxes, None, gt_labels)
    sr = sampler.sample(ar, boxes, gt_bboxes, gt_labels)
    return SimpleNamespace(
        pos_inds=sr.pos_inds, neg_inds=sr.neg_inds, pos_bboxes=sr.pos_bboxes,
        neg_bboxes=sr.neg_bboxes, pos_is_gt=sr.pos_is_gt,
        pos_assigned_gt_inds=sr.pos_assigned_gt_inds, pos_gt_bboxes=sr.pos_gt_bboxes,
        pos_gt_labels=sr.pos_gt_labels if sr.pos_gt_labels is not None
        else torch.zeros(0, dtype=torch.long),
        bboxes=sr.bboxes, gt_inds=ar.gt_inds, max_overlaps=ar.max_overlaps,
        cand=torch.cat([src[sr.pos_inds], src[sr.neg_inds]]),
        npos_cand=int((ar.gt_inds > 0).sum()), nneg_cand=int((ar.gt_inds == 0).sum()))


def ref_forward_train_assigned(head, x, proposals, gt_bboxes, gt_labels, keys, cfgs, img_shapes, G):
    """The reference's own ``HTDRoIHead.forward_train`` (htd_roi_head.py:217-317), unmodified,
    assigners and samplers included.  Only ``RandomSampler.random_choice`` is redirected to the
    key rule (restate.choose_by_keys) through a wrapper around ``sampler.sample`` that knows which
    (stage, image) is being sampled; ``keys`` as in restate.forward_train_assigned."""
    from . import restate
    img_metas = [dict(img_shape=s) for s in img_shapes]
    rec = {0: [], 1: []}
    keep0 = {}
    refined = []
    saved = []
    for st in (0, 1):
        sampler, assigner = head.bbox_sampler[st], head.bbox_assigner[st]
        a, s = cfgs[st]['assigner'], cfgs[st]['sampler']
        saved.append((sampler, sampler.num, sampler.pos_fraction))
        sampler.num, sampler.pos_fraction = s['num'], s['pos_fraction']
        assert (assigner.pos_iou_thr, assigner.neg_iou_thr, assigner.min_pos_iou,
                assigner.match_low_quality) == (a['pos_iou_thr'], a['neg_iou_thr'],
                                                a['min_pos_iou'], a['match_low_quality'])

        def sample(assign_result, bboxes, gtb, gtl=None, _st=st, _sampler=sampler, **kw):
            j = len(rec[_st])
            k, ng = keys[_st][j], gtb.size(0)
            if _st == 0:
                pk = k[G:]
            else:
                pk = k[G:G + keep0[j].numel()][keep0[j]]
                refined.append(bboxes)
            ck = torch.cat([k[:ng], pk]) if (_sampler.add_gt_as_proposals and ng > 0) else pk
            _sampler.random_choice = lambda gallery, num: restate.choose_by_keys(gallery, num, ck)
            sr = type(_sampler).sample(_sampler, assign_result, bboxes, gtb, gtl, **kw)
            if _st == 0:
                keep0[j] = torch.cat([1 - sr.pos_is_gt,
                                      sr.pos_is_gt.new_ones(sr.neg_bboxes.size(0))]).bool()
            rec[_st].append(sr)
            return sr
        sampler.sample = sample
    try:
        losses = head.forward_train(x, img_metas, proposals, gt_bboxes, gt_labels)
    finally:
        for sampler, num, frac in saved:
            sampler.num, sampler.pos_fraction = num, frac
            del sampler.sample
            if 'random_choice' in sampler.__dict__:
                del sampler.random_choice
    return losses, dict(samp0=rec[0], samp1=rec[1], refined=refined)


def ref_aug_test(head, features, proposals, img_metas):
    """The reference's own HTDRoIHead.aug_test (htd_roi_head.py:388-440), unmodified; its
    multiclass_nms call is wrapped only to capture the merged boxes / scores it receives.
    Returns (merged_bboxes, merged_scores, det_bboxes, det_labels)."""
    import sys
    refshim.load()
    rh = sys.modules['mmdet.models.roi_heads.htd_roi_head']
    seen = {}
    real = rh.multiclass_nms

    def capture(bboxes, scores, *a, **k):
        seen['bboxes'], seen['scores'] = bboxes, scores
        det = real(bboxes, scores, *a, **k)
        seen['det'] = det
        return det

    rh.multiclass_nms = capture
    try:
        head.aug_test(features, [proposals], img_metas)
    finally:
        rh.multiclass_nms = real
    return seen['bboxes'], seen['scores'], seen['det'][0], seen['det'][1]


def ref_rpn_function():
    """The reference's RPNHead._get_bboxes_single (dense_heads/rpn_head.py:77-168), taken from
    its source file and executed unmodified.  Importing the class would pull in the whole
    anchor-head / loss stack of mmdet; the method only needs torch and mmcv.ops.batched_nms
    (mmcv is un-vendored: refshim's torchvision-based stand-in, as for the RoI head's NMS)."""
    import ast
    path = refshim.REF + '/mmdet/models/dense_heads/rpn_head.py'
    src = open(path).read()
    fn = [n for n in ast.walk(ast.parse(src))
          if isinstance(n, ast.FunctionDef) and n.name == '_get_bboxes_single'][0]
    code = ast.get_source_segment(src, fn)
    import textwrap
    g = {'torch': torch, 'batched_nms': refshim._batched_nms}
    exec(compile(textwrap.dedent(code), path, 'exec'), g)
    return g['_get_bboxes_single']


def ref_rpn_proposals(cls_scores, bbox_preds, img_shape, cfg):
    """One image through the reference's _get_bboxes_single with the reference's own
    AnchorGenerator and DeltaXYWHBBoxCoder (configs/htd/htd_resnet50_1x.py:22-37)."""
    import importlib
    from types import SimpleNamespace
    ns = refshim.load()
    refshim._mod('mmdet.core.anchor', path=refshim.REF + '/mmdet/core/anchor')
    ag = importlib.import_module('mmdet.core.anchor.anchor_generator').AnchorGenerator(
        strides=[4, 8, 16, 32, 64], ratios=[0.5, 1.0, 2.0], scales=[8])
    anchors = ag.grid_anchors([c.shape[-2:] for c in cls_scores], device='cpu')
    anchors = [a.to(cls_scores[0].dtype) for a in anchors]
    self = SimpleNamespace(use_sigmoid_cls=True, test_cfg=cfg, bbox_coder=ns.DeltaXYWHBBoxCoder(
        target_means=[.0, .0, .0, .0], target_stds=[1.0, 1.0, 1.0, 1.0]))
    return ref_rpn_function()(self, cls_scores, bbox_preds, anchors, img_shape, 1.0,
                              refshim.AttrDict.wrap(dict(cfg)))

