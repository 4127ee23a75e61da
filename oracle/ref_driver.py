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
