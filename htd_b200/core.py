"""Host-side helpers adjacent to the RoI-head path: box/RoI
conversion, the delta coder, IoU assignment, random sampling, the two losses and the NMS wrapper.

These are small tensor programs that the reference also runs in plain PyTorch
(``mmdet/core/bbox/*``, ``mmdet/models/losses/*``); they are kept in PyTorch here (device
plumbing), with the reference's names, arguments and conventions so the head modules read like
the reference's.  The hot arithmetic (RoIAlign / BA / PGraph) is NOT here - see ops.py/pgraph.py.
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .registry import BBOX_ASSIGNERS, BBOX_CODERS, BBOX_SAMPLERS, LOSSES


class AttrDict(dict):
    """Minimal stand-in for mmcv.Config nodes: ``cfg.assigner`` == ``cfg['assigner']``."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError:
            raise AttributeError(k) from None
        return AttrDict(v) if isinstance(v, dict) and not isinstance(v, AttrDict) else v

    def get(self, k, default=None):
        return getattr(self, k) if k in self else default


def as_cfg(obj):
    if isinstance(obj, dict) and not isinstance(obj, AttrDict):
        return AttrDict(obj)
    if isinstance(obj, (list, tuple)):
        return [as_cfg(o) for o in obj]
    return obj


# ----------------------------------------------------------------------------------------------
# boxes
# ----------------------------------------------------------------------------------------------
def bbox2roi(bbox_list):
    """core/bbox/transforms.py:58-77 - [K,5] = (image index, x1, y1, x2, y2)."""
    sizes = tuple(int(b.size(0)) for b in bbox_list)
    b0 = bbox_list[0]
    key = (sizes, str(b0.device), b0.dtype)
    idx = _ROI_INDEX_CACHE.get(key)
    if idx is None:                     # the image-index column depends on the counts only
        if len(_ROI_INDEX_CACHE) > 64:
            _ROI_INDEX_CACHE.clear()
        col = torch.cat([torch.full((n, 1), float(i)) for i, n in enumerate(sizes)]) if sum(sizes) \
            else torch.zeros((0, 1))
        idx = _ROI_INDEX_CACHE[key] = col.to(device=b0.device, dtype=b0.dtype)
    boxes = torch.cat([b[:, :4] for b in bbox_list], 0) if len(bbox_list) > 1 else b0[:, :4]
    return torch.cat([idx, boxes], dim=1)


_ROI_INDEX_CACHE = {}


def bbox2result(bboxes, labels, num_classes):
    """core/bbox/transforms.py:80-97 - per-class numpy arrays."""
    if bboxes.shape[0] == 0:
        return [np.zeros((0, 5), dtype=np.float32) for _ in range(num_classes)]
    b = bboxes.detach().float().cpu().numpy()
    l = labels.detach().cpu().numpy()
    return [b[l == i, :] for i in range(num_classes)]


_FLIP_AXES = {'horizontal': (1,), 'vertical': (0,), 'diagonal': (1, 0)}


def bbox_flip(bboxes, img_shape, direction='horizontal'):
    """core/bbox/transforms.py:5-32 - mirror (..., 4k) boxes inside an image of img_shape (h, w):
    per mirrored axis the two coordinates swap and become extent - coordinate."""
    assert bboxes.shape[-1] % 4 == 0
    out = bboxes.clone()
    for axis in _FLIP_AXES[direction]:            # axis 1 = x (image width), axis 0 = y
        lo, hi = (0, 2) if axis == 1 else (1, 3)
        out[..., lo::4] = img_shape[axis] - bboxes[..., hi::4]
        out[..., hi::4] = img_shape[axis] - bboxes[..., lo::4]
    return out


def bbox_mapping(bboxes, img_shape, scale_factor, flip, flip_direction='horizontal'):
    """core/bbox/transforms.py:35-44 - original image -> augmented test image."""
    scaled = bboxes * bboxes.new_tensor(scale_factor)
    return bbox_flip(scaled, img_shape, flip_direction) if flip else scaled


def bbox_mapping_back(bboxes, img_shape, scale_factor, flip, flip_direction='horizontal'):
    """core/bbox/transforms.py:47-56 - augmented test image -> original image."""
    unflipped = bbox_flip(bboxes, img_shape, flip_direction) if flip else bboxes
    return (unflipped.view(-1, 4) / unflipped.new_tensor(scale_factor)).view(bboxes.shape)


def merge_aug_bboxes(aug_bboxes, aug_scores, img_metas, rcnn_test_cfg=None):
    """core/post_processing/merge_augs.py:50-76 - boxes of every augmentation mapped back to the
    original image and averaged; scores averaged."""
    back = []
    for boxes, meta in zip(aug_bboxes, img_metas):
        m = meta[0]
        back.append(bbox_mapping_back(boxes, m['img_shape'], m['scale_factor'], m['flip'],
                                      m.get('flip_direction', 'horizontal')))
    merged = torch.stack(back).mean(dim=0)
    if aug_scores is None:
        return merged
    return merged, torch.stack(aug_scores).mean(dim=0)


def bbox_overlaps(bboxes1, bboxes2, mode='iou', eps=1e-6):
    """core/bbox/iou_calculators/iou2d_calculator.py:43-158, non-aligned 'iou' / 'iof'."""
    assert mode in ('iou', 'iof')
    rows, cols = bboxes1.size(0), bboxes2.size(0)
    if rows * cols == 0:
        return bboxes1.new_zeros((rows, cols))
    area1 = (bboxes1[:, 2] - bboxes1[:, 0]) * (bboxes1[:, 3] - bboxes1[:, 1])
    area2 = (bboxes2[:, 2] - bboxes2[:, 0]) * (bboxes2[:, 3] - bboxes2[:, 1])
    lt = torch.max(bboxes1[:, None, :2], bboxes2[None, :, :2])
    rb = torch.min(bboxes1[:, None, 2:4], bboxes2[None, :, 2:4])
    wh = (rb - lt).clamp(min=0)
    overlap = wh[..., 0] * wh[..., 1]
    union = area1[:, None] + area2[None, :] - overlap if mode == 'iou' else area1[:, None]
    union = torch.max(union, union.new_tensor([eps]))
    return overlap / union


@BBOX_CODERS.register_module()
class DeltaXYWHBBoxCoder:
    """core/bbox/coder/delta_xywh_bbox_coder.py."""

    def __init__(self, target_means=(0., 0., 0., 0.), target_stds=(1., 1., 1., 1.),
                 clip_border=True):
        self.means = tuple(target_means)
        self.stds = tuple(target_stds)
        self.clip_border = clip_border

    def encode(self, bboxes, gt_bboxes):
        assert bboxes.size(0) == gt_bboxes.size(0)
        p, g = bboxes.float(), gt_bboxes.float()
        pw, ph = p[..., 2] - p[..., 0], p[..., 3] - p[..., 1]
        gw, gh = g[..., 2] - g[..., 0], g[..., 3] - g[..., 1]
        dx = ((g[..., 0] + g[..., 2]) * 0.5 - (p[..., 0] + p[..., 2]) * 0.5) / pw
        dy = ((g[..., 1] + g[..., 3]) * 0.5 - (p[..., 1] + p[..., 3]) * 0.5) / ph
        cols = [dx, dy, torch.log(gw / pw), torch.log(gh / ph)]
        # python-scalar means / stds: no host->device tensor upload (CUDA-graph capturable)
        return torch.stack([(c - m) / sd for c, m, sd in zip(cols, self.means, self.stds)], dim=-1)

    def decode(self, bboxes, pred_bboxes, max_shape=None, wh_ratio_clip=16 / 1000):
        assert pred_bboxes.size(0) == bboxes.size(0)
        d = pred_bboxes
        dx, dy, dw, dh = (d[:, i::4] * self.stds[i] + self.means[i] for i in range(4))
        max_ratio = abs(math.log(wh_ratio_clip))
        dw = dw.clamp(min=-max_ratio, max=max_ratio)
        dh = dh.clamp(min=-max_ratio, max=max_ratio)
        px = ((bboxes[:, 0] + bboxes[:, 2]) * 0.5).unsqueeze(1)
        py = ((bboxes[:, 1] + bboxes[:, 3]) * 0.5).unsqueeze(1)
        pw = (bboxes[:, 2] - bboxes[:, 0]).unsqueeze(1)
        ph = (bboxes[:, 3] - bboxes[:, 1]).unsqueeze(1)
        gw, gh = pw * dw.exp(), ph * dh.exp()
        gx, gy = px + pw * dx, py + ph * dy
        x1, y1, x2, y2 = gx - gw * 0.5, gy - gh * 0.5, gx + gw * 0.5, gy + gh * 0.5
        if self.clip_border and max_shape is not None:
            x1 = x1.clamp(min=0, max=max_shape[1])
            y1 = y1.clamp(min=0, max=max_shape[0])
            x2 = x2.clamp(min=0, max=max_shape[1])
            y2 = y2.clamp(min=0, max=max_shape[0])
        return torch.stack([x1, y1, x2, y2], dim=-1).view(pred_bboxes.size())


# ----------------------------------------------------------------------------------------------
# assign + sample
# ----------------------------------------------------------------------------------------------
class AssignResult:
    """core/bbox/assigners/assign_result.py (fields used by the head)."""

    def __init__(self, num_gts, gt_inds, max_overlaps, labels=None):
        self.num_gts, self.gt_inds, self.max_overlaps, self.labels = \
            num_gts, gt_inds, max_overlaps, labels

    def add_gt_(self, gt_labels):
        n = len(gt_labels)
        self_inds = torch.arange(1, n + 1, dtype=torch.long, device=gt_labels.device)
        self.gt_inds = torch.cat([self_inds, self.gt_inds])
        self.max_overlaps = torch.cat([self.max_overlaps.new_ones(n), self.max_overlaps])
        if self.labels is not None:
            self.labels = torch.cat([gt_labels, self.labels])


@BBOX_ASSIGNERS.register_module()
class MaxIoUAssigner:
    """core/bbox/assigners/max_iou_assigner.py:127-212 (no ignore regions, GPU assignment)."""

    def __init__(self, pos_iou_thr, neg_iou_thr, min_pos_iou=.0, gt_max_assign_all=True,
                 ignore_iof_thr=-1, ignore_wrt_candidates=True, match_low_quality=True,
                 gpu_assign_thr=-1, iou_calculator=None):
        self.pos_iou_thr, self.neg_iou_thr, self.min_pos_iou = pos_iou_thr, neg_iou_thr, min_pos_iou
        self.gt_max_assign_all = gt_max_assign_all
        self.match_low_quality = match_low_quality
        self.ignore_iof_thr = ignore_iof_thr

    def assign(self, bboxes, gt_bboxes, gt_bboxes_ignore=None, gt_labels=None):
        overlaps = bbox_overlaps(gt_bboxes, bboxes)
        return self.assign_wrt_overlaps(overlaps, gt_labels)

    def assign_wrt_overlaps(self, overlaps, gt_labels=None):
        num_gts, num_bboxes = overlaps.shape
        gt_inds = overlaps.new_full((num_bboxes,), -1, dtype=torch.long)
        if num_gts == 0 or num_bboxes == 0:
            if num_gts == 0:
                gt_inds[:] = 0
            labels = None if gt_labels is None else gt_inds.new_full((num_bboxes,), -1)
            return AssignResult(num_gts, gt_inds, overlaps.new_zeros((num_bboxes,)), labels)
        max_ov, argmax_ov = overlaps.max(dim=0)
        if isinstance(self.neg_iou_thr, (tuple, list)):
            lo, hi = self.neg_iou_thr
            gt_inds[(max_ov >= lo) & (max_ov < hi)] = 0
        else:
            gt_inds[(max_ov >= 0) & (max_ov < self.neg_iou_thr)] = 0
        pos = max_ov >= self.pos_iou_thr
        gt_inds[pos] = argmax_ov[pos] + 1
        if self.match_low_quality:
            gt_max, gt_argmax = overlaps.max(dim=1)
            for i in range(num_gts):
                if gt_max[i] >= self.min_pos_iou:
                    if self.gt_max_assign_all:
                        gt_inds[overlaps[i, :] == gt_max[i]] = i + 1
                    else:
                        gt_inds[gt_argmax[i]] = i + 1
        labels = None
        if gt_labels is not None:
            labels = gt_inds.new_full((num_bboxes,), -1)
            p = gt_inds > 0
            labels[p] = gt_labels[gt_inds[p] - 1]
        return AssignResult(num_gts, gt_inds, max_ov, labels)


class SamplingResult:
    """core/bbox/samplers/sampling_result.py - positives first in ``bboxes`` (:52-54)."""

    def __init__(self, pos_inds, neg_inds, bboxes, gt_bboxes, assign_result, gt_flags):
        self.pos_inds, self.neg_inds = pos_inds, neg_inds
        self.pos_bboxes, self.neg_bboxes = bboxes[pos_inds], bboxes[neg_inds]
        self.pos_is_gt = gt_flags[pos_inds]
        self.num_gts = gt_bboxes.shape[0]
        self.pos_assigned_gt_inds = assign_result.gt_inds[pos_inds] - 1
        if gt_bboxes.numel() == 0:
            self.pos_gt_bboxes = torch.empty_like(gt_bboxes).view(-1, 4)
        else:
            self.pos_gt_bboxes = gt_bboxes.view(-1, 4)[self.pos_assigned_gt_inds, :]
        self.pos_gt_labels = None if assign_result.labels is None else assign_result.labels[pos_inds]

    @property
    def bboxes(self):
        return torch.cat([self.pos_bboxes, self.neg_bboxes])


@BBOX_SAMPLERS.register_module()
class RandomSampler:
    """core/bbox/samplers/{base_sampler.py:34-101, random_sampler.py:31-78}."""

    def __init__(self, num, pos_fraction, neg_pos_ub=-1, add_gt_as_proposals=True, **kwargs):
        self.num, self.pos_fraction = num, pos_fraction
        self.neg_pos_ub, self.add_gt_as_proposals = neg_pos_ub, add_gt_as_proposals

    @staticmethod
    def random_choice(gallery, num):
        perm = torch.randperm(gallery.numel(), device=gallery.device)[:num]
        return gallery[perm]

    def sample(self, assign_result, bboxes, gt_bboxes, gt_labels=None, **kwargs):
        bboxes = bboxes[:, :4]
        gt_flags = bboxes.new_zeros((bboxes.shape[0],), dtype=torch.uint8)
        if self.add_gt_as_proposals and len(gt_bboxes) > 0:
            bboxes = torch.cat([gt_bboxes, bboxes], dim=0)
            assign_result.add_gt_(gt_labels)
            gt_flags = torch.cat([bboxes.new_ones(gt_bboxes.shape[0], dtype=torch.uint8), gt_flags])
        num_pos = int(self.num * self.pos_fraction)
        pos = torch.nonzero(assign_result.gt_inds > 0, as_tuple=False).flatten()
        if pos.numel() > num_pos:
            pos = self.random_choice(pos, num_pos)
        pos = pos.unique()
        num_neg = self.num - pos.numel()
        if self.neg_pos_ub >= 0:
            num_neg = min(num_neg, int(self.neg_pos_ub * max(1, pos.numel())))
        neg = torch.nonzero(assign_result.gt_inds == 0, as_tuple=False).flatten()
        if neg.numel() > num_neg:
            neg = self.random_choice(neg, num_neg)
        neg = neg.unique()
        return SamplingResult(pos, neg, bboxes, gt_bboxes, assign_result, gt_flags)


# ----------------------------------------------------------------------------------------------
# losses
# ----------------------------------------------------------------------------------------------
def _reduce(loss, weight, reduction, avg_factor):
    """models/losses/utils.py:26-52."""
    if weight is not None:
        loss = loss * weight
    if avg_factor is None:
        return loss.mean() if reduction == 'mean' else loss.sum() if reduction == 'sum' else loss
    if reduction == 'mean':
        return loss.sum() / avg_factor          # avg_factor: python number or 0-dim device tensor
    if reduction == 'none':
        return loss
    raise ValueError('avg_factor can not be used with reduction="sum"')


@LOSSES.register_module()
class CrossEntropyLoss(nn.Module):
    """models/losses/cross_entropy_loss.py (softmax variant, the one configs/htd uses)."""

    def __init__(self, use_sigmoid=False, use_mask=False, reduction='mean', class_weight=None,
                 loss_weight=1.0):
        super().__init__()
        if use_sigmoid or use_mask:
            raise NotImplementedError('HTD uses the softmax cross entropy')
        self.reduction, self.loss_weight, self.class_weight = reduction, loss_weight, class_weight

    def forward(self, cls_score, label, weight=None, avg_factor=None, reduction_override=None):
        red = reduction_override or self.reduction
        loss = F.cross_entropy(cls_score, label, reduction='none')
        if weight is not None:
            weight = weight.float()
        return self.loss_weight * _reduce(loss, weight, red, avg_factor)


@LOSSES.register_module()
class SmoothL1Loss(nn.Module):
    """models/losses/smooth_l1_loss.py."""

    def __init__(self, beta=1.0, reduction='mean', loss_weight=1.0):
        super().__init__()
        self.beta, self.reduction, self.loss_weight = beta, reduction, loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None):
        red = reduction_override or self.reduction
        diff = torch.abs(pred - target)
        loss = torch.where(diff < self.beta, 0.5 * diff * diff / self.beta, diff - 0.5 * self.beta)
        return self.loss_weight * _reduce(loss, weight, red, avg_factor)


def accuracy(pred, target, topk=1):
    """models/losses/accuracy.py:4-48 (single k)."""
    if pred.size(0) == 0:
        return pred.new_tensor(0.)
    top = pred.topk(topk, dim=1)[1].t()
    correct = top.eq(target.view(1, -1).expand_as(top))
    return correct[:topk].reshape(-1).float().sum(0, keepdim=True).mul_(100.0 / pred.size(0))


# ----------------------------------------------------------------------------------------------
# post-processing (the step right after the path at inference)
# ----------------------------------------------------------------------------------------------
def multiclass_nms(multi_bboxes, multi_scores, score_thr, nms_cfg, max_num=-1):
    """core/post_processing/bbox_nms.py:7-71.  CUDA tensors: csrc/nms.cu (three launches, no host
    sync until the final read of the detection count that sizes the returned tensors); CPU
    tensors are not supported (no fallback)."""
    cfg = dict(nms_cfg)
    kind = cfg.pop('type', 'nms')
    if kind not in ('nms', 'soft_nms'):
        raise NotImplementedError(f'nms type {kind!r}: configs/htd use nms (htd_resnet50_1x.py:166) '
                                  'and soft_nms (htd_resnet101_2x.py:298)')
    from . import ops
    soft = None
    if kind == 'soft_nms':
        # mmcv.ops.soft_nms defaults: iou_threshold 0.3, sigma 0.5, min_score 1e-3, method 'linear',
        # offset 0 (`iou_thr` is its deprecated alias of `iou_threshold`)
        thr = cfg.get('iou_threshold', cfg.get('iou_thr', 0.3))
        if cfg.get('offset', 0) != 0:
            raise NotImplementedError('soft_nms with offset != 0')
        soft = dict(min_score=cfg.get('min_score', 1e-3), method=cfg.get('method', 'linear'))
    else:
        thr = cfg.get('iou_threshold', cfg.get('iou_thr', 0.5))
    det, labels, count = ops.multiclass_nms(multi_bboxes, multi_scores, score_thr, thr, max_num,
                                            soft=soft)
    n = int(count)                                   # the one host read: output sizes
    return det[:n].to(multi_bboxes.dtype), labels[:n]
