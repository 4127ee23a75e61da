"""ORACLE (test infrastructure, not product code).

CPU restatement, in plain PyTorch + the C RoIAlign of ``oracle/roi_align_ref.c``, of the
reference's HTD RoI-head hot path.  Every function cites the reference file:line it follows
(paths relative to /root/reference).  It exists because the reference tree does not travel to
the GPU box; it is PINNED to the reference in two ways:

  * ``tests/test_oracle_cpu.py::test_restatement_equals_reference*`` run the reference's own
    unmodified modules (``oracle/refshim.py``) next to this file on identical inputs (only
    where /root/reference exists), and
  * ``tests/golden/*.npz`` hold outputs produced by the reference itself
    (``oracle/gen_golden.py``); ``tests/test_oracle_cpu.py::test_restatement_matches_golden*``
    re-check this file against them anywhere.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this module.  State-dict key names equal the reference's so that
``htd_b200.synth.fill_params_`` gives both sides identical weights.
"""
import ctypes
import os
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, '_build', 'liboracle.so')
        if not os.path.isfile(path):
            import subprocess
            subprocess.check_call(['make', '-C', _HERE, '-s'])
        _LIB = ctypes.CDLL(path)
    return _LIB


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


# --------------------------------------------------------------------------------------
# RoIAlign  (mmcv.ops.RoIAlign, aligned=True, sampling_ratio=0, avg) - SURVEY §8 a4/a5
# --------------------------------------------------------------------------------------
class _RoIAlignFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, rois, out_h, out_w, spatial_scale, sampling_ratio, aligned):
        assert x.device.type == 'cpu' and x.dtype in (torch.float32, torch.float64)
        x = x.contiguous()
        rois = rois.to(x.dtype).contiguous()
        B, C, H, W = x.shape
        K = rois.shape[0]
        out = x.new_zeros(K, C, out_h, out_w)
        sfx, ct = ('f32', ctypes.c_float) if x.dtype == torch.float32 else ('f64', ctypes.c_double)
        if K > 0:
            getattr(lib(), 'roi_align_fwd_' + sfx)(
                _ptr(x), _ptr(rois), _ptr(out), B, C, H, W, K, out_h, out_w,
                ct(spatial_scale), int(sampling_ratio), int(bool(aligned)))
        ctx.save_for_backward(rois)
        ctx.cfg = (x.shape, out_h, out_w, spatial_scale, sampling_ratio, aligned, sfx, ct)
        return out

    @staticmethod
    def backward(ctx, go):
        (rois,) = ctx.saved_tensors
        shape, out_h, out_w, spatial_scale, sampling_ratio, aligned, sfx, ct = ctx.cfg
        B, C, H, W = shape
        go = go.contiguous()
        gi = go.new_zeros(shape)
        K = rois.shape[0]
        if K > 0:
            getattr(lib(), 'roi_align_bwd_' + sfx)(
                _ptr(go), _ptr(rois), _ptr(gi), B, C, H, W, K, out_h, out_w,
                ct(spatial_scale), int(sampling_ratio), int(bool(aligned)))
        return gi, None, None, None, None, None, None


class RoIAlign(nn.Module):
    """Signature of mmcv.ops.RoIAlign as used at base_roi_extractor.py:49-55."""

    def __init__(self, output_size, spatial_scale=1.0, sampling_ratio=0, pool_mode='avg',
                 aligned=True, use_torchvision=False):
        super().__init__()
        self.output_size = (output_size, output_size) if isinstance(output_size, int) \
            else tuple(output_size)
        self.spatial_scale = float(spatial_scale)
        self.sampling_ratio = int(sampling_ratio)
        assert pool_mode == 'avg'
        self.aligned = aligned

    def forward(self, x, rois):
        return _RoIAlignFn.apply(x, rois, self.output_size[0], self.output_size[1],
                                 self.spatial_scale, self.sampling_ratio, self.aligned)


def map_roi_levels(rois, num_levels, finest_scale=56):
    """single_level_roi_extractor.py:32-51 (duplicate: htd_bbox_head.py:129-135)."""
    scale = torch.sqrt((rois[:, 3] - rois[:, 1]) * (rois[:, 4] - rois[:, 2]))
    target_lvls = torch.floor(torch.log2(scale / finest_scale + 1e-6))
    return target_lvls.clamp(min=0, max=num_levels - 1).long()


def bbox2roi(bbox_list):
    """core/bbox/transforms.py:58-77."""
    rois_list = []
    for img_id, bboxes in enumerate(bbox_list):
        if bboxes.size(0) > 0:
            img_inds = bboxes.new_full((bboxes.size(0), 1), img_id)
            rois = torch.cat([img_inds, bboxes[:, :4]], dim=-1)
        else:
            rois = bboxes.new_zeros((0, 5))
        rois_list.append(rois)
    return torch.cat(rois_list, 0)


def bbox_overlaps(b1, b2, eps=1e-6):
    """iou2d_calculator.py:43-158, mode='iou', is_aligned=False branch (:129-150)."""
    rows, cols = b1.size(0), b2.size(0)
    if rows * cols == 0:
        return b1.new_zeros((rows, cols))
    area1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    area2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    lt = torch.max(b1[:, None, :2], b2[None, :, :2])
    rb = torch.min(b1[:, None, 2:], b2[None, :, 2:])
    wh = (rb - lt).clamp(min=0)
    overlap = wh[..., 0] * wh[..., 1]
    union = area1[:, None] + area2[None, :] - overlap
    union = torch.max(union, union.new_tensor([eps]))
    return overlap / union


def bbox2delta(proposals, gt, means, stds):
    """delta_xywh_bbox_coder.py:78-120."""
    proposals = proposals.float()      # the reference computes targets in fp32 (:98-99)
    gt = gt.float()
    px = (proposals[..., 0] + proposals[..., 2]) * 0.5
    py = (proposals[..., 1] + proposals[..., 3]) * 0.5
    pw = proposals[..., 2] - proposals[..., 0]
    ph = proposals[..., 3] - proposals[..., 1]
    gx = (gt[..., 0] + gt[..., 2]) * 0.5
    gy = (gt[..., 1] + gt[..., 3]) * 0.5
    gw = gt[..., 2] - gt[..., 0]
    gh = gt[..., 3] - gt[..., 1]
    deltas = torch.stack([(gx - px) / pw, (gy - py) / ph, torch.log(gw / pw),
                          torch.log(gh / ph)], dim=-1)
    means = deltas.new_tensor(means).unsqueeze(0)
    stds = deltas.new_tensor(stds).unsqueeze(0)
    return deltas.sub_(means).div_(stds)


def delta2bbox(rois, deltas, means, stds, max_shape=None, wh_ratio_clip=16 / 1000):
    """delta_xywh_bbox_coder.py:123-204 (class-agnostic: deltas [N,4])."""
    means = deltas.new_tensor(means).view(1, -1)
    stds = deltas.new_tensor(stds).view(1, -1)
    d = deltas * stds + means
    dx, dy, dw, dh = d[:, 0::4], d[:, 1::4], d[:, 2::4], d[:, 3::4]
    max_ratio = np.abs(np.log(wh_ratio_clip))
    dw = dw.clamp(min=-max_ratio, max=max_ratio)
    dh = dh.clamp(min=-max_ratio, max=max_ratio)
    px = ((rois[:, 0] + rois[:, 2]) * 0.5).unsqueeze(1).expand_as(dx)
    py = ((rois[:, 1] + rois[:, 3]) * 0.5).unsqueeze(1).expand_as(dy)
    pw = (rois[:, 2] - rois[:, 0]).unsqueeze(1).expand_as(dw)
    ph = (rois[:, 3] - rois[:, 1]).unsqueeze(1).expand_as(dh)
    gw = pw * dw.exp()
    gh = ph * dh.exp()
    gx = px + pw * dx
    gy = py + ph * dy
    x1, y1, x2, y2 = gx - gw * 0.5, gy - gh * 0.5, gx + gw * 0.5, gy + gh * 0.5
    if max_shape is not None:
        x1 = x1.clamp(min=0, max=max_shape[1])
        y1 = y1.clamp(min=0, max=max_shape[0])
        x2 = x2.clamp(min=0, max=max_shape[1])
        y2 = y2.clamp(min=0, max=max_shape[0])
    return torch.stack([x1, y1, x2, y2], dim=-1).view(deltas.size())


# --------------------------------------------------------------------------------------
# extractors
# --------------------------------------------------------------------------------------
class SingleRoIExtractor(nn.Module):
    """single_level_roi_extractor.py:53-99."""

    def __init__(self, out_channels=256, featmap_strides=(4, 8, 16, 32), finest_scale=56,
                 output_size=7, sampling_ratio=0):
        super().__init__()
        self.roi_layers = nn.ModuleList(
            [RoIAlign(output_size, 1 / s, sampling_ratio) for s in featmap_strides])
        self.out_channels = out_channels
        self.featmap_strides = featmap_strides
        self.finest_scale = finest_scale

    @property
    def num_inputs(self):
        return len(self.featmap_strides)

    def map_roi_levels(self, rois, num_levels):
        return map_roi_levels(rois, num_levels, self.finest_scale)

    def forward(self, feats, rois):
        out_size = self.roi_layers[0].output_size
        num_levels = len(feats)
        roi_feats = feats[0].new_zeros(rois.size(0), self.out_channels, *out_size)
        target_lvls = map_roi_levels(rois, num_levels, self.finest_scale)
        for i in range(num_levels):
            inds = (target_lvls == i).nonzero(as_tuple=False).squeeze(1)
            if inds.numel() > 0:
                roi_feats[inds] = self.roi_layers[i](feats[i], rois[inds])
            else:
                roi_feats = roi_feats + feats[i].sum() * 0.
        return roi_feats


class AdptRoIExtractor(nn.Module):
    """BA extractor, adaptative_roi_extractor.py:24-91 (aggregation='sum', edge from cfg)."""

    def __init__(self, out_channels=256, featmap_strides=(4, 8, 16, 32), edge=1, output_size=7,
                 sampling_ratio=0):
        super().__init__()
        self.roi_layers = nn.ModuleList(
            [RoIAlign(output_size, 1 / s, sampling_ratio) for s in featmap_strides])
        self.out_channels = out_channels
        self.featmap_strides = featmap_strides
        self.edge = edge
        self.pool = nn.AdaptiveAvgPool2d(1)
        self.conv1 = nn.Conv2d(256, 128, 1)
        self.conv2 = nn.Conv2d(128, 1, 1)
        self.att = nn.Sequential(self.pool, self.conv1, nn.Tanh(), self.conv2)

    @property
    def num_inputs(self):
        return len(self.featmap_strides)

    def forward(self, feats, rois):
        out_size = self.roi_layers[0].output_size
        num_levels = len(feats)
        if rois.size(0) == 0:
            return feats[0].new_zeros(0, self.out_channels, *out_size)
        roi_feat, atts = [], []
        for i in range(num_levels):
            t = self.roi_layers[i](feats[i], rois)
            # reference: self.att(t).squeeze().unsqueeze(0) (:73) - for n == 1 the reference
            # crashes; the restatement keeps the RoI dim (documented).
            atts.append(self.att(t).reshape(1, -1))
            roi_feat.append(t.unsqueeze(0))
        roi_feat = torch.cat(roi_feat, dim=0)
        lvl, n, c, x, y = roi_feat.size()
        atts = torch.cat(atts, dim=0).softmax(0)
        atts = atts.unsqueeze(-1).repeat(1, 1, c * x * y).view(lvl, n, c, x, y)
        fused = (atts * roi_feat).sum(0)
        enh = self.roi_layers[0](feats[0], rois)
        e = self.edge
        mask = torch.ones_like(enh)
        mask[:, :, e:-e, e:-e] = 0     # in-place zeroing at :88, expressed as a mask
        return fused + enh * mask


# --------------------------------------------------------------------------------------
# heads
# --------------------------------------------------------------------------------------
class ConvModule(nn.Module):
    """mmcv ConvModule defaults used by the reference: conv -> GN -> ReLU, bias='auto'."""

    def __init__(self, i, o, k, padding=0, norm_groups=None, bias='auto', act=True):
        super().__init__()
        with_norm = norm_groups is not None
        if bias == 'auto':
            bias = not with_norm
        self.conv = nn.Conv2d(i, o, k, 1, padding, bias=bias)
        self.gn = nn.GroupNorm(norm_groups, o) if with_norm else None
        self.act = act                    # act_cfg=None (FPN): no activation

    def forward(self, x):
        x = self.conv(x)
        if self.gn is not None:
            x = self.gn(x)
        return F.relu(x) if self.act else x


class GlobalContextHead(nn.Module):
    """SFA, global_context_head.py:323-401."""

    def __init__(self, num_convs=4, in_channels=256, conv_out_channels=256, num_classes=81,
                 loss_weight=3.0):
        super().__init__()
        self.convs = nn.ModuleList(
            [ConvModule(in_channels if i == 0 else conv_out_channels, conv_out_channels, 3, 1)
             for i in range(num_convs)])
        self.fc = nn.Linear(conv_out_channels, num_classes)
        self.loss_weight = loss_weight

    def forward(self, feats):
        x = feats[-1]
        for conv in self.convs:
            x = conv(x)
        x = F.adaptive_avg_pool2d(x, 1)
        return self.fc(x.reshape(x.size(0), -1)), x

    def loss(self, pred, labels):
        labels = [lbl.unique() for lbl in labels]
        targets = pred.new_zeros(pred.size())
        for i, label in enumerate(labels):
            targets[i, label] = 1.0
        return self.loss_weight * F.binary_cross_entropy_with_logits(pred, targets)


def fuse_global(roi_feats, global_feat, rois):
    """htd_roi_head.py:133-141 / htd_bbox_head.py:147-155."""
    img_inds = torch.unique(rois[:, 0], sorted=True).long()
    fused = torch.zeros_like(roi_feats)
    for img_id in img_inds:
        inds = rois[:, 0] == img_id.item()
        fused[inds] = roi_feats[inds] + global_feat[img_id]
    return fused


def cross_entropy_loss(cls_score, labels, label_weights, avg_factor):
    """losses/cross_entropy_loss.py:9-39 + utils.py:26-52."""
    loss = F.cross_entropy(cls_score, labels, reduction='none')
    return (loss * label_weights.to(loss.dtype)).sum() / avg_factor


def smooth_l1(pred, target, weight, avg_factor, beta=1.0):
    """losses/smooth_l1_loss.py:9-27."""
    diff = torch.abs(pred - target)
    loss = torch.where(diff < beta, 0.5 * diff * diff / beta, diff - 0.5 * beta)
    return (loss * weight).sum() / avg_factor


def accuracy(pred, target):
    """losses/accuracy.py:4-48, topk=1."""
    if pred.size(0) == 0:
        return pred.new_tensor(0.)
    lab = pred.topk(1, dim=1)[1].t()
    correct = lab.eq(target.view(1, -1))
    return correct[:1].reshape(-1).float().sum(0, keepdim=True).mul_(100.0 / pred.size(0))


class BBoxHeadBase(nn.Module):
    """bbox_head.py:85-335 (targets, loss, refine, regress) for class-agnostic regression."""
    num_classes = 80
    target_means = (0., 0., 0., 0.)
    target_stds = (0.1, 0.1, 0.2, 0.2)

    def get_targets(self, samp, pos_weight=-1):
        labels, lw, bt, bw = [], [], [], []
        for res in samp:
            num_pos, num_neg = res.pos_bboxes.size(0), res.neg_bboxes.size(0)
            n = num_pos + num_neg
            lab = res.pos_bboxes.new_full((n,), self.num_classes, dtype=torch.long)
            w = res.pos_bboxes.new_zeros(n)
            t = res.pos_bboxes.new_zeros(n, 4)
            tw = res.pos_bboxes.new_zeros(n, 4)
            if num_pos > 0:
                lab[:num_pos] = res.pos_gt_labels
                w[:num_pos] = 1.0 if pos_weight <= 0 else pos_weight
                t[:num_pos] = bbox2delta(res.pos_bboxes, res.pos_gt_bboxes, self.target_means,
                                         self.target_stds)
                tw[:num_pos] = 1
            if num_neg > 0:
                w[-num_neg:] = 1.0
            labels.append(lab), lw.append(w), bt.append(t), bw.append(tw)
        return torch.cat(labels), torch.cat(lw), torch.cat(bt), torch.cat(bw)

    def loss(self, cls_score, bbox_pred, rois, labels, label_weights, bbox_targets,
             bbox_weights):
        losses = {}
        avg_factor = max(torch.sum(label_weights > 0).float().item(), 1.)
        losses['loss_cls'] = cross_entropy_loss(cls_score, labels, label_weights, avg_factor)
        losses['acc'] = accuracy(cls_score, labels)
        pos = (labels >= 0) & (labels < self.num_classes)
        if pos.any():
            losses['loss_bbox'] = smooth_l1(bbox_pred.view(bbox_pred.size(0), 4)[pos],
                                            bbox_targets[pos], bbox_weights[pos],
                                            avg_factor=bbox_targets.size(0))
        else:
            losses['loss_bbox'] = bbox_pred[pos].sum()
        return losses

    def regress_by_class(self, rois, label, bbox_pred, img_shape):
        bboxes = delta2bbox(rois[:, 1:], bbox_pred, self.target_means, self.target_stds,
                            max_shape=img_shape)
        return torch.cat((rois[:, [0]], bboxes), dim=1)

    def refine_bboxes(self, rois, labels, bbox_preds, pos_is_gts, img_shapes):
        out = []
        for i in range(len(img_shapes)):
            inds = torch.nonzero(rois[:, 0] == i, as_tuple=False).squeeze(1)
            b = self.regress_by_class(rois[inds], labels[inds], bbox_preds[inds],
                                      img_shapes[i])[:, 1:]
            keep = pos_is_gts[i].new_ones(inds.numel())
            keep[:len(pos_is_gts[i])] = 1 - pos_is_gts[i]
            out.append(b[keep.type(torch.bool)])
        return out


class Shared2FCBBoxHead(BBoxHeadBase):
    """convfc_bbox_head.py:135-189 with the htd_resnet50_1x.py:57-74 arguments."""

    def __init__(self):
        super().__init__()
        self.shared_fcs = nn.ModuleList([nn.Linear(256 * 49, 1024), nn.Linear(1024, 1024)])
        self.fc_cls = nn.Linear(1024, 81)
        self.fc_reg = nn.Linear(1024, 4)

    def forward(self, x):
        x = x.flatten(1)
        for fc in self.shared_fcs:
            x = F.relu(fc(x))
        return self.fc_cls(x), self.fc_reg(x)


class HTDBBoxHead(BBoxHeadBase):
    """htd_bbox_head.py:34-230 with relpace=False, average=False, alpha=1 (all HTD configs)."""
    target_stds = (0.05, 0.05, 0.1, 0.1)

    def __init__(self):
        super().__init__()
        mid = 16 * 36
        self.fc_cls = nn.Linear(1024, 81)
        self.fc_reg = nn.Linear(1024, 4)
        self.convs = nn.Sequential(
            ConvModule(256, mid, 3, 1, norm_groups=36, bias=False),
            ConvModule(mid, mid, 3, 1, norm_groups=36, bias=False),
            ConvModule(mid, mid, 3, 1, norm_groups=36, bias=False),
            ConvModule(mid, 1024, 3, 1, norm_groups=None, bias=False))
        self.fcs = nn.Sequential(nn.Linear(256 * 49, 1024), nn.ReLU(), nn.Linear(1024, 1024),
                                 nn.ReLU())
        for i in range(4):
            setattr(self, f'graph_lvl{i}_cls', nn.Linear(1024, 1024))

    def graph_group(self, rois_, x_, sam_, lvl):
        """One (image, level) group, htd_bbox_head.py:204-217.  Returns new_cls and the mask."""
        M = bbox_overlaps(rois_, rois_).fill_diagonal_(1.)
        M[M > 0] = 1.
        D = torch.diag(torch.sum(M, dim=-1).pow(-0.5))
        A_local = torch.mm(torch.mm(D, M), D)
        G = 1. - M
        mixed = torch.mm(A_local, x_)
        sim = torch.mm(sam_, sam_.t())
        A_global = (G * sim).softmax(-1)
        lin = getattr(self, f'graph_lvl{lvl}_cls')
        return F.relu(lin(torch.matmul(A_global, mixed))), M

    def forward(self, x_cls, x_reg, feat, rois, fc_cls_0, enhanced_feat=None, pos_rois=None,
                global_feat=None, return_masks=False):
        prototype = torch.cat((fc_cls_0.weight, fc_cls_0.bias.unsqueeze(1)), 1).detach()
        bs = int(torch.max(rois[..., 0])) + 1
        if global_feat is not None:
            x_cls_glb = fuse_global(x_cls, global_feat, rois)
            x_reg = fuse_global(x_reg, global_feat, pos_rois)
            x_cls_glb = self.fcs(x_cls_glb.flatten(1))
        x_reg = x_reg + enhanced_feat
        x_reg = self.convs(x_reg)
        x_reg = F.avg_pool2d(x_reg, 7).view(x_reg.size(0), -1)
        x_cls = self.fcs(x_cls.flatten(1))
        sam = torch.mm(fc_cls_0(x_cls).softmax(-1), prototype)
        target_lvls = map_roi_levels(rois, len(feat))
        refined = x_cls.new_zeros(x_cls.size(0), 1024)
        masks = {}
        for b in range(bs):
            bs_indx = rois[..., 0] == b
            for i in range(len(feat)):
                sel = torch.logical_and(target_lvls == i, bs_indx)
                if sel.any():
                    new_cls, M = self.graph_group(rois[sel, 1:5], x_cls[sel, :], sam[sel, :], i)
                    refined = refined.index_put((sel.nonzero(as_tuple=True)[0],), new_cls)
                    masks[(b, i)] = (sel.nonzero(as_tuple=True)[0], M)
        feat_cls_new = (x_cls_glb if global_feat is not None else x_cls) + refined
        out = self.fc_cls(feat_cls_new), self.fc_reg(x_reg)
        return out + (masks,) if return_masks else out


# --------------------------------------------------------------------------------------
# multi-class NMS - checker of csrc/nms.cu
# --------------------------------------------------------------------------------------
class FPN(nn.Module):
    """necks/fpn.py:9-216 with the configs/htd arguments (htd_resnet50_1x.py:17-21: no extra convs,
    no norm / activation, nearest top-down by target size, extra levels = max_pool2d(x, 1, 2))."""

    def __init__(self, in_channels=(256, 512, 1024, 2048), out_channels=256, num_outs=5):
        super().__init__()
        self.num_outs = num_outs
        self.lateral_convs = nn.ModuleList()
        self.fpn_convs = nn.ModuleList()
        for c in in_channels:                                    # fpn.py:110-131
            self.lateral_convs.append(ConvModule(c, out_channels, 1, act=False))
            self.fpn_convs.append(ConvModule(out_channels, out_channels, 3, padding=1, act=False))

    def forward(self, inputs):
        lat = [conv(x) for conv, x in zip(self.lateral_convs, inputs)]           # fpn.py:170-174
        for i in range(len(lat) - 1, 0, -1):                                     # fpn.py:177-190
            lat[i - 1] = lat[i - 1] + F.interpolate(lat[i], size=lat[i - 1].shape[2:], mode='nearest')
        outs = [conv(x) for conv, x in zip(self.fpn_convs, lat)]                 # fpn.py:194-196
        while len(outs) < self.num_outs:                                         # fpn.py:198-201
            outs.append(F.max_pool2d(outs[-1], 1, stride=2))
        return tuple(outs)


def anchor_grid(featmap_sizes, strides=(4, 8, 16, 32, 64), ratios=(0.5, 1.0, 2.0), scales=(8.,)):
    """core/anchor/anchor_generator.py:142-187 (base anchors, scale_major, center_offset 0) and
    :232-272 (grid): per level [H*W*A, 4], rows ordered (y, x, anchor)."""
    out = []
    r, sc = torch.Tensor(ratios), torch.Tensor(scales)
    for (h, w), stride in zip(featmap_sizes, strides):
        hr = torch.sqrt(r)
        wr = 1 / hr
        ws = (stride * wr[:, None] * sc[None, :]).view(-1)
        hs = (stride * hr[:, None] * sc[None, :]).view(-1)
        base = torch.stack([-0.5 * ws, -0.5 * hs, 0.5 * ws, 0.5 * hs], dim=-1)
        sx = torch.arange(0, int(w)) * stride
        sy = torch.arange(0, int(h)) * stride
        xx, yy = sx.repeat(len(sy)), sy.view(-1, 1).repeat(1, len(sx)).view(-1)
        shifts = torch.stack([xx, yy, xx, yy], dim=-1).type_as(base)
        out.append((base[None] + shifts[:, None]).view(-1, 4))
    return out


def rpn_proposals_single(cls_scores, bbox_preds, mlvl_anchors, img_shape, nms_pre, nms_post, nms_thr,
                         min_bbox_size=0):
    """dense_heads/rpn_head.py:77-168 (use_sigmoid_cls): per level the nms_pre best anchors by
    sigmoid score, delta2bbox with means 0 / stds 1 clipped to the image, batched NMS with the
    level index as the class (mmcv batched_nms: boxes shifted by level * (max coordinate + 1),
    greedy NMS in descending score order), first nms_post.  Returns [n, 5]."""
    scores, boxes, ids = [], [], []
    for l, (c, d, a) in enumerate(zip(cls_scores, bbox_preds, mlvl_anchors)):
        s = c.permute(1, 2, 0).reshape(-1).sigmoid()
        d = d.permute(1, 2, 0).reshape(-1, 4)
        if nms_pre > 0 and s.shape[0] > nms_pre:
            rs, ri = s.sort(descending=True)
            s, d, a = rs[:nms_pre], d[ri[:nms_pre]], a[ri[:nms_pre]]
        scores.append(s)
        boxes.append(delta2bbox(a, d, (0., 0., 0., 0.), (1., 1., 1., 1.), max_shape=img_shape))
        ids.append(torch.full((s.shape[0],), l, dtype=torch.long))
    scores, boxes, ids = torch.cat(scores), torch.cat(boxes), torch.cat(ids)
    if min_bbox_size > 0:
        ok = ((boxes[:, 2] - boxes[:, 0]) >= min_bbox_size) & ((boxes[:, 3] - boxes[:, 1]) >= min_bbox_size)
        scores, boxes, ids = scores[ok], boxes[ok], ids[ok]
    shifted = boxes + (ids.to(boxes) * (boxes.max() + 1))[:, None]
    keep = nms_greedy(shifted, scores, nms_thr)
    return torch.cat([boxes[keep], scores[keep, None]], dim=1)[:nms_post]


def nms_greedy(boxes, scores, iou_thr):
    """mmcv.ops.nms (mmcv-full 1.2.1, un-vendored; the same published algorithm as
    torchvision.ops.nms, against which tests/test_oracle_cpu.py pins this function): visit boxes
    in descending score order (stable: ties keep index order), keep a box unless an earlier kept
    box has IoU > iou_thr with it; IoU = inter / (Sa + Sb - inter), fp32, offset 0.  Returns the
    kept indices in visiting order."""
    b = boxes.detach().float().cpu().numpy()
    order = torch.argsort(scores.detach().float().cpu(), descending=True, stable=True).numpy()
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    supp = np.zeros(len(b), dtype=bool)
    keep = []
    for pos, i in enumerate(order):
        if supp[i]:
            continue
        keep.append(int(i))
        rest = order[pos + 1:]
        w = np.maximum(np.minimum(b[i, 2], b[rest, 2]) - np.maximum(b[i, 0], b[rest, 0]),
                       np.float32(0))
        h = np.maximum(np.minimum(b[i, 3], b[rest, 3]) - np.maximum(b[i, 1], b[rest, 1]),
                       np.float32(0))
        inter = w * h
        with np.errstate(divide='ignore', invalid='ignore'):
            iou = inter / (area[i] + area[rest] - inter)
        supp[rest[iou > np.float32(iou_thr)]] = True
    return torch.tensor(keep, dtype=torch.long)


def soft_nms(boxes, scores, iou_threshold=0.3, sigma=0.5, min_score=1e-3, method='linear', offset=0):
    """mmcv.ops.soft_nms of mmcv-full 1.2.1 (un-vendored dependency, README.md:11): the published
    sequential loop, restated literally in oracle/soft_nms_ref.c.  Returns (dets [m,5], inds [m])
    in selection order; dets hold the boxes as given and the decayed scores."""
    n = boxes.size(0)
    b = boxes.detach().to(torch.float32).contiguous().cpu()
    s = scores.detach().to(torch.float32).contiguous().cpu()
    dets = torch.empty((n, 5), dtype=torch.float32)
    inds = torch.empty(n, dtype=torch.int64)
    fn = lib().oracle_soft_nms
    fn.restype = ctypes.c_longlong
    fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_float,
                   ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                   ctypes.c_void_p]
    m = fn(_ptr(b), _ptr(s), n, float(iou_threshold), float(sigma), float(min_score),
           {'naive': 0, 'linear': 1, 'gaussian': 2}[method], int(offset), _ptr(dets), _ptr(inds))
    return dets[:m], inds[:m]


def soft_nms_vectorised(boxes, scores, iou_threshold, sigma, min_score, method='linear',
                        max_out=-1):
    """The same loop with every inner pass as one array operation - the formulation the CUDA kernel
    uses (csrc/nms.cu): first-occurrence argmax, swap, decay of all later boxes at once, and the
    'overwrite with the last box' removals as ONE unstable compaction: with m survivors among the
    boxes after position i, the k-th dead slot (ascending) among the first m is filled by the k-th
    live box from the end (descending).  Checked against the literal loop in tests/test_oracle_cpu."""
    import numpy as np
    b = boxes.detach().cpu().numpy().astype(np.float32).copy()
    sc = scores.detach().cpu().numpy().astype(np.float32).copy()
    area = ((b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])).astype(np.float32)
    ind = np.arange(b.shape[0], dtype=np.int64)
    n = b.shape[0]
    dets, keep = [], []
    i = 0
    while i < n and (max_out < 0 or i < max_out):
        mp = i + int(np.argmax(sc[i:n]))               # numpy: first occurrence of the maximum
        for arr in (b, sc, area, ind):
            arr[[i, mp]] = arr[[mp, i]]
        dets.append(np.concatenate([b[i], sc[i:i + 1]]))
        keep.append(ind[i])
        r = slice(i + 1, n)
        w = np.maximum(np.float32(0), np.minimum(b[i, 2], b[r, 2]) - np.maximum(b[i, 0], b[r, 0]))
        h = np.maximum(np.float32(0), np.minimum(b[i, 3], b[r, 3]) - np.maximum(b[i, 1], b[r, 1]))
        inter = (w * h).astype(np.float32)
        ovr = (inter / ((area[i] + area[r]).astype(np.float32) - inter)).astype(np.float32)
        if method == 'linear':
            wt = np.where(ovr >= np.float32(iou_threshold), np.float32(1) - ovr, np.float32(1))
        elif method == 'naive':
            wt = np.where(ovr >= np.float32(iou_threshold), np.float32(0), np.float32(1))
        else:
            wt = np.exp(-(ovr * ovr) / np.float32(sigma)).astype(np.float32)
        sc[r] = (sc[r] * wt.astype(np.float32)).astype(np.float32)
        live = sc[r] >= np.float32(min_score)
        m = int(live.sum())
        holes = np.nonzero(~live[:m])[0]
        fill = np.nonzero(live[m:])[0][::-1] + m
        for arr in (b, sc, area, ind):
            v = arr[r]
            v[holes] = v[fill]
        n = i + 1 + m
        i += 1
    if not dets:
        return torch.zeros((0, 5)), torch.zeros((0,), dtype=torch.long)
    return torch.from_numpy(np.stack(dets)), torch.from_numpy(np.asarray(keep, dtype=np.int64))


def multiclass_nms(multi_bboxes, multi_scores, score_thr, iou_thr, max_num=-1, nms_type='nms',
                   min_score=1e-3, sigma=0.5, method='linear'):
    """core/post_processing/bbox_nms.py:7-71 with nms_cfg = dict(type='nms', iou_threshold=...)
    and mmcv's batched_nms (class_agnostic=False, fewer than split_thr boxes): shift the boxes of
    class c by c * (max coordinate + 1), one NMS over all of them, ``dets`` in descending score
    order, first ``max_num``.  ``nms_type='soft_nms'`` (configs/htd/htd_resnet101_2x.py:298):
    mmcv's soft_nms on the shifted boxes, scores replaced by the decayed ones (batched_nms returns
    ``dets[:, -1]``).  Returns (dets [n,5], labels [n])."""
    num_classes = multi_scores.size(1) - 1
    if multi_bboxes.shape[1] > 4:
        bboxes = multi_bboxes.view(multi_scores.size(0), -1, 4)
    else:
        bboxes = multi_bboxes[:, None].expand(multi_scores.size(0), num_classes, 4)
    scores = multi_scores[:, :-1]
    valid = scores > score_thr
    bboxes = bboxes[valid]
    scores = scores[valid]
    labels = valid.nonzero(as_tuple=False)[:, 1]
    if bboxes.numel() == 0:
        return multi_bboxes.new_zeros((0, 5)), multi_bboxes.new_zeros((0,), dtype=torch.long)
    max_coordinate = bboxes.max()
    offsets = labels.to(bboxes) * (max_coordinate + 1)
    if nms_type == 'soft_nms':
        dets, keep = soft_nms(bboxes + offsets[:, None], scores, iou_thr, sigma, min_score, method)
        if max_num > 0:
            dets, keep = dets[:max_num], keep[:max_num]
        return torch.cat([bboxes[keep], dets[:, 4:5].to(bboxes)], -1), labels[keep]
    assert nms_type == 'nms', nms_type
    keep = nms_greedy(bboxes + offsets[:, None], scores, iou_thr)
    if max_num > 0:
        keep = keep[:max_num]
    return torch.cat([bboxes[keep], scores[keep, None]], -1), labels[keep]


# --------------------------------------------------------------------------------------
# assign + sample - checker of csrc/assign_sample.cu
# --------------------------------------------------------------------------------------
def max_iou_assign(bboxes, gt_bboxes, gt_labels, pos_iou_thr, neg_iou_thr, min_pos_iou=0.,
                   match_low_quality=True):
    """MaxIoUAssigner.assign / assign_wrt_overlaps (max_iou_assigner.py:84-212) without ignore
    regions, float ``neg_iou_thr``, ``gt_max_assign_all=True``.  Returns (gt_inds int64 with
    -1 ignore / 0 negative / i+1 positive, max_overlaps, labels)."""
    overlaps = bbox_overlaps(gt_bboxes, bboxes)                       # :107
    num_gts, num_bboxes = overlaps.size(0), overlaps.size(1)
    gt_inds = overlaps.new_full((num_bboxes,), -1, dtype=torch.long)  # :142
    if num_gts == 0 or num_bboxes == 0:                               # :146-161
        max_overlaps = overlaps.new_zeros((num_bboxes,))
        if num_gts == 0:
            gt_inds[:] = 0
        labels = overlaps.new_full((num_bboxes,), -1, dtype=torch.long)
        return gt_inds, max_overlaps, labels
    max_overlaps, argmax_overlaps = overlaps.max(dim=0)               # :165
    gt_max_overlaps, _ = overlaps.max(dim=1)                          # :168
    gt_inds[(max_overlaps >= 0) & (max_overlaps < neg_iou_thr)] = 0   # :171-173
    pos = max_overlaps >= pos_iou_thr                                 # :181-182
    gt_inds[pos] = argmax_overlaps[pos] + 1
    if match_low_quality:                                             # :184-198
        for i in range(num_gts):
            if gt_max_overlaps[i] >= min_pos_iou:
                gt_inds[overlaps[i, :] == gt_max_overlaps[i]] = i + 1
    labels = gt_inds.new_full((num_bboxes,), -1)                      # :200-209
    p = torch.nonzero(gt_inds > 0, as_tuple=False).squeeze(1)
    if p.numel() > 0:
        labels[p] = gt_labels[gt_inds[p] - 1]
    return gt_inds, max_overlaps, labels


def choose_by_keys(gallery, num, keys):
    """Stand-in for RandomSampler.random_choice (random_sampler.py:31-54: ``gallery[randperm(n)
    [:num]]``): the ``num`` members of ``gallery`` with the smallest (key, index).  Same
    distribution for i.i.d. uniform keys; deterministic given the keys."""
    k = keys[gallery].double() * (2.0 ** 40) + gallery.double()      # exact: keys are fp32 < 2
    order = torch.argsort(k, stable=True)
    return gallery[order[:num]]


def assign_sample_image(bboxes, gt_bboxes, gt_labels, keys, cfg, valid=None):
    """One image of HTDRoIHead.forward_train's assign + sample (htd_roi_head.py:254-264):
    MaxIoUAssigner.assign, then BaseSampler.sample (base_sampler.py:34-101) with
    RandomSampler._sample_pos/_sample_neg (random_sampler.py:56-78) drawing by ``keys`` (indexed
    like cat([gt_bboxes, bboxes]) when gts are added, else like bboxes).  ``valid`` (optional bool
    [n]) restricts the proposals that exist.  Returns a namespace with the SamplingResult fields
    (sampling_result.py) plus ``pos_inds`` / ``neg_inds`` / ``gt_inds`` / ``max_overlaps``."""
    a, s = cfg['assigner'], cfg['sampler']
    bboxes = bboxes[:, :4]
    keep = torch.arange(bboxes.size(0)) if valid is None else torch.nonzero(valid).squeeze(1)
    boxes = bboxes[keep]
    gt_inds, max_ov, _ = max_iou_assign(boxes, gt_bboxes, gt_labels, a['pos_iou_thr'],
                                        a['neg_iou_thr'], a.get('min_pos_iou', 0.),
                                        a.get('match_low_quality', True))
    gt_flags = torch.zeros(boxes.size(0), dtype=torch.uint8)
    g = gt_bboxes.size(0)
    src = keep.clone()                       # index of every candidate in cat([gt, bboxes])
    if s.get('add_gt_as_proposals', True) and g > 0:                  # base_sampler.py:73-81
        boxes = torch.cat([gt_bboxes, boxes], 0)
        gt_inds = torch.cat([torch.arange(1, g + 1), gt_inds])        # AssignResult.add_gt_
        max_ov = torch.cat([max_ov.new_ones(g), max_ov])
        gt_flags = torch.cat([torch.ones(g, dtype=torch.uint8), gt_flags])
        src = torch.cat([torch.arange(g), keep + g])
    ckeys = keys[src]
    num_expected_pos = int(s['num'] * s['pos_fraction'])             # :83
    pos_inds = torch.nonzero(gt_inds > 0, as_tuple=False).squeeze(1)
    if pos_inds.numel() > num_expected_pos:
        pos_inds = choose_by_keys(pos_inds, num_expected_pos, ckeys)
    pos_inds = pos_inds.unique()                                      # :88
    num_expected_neg = s['num'] - pos_inds.numel()
    if s.get('neg_pos_ub', -1) >= 0:                                  # :91-95
        num_expected_neg = min(num_expected_neg, int(s['neg_pos_ub'] * max(1, pos_inds.numel())))
    neg_inds = torch.nonzero(gt_inds == 0, as_tuple=False).squeeze(1)
    if neg_inds.numel() > num_expected_neg:
        neg_inds = choose_by_keys(neg_inds, num_expected_neg, ckeys)
    neg_inds = neg_inds.unique()
    pos_assigned = gt_inds[pos_inds] - 1
    return SimpleNamespace(
        pos_inds=pos_inds, neg_inds=neg_inds, pos_bboxes=boxes[pos_inds],
        neg_bboxes=boxes[neg_inds], pos_is_gt=gt_flags[pos_inds],
        pos_assigned_gt_inds=pos_assigned,
        pos_gt_bboxes=gt_bboxes.view(-1, 4)[pos_assigned] if g > 0 else gt_bboxes.new_zeros((0, 4)),
        pos_gt_labels=gt_labels[pos_assigned] if g > 0 else gt_labels.new_zeros((0,)),
        bboxes=torch.cat([boxes[pos_inds], boxes[neg_inds]]), gt_inds=gt_inds, max_overlaps=max_ov,
        cand=torch.cat([src[pos_inds], src[neg_inds]]),
        npos_cand=int((gt_inds > 0).sum()), nneg_cand=int((gt_inds == 0).sum()))


def make_sampling(bboxes, num_pos, gt):
    """Synthetic SamplingResult (positives first, sampling_result.py:52-54): the first
    ``num_pos`` rows of ``bboxes`` are positives with the targets in ``gt``."""
    n = min(num_pos, bboxes.size(0))
    return SimpleNamespace(pos_bboxes=bboxes[:n], neg_bboxes=bboxes[n:],
                           pos_gt_bboxes=gt['pos_gt_bboxes'][:n].to(bboxes.dtype),
                           pos_gt_labels=gt['pos_gt_labels'][:n],
                           pos_is_gt=torch.zeros(n, dtype=torch.uint8),
                           bboxes=bboxes)


class HTDRoIHead(nn.Module):
    """htd_roi_head.py:13-386 restricted to the box path of configs/htd/*.py."""

    def __init__(self):
        super().__init__()
        self.bbox_roi_extractor = nn.ModuleList([SingleRoIExtractor(), AdptRoIExtractor()])
        self.bbox_head = nn.ModuleList([Shared2FCBBoxHead(), HTDBBoxHead()])
        self.glbctx_head = GlobalContextHead()
        self.stage_loss_weights = [1, 0.5]

    def _bbox_forward(self, stage, x, rois, global_feat, samp=None):
        """htd_roi_head.py:143-201."""
        ext, enh = self.bbox_roi_extractor
        x4 = x[:ext.num_inputs]
        if stage == 0:
            feats = fuse_global(ext(x4, rois), global_feat, rois)
            cls_score, bbox_pred = self.bbox_head[0](feats)
            return dict(cls_score=cls_score, bbox_pred=bbox_pred, bbox_feats=feats)
        head = self.bbox_head[1]
        if samp is not None:
            pos_rois = bbox2roi([r.pos_bboxes for r in samp])
            feats = ext(x4, rois)
            sef = enh(x4, pos_rois)
            # generalisation of the reference's hard-coded <=2 images (:157-170):
            # positives are the prefix of every image's block
            pos_idx, off = [], 0
            for r in samp:
                pos_idx.append(torch.arange(off, off + r.pos_bboxes.size(0)))
                off += r.bboxes.size(0)
            pos_idx = torch.cat(pos_idx)
            cls_score, bbox_pred = head(feats, feats[pos_idx], x4, rois,
                                        self.bbox_head[0].fc_cls, sef, pos_rois, global_feat)
            full = cls_score.new_zeros(cls_score.size(0), 4)
            full = full.index_put((pos_idx,), bbox_pred)
            return dict(cls_score=cls_score, bbox_pred=full)
        feats = ext(x4, rois)
        sef = enh(x4, rois)
        cls_score, bbox_pred = head(feats, feats, x4, rois, self.bbox_head[0].fc_cls, sef, rois,
                                    global_feat)
        return dict(cls_score=cls_score, bbox_pred=bbox_pred)

    def forward_train_sampled(self, x, proposals, gts, img_shapes, num_pos=128):
        """htd_roi_head.py:217-317 with the assign+sample steps (:254-264, :300-310) replaced
        by the synthetic positives-first sampling of SURVEY §8d."""
        losses = {}
        mc_pred, g = self.glbctx_head(x)
        losses['loss_global'] = self.glbctx_head.loss(mc_pred,
                                                      [gt['gt_labels_unique'] for gt in gts])
        samp = [make_sampling(p, num_pos, gt) for p, gt in zip(proposals, gts)]
        rois = bbox2roi([r.bboxes for r in samp])
        res = self._bbox_forward(0, x, rois, g)
        targets = self.bbox_head[0].get_targets(samp)
        l0 = self.bbox_head[0].loss(res['cls_score'], res['bbox_pred'], rois, *targets)
        for k, v in l0.items():
            losses[f's0.{k}'] = v * self.stage_loss_weights[0] if 'loss' in k else v
        with torch.no_grad():
            roi_labels = torch.where(targets[0] == 80, res['cls_score'][:, :-1].argmax(1),
                                     targets[0])
            refined = self.bbox_head[0].refine_bboxes(rois, roi_labels, res['bbox_pred'],
                                                      [r.pos_is_gt for r in samp], img_shapes)
        samp = [make_sampling(p, num_pos, gt) for p, gt in zip(refined, gts)]
        rois = bbox2roi([r.bboxes for r in samp])
        res = self._bbox_forward(1, x, rois, g, samp)
        targets = self.bbox_head[1].get_targets(samp)
        l1 = self.bbox_head[1].loss(res['cls_score'], res['bbox_pred'], rois, *targets)
        for k, v in l1.items():
            losses[f's1.{k}'] = v * self.stage_loss_weights[1] if 'loss' in k else v
        return losses

    def forward_train_assigned(self, x, proposals, gt_bboxes, gt_labels, keys, cfgs, img_shapes, G):
        """htd_roi_head.py:217-317 INCLUDING the assign + sample steps (:254-264, :300-310), the
        sampler drawing by ``keys`` instead of randperm.  ``keys[stage]`` [B, G + n_stage] is laid
        out like htd_assign_sample's: G gt slots, then one key per stage candidate - for stage 1
        the candidates are the stage-0 rows, of which refine_bboxes (bbox_head.py:227-303) drops
        the gt boxes.  Assignment runs in fp32 like the reference.  Returns (losses, dict of the
        per-stage sampling results and the refined boxes)."""
        dt, dev = x[0].dtype, x[0].device
        losses = {}
        mc_pred, g = self.glbctx_head(x)
        losses['loss_global'] = self.glbctx_head.loss(mc_pred, [l.unique() for l in gt_labels])
        B = len(proposals)

        def sample(stage, boxes, keep):
            out = []
            for b in range(B):
                ng = gt_bboxes[b].size(0)
                k = keys[stage][b]
                pk = k[G:G + keep[b].numel()][keep[b].cpu()] if keep is not None else k[G:]
                add = cfgs[stage]['sampler'].get('add_gt_as_proposals', True) and ng > 0
                kb = torch.cat([k[:ng], pk]) if add else pk
                r = assign_sample_image(boxes[b].detach().float().cpu(), gt_bboxes[b].float().cpu(),
                                        gt_labels[b].cpu(), kb.cpu(), cfgs[stage])
                for f, v in list(vars(r).items()):
                    if torch.is_tensor(v):
                        setattr(r, f, v.to(dev))
                for f in ('pos_bboxes', 'neg_bboxes', 'pos_gt_bboxes', 'bboxes'):
                    setattr(r, f, getattr(r, f).to(dt))
                out.append(r)
            return out

        samp0 = sample(0, proposals, None)
        rois = bbox2roi([r.bboxes for r in samp0])
        res = self._bbox_forward(0, x, rois, g)
        targets = self.bbox_head[0].get_targets(samp0)
        l0 = self.bbox_head[0].loss(res['cls_score'], res['bbox_pred'], rois, *targets)
        for k, v in l0.items():
            losses[f's0.{k}'] = v * self.stage_loss_weights[0] if 'loss' in k else v
        with torch.no_grad():
            roi_labels = torch.where(targets[0] == 80, res['cls_score'][:, :-1].argmax(1),
                                     targets[0])
            refined = self.bbox_head[0].refine_bboxes(rois, roi_labels, res['bbox_pred'],
                                                      [r.pos_is_gt for r in samp0], img_shapes)
            keep = [torch.cat([1 - r.pos_is_gt, r.pos_is_gt.new_ones(r.neg_bboxes.size(0))]).bool()
                    for r in samp0]
        samp1 = sample(1, refined, keep)
        rois = bbox2roi([r.bboxes for r in samp1])
        res = self._bbox_forward(1, x, rois, g, samp1)
        targets = self.bbox_head[1].get_targets(samp1)
        l1 = self.bbox_head[1].loss(res['cls_score'], res['bbox_pred'], rois, *targets)
        for k, v in l1.items():
            losses[f's1.{k}'] = v * self.stage_loss_weights[1] if 'loss' in k else v
        return losses, dict(samp0=samp0, samp1=samp1, refined=refined)

    def simple_test_scores(self, x, proposals, img_shapes):
        """htd_roi_head.py:319-366 up to (and excluding) get_bboxes/NMS: returns the refined
        rois, the stage-averaged cls_score and the stage-1 bbox_pred."""
        rois = bbox2roi(proposals)
        _, g = self.glbctx_head(x)
        r0 = self._bbox_forward(0, x, rois, g)
        n = [len(p) for p in proposals]
        label = r0['cls_score'][:, :-1].argmax(1)
        new_rois = torch.cat([
            self.bbox_head[0].regress_by_class(r, l, p, s)
            for r, l, p, s in zip(rois.split(n), label.split(n), r0['bbox_pred'].split(n),
                                  img_shapes)])
        r1 = self._bbox_forward(1, x, new_rois, g)
        return new_rois, (r0['cls_score'] + r1['cls_score']) / 2.0, r1['bbox_pred']

    def aug_test_merged(self, features, proposals, img_metas):
        """htd_roi_head.py:388-433 with core/bbox/transforms.py:5-56 (bbox_flip / bbox_mapping /
        bbox_mapping_back), bbox_head.py:189-219 (get_bboxes with cfg=None: softmax scores, class
        boxes decoded and clipped, no rescale) and core/post_processing/merge_augs.py:50-76: the
        class boxes of every augmented view mapped back to the original image and averaged,
        scores averaged.  `proposals` is the [n, >=4] tensor of the (one) image, in the original
        image's coordinates."""
        def flip(b, shape, direction):
            out = b.clone()
            if direction in ('horizontal', 'diagonal'):
                out[..., 0::4] = shape[1] - b[..., 2::4]
                out[..., 2::4] = shape[1] - b[..., 0::4]
            if direction in ('vertical', 'diagonal'):
                out[..., 1::4] = shape[0] - b[..., 3::4]
                out[..., 3::4] = shape[0] - b[..., 1::4]
            return out

        boxes, scores = [], []
        for x, meta in zip(features, img_metas):
            m = meta[0]
            sf = proposals.new_tensor(m['scale_factor'])
            direction = m.get('flip_direction', 'horizontal')
            pr = proposals[:, :4] * sf
            if m['flip']:
                pr = flip(pr, m['img_shape'], direction)
            rois, cls_score, bbox_pred = self.simple_test_scores(x, [pr], [m['img_shape']])
            sc = torch.softmax(cls_score, dim=1)
            head = self.bbox_head[-1]
            bx = delta2bbox(rois[:, 1:], bbox_pred, head.target_means, head.target_stds,
                            max_shape=m['img_shape'])
            if m['flip']:
                bx = flip(bx, m['img_shape'], direction)
            boxes.append((bx.view(-1, 4) / sf).view(bx.shape))
            scores.append(sc)
        return torch.stack(boxes).mean(dim=0), torch.stack(scores).mean(dim=0)

