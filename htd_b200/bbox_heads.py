"""Box heads of the HTD RoI head - host-side mirror of the reference plugin surface.

  BBoxHead            mmdet/models/roi_heads/bbox_heads/bbox_head.py
  Shared2FCBBoxHead   .../convfc_bbox_head.py:176-189 (ConvFCBBoxHead with 2 shared FCs)
  HTDBBoxHead         .../htd_bbox_head.py   (PGraph cls branch + BA/SFA reg branch)
  GlobalContextHead   .../global_context_head.py:323-401 (SFA)

Same constructor arguments, forward signatures and state-dict keys as the reference, so
``configs/htd/*.py`` build them unchanged and released checkpoints load.  The dense library work
(FC stacks, the 3x3 conv tower, GroupNorm) stays in cuBLAS/cuDNN through PyTorch; the graph part
of HTDBBoxHead runs on this package's kernels (``htd_b200.pgraph``), and feature fusion is folded
into the extraction kernels where the caller allows it.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.modules.utils import _pair

from . import dense, ops, pgraph
from .core import accuracy, as_cfg, multiclass_nms
from .registry import HEADS, build_bbox_coder, build_loss


class ConvModule(nn.Module):
    fused_gn = True      # class switch (diagnostics): GN + ReLU through csrc/gn_relu.cu
    """The subset of mmcv.cnn.ConvModule the path uses: conv -> (GN) -> ReLU, bias='auto'
    (no conv bias when a norm layer follows); sub-module names ``conv`` / ``gn`` as in mmcv."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias='auto',
                 conv_cfg=None, norm_cfg=None, act_cfg=dict(type='ReLU')):
        super().__init__()
        if conv_cfg is not None:
            raise NotImplementedError('plain Conv2d only')
        with_norm = norm_cfg is not None
        if bias == 'auto':
            bias = not with_norm
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, bias=bias)
        nn.init.kaiming_normal_(self.conv.weight, a=0, mode='fan_out', nonlinearity='relu')
        if bias:
            nn.init.constant_(self.conv.bias, 0)
        self.gn = None
        if with_norm:
            if norm_cfg['type'] != 'GN':
                raise NotImplementedError('GroupNorm only (htd_bbox_head.py:48)')
            self.gn = nn.GroupNorm(norm_cfg['num_groups'], out_channels)
        self.with_act = act_cfg is not None
        # channels-last weights: cuDNN then runs NHWC end to end instead of wrapping every conv
        # in nchw<->nhwc conversion kernels (63 launches / 0.66 ms per step in the round-1 profile)
        self.conv.to(memory_format=torch.channels_last)

    own_dense = True     # class switch (diagnostics): 3x3 tower convs through csrc/dense_gemm.cu

    def _conv(self, x):
        c = self.conv
        if ConvModule.own_dense and c.bias is None and c.kernel_size == (3, 3) and \
                c.padding == (1, 1) and c.stride == (1, 1) and c.dilation == (1, 1) and \
                c.groups == 1 and x.dim() == 4 and x.shape[2:] == (7, 7) and \
                c.in_channels % 64 == 0 and c.out_channels % 64 == 0 and dense.usable(x, c.weight):
            # implicit-GEMM 3x3 conv on the tcgen05 tensor cores (bf16; halo = TMA zero fill)
            return dense.conv3x3(x, c.weight)
        return c(x)

    def forward(self, x):
        x = self._conv(x)
        if self.gn is not None:
            if self.with_act and x.is_cuda and (x.size(1) // self.gn.num_groups) % 8 == 0 \
                    and ConvModule.fused_gn:
                # GN + ReLU as one kernel (csrc/gn_relu.cu)
                return ops.group_norm_relu(x, self.gn.weight, self.gn.bias, self.gn.num_groups,
                                           self.gn.eps)
            x = self.gn(x)
        return F.relu(x, inplace=True) if self.with_act else x


def _pad8(n):
    return (-n) % 8


def fc(m, x, relu=False):
    """``act(m(x))`` for an nn.Linear: bf16 CUDA tensors run the package's own tcgen05 GEMMs
    (forward with bias / ReLU in the epilogue, data and weight gradient, bias gradient fused with
    the ReLU backward - csrc/dense_gemm.cu); anything else (the fp32 parity configuration) the
    library call."""
    if BBoxHead.own_dense and x.dim() == 2 and dense.usable(x, m.weight):
        return dense.linear(x, m.weight, m.bias, relu)
    y = m(x)
    return F.relu(y) if relu else y


def linear_aligned(m, x):
    """``m(x)`` for an nn.Linear.  cuBLAS has no fast bf16 kernel for an output width that is not
    a multiple of 8 elements (fc_cls: 81, fc_reg: 4 - rows of 162 / 8 bytes; it falls back to
    sm_75-era kernels, 22 us for a 0.17 GFLOP product): the weight is zero-padded to the next
    multiple of 8 rows and the result sliced.  Values are unchanged; fp32 runs the plain call."""
    n = m.out_features
    if BBoxHead.own_dense and x.dim() == 2 and dense.usable(x, m.weight):
        return dense.linear(x, m.weight, m.bias, False)      # any N: the TMA unit pads the tiles
    if not (x.is_cuda and x.dtype == torch.bfloat16 and n % 8 and BBoxHead.aligned_small_fc):
        return m(x)
    w = F.pad(m.weight, (0, 0, 0, _pad8(n)))
    b = F.pad(m.bias, (0, _pad8(n))) if m.bias is not None else None
    return F.linear(x, w, b)[:, :n]


def cls_reg_outputs(head, x):
    """(fc_cls(x), fc_reg(x)) of a head whose two output layers read the same input: ONE GEMM
    over the concatenated (and 8-aligned) weights in bf16, the plain calls otherwise."""
    if not (head.with_cls and head.with_reg and x.is_cuda and x.dtype == torch.bfloat16
            and BBoxHead.aligned_small_fc):
        return (head.fc_cls(x) if head.with_cls else None, head.fc_reg(x) if head.with_reg else None)
    nc, nr = head.fc_cls.out_features, head.fc_reg.out_features
    pad = _pad8(nc + nr)
    w = torch.cat([head.fc_cls.weight, head.fc_reg.weight], 0)
    b = torch.cat([head.fc_cls.bias, head.fc_reg.bias], 0)
    if BBoxHead.own_dense and x.dim() == 2 and dense.usable(x, w):
        out = dense.linear(x, w, b, False)
        return out[:, :nc], out[:, nc:nc + nr]
    if pad:
        w, b = F.pad(w, (0, 0, 0, pad)), F.pad(b, (0, pad))
    out = F.linear(x, w, b)
    return out[:, :nc], out[:, nc:nc + nr]


@HEADS.register_module()
class BBoxHead(nn.Module):
    """bbox_head.py:12-335."""

    def __init__(self, with_avg_pool=False, with_cls=True, with_reg=True, roi_feat_size=7,
                 in_channels=256, num_classes=80,
                 bbox_coder=dict(type='DeltaXYWHBBoxCoder', clip_border=True,
                                 target_means=[0., 0., 0., 0.], target_stds=[0.1, 0.1, 0.2, 0.2]),
                 reg_class_agnostic=False, reg_decoded_bbox=False,
                 loss_cls=dict(type='CrossEntropyLoss', use_sigmoid=False, loss_weight=1.0),
                 loss_bbox=dict(type='SmoothL1Loss', beta=1.0, loss_weight=1.0)):
        super().__init__()
        assert with_cls or with_reg
        self.with_avg_pool, self.with_cls, self.with_reg = with_avg_pool, with_cls, with_reg
        self.roi_feat_size = _pair(roi_feat_size)
        self.roi_feat_area = self.roi_feat_size[0] * self.roi_feat_size[1]
        self.in_channels, self.num_classes = in_channels, num_classes
        self.reg_class_agnostic, self.reg_decoded_bbox = reg_class_agnostic, reg_decoded_bbox
        self.fp16_enabled = False
        self.bbox_coder = build_bbox_coder(bbox_coder)
        self.loss_cls = build_loss(loss_cls)
        self.loss_bbox = build_loss(loss_bbox)
        in_ch = in_channels
        if with_avg_pool:
            self.avg_pool = nn.AvgPool2d(self.roi_feat_size)
        else:
            in_ch *= self.roi_feat_area
        if with_cls:
            self.fc_cls = nn.Linear(in_ch, num_classes + 1)
        if with_reg:
            self.fc_reg = nn.Linear(in_ch, 4 if reg_class_agnostic else 4 * num_classes)

    def init_weights(self):
        if self.with_cls:
            nn.init.normal_(self.fc_cls.weight, 0, 0.01)
            nn.init.constant_(self.fc_cls.bias, 0)
        if self.with_reg:
            nn.init.normal_(self.fc_reg.weight, 0, 0.001)
            nn.init.constant_(self.fc_reg.bias, 0)

    def forward(self, x):
        if self.with_avg_pool:
            x = self.avg_pool(x)
        x = x.reshape(x.size(0), -1)
        return cls_reg_outputs(self, x)

    # ---- targets / loss (bbox_head.py:85-186) -------------------------------------------------
    def _get_target_single(self, pos_bboxes, neg_bboxes, pos_gt_bboxes, pos_gt_labels, cfg):
        num_pos, num_neg = pos_bboxes.size(0), neg_bboxes.size(0)
        n = num_pos + num_neg
        labels = pos_bboxes.new_full((n,), self.num_classes, dtype=torch.long)
        label_weights = pos_bboxes.new_zeros(n)
        bbox_targets = pos_bboxes.new_zeros(n, 4)
        bbox_weights = pos_bboxes.new_zeros(n, 4)
        if num_pos > 0:
            labels[:num_pos] = pos_gt_labels
            pw = cfg.get('pos_weight', -1) if cfg is not None else -1
            label_weights[:num_pos] = 1.0 if pw <= 0 else pw
            bbox_targets[:num_pos] = pos_gt_bboxes if self.reg_decoded_bbox else \
                self.bbox_coder.encode(pos_bboxes, pos_gt_bboxes).to(bbox_targets.dtype)
            bbox_weights[:num_pos] = 1
        if num_neg > 0:
            label_weights[-num_neg:] = 1.0
        return labels, label_weights, bbox_targets, bbox_weights

    def get_targets(self, sampling_results, gt_bboxes, gt_labels, rcnn_train_cfg, concat=True):
        cfg = as_cfg(rcnn_train_cfg)
        if concat and self.fused_glue and not self.reg_decoded_bbox and sampling_results and \
                sampling_results[0].pos_bboxes.is_cuda:
            return self._get_targets_fused(sampling_results, cfg)
        sizes = {(r.pos_bboxes.size(0), r.neg_bboxes.size(0)) for r in sampling_results}
        if concat and len(sizes) == 1 and len(sampling_results) > 1 and min(next(iter(sizes))) > 0 \
                and not self.reg_decoded_bbox:
            # same (positives, negatives) in every image: one batched pass instead of a per-image
            # loop (identical values; the op count no longer grows with images per GPU)
            npos, nneg = next(iter(sizes))
            nb, n = len(sampling_results), npos + nneg
            pos = torch.stack([r.pos_bboxes for r in sampling_results])
            gtb = torch.stack([r.pos_gt_bboxes for r in sampling_results])
            lab = torch.stack([r.pos_gt_labels for r in sampling_results])
            labels = pos.new_full((nb, n), self.num_classes, dtype=torch.long)
            labels[:, :npos] = lab
            pw = cfg.get('pos_weight', -1) if cfg is not None else -1
            label_weights = pos.new_ones((nb, n))
            if pw > 0:
                label_weights[:, :npos] = pw
            bbox_targets = pos.new_zeros((nb, n, 4))
            bbox_targets[:, :npos] = self.bbox_coder.encode(
                pos.reshape(-1, 4), gtb.reshape(-1, 4)).to(pos.dtype).view(nb, npos, 4)
            bbox_weights = pos.new_zeros((nb, n, 4))
            bbox_weights[:, :npos] = 1
            return (labels.view(-1), label_weights.view(-1), bbox_targets.view(-1, 4),
                    bbox_weights.view(-1, 4))
        outs = [self._get_target_single(r.pos_bboxes, r.neg_bboxes, r.pos_gt_bboxes,
                                        r.pos_gt_labels, cfg) for r in sampling_results]
        cols = list(zip(*outs))
        return tuple(torch.cat(c, 0) for c in cols) if concat else tuple(list(c) for c in cols)

    def _get_targets_fused(self, sampling_results, cfg):
        """All images in ONE htd_bbox_targets launch: the gt of the positives is laid out in the
        sampled-RoI order (positives are the prefix of each image's block)."""
        dev = sampling_results[0].pos_bboxes.device
        sizes = tuple((r.pos_bboxes.size(0), r.neg_bboxes.size(0)) for r in sampling_results)
        key = (sizes, str(dev))
        cache = BBoxHead._pos_mask_cache
        is_pos = cache.get(key)
        if is_pos is None:
            # depends on the counts only.  Fixed-count protocols (bench, CUDA graphs) hit the cache
            # every step; with the reference's random sampler the counts change almost every
            # iteration, so the cache is bounded (least recently used entry dropped)
            m = torch.cat([torch.cat([torch.ones(p, dtype=torch.uint8), torch.zeros(n, dtype=torch.uint8)])
                           for p, n in sizes])
            is_pos = cache[key] = m.to(dev)
            while len(cache) > BBoxHead._pos_mask_cache_max:
                cache.pop(next(iter(cache)))
        else:
            cache[key] = cache.pop(key)         # most recently used last
        boxes = torch.cat([t for r in sampling_results for t in (r.pos_bboxes, r.neg_bboxes)], 0)
        # gt of the positives in the sampled-RoI order; the rows of the negatives are never read
        # (is_pos == 0), so a cached zero block fills them: one cat per tensor instead of a fill and
        # a slice copy per image
        zb, zl = BBoxHead._zero_rows(max(n for _, n in sizes), dev)
        gtb = torch.cat([t for r, (p, n) in zip(sampling_results, sizes)
                         for t in (r.pos_gt_bboxes.float(), zb[:n])], 0)
        gtl = torch.cat([t for r, (p, n) in zip(sampling_results, sizes)
                         for t in (r.pos_gt_labels.long(), zl[:n])], 0)
        pw = cfg.get('pos_weight', -1) if cfg is not None else -1
        return ops.bbox_targets(boxes, gtb, gtl, is_pos, self.num_classes, pw, self.bbox_coder.means,
                                self.bbox_coder.stds)

    def loss(self, cls_score, bbox_pred, rois, labels, label_weights, bbox_targets, bbox_weights,
             reduction_override=None, pad_rows=False):
        """bbox_head.py:141-186 in static-shape form: the reference selects the positive rows
        with boolean masks (`.any()`, `.item()` host syncs, data-dependent shapes); here the same
        sums are taken over ALL rows with the positive mask as a 0/1 factor and `avg_factor` kept
        on the device - identical values, no host sync, capturable in a CUDA graph."""
        losses = dict()
        if self._fused_loss_ok(cls_score, bbox_pred, reduction_override):
            # one kernel pair for CE + accuracy + smooth-L1 and their gradients (csrc/rcnn_glue.cu)
            loss_cls, acc, loss_bbox = ops.rcnn_loss(
                cls_score, bbox_pred, labels, label_weights, bbox_targets, bbox_weights,
                self.num_classes, self.loss_bbox.beta, self.loss_cls.loss_weight,
                self.loss_bbox.loss_weight, pad_rows=pad_rows)
            return dict(loss_cls=loss_cls, acc=acc, loss_bbox=loss_bbox)
        if pad_rows:
            raise NotImplementedError('pad rows (static-shape sampling) need the fused loss kernels')
        if cls_score is not None:
            cls_score = cls_score.float()                      # force_fp32 (bbox_head.py:141)
            avg_factor = torch.sum(label_weights > 0).float().clamp(min=1.)
            if cls_score.numel() > 0:
                losses['loss_cls'] = self.loss_cls(cls_score, labels, label_weights,
                                                   avg_factor=avg_factor,
                                                   reduction_override=reduction_override)
                losses['acc'] = accuracy(cls_score, labels)
        if bbox_pred is not None:
            bbox_pred = bbox_pred.float()
            pos = ((labels >= 0) & (labels < self.num_classes)).to(bbox_pred.dtype)
            if self.reg_decoded_bbox:
                bbox_pred = self.bbox_coder.decode(rois[:, 1:], bbox_pred)
            if self.reg_class_agnostic:
                pred = bbox_pred.view(bbox_pred.size(0), 4)
            else:
                idx = labels.clamp(0, self.num_classes - 1)
                pred = bbox_pred.view(bbox_pred.size(0), -1, 4)[
                    torch.arange(bbox_pred.size(0), device=bbox_pred.device), idx]
            losses['loss_bbox'] = self.loss_bbox(pred, bbox_targets.float(),
                                                 bbox_weights.float() * pos[:, None],
                                                 avg_factor=bbox_targets.size(0),
                                                 reduction_override=reduction_override)
        return losses

    def _fused_loss_ok(self, cls_score, bbox_pred, reduction_override):
        from .core import CrossEntropyLoss, SmoothL1Loss
        return (cls_score is not None and bbox_pred is not None and cls_score.is_cuda
                and cls_score.numel() > 0 and self.reg_class_agnostic and not self.reg_decoded_bbox
                and reduction_override is None and bbox_pred.size(-1) == 4
                and type(self.loss_cls) is CrossEntropyLoss and type(self.loss_bbox) is SmoothL1Loss
                and self.loss_cls.reduction == 'mean' and self.loss_bbox.reduction == 'mean'
                and self.loss_cls.class_weight is None and self.fused_glue)

    own_dense = True         # class switch (diagnostics): FC layers through csrc/dense_gemm.cu (bf16)
    fused_glue = True        # class switch (diagnostics): targets / loss / decode via csrc/rcnn_glue.cu
    aligned_small_fc = True  # class switch (diagnostics): 8-aligned fc_cls / fc_reg GEMMs in bf16
    _pos_mask_cache = {}
    _pos_mask_cache_max = 16
    _zero_cache = {}

    @staticmethod
    def _zero_rows(n, dev):
        z = BBoxHead._zero_cache.get(str(dev))
        if z is None or z[0].size(0) < n:
            m = max(n, 2048)
            z = BBoxHead._zero_cache[str(dev)] = (torch.zeros((m, 4), dtype=torch.float32, device=dev),
                                                  torch.zeros(m, dtype=torch.long, device=dev))
        return z

    # ---- decoding (bbox_head.py:188-335) ------------------------------------------------------
    def get_bboxes(self, rois, cls_score, bbox_pred, img_shape, scale_factor, rescale=False,
                   cfg=None):
        if isinstance(cls_score, list):
            cls_score = sum(cls_score) / float(len(cls_score))
        scores = F.softmax(cls_score.float(), dim=1) if cls_score is not None else None
        if bbox_pred is not None:
            bboxes = self.bbox_coder.decode(rois[:, 1:].float(), bbox_pred.float(),
                                            max_shape=img_shape)
        else:
            bboxes = rois[:, 1:].clone()
            if img_shape is not None:
                bboxes[:, [0, 2]] = bboxes[:, [0, 2]].clamp(min=0, max=img_shape[1])
                bboxes[:, [1, 3]] = bboxes[:, [1, 3]].clamp(min=0, max=img_shape[0])
        if rescale and bboxes.size(0) > 0:
            if isinstance(scale_factor, float):
                bboxes = bboxes / scale_factor
            else:
                sf = bboxes.new_tensor(scale_factor)
                bboxes = (bboxes.view(bboxes.size(0), -1, 4) / sf).view(bboxes.size(0), -1)
        if cfg is None:
            return bboxes, scores
        cfg = as_cfg(cfg)
        return multiclass_nms(bboxes, scores, cfg.score_thr, cfg.nms, cfg.max_per_img)

    def regress_by_class(self, rois, label, bbox_pred, img_meta):
        assert rois.size(1) in (4, 5), repr(rois.shape)
        bbox_pred = bbox_pred.float()
        if not self.reg_class_agnostic:
            label = label * 4
            inds = torch.stack((label, label + 1, label + 2, label + 3), 1)
            bbox_pred = torch.gather(bbox_pred, 1, inds)
        assert bbox_pred.size(1) == 4
        if self.fused_glue and rois.is_cuda and getattr(self.bbox_coder, 'clip_border', True):
            return ops.bbox_decode(rois, bbox_pred, self.bbox_coder.means, self.bbox_coder.stds,
                                   max_shape=img_meta['img_shape'])
        if rois.size(1) == 4:
            return self.bbox_coder.decode(rois, bbox_pred, max_shape=img_meta['img_shape'])
        boxes = self.bbox_coder.decode(rois[:, 1:], bbox_pred, max_shape=img_meta['img_shape'])
        return torch.cat((rois[:, [0]], boxes), dim=1)

    def refine_bboxes(self, rois, labels, bbox_preds, pos_is_gts, img_metas, num_per_img=None):
        """bbox_head.py:227-303.  ``num_per_img`` (RoIs of every image, in order) replaces the
        reference's per-image ``nonzero`` (host sync); ``pos_is_gts[i] is None`` means "no sampled
        box of image i is a ground-truth box" and skips the boolean filter (static shapes)."""
        shapes = {tuple(m['img_shape'][:2]) for m in img_metas}
        if num_per_img is not None and len(shapes) == 1 and all(g is None for g in pos_is_gts) \
                and sum(num_per_img) == rois.size(0):
            # one decode for the whole batch (same image shape, nothing to filter)
            boxes = self.regress_by_class(rois[:, 1:], labels, bbox_preds, img_metas[0])
            return list(boxes.split(list(num_per_img), 0))
        out, off = [], 0
        for i in range(len(img_metas)):
            if num_per_img is not None:
                inds = slice(off, off + num_per_img[i])
                off += num_per_img[i]
                n = num_per_img[i]
            else:
                inds = torch.nonzero(rois[:, 0] == i, as_tuple=False).squeeze(dim=1)
                n = inds.numel()
            boxes = self.regress_by_class(rois[inds, 1:], labels[inds], bbox_preds[inds],
                                          img_metas[i])
            if pos_is_gts[i] is not None:
                keep = pos_is_gts[i].new_ones(n)
                keep[:len(pos_is_gts[i])] = 1 - pos_is_gts[i]
                boxes = boxes[keep.type(torch.bool)]
            out.append(boxes)
        return out


@HEADS.register_module()
class ConvFCBBoxHead(BBoxHead):
    """convfc_bbox_head.py:9-173 restricted to what HTD instantiates: shared FCs only."""

    def __init__(self, num_shared_convs=0, num_shared_fcs=0, num_cls_convs=0, num_cls_fcs=0,
                 num_reg_convs=0, num_reg_fcs=0, conv_out_channels=256, fc_out_channels=1024,
                 conv_cfg=None, norm_cfg=None, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if num_shared_convs or num_cls_convs or num_cls_fcs or num_reg_convs or num_reg_fcs:
            raise NotImplementedError('HTD stage 0 is Shared2FCBBoxHead (shared FCs only)')
        assert num_shared_fcs > 0 and not self.with_avg_pool
        self.num_shared_fcs, self.fc_out_channels = num_shared_fcs, fc_out_channels
        self.shared_fcs = nn.ModuleList()
        last = self.in_channels * self.roi_feat_area
        for _ in range(num_shared_fcs):
            self.shared_fcs.append(nn.Linear(last, fc_out_channels))
            last = fc_out_channels
        if self.with_cls:
            self.fc_cls = nn.Linear(last, self.num_classes + 1)
        if self.with_reg:
            self.fc_reg = nn.Linear(last, 4 if self.reg_class_agnostic else 4 * self.num_classes)

    def init_weights(self):
        super().init_weights()
        for m in self.shared_fcs:
            nn.init.xavier_uniform_(m.weight)
            nn.init.constant_(m.bias, 0)

    def forward(self, x, sfa_bias=None, rois=None):
        """``sfa_bias`` ([B,C,1,1]) + ``rois`` (optional): the global-context vector the reference
        adds to the RoI features before this head (``_fuse_global``, htd_roi_head.py:133-141),
        added here while the features are flattened - the extraction that produced ``x`` then
        does not depend on the global-context head."""
        x = ops.flatten_roi_feats(x, sfa_bias, rois)
        for m in self.shared_fcs:
            x = fc(m, x, relu=True)
        return cls_reg_outputs(self, x)


@HEADS.register_module()
class Shared2FCBBoxHead(ConvFCBBoxHead):

    def __init__(self, fc_out_channels=1024, *args, **kwargs):
        super().__init__(num_shared_fcs=2, fc_out_channels=fc_out_channels, *args, **kwargs)


@HEADS.register_module()
class HTDBBoxHead(BBoxHead):
    """htd_bbox_head.py:24-230.  cls branch: 2 FCs + PGraph; reg branch: RoI feature + SFA + BA
    feature -> 4 convs (GN-36) -> avg-pool -> fc_reg."""

    def __init__(self, num_shared_convs=0, num_shared_fcs=0, num_cls_convs=0, num_cls_fcs=2,
                 num_reg_convs=4, num_reg_fcs=0, alpha=1, relpace=False, average=False, edge=1,
                 conv_out_channels=256, fc_out_channels=1024, conv_cfg=None,
                 norm_cfg=dict(type='GN', num_groups=36), *args, **kwargs):
        kwargs.setdefault('with_avg_pool', True)
        super().__init__(*args, **kwargs)
        if relpace or average:
            raise NotImplementedError('configs/htd use relpace=False, average=False')
        self.num_cls_fcs, self.num_reg_convs = num_cls_fcs, num_reg_convs
        self.alpha, self.relpace, self.average, self.edge = alpha, relpace, average, edge
        self.conv_out_channels = self.fc_out_channels = 1024
        self.gcn_in = self.gcn_out = 1024
        self.norm_cfg = norm_cfg
        self.fc_cls = nn.Linear(self.fc_out_channels, self.num_classes + 1)
        self.fc_reg = nn.Linear(self.conv_out_channels, 4)
        mid = self.middle_channel = 16 * 36
        convs = []
        for i in range(num_reg_convs):
            cin = self.in_channels if i == 0 else mid
            last = (i == num_reg_convs - 1)
            convs.append(ConvModule(cin, 1024 if last else mid, 3, padding=1, conv_cfg=conv_cfg,
                                    norm_cfg=None if last else norm_cfg, bias=False))
        self.convs = nn.Sequential(*convs)
        fcs = []
        for i in range(num_cls_fcs):
            fcs += [nn.Linear(self.in_channels * self.roi_feat_area if i == 0
                              else self.fc_out_channels, self.fc_out_channels), nn.ReLU(inplace=True)]
        self.fcs = nn.Sequential(*fcs)
        self.avg_pool = nn.AvgPool2d(self.roi_feat_size)
        for i in range(4):
            setattr(self, f'graph_lvl{i}_cls', nn.Linear(self.gcn_in, self.gcn_out))
        self.finest_scale = 56
        self.last_plan = None            # GraphPlan of the latest forward (inspection / tests)

    @property
    def graph_layer_cls(self):
        return [getattr(self, f'graph_lvl{i}_cls') for i in range(4)]

    def init_weights(self):
        super().init_weights()
        nn.init.normal_(self.fc_cls.weight, 0, 0.01)
        nn.init.normal_(self.fc_reg.weight, 0, 0.001)
        for m in list(self.fcs.modules()) + self.graph_layer_cls:
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                nn.init.constant_(m.bias, 0)

    def map_roi_levels(self, rois, num_levels):
        """htd_bbox_head.py:129-135 (bit-exact level index, int64 like the reference)."""
        return ops.level_assign(rois, num_levels, self.finest_scale).long()

    @staticmethod
    def _img_onehot(rois, num_imgs, dtype):
        """[K, B] one-hot of the RoIs' image index: per-image vectors are broadcast to RoIs as
        ``onehot @ v`` - a tiny GEMM whose backward is a GEMM too, instead of advanced indexing
        whose backward (indexing_backward_kernel) cost 0.75 ms per step in the round-1 profile."""
        b = torch.arange(num_imgs, device=rois.device, dtype=rois.dtype)
        return (rois[:, :1] == b[None, :]).to(dtype)

    def forward(self, x_cls, x_reg, feat, rois, fc_cls_0, enhanced_feat=None, pos_rois=None,
                global_feat=None, num_imgs=None, max_rois_per_img=None, row_valid=None,
                x_cls_flat=None, reg_stream=None):
        """Reference signature (htd_bbox_head.py:157).  ``num_imgs`` (optional) avoids the host
        sync of ``int(max(rois[:,0])) + 1`` (:159); ``global_feat`` gives it otherwise.
        ``max_rois_per_img`` (optional) bounds the PGraph group size (default: all RoIs);
        ``row_valid`` ([K] bool, optional) keeps pad rows of the static sampler out of the graph;
        ``x_cls_flat`` (optional) is ``x_cls`` already flattened (``ops.flatten_with_prefix``);
        ``reg_stream`` (optional CUDA stream, the one ``enhanced_feat`` was produced on): the reg
        branch - independent of the cls branch until the loss - runs there, as a parallel branch
        of the step."""
        if num_imgs is None:
            num_imgs = global_feat.size(0) if global_feat is not None \
                else int(torch.max(rois[..., 0])) + 1
        d = self.fc_out_channels
        prototype = torch.cat((fc_cls_0.weight, fc_cls_0.bias.unsqueeze(1)), 1).detach()
        g = global_feat.reshape(global_feat.size(0), -1) if global_feat is not None else None
        # ---- reg branch: (x_reg + SFA) + alpha * BA  -> conv tower -> avg pool -> fc_reg
        # (:161-189).  Independent of the cls branch: with ``reg_stream`` it is a parallel branch
        def reg_branch(x_reg, enhanced_feat):
            if callable(enhanced_feat):
                enhanced_feat = enhanced_feat()
            if BBoxHead.own_dense and dense.usable(x_reg, enhanced_feat) and x_reg.size(1) % 8 == 0 \
                    and (g is None or g.dtype == x_reg.dtype):
                # x_reg + global_feat[image] + alpha * BA in ONE pass (csrc/dense_gemm.cu
                # add3_kernel) instead of a one-hot GEMM, three elementwise launches and their
                # backward ops
                x_reg = dense.add3(x_reg, enhanced_feat, g, pos_rois, self.alpha)
            else:
                if global_feat is not None:
                    x_reg = x_reg + (self._img_onehot(pos_rois, g.size(0), g.dtype) @ g)[:, :, None, None]
                x_reg = x_reg + self.alpha * enhanced_feat
                x_reg = x_reg.contiguous(memory_format=torch.channels_last)
            last = self.convs[-1] if len(self.convs) else None
            if last is not None and last.gn is None and last.with_act and x_reg.is_cuda and \
                    ConvModule.fused_gn and last.conv.out_channels % 8 == 0:
                for m in self.convs[:-1]:
                    x_reg = m(x_reg)
                # last conv -> ReLU -> AvgPool2d(7) on a 7x7 map (htd_bbox_head.py:109-113,188-189):
                # activation and pool in one pass over the largest activation of the head
                x_reg = ops.relu_mean_pool(last._conv(x_reg))
            else:
                x_reg = self.convs(x_reg).mean((2, 3))
            return linear_aligned(self.fc_reg, x_reg) if self.with_reg else None

        bbox_pred = None
        if reg_stream is not None:
            cur = torch.cuda.current_stream()
            reg_stream.wait_stream(cur)                      # x_reg, g, pos_rois come from `cur`
            for t in (x_reg, g, pos_rois):
                if torch.is_tensor(t):
                    t.record_stream(reg_stream)
            with torch.cuda.stream(reg_stream):
                bbox_pred = reg_branch(x_reg, enhanced_feat)
        # ---- cls branch: fcs on x_cls and on x_cls + SFA.  fcs.0 is linear, so
        # fcs.0(x + g (x) 1_49) = fcs.0(x) + g W_sum^T with W_sum = sum of W over the 49 bins:
        # one [K,12544]x[12544,1024] GEMM instead of the reference's two (:164 and :192).
        fc0, fc1 = self.fcs[0], self.fcs[2]
        x_flat = x_cls_flat if x_cls_flat is not None else ops.flatten_roi_feats(x_cls)
        x_glb = None
        if BBoxHead.own_dense and global_feat is not None and dense.usable(x_flat, fc0.weight):
            # both FC inputs of the reference from ONE product: the epilogue of the fcs.0 GEMM
            # writes relu(v) and relu(v + corr[image of the RoI]); fcs.2 then runs once on the
            # 2K stacked rows
            g49 = g.to(fc0.weight.dtype).repeat_interleave(self.roi_feat_area, dim=1)
            corr = dense.linear(g49, fc0.weight, None, False)
            h = dense.linear_dual(x_flat, fc0.weight, fc0.bias, corr, rois[:, 0])
            both = fc(fc1, h, relu=True)
            K_ = x_flat.size(0)
            x_c, x_glb = both[:K_], both[K_:]
        else:
            pre = fc(fc0, x_flat)
            x_c = fc(fc1, F.relu(pre), relu=True)
            if global_feat is not None:
                # g (x) 1_49 through fcs.0's weight: a [B, 12544] x [12544, 1024] GEMM that reads W
                # once (summing W over the 49 bins first is a 25 MB strided reduction)
                g49 = g.to(fc0.weight.dtype).repeat_interleave(self.roi_feat_area, dim=1)
                corr = F.linear(g49, fc0.weight)
                x_glb = fc(fc1, F.relu(pre + self._img_onehot(rois, g.size(0), corr.dtype) @ corr),
                           relu=True)
        # ---- semantic vectors and the graph (:194-219)
        probs = linear_aligned(fc_cls_0, x_c).softmax(-1)
        if BBoxHead.own_dense and dense.usable(probs, prototype):
            sam = dense.mm(probs, prototype)
        else:
            sam = torch.mm(probs, prototype)
        with torch.no_grad():
            levels = ops.level_assign(rois, len(feat), self.finest_scale)
            if row_valid is not None:
                levels = torch.where(row_valid, levels, torch.full_like(levels, -1))
            plan = pgraph.GraphPlan(rois, levels, num_imgs, len(feat), x_c.dtype,
                                    max_group=max_rois_per_img, d=d, ds=prototype.size(1))
        self.last_plan = plan
        layers = self.graph_layer_cls
        refined = pgraph.pgraph_refine(x_c, sam, [m.weight for m in layers],
                                       [m.bias for m in layers], plan)
        feat_cls_new = (x_glb if x_glb is not None else x_c) + refined
        cls_score = linear_aligned(self.fc_cls, feat_cls_new) if self.with_cls else None
        if reg_stream is None:
            bbox_pred = reg_branch(x_reg, enhanced_feat)
        else:
            cur.wait_stream(reg_stream)
            if bbox_pred is not None:
                bbox_pred.record_stream(cur)
        return cls_score, bbox_pred


@HEADS.register_module()
class GlobalContextHead(nn.Module):
    """SFA, global_context_head.py:323-401: convs on the LAST pyramid level -> global average
    pool -> [B,C,1,1] context vector (+ multi-label logits and BCE loss)."""

    def __init__(self, num_ins, num_convs=4, in_channels=256, conv_out_channels=256,
                 num_classes=81, loss_weight=1.0, conv_cfg=None, norm_cfg=None, conv_to_res=False):
        super().__init__()
        if conv_to_res:
            raise NotImplementedError('HTD builds GlobalContextHead with conv_to_res=False')
        self.num_ins, self.num_convs = num_ins, num_convs
        self.in_channels, self.conv_out_channels = in_channels, conv_out_channels
        self.num_classes, self.loss_weight = num_classes, loss_weight
        self.fp16_enabled = False
        self.convs = nn.ModuleList([
            ConvModule(in_channels if i == 0 else conv_out_channels, conv_out_channels, 3,
                       padding=1, conv_cfg=conv_cfg, norm_cfg=norm_cfg) for i in range(num_convs)])
        self.pool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Linear(conv_out_channels, num_classes)
        self.criterion = nn.BCEWithLogitsLoss()

    def init_weights(self):
        nn.init.normal_(self.fc.weight, 0, 0.01)
        nn.init.constant_(self.fc.bias, 0)

    def forward(self, feats):
        x = feats[-1]
        w = self.convs[0].conv.weight
        if x.dtype != w.dtype:
            x = x.to(w.dtype)
        for conv in self.convs:
            x = conv(x)
        x = self.pool(x)
        return self.fc(x.reshape(x.size(0), -1)), x

    def loss_multihot(self, pred, multihot):
        """``loss`` with the multi-hot target given ([B,num_classes] bool; static shapes)."""
        return self.loss_weight * self.criterion(pred.float(), multihot.float())

    def loss(self, pred, labels):
        pred = pred.float()
        targets = pred.new_zeros(pred.size())
        for i, label in enumerate(labels):          # multi-hot; duplicates are harmless, so the
            if label.numel():                       # reference's `.unique()` (host sync) is not needed
                targets[i].index_fill_(0, label.long(), 1.0)
        return self.loss_weight * self.criterion(pred, targets)
