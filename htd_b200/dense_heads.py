"""RPN proposal generation in front of the RoI head (the producer side;
reference: ``mmdet/models/dense_heads/rpn_head.py:77-168`` (``_get_bboxes_single``),
``anchor_head.py:487-583`` (``get_bboxes``), ``mmdet/core/anchor/anchor_generator.py:10-275``, built
by ``configs/htd/htd_resnet50_1x.py:22-37`` with the ``rpn_proposal`` / test ``rpn`` settings of
lines 115-121 / 157-163).

Per image: per level the ``nms_pre`` best anchors by objectness, decoded with the zero-mean /
unit-std DeltaXYWH coder and clipped (``htd_bbox_decode``), then ONE class-aware NMS over the levels
(mmcv ``batched_nms`` with the level index as the class: exactly what ``htd_multiclass_nms`` computes
when candidate row k holds the rank-k box of every level) and the ``nms_post`` best survivors - no
host synchronisation before the final count.  The ranking is ``htd_topk_sorted`` (radix select +
shared-memory bitonic sort, one CTA per level) on the LOGITS: the sigmoid is monotonic, so the order
is the reference's wherever that is defined (equal scores have no defined order there: an unstable
sort; here they come in anchor order).  The anchor-target / loss side of the
RPN is training of another head and is not part of this path.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .core import as_cfg
from .registry import HEADS, build_bbox_coder

MAX_CANDIDATES = 4096          # HTD_NMS_MAX_ROIS: candidates per level


class AnchorGenerator:
    """core/anchor/anchor_generator.py:10-275 for the arguments the configs use (scales given,
    scale_major, centers None)."""

    def __init__(self, strides, ratios, scales=None, base_sizes=None, scale_major=True,
                 octave_base_scale=None, scales_per_octave=None, centers=None, center_offset=0.):
        if centers is not None or octave_base_scale is not None or scales_per_octave is not None \
                or scales is None:
            raise NotImplementedError('configs/htd give `scales` and no centers / octaves')
        self.strides = [(s, s) if not isinstance(s, (tuple, list)) else tuple(s) for s in strides]
        self.base_sizes = [min(s) for s in self.strides] if base_sizes is None else list(base_sizes)
        self.scales, self.ratios = torch.Tensor(scales), torch.Tensor(ratios)
        self.scale_major, self.center_offset = scale_major, center_offset
        self.base_anchors = [self.gen_single_level_base_anchors(b, self.scales, self.ratios)
                             for b in self.base_sizes]
        self._grid = {}

    @property
    def num_base_anchors(self):
        return [b.size(0) for b in self.base_anchors]

    @property
    def num_levels(self):
        return len(self.strides)

    def gen_single_level_base_anchors(self, base_size, scales, ratios, center=None):
        """anchor_generator.py:142-187: float anchors centred on center_offset * base_size."""
        w = h = base_size
        xc, yc = self.center_offset * w, self.center_offset * h
        h_ratios = torch.sqrt(ratios)
        w_ratios = 1 / h_ratios
        if self.scale_major:
            ws = (w * w_ratios[:, None] * scales[None, :]).view(-1)
            hs = (h * h_ratios[:, None] * scales[None, :]).view(-1)
        else:
            ws = (w * scales[:, None] * w_ratios[None, :]).view(-1)
            hs = (h * scales[:, None] * h_ratios[None, :]).view(-1)
        return torch.stack([xc - 0.5 * ws, yc - 0.5 * hs, xc + 0.5 * ws, yc + 0.5 * hs], dim=-1)

    def single_level_grid_anchors(self, base_anchors, featmap_size, stride=(16, 16), device='cuda'):
        """anchor_generator.py:232-272: [H*W*A, 4], rows ordered (y, x, anchor)."""
        feat_h, feat_w = int(featmap_size[0]), int(featmap_size[1])
        sx = torch.arange(0, feat_w, device=device) * stride[0]
        sy = torch.arange(0, feat_h, device=device) * stride[1]
        xx, yy = sx.repeat(feat_h), sy.view(-1, 1).repeat(1, feat_w).view(-1)
        shifts = torch.stack([xx, yy, xx, yy], dim=-1).type_as(base_anchors)
        return (base_anchors[None, :, :] + shifts[:, None, :]).view(-1, 4)

    def grid_anchors(self, featmap_sizes, device='cuda'):
        """Cached per (sizes, device): the anchors depend on the feature-map sizes only."""
        assert self.num_levels == len(featmap_sizes)
        key = (tuple((int(h), int(w)) for h, w in featmap_sizes), str(device))
        if key not in self._grid:
            if len(self._grid) > 16:
                self._grid.clear()
            self._grid[key] = [self.single_level_grid_anchors(self.base_anchors[i].to(device),
                                                              featmap_sizes[i], self.strides[i], device)
                               for i in range(self.num_levels)]
        return self._grid[key]


@HEADS.register_module()
class RPNHead(nn.Module):
    """dense_heads/rpn_head.py: the layers, the forward and the proposal side (get_bboxes)."""

    def __init__(self, in_channels, feat_channels=256, anchor_generator=None, bbox_coder=None,
                 loss_cls=None, loss_bbox=None, train_cfg=None, test_cfg=None, **kwargs):
        super().__init__()
        ag = dict(anchor_generator or dict(type='AnchorGenerator', scales=[8], ratios=[0.5, 1.0, 2.0],
                                           strides=[4, 8, 16, 32, 64]))
        if ag.pop('type', 'AnchorGenerator') != 'AnchorGenerator':
            raise NotImplementedError('configs/htd use AnchorGenerator')
        self.anchor_generator = AnchorGenerator(**ag)
        self.bbox_coder = build_bbox_coder(bbox_coder or dict(
            type='DeltaXYWHBBoxCoder', target_means=[.0, .0, .0, .0], target_stds=[1.0, 1.0, 1.0, 1.0]))
        self.use_sigmoid_cls = True if loss_cls is None else loss_cls.get('use_sigmoid', False)
        if not self.use_sigmoid_cls:
            raise NotImplementedError('configs/htd train the RPN with use_sigmoid=True')
        self.in_channels, self.feat_channels = in_channels, feat_channels
        self.num_anchors = self.anchor_generator.num_base_anchors[0]
        self.cls_out_channels = 1
        self.train_cfg, self.test_cfg = train_cfg, test_cfg
        self.rpn_conv = nn.Conv2d(in_channels, feat_channels, 3, padding=1)
        self.rpn_cls = nn.Conv2d(feat_channels, self.num_anchors * self.cls_out_channels, 1)
        self.rpn_reg = nn.Conv2d(feat_channels, self.num_anchors * 4, 1)
        self.init_weights()

    def init_weights(self):
        for m in (self.rpn_conv, self.rpn_cls, self.rpn_reg):        # rpn_head.py:33-37
            nn.init.normal_(m.weight, 0, 0.01)
            nn.init.constant_(m.bias, 0)

    def forward_single(self, x):
        x = F.relu(self.rpn_conv(x), inplace=True)
        return self.rpn_cls(x), self.rpn_reg(x)

    def forward(self, feats):
        outs = [self.forward_single(x) for x in feats]
        return [o[0] for o in outs], [o[1] for o in outs]

    def loss(self, *a, **k):
        raise NotImplementedError('the anchor-target / loss side of the RPN is outside the accelerated '
                                  'path (SURVEY.md §8): proposals only')

    def _rank(self, cls_scores, cfg):
        """Per level the nms_pre best logits of EVERY image in one launch ([B, nms_pre] values and
        positions), None for a level that is not ranked (rpn_head.py:125)."""
        out = []
        for c in cls_scores:
            B = c.shape[0]
            logit = c.permute(0, 2, 3, 1).reshape(B, -1)
            if cfg.nms_pre > 0 and logit.shape[1] > cfg.nms_pre:
                logit = logit.float().contiguous()
                if cfg.nms_pre <= MAX_CANDIDATES:
                    out.append(ops.topk_sorted(logit, cfg.nms_pre))
                else:
                    ranked, idx = logit.sort(dim=1, descending=True)
                    out.append((ranked[:, :cfg.nms_pre], idx[:, :cfg.nms_pre]))
            else:
                out.append(None)
        return out

    def _get_bboxes_single(self, cls_scores, bbox_preds, mlvl_anchors, img_shape, scale_factor, cfg,
                           rescale=False, ranked=None):
        """rpn_head.py:77-168 for one image.  Returns [n, 5] (x1, y1, x2, y2, score), n <= nms_post,
        in descending score order.  ``ranked``: this image's rows of ``_rank`` (computed here when
        absent)."""
        cfg = as_cfg(self.test_cfg if cfg is None else cfg)
        L = len(cls_scores)
        dev = cls_scores[0].device
        if ranked is None:
            ranked = [None if r is None else (r[0][0], r[1][0])
                      for r in self._rank([c[None] for c in cls_scores], cfg)]
        per_level = []
        for l in range(L):
            delta = bbox_preds[l].permute(1, 2, 0).reshape(-1, 4)
            anchors = mlvl_anchors[l]
            if ranked[l] is not None:
                logit, idx = ranked[l]
                delta, anchors = delta[idx], anchors[idx]
            else:
                logit = cls_scores[l].permute(1, 2, 0).reshape(-1).float()
            if logit.shape[0] > MAX_CANDIDATES:
                raise NotImplementedError(f'{logit.shape[0]} candidates on level {l}: nms_pre <= '
                                          f'{MAX_CANDIDATES} (configs/htd: 2000 / 1000)')
            per_level.append((logit.sigmoid(), delta, anchors))
        K = max(p[0].shape[0] for p in per_level)
        boxes = torch.zeros((K, L, 4), dtype=torch.float32, device=dev)
        scores = torch.zeros((K, L + 1), dtype=torch.float32, device=dev)
        for l, (s, delta, anchors) in enumerate(per_level):
            n = s.shape[0]
            if n == 0:
                continue
            dec = ops.bbox_decode(anchors.float().contiguous(), delta.contiguous(),
                                  self.bbox_coder.means, self.bbox_coder.stds, max_shape=img_shape)
            if cfg.min_bbox_size > 0:                                # rpn_head.py:151-161
                ok = ((dec[:, 2] - dec[:, 0]) >= cfg.min_bbox_size) & \
                     ((dec[:, 3] - dec[:, 1]) >= cfg.min_bbox_size)
                s = s * ok
            boxes[:n, l] = dec
            scores[:n, l] = s
        # candidates = entries with score > 0 (a sigmoid that underflowed to exactly 0 is dropped)
        det, _, count = ops.multiclass_nms(boxes.view(K, L * 4), scores, 0.0, cfg.nms_thr, cfg.nms_post)
        return det[:int(count)]

    def get_bboxes(self, cls_scores, bbox_preds, img_metas, cfg=None, rescale=False, with_nms=True):
        """anchor_head.py:487-583."""
        assert len(cls_scores) == len(bbox_preds) and with_nms
        sizes = [c.shape[-2:] for c in cls_scores]
        anchors = self.anchor_generator.grid_anchors(sizes, device=cls_scores[0].device)
        ranked = self._rank([c.detach() for c in cls_scores], as_cfg(self.test_cfg if cfg is None else cfg))
        out = []
        for i, meta in enumerate(img_metas):
            out.append(self._get_bboxes_single([c[i].detach() for c in cls_scores],
                                               [b[i].detach() for b in bbox_preds], anchors,
                                               meta['img_shape'], meta.get('scale_factor', 1.0), cfg,
                                               rescale, ranked=[None if r is None else (r[0][i], r[1][i])
                                                                for r in ranked]))
        return out

    def simple_test_rpn(self, x, img_metas):
        """rpn_test_mixin.py:25-37."""
        return self.get_bboxes(*self(x), img_metas)
