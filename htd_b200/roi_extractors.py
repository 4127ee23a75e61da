"""RoI feature extraction of the HTD head: host-side mirror of the reference plugin surface.

Same class names, constructor arguments, attributes and forward signatures as
  mmcv.ops.RoIAlign                      (constructed at roi_extractors/base_roi_extractor.py:49-55)
  BaseRoIExtractor                       mmdet/models/roi_heads/roi_extractors/base_roi_extractor.py
  SingleRoIExtractor                     .../single_level_roi_extractor.py
  AdptRoIExtractor (BA)                  .../adaptative_roi_extractor.py
so that ``configs/htd/htd_resnet50_1x.py:43-55`` builds them unchanged and released checkpoints
load (state-dict keys ``conv1/conv2`` and the aliases ``att.1/att.3``).  All arithmetic is done by
the sm_100a kernels behind ``htd_b200.ops``; there is no CPU path.
"""
import torch
import torch.nn as nn
from torch.nn.modules.utils import _pair

from . import ops
from .registry import ROI_EXTRACTORS, ROI_LAYERS


@ROI_LAYERS.register_module()
class RoIAlign(nn.Module):
    """mmcv.ops.RoIAlign signature.  Only the configuration the reference uses is implemented in
    CUDA: pool_mode='avg', aligned=True (sampling_ratio 0 = adaptive, or > 0)."""

    def __init__(self, output_size, spatial_scale=1.0, sampling_ratio=0, pool_mode='avg',
                 aligned=True, use_torchvision=False):
        super().__init__()
        self.output_size = _pair(output_size)
        self.spatial_scale = float(spatial_scale)
        self.sampling_ratio = int(sampling_ratio)
        self.pool_mode = pool_mode
        self.aligned = aligned
        self.use_torchvision = use_torchvision
        if pool_mode != 'avg' or not aligned or use_torchvision:
            raise NotImplementedError(
                'htd_b200.RoIAlign implements pool_mode="avg", aligned=True (the HTD configuration)')
        if self.output_size[0] != self.output_size[1] or self.output_size[0] > 8:
            raise NotImplementedError('square output_size <= 8 only')

    def forward(self, input, rois):
        """input [B,C,H,W], rois [k,5] -> [k,C,oh,ow] (contiguous NCHW, like mmcv)."""
        x = ops.to_channels_last(input)
        out = ops.roi_align_levels([x], rois, [self.spatial_scale], self.output_size[0],
                                   self.sampling_ratio)
        return out[0].contiguous()

    def __repr__(self):
        return (f'{self.__class__.__name__}(output_size={self.output_size}, '
                f'spatial_scale={self.spatial_scale}, sampling_ratio={self.sampling_ratio}, '
                f'pool_mode={self.pool_mode}, aligned={self.aligned})')


class BaseRoIExtractor(nn.Module):
    """base_roi_extractor.py:9-84."""

    def __init__(self, roi_layer, out_channels, featmap_strides):
        super().__init__()
        self.roi_layers = self.build_roi_layers(roi_layer, featmap_strides)
        self.out_channels = out_channels
        self.featmap_strides = featmap_strides
        self.fp16_enabled = False
        self.compute_dtype = None        # None: keep the dtype of the incoming feature maps

    @property
    def num_inputs(self):
        return len(self.featmap_strides)

    def init_weights(self):
        pass

    def build_roi_layers(self, layer_cfg, featmap_strides):
        cfg = layer_cfg.copy()
        layer_type = cfg.pop('type')
        layer_cls = ROI_LAYERS.get(layer_type)
        if layer_cls is None:
            raise KeyError(f'RoI layer {layer_type!r} is not provided by htd_b200 '
                           f'(available: {sorted(ROI_LAYERS.module_dict)})')
        return nn.ModuleList([layer_cls(spatial_scale=1 / s, **cfg) for s in featmap_strides])

    def roi_rescale(self, rois, scale_factor):
        """base_roi_extractor.py:58-80."""
        cx = (rois[:, 1] + rois[:, 3]) * 0.5
        cy = (rois[:, 2] + rois[:, 4]) * 0.5
        w = rois[:, 3] - rois[:, 1]
        h = rois[:, 4] - rois[:, 2]
        new_w = w * scale_factor
        new_h = h * scale_factor
        return torch.stack((rois[:, 0], cx - new_w * 0.5, cy - new_h * 0.5, cx + new_w * 0.5,
                            cy + new_h * 0.5), dim=-1)

    def _prep(self, feats):
        scales = [l.spatial_scale for l in self.roi_layers[:len(feats)]]
        if isinstance(feats, ops.Pyramid):      # converted once per step by the RoI head
            return feats, scales
        return [ops.to_channels_last(f, self.compute_dtype) for f in feats], scales


@ROI_EXTRACTORS.register_module()
class SingleRoIExtractor(BaseRoIExtractor):
    """single_level_roi_extractor.py:9-99: FPN level assignment + RoIAlign on the assigned level,
    as ONE kernel launch instead of the per-level nonzero / gather / RoIAlign / scatter loop."""

    def __init__(self, roi_layer, out_channels, featmap_strides, finest_scale=56):
        super().__init__(roi_layer, out_channels, featmap_strides)
        self.finest_scale = finest_scale

    def map_roi_levels(self, rois, num_levels):
        """single_level_roi_extractor.py:32-51; int64 like the reference."""
        return ops.level_assign(rois, num_levels, self.finest_scale).long()

    def forward(self, feats, rois, roi_scale_factor=None, bias=None):
        """``bias`` (optional [B,C,1,1]) folds ``_fuse_global`` (htd_roi_head.py:133-141)."""
        l0 = self.roi_layers[0]
        feats_cl, scales = self._prep(feats)
        if len(feats) == 1:
            return ops.roi_align_levels(feats_cl, rois, scales, l0.output_size[0],
                                        l0.sampling_ratio, bias=bias)[0]
        levels = ops.level_assign(rois, len(feats), self.finest_scale)
        if roi_scale_factor is not None:
            rois = self.roi_rescale(rois, roi_scale_factor)
        return ops.roi_align_levels(feats_cl, rois, scales, l0.output_size[0], l0.sampling_ratio,
                                    roi_level=levels, bias=bias)


@ROI_EXTRACTORS.register_module()
class AdptRoIExtractor(BaseRoIExtractor):
    """BA extractor, adaptative_roi_extractor.py:9-91.  Unlike the reference it also works for a
    single RoI (the reference's ``.squeeze()`` at :73 drops the RoI dim and crashes)."""

    def __init__(self, aggregation='sum', pre_cfg=None, post_cfg=None, edge=2, **kwargs):
        super().__init__(**kwargs)
        assert aggregation in ['sum', 'concat']
        if aggregation != 'sum' or pre_cfg is not None or post_cfg is not None:
            raise NotImplementedError('HTD uses aggregation="sum" without pre/post modules')
        self.aggregation = aggregation
        self.with_post = False
        self.with_pre = False
        self.edge = edge
        self.pool = nn.AdaptiveAvgPool2d(1)
        self.conv1 = nn.Conv2d(in_channels=256, out_channels=128, kernel_size=1, stride=1)
        self.conv2 = nn.Conv2d(in_channels=128, out_channels=1, kernel_size=1, stride=1)
        self.att = nn.Sequential(self.pool, self.conv1, nn.Tanh(), self.conv2)

    def forward(self, feats, rois, roi_scale_factor=None, add=None, bias=None):
        """``add`` ([K,C,7,7]) and ``bias`` ([B,C,1,1]) are optional residual terms fused into the
        output pass (HTDBBoxHead adds x_reg + global feature + this, htd_bbox_head.py:161-184)."""
        l0 = self.roi_layers[0]
        feats_cl, scales = self._prep(feats)
        if len(feats) == 1:
            return ops.roi_align_levels(feats_cl, rois, scales, l0.output_size[0],
                                        l0.sampling_ratio)[0]
        if rois.size(0) == 0:
            # no RoI on this rank: the result is empty but stays attached to the attention
            # parameters (zero gradients instead of none), as the reference's `0 * sum(params)` does
            # (single_level_roi_extractor.py:96-98) - every data-parallel rank then reduces the
            # same set of gradients in the same order
            out = feats[0].new_zeros(0, self.out_channels, *l0.output_size)
            return out + 0 * sum(p.sum() for p in (self.conv1.weight, self.conv1.bias,
                                                   self.conv2.weight, self.conv2.bias)).to(out.dtype)
        if roi_scale_factor is not None:
            rois = self.roi_rescale(rois, roi_scale_factor)
        return ops.ba_extract(feats_cl, rois, scales, self.conv1, self.conv2, l0.output_size[0],
                              l0.sampling_ratio, self.edge, add=add, bias=bias)
