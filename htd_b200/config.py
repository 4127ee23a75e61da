"""The RoI-head block of the reference configs (identical in all five ``configs/htd/*.py``;
values from ``configs/htd/htd_resnet50_1x.py:38-95`` and the rcnn train/test settings at
``:122-168``), as plain dicts, plus a builder.  With a real mmdet the same dicts are what
``Config.fromfile`` yields."""
import copy

from .registry import build_head

_ROI_LAYER = dict(type='RoIAlign', output_size=7, sampling_ratio=0)


def _rcnn_stage(iou):
    return dict(
        assigner=dict(type='MaxIoUAssigner', pos_iou_thr=iou, neg_iou_thr=iou, min_pos_iou=iou,
                      match_low_quality=False, ignore_iof_thr=-1),
        sampler=dict(type='RandomSampler', num=512, pos_fraction=0.25, neg_pos_ub=-1,
                     add_gt_as_proposals=True),
        pos_weight=-1, debug=False)


def _bbox_head(type_, stds, **extra):
    return dict(type=type_, in_channels=256, fc_out_channels=1024, roi_feat_size=7, num_classes=80,
                bbox_coder=dict(type='DeltaXYWHBBoxCoder', target_means=[0., 0., 0., 0.],
                                target_stds=stds),
                reg_class_agnostic=True,
                loss_cls=dict(type='CrossEntropyLoss', use_sigmoid=False, loss_weight=1.0),
                loss_bbox=dict(type='SmoothL1Loss', beta=1.0, loss_weight=1.0), **extra)


def htd_roi_head_cfg():
    return dict(
        type='HTDRoIHead', num_stages=2, with_global=True, stage_loss_weights=[1, 0.5],
        bbox_roi_extractor=[
            dict(type='SingleRoIExtractor', roi_layer=dict(_ROI_LAYER), out_channels=256,
                 featmap_strides=[4, 8, 16, 32]),
            dict(type='AdptRoIExtractor', edge=1, roi_layer=dict(_ROI_LAYER), out_channels=256,
                 featmap_strides=[4, 8, 16, 32])],
        bbox_head=[_bbox_head('Shared2FCBBoxHead', [0.1, 0.1, 0.2, 0.2]),
                   _bbox_head('HTDBBoxHead', [0.05, 0.05, 0.1, 0.1], relpace=False, edge=1)],
        train_cfg=[_rcnn_stage(0.5), _rcnn_stage(0.6)],
        test_cfg=dict(score_thr=0.05, nms=dict(type='nms', iou_threshold=0.5), max_per_img=100))


def build_htd_roi_head(**overrides):
    cfg = copy.deepcopy(htd_roi_head_cfg())
    cfg.update(overrides)
    return build_head(cfg)
