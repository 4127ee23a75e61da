"""ORACLE (test infrastructure, not product code).

Stub harness that imports the reference's HTD hot-path modules UNMODIFIED from
``/root/reference`` (read in place, nothing copied) so that they can be executed on CPU:

  * ``mmcv`` (mmcv-full 1.2.1, pinned at /root/reference/README.md:11, not installed and not
    vendored) is replaced by a ~60-line stand-in: ``Registry``/``build_from_cfg``,
    no-op ``auto_fp16``/``force_fp32``, ``ConvModule`` (conv -> GN -> ReLU, bias='auto',
    kaiming init), ``normal_init``/``xavier_init`` and ``mmcv.ops.RoIAlign`` realised by
    ``torchvision.ops.roi_align(..., sampling_ratio, aligned=True)`` (same detectron2
    algorithm).
  * heavy ``mmdet`` package ``__init__``s are skipped by registering bare parent packages.

Only usable where ``/root/reference`` exists (the authoring container).  It is used by
``oracle/gen_golden.py`` to produce ``tests/golden/*.npz`` and by ``tests/test_oracle_cpu.py``
(skipped when the tree is absent) to prove ``oracle/restate.py`` equals the reference.
"""
import importlib
import os
import sys
import types

import torch
import torch.nn as nn

REF = os.environ.get('HTD_REFERENCE_ROOT', '/root/reference')


def available():
    return os.path.isfile(os.path.join(REF, 'mmdet/models/roi_heads/htd_roi_head.py'))


class AttrDict(dict):
    """Minimal stand-in for mmcv.Config dict nodes (attribute access + .get)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    @staticmethod
    def wrap(o):
        if isinstance(o, dict):
            return AttrDict({k: AttrDict.wrap(v) for k, v in o.items()})
        if isinstance(o, (list, tuple)):
            return type(o)(AttrDict.wrap(v) for v in o)
        return o


def _mod(name, path=None, **attrs):
    m = types.ModuleType(name)
    if path is not None:
        m.__path__ = [path]
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Registry:
    def __init__(self, name):
        self.name = name
        self.module_dict = {}

    def get(self, k):
        return self.module_dict.get(k)

    def register_module(self, name=None, force=False, module=None):
        def _reg(cls):
            self.module_dict[name or cls.__name__] = cls
            return cls
        return _reg


def _build_from_cfg(cfg, registry, default_args=None):
    args = dict(cfg)
    t = args.pop('type')
    if default_args:
        for k, v in default_args.items():
            args.setdefault(k, v)
    cls = registry.get(t) if isinstance(t, str) else t
    assert cls is not None, t
    return cls(**args)


def _noop_deco(*a, **k):
    def d(f):
        return f
    return d


class _ConvModule(nn.Module):
    def __init__(self, i, o, k, stride=1, padding=0, dilation=1, groups=1, bias='auto',
                 conv_cfg=None, norm_cfg=None, act_cfg=dict(type='ReLU'), inplace=True, **kw):
        super().__init__()
        with_norm = norm_cfg is not None
        if bias == 'auto':
            bias = not with_norm
        self.conv = nn.Conv2d(i, o, k, stride, padding, dilation, groups, bias=bias)
        nn.init.kaiming_normal_(self.conv.weight, a=0, mode='fan_out', nonlinearity='relu')
        if bias:
            nn.init.constant_(self.conv.bias, 0)
        self.gn = None
        if with_norm:
            assert norm_cfg['type'] == 'GN'
            self.gn = nn.GroupNorm(norm_cfg['num_groups'], o)
        self.activate = nn.ReLU(inplace=inplace) if act_cfg is not None else None

    def forward(self, x):
        x = self.conv(x)
        if self.gn is not None:
            x = self.gn(x)
        if self.activate is not None:
            x = self.activate(x)
        return x


def _normal_init(m, mean=0, std=1, bias=0):
    nn.init.normal_(m.weight, mean, std)
    if getattr(m, 'bias', None) is not None:
        nn.init.constant_(m.bias, bias)


def _xavier_init(m, gain=1, bias=0, distribution='normal'):
    (nn.init.xavier_uniform_ if distribution == 'uniform' else nn.init.xavier_normal_)(
        m.weight, gain=gain)
    if getattr(m, 'bias', None) is not None:
        nn.init.constant_(m.bias, bias)


class _RoIAlign(nn.Module):
    def __init__(self, output_size, spatial_scale=1.0, sampling_ratio=0, pool_mode='avg',
                 aligned=True, use_torchvision=False):
        super().__init__()
        from torch.nn.modules.utils import _pair
        self.output_size = _pair(output_size)
        self.spatial_scale = float(spatial_scale)
        self.sampling_ratio = int(sampling_ratio)
        self.aligned = aligned

    def forward(self, x, rois):
        from torchvision.ops import roi_align as tv_roi_align
        return tv_roi_align(x, rois, self.output_size, self.spatial_scale,
                            self.sampling_ratio, self.aligned)


_LOADED = None


def _batched_nms(boxes, scores, idxs, nms_cfg, class_agnostic=False):
    """mmcv.ops.nms.batched_nms of mmcv-full 1.2.1 below split_thr: coordinate shift per class,
    then the nms op named by the config - type='nms': torchvision.ops.nms (the same greedy
    algorithm); type='soft_nms': mmcv's sequential soft-NMS loop as restated in
    oracle/soft_nms_ref.c (mmcv is not installed; `iou_thr` is mmcv's deprecated alias of
    `iou_threshold`); the returned scores are the op's (`dets[:, -1]`)."""
    from torchvision.ops import nms
    cfg = dict(nms_cfg)
    kind = cfg.pop('type', 'nms')
    cfg.pop('split_thr', None)
    class_agnostic = cfg.pop('class_agnostic', class_agnostic)
    thr = cfg.pop('iou_threshold', cfg.pop('iou_thr', None))
    if class_agnostic:
        b = boxes
    else:
        b = boxes + (idxs.to(boxes) * (boxes.max() + 1))[:, None]
    if kind == 'soft_nms':
        from . import restate
        dets, keep = restate.soft_nms(b, scores, 0.3 if thr is None else thr, **cfg)
        return torch.cat([boxes[keep], dets[:, 4:5].to(boxes)], -1), keep
    assert kind == 'nms' and not cfg, (kind, cfg)
    keep = nms(b, scores, thr)
    return torch.cat([boxes[keep], scores[keep, None]], -1), keep


def load():
    """Import the reference modules; returns a namespace of the classes/functions used."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not available():
        raise RuntimeError(f'reference tree not found at {REF}')
    mmcv = _mod('mmcv', path='/nonexistent', __version__='1.2.1')
    _mod('mmcv.ops', path='/nonexistent', RoIAlign=_RoIAlign, nms=None)
    mmcv.ops = sys.modules['mmcv.ops']
    _mod('mmcv.ops.nms', batched_nms=_batched_nms)
    _mod('mmcv.cnn', path='/nonexistent', ConvModule=_ConvModule, normal_init=_normal_init,
         xavier_init=_xavier_init)
    _mod('mmcv.cnn.bricks', ConvModule=_ConvModule, build_plugin_layer=None)
    _mod('mmcv.runner', auto_fp16=_noop_deco, force_fp32=_noop_deco)
    _mod('mmcv.utils', Registry=_Registry, build_from_cfg=_build_from_cfg)
    if 'matplotlib' not in sys.modules:
        _mod('matplotlib', path='/nonexistent')
        _mod('matplotlib.pyplot')
    for pkg in ['mmdet', 'mmdet.models', 'mmdet.models.roi_heads',
                'mmdet.models.roi_heads.bbox_heads', 'mmdet.models.roi_heads.roi_extractors',
                'mmdet.models.losses', 'mmdet.models.backbones', 'mmdet.models.utils',
                'mmdet.core', 'mmdet.core.bbox', 'mmdet.core.bbox.coder',
                'mmdet.core.bbox.iou_calculators', 'mmdet.core.bbox.assigners',
                'mmdet.core.bbox.samplers', 'mmdet.core.utils', 'mmdet.utils',
                'mmdet.core.post_processing']:
        _mod(pkg, path=REF + '/' + pkg.replace('.', '/'))
    _mod('mmdet.models.backbones.resnet', Bottleneck=object)
    sys.modules['mmdet.models.utils'].ResLayer = object
    sys.modules['mmdet.models.utils'].SimplifiedBasicBlock = object
    imp = importlib.import_module
    um = imp('mmdet.utils.util_mixins')
    sys.modules['mmdet.utils'].util_mixins = um
    builder = imp('mmdet.models.builder')
    sys.modules['mmdet.models'].builder = builder
    bb = imp('mmdet.core.bbox.builder')
    iou = imp('mmdet.core.bbox.iou_calculators.iou2d_calculator')
    ic = sys.modules['mmdet.core.bbox.iou_calculators']
    ic.BboxOverlaps2D = iou.BboxOverlaps2D
    ic.bbox_overlaps = iou.bbox_overlaps
    ioub = imp('mmdet.core.bbox.iou_calculators.builder')
    ic.build_iou_calculator = ioub.build_iou_calculator
    tr = imp('mmdet.core.bbox.transforms')
    misc = imp('mmdet.core.utils.misc')
    sys.modules['mmdet.core.bbox.coder'].BaseBBoxCoder = imp(
        'mmdet.core.bbox.coder.base_bbox_coder').BaseBBoxCoder
    coder = imp('mmdet.core.bbox.coder.delta_xywh_bbox_coder')
    core = sys.modules['mmdet.core']
    for k in ['bbox2result', 'bbox2roi', 'bbox_mapping']:
        setattr(core, k, getattr(tr, k))
    core.multi_apply = misc.multi_apply
    core.build_bbox_coder = bb.build_bbox_coder
    core.build_assigner = bb.build_assigner
    core.build_sampler = bb.build_sampler
    pp = imp('mmdet.core.post_processing.bbox_nms')
    core.multiclass_nms = pp.multiclass_nms
    for k in ['bbox_mapping_back', 'bbox_flip']:
        setattr(sys.modules['mmdet.core.bbox'], k, getattr(tr, k))
    ma = imp('mmdet.core.post_processing.merge_augs')
    core.merge_aug_bboxes = ma.merge_aug_bboxes
    core.merge_aug_masks = None
    sys.modules['mmdet.core.bbox'].demodata = imp('mmdet.core.bbox.demodata')
    ar = imp('mmdet.core.bbox.assigners.assign_result')
    sys.modules['mmdet.core.bbox.assigners'].AssignResult = ar.AssignResult
    imp('mmdet.core.bbox.assigners.base_assigner')
    mia = imp('mmdet.core.bbox.assigners.max_iou_assigner')
    sr = imp('mmdet.core.bbox.samplers.sampling_result')
    imp('mmdet.core.bbox.samplers.base_sampler')
    rs = imp('mmdet.core.bbox.samplers.random_sampler')
    losses = sys.modules['mmdet.models.losses']
    acc = imp('mmdet.models.losses.accuracy')
    losses.accuracy = acc.accuracy
    losses.Accuracy = acc.Accuracy
    imp('mmdet.models.losses.utils')
    ce = imp('mmdet.models.losses.cross_entropy_loss')
    sl1 = imp('mmdet.models.losses.smooth_l1_loss')
    bh = imp('mmdet.models.roi_heads.bbox_heads.bbox_head')
    cf = imp('mmdet.models.roi_heads.bbox_heads.convfc_bbox_head')
    htdh = imp('mmdet.models.roi_heads.bbox_heads.htd_bbox_head')
    gch = imp('mmdet.models.roi_heads.bbox_heads.global_context_head')
    imp('mmdet.models.roi_heads.roi_extractors.base_roi_extractor')
    sle = imp('mmdet.models.roi_heads.roi_extractors.single_level_roi_extractor')
    ada = imp('mmdet.models.roi_heads.roi_extractors.adaptative_roi_extractor')
    imp('mmdet.models.roi_heads.base_roi_head')
    _mod('mmdet.models.roi_heads.test_mixins', BBoxTestMixin=type('BBoxTestMixin', (), {}),
         MaskTestMixin=type('MaskTestMixin', (), {}))
    rh = imp('mmdet.models.roi_heads.htd_roi_head')
    _mod('mmdet.models.necks', path=REF + '/mmdet/models/necks')
    fpn = imp('mmdet.models.necks.fpn')
    ns = types.SimpleNamespace(
        FPN=fpn.FPN, merge_aug_bboxes=ma.merge_aug_bboxes,
        HTDRoIHead=rh.HTDRoIHead, HTDBBoxHead=htdh.HTDBBoxHead,
        AdptRoIExtractor=ada.AdptRoIExtractor, SingleRoIExtractor=sle.SingleRoIExtractor,
        GlobalContextHead=gch.GlobalContextHead, Shared2FCBBoxHead=cf.Shared2FCBBoxHead,
        BBoxHead=bh.BBoxHead, bbox_overlaps=iou.bbox_overlaps, bbox2roi=tr.bbox2roi,
        delta2bbox=coder.delta2bbox, bbox2delta=coder.bbox2delta,
        DeltaXYWHBBoxCoder=coder.DeltaXYWHBBoxCoder, MaxIoUAssigner=mia.MaxIoUAssigner,
        RandomSampler=rs.RandomSampler, SamplingResult=sr.SamplingResult,
        AssignResult=ar.AssignResult, multiclass_nms=pp.multiclass_nms,
        CrossEntropyLoss=ce.CrossEntropyLoss,
        SmoothL1Loss=sl1.SmoothL1Loss, accuracy=acc.accuracy, RoIAlign=_RoIAlign)
    _LOADED = ns
    return ns


def load_config(name='htd_resnet50_1x.py'):
    """exec a configs/htd/*.py file; returns (roi_head_cfg, train_cfg.rcnn, test_cfg.rcnn)."""
    path = os.path.join(REF, 'configs/htd', name)
    g = {}
    src = open(path).read()
    # _base_ inheritance is irrelevant for the roi_head block; exec the file as Python
    exec(compile(src, path, 'exec'), g)
    model = g['model']
    return (AttrDict.wrap(model['roi_head']), AttrDict.wrap(g['train_cfg']['rcnn']),
            AttrDict.wrap(g['test_cfg']['rcnn']))


def build_head(double=False, config='htd_resnet50_1x.py'):
    """Build the reference HTDRoIHead exactly as configs/htd/htd_resnet50_1x.py:38-95 says."""
    ns = load()
    roi_head, train_rcnn, test_rcnn = load_config(config)
    cfg = dict(roi_head)
    cfg.pop('type')
    head = ns.HTDRoIHead(**cfg, train_cfg=train_rcnn, test_cfg=test_rcnn)
    head.init_weights(None)
    if double:
        head = head.double()
    return head


if __name__ == '__main__':
    h = build_head()
    print('reference HTDRoIHead built; params =', sum(p.numel() for p in set(h.parameters())))
