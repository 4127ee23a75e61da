"""Plugin registries.  The reference instantiates the hot-path classes by type name through
mmcv registries (``mmdet/models/builder.py:4-57``: ``HEADS``, ``ROI_EXTRACTORS``, ``LOSSES``) and
looks the RoI op up with ``getattr(mmcv.ops, layer_type)``
(``roi_extractors/base_roi_extractor.py:49-55``).  The same names are registered here; when a
real mmdet/mmcv is importable ``register_into_mmdet()`` overrides its entries so
``configs/htd/*.py`` build this package's classes unchanged.
"""


class Registry:
    def __init__(self, name):
        self.name = name
        self.module_dict = {}

    def get(self, key):
        return self.module_dict.get(key)

    def register_module(self, name=None, force=False, module=None):
        def _reg(cls):
            key = name or cls.__name__
            if key in self.module_dict and not force and self.module_dict[key] is not cls:
                raise KeyError(f'{key} is already registered in {self.name}')
            self.module_dict[key] = cls
            return cls
        if module is not None:
            return _reg(module)
        return _reg


HEADS = Registry('head')
ROI_EXTRACTORS = Registry('roi_extractor')
ROI_LAYERS = Registry('roi_layer')          # stands in for the ``mmcv.ops`` namespace lookup
BBOX_ASSIGNERS = Registry('bbox_assigner')
BBOX_SAMPLERS = Registry('bbox_sampler')
BBOX_CODERS = Registry('bbox_coder')
LOSSES = Registry('loss')


def build_from_cfg(cfg, registry, default_args=None):
    if not isinstance(cfg, dict) or 'type' not in cfg:
        raise TypeError(f'cfg must be a dict with a "type" key, got {cfg!r}')
    args = dict(cfg)
    t = args.pop('type')
    if default_args:
        for k, v in default_args.items():
            args.setdefault(k, v)
    cls = registry.get(t) if isinstance(t, str) else t
    if cls is None:
        raise KeyError(f'{t} is not in the {registry.name} registry')
    return cls(**args)


def build_head(cfg):
    return build_from_cfg(cfg, HEADS)


def build_roi_extractor(cfg):
    return build_from_cfg(cfg, ROI_EXTRACTORS)


def build_loss(cfg):
    return build_from_cfg(cfg, LOSSES)


def build_assigner(cfg, **kw):
    return build_from_cfg(cfg, BBOX_ASSIGNERS, kw)


def build_sampler(cfg, **kw):
    return build_from_cfg(cfg, BBOX_SAMPLERS, kw)


def build_bbox_coder(cfg, **kw):
    return build_from_cfg(cfg, BBOX_CODERS, kw)


def register_into_mmdet():
    """Override mmdet's registry entries (and mmcv.ops.RoIAlign) with this package's classes.
    Returns the list of names overridden; raises ImportError when mmdet/mmcv are absent."""
    import mmcv.ops
    from mmdet.models import builder as mb
    done = []
    for src, dst in ((HEADS, mb.HEADS), (ROI_EXTRACTORS, mb.ROI_EXTRACTORS)):
        for name, cls in src.module_dict.items():
            dst.register_module(name=name, force=True, module=cls)
            done.append(name)
    for name, cls in ROI_LAYERS.module_dict.items():
        setattr(mmcv.ops, name, cls)
        done.append('mmcv.ops.' + name)
    return done
