"""htd_b200 - B200-native (sm_100a) implementation of the HTD RoI-head hot path behind the
reference's mmdet-2.7 plugin surface.  Importing the package registers the plugin classes
(``registry.HEADS`` / ``ROI_EXTRACTORS`` / ``ROI_LAYERS`` ...); kernels live in
``htd_b200/_lib/libhtd_b200.so`` (built by ``python -m htd_b200.build``) and are reached through
the C ABI of ``include/htd_b200.h``.  There is no CPU or PyTorch fallback for the hot ops."""
from . import registry  # noqa: F401
from .roi_extractors import AdptRoIExtractor, RoIAlign, SingleRoIExtractor  # noqa: F401
from .core import (CrossEntropyLoss, DeltaXYWHBBoxCoder, MaxIoUAssigner,  # noqa: F401
                   RandomSampler, SmoothL1Loss, bbox2result, bbox2roi, multiclass_nms)
from .bbox_heads import (BBoxHead, ConvFCBBoxHead, GlobalContextHead,  # noqa: F401
                         HTDBBoxHead, Shared2FCBBoxHead)
from .roi_head import HTDRoIHead  # noqa: F401
from .necks import FPN  # noqa: F401
from .dense_heads import AnchorGenerator, RPNHead  # noqa: F401
from .config import htd_roi_head_cfg, build_htd_roi_head  # noqa: F401

__version__ = '0.1.0'
