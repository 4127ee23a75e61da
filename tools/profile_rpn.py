"""Where the time of RPNHead.get_bboxes goes (torch profiler, CUDA + CPU totals)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
from htd_b200.dense_heads import RPNHead

rpn = RPNHead(256, 256).cuda().bfloat16().to(memory_format=torch.channels_last)
g = torch.Generator().manual_seed(0)
feats = [torch.randn(2, 256, h, w, generator=g).cuda().bfloat16().contiguous(memory_format=torch.channels_last)
         for h, w in ((200, 336), (100, 168), (50, 84), (25, 42), (13, 21))]
metas = [dict(img_shape=(800, 1333, 3), scale_factor=1.0)] * 2
cfg = dict(nms_across_levels=False, nms_pre=2000, nms_post=2000, max_num=2000, nms_thr=0.7, min_bbox_size=0)
with torch.no_grad():
    cls, reg = rpn(feats)
    for _ in range(3):
        rpn.get_bboxes(cls, reg, metas, cfg)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(5):
            rpn.get_bboxes(cls, reg, metas, cfg)
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=14, max_name_column_width=60))
