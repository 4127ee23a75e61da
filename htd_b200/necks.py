"""FPN neck emitting the pyramid in the layout the RoI head reads (the
producer side; reference: ``mmdet/models/necks/fpn.py:9-216`` as built by
``configs/htd/htd_resnet50_1x.py:17-21`` - in_channels [256, 512, 1024, 2048], out 256, five
outputs, no extra convs, nearest top-down).

The reference hands the head five NCHW fp32 maps; this package's extractors read channels-last
maps (bf16 in the benchmarked configuration), so with the reference's FPN in front every step
pays a layout / cast pass over the whole pyramid in both directions.  This FPN produces its
outputs channels-last in the compute dtype - ``ops.to_channels_last`` / ``HTDRoIHead._pyramid``
then take them as they are (no copy), and the gradient of the pyramid comes back the same way.

  * lateral 1x1 convs: rows-by-channels products on the package's dense tcgen05 kernel
    (``dense.linear``: forward, dgrad, wgrad, bias gradient) when the maps are bf16;
  * top-down merge ``lat[i-1] += interpolate(lat[i], size=..., mode='nearest')`` and the extra
    level ``max_pool2d(out, 1, stride=2)``: ``csrc/fpn.cu`` (forward + backward);
  * the 3x3 output convs on the full-size maps: cuDNN channels-last (library).

Parameter names follow the reference (``lateral_convs.N.conv.weight`` ...), so its checkpoints
load unchanged.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, dense, ops
from ._lib import check, lib, ptr, stream
from .registry import Registry

NECKS = Registry('neck')


def _cl_view(x):
    """[B,C,H,W] channels-last tensor -> its [B,H,W,C] memory as a contiguous tensor."""
    return x.permute(0, 2, 3, 1)


class _TopDown(torch.autograd.Function):
    """fine + nearest-upsample(coarse) on channels-last maps (fpn.py:187-190)."""

    @staticmethod
    def forward(ctx, fine, coarse):
        _lib.require_cuda(fine, coarse)
        B, C, Hf, Wf = fine.shape
        Hc, Wc = coarse.shape[2:]
        out = torch.empty_like(fine, memory_format=torch.channels_last)
        check(lib().htd_fpn_topdown_fwd(ptr(fine), ptr(coarse), ptr(out), _lib.dt(fine), B, Hf, Wf,
                                        Hc, Wc, C, stream()), 'htd_fpn_topdown_fwd')
        ctx.shape_c = coarse.shape
        return out

    @staticmethod
    def backward(ctx, dout):
        dout = ops.to_channels_last(dout)
        B, C, Hf, Wf = dout.shape
        _, _, Hc, Wc = ctx.shape_c
        dcoarse = torch.empty(ctx.shape_c, dtype=dout.dtype, device=dout.device).contiguous(
            memory_format=torch.channels_last)
        check(lib().htd_fpn_topdown_bwd(ptr(dout), ptr(dcoarse), _lib.dt(dout), B, Hf, Wf, Hc, Wc, C,
                                        stream()), 'htd_fpn_topdown_bwd')
        return dout, dcoarse


class _Subsample(torch.autograd.Function):
    """x[:, :, ::2, ::2] = max_pool2d(x, 1, stride=2) on a channels-last map (fpn.py:201)."""

    @staticmethod
    def forward(ctx, x):
        _lib.require_cuda(x)
        B, C, H, W = x.shape
        out = torch.empty((B, C, (H - 1) // 2 + 1, (W - 1) // 2 + 1), dtype=x.dtype,
                          device=x.device).contiguous(memory_format=torch.channels_last)
        check(lib().htd_fpn_subsample(ptr(x), ptr(out), _lib.dt(x), B, H, W, C, 0, stream()),
              'htd_fpn_subsample')
        ctx.shape = x.shape
        return out

    @staticmethod
    def backward(ctx, dout):
        dout = ops.to_channels_last(dout)
        B, C, H, W = ctx.shape
        dx = torch.empty(ctx.shape, dtype=dout.dtype, device=dout.device).contiguous(
            memory_format=torch.channels_last)
        check(lib().htd_fpn_subsample(ptr(dout), ptr(dx), _lib.dt(dout), B, H, W, C, 1, stream()),
              'htd_fpn_subsample')
        return dx


class _Conv(nn.Module):
    """mmcv ConvModule without norm / activation: the one sub-module ``conv`` (bias=True)."""

    def __init__(self, cin, cout, k, padding=0):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, padding=padding)
        self.conv.to(memory_format=torch.channels_last)

    def forward(self, x):
        c = self.conv
        if c.kernel_size == (1, 1) and dense.usable(x, c.weight) and c.in_channels % 64 == 0 and \
                c.out_channels % 64 == 0:
            B, _, H, W = x.shape
            rows = _cl_view(x).reshape(B * H * W, c.in_channels)
            y = dense.linear(rows, c.weight.view(c.out_channels, c.in_channels), c.bias)
            return y.view(B, H, W, c.out_channels).permute(0, 3, 1, 2)
        return F.conv2d(x, c.weight, c.bias, c.stride, c.padding)


@NECKS.register_module()
class FPN(nn.Module):
    """necks/fpn.py:9-216 for the configurations of configs/htd (no extra convs, no norm, nearest
    top-down, extra levels by stride-2 subsampling)."""

    def __init__(self, in_channels, out_channels, num_outs, start_level=0, end_level=-1,
                 add_extra_convs=False, extra_convs_on_inputs=True, relu_before_extra_convs=False,
                 no_norm_on_lateral=False, conv_cfg=None, norm_cfg=None, act_cfg=None,
                 upsample_cfg=dict(mode='nearest'), compute_dtype=None):
        super().__init__()
        if add_extra_convs or conv_cfg is not None or norm_cfg is not None or act_cfg is not None:
            raise NotImplementedError('configs/htd build the plain FPN (fpn.py defaults)')
        if dict(upsample_cfg) != dict(mode='nearest'):
            raise NotImplementedError("upsample_cfg: configs/htd use dict(mode='nearest')")
        assert isinstance(in_channels, list)
        self.in_channels, self.out_channels = in_channels, out_channels
        self.num_ins, self.num_outs = len(in_channels), num_outs
        if end_level == -1:
            self.backbone_end_level = self.num_ins
            assert num_outs >= self.num_ins - start_level
        else:
            self.backbone_end_level = end_level
            assert end_level <= len(in_channels) and num_outs == end_level - start_level
        self.start_level, self.end_level = start_level, end_level
        self.compute_dtype = compute_dtype          # None: the dtype of the weights
        self.lateral_convs = nn.ModuleList()
        self.fpn_convs = nn.ModuleList()
        for i in range(self.start_level, self.backbone_end_level):
            self.lateral_convs.append(_Conv(in_channels[i], out_channels, 1))
            self.fpn_convs.append(_Conv(out_channels, out_channels, 3, padding=1))
        self.init_weights()

    def init_weights(self):
        for m in self.modules():                    # fpn.py:159-163: xavier_init(uniform), bias 0
            if isinstance(m, nn.Conv2d):
                nn.init.xavier_uniform_(m.weight, gain=1)
                nn.init.constant_(m.bias, 0)

    def forward(self, inputs):
        assert len(inputs) == len(self.in_channels)
        dt = self.compute_dtype or self.lateral_convs[0].conv.weight.dtype    # the module's dtype
        xs = [ops.to_channels_last(x, dt) for x in inputs]
        laterals = [conv(xs[i + self.start_level]) for i, conv in enumerate(self.lateral_convs)]
        n = len(laterals)
        for i in range(n - 1, 0, -1):               # fpn.py:179-190
            laterals[i - 1] = _TopDown.apply(ops.to_channels_last(laterals[i - 1]),
                                             ops.to_channels_last(laterals[i]))
        outs = [self.fpn_convs[i](laterals[i]) for i in range(n)]
        for _ in range(self.num_outs - n):          # fpn.py:196-201
            outs.append(_Subsample.apply(ops.to_channels_last(outs[-1])))
        return tuple(outs)
