"""Deterministic synthetic inputs for the HTD RoI-head path.

Everything is generated on the CPU with explicitly seeded ``torch.Generator`` objects so the
authoring container (which produces ``tests/golden``) and the GPU box (same image, same torch
build) see bit-identical tensors.  Used by ``bench.py``, the tests and ``oracle/gen_golden.py``.

Pyramid shapes for an 800x1333 image padded to 800x1344 are the reference's own test shapes
(``tests/test_models/test_roi_extractor.py:31-36``) plus P6 = stride-2 max-pool of P5
(``mmdet/models/necks/fpn.py:199``).
"""
import math
import zlib

import torch

STRIDES = (4, 8, 16, 32, 64)


def pyramid_shapes(img_h=800, img_w=1333, pad=32):
    """Feature-map (H, W) of P2..P6 for an image padded to a multiple of ``pad``."""
    ph = int(math.ceil(img_h / pad) * pad)
    pw = int(math.ceil(img_w / pad) * pad)
    shapes = [(ph // s, pw // s) for s in STRIDES[:4]]
    h5, w5 = shapes[-1]
    shapes.append(((h5 + 1) // 2, (w5 + 1) // 2))  # max_pool2d(1, stride=2)
    return shapes


def make_pyramid(num_imgs, img_h=800, img_w=1333, channels=256, seed=1000, dtype=torch.float32):
    """List of 5 NCHW tensors; image ``i`` is drawn from seed ``seed + i`` (N(0,1))."""
    shapes = pyramid_shapes(img_h, img_w)
    levels = []
    for li, (h, w) in enumerate(shapes):
        per_img = []
        for i in range(num_imgs):
            g = torch.Generator().manual_seed(seed + i + 7919 * li)
            per_img.append(torch.randn(1, channels, h, w, generator=g, dtype=torch.float32))
        levels.append(torch.cat(per_img, 0).to(dtype))
    return levels


def make_proposals(num_imgs, num_rois=512, img_h=800, img_w=1333, seed=1234,
                   min_scale=16.0, max_scale=800.0):
    """Per-image [num_rois, 4] boxes: scale ~ logU(min,max), aspect ~ logU(.5, 2), centre
    uniform over the image, corners clipped to the image."""
    out = []
    for i in range(num_imgs):
        g = torch.Generator().manual_seed(seed + i)
        u = torch.rand(num_rois, 4, generator=g, dtype=torch.float64)
        s = torch.exp(math.log(min_scale) + u[:, 0] * (math.log(max_scale) - math.log(min_scale)))
        r = torch.exp(math.log(0.5) + u[:, 1] * (math.log(2.0) - math.log(0.5)))
        w = s * torch.sqrt(r)
        h = s / torch.sqrt(r)
        cx = u[:, 2] * img_w
        cy = u[:, 3] * img_h
        x1 = (cx - w / 2).clamp(0, img_w)
        x2 = (cx + w / 2).clamp(0, img_w)
        y1 = (cy - h / 2).clamp(0, img_h)
        y2 = (cy + h / 2).clamp(0, img_h)
        out.append(torch.stack([x1, y1, x2, y2], 1).float())
    return out


def make_gt(num_imgs, proposals, num_pos=128, num_classes=80, seed=4321):
    """Synthetic sampling outcome: the first ``num_pos`` proposals of every image are the
    positives (``SamplingResult.bboxes`` puts positives first,
    ``mmdet/core/bbox/samplers/sampling_result.py:52-54``).  Returns per-image dicts with
    ``pos_gt_labels`` (U{0..79}), ``pos_gt_bboxes`` (the positive box jittered so regression
    targets are O(1)) and ``gt_labels_unique`` (5 classes per image for the SFA loss)."""
    out = []
    for i in range(num_imgs):
        g = torch.Generator().manual_seed(seed + i)
        p = proposals[i][:num_pos]
        labels = torch.randint(0, num_classes, (p.shape[0],), generator=g)
        jit = 1.0 + 0.1 * torch.randn(p.shape[0], 4, generator=g)
        w = (p[:, 2] - p[:, 0]).clamp(min=1.0)
        h = (p[:, 3] - p[:, 1]).clamp(min=1.0)
        cx = (p[:, 0] + p[:, 2]) * 0.5 + 0.05 * w * torch.randn(p.shape[0], generator=g)
        cy = (p[:, 1] + p[:, 3]) * 0.5 + 0.05 * h * torch.randn(p.shape[0], generator=g)
        gw = w * jit[:, 0].abs()
        gh = h * jit[:, 1].abs()
        gtb = torch.stack([cx - gw / 2, cy - gh / 2, cx + gw / 2, cy + gh / 2], 1)
        uniq = torch.randperm(num_classes, generator=g)[:5].sort().values
        out.append(dict(pos_gt_labels=labels, pos_gt_bboxes=gtb, gt_labels_unique=uniq))
    return out


def make_detection_batch(num_imgs, num_props=2000, num_gt=(7, 12), gt_capacity=32, img_h=800,
                         img_w=1333, near=0.2, jitter=0.15, num_classes=80, seed=77):
    """Inputs of the complete training step (``HTDRoIHead.forward_train_static``): RPN-like
    proposals [B,N,4] of which a fraction ``near`` are jittered copies of the ground-truth boxes
    (so that assignment finds positives), gt boxes [B,G,4] / labels [B,G] padded to
    ``gt_capacity`` slots, and the per-image gt counts [B] (int32)."""
    g = torch.Generator().manual_seed(seed)
    B, N, G = num_imgs, num_props, gt_capacity
    props = torch.stack(make_proposals(B, N, img_h, img_w, seed=seed * 31))
    gt = torch.zeros(B, G, 4)
    labels = torch.zeros(B, G, dtype=torch.long)
    counts = torch.tensor([num_gt[b % len(num_gt)] for b in range(B)], dtype=torch.int32)
    lim = torch.tensor([img_w, img_h, img_w, img_h], dtype=torch.float32)
    for b in range(B):
        ng = int(counts[b])
        wh = torch.exp(torch.rand(ng, 2, generator=g) * 2.5 + 3.0)
        ctr = torch.rand(ng, 2, generator=g) * lim[:2]
        box = torch.min(torch.cat([ctr - wh / 2, ctr + wh / 2], 1).clamp(min=0), lim)
        gt[b, :ng] = box
        labels[b, :ng] = torch.randint(0, num_classes, (ng,), generator=g)
        k = int(N * near)
        if ng and k:
            src = box[torch.randint(0, ng, (k,), generator=g)]
            swh = (src[:, 2:] - src[:, :2]).repeat(1, 2)
            q = src + torch.randn(k, 4, generator=g) * jitter * swh
            q = torch.cat([torch.min(q[:, :2], q[:, 2:]), torch.max(q[:, :2], q[:, 2:])], 1)
            props[b, torch.randperm(N, generator=g)[:k]] = torch.min(q.clamp(min=0), lim)
    return props, gt, labels, counts


def _gen_for(name, seed):
    return torch.Generator().manual_seed((zlib.crc32(name.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)


def _alternating(n, dtype=torch.float32):
    return torch.where(torch.arange(n) % 2 == 0, 1.0, -1.0).to(dtype)


def _stable_value(name, p, z):
    """Parameter value of the 'stable' scheme (see ``fill_params_``); z ~ N(0,1) of p's shape."""
    relu_fc = ('shared_fcs.' in name or '.fcs.' in name or 'graph_lvl' in name)
    if name.endswith('gn.weight'):
        return 1.0 + 0.05 * z
    if name.endswith('gn.bias'):                      # GN output has unit variance: +-6 sigma
        return 6.0 * _alternating(p.shape[0]) + 0.1 * z
    if 'glbctx_head.convs' in name:                   # SFA: conv + bias + ReLU
        return 3.0 * _alternating(p.shape[0]) + 0.1 * z if name.endswith('bias') else 0.005 * z
    if relu_fc and name.endswith('bias'):
        return 1.5 * _alternating(p.shape[0]) + 0.05 * z
    if relu_fc:
        return (0.001 if p.shape[1] > 2048 else 0.006) * z
    if 'convs.3.conv.weight' in name:                 # last tower conv: no bias, no norm -> ReLU
        return 0.005 * z + 0.001 * _alternating(p.shape[0]).view(-1, 1, 1, 1)
    if name.endswith('fc_reg.weight'):
        return 0.001 * z
    if name.endswith('bbox_head.0.fc_cls.weight'):
        # stage-0 classifier = the PGraph prototype (htd_bbox_head.py:158,194): logits of O(3) make
        # the semantic vectors `sam` differ between RoIs by far more than a bf16 ulp, so the global
        # graph softmax - and the gradient that reaches this layer through it - is well resolved
        return 0.1 * z
    if name.endswith('fc_cls.weight') or name.endswith('glbctx_head.fc.weight'):
        return 0.01 * z
    if name.endswith('bias'):
        return 0.01 * z
    if p.dim() == 4 and p.shape[-1] == 3:             # tower convs followed by GroupNorm
        return math.sqrt(2.0 / (p.shape[0] * 9)) * z
    return 0.05 * z                                   # BA attention convs


def fill_params_(module, scheme='n005', seed=0):
    """Deterministically (re)initialise every parameter of ``module`` from its state-dict
    name, so the reference head (oracle side) and this package's head (product side) get
    identical weights without shipping a 189 MB checkpoint.

    scheme 'n005': every tensor ~ N(0, 0.05) (GroupNorm weight 1 + N(0, 0.05)); makes the
    PGraph softmax non-trivial.  scheme 'init': magnitudes of the reference's
    own initialisers (normal 0.01 / 0.001 for fc_cls / fc_reg, xavier-uniform Linear,
    kaiming-normal conv, zero biases) - ``htd_bbox_head.py:136-145``.  scheme 'stable'
    ("gate-stable"): every ReLU on the path gets a pre-activation whose sign is fixed by a large
    alternating bias (FC / SFA conv bias, GroupNorm beta; the bias-free last tower conv gets a
    per-output-channel mean shift of its weights) with small weights around it, so that bf16 / fp32
    rounding cannot flip a gate - end-to-end gradients can then be gated in the max-norm like the
    forward values (a flipped gate changes a gradient element by 100 %, whatever the arithmetic).
    """
    seen = set()
    with torch.no_grad():
        for name, p in module.named_parameters(remove_duplicate=False):
            if id(p) in seen:      # AdptRoIExtractor registers conv1/conv2 twice (att.1/att.3)
                continue
            seen.add(id(p))
            g = _gen_for(name, seed)
            z = torch.randn(p.shape, generator=g, dtype=torch.float32)
            if scheme == 'n005':
                v = 0.05 * z
                if name.endswith('gn.weight'):
                    v = 1.0 + v
                if '.fc_reg.' in name:      # keep refined boxes non-degenerate
                    v = 0.001 * z
            elif scheme == 'init':
                if name.endswith('bias'):
                    v = torch.zeros_like(z)
                elif name.endswith('gn.weight'):
                    v = torch.ones_like(z)
                elif name.endswith('fc_cls.weight') or name.endswith('glbctx_head.fc.weight'):
                    v = 0.01 * z
                elif name.endswith('fc_reg.weight'):
                    v = 0.001 * z
                elif p.dim() == 2:
                    bound = math.sqrt(6.0 / (p.shape[0] + p.shape[1]))
                    u = torch.rand(p.shape, generator=g, dtype=torch.float32)
                    v = (2 * u - 1) * bound
                else:
                    fan_out = p.shape[0] * p[0][0].numel()
                    v = math.sqrt(2.0 / fan_out) * z
            elif scheme == 'stable':
                v = _stable_value(name, p, z)
            else:
                raise ValueError(scheme)
            p.copy_(v.to(p.dtype))
    return module


def level_histogram(rois, finest_scale=56.0, num_levels=4):
    scale = torch.sqrt((rois[:, 2] - rois[:, 0]) * (rois[:, 3] - rois[:, 1]))
    lv = torch.floor(torch.log2(scale / finest_scale + 1e-6)).clamp(0, num_levels - 1).long()
    return torch.bincount(lv, minlength=num_levels).tolist()


def make_sampling(bboxes, num_pos, gt):
    """Synthetic sampling result in the reference's layout (positives first,
    ``sampling_result.py:52-54``): the first ``num_pos`` rows of ``bboxes`` are the positives,
    their targets come from ``make_gt``.  Duck-types ``core.SamplingResult``."""
    from types import SimpleNamespace
    n = min(num_pos, bboxes.size(0))
    dev = bboxes.device
    return SimpleNamespace(
        pos_bboxes=bboxes[:n], neg_bboxes=bboxes[n:],
        pos_gt_bboxes=gt['pos_gt_bboxes'][:n].to(device=dev, dtype=bboxes.dtype),
        pos_gt_labels=gt['pos_gt_labels'][:n].to(dev),
        pos_is_gt=None, bboxes=bboxes)      # None: no sampled box is a gt box (static shapes)


def sampled_forward_train(head, x, proposals, gts, img_shapes, num_pos=128):
    """``HTDRoIHead.forward_train`` with the random assign+sample steps replaced by the
    positives-first synthetic sampling above - the protocol of bench.py and of
    the parity tests (oracle side: ``restate.HTDRoIHead.forward_train_sampled``)."""
    metas = [dict(img_shape=s) for s in img_shapes]
    dev = proposals[0].device
    labels = [g['gt_labels_unique'].to(dev) for g in gts]

    def sampling_fn(stage, plist):
        return [make_sampling(p, num_pos, g) for p, g in zip(plist, gts)]

    return head.forward_train(x, metas, proposals, None, labels, sampling_fn=sampling_fn)
