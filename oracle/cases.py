"""ORACLE (test infrastructure).  Named parity cases and ONE driver that runs any
implementation of the HTD RoI-head plugin surface through them.

The reference's modules (via refshim), the CPU restatement (oracle/restate.py) and the CUDA
product (htd_b200) expose the same structure -
``head.bbox_roi_extractor[0|1](feats, rois)``, ``head.bbox_head[1](x_cls, x_reg, feat, rois,
fc_cls_0, enhanced, pos_rois, global_feat)``, ``head.glbctx_head(x)`` - so the driver is
implementation-agnostic; only the training/test entry point is passed in as a callable.
Outputs are a flat ``{name: tensor}`` dict that ``summarize``/``compare`` turn into small
golden fixtures (strided samples + checksums) under tests/golden/.
"""
import numpy as np
import torch

from htd_b200 import synth

CASES = {
    # name: images, (H, W), rois/img, positives/img, (min,max) scale, weight scheme, seed
    'small': dict(B=2, hw=(320, 448), K=48, P=12, scales=(8.0, 600.0), scheme='n005', seed=0),
    'mid': dict(B=1, hw=(512, 640), K=96, P=24, scales=(8.0, 900.0), scheme='init', seed=1),
    # gate-stable weights (synth.fill_params_ 'stable'): no ReLU gate can flip under bf16 / fp32
    # rounding, so END-TO-END gradients are gated in the max-norm (2e-2 bf16) like forward values
    'small_s': dict(B=2, hw=(320, 448), K=64, P=24, scales=(8.0, 600.0), scheme='stable', seed=0),
    # BASELINE.json configs[1] = the benchmarked size: 2 images x 512 RoIs (128 positives),
    # 800x1333 pyramid (P2 rows of 336 px: the wide-footprint paths only occur at this size)
    'c2': dict(B=2, hw=(800, 1333), K=512, P=128, scales=(16.0, 800.0), scheme='stable', seed=0),
}


def case_inputs(name, dtype=torch.float32, device='cpu'):
    c = CASES[name]
    H, W = c['hw']
    x = [t.to(dtype).to(device) for t in synth.make_pyramid(c['B'], H, W, seed=1000 + c['seed'])]
    pdt = dtype if dtype in (torch.float32, torch.float64) else torch.float32   # boxes stay fp32
    props = [p.to(pdt).to(device) for p in synth.make_proposals(
        c['B'], c['K'], H, W, seed=1234 + c['seed'], min_scale=c['scales'][0],
        max_scale=c['scales'][1])]
    gts = synth.make_gt(c['B'], [p.float().cpu() for p in props], num_pos=c['P'],
                        seed=4321 + c['seed'])
    gts = [{k: v.to(device) for k, v in g.items()} for g in gts]
    shapes = [(H, W, 3)] * c['B']
    return c, x, props, gts, shapes


def seeded_like(t, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(t.shape, generator=g, dtype=torch.float32).to(t.dtype).to(t.device)


def _rois(props):
    return torch.cat([torch.cat([p.new_full((p.size(0), 1), i), p], 1)
                      for i, p in enumerate(props)], 0)


def graph_masks(bbox_overlaps_fn, levels, rois):
    """h_local_mask and degree of every (image, level) group, htd_bbox_head.py:198-209; row
    order = ascending original RoI index (boolean-mask semantics)."""
    out = {}
    bs = int(rois[:, 0].max().item()) + 1
    for b in range(bs):
        for i in range(4):
            sel = (rois[:, 0] == b) & (levels == i)
            if sel.any():
                r = rois[sel, 1:5]
                M = bbox_overlaps_fn(r, r).fill_diagonal_(1.)
                M[M > 0] = 1.
                out[(b, i)] = (sel.nonzero(as_tuple=True)[0], M, M.sum(-1))
    return out


def run_extractors(head, name, dtype, device='cpu'):
    c, x, props, gts, shapes = case_inputs(name, dtype, device)
    out = {}
    rois = _rois(props)
    pos_rois = _rois([p[:c['P']] for p in props])
    ext0, ext1 = head.bbox_roi_extractor[0], head.bbox_roi_extractor[1]
    out['levels'] = ext0.map_roi_levels(rois, 4)
    xs = [t.clone().requires_grad_(True) for t in x[:4]]
    y = ext0(xs, rois)
    out['sle.out'] = y
    gx = torch.autograd.grad((y * seeded_like(y, 11)).sum(), xs, allow_unused=True)
    for l, g in enumerate(gx):
        out[f'sle.dx{l}'] = g if g is not None else torch.zeros_like(xs[l])
    yb = ext1(xs, pos_rois)
    out['ba.out'] = yb
    params = [ext1.conv1.weight, ext1.conv1.bias, ext1.conv2.weight]
    gb = torch.autograd.grad((yb * seeded_like(yb, 12)).sum(), xs + params)
    for l in range(4):
        out[f'ba.dx{l}'] = gb[l]
    out['ba.dconv1_w'], out['ba.dconv1_b'], out['ba.dconv2_w'] = gb[4], gb[5], gb[6]
    return out


def run_head(head, name, dtype, device='cpu'):
    """HTDBBoxHead forward/backward on seeded RoI features (golden item 5)."""
    c, x, props, gts, shapes = case_inputs(name, dtype, device)
    rois = _rois(props)
    pos_rois = _rois([p[:c['P']] for p in props])
    K, P = rois.size(0), pos_rois.size(0)
    pos_idx = torch.cat([torch.arange(c['P']) + i * c['K'] for i in range(c['B'])]).to(device)
    x_cls = (0.5 * seeded_like(torch.empty(K, 256, 7, 7, dtype=dtype, device=device), 21)
             ).requires_grad_(True)
    enh = (0.5 * seeded_like(torch.empty(P, 256, 7, 7, dtype=dtype, device=device), 22)
           ).requires_grad_(True)
    g = (0.5 * seeded_like(torch.empty(c['B'], 256, 1, 1, dtype=dtype, device=device), 23)
         ).requires_grad_(True)
    h1, fc0 = head.bbox_head[1], head.bbox_head[0].fc_cls
    cls_score, bbox_pred = h1(x_cls, x_cls[pos_idx], x[:4], rois, fc0, enh, pos_rois, g)
    out = {'head.cls_score': cls_score, 'head.bbox_pred': bbox_pred}
    loss = (cls_score * seeded_like(cls_score, 24)).sum() + \
        (bbox_pred * seeded_like(bbox_pred, 25)).sum()
    names, params = zip(*[(n, p) for n, p in list(h1.named_parameters()) +
                          [('fc0.' + n, p) for n, p in fc0.named_parameters()]])
    grads = torch.autograd.grad(loss, [x_cls, enh, g] + list(params), allow_unused=True)
    out['head.dx_cls'], out['head.denh'], out['head.dg'] = grads[:3]
    for n, gr, p in zip(names, grads[3:], params):
        out['head.d.' + n] = gr if gr is not None else torch.zeros_like(p)
    return out


def run_train(head, train_fn, test_fn, name, dtype, device='cpu'):
    """Full sampled forward_train (losses + gradients) and simple_test scores (item 6)."""
    c, x, props, gts, shapes = case_inputs(name, dtype, device)
    xs = [t.clone().requires_grad_(True) for t in x]
    losses = train_fn(head, xs, props, gts, shapes, c['P'])
    out = {'train.' + k: v.reshape(1) for k, v in losses.items()}
    total = sum(v for k, v in losses.items() if 'loss' in k)
    named = [(n, p) for n, p in head.named_parameters()]
    grads = torch.autograd.grad(total, xs + [p for _, p in named], allow_unused=True)
    for l in range(5):
        out[f'train.dx{l}'] = grads[l] if grads[l] is not None else torch.zeros_like(xs[l])
    for (n, p), gr in zip(named, grads[5:]):
        out['train.d.' + n] = gr if gr is not None else torch.zeros_like(p)
    with torch.no_grad():
        r, s, bp = test_fn(head, x, props, shapes)
    out['test.rois'], out['test.cls_score'], out['test.bbox_pred'] = r, s, bp
    return out


def aug_inputs(name, dtype=torch.float32, device='cpu'):
    """Test-time augmentation case built on case `name` (its first image): three views - plain,
    horizontally flipped, vertically flipped with another feature content - all at scale factor
    0.5 of the 'original' image, so that bbox_mapping / bbox_mapping_back do real work."""
    import numpy as np
    c, x, props, gts, shapes = case_inputs(name, dtype, device)
    sf = np.array([0.5, 0.5, 0.5, 0.5], dtype=np.float32)
    x0 = [t[:1] for t in x]
    x1 = [torch.flip(t[:1], dims=[3]).contiguous() for t in x]
    x2 = [(0.5 * torch.flip(t[-1:], dims=[2])).contiguous() for t in x]
    metas = [[dict(img_shape=shapes[0], scale_factor=sf, flip=False, flip_direction='horizontal')],
             [dict(img_shape=shapes[0], scale_factor=sf, flip=True, flip_direction='horizontal')],
             [dict(img_shape=shapes[0], scale_factor=sf, flip=True, flip_direction='vertical')]]
    proposals = props[0] / props[0].new_tensor(sf)          # 'original image' coordinates
    return [x0, x1, x2], proposals, metas


def run_aug(head, aug_fn, name, dtype, device='cpu'):
    feats, proposals, metas = aug_inputs(name, dtype, device)
    with torch.no_grad():
        out = aug_fn(head, feats, proposals, metas)
    return {'aug.bboxes': out[0], 'aug.scores': out[1]}


FPN_SIZES = ((37, 53), (19, 27), (10, 14), (5, 7))     # odd sizes: non-integer nearest scales


def fpn_fill_(fpn, seed=5):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in sorted(fpn.named_parameters()):
            std = 0.1 if n.endswith('bias') else (2.0 / p[0].numel()) ** 0.5
            p.copy_((std * torch.randn(p.shape, generator=g)).to(p.dtype))
    return fpn


def fpn_inputs(dtype=torch.float32, device='cpu', B=2, in_channels=(256, 512, 1024, 2048)):
    return [seeded_like(torch.empty(B, c, h, w), 70 + i).to(dtype).to(device)
            for i, (c, (h, w)) in enumerate(zip(in_channels, FPN_SIZES))]


def run_fpn(fpn, dtype, device='cpu'):
    """FPN forward (five outputs) and the gradients of a seeded linear loss w.r.t. inputs / weights."""
    xs = [x.requires_grad_(True) for x in fpn_inputs(dtype, device)]
    outs = fpn(xs)
    res = {f'fpn.out{i}': o for i, o in enumerate(outs)}
    loss = sum((o.float() * seeded_like(o, 90 + i).float()).sum() for i, o in enumerate(outs))
    named = sorted(fpn.named_parameters())
    grads = torch.autograd.grad(loss, xs + [p for _, p in named])
    for i in range(len(xs)):
        res[f'fpn.dx{i}'] = grads[i]
    for (n, _), g in zip(named, grads[len(xs):]):
        res['fpn.d.' + n] = g
    return res


RPN_CASES = {
    # name: image (h, w), nms_pre, nms_post, nms_thr, min_bbox_size, seed
    'train': dict(hw=(200, 304), nms_pre=2000, nms_post=2000, nms_thr=0.7, min_bbox_size=0, seed=11),
    'test': dict(hw=(160, 256), nms_pre=1000, nms_post=1000, nms_thr=0.7, min_bbox_size=0, seed=12),
    'minsize': dict(hw=(96, 128), nms_pre=300, nms_post=100, nms_thr=0.5, min_bbox_size=12, seed=13),
}


def rpn_inputs(name, dtype=torch.float32, device='cpu'):
    """RPN head outputs of one image on the five levels: objectness logits [3, H, W] (distinct
    values: a random permutation of an evenly spaced grid, so the ranking has no ties) and box
    deltas [12, H, W]."""
    c = RPN_CASES[name]
    H, W = c['hw']
    g = torch.Generator().manual_seed(c['seed'])
    cls, reg = [], []
    for s in (4, 8, 16, 32, 64):
        h, w = -(-H // s), -(-W // s)
        n = 3 * h * w
        vals = torch.linspace(-6.0, 4.0, n, dtype=torch.float64)[torch.randperm(n, generator=g)]
        cls.append(vals.view(3, h, w).to(dtype).to(device))
        reg.append((0.35 * torch.randn(12, h, w, generator=g)).to(dtype).to(device))
    return cls, reg, (H, W, 3), dict(nms_across_levels=False, nms_pre=c['nms_pre'],
                                     nms_post=c['nms_post'], max_num=c['nms_post'],
                                     nms_thr=c['nms_thr'], min_bbox_size=c['min_bbox_size'])


# ------------------------------------------------------------------ fixtures
def summarize(t, nsample=1024):
    t = t.detach().to('cpu')
    if t.dtype in (torch.int64, torch.int32, torch.uint8, torch.bool, torch.int8):
        return dict(kind='int', shape=np.array(t.shape), full=t.numpy().astype(np.int64))
    t = t.to(torch.float64).reshape(-1)
    n = t.numel()
    idx = np.unique(np.linspace(0, max(n - 1, 0), num=min(nsample, n)).astype(np.int64))
    return dict(kind='float', n=np.array(n), idx=idx, sample=t.numpy()[idx],
                sum=np.array(t.sum().item()), abssum=np.array(t.abs().sum().item()),
                maxabs=np.array(t.abs().max().item() if n else 0.0))


def save_fixture(path, outs):
    flat = {}
    for name, t in outs.items():
        for k, v in summarize(t).items():
            flat[f'{name}|{k}'] = np.asarray(v)
    np.savez_compressed(path, **flat)


def load_fixture(path):
    z = np.load(path, allow_pickle=False)
    d = {}
    for key in z.files:
        name, k = key.split('|')
        d.setdefault(name, {})[k] = z[key]
    return d


def rel_err(a, b):
    """max|a-b| / max|b| - the parity metric of SURVEY F12."""
    a = a.detach().to('cpu', torch.float64)
    b = b.detach().to('cpu', torch.float64)
    den = max(b.abs().max().item(), 1e-9) if b.numel() else 1.0
    return ((a - b).abs().max().item() / den) if b.numel() else 0.0


def compare_to_fixture(outs, fix, tol, names=None, sum_tol=None, floor=1e-9):
    """Returns {name: err}; raises AssertionError listing every tensor above ``tol``."""
    errs, bad = {}, []
    for name, t in outs.items():
        if names is not None and name not in names:
            continue
        f = fix[name]
        kind = str(f['kind'])
        t = t.detach().to('cpu')
        if kind == 'int':
            ok = tuple(f['shape']) == tuple(t.shape) and np.array_equal(
                f['full'], t.numpy().astype(np.int64))
            errs[name] = 0.0 if ok else float('inf')
        else:
            flat = t.to(torch.float64).reshape(-1).numpy()
            assert flat.size == int(f['n']), (name, flat.size, int(f['n']))
            den = max(float(f['maxabs']), floor)   # ~0 tensors (e.g. d conv2.bias) compare absolutely
            e = float(np.abs(flat[f['idx']] - f['sample']).max() / den) if flat.size else 0.0
            if sum_tol is not None and flat.size:
                e = max(e, abs(float(flat.sum()) - float(f['sum'])) /
                        max(float(f['abssum']), 1e-30) / sum_tol * tol)
            errs[name] = e
        if not errs[name] <= tol:
            bad.append((name, errs[name]))
    assert not bad, f'parity failures (tol {tol}): {bad[:12]}'
    return errs


# ---------------------------------------------------------------------------------------------
# assign + sample cases: seeded inputs shared by the generator of
# tests/golden/assign_sample.npz (reference classes), the CPU tests of the restatement and the
# GPU tests of csrc/assign_sample.cu.
# ---------------------------------------------------------------------------------------------
def _rcnn_cfg(iou, num=512, mlq=False, min_pos=None, ub=-1, add_gt=True, frac=0.25):
    return dict(assigner=dict(type='MaxIoUAssigner', pos_iou_thr=iou, neg_iou_thr=iou,
                              min_pos_iou=iou if min_pos is None else min_pos,
                              match_low_quality=mlq, ignore_iof_thr=-1),
                sampler=dict(type='RandomSampler', num=num, pos_fraction=frac, neg_pos_ub=ub,
                             add_gt_as_proposals=add_gt))


ASSIGN_CASES = {
    # reference tests/test_assigner.py:14-36 (expected gt_inds [1,0,2,0]) and :66-83 (no gt)
    'kat': dict(kind='kat', cfg=dict(
        assigner=dict(type='MaxIoUAssigner', pos_iou_thr=0.5, neg_iou_thr=0.5),
        sampler=dict(type='RandomSampler', num=4, pos_fraction=0.5, neg_pos_ub=-1,
                     add_gt_as_proposals=False))),
    # configs/htd/htd_resnet50_1x.py:122-138: stage 0 on RPN-like proposals
    'stage0': dict(kind='synth', B=3, N=1000, G=16, gts=(7, 16, 1), jitter=0.25, near=0.15,
                   seed=11, cfg=_rcnn_cfg(0.5)),
    # :139-155: stage 1 on refined boxes - many positives, some proposals removed -> pad rows
    'stage1': dict(kind='synth', B=3, N=512, G=16, gts=(5, 9, 3), jitter=0.08, near=0.6, seed=12,
                   drop=0.05, cfg=_rcnn_cfg(0.6)),
    'nogt': dict(kind='synth', B=2, N=300, G=8, gts=(0, 2), jitter=0.2, near=0.2, seed=13,
                 cfg=_rcnn_cfg(0.5, num=128)),
    'mlq': dict(kind='synth', B=2, N=700, G=24, gts=(24, 11), jitter=0.5, near=0.1, seed=14,
                cfg=_rcnn_cfg(0.7, num=256, mlq=True, min_pos=0.3, ub=3, add_gt=False, frac=0.5)),
    'few': dict(kind='synth', B=2, N=100, G=4, gts=(3, 4), jitter=0.2, near=0.3, seed=15,
                cfg=_rcnn_cfg(0.5)),
    # keys quantised to 1/8: many ties, broken by candidate index
    'ties': dict(kind='synth', B=2, N=900, G=8, gts=(6, 8), jitter=0.2, near=0.3, seed=16,
                 quant=8, cfg=_rcnn_cfg(0.5, num=256)),
}


def assign_case_inputs(name):
    """Returns dict(props [B,N,4], valid [B,N] bool, gt_boxes [B,G,4], gt_labels [B,G],
    num_gt [B] int32, keys [B,G+N] fp32, cfg).  CPU tensors."""
    c = ASSIGN_CASES[name] if isinstance(name, str) else name
    if c['kind'] == 'kat':
        props = torch.tensor([[[0, 0, 10, 10], [10, 10, 20, 20], [5, 5, 15, 15], [32, 32, 38, 42]],
                              [[0, 0, 10, 10], [10, 10, 20, 20], [5, 5, 15, 15], [32, 32, 38, 42]]],
                             dtype=torch.float32)
        gt = torch.tensor([[[0, 0, 10, 9], [0, 10, 10, 19]], [[0, 0, 0, 0], [0, 0, 0, 0]]],
                          dtype=torch.float32)
        labels = torch.tensor([[2, 3], [0, 0]])
        return dict(props=props, valid=torch.ones(2, 4, dtype=torch.bool), gt_boxes=gt,
                    gt_labels=labels, num_gt=torch.tensor([2, 0], dtype=torch.int32),
                    keys=torch.linspace(0.1, 0.9, 12).view(2, 6), cfg=c['cfg'])
    g = torch.Generator().manual_seed(c['seed'])
    B, N, G = c['B'], c['N'], c['G']
    H, W = (float(v) for v in c.get('hw', (800, 1333)))
    gt = torch.zeros(B, G, 4)
    labels = torch.zeros(B, G, dtype=torch.long)
    props = torch.zeros(B, N, 4)
    for b in range(B):
        ng = c['gts'][b]
        wh = torch.exp(torch.rand(G, 2, generator=g) * 2.5 + 3.0)             # 20 .. 245 px
        ctr = torch.rand(G, 2, generator=g) * torch.tensor([W, H])
        box = torch.cat([ctr - wh / 2, ctr + wh / 2], 1)
        box[:, 0::2] = box[:, 0::2].clamp(0, W)
        box[:, 1::2] = box[:, 1::2].clamp(0, H)
        gt[b, :ng] = box[:ng]
        labels[b, :ng] = torch.randint(0, 80, (ng,), generator=g)
        p = synth.make_proposals(1, N, int(H), int(W), seed=c['seed'] * 100 + b,
                                 max_scale=min(800.0, H))[0]
        near = int(N * c['near']) if ng > 0 else 0
        if near:                                           # jittered copies of the gt boxes
            src = box[:ng][torch.randint(0, ng, (near,), generator=g)]
            swh = (src[:, 2:] - src[:, :2]).repeat(1, 2)
            q = src + torch.randn(near, 4, generator=g) * c['jitter'] * swh
            q = torch.stack([torch.min(q[:, 0], q[:, 2]), torch.min(q[:, 1], q[:, 3]),
                             torch.max(q[:, 0], q[:, 2]), torch.max(q[:, 1], q[:, 3])], 1)
            q[:, 0::2] = q[:, 0::2].clamp(0, W)
            q[:, 1::2] = q[:, 1::2].clamp(0, H)
            idx = torch.randperm(N, generator=g)[:near]
            p[idx] = q
            if ng > 1:                                     # exact duplicates: equal IoU maxima
                p[idx[0]] = p[idx[1]]
        props[b] = p
    valid = torch.ones(B, N, dtype=torch.bool)
    if c.get('drop'):
        valid = torch.rand(B, N, generator=g) >= c['drop']
    keys = torch.rand(B, G + N, generator=g)
    if c.get('quant'):
        keys = torch.floor(keys * c['quant']) / c['quant']
    return dict(props=props, valid=valid, gt_boxes=gt, gt_labels=labels,
                num_gt=torch.tensor(c['gts'], dtype=torch.int32), keys=keys, cfg=c['cfg'])


def assign_static_layout(per_image, num, B):
    """Lays per-image sampling results (namespaces with pos/neg fields in the reference's
    candidate index space) out like htd_assign_sample: ``num`` rows per image, positives,
    negatives, pad rows.  Integer tensors only (what must match bit for bit) + the boxes."""
    K = B * num
    out = dict(rois=torch.zeros(K, 5), kind=torch.full((K,), 2, dtype=torch.uint8),
               gt_boxes=torch.zeros(K, 4), gt_labels=torch.zeros(K, dtype=torch.long),
               gt_index=torch.full((K,), -1, dtype=torch.int32),
               is_gt=torch.zeros(K, dtype=torch.uint8),
               cand=torch.full((K,), -1, dtype=torch.int32),
               counts=torch.zeros(B, 4, dtype=torch.int32))
    for b, r in enumerate(per_image):
        o = b * num
        out['rois'][o:o + num, 0] = b
        npos, nneg = r.pos_inds.numel(), r.neg_inds.numel()
        out['rois'][o:o + npos + nneg, 1:] = r.bboxes
        out['kind'][o:o + npos] = 1
        out['kind'][o + npos:o + npos + nneg] = 0
        out['gt_boxes'][o:o + npos] = r.pos_gt_bboxes
        out['gt_labels'][o:o + npos] = r.pos_gt_labels
        out['gt_index'][o:o + npos] = r.pos_assigned_gt_inds.to(torch.int32)
        out['is_gt'][o:o + npos] = r.pos_is_gt
        out['cand'][o:o + npos + nneg] = r.cand.to(torch.int32)
        out['counts'][b] = torch.tensor([npos, nneg, r.npos_cand, r.nneg_cand], dtype=torch.int32)
    return out


def run_assign_case(name, image_fn):
    """``image_fn(bboxes[n,4], gt_bboxes[g,4], gt_labels[g], keys[g+n or n], cfg, valid[n])`` ->
    namespace (oracle/restate.assign_sample_image or the reference-backed equivalent)."""
    d = assign_case_inputs(name)
    B, N = d['props'].shape[:2]
    G = d['gt_boxes'].shape[1]
    add_gt = d['cfg']['sampler'].get('add_gt_as_proposals', True)
    res = []
    for b in range(B):
        g = int(d['num_gt'][b])
        keys = d['keys'][b]
        keys = torch.cat([keys[:g], keys[G:]]) if (add_gt and g > 0) else keys[G:]
        res.append(image_fn(d['props'][b], d['gt_boxes'][b, :g], d['gt_labels'][b, :g], keys,
                            d['cfg'], d['valid'][b]))
    return assign_static_layout(res, d['cfg']['sampler']['num'], B), res


# full training step with assign + sample inside (forward_train, htd_roi_head.py:217-317)
TRAIN_ASSIGNED = dict(kind='synth', B=2, N=200, G=6, gts=(4, 6), jitter=0.12, near=0.45, seed=17,
                      hw=(320, 448), pyramid_seed=1000, scheme='n005', wseed=0, num=64,
                      cfg=_rcnn_cfg(0.5, num=64))


def train_assigned_inputs(dtype=torch.float32, device='cpu', n_props=None):
    """x (pyramid), proposals [B,N,4], gt_boxes [B,G,4], gt_labels [B,G], num_gt [B], keys
    (two tensors [B,G+N], [B,G+num]), cfgs (two rcnn stage cfgs), img_shapes.  ``n_props`` keeps
    only the first proposals of every image (fewer than ``num``: pad rows in both stages)."""
    c = TRAIN_ASSIGNED
    d = assign_case_inputs(c)
    if n_props is not None:
        d['props'] = d['props'][:, :n_props].contiguous()
        d['keys'] = d['keys'][:, :c['G'] + n_props].contiguous()
    H, W = c['hw']
    x = [t.to(dtype).to(device) for t in synth.make_pyramid(c['B'], H, W, seed=c['pyramid_seed'])]
    g = torch.Generator().manual_seed(c['seed'] + 1)
    keys1 = torch.rand(c['B'], c['G'] + c['num'], generator=g)
    cfgs = [_rcnn_cfg(0.5, num=c['num']), _rcnn_cfg(0.6, num=c['num'])]
    return dict(x=x, props=d['props'], gt_boxes=d['gt_boxes'], gt_labels=d['gt_labels'],
                num_gt=d['num_gt'], keys=[d['keys'], keys1], cfgs=cfgs,
                img_shapes=[(H, W, 3)] * c['B'])


def run_train_assigned(train_fn, head, dtype=torch.float32, device='cpu', n_props=None):
    """``train_fn(head, x, proposals(list), gt_bboxes(list), gt_labels(list), keys, cfgs,
    img_shapes, G) -> (losses, info)``; returns the flat dict of losses, input / parameter
    gradients and the sampled candidate indices (the fixture content)."""
    d = train_assigned_inputs(dtype, n_props=n_props)
    B, G = d['props'].shape[0], d['gt_boxes'].shape[1]
    xs = [t.clone().to(device).requires_grad_(True) for t in d['x']]
    ng = [int(v) for v in d['num_gt']]
    losses, info = train_fn(head, xs, [d['props'][b].to(dtype).to(device) for b in range(B)],
                            [d['gt_boxes'][b, :ng[b]].to(dtype).to(device) for b in range(B)],
                            [d['gt_labels'][b, :ng[b]].to(device) for b in range(B)], d['keys'],
                            d['cfgs'], d['img_shapes'], G)
    head.zero_grad()
    sum(v for k, v in losses.items() if 'loss' in k).backward()
    out = {f'assigned.{k}': v.detach().reshape(-1).cpu() for k, v in losses.items()}
    for i, t in enumerate(xs):
        out[f'assigned.dx{i}'] = t.grad.cpu()
    seen = set()
    for k, p in head.named_parameters():
        if p.grad is not None and id(p) not in seen:
            seen.add(id(p))
            out[f'assigned.grad.{k}'] = p.grad.cpu()
    for st in (0, 1):
        for b, r in enumerate(info[f'samp{st}']):
            out[f'assigned.s{st}.cand{b}'] = r.cand.to(torch.int32).cpu() if hasattr(r, 'cand') \
                else torch.cat([r.pos_inds, r.neg_inds]).to(torch.int32).cpu()
            out[f'assigned.s{st}.npos{b}'] = torch.tensor([r.pos_bboxes.size(0)], dtype=torch.int32)
    out['assigned.refined'] = torch.cat(info['refined']).cpu()
    return out


# ---------------------------------------------------------------------------------------------
# multi-class NMS cases
# ---------------------------------------------------------------------------------------------
NMS_CASES = {
    # configs/htd/htd_resnet50_1x.py:164-168: 1000 RoIs, 80 classes, class-agnostic boxes
    'htd': dict(K=1000, C=80, per_class=False, score_thr=0.05, iou_thr=0.5, max_num=100, seed=41,
                dup=0.5, temp=3.0),
    'perclass': dict(K=300, C=20, per_class=True, score_thr=0.05, iou_thr=0.5, max_num=100, seed=42,
                     dup=0.4, temp=2.5),
    'dense_all': dict(K=400, C=6, per_class=False, score_thr=0.01, iou_thr=0.3, max_num=-1, seed=43,
                      dup=0.9, temp=1.0),
    'empty': dict(K=64, C=80, per_class=False, score_thr=0.9, iou_thr=0.5, max_num=100, seed=44,
                  dup=0.0, temp=0.1),
    'ties': dict(K=500, C=10, per_class=False, score_thr=0.05, iou_thr=0.5, max_num=200, seed=45,
                 dup=0.6, temp=2.0, quant=32),
}


def nms_case_inputs(name):
    """(multi_bboxes [K,4] or [K,C*4], multi_scores [K,C+1]) on the CPU + the case dict."""
    c = NMS_CASES[name]
    g = torch.Generator().manual_seed(c['seed'])
    K, C = c['K'], c['C']
    ctr = torch.rand(K, 2, generator=g) * torch.tensor([1333.0, 800.0])
    wh = torch.exp(torch.rand(K, 2, generator=g) * 2.5 + 3.0)
    boxes = torch.cat([ctr - wh / 2, ctr + wh / 2], 1)
    nd = int(K * c['dup'])
    if nd:                                             # clusters of near-duplicates
        src = boxes[torch.randint(0, K - nd, (nd,), generator=g)]
        boxes[K - nd:] = src + torch.randn(nd, 4, generator=g) * 0.06 * (src[:, 2:] - src[:, :2]).repeat(1, 2)
    lim = torch.tensor([1333.0, 800.0, 1333.0, 800.0])
    boxes = torch.min(boxes.clamp(min=0), lim)
    scores = torch.softmax(torch.randn(K, C + 1, generator=g) * c['temp'], 1)
    if c.get('quant'):
        scores = torch.floor(scores * c['quant']) / c['quant']
    if c['per_class']:
        boxes = (boxes[:, None, :] + torch.randn(K, C, 4, generator=g) * 2.0).reshape(K, C * 4)
        boxes = torch.min(boxes.clamp(min=0), lim.repeat(C))
    return boxes, scores, c
