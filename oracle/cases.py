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
