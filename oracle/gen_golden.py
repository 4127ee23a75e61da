"""ORACLE (test infrastructure).  Generates tests/golden/*.npz by running the REFERENCE's own
unmodified modules (oracle/refshim.py, reading /root/reference in place) on the seeded cases
of oracle/cases.py.  Run here (the authoring container); the fixtures travel, the reference
does not.

    python -m oracle.gen_golden            # all cases, fp64 + fp32
    python -m oracle.gen_golden assign     # only tests/golden/assign_sample.npz
    python -m oracle.gen_golden case c2    # only the module / head / training fixtures of one case

Fixtures hold, per tensor, a strided sample of <=1024 elements plus sum / abs-sum / max-abs
(float tensors) or the full tensor (integer tensors: level indices, mask bitsets, degrees).
"""
import os
import sys

import numpy as np
import torch

from htd_b200 import synth
from . import cases, ref_driver, refshim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def boundary_rois():
    """RoIs whose sqrt(w*h)/56 + 1e-6 sits on / one ulp around the level boundaries
    (scale 112, 224, 448), plus degenerate (zero-area) and huge boxes."""
    rows = []
    for s in (112.0, 224.0, 448.0, 56.0, 896.0):
        base = np.float32(s)
        for k in range(-3, 4):
            w = base
            for _ in range(abs(k)):
                w = np.nextafter(w, np.float32(np.inf if k > 0 else -np.inf), dtype=np.float32)
            rows.append([0.0, 10.0, 20.0, 10.0 + float(w), 20.0 + float(base)])
            rows.append([0.0, 0.0, 0.0, float(w), float(w)])
    rows += [[0, 5, 5, 5, 5], [0, 5, 5, 5, 50], [0, 0, 0, 1333, 800], [0, 3, 4, 3.5, 4.25],
             [0, 100, 100, 100 + 111.99999, 100 + 112.00001]]
    return torch.tensor(rows, dtype=torch.float32)


def gen_assign_sample():
    """tests/golden/assign_sample.npz: the reference's MaxIoUAssigner + RandomSampler (keys in place
    of randperm) on oracle/cases.ASSIGN_CASES, in the static layout of htd_assign_sample."""
    os.makedirs(OUT, exist_ok=True)
    flat = {}
    for name in cases.ASSIGN_CASES:
        out, res = cases.run_assign_case(name, ref_driver.ref_assign_sample_image)
        for k, v in out.items():
            flat[f'{name}|{k}'] = v.numpy()
        for b, r in enumerate(res):
            flat[f'{name}|gt_inds{b}'] = r.gt_inds.numpy().astype(np.int32)
            flat[f'{name}|max_overlaps{b}'] = r.max_overlaps.numpy()
        print('assign', name, out['counts'].tolist())
    np.savez_compressed(os.path.join(OUT, 'assign_sample.npz'), **flat)
    # the reference's whole forward_train (its own assigners + samplers, keyed) on a small case
    c = cases.TRAIN_ASSIGNED
    head = refshim.build_head()
    synth.fill_params_(head, c['scheme'], c['wseed'])
    outs = cases.run_train_assigned(ref_driver.ref_forward_train_assigned, head)
    path = os.path.join(OUT, 'train_assigned_f32.npz')
    cases.save_fixture(path, outs)
    print('train_assigned', len(outs), 'tensors ->', path, os.path.getsize(path) // 1024, 'KiB',
          {k: float(v) for k, v in outs.items() if v.numel() == 1 and v.is_floating_point()})


def gen_nms():
    """tests/golden/nms.npz: the reference's own multiclass_nms (bbox_nms.py:7-71; its mmcv
    batched_nms realised by torchvision.ops.nms in oracle/refshim.py) on oracle/cases.NMS_CASES."""
    ns = refshim.load()
    flat = {}
    for name in cases.NMS_CASES:
        boxes, scores, c = cases.nms_case_inputs(name)
        dets, labels = ns.multiclass_nms(boxes, scores, c['score_thr'],
                                         dict(type='nms', iou_threshold=c['iou_thr']), c['max_num'])
        flat[f'{name}|dets'] = dets.numpy()
        flat[f'{name}|labels'] = labels.numpy()
        print('nms', name, tuple(dets.shape))
    np.savez_compressed(os.path.join(OUT, 'nms.npz'), **flat)
    # soft-NMS (configs/htd/htd_resnet101_2x.py:298): the reference's multiclass_nms with its own
    # nms_cfg; mmcv's soft_nms op = the literal restatement oracle/soft_nms_ref.c (mmcv is absent)
    flat = {}
    for name in cases.NMS_CASES:
        boxes, scores, c = cases.nms_case_inputs(name)
        dets, labels = ns.multiclass_nms(boxes, scores, c['score_thr'],
                                         dict(type='soft_nms', iou_thr=c['iou_thr'],
                                              min_score=c['score_thr']), c['max_num'])
        flat[f'{name}|dets'] = dets.numpy()
        flat[f'{name}|labels'] = labels.numpy()
        print('soft_nms', name, tuple(dets.shape))
    np.savez_compressed(os.path.join(OUT, 'nms_soft.npz'), **flat)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == 'assign':
        return gen_assign_sample()
    if len(sys.argv) > 1 and sys.argv[1] == 'nms':
        return gen_nms()
    os.makedirs(OUT, exist_ok=True)
    ns = refshim.load()
    torch.set_num_threads(os.cpu_count())
    ext = ns.SingleRoIExtractor(dict(type='RoIAlign', output_size=7, sampling_ratio=0), 256,
                                [4, 8, 16, 32])
    only = sys.argv[2:] if len(sys.argv) > 2 and sys.argv[1] == 'case' else None
    if only is None:
        gen_levels(ext)
    gen_cases(ns, ext, only)
    if only is None:
        gen_assign_sample()
        gen_nms()
        gen_fpn(ns)
        gen_rpn()


def gen_fpn(ns):
    # --- FPN neck: the reference's own FPN, fp64 ---------------------------------
    fpn = cases.fpn_fill_(ns.FPN([256, 512, 1024, 2048], 256, 5).double())
    outs = cases.run_fpn(fpn, torch.float64)
    cases.save_fixture(os.path.join(OUT, 'fpn_f64.npz'), outs)
    print('fpn:', len(outs), 'tensors')


def gen_rpn():
    # --- RPN proposals: the reference's own _get_bboxes_single -------------------
    flat = {}
    for name in cases.RPN_CASES:
        cls, reg, shape, cfg = cases.rpn_inputs(name)
        det = ref_driver.ref_rpn_proposals(cls, reg, shape, cfg)
        flat[name] = det.numpy()
        print('rpn', name, tuple(det.shape))
    np.savez_compressed(os.path.join(OUT, 'rpn_proposals.npz'), **flat)


def gen_levels(ext):
    # --- level assignment (bit-exact integer fixture) --------------------------------------
    props = synth.make_proposals(8, 512, seed=99)
    rois = torch.cat([torch.cat([p.new_full((p.size(0), 1), i), p], 1)
                      for i, p in enumerate(props)], 0)
    rois = torch.cat([rois, boundary_rois()], 0)
    lv = ext.map_roi_levels(rois, 4)
    np.savez_compressed(os.path.join(OUT, 'levels.npz'), rois=rois.numpy(),
                        levels=lv.numpy().astype(np.int8))
    print('levels:', torch.bincount(lv).tolist())


def gen_cases(ns, ext, only=None):
    names = [n for n in cases.CASES if only is None or n in only]
    # --- graph masks (bit-exact) --------------------------------------------------------------
    for name in names:
        c, x, pr, gts, shapes = cases.case_inputs(name)
        r = cases._rois(pr)
        masks = cases.graph_masks(ns.bbox_overlaps, ext.map_roi_levels(r, 4), r)
        flat = {}
        for (b, i), (idx, M, deg) in masks.items():
            flat[f'{b}_{i}|idx'] = idx.numpy()
            flat[f'{b}_{i}|bits'] = np.packbits(M.numpy().astype(np.uint8), axis=1)
            flat[f'{b}_{i}|deg'] = deg.numpy().astype(np.int32)
        np.savez_compressed(os.path.join(OUT, f'masks_{name}.npz'), **flat)
    # --- module / head / training fixtures ----------------------------------------------------
    for name in names:
        c = cases.CASES[name]
        for dt, tag in ((torch.float64, 'f64'), (torch.float32, 'f32')):
            head = refshim.build_head(double=(dt == torch.float64))
            synth.fill_params_(head, c['scheme'], c['seed'])
            outs = {}
            outs.update(cases.run_extractors(head, name, dt))
            outs.update(cases.run_head(head, name, dt))
            outs.update(cases.run_train(head, ref_driver.ref_forward_train_sampled,
                                        ref_driver.ref_simple_test_scores, name, dt))
            if name == 'small':                      # test-time augmentation: the reference's aug_test
                aug = cases.run_aug(head, lambda h, *a: ref_driver.ref_aug_test(h, *a)[:2], name, dt)
                cases.save_fixture(os.path.join(OUT, f'aug_{name}_{tag}.npz'), aug)
            path = os.path.join(OUT, f'{name}_{tag}.npz')
            cases.save_fixture(path, outs)
            print(name, tag, len(outs), 'tensors ->', path, os.path.getsize(path) // 1024, 'KiB',
                  {k: float(v) for k, v in outs.items() if k.startswith('train.') and v.numel() == 1})


if __name__ == '__main__':
    sys.exit(main())
