"""HTDRoIHead - host-side mirror of ``mmdet/models/roi_heads/htd_roi_head.py``.

Two stages plus SFA:  stage 0 = SingleRoIExtractor + Shared2FCBBoxHead (common head), box
refinement, stage 1 = SingleRoIExtractor (cls) + AdptRoIExtractor/BA (reg) + HTDBBoxHead
(PGraph).  Same constructor arguments, method names, result-dict keys and loss keys as the
reference.  Differences (documented in DESIGN.md):
  * the pyramid is converted ONCE per call to the channels-last layout the kernels read and
    shared by the three extractor calls (the reference re-reads NCHW maps 13 times);
  * ``_fuse_global`` (htd_roi_head.py:133-141) is folded into the stage-0 extraction launch;
  * stage-1 training gathers the positive prefix of EVERY image's block (the reference
    hard-codes <= 2 images per GPU, htd_roi_head.py:157-170,180-182 - same result for B <= 2).
"""
import torch
import torch.nn as nn

from . import ops
from .core import (as_cfg, bbox2result, bbox2roi, bbox_mapping, merge_aug_bboxes,
                   multiclass_nms)
from .registry import HEADS, build_assigner, build_head, build_roi_extractor, build_sampler


class _StaticRows:
    """Sampling-result view of one image of a static sample: the regression slots and the rest."""
    __slots__ = ('pos_bboxes', 'neg_bboxes')

    def __init__(self, pos_bboxes, neg_bboxes):
        self.pos_bboxes, self.neg_bboxes = pos_bboxes, neg_bboxes



def _hand_over(obj, stream, depth=0):
    """Tensors made on a branch stream and used on ``stream`` afterwards: tell the allocator
    (the join itself is a ``wait_stream``; this is about when their memory may be reused)."""
    if torch.is_tensor(obj):
        if obj.is_cuda:
            obj.record_stream(stream)
    elif isinstance(obj, dict):
        for v in obj.values():
            _hand_over(v, stream, depth + 1)
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            _hand_over(v, stream, depth + 1)
    elif hasattr(obj, '__dict__') and depth < 3:
        _hand_over(vars(obj), stream, depth + 1)


class _Branch:
    """``with _Branch(stream, uses) as b:`` runs the body on ``stream`` after everything issued so
    far on the current stream; ``b.join(results)`` makes the current stream wait for it.
    ``stream=None``: no-op (the body runs in line)."""

    def __init__(self, stream, uses=()):
        self.stream, self.uses = stream, uses

    def __enter__(self):
        if self.stream is not None:
            self.cur = torch.cuda.current_stream(self.stream.device)
            self.stream.wait_stream(self.cur)
            _hand_over(self.uses, self.stream)      # saved for a backward that runs over there
            self.ctx = torch.cuda.stream(self.stream)
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.stream is not None:
            self.ctx.__exit__(*exc)
        return False

    def join(self, results):
        if self.stream is not None:
            self.cur.wait_stream(self.stream)
            _hand_over(results, self.cur)
        return results


@HEADS.register_module()
class HTDRoIHead(nn.Module):

    def __init__(self, num_stages, stage_loss_weights, with_global=False, bbox_roi_extractor=None,
                 bbox_head=None, mask_roi_extractor=None, mask_head=None, shared_head=None,
                 train_cfg=None, test_cfg=None, compute_dtype=None):
        super().__init__()
        assert bbox_roi_extractor is not None and bbox_head is not None
        assert shared_head is None, 'Shared head is not supported'
        if mask_head is not None or mask_roi_extractor is not None:
            raise NotImplementedError('configs/htd define no mask branch')
        self.num_stages = num_stages
        self.stage_loss_weights = stage_loss_weights
        self.with_global = with_global
        self.train_cfg = as_cfg(train_cfg)
        self.test_cfg = as_cfg(test_cfg)
        self.compute_dtype = compute_dtype       # dtype of the channels-last pyramid copy
        self.init_bbox_head(bbox_roi_extractor, bbox_head)
        self.init_assigner_sampler()

    with_bbox, with_mask, with_shared_head = True, False, False

    def init_bbox_head(self, bbox_roi_extractor, bbox_head):
        """htd_roi_head.py:42-71."""
        self.bbox_roi_extractor = nn.ModuleList()
        self.bbox_head = nn.ModuleList()
        if not isinstance(bbox_roi_extractor, list):
            bbox_roi_extractor = [bbox_roi_extractor for _ in range(self.num_stages)]
        if not isinstance(bbox_head, list):
            bbox_head = [bbox_head for _ in range(self.num_stages)]
        assert len(bbox_roi_extractor) == len(bbox_head) == self.num_stages
        for ext, head in zip(bbox_roi_extractor, bbox_head):
            self.bbox_roi_extractor.append(build_roi_extractor(ext))
            self.bbox_head.append(build_head(head))
        if self.with_global:
            self.glbctx_head = build_head(dict(
                type='GlobalContextHead', num_ins=5, num_convs=4, in_channels=256,
                conv_out_channels=256, num_classes=self.bbox_head[0].num_classes + 1,
                loss_weight=3.0))

    def init_assigner_sampler(self):
        """htd_roi_head.py:101-111."""
        self.bbox_assigner, self.bbox_sampler = [], []
        if self.train_cfg is not None:
            for cfg in self.train_cfg:
                self.bbox_assigner.append(build_assigner(dict(cfg.assigner)))
                self.bbox_sampler.append(build_sampler(dict(cfg.sampler), context=self))

    def init_weights(self, pretrained=None):
        for i in range(self.num_stages):
            self.bbox_roi_extractor[i].init_weights()
            self.bbox_head[i].init_weights()
        if self.with_global:
            self.glbctx_head.init_weights()

    # ------------------------------------------------------------------------------------------
    # Independent parts of the step run as parallel branches (streams in eager mode, parallel
    # branches of the graph once captured; autograd mirrors every branch in backward):
    #   overlap_ba      BA extraction + the reg branch that consumes it, next to the single-level
    #                   extraction, the cls-branch GEMMs and the PGraph
    #   overlap_global  global-context head next to the pyramid conversion (forward) and the
    #                   backward gather (backward), on a prioritised stream: its launches are tiny
    #                   and would otherwise queue behind the thousands of CTAs of those kernels
    #   overlap_stages  stage 0 on its own stream: forward is serial (stage 1 samples the refined
    #                   boxes), but the two stages' backward passes are independent
    # Class switches (diagnostics; ``overlap = False`` turns all of them off).
    overlap = True
    overlap_ba = True
    overlap_global = True
    overlap_stages = True
    inputs_consumed_event = None   # optional CUDA event recorded once forward_train has read `x`

    def _stream(self, name, device, priority=0):
        pool = self.__dict__.setdefault('_streams', {})
        st = pool.get(name)
        if st is None or st.device != device:
            st = pool[name] = torch.cuda.Stream(device=device, priority=priority)
        return st

    def _side_stream(self, device):
        return self._stream('side', device)

    def _on(self, which, t):
        return self.overlap and getattr(self, which) and t.is_cuda

    def _global_context(self, x, loss_fn, defer=False):
        """Global-context head + its loss (htd_roi_head.py:245-249) next to the channels-last
        conversion of the pyramid: both only read ``x``, the head's launches are small
        (2 x 256 x 13 x 21 maps), so on a side stream it costs nothing on the forward critical
        path, and autograd mirrors the branch in backward (next to the backward gather).
        Returns (channels-last pyramid, loss, global_feat, join).  ``defer``: the current stream
        does NOT wait for the branch here; ``global_feat.ready`` is an event a consumer stream
        waits on before it reads the vector (stage 0 does so after its extraction), and ``join()``
        makes the current stream wait for the whole branch (vector and loss)."""
        if not self._on('overlap_global', x[0]):
            x_cl = self._pyramid(x)
            mc_pred, global_feat = self.glbctx_head(x)
            return x_cl, loss_fn(mc_pred), global_feat, lambda: None
        cur, side = torch.cuda.current_stream(), self._stream('global', x[0].device, priority=-1)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            mc_pred, global_feat = self.glbctx_head(x)
            ready = torch.cuda.Event()
            ready.record()
            loss = loss_fn(mc_pred)
        x_cl = self._pyramid(x)

        def join():
            cur.wait_stream(side)
            for t in (loss, global_feat):
                t.record_stream(cur)
        if defer:
            global_feat.ready = ready
        else:
            join()
        return x_cl, loss, global_feat, join

    def _pyramid(self, x):
        """Channels-last copy of the levels the extractors read, made once per call."""
        n = self.bbox_roi_extractor[0].num_inputs
        dtype = self.compute_dtype or x[0].dtype
        return ops.make_pyramid(list(x[:n]), dtype)

    def _fuse_global(self, roi_feats, global_feat, rois):
        """htd_roi_head.py:133-141 as a broadcast add (equal whenever every image index is valid)."""
        assert roi_feats.size(0) == rois.size(0)
        g = global_feat.reshape(global_feat.size(0), -1)
        return roi_feats + g[rois[:, 0].long()][:, :, None, None].to(roi_feats.dtype)

    def _bbox_forward(self, stage, x, rois, global_feat=None, sampling_results=None,
                      img_metas=None, x_cl=None, row_valid=None):
        """htd_roi_head.py:143-201.  ``row_valid`` ([K] bool, optional): False marks the pad rows of
        the static-shape sampler; they are kept out of the PGraph groups."""
        ext, enh = self.bbox_roi_extractor[0], self.bbox_roi_extractor[1]
        if x_cl is None:
            x_cl = self._pyramid(x)
        g = global_feat if self.with_global else None
        if stage == 0:
            late = getattr(global_feat, 'ready', None) if g is not None else None
            if late is not None and ops.flatten_fuses_bias_shape(ext.out_channels,
                                                                  ext.roi_layers[0].output_size):
                # the global-context head is still running on its own stream: extract without the
                # SFA vector, wait for it only now, and add it while the features are flattened
                # for the FCs (same sum, one more rounding in bf16); `bbox_feats` is returned
                # WITHOUT the vector in this mode
                bbox_feats = ext(x_cl, rois)
                torch.cuda.current_stream().wait_event(late)
                cls_score, bbox_pred = self.bbox_head[0](bbox_feats, sfa_bias=g, rois=rois)
                return dict(cls_score=cls_score, bbox_pred=bbox_pred, bbox_feats=bbox_feats)
            if late is not None:
                torch.cuda.current_stream().wait_event(late)
            bbox_feats = ext(x_cl, rois, bias=g)          # RoIAlign + SFA add in one launch
            cls_score, bbox_pred = self.bbox_head[0](bbox_feats)
            return dict(cls_score=cls_score, bbox_pred=bbox_pred, bbox_feats=bbox_feats)
        head = self.bbox_head[stage]
        fc0 = self.bbox_head[0].fc_cls
        nimg = g.size(0) if g is not None else None
        if sampling_results:
            pos_rois = bbox2roi([res.pos_bboxes for res in sampling_results])
            reg_stream = None
            if self._on('overlap_ba', rois):
                # BA extraction (plan, 4-level gather, attention, fuse: latency-bound kernels) and
                # the reg branch that consumes it (conv tower) on a side stream next to the
                # single-level extraction, the cls-branch GEMMs and the PGraph; the head joins
                # the stream once both branches are issued.  Inside a CUDA graph this becomes a
                # parallel branch; autograd mirrors it in backward.
                cur, reg_stream = torch.cuda.current_stream(), self._side_stream(rois.device)
                reg_stream.wait_stream(cur)
                with torch.cuda.stream(reg_stream):
                    enhanced = enh(x_cl, pos_rois)
            else:
                enhanced = enh(x_cl, pos_rois)
            bbox_feats = ext(x_cl, rois)
            # positives are the prefix of each image's block: slices, not index tensors, so the
            # gather and the scatter of bbox_pred (htd_roi_head.py:169-170,180-182) are plain copies
            spans, off = [], 0
            for res in sampling_results:
                spans.append((off, res.pos_bboxes.size(0), res.neg_bboxes.size(0)))
                off += res.pos_bboxes.size(0) + res.neg_bboxes.size(0)
            # one autograd node for both uses of bbox_feats (flattened cls input, positive rows)
            flat, pos_feats = ops.flatten_with_prefix(bbox_feats, [(o, n_) for o, n_, _ in spans])
            cls_score, bbox_pred = head(bbox_feats, pos_feats, x_cl, rois, fc0, enhanced,
                                        pos_rois, g, num_imgs=nimg or len(sampling_results),
                                        max_rois_per_img=max(p_ + n_ for _, p_, n_ in spans),
                                        row_valid=row_valid, x_cls_flat=flat, reg_stream=reg_stream)
            parts, o2 = [], 0
            for _, npos, nneg in spans:
                parts += [bbox_pred[o2:o2 + npos], bbox_pred.new_zeros(nneg, bbox_pred.size(1))]
                o2 += npos
            return dict(cls_score=cls_score, bbox_pred=torch.cat(parts, 0))
        bbox_feats = ext(x_cl, rois)
        enhanced = enh(x_cl, rois)
        cls_score, bbox_pred = head(bbox_feats, bbox_feats, x_cl, rois, fc0, enhanced, rois, g,
                                    num_imgs=nimg)
        return dict(cls_score=cls_score, bbox_pred=bbox_pred)

    def _bbox_forward_train(self, stage, x, sampling_results, gt_bboxes, gt_labels, rcnn_train_cfg,
                            img_metas, global_feat=None, x_cl=None):
        """htd_roi_head.py:203-215."""
        rois = bbox2roi([res.bboxes for res in sampling_results])
        res = self._bbox_forward(stage, x, rois, global_feat, sampling_results, img_metas, x_cl)
        targets = self.bbox_head[stage].get_targets(sampling_results, gt_bboxes, gt_labels,
                                                    rcnn_train_cfg)
        loss = self.bbox_head[stage].loss(res['cls_score'], res['bbox_pred'], rois, *targets)
        res.update(loss_bbox=loss, rois=rois, bbox_targets=targets)
        return res

    def _assign_sample(self, stage, proposal_list, gt_bboxes, gt_labels, gt_bboxes_ignore, x):
        out = []
        for j in range(len(proposal_list)):
            a = self.bbox_assigner[stage].assign(proposal_list[j], gt_bboxes[j],
                                                 gt_bboxes_ignore[j], gt_labels[j])
            out.append(self.bbox_sampler[stage].sample(a, proposal_list[j], gt_bboxes[j],
                                                       gt_labels[j]))
        return out

    def forward_train(self, x, img_metas, proposal_list, gt_bboxes, gt_labels,
                      gt_bboxes_ignore=None, gt_masks=None, sampling_fn=None):
        """htd_roi_head.py:217-317.  ``sampling_fn(stage, proposal_list) -> list of sampling
        results`` replaces the random assign+sample steps when given (deterministic benches and
        parity tests; the reference's RandomSampler draws from the global torch RNG)."""
        losses = dict()
        num_imgs = len(img_metas)
        if gt_bboxes_ignore is None:
            gt_bboxes_ignore = [None] * num_imgs
        x_cl = None if self.with_global else self._pyramid(x)

        def sample(stage, props):
            if sampling_fn is not None:
                return sampling_fn(stage, props)
            return self._assign_sample(stage, props, gt_bboxes, gt_labels, gt_bboxes_ignore, x)

        if self.inputs_consumed_event is not None:
            proposal_list = [p.clone() for p in proposal_list]   # private copy: see the event below
        samp = sample(0, proposal_list)
        global_feat = None
        st0 = self._stream('stage0', x[0].device) if self._on('overlap_stages', x[0]) else None
        join_global = lambda: None
        if self.with_global:
            x_cl, losses['loss_global'], global_feat, join_global = self._global_context(
                x, lambda mc_pred: self.glbctx_head.loss(mc_pred, gt_labels), defer=st0 is not None)

        def inputs_consumed():
            if self.inputs_consumed_event is not None and x[0].is_cuda:
                # everything downstream reads the channels-last copy / the SFA head's cast of P6
                # only (and the proposals through copies made above): the caller may refill `x`
                # from here on - graphed.GraphedTrainStep(flat_inputs=True) overlaps the next
                # upload with this step
                self.inputs_consumed_event.record()
        if st0 is None:
            inputs_consumed()
        with _Branch(st0, (list(x_cl), global_feat)) as br:
            res = self._bbox_forward_train(0, x, samp, gt_bboxes, gt_labels, self.train_cfg[0],
                                           img_metas, global_feat, x_cl)
            # the loss weights belong to the branch too: their backward nodes are the entry of
            # stage 0's backward, and on the main stream they would be queued behind all of
            # stage 1's (autograd issues later forward ops first)
            lw = self.stage_loss_weights[0]
            for name, value in res['loss_bbox'].items():
                losses[f's0.{name}'] = value * lw if 'loss' in name else value
        if st0 is not None:                 # stage 0 is issued: now this stream needs the vector
            join_global()
            if global_feat is not None and hasattr(global_feat, 'ready'):
                del global_feat.ready
            inputs_consumed()
        br.join((res, losses))
        roi_labels = res['bbox_targets'][0]
        with torch.no_grad():
            roi_labels = torch.where(roi_labels == self.bbox_head[0].num_classes,
                                     res['cls_score'][:, :-1].argmax(1), roi_labels)
            proposal_list = self.bbox_head[0].refine_bboxes(
                res['rois'], roi_labels, res['bbox_pred'], [r.pos_is_gt for r in samp], img_metas,
                num_per_img=[r.pos_bboxes.size(0) + r.neg_bboxes.size(0) for r in samp])
        samp = sample(1, proposal_list)
        res = self._bbox_forward_train(1, x, samp, gt_bboxes, gt_labels, self.train_cfg[1],
                                       img_metas, global_feat, x_cl)
        lw = self.stage_loss_weights[1]
        for name, value in res['loss_bbox'].items():
            losses[f's1.{name}'] = value * lw if 'loss' in name else value
        return losses

    def forward_train_static(self, x, img_metas, proposals, gt_bboxes, gt_labels, num_gt, keys=None):
        """``forward_train`` (htd_roi_head.py:217-317) with static shapes and no host sync, so that
        the whole step INCLUDING the assign + sample steps (:254-264, :300-310) is one CUDA graph.

        ``proposals`` [B,N,4], ``gt_bboxes`` [B,G,4] / ``gt_labels`` [B,G] padded to G slots,
        ``num_gt`` [B] int32 on the device; ``keys`` = two tensors [B,G+N] and [B,G+num] of
        uniform random numbers (default: drawn here) that take the place of the reference
        sampler's ``randperm``.  Every image contributes exactly ``sampler.num`` rows per stage:
        positives, negatives, then pad rows (``ops.assign_sample``).  The first ``num *
        pos_fraction`` rows of an image are the stage-1 regression rows; where an image has fewer
        positives the remaining slots hold its first negatives, which the loss masks out - the
        sampled set, its order and every loss / gradient equal the reference's.  Pad rows (an image
        that cannot give ``num`` samples - usual in stage 1, where more than 128 refined boxes
        are positive) are inert: no label weight, no box target, no PGraph group."""
        B, N = proposals.shape[:2]
        G = gt_bboxes.shape[1]
        dev = proposals.device
        losses = dict()
        global_feat = None
        if self.with_global:
            def multihot_loss(mc_pred):
                real = torch.arange(G, device=dev)[None, :] < num_gt[:, None]
                nc1 = mc_pred.size(1)
                hot = (gt_labels[:, :, None] == torch.arange(nc1, device=dev)) & real[:, :, None]
                return self.glbctx_head.loss_multihot(mc_pred, hot.any(1))
            x_cl, losses['loss_global'], global_feat, _ = self._global_context(x, multihot_loss)
        else:
            x_cl = self._pyramid(x)
        cand, valid = proposals, None
        self.last_static = []
        for stage in range(self.num_stages):
            st0 = self._stream('stage0', dev) if stage == 0 and self.num_stages > 1 and \
                self._on('overlap_stages', proposals) else None
            with _Branch(st0, (list(x_cl), global_feat)) as br:
                cfg = self.train_cfg[stage]
                a, s = dict(cfg.assigner), dict(cfg.sampler)
                head = self.bbox_head[stage]
                assert head.reg_class_agnostic and not head.reg_decoded_bbox and \
                    a.get('ignore_iof_thr', -1) <= 0 and a.get('gt_max_assign_all', True) and \
                    not isinstance(a['neg_iou_thr'], (tuple, list)), \
                    'forward_train_static covers the configs/htd settings'
                k = keys[stage] if keys is not None else \
                    torch.rand((B, G + cand.shape[1]), device=dev, dtype=torch.float32)
                S = ops.assign_sample(cand, gt_bboxes, gt_labels, num_gt, k, valid=valid,
                                      pos_iou_thr=a['pos_iou_thr'], neg_iou_thr=a['neg_iou_thr'],
                                      min_pos_iou=a.get('min_pos_iou', 0.),
                                      match_low_quality=a.get('match_low_quality', True),
                                      add_gt_as_proposals=s.get('add_gt_as_proposals', True),
                                      num=s['num'], pos_fraction=s['pos_fraction'],
                                      neg_pos_ub=s.get('neg_pos_ub', -1))
                self.last_static.append(S)
                num, npos = S.num, S.num_pos
                rois = S.rois
                row_valid = S.kind != 2
                samp = None
                if stage > 0:
                    r3 = rois.view(B, num, 5)
                    samp = [_StaticRows(r3[b, :npos, 1:], r3[b, npos:, 1:]) for b in range(B)]
                res = self._bbox_forward(stage, x, rois, global_feat, samp, img_metas, x_cl,
                                         row_valid=row_valid)
                pw = cfg.get('pos_weight', -1)
                targets = ops.bbox_targets(rois[:, 1:], S.gt_boxes, S.gt_labels, S.kind,
                                           head.num_classes, pw, head.bbox_coder.means,
                                           head.bbox_coder.stds)
                loss = head.loss(res['cls_score'], res['bbox_pred'], rois, *targets, pad_rows=True)
                lw = self.stage_loss_weights[stage]
                for name, value in loss.items():
                    losses[f's{stage}.{name}'] = value * lw if 'loss' in name else value
                if stage < self.num_stages - 1:
                    with torch.no_grad():                      # refine_bboxes (bbox_head.py:227-303):
                        if len({tuple(m['img_shape'][:2]) for m in img_metas}) == 1:
                            cand = head.regress_by_class(rois[:, 1:], None, res['bbox_pred'],
                                                         img_metas[0]).view(B, num, 4)
                        else:                                  # gt rows are masked, not dropped
                            r3, bp = rois.view(B, num, 5), res['bbox_pred'].view(B, num, -1)
                            cand = torch.stack([head.regress_by_class(r3[b, :, 1:], None, bp[b],
                                                                      img_metas[b]) for b in range(B)])
                        valid = (row_valid & (S.is_gt == 0)).view(B, num)
                        self.last_refined = cand
            br.join((losses, cand, valid, S, getattr(self, 'last_refined', None)))
        return losses

    def simple_test_scores(self, x, proposal_list, img_metas):
        """Body of simple_test up to (excluding) get_bboxes/NMS (htd_roi_head.py:319-366):
        refined rois, stage-averaged cls_score, stage-1 bbox_pred."""
        num_imgs = len(proposal_list)
        rois = bbox2roi(proposal_list)
        x_cl = self._pyramid(x)
        global_feat = self.glbctx_head(x)[1] if self.with_global else None
        r0 = self._bbox_forward(0, x, rois, global_feat, x_cl=x_cl)
        n = tuple(len(p) for p in proposal_list)
        cls0, bp0, rs = r0['cls_score'].split(n, 0), r0['bbox_pred'].split(n, 0), rois.split(n, 0)
        label = [s[:, :-1].argmax(dim=1) for s in cls0]
        rois = torch.cat([self.bbox_head[0].regress_by_class(rs[j], label[j], bp0[j], img_metas[j])
                          for j in range(num_imgs)])
        r1 = self._bbox_forward(1, x, rois, global_feat, x_cl=x_cl)
        return rois, (r0['cls_score'] + r1['cls_score']) / 2.0, r1['bbox_pred']

    def simple_test(self, x, proposal_list, img_metas, rescale=False):
        """htd_roi_head.py:319-386."""
        n = tuple(len(p) for p in proposal_list)
        rois, cls_score, bbox_pred = self.simple_test_scores(x, proposal_list, img_metas)
        rois, cls_score, bbox_pred = rois.split(n, 0), cls_score.split(n, 0), bbox_pred.split(n, 0)
        results = []
        for i in range(len(proposal_list)):
            det_bbox, det_label = self.bbox_head[-1].get_bboxes(
                rois[i], cls_score[i], bbox_pred[i], img_metas[i]['img_shape'],
                img_metas[i].get('scale_factor', 1.0), rescale=rescale, cfg=self.test_cfg)
            results.append(bbox2result(det_bbox, det_label, self.bbox_head[-1].num_classes))
        return results

    def aug_test_merged(self, features, proposal_list, img_metas):
        """htd_roi_head.py:388-433: every augmented view (one image each) through both stages, the
        class boxes mapped back to the original image and averaged over the views, scores
        averaged.  Returns (merged_bboxes [n, 4*classes], merged_scores [n, classes+1])."""
        aug_bboxes, aug_scores = [], []
        for x, meta in zip(features, img_metas):
            m = meta[0]
            proposals = bbox_mapping(proposal_list[0][:, :4], m['img_shape'], m['scale_factor'],
                                     m['flip'], m.get('flip_direction', 'horizontal'))
            rois, cls_score, bbox_pred = self.simple_test_scores(x, [proposals], [m])
            bboxes, scores = self.bbox_head[-1].get_bboxes(rois, cls_score, bbox_pred,
                                                           m['img_shape'], m['scale_factor'],
                                                           rescale=False, cfg=None)
            aug_bboxes.append(bboxes)
            aug_scores.append(scores)
        return merge_aug_bboxes(aug_bboxes, aug_scores, img_metas, self.test_cfg)

    def aug_test(self, features, proposal_list, img_metas, rescale=False):
        """htd_roi_head.py:388-440 (configs/htd define no mask branch).  As in the reference the
        merged boxes are in the original image's scale whatever `rescale` says, and the result is
        the per-class list of ONE image."""
        merged_bboxes, merged_scores = self.aug_test_merged(features, proposal_list, img_metas)
        cfg = as_cfg(self.test_cfg)
        det_bboxes, det_labels = multiclass_nms(merged_bboxes, merged_scores, cfg.score_thr,
                                                cfg.nms, cfg.max_per_img)
        return bbox2result(det_bboxes, det_labels, self.bbox_head[-1].num_classes)
