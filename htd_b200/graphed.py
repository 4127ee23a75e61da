"""CUDA-graph capture of the RoI-head training step.

The step issues ~700 kernels of a few microseconds each (the library FC / conv / GroupNorm /
loss glue around the own kernels), so eager execution is bound by host launch latency, not by
the GPU.  Everything on the path is host-sync free and shape-static once the RoI
counts are fixed (static-shape losses, device-scheduled PGraph), so forward + losses + backward
are captured ONCE into a CUDA graph and replayed: one host call per step.

    step = GraphedTrainStep(head, x, proposals, gts, img_shapes, num_pos)
    losses = step(x_new, proposals_new, gts_new)      # copies into static buffers, replays
    head.parameters() .grad / step.x[i].grad          # static gradient tensors, rewritten per replay

``GraphedTrainStep``: the sampled-RoI protocol with fixed counts per image
(``synth.sampled_forward_train``; bench.py); a different (K, P) needs a new capture.
``GraphedStaticTrainStep``: the complete step including assignment and sampling, static shapes.
"""
import torch

from . import synth


class _GraphedStep:
    """Warm-up on a side stream, capture ``_zero`` + ``_run`` once, replay per step."""

    def _init_grads(self, head, dev, flat_grads, early_modules=None):
        """``early_modules`` (with ``flat_grads``): modules whose parameter gradients are complete
        before backward ends (stage-1 head and BA extractor: backward reaches them first).  Their
        gradients are laid out first in the flat buffer (``self.early_grad``; the rest is
        ``self.late_grad``) and ``self.early_event`` - an EXTERNAL CUDA event recorded inside the
        graph once the last of them has accumulated - lets the data-parallel exchange of that part
        start on another stream while the replay is still running the rest of backward."""
        self.head = head
        self.flat_grad = self.early_grad = self.late_grad = self.early_event = None
        if flat_grads:
            params = list(head.parameters())
            early = []
            if early_modules:
                ids = set()
                for m in early_modules:
                    for p in m.parameters():
                        if id(p) not in ids:
                            ids.add(id(p))
                            early.append(p)
                params = early + [p for p in params if id(p) not in ids]
            dtypes = {p.dtype for p in params}
            assert len(dtypes) == 1, 'flat_grads needs a single parameter dtype'
            self.flat_grad = torch.zeros(sum(p.numel() for p in params), dtype=params[0].dtype,
                                         device=dev)
            self._views, off = [], 0
            for p in params:
                # same strides as the parameter (conv weights are channels-last): AccumulateGrad
                # then accumulates in place instead of re-laying the gradient out
                dense = p.is_contiguous() or p.is_contiguous(memory_format=torch.channels_last)
                v = torch.as_strided(self.flat_grad, p.size(), p.stride(), off) if dense \
                    else self.flat_grad[off:off + p.numel()].view_as(p)
                self._views.append((p, v))
                off += p.numel()
            if early:
                n_early = sum(p.numel() for p in early)
                self.early_grad, self.late_grad = self.flat_grad[:n_early], self.flat_grad[n_early:]
                self.early_event = torch.cuda.Event(external=True)
                self._early_seen, self.early_fired = 0, 0
                self._early_marks = []
                self._armed = False
                self._join_stream = torch.cuda.Stream(device=dev)

                def hook(_p, n=len(early)):
                    # The step runs as parallel branches (streams): each early gradient is
                    # accumulated on the stream of its own branch.  Every hook leaves a marker on
                    # that stream; the last one makes a join stream wait for all of them and
                    # records the external event there - "every early gradient is complete"
                    # whatever the branch structure.  (Captured: markers and waits become edges.)
                    if not self._armed:               # another step object is running backward
                        return
                    mark = torch.cuda.Event()
                    mark.record()
                    self._early_marks.append(mark)
                    self._early_seen += 1
                    if self._early_seen == n:
                        self._early_seen = 0
                        self.early_fired += 1
                        for m in self._early_marks:
                            self._join_stream.wait_event(m)
                        self._early_marks = []
                        self.early_event.record(self._join_stream)
                self._early_handles = [p.register_post_accumulate_grad_hook(hook) for p in early]
        self.losses = None
        self.total = None

    def _capture(self, dev, warmup):
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                 # warm-up off the capture stream
            for _ in range(warmup):
                self._zero()
                self._run()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if self.early_event is not None:
            assert self.early_fired == warmup, 'an early module has a parameter without gradient'
        self._zero()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            if self.flat_grad is not None:
                self.flat_grad.zero_()
            self._run()
        torch.cuda.synchronize()

    def _zero(self):
        if self.flat_grad is not None:
            self.flat_grad.zero_()                 # captured: part of every replay
            for p, v in self._views:
                p.grad = v
        else:
            for p in self.head.parameters():
                p.grad = None
        for t in self.x:
            t.grad = None

    def _finish(self, losses):
        total = sum(v for k, v in losses.items() if 'loss' in k)
        total.backward()
        self.losses, self.total = losses, total
        # all loss / accuracy scalars in ONE device vector (dict order): one D2H per step
        self.loss_vec = torch.stack([v.detach().float().reshape(()) for v in losses.values()])


class GraphedTrainStep(_GraphedStep):

    def __init__(self, head, x, proposals, gts, img_shapes, num_pos, warmup=3, flat_grads=False,
                 early_modules=None, flat_inputs=False):
        """``flat_grads``: parameter gradients live in ONE flat buffer (``self.flat_grad``; every
        ``p.grad`` is a view of it) that the captured step zeroes and accumulates into - the
        data-parallel exchange is then an all-reduce of that buffer, no packing copies.
        ``early_modules``: see ``_init_grads``.
        ``flat_inputs``: the static pyramid levels and proposals are views of ONE device buffer
        (``self.input_flat``, fp32) so that a step's inputs arrive with a single host-to-device
        copy straight into the buffers the graph reads - no staging tensor, no device-to-device
        copy.  ``self.inputs_consumed`` (an EXTERNAL event recorded inside the graph once the
        layout conversion and the SFA head have read the inputs, ~0.1 ms into the step) tells a
        copy stream when the next step's upload may overwrite them."""
        self.img_shapes, self.num_pos = img_shapes, num_pos
        dev = x[0].device
        self._init_grads(head, dev, flat_grads, early_modules)
        self.input_flat = self.inputs_consumed = None
        if flat_inputs:
            assert all(t.dtype == torch.float32 for t in list(x) + list(proposals))
            parts = list(x) + list(proposals)
            self.input_flat = torch.empty(sum(t.numel() for t in parts), dtype=torch.float32, device=dev)
            views, off = [], 0
            for t in parts:
                v = self.input_flat[off:off + t.numel()].view(t.shape)
                v.copy_(t.detach())
                views.append(v)
                off += t.numel()
            self.x = [v.requires_grad_(True) for v in views[:len(x)]]
            self.proposals = views[len(x):]
            self.inputs_consumed = torch.cuda.Event(external=True)
        else:
            self.x = [t.detach().clone().requires_grad_(True) for t in x]
            self.proposals = [p.detach().clone() for p in proposals]
        self.gts = [{k: v.detach().clone().to(dev) for k, v in g.items()} for g in gts]
        self._capture(dev, warmup)

    def _run(self):
        self.head.inputs_consumed_event = self.inputs_consumed
        self._armed = True
        try:
            self._finish(synth.sampled_forward_train(self.head, self.x, self.proposals, self.gts,
                                                     self.img_shapes, self.num_pos))
            if self.early_event is not None:          # the join stream is part of the step
                torch.cuda.current_stream().wait_stream(self._join_stream)
        finally:
            self._armed = False
            self.head.inputs_consumed_event = None

    def load(self, x=None, proposals=None, gts=None, non_blocking=True):
        """Copy new inputs (device or pinned-host tensors) into the static buffers."""
        with torch.no_grad():
            if x is not None:
                for d, s in zip(self.x, x):
                    d.copy_(s, non_blocking=non_blocking)
            if proposals is not None:
                for d, s in zip(self.proposals, proposals):
                    d.copy_(s, non_blocking=non_blocking)
            if gts is not None:
                for d, s in zip(self.gts, gts):
                    for k in d:
                        d[k].copy_(s[k], non_blocking=non_blocking)

    def __call__(self, x=None, proposals=None, gts=None):
        self.load(x, proposals, gts)
        self.graph.replay()
        return self.losses


class GraphedStaticTrainStep(_GraphedStep):
    """The REAL training step - ``HTDRoIHead.forward_train_static``: assign + sample on the
    device (csrc/assign_sample.cu), both stages, losses, backward - as one CUDA graph.

        step = GraphedStaticTrainStep(head, x, img_metas, proposals, gt_bboxes, gt_labels, num_gt)
        losses = step(x=..., proposals=..., gt_bboxes=..., gt_labels=..., num_gt=...)

    ``proposals`` [B,N,4], ``gt_bboxes`` [B,G,4], ``gt_labels`` [B,G], ``num_gt`` [B] int32 are
    static buffers (``load`` copies new values in; G is the per-image gt capacity).  ``keys``:
    two static tensors of uniform random numbers, or None to draw them inside the graph
    (torch's CUDA generator is graph-aware: every replay gets fresh numbers)."""

    def __init__(self, head, x, img_metas, proposals, gt_bboxes, gt_labels, num_gt, keys=None,
                 warmup=3, flat_grads=False):
        dev = x[0].device
        self._init_grads(head, dev, flat_grads)
        self.img_metas = img_metas
        self.x = [t.detach().clone().requires_grad_(True) for t in x]
        self.proposals = proposals.detach().clone()
        self.gt_bboxes = gt_bboxes.detach().clone()
        self.gt_labels = gt_labels.detach().clone()
        self.num_gt = num_gt.detach().to(torch.int32).clone()
        self.keys = None if keys is None else [k.detach().clone() for k in keys]
        self._capture(dev, warmup)

    def _run(self):
        self._finish(self.head.forward_train_static(self.x, self.img_metas, self.proposals,
                                                    self.gt_bboxes, self.gt_labels, self.num_gt,
                                                    keys=self.keys))

    def load(self, x=None, proposals=None, gt_bboxes=None, gt_labels=None, num_gt=None, keys=None,
             non_blocking=True):
        with torch.no_grad():
            if x is not None:
                for d, s in zip(self.x, x):
                    d.copy_(s, non_blocking=non_blocking)
            for d, s in ((self.proposals, proposals), (self.gt_bboxes, gt_bboxes),
                         (self.gt_labels, gt_labels), (self.num_gt, num_gt)):
                if s is not None:
                    d.copy_(s, non_blocking=non_blocking)
            if keys is not None:
                assert self.keys is not None, 'captured with in-graph random keys'
                for d, s in zip(self.keys, keys):
                    d.copy_(s, non_blocking=non_blocking)

    def __call__(self, **inputs):
        self.load(**inputs)
        self.graph.replay()
        return self.losses
