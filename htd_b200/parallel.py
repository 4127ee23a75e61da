"""Data parallelism of the RoI head: one process per GPU, images sharded across ranks, and ONE
exchange step - the gradient all-reduce of the head parameters (the reference wraps the whole
detector in MMDistributedDataParallel, ``mmdet/apis/train.py:72-80``, NCCL backend
``configs/_base_/default_runtime.py:11``).  Forward needs no communication: RoIAlign/BA read
only their image's pyramid and PGraph groups never span images (htd_bbox_head.py:198-202).

``GradAllReducer`` buckets the parameters in reverse registration order (~backward order) and
launches an asynchronous ``all_reduce`` for a bucket once the last gradient of that bucket AND of
every earlier bucket has been accumulated, so NCCL traffic over NVLink overlaps the rest of the
backward pass while every rank issues its collectives in the SAME order (NCCL pairs collectives
by issue order: a rank-dependent order - e.g. one rank with no positive RoI, whose BA attention
parameters finish at a different time or not at all - would pair buckets of different sizes).
``allreduce()`` after ``backward()`` launches what is left in index order, waits, averages and
writes the results back into ``p.grad``.  Gradient accumulation works as with DDP: every backward
pass but the last runs under ``no_sync()`` (hooks off), so each rank issues exactly one collective
per bucket and step; a second un-declared backward is reported instead of silently issuing a
rank-dependent number of collectives.  Works with any torch.distributed backend (gloo in the CPU
tests).
"""
import contextlib

import torch
import torch.distributed as dist


class GradAllReducer:

    def __init__(self, params, world_size=None, bucket_mb=32, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.world = world_size or dist.get_world_size(group)
        self.group = group
        self.buckets = []           # lists of params, same dtype/device per bucket
        cap = bucket_mb * 1024 * 1024
        cur, size = [], 0
        for p in reversed(self.params):
            nbytes = p.numel() * p.element_size()
            if cur and (size + nbytes > cap or p.dtype != cur[0].dtype):
                self.buckets.append(cur)
                cur, size = [], 0
            cur.append(p)
            size += nbytes
        if cur:
            self.buckets.append(cur)
        self._bucket_of = {}
        for bi, b in enumerate(self.buckets):
            for p in b:
                self._bucket_of[id(p)] = bi
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        self._reset()

    def _reset(self):
        self._pending = [len(b) for b in self.buckets]
        self._inflight = {}
        self._next = 0              # buckets [0, _next) have been launched, in index order
        self._stale = False         # a gradient changed after its bucket was launched
        self._sync = True

    @contextlib.contextmanager
    def no_sync(self):
        """Backward passes inside this context only accumulate into ``p.grad`` (no collective is
        launched); the last backward of the step runs outside it, then ``allreduce()``."""
        prev, self._sync = self._sync, False
        try:
            yield
        finally:
            self._sync = prev

    def _launch(self, bi):
        bucket = self.buckets[bi]
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in bucket]
        flat = torch.cat([g.reshape(-1) for g in grads])
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self._inflight[bi] = (flat, work)

    def _on_grad(self, p):
        if not self._sync:
            return
        bi = self._bucket_of[id(p)]
        self._pending[bi] -= 1
        if self._pending[bi] < 0 or bi < self._next:
            # a second backward() outside no_sync() before allreduce(): what was launched for this
            # bucket is out of date (reported in allreduce())
            self._stale = True
            return
        # fixed issue order: bucket i only after buckets 0..i-1
        while self._next < len(self.buckets) and self._pending[self._next] == 0:
            self._launch(self._next)
            self._next += 1

    def allreduce(self):
        """Finish the step: launch the buckets that were not launched from the hooks (gradients
        that arrived late or never - unused parameters) in index order, wait for every bucket,
        average, write back."""
        if self._stale:
            for _, work in self._inflight.values():
                work.wait()
            self._reset()
            raise RuntimeError('GradAllReducer: more than one backward() since the last allreduce(); '
                               'run all but the last one under `with reducer.no_sync():`')
        for bi in range(self._next, len(self.buckets)):
            self._launch(bi)
        self._next = len(self.buckets)
        inv = 1.0 / self.world
        for bi, bucket in enumerate(self.buckets):
            flat, work = self._inflight[bi]
            work.wait()
            flat.mul_(inv)
            off = 0
            for p in bucket:
                n = p.numel()
                view = flat[off:off + n].view_as(p)
                if p.grad is None:
                    p.grad = view.clone()
                else:
                    p.grad.copy_(view)
                off += n
        self._reset()

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []


def shard_images(num_images, rank, world_size):
    """Contiguous image shard of a rank (data-parallel by image)."""
    per = (num_images + world_size - 1) // world_size
    lo = min(rank * per, num_images)
    return range(lo, min(lo + per, num_images))
