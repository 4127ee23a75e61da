"""world_size-2 gloo tests (CPU) of the one exchange step of the path: the bucketed, overlapped
gradient all-reduce (htd_b200/parallel.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from htd_b200.parallel import GradAllReducer, shard_images
    torch.manual_seed(0)
    net = nn.Sequential(nn.Linear(40, 300), nn.ReLU(), nn.Linear(300, 300), nn.ReLU(),
                        nn.Linear(300, 7))
    unused = nn.Linear(5, 5)                      # never gets a gradient
    params = list(net.parameters()) + list(unused.parameters())
    red = GradAllReducer(params, world, bucket_mb=0.2)     # several buckets
    assert len(red.buckets) > 2
    data = torch.randn(8, 40, generator=torch.Generator().manual_seed(1))
    imgs = shard_images(8, rank, world)
    for step in range(2):                          # two steps: state resets correctly
        for p in params:
            p.grad = None
        net(data[imgs.start:imgs.stop]).pow(2).sum().backward()
        red.allreduce()
    torch.save([p.grad for p in params], os.path.join(out_dir, f'g{rank}.pt'))
    dist.destroy_process_group()


def test_bucketed_grad_allreduce_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    g0 = torch.load(tmp_path / 'g0.pt')
    g1 = torch.load(tmp_path / 'g1.pt')
    # reference: average of the per-shard gradients computed in one process
    torch.manual_seed(0)
    net = nn.Sequential(nn.Linear(40, 300), nn.ReLU(), nn.Linear(300, 300), nn.ReLU(),
                        nn.Linear(300, 7))
    data = torch.randn(8, 40, generator=torch.Generator().manual_seed(1))
    want = None
    for lo in (0, 4):
        for p in net.parameters():
            p.grad = None
        net(data[lo:lo + 4]).pow(2).sum().backward()
        gs = [p.grad.clone() for p in net.parameters()]
        want = gs if want is None else [a + b for a, b in zip(want, gs)]
    want = [w / 2 for w in want]
    for a, b, w in zip(g0, g1, want):
        assert torch.equal(a, b)
        assert torch.allclose(a, w, rtol=1e-5, atol=1e-6)
    assert all((g == 0).all() for g in g0[len(want):])      # unused params: zero, not missing


def _worker_uneven(rank, world, port, out_dir):
    """Rank 1 never uses `side` (a rank without positive RoIs never runs the BA attention convs):
    its gradients are missing there, so the hooks fire in a different pattern on the two ranks.
    The collectives must still pair bucket by bucket; then a step with two backward passes before
    the exchange (gradient accumulation)."""
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from htd_b200.parallel import GradAllReducer
    torch.manual_seed(0)
    trunk = nn.Linear(30, 200)
    side = nn.Linear(200, 200)                    # registered between the two: a middle bucket
    top = nn.Linear(200, 3)
    params = list(trunk.parameters()) + list(side.parameters()) + list(top.parameters())
    red = GradAllReducer(params, world, bucket_mb=0.05)
    assert len(red.buckets) >= 3
    data = torch.randn(4, 30, generator=torch.Generator().manual_seed(2 + rank))

    def loss(scale=1.0):
        h = torch.relu(trunk(data))
        if rank == 0:
            h = h + side(h)
        return top(h).pow(2).sum() * scale
    for p in params:
        p.grad = None
    loss().backward()
    red.allreduce()
    g_single = [p.grad.clone() for p in params]
    for p in params:
        p.grad = None
    with red.no_sync():                           # accumulation: two backward passes, one exchange
        loss(0.25).backward()
    loss(0.75).backward()
    red.allreduce()
    torch.save(dict(single=g_single, accum=[p.grad.clone() for p in params]),
               os.path.join(out_dir, f'u{rank}.pt'))
    dist.destroy_process_group()


def test_collectives_pair_when_a_rank_misses_gradients_and_with_accumulation(tmp_path):
    world = 2
    mp.spawn(_worker_uneven, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0, r1 = torch.load(tmp_path / 'u0.pt'), torch.load(tmp_path / 'u1.pt')
    for a, b in zip(r0['single'], r1['single']):
        assert torch.equal(a, b)                                  # both ranks hold the same average
    # reference in one process
    torch.manual_seed(0)
    trunk, side, top = nn.Linear(30, 200), nn.Linear(200, 200), nn.Linear(200, 3)
    params = list(trunk.parameters()) + list(side.parameters()) + list(top.parameters())
    want = [torch.zeros_like(p) for p in params]
    for rank in range(2):
        data = torch.randn(4, 30, generator=torch.Generator().manual_seed(2 + rank))
        for p in params:
            p.grad = None
        h = torch.relu(trunk(data))
        if rank == 0:
            h = h + side(h)
        top(h).pow(2).sum().backward()
        want = [w + (p.grad if p.grad is not None else 0) / 2 for w, p in zip(want, params)]
    for a, w in zip(r0['single'], want):
        assert torch.allclose(a, w, rtol=1e-5, atol=1e-6)
    for a, b, w in zip(r0['accum'], r1['accum'], want):           # 0.25 + 0.75 of the same loss
        assert torch.equal(a, b) and torch.allclose(a, w, rtol=1e-5, atol=1e-6)


def test_shard_images_covers_everything():
    from htd_b200.parallel import shard_images
    for n, w in ((16, 8), (5, 2), (3, 4), (0, 2)):
        seen = [i for r in range(w) for i in shard_images(n, r, w)]
        assert seen == list(range(n))
