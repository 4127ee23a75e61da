"""CPU check of the kernels' separable RoIAlign math (htd_b200/csrc/roi_axis.h compiled for the
host, tests/emu/roi_emu.cpp) against the oracle - no GPU involved, no product path involved."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from htd_b200 import synth
from oracle import restate

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope='module')
def emu():
    subprocess.check_call(['make', '-C', os.path.join(HERE, 'emu'), '-s'])
    return ctypes.CDLL(os.path.join(HERE, 'emu', '_build', 'libemu.so'))


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _rois(n, H, W, stride, seed):
    g = torch.Generator().manual_seed(seed)
    props = synth.make_proposals(2, n, H * stride, W * stride, seed=seed, min_scale=2.0,
                                 max_scale=1.5 * max(H, W) * stride)
    r = torch.cat([torch.cat([p.new_full((p.size(0), 1), i), p], 1) for i, p in enumerate(props)])
    # un-clipped / negative / degenerate / sub-pixel extras
    extra = torch.tensor([
        [0, -30.0, -20.0, 50.0, 40.0], [1, W * stride - 10.0, H * stride - 8.0, W * stride + 90.0,
                                        H * stride + 70.0],
        [0, 12.0, 12.0, 12.0, 12.0], [1, 12.0, 12.0, 12.0, 40.0], [0, 9.0, 7.0, 9.5, 7.25],
        [1, 0.0, 0.0, W * stride, H * stride], [0, 40.0, 30.0, 20.0, 10.0],
        [0, -500.0, -500.0, -300.0, -300.0], [1, 3.0, 5.0, 3.0 + 7 * stride, 5.0 + 14 * stride],
        [0, -1e4, -1e4, 1e4, 1e4], [0, 2.0, 2.0, 2.0 + 28 * stride, 2.0 + 28 * stride]])
    jit = torch.rand(r.shape[0], 4, generator=g)
    r[:, 1:] += (jit - 0.5) * 1e-3
    return torch.cat([r, extra]).float().contiguous()


@pytest.mark.parametrize('H,W,stride', [(40, 56, 4), (20, 28, 8), (5, 7, 32), (13, 9, 16)])
def test_emu_forward_backward_match_oracle(emu, H, W, stride):
    C, P = 8, 7
    g = torch.Generator().manual_seed(H * 100 + W)
    x = torch.randn(2, C, H, W, generator=g, dtype=torch.float64)
    rois = _rois(40, H, W, stride, seed=H)
    K = rois.shape[0]
    xr = x.clone().requires_grad_(True)
    y = restate.RoIAlign(P, 1.0 / stride, 0)(xr, rois.double())          # fp64 oracle, NCHW
    dy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    dx, = torch.autograd.grad((y * dy).sum(), xr)
    feat = x.permute(0, 2, 3, 1).float().contiguous()
    out = torch.empty(K, P * P, C)
    emu.emu_roi_align_fwd(_p(feat), _p(rois), _p(out), 2, C, H, W, K, P,
                          ctypes.c_double(1.0 / stride), 0)
    want = y.permute(0, 2, 3, 1).reshape(K, P * P, C)
    err = (out.double() - want).abs().max().item() / want.abs().max().item()
    assert err <= 1e-5, err                                               # fp32 gate
    assert err <= 2e-6, err                                               # what we actually get
    dyh = dy.permute(0, 2, 3, 1).reshape(K, P * P, C).float().contiguous()
    dxe = torch.empty(2, H, W, C)
    emu.emu_roi_align_bwd(_p(dyh), _p(rois), _p(dxe), 2, C, H, W, K, P,
                          ctypes.c_double(1.0 / stride), 0)
    wantdx = dx.permute(0, 2, 3, 1)
    err = (dxe.double() - wantdx).abs().max().item() / wantdx.abs().max().item()
    assert err <= 2e-6, err
    assert emu.emu_check_ranges(_p(rois), K, P, H, W, ctypes.c_double(1.0 / stride), 0) == 0


def test_emu_sampling_ratio_2(emu):
    C, P, H, W, stride = 4, 7, 24, 30, 4
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, C, H, W, generator=g, dtype=torch.float64)
    rois = _rois(30, H, W, stride, seed=2)
    rois = rois[(rois[:, 3] >= rois[:, 1]) & (rois[:, 4] >= rois[:, 2])].contiguous()
    K = rois.shape[0]
    y = restate.RoIAlign(P, 1.0 / stride, 2)(x, rois.double())
    feat = x.permute(0, 2, 3, 1).float().contiguous()
    out = torch.empty(K, P * P, C)
    emu.emu_roi_align_fwd(_p(feat), _p(rois), _p(out), 2, C, H, W, K, P,
                          ctypes.c_double(1.0 / stride), 2)
    want = y.permute(0, 2, 3, 1).reshape(K, P * P, C)
    assert (out.double() - want).abs().max().item() / want.abs().max().item() <= 2e-6


def test_emu_levels_match_golden(emu):
    z = np.load(os.path.join(HERE, 'golden', 'levels.npz'))
    rois = torch.from_numpy(z['rois']).contiguous()
    out = torch.empty(rois.shape[0], dtype=torch.int32)
    emu.emu_levels(_p(rois), rois.shape[0], 4, ctypes.c_float(56.0), _p(out))
    assert np.array_equal(out.numpy().astype(np.int8), z['levels'])


def test_forward_ring_allocator_never_overlaps_or_stalls():
    """Host model of the byte-ring carve-out of roi_align_fwd_persist_kernel's producer warp
    (csrc/roi_align.cu): strips of arbitrary sizes (empty units, strips larger than half the ring)
    never overlap a strip that is still in flight, never exceed the ring, and the producer never
    waits for space that cannot appear (units are released in order)."""
    import random
    CAP, Q = 200 * 1024, 8
    rnd = random.Random(0)
    for _ in range(1500):
        needs = [rnd.choice([0, 16 * rnd.randint(1, 800), 16 * rnd.randint(4000, 12800),
                             16 * rnd.randint(1, 3000)]) for _ in range(rnd.randint(1, 80))]
        head, free, oldest, used, live = 0, CAP, 0, [0] * Q, []
        for i, need in enumerate(needs):
            spins = 0
            while True:
                wrap = head + need > CAP
                waste = CAP - head if wrap else 0
                if oldest + Q > i and free >= need + waste:
                    break
                if oldest == i:
                    head = 0
                    spins += 1
                    assert spins < 3
                    continue
                free += used[oldest % Q]
                live.pop(0)
                oldest += 1
            if wrap:
                head = 0
            buf = head
            head += need
            free -= need + waste
            used[i % Q] = need + waste
            assert buf + need <= CAP and free >= 0
            assert all(not (buf < b and a < buf + need) for a, b in live)
            live.append((buf, buf + need))
