"""GPU parity tests of the FPN neck (producer side): the product FPN
(htd_b200/necks.py: dense tcgen05 lateral convs in bf16, csrc/fpn.cu top-down / subsample kernels)
against the fp64 CPU oracle (oracle/restate.py FPN) and the fixture written by the reference's own
FPN.  fp32 <= 1e-5, bf16 <= 2e-2 (max|a-b| / max|b|)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _oracle():
    from oracle import cases, restate
    return cases.run_fpn(cases.fpn_fill_(restate.FPN().double()), torch.float64)


def _product(dtype):
    from htd_b200.necks import FPN
    from oracle import cases
    fpn = cases.fpn_fill_(FPN([256, 512, 1024, 2048], 256, 5)).cuda().to(dtype)
    return fpn


@pytest.mark.parametrize('dtype,tol', [(torch.float32, 1e-5), (torch.bfloat16, 2e-2)])
def test_fpn_forward_backward_vs_oracle_and_reference_fixture(dtype, tol):
    from oracle import cases
    saved = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        want = _oracle()
        got = cases.run_fpn(_product(dtype), dtype, 'cuda')
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved
    assert set(got) == set(want)
    errs = {k: cases.rel_err(got[k].float(), want[k]) for k in want}
    bad = {k: v for k, v in errs.items() if not v <= tol}
    assert not bad, bad
    if dtype == torch.float32:
        cases.compare_to_fixture(got, cases.load_fixture(os.path.join(GOLD, 'fpn_f64.npz')), tol)
    print(dtype, max(errs.values()))


def test_topdown_and_subsample_kernels_match_aten_bitwise():
    """csrc/fpn.cu against F.interpolate(mode='nearest') + add and max_pool2d(x, 1, 2), forward and
    backward, on sizes with non-integer scales; fp32 results are bit-identical."""
    import torch.nn.functional as F
    from htd_b200.necks import _Subsample, _TopDown
    g = torch.Generator().manual_seed(0)
    for (hf, wf), (hc, wc) in (((37, 53), (19, 27)), ((10, 14), (5, 7)), ((25, 42), (13, 21)),
                               ((7, 9), (7, 9)), ((50, 31), (7, 5))):
        fine = torch.randn(2, 16, hf, wf, generator=g).cuda().contiguous(memory_format=torch.channels_last)
        coarse = torch.randn(2, 16, hc, wc, generator=g).cuda().contiguous(memory_format=torch.channels_last)
        a, b = fine.clone().requires_grad_(True), coarse.clone().requires_grad_(True)
        c, d = fine.clone().requires_grad_(True), coarse.clone().requires_grad_(True)
        y = _TopDown.apply(a, b)
        yr = c + F.interpolate(d, size=(hf, wf), mode='nearest')
        assert torch.equal(y, yr)
        dy = torch.randn(y.shape, generator=g).cuda()
        y.backward(dy)
        yr.backward(dy)
        assert torch.equal(a.grad, c.grad)
        assert torch.allclose(b.grad, d.grad, rtol=1e-6, atol=1e-6)       # summation order differs
        s, sr = _Subsample.apply(a), F.max_pool2d(c, 1, stride=2)
        assert torch.equal(s, sr)
        ds = torch.randn(s.shape, generator=g).cuda()
        ga, = torch.autograd.grad(s, a, ds)
        gc, = torch.autograd.grad(sr, c, ds)
        assert torch.equal(ga, gc)


def test_fpn_outputs_feed_the_head_without_a_layout_pass():
    """The outputs are channels-last in the compute dtype: the head's pyramid conversion returns
    the SAME storage (no copy), which is the point of emitting them this way."""
    from htd_b200 import ops
    from oracle import cases
    fpn = _product(torch.bfloat16)
    with torch.no_grad():
        outs = fpn(cases.fpn_inputs(torch.float32, 'cuda'))
    assert len(outs) == 5
    for o in outs:
        assert o.dtype == torch.bfloat16 and o.shape[1] == 256
        assert o.is_contiguous(memory_format=torch.channels_last)
        assert ops.to_channels_last(o, torch.bfloat16).data_ptr() == o.data_ptr()


def test_fpn_small_odd_configuration_falls_back_for_narrow_channels():
    """in_channels that are not multiples of 64 (the dense kernel does not apply), one image, three
    outputs of which one is the subsampled extra level, bf16 and fp32, against torch ops."""
    import torch.nn.functional as F
    from htd_b200.necks import FPN
    g = torch.Generator().manual_seed(1)
    for dtype, tol in ((torch.float32, 1e-5), (torch.bfloat16, 2e-2)):
        fpn = FPN([32, 48], 32, 3).cuda()
        xs = [torch.randn(1, 32, 21, 30, generator=g).cuda(), torch.randn(1, 48, 11, 15, generator=g).cuda()]
        with torch.no_grad():
            lat = [F.conv2d(x, c.conv.weight, c.conv.bias) for x, c in zip(xs, fpn.lateral_convs)]
            lat[0] = lat[0] + F.interpolate(lat[1], size=lat[0].shape[2:], mode='nearest')
            want = [F.conv2d(l, c.conv.weight, c.conv.bias, padding=1) for l, c in zip(lat, fpn.fpn_convs)]
            want.append(F.max_pool2d(want[-1], 1, stride=2))
            got = fpn.to(dtype)([x.to(dtype) for x in xs])
        assert len(got) == 3
        for a, b in zip(got, want):
            assert a.shape == b.shape
            err = (a.float() - b).abs().max().item() / b.abs().max().item()
            assert err <= tol, (dtype, err)
