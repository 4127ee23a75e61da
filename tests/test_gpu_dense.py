"""GPU tests (through the C-ABI) of the own dense tcgen05 kernels (csrc/dense_gemm.cu, SURVEY §8
row f1): every problem kind against fp32 PyTorch on the same bf16-rounded operands - the three
GEMM layouts (K-major / MN-major operand descriptors), the implicit-GEMM 3x3 convolution on 7x7
RoI maps (forward, data gradient, weight gradient; the halo is TMA out-of-bounds fill), split-K,
the fused epilogues, and the autograd wrappers the heads call."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
BF16 = torch.bfloat16


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-9))


def _rand(g, *shape, scale=1.0):
    return (scale * torch.randn(*shape, generator=g)).to(BF16).cuda()


@pytest.mark.parametrize('M,N,K,splits', [(1024, 1024, 12544, 0), (1000, 85, 1024, 0), (130, 260, 200, 1),
                                          (64, 16, 64, 1), (300, 1024, 1096, 3), (2048, 1024, 1024, 0)])
def test_gemm_nt_bias_relu_second_output(M, N, K, splits):
    from htd_b200 import _lib, dense
    g = torch.Generator().manual_seed(M + N + K)
    ldk = (K + 7) // 8 * 8
    A = torch.zeros(M, ldk, dtype=BF16, device='cuda')
    A[:, :K] = _rand(g, M, K)
    B = torch.zeros(N, ldk, dtype=BF16, device='cuda')
    B[:, :K] = _rand(g, N, K, scale=0.05)
    bias = torch.randn(N, generator=g).cuda()
    rb = torch.randn(3, N, generator=g).cuda()
    rc = torch.randint(0, 3, (M,), generator=g).int().cuda()
    ldd = (N + 7) // 8 * 8
    D = torch.full((M, ldd), -7.0, dtype=BF16, device='cuda')
    D2 = torch.full((M, ldd), -7.0, dtype=BF16, device='cuda')
    dense.gemm(_lib.DENSE_NT, A, B, D, M=M, N=N, K=K, lda=ldk, ldb=ldk, ldd=ldd, bias=bias, relu=True,
               D2=D2, row_bias=rb, row_class=rc, splits=splits)
    acc = A[:, :K].float() @ B[:, :K].float().t() + bias
    assert _rel(D[:, :N], torch.relu(acc)) <= 1e-2
    assert _rel(D2[:, :N], torch.relu(acc + rb[rc.long()])) <= 1e-2
    assert (D[:, N:] == -7.0).all()
    # fp32 output, no epilogue
    Df = torch.empty((M, N), dtype=torch.float32, device='cuda')
    dense.gemm(_lib.DENSE_NT, A, B, Df, M=M, N=N, K=K, lda=ldk, ldb=ldk, ldd=N, splits=splits)
    assert _rel(Df, acc - bias) <= 2e-5 * max(1.0, K / 1024) + 1e-6


@pytest.mark.parametrize('M,N,K,splits', [(1024, 12544, 1024, 0), (1000, 1024, 85, 0), (130, 200, 260, 1),
                                          (256, 576, 1024, 2)])
def test_gemm_nn_with_relu_gate(M, N, K, splits):
    """D = A[M,K] . B[K,N] (B read MN-major) * [gate > 0]: the FC data gradient with the ReLU
    backward of the layer below fused into the epilogue."""
    from htd_b200 import _lib, dense
    g = torch.Generator().manual_seed(M * 3 + N + K)
    ldk = (K + 7) // 8 * 8
    A = torch.zeros(M, ldk, dtype=BF16, device='cuda')
    A[:, :K] = _rand(g, M, K)
    ldn = (N + 7) // 8 * 8
    B = torch.zeros(K, ldn, dtype=BF16, device='cuda')
    B[:, :N] = _rand(g, K, N, scale=0.05)
    gate = _rand(g, M, ldn)
    D = torch.empty((M, ldn), dtype=BF16, device='cuda')
    dense.gemm(_lib.DENSE_NN, A, B, D, M=M, N=N, K=K, lda=ldk, ldb=ldn, ldd=ldn, gate=gate, ldg=ldn,
               splits=splits)
    want = (A[:, :K].float() @ B[:, :N].float()) * (gate[:, :N] > 0)
    assert _rel(D[:, :N], want) <= 1e-2


@pytest.mark.parametrize('M,N,K,splits', [(1024, 12544, 1024, 0), (85, 1024, 1000, 0), (200, 130, 260, 1),
                                          (1024, 1024, 2048, 0), (576, 256, 300, 4)])
def test_gemm_tn(M, N, K, splits):
    """D = A[K,M]^T . B[K,N] (both MN-major): the FC weight gradient dW = dY^T X."""
    from htd_b200 import _lib, dense
    g = torch.Generator().manual_seed(M + N * 5 + K)
    ldm, ldn = (M + 7) // 8 * 8, (N + 7) // 8 * 8
    A = torch.zeros(K, ldm, dtype=BF16, device='cuda')
    A[:, :M] = _rand(g, K, M)
    B = torch.zeros(K, ldn, dtype=BF16, device='cuda')
    B[:, :N] = _rand(g, K, N, scale=0.05)
    D = torch.empty((M, N), dtype=torch.float32, device='cuda')
    dense.gemm(_lib.DENSE_TN, A, B, D, M=M, N=N, K=K, lda=ldm, ldb=ldn, ldd=N, splits=splits)
    want = A[:, :M].float().t() @ B[:, :N].float()
    assert _rel(D, want) <= 2e-5 * max(1.0, K / 1024) + 1e-6
    Db = torch.empty((M, ldn), dtype=BF16, device='cuda')
    dense.gemm(_lib.DENSE_TN, A, B, Db, M=M, N=N, K=K, lda=ldm, ldb=ldn, ldd=ldn, splits=splits)
    assert _rel(Db[:, :N], want) <= 1e-2


@pytest.mark.parametrize('P,Cin,Cout', [(1, 64, 64), (7, 256, 576), (256, 576, 576), (23, 576, 1024),
                                        (256, 256, 576)])
def test_conv3x3_forward_dgrad_wgrad_vs_torch(P, Cin, Cout):
    """The three convolution kinds against F.conv2d and its autograd (fp32, same bf16 operands):
    borders (halo = TMA zero fill), a RoI count that is not a multiple of the 5-RoI tile, channel
    counts that are not multiples of the 128-row tile (576)."""
    from htd_b200 import dense
    g = torch.Generator().manual_seed(P + Cin + Cout)
    x = _rand(g, P, Cin, 7, 7).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    w = _rand(g, Cout, Cin, 3, 3, scale=0.03).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    dy = _rand(g, P, Cout, 7, 7).contiguous(memory_format=torch.channels_last)
    y = dense.conv3x3(x, w)
    assert y.shape == (P, Cout, 7, 7) and y.is_contiguous(memory_format=torch.channels_last)
    gx, gw = torch.autograd.grad(y, [x, w], dy)
    xf, wf = x.detach().float().requires_grad_(True), w.detach().float().requires_grad_(True)
    torch.backends.cudnn.allow_tf32 = False
    yf = F.conv2d(xf, wf, padding=1)
    fx, fw = torch.autograd.grad(yf, [xf, wf], dy.float())
    assert _rel(y, yf) <= 1e-2
    assert _rel(gx, fx) <= 1e-2
    assert _rel(gw, fw) <= 1e-2
    assert gw.shape == w.shape and gx.shape == x.shape


def test_conv_epilogue_relu_and_gate():
    from htd_b200 import _lib, dense
    g = torch.Generator().manual_seed(5)
    P, Cin, Cout = 12, 128, 192
    x = _rand(g, P, 7, 7, Cin)
    w = _rand(g, Cout, 3, 3, Cin, scale=0.05)
    gate = _rand(g, P, 7, 7, Cout)
    y = torch.empty((P, 7, 7, Cout), dtype=BF16, device='cuda')
    dense.gemm(_lib.DENSE_CONV_FPROP, w, x, y, P=P, Cin=Cin, Cout=Cout, ldd=Cout, relu=True, gate=gate,
               ldg=Cout)
    want = F.conv2d(x.permute(0, 3, 1, 2).float(), w.permute(0, 3, 1, 2).float(), padding=1)
    want = (torch.relu(want) * (gate.permute(0, 3, 1, 2) > 0)).permute(0, 2, 3, 1)
    assert _rel(y, want) <= 1e-2


@pytest.mark.parametrize('M,K,N,relu', [(1024, 12544, 1024, True), (1024, 1024, 85, False),
                                        (77, 200, 130, True)])
def test_linear_autograd_vs_torch(M, K, N, relu):
    from htd_b200 import dense
    g = torch.Generator().manual_seed(M + K + N)
    x = _rand(g, M, K).requires_grad_(True)
    w = _rand(g, N, K, scale=1.0 / K ** 0.5).requires_grad_(True)
    b = _rand(g, N).requires_grad_(True)
    dy = _rand(g, M, N)
    y = dense.linear(x, w, b, relu)
    gx, gw, gb = torch.autograd.grad(y, [x, w, b], dy)
    xf, wf, bf = (t.detach().float().requires_grad_(True) for t in (x, w, b))
    yf = F.linear(xf, wf, bf)
    if relu:
        # the product's gate is its own bf16 output; use the same mask so that elements whose
        # pre-activation rounds across zero do not count as arithmetic errors
        yf = yf * (y.detach() > 0)
    fx, fw, fb = torch.autograd.grad(yf, [xf, wf, bf], dy.float())
    assert _rel(y, yf) <= 1e-2
    assert _rel(gx, fx) <= 1e-2 and _rel(gw, fw) <= 1e-2 and _rel(gb, fb) <= 1e-2
    assert y.dtype == BF16 and gw.shape == w.shape and gb.shape == b.shape


def test_gate_colsum():
    from htd_b200 import dense
    g = torch.Generator().manual_seed(1)
    dy, y = _rand(g, 1000, 1030), _rand(g, 1000, 1030)
    dy8 = torch.zeros(1000, 1032, dtype=BF16, device='cuda')[:, :1030]
    dy8.copy_(dy)
    dz, cs = dense.gate_colsum(dy8, y)
    want = dy.float() * (y > 0)
    assert torch.equal(dz, want.to(BF16))
    assert _rel(cs, want.sum(0)) <= 1e-5
    _, cs2 = dense.gate_colsum(dy8, None)
    assert _rel(cs2, dy.float().sum(0)) <= 1e-5


def test_dense_results_are_deterministic():
    from htd_b200 import _lib, dense
    g = torch.Generator().manual_seed(2)
    A, B = _rand(g, 512, 4096), _rand(g, 1024, 4096, scale=0.05)
    outs = []
    for _ in range(3):
        D = torch.empty((512, 1024), dtype=BF16, device='cuda')
        dense.gemm(_lib.DENSE_NT, A, B, D, M=512, N=1024, K=4096, lda=4096, ldb=4096, ldd=1024)
        outs.append(D)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


def test_linear_dual_autograd_vs_torch():
    """H = [relu(x w^T + b); relu(x w^T + b + corr[cls])] and its backward (htd_dual_gate: gate,
    bias gradient and per-class corr gradient in one pass) against plain PyTorch."""
    from htd_b200 import dense
    g = torch.Generator().manual_seed(3)
    M, K, N, R = 1000, 520, 1024, 2
    x = _rand(g, M, K).requires_grad_(True)
    w = _rand(g, N, K, scale=1.0 / K ** 0.5).requires_grad_(True)
    b = _rand(g, N).requires_grad_(True)
    corr = _rand(g, R, N).requires_grad_(True)
    cls = torch.randint(0, R, (M,), generator=g).float().cuda()
    dH = _rand(g, 2 * M, N)
    H = dense.linear_dual(x, w, b, corr, cls)
    gx, gw, gb, gc = torch.autograd.grad(H, [x, w, b, corr], dH)
    xf, wf, bf_, cf = (t.detach().float().requires_grad_(True) for t in (x, w, b, corr))
    v = xf @ wf.t() + bf_
    Hf = torch.cat([v, v + cf[cls.long()]], 0) * (H.detach() > 0)      # the product's own gates
    fx, fw, fb, fc = torch.autograd.grad(Hf, [xf, wf, bf_, cf], dH.float())
    assert _rel(H, Hf) <= 1e-2
    for a, e in ((gx, fx), (gw, fw), (gb, fb), (gc, fc)):
        assert _rel(a, e) <= 1e-2


def test_add3_vs_torch():
    from htd_b200 import dense
    g = torch.Generator().manual_seed(4)
    P, C, B = 37, 256, 3
    a = _rand(g, P, C, 7, 7).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    b = _rand(g, P, C, 7, 7).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    gg = _rand(g, B, C, 1, 1).requires_grad_(True)
    rois = torch.zeros(P, 5, device='cuda')
    rois[:, 0] = torch.randint(0, B, (P,), generator=g).float().cuda()
    dy = _rand(g, P, C, 7, 7).contiguous(memory_format=torch.channels_last)
    out = dense.add3(a, b, gg, rois, 1.0)
    ga, gb, g3 = torch.autograd.grad(out, [a, b, gg], dy)
    want = a.float() + b.float() + gg.float()[rois[:, 0].long()]
    assert _rel(out, want) <= 1e-2
    assert torch.equal(ga, dy) and torch.equal(gb, dy)
    w3 = torch.zeros(B, C, device='cuda').index_add_(0, rois[:, 0].long(), dy.float().sum((2, 3)))
    assert _rel(g3.reshape(B, C), w3) <= 1e-2


@pytest.mark.parametrize('mode', [1, 2])
def test_cta_pair_form_matches(mode):
    """The CTA-pair (cta_group::2) form of the kernel - opt-in, hooks build only - against the
    same torch references: conv fprop / dgrad (mode 1), and the three GEMM kinds too (mode 2)."""
    from htd_b200 import _lib
    with _lib.hooks_library() as L:
        L.htd_debug_set_option(b'dense_pair', mode)
        try:
            for P, Cin, Cout in ((7, 256, 576), (256, 576, 576), (23, 576, 1024), (2, 64, 128)):
                test_conv3x3_forward_dgrad_wgrad_vs_torch(P, Cin, Cout)
            test_conv_epilogue_relu_and_gate()
            if mode == 2:
                test_gemm_nt_bias_relu_second_output(1024, 1024, 12544, 0)
                test_gemm_nt_bias_relu_second_output(1000, 85, 1024, 0)
                test_gemm_nn_with_relu_gate(1024, 12544, 1024, 0)
                test_gemm_nn_with_relu_gate(130, 200, 260, 1)
                test_gemm_tn(1024, 12544, 1024, 0)
                test_gemm_tn(200, 130, 260, 1)
        finally:
            L.htd_debug_set_option(b'dense_pair', 0)
