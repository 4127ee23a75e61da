"""Throwaway GPU probe: (1) which kernel faults in the bf16 head, (2) conv-tower backward accuracy."""
import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
import torch.nn.functional as F

def conv_probe():
    from htd_b200.bbox_heads import ConvModule
    torch.manual_seed(0)
    net = nn.Sequential(ConvModule(256, 576, 3, padding=1, norm_cfg=dict(type='GN', num_groups=36), bias=False),
                        ConvModule(576, 576, 3, padding=1, norm_cfg=dict(type='GN', num_groups=36), bias=False),
                        ConvModule(576, 1024, 3, padding=1, norm_cfg=None, bias=False))
    x = torch.randn(24, 256, 7, 7)
    dy = torch.randn(24, 1024, 7, 7)
    def run(net, x, dy, cl):
        x = x.clone().requires_grad_(True)
        xi = x.contiguous(memory_format=torch.channels_last) if cl else x
        y = net(xi)
        g = torch.autograd.grad((y * dy).sum(), [x] + list(net.parameters()))
        return [y] + list(g)
    ref = run(net.double(), x.double(), dy.double(), False)
    net = net.float().cuda()
    def rel(a, b):
        return ((a.double().cpu() - b).abs().max() / b.abs().max()).item()
    for label, setup in [
        ('default', lambda: None),
        ('allow_tf32=False', lambda: setattr(torch.backends.cudnn, 'allow_tf32', False)),
        ('conv.fp32_precision=ieee', lambda: setattr(torch.backends.cudnn.conv, 'fp32_precision', 'ieee')),
        ('deterministic', lambda: setattr(torch.backends.cudnn, 'deterministic', True)),
        ('benchmark', lambda: setattr(torch.backends.cudnn, 'benchmark', True)),
    ]:
        try:
            setup()
        except Exception as e:
            print(label, 'setup failed', e)
        for cl in (True, False):
            out = run(net, x.cuda(), dy.cuda(), cl)
            print(label, 'channels_last' if cl else 'nchw', ['%.1e' % rel(a, b) for a, b in zip(out, ref)], flush=True)
    print('flags', torch.backends.cudnn.allow_tf32, getattr(torch.backends.cudnn.conv, 'fp32_precision', None))

def bf16_probe():
    os.environ['CUDA_LAUNCH_BLOCKING'] = '1'
    import htd_b200
    from htd_b200 import synth, _lib
    from oracle import cases
    orig = _lib.check
    def check(rc, what=''):
        torch.cuda.synchronize()
        print('ok', what, flush=True)
        return orig(rc, what)
    _lib.check = check
    import htd_b200.ops as ops, htd_b200.pgraph as pg
    ops.check = check; pg.check = check
    c = cases.CASES['small']
    head = htd_b200.build_htd_roi_head()
    synth.fill_params_(head, c['scheme'], c['seed'])
    head = head.cuda().to(torch.bfloat16)
    head.compute_dtype = torch.bfloat16
    out = cases.run_head(head, 'small', torch.bfloat16, 'cuda')
    torch.cuda.synchronize()
    print('bf16 head ok', {k: float(v.float().abs().max()) for k, v in list(out.items())[:4]})

def fp32_probe():
    import htd_b200
    from htd_b200 import synth
    from oracle import cases, restate
    name = 'small'
    c = cases.CASES[name]
    oh = restate.HTDRoIHead().double()
    synth.fill_params_(oh, c['scheme'], c['seed'])
    want = cases.run_head(oh, name, torch.float64)
    def run(label):
        head = htd_b200.build_htd_roi_head()
        synth.fill_params_(head, c['scheme'], c['seed'])
        head = head.cuda()
        got = cases.run_head(head, name, torch.float32, 'cuda')
        errs = {k: cases.rel_err(got[k], want[k]) for k in want}
        worst = sorted(errs.items(), key=lambda kv: -kv[1])[:6]
        print(label, [(k, '%.1e' % v) for k, v in worst], flush=True)
    run('default')
    torch.backends.cudnn.allow_tf32 = False
    run('cudnn tf32 off')
    torch.backends.cuda.matmul.allow_tf32 = False
    run('+matmul tf32 off')
    torch.backends.cudnn.enabled = False
    run('cudnn disabled')


if __name__ == '__main__':
    what = sys.argv[1]
    try:
        {'conv': conv_probe, 'bf16': bf16_probe, 'fp32': fp32_probe}[what]()
    except Exception:
        traceback.print_exc()
