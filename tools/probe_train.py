"""Throwaway GPU probe: ranked parity errors of run_head / run_train for a case and dtype."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import htd_b200
from htd_b200 import synth
from oracle import cases, restate

name, dt, what = sys.argv[1], sys.argv[2], sys.argv[3]
dtype = dict(f32=torch.float32, bf16=torch.bfloat16)[dt]
if dtype == torch.float32:
    torch.backends.cudnn.enabled = False
c = cases.CASES[name]
oh = restate.HTDRoIHead().double()
synth.fill_params_(oh, c['scheme'], c['seed'])
import htd_b200.bbox_heads as bh
bh.ConvModule.fused_gn = os.environ.get('FUSED_GN', '1') == '1'
head = htd_b200.build_htd_roi_head()
synth.fill_params_(head, c['scheme'], c['seed'])
head = head.cuda().to(dtype)
head.compute_dtype = dtype
if what == 'head':
    want = cases.run_head(oh, name, torch.float64)
    got = cases.run_head(head, name, dtype, 'cuda')
else:
    want = cases.run_train(oh, lambda h, *a: h.forward_train_sampled(*a),
                           lambda h, *a: h.simple_test_scores(*a), name, torch.float64)
    got = cases.run_train(head, lambda h, xs, p, g, s, P: synth.sampled_forward_train(h, xs, p, g, s, P),
                          lambda h, x, p, s: h.simple_test_scores(x, p, [dict(img_shape=t) for t in s]),
                          name, dtype, 'cuda')
def l2(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return float((a - b).norm() / max(b.norm(), 1e-30))
errs = sorted(((cases.rel_err(got[k].float(), want[k]), k, float(want[k].abs().max()), l2(got[k], want[k])) for k in want), reverse=True)
for e, k, m, l in errs[:45]:
    print('%-55s max=%.2e l2=%.2e maxabs=%.2e' % (k, e, l, m))
