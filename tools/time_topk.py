import sys, torch, time
sys.path.insert(0, '/root/repo')
from htd_b200 import ops
from htd_b200.dense_heads import RPNHead
torch.manual_seed(0)
for n, k in ((201600, 2000), (50400, 2000), (12600, 2000), (3150, 2000)):
    x = torch.randn(2, n, device='cuda')
    for _ in range(3): ops.topk_sorted(x, k)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): ops.topk_sorted(x, k)
    b.record(); torch.cuda.synchronize()
    t1 = a.elapsed_time(b) / 10
    a.record()
    for _ in range(10): torch.sort(x, dim=1, descending=True)
    b.record(); torch.cuda.synchronize()
    print(n, k, 'own %.3f ms' % t1, 'torch.sort %.3f ms' % (a.elapsed_time(b) / 10))
