import torch, time
x = torch.empty(183359488 // 4, dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device='cuda')
for _ in range(3): d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): d.copy_(x, non_blocking=True)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
print('H2D 183 MB pinned: %.3f ms  %.1f GB/s' % (ms, 183.359488 / ms))
