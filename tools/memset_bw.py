import torch
x = torch.empty(2<<30, dtype=torch.uint8, device="cuda")
for _ in range(3): x.zero_()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): x.zero_()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)/10
print("memset 2GiB: %.3f ms  %.1f GB/s" % (ms, (2<<30)/ms/1e6))
y = torch.empty(1<<30, dtype=torch.uint8, device="cuda"); z = torch.empty_like(y)
for _ in range(3): z.copy_(y)
torch.cuda.synchronize(); e0.record()
for _ in range(10): z.copy_(y)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)/10
print("copy 1GiB: %.3f ms  %.1f GB/s (r+w)" % (ms, 2*(1<<30)/ms/1e6))
s = torch.empty(1, device="cuda")
for _ in range(3): y.float().sum() if False else None
