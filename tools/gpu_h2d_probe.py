"""Raw pinned host->device bandwidth on this box (context for the e2e number)."""
import torch
x = torch.empty(256, 196, 2048).pin_memory()
d = torch.empty_like(x, device="cuda")
for _ in range(2):
    d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    d.copy_(x, non_blocking=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("H2D %.0f MB pinned: %.2f ms -> %.1f GB/s" % (x.numel() * 4 / 1e6, ms, x.numel() * 4 / ms / 1e6))
