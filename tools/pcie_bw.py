"""GPU: raw pinned-host -> device copy bandwidth for one 256-image FP32 batch (154 MB), to bound the end-to-end number."""
import time
import torch
x = torch.empty(256, 3, 224, 224, dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device="cuda")
for _ in range(3):
    d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    d.copy_(x, non_blocking=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"H2D {x.numel() * 4 / 1e6:.0f} MB in {ms:.3f} ms = {x.numel() * 4 / ms / 1e6:.1f} GB/s -> {256 / ms:.1f} k img/s ceiling")
u = torch.empty(256, 224, 224, 3, dtype=torch.uint8).pin_memory()
du = torch.empty_like(u, device="cuda")
du.copy_(u, non_blocking=True)
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    du.copy_(u, non_blocking=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"H2D u8 {u.numel() / 1e6:.0f} MB in {ms:.3f} ms = {u.numel() / ms / 1e6:.1f} GB/s")
