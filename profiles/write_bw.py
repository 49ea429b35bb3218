"""Pure-write / copy bandwidth of the box, for context beside the rollout kernel's roofline (which is 99 % stores)."""
import torch
def t(fn, reps=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps
for mb in (280, 1024, 4096):
    n = mb * 1024 * 1024 // 4
    a = torch.empty(n, device="cuda"); b = torch.empty(n, device="cuda")
    dt = t(lambda: a.fill_(1.0)); print(f"fill_  {mb:5d} MB: {mb * 1.048576 / dt / 1e3:8.1f} GB/s")
    dt = t(lambda: a.zero_());    print(f"zero_  {mb:5d} MB: {mb * 1.048576 / dt / 1e3:8.1f} GB/s")
    dt = t(lambda: b.copy_(a));   print(f"copy_  {mb:5d} MB: {2 * mb * 1.048576 / dt / 1e3:8.1f} GB/s (read+write)")
