import torch
x = torch.empty(393216 * 1024, dtype=torch.float32, device="cuda")
y = torch.empty_like(x)
for name, fn, nbytes in (("fill", lambda: x.fill_(1.0), x.numel() * 4), ("copy", lambda: y.copy_(x), x.numel() * 8)):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name}: {ms:.3f} ms  {nbytes / ms / 1e6:.0f} GB/s")
