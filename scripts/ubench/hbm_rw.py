"""HBM ceilings for the window-repaint traffic pattern (torch elementwise kernels as the yardstick):
write-only (forward sweep: alpha rows out), in-place read-modify-write (backward sweep: posterior rows), copy."""
import torch
n = 5 * 1000 * 1000 * 1000 // 4
x = torch.empty(n, dtype=torch.float32, device="cuda")
y = torch.empty(n, dtype=torch.float32, device="cuda")
def t(f, reps=5):
    f(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
gb = n * 4 / 1e9
ms = t(lambda: x.fill_(1.0)); print(f"write-only  {gb:.1f} GB in {ms:.3f} ms = {gb/ms:.2f} TB/s")
ms = t(lambda: x.mul_(1.0001)); print(f"in-place rw {2*gb:.1f} GB in {ms:.3f} ms = {2*gb/ms:.2f} TB/s")
ms = t(lambda: y.copy_(x)); print(f"copy        {2*gb:.1f} GB in {ms:.3f} ms = {2*gb/ms:.2f} TB/s")
ms = t(lambda: x.sum()); print(f"read-only   {gb:.1f} GB in {ms:.3f} ms = {gb/ms:.2f} TB/s")
