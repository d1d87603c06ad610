"""Write-only / read-only / copy bandwidth of this B200 at the size of the conv operand (1.67 GB), CUDA-event timed."""
import torch

n = 1675 * 1000 * 1000 // 4
a = torch.empty(n, device="cuda")
b = torch.empty(n, device="cuda")


def t(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


gb = n * 4 / 1e9
ms = t(lambda: a.zero_())
print(f"write-only (memset) {gb / ms * 1e3:7.0f} GB/s  {ms:.3f} ms")
ms = t(lambda: a.fill_(1.5))
print(f"write-only (fill kernel) {gb / ms * 1e3:7.0f} GB/s  {ms:.3f} ms")
ms = t(lambda: a.sum())
print(f"read-only (sum) {gb / ms * 1e3:7.0f} GB/s  {ms:.3f} ms")
ms = t(lambda: b.copy_(a))
print(f"copy {2 * gb / ms * 1e3:7.0f} GB/s (read+write)  {ms:.3f} ms")
