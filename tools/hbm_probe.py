"""Measure what HBM gives for the access mixes of this path on the current GPU (context for the roofline):
pure write (fill), copy (read+write), and a write-heavy 5:1 mix like the window gather."""
import torch

dev = torch.device("cuda")
n = (5 * 1024 ** 3) // 4
a = torch.empty(n, dtype=torch.float32, device=dev)
b = torch.empty(n // 5, dtype=torch.float32, device=dev)
c = torch.empty(n // 5, dtype=torch.float32, device=dev)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


t = timeit(lambda: a.zero_())
print(f"fill  5 GiB: {t:.3f} ms  {a.numel() * 4 / t / 1e6:.0f} GB/s (write only)")
h = a[: n // 2]
g = a[n // 2: 2 * (n // 2)]
t = timeit(lambda: g.copy_(h))
print(f"copy  2.5 GiB -> 2.5 GiB: {t:.3f} ms  {2 * h.numel() * 4 / t / 1e6:.0f} GB/s (read+write)")
t = timeit(lambda: (a.zero_(), c.copy_(b)))
print(f"mix   5 GiB fill + 1 GiB copy: {t:.3f} ms  {(a.numel() + 2 * b.numel()) * 4 / t / 1e6:.0f} GB/s")
