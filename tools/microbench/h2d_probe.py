import torch, time, ctypes
n = 1 << 30
host = torch.empty(n, dtype=torch.uint8).pin_memory()
dev = torch.empty(n + (1 << 20), dtype=torch.uint8, device="cuda")
s = torch.cuda.Stream()
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(s)
    for _ in range(reps): fn()
    b.record(s); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
with torch.cuda.stream(s):
    ms = timeit(lambda: dev[:n].copy_(host, non_blocking=True))
    print("plain 1 GiB: %.2f ms = %.1f GB/s" % (ms, n / ms / 1e6))
    h2 = host.view(1024, 1 << 20)
    d2 = dev[: 1024 * ((1 << 20) + 16)].view(1024, (1 << 20) + 16)[:, : 1 << 20]
    ms = timeit(lambda: d2.copy_(h2, non_blocking=True))
    print("pitched 1024 x 1 MiB: %.2f ms = %.1f GB/s" % (ms, n / ms / 1e6))
    half = n // 2
    ms = timeit(lambda: dev[:half].copy_(host[:half], non_blocking=True))
    print("plain 0.5 GiB: %.2f ms = %.1f GB/s" % (ms, half / ms / 1e6))
