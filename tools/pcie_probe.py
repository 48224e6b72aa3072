"""PCIe probe: pinned host <-> device copy bandwidth at the e2e leg's message sizes, alone and in both directions at once."""
import torch, time
dev = torch.device("cuda:0")
for mb in (1, 2, 4, 8, 32):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    d2 = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def run(fn, reps=50):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps): fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps
    def d2h():
        with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
    def h2d():
        with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
    def both():
        d2h(); h2d()
    def d2h_sync():
        with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
        s1.synchronize()
    a, b, c, e = run(d2h), run(h2d), run(both), run(d2h_sync)
    print("%2d MiB: D2H %.1f GB/s  H2D %.1f GB/s  both %.1f + %.1f GB/s   D2H+sync per copy %.1f us (%.1f GB/s)" % (
        mb, n / a / 1e9, n / b / 1e9, n / c / 1e9, n / c / 1e9, e * 1e6, n / e / 1e9))
