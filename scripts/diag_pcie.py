"""Raw PCIe probe with pinned buffers (torch only): H2D alone, D2H alone, both directions at once, in chunk sizes of the
host pipeline's slices.  Explains the e2e floor of bench.py on this box."""
import time, torch
MB = 1 << 20
tot = 328 * MB
res = {}
for chunk_mb in (4, 16, 32, 328):
    chunk = chunk_mb * MB
    n = tot // chunk
    h_in = torch.empty(tot, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(tot, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(tot, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(tot, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def run(up, down):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for i in range(n):
            sl = slice(i * chunk, (i + 1) * chunk)
            if up:
                with torch.cuda.stream(s1): d_in[sl].copy_(h_in[sl], non_blocking=True)
            if down:
                with torch.cuda.stream(s2): h_out[sl].copy_(d_out[sl], non_blocking=True)
        torch.cuda.synchronize()
        return time.perf_counter() - t0
    for _ in range(2): run(True, True)
    a = min(run(True, False) for _ in range(3)); b = min(run(False, True) for _ in range(3)); c = min(run(True, True) for _ in range(3))
    print("chunk %3d MB: H2D %.1f GB/s (%.2f ms)  D2H %.1f GB/s (%.2f ms)  both %.1f GB/s total (%.2f ms)" %
          (chunk_mb, tot / a / 1e9, a * 1e3, tot / b / 1e9, b * 1e3, 2 * tot / c / 1e9, c * 1e3))
