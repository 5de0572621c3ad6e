"""PCIe copy probe (pinned host memory): bandwidth vs transfer size, both directions.  profiles/ aid."""
import time
import torch

d = torch.device("cuda")
big = torch.empty(64 * 2**20 // 4, dtype=torch.float32).pin_memory()
gbig = torch.empty_like(big, device=d)
for kb in (64, 256, 512, 1024, 2048, 4096, 8192, 16384, 65536):
    n = kb * 1024 // 4
    h, g = big[:n], gbig[:n]
    row = [f"{kb:6d} KiB"]
    for name, fn in (("h2d", lambda: g.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(g, non_blocking=True))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 20
        row.append(f"{name} {kb * 1024 / dt / 1e9:6.1f} GB/s {dt * 1e6:7.1f} us")
    print(" | ".join(row))
