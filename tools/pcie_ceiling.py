"""What can the box's host memory take from N GPUs at once?  Plain cudaMemcpyAsync device->pinned host from N GPUs
concurrently (one stream per GPU, one process), aggregate GB/s for N = 1, 2, 4, 8 -- the platform ceiling under the
N-GPU end-to-end path (every GPU storing its part of the frame into one host frame over its own PCIe link).
    python tools/pcie_ceiling.py [MiB per GPU]"""
import sys
import time

import torch

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n_all = torch.cuda.device_count()
dev = [torch.empty(mib << 20, dtype=torch.uint8, device="cuda:%d" % i) for i in range(n_all)]
host = [torch.empty(mib << 20, dtype=torch.uint8).pin_memory() for _ in range(n_all)]
streams = [torch.cuda.Stream(device=i) for i in range(n_all)]
for n in [k for k in (1, 2, 4, 8) if k <= n_all]:
    best = 0.0
    for rep in range(5):
        for i in range(n):
            torch.cuda.synchronize(i)
        t0 = time.perf_counter()
        for i in range(n):
            with torch.cuda.stream(streams[i]):
                host[i].copy_(dev[i], non_blocking=True)
        for i in range(n):
            streams[i].synchronize()
        dt = time.perf_counter() - t0
        best = max(best, n * (mib << 20) / dt * 1e-9)
    print("D2H from %d GPU(s) at once, %d MiB each: %.1f GB/s aggregate (%.1f per GPU)" % (n, mib, best, best / n), flush=True)
