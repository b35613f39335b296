#!/usr/bin/env python3
"""Does host->device bandwidth improve with several concurrent copy streams / smaller pieces?"""
import sys, time, os
import torch
dev = torch.device("cuda", 0)
nbytes = 151 * (1 << 20)
h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
def run(k, pieces):
    streams = [torch.cuda.Stream(dev) for _ in range(k)]
    step = nbytes // pieces
    def once():
        for i in range(pieces):
            with torch.cuda.stream(streams[i % k]):
                d[i * step:(i + 1) * step].copy_(h[i * step:(i + 1) * step], non_blocking=True)
    once(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        once()
    torch.cuda.synchronize()
    t = (time.perf_counter() - t0) / 10
    print(f"streams={k} pieces={pieces}: {t*1e3:.3f} ms = {nbytes/t/1e9:.1f} GB/s", flush=True)
for k, p in ((1, 1), (1, 12), (2, 2), (2, 12), (4, 4), (4, 24)):
    run(k, p)
