#!/usr/bin/env python
"""Developer aid: pinned H2D / D2H bandwidth (both directions at once, 20 MB pieces like the streamed coder's)
with and without an HBM-saturating kernel loop running on a third stream."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ivclab_b200 as ivc  # noqa: E402

n, pieces = 20 << 20, 16
h_in = torch.empty(n * pieces, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n * pieces, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n * pieces, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n * pieces, dtype=torch.uint8, device="cuda")
s1, s2, s3 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
img = torch.rand((16, 1080, 1920, 3), device="cuda", dtype=torch.float64) * 255
coder = ivc.IntraBlockCoder(1.0)
big_a = torch.empty(1 << 28, dtype=torch.float32, device="cuda")
big_b = torch.empty_like(big_a)


def run(load):
    torch.cuda.synchronize()
    stop = torch.cuda.Event()
    t0 = time.perf_counter()
    if load == "k1k2":
        with torch.cuda.stream(s3):
            for _ in range(12):
                coder.inverse(coder.forward(img))
    elif load == "copy":
        with torch.cuda.stream(s3):
            for _ in range(12):
                big_b.copy_(big_a)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s1):
        e0.record()
        for i in range(pieces):
            d_in[i * n:(i + 1) * n].copy_(h_in[i * n:(i + 1) * n], non_blocking=True)
        e1.record()
    with torch.cuda.stream(s2):
        f0.record()
        for i in range(pieces):
            h_out[i * n:(i + 1) * n].copy_(d_out[i * n:(i + 1) * n], non_blocking=True)
        f1.record()
    torch.cuda.synchronize()
    return n * pieces / e0.elapsed_time(e1) / 1e6, n * pieces / f0.elapsed_time(f1) / 1e6


for load in ("none", "none", "copy", "k1k2"):
    up, down = run(load)
    print(f"load={load:5s}: H2D {up:5.1f} GB/s, D2H {down:5.1f} GB/s (both directions at once, 16 x 20 MB each)")
