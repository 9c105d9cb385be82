#!/usr/bin/env python
"""Developer aid: timeline of one StreamedCoder.run (device intervals per stream and host enqueue times)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ivclab_b200 as ivc  # noqa: E402
import bench_configs as BC  # noqa: E402

F, chunk = int(sys.argv[1]), int(sys.argv[2])
g = torch.Generator(device="cuda").manual_seed(0)
rgb = (torch.nn.functional.avg_pool2d(torch.rand((F, 3, 1080, 1920), generator=g, device="cuda") * 255, 5, 1, 2)
       .permute(0, 2, 3, 1).contiguous().to(torch.uint8).cpu().pin_memory())
first = rgb[-1].clone().pin_memory()
sc = ivc.StreamedCoder(1.0, 4, chunk_frames=chunk, ramp=tuple(int(v) for v in os.environ.get('RAMP', '').split(',') if v),
                       slots=int(os.environ.get('SLOTS', '3')))
run = lambda: sc.run(rgb, first_ref=first)                  # the bench's e2e form: RGB only, luma derived on the device
for _ in range(3):
    run()
torch.cuda.synchronize()
sc.trace = []
origin = torch.cuda.Event(enable_timing=True)
origin.record()
t0 = time.perf_counter()
run()
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"run: {(t1 - t0) * 1e3:.3f} ms host wall")
for label, k, a, b, h0, h1 in sorted(sc.trace, key=lambda r: origin.elapsed_time(r[2])):
    print(f"{label:8s} chunk {k:2d}  device {origin.elapsed_time(a):7.3f} -> {origin.elapsed_time(b):7.3f} ms"
          f"   host enqueue {(h0 - t0) * 1e3:7.3f} -> {(h1 - t0) * 1e3:7.3f} ms")
busy = {}
for label, k, a, b, h0, h1 in sc.trace:
    busy[label] = busy.get(label, 0.0) + a.elapsed_time(b)
print("busy ms per stage:", {k: round(v, 3) for k, v in busy.items()})
