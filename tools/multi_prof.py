#!/usr/bin/env python
"""Profiling target: the multi-table forward (one transform, ten quantisations) on 8 x 1080p uint8 RGB frames."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ivclab_b200 as ivc  # noqa: E402
import bench_configs as BC  # noqa: E402

dev = torch.device("cuda", 0)
rgb = BC.rgb_frames(torch, dev, range(3000, 3008), 1080, 1920)
coders = [ivc.IntraBlockCoder(q) for q in (0.07, 0.2, 0.4, 0.8, 1.0, 1.5, 2, 3, 4, 4.5)]
zz = ivc.forward_rgb_multi(coders, rgb)
torch.cuda.synchronize()
cx = BC.Ctx(torch, None, dev, 0, 1, 6542.1)
t = cx.timed(lambda: ivc.forward_rgb_multi(coders, rgb), 10, warm=2)
print(f"forward_rgb_multi, 10 tables, 8 frames: {t:.3f} ms = {t / 80 * 1e3:.2f} us per frame and table; "
      f"{(8 * 1080 * 1920 * (3 + 10 * 12)) / t / 1e6:.0f} GB/s of {cx.peak:.0f}")
