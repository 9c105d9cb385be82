#!/usr/bin/env python
"""Developer aid: the fused +-4 search + P-frame forward on float64 frames against uint8 planes (+ the conversion pass)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ivclab_b200 as ivc  # noqa: E402
import bench_configs as BC  # noqa: E402

dev = torch.device("cuda", 0)
cx = BC.Ctx(torch, None, dev, 0, 1, 6542.1)
s = BC.luma_seq(torch, dev, 33, 1080, 1920, 5000)
s8 = s.to(torch.uint8)
pc = ivc.PFrameBlockCoder(1.0, 4, me_mode="auto")
a = pc.estimate_forward(s[:-1], s[1:])
b = pc.estimate_forward(s8[:-1], s8[1:])
assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
print(f"fused, float64 frames: {cx.timed(lambda: pc.estimate_forward(s[:-1], s[1:]), 20, warm=3):.3f} ms")
print(f"fused, uint8 planes:   {cx.timed(lambda: pc.estimate_forward(s8[:-1], s8[1:]), 20, warm=3):.3f} ms")
print(f"float64 -> uint8 (torch): {cx.timed(lambda: s.to(torch.uint8), 20, warm=3):.3f} ms for 33 frames")
print(f"search alone, float64: {cx.timed(lambda: pc.estimate(s[:-1], s[1:]), 20, warm=3):.3f} ms; uint8: "
      f"{cx.timed(lambda: pc.estimate(s8[:-1], s8[1:]), 20, warm=3):.3f} ms")
