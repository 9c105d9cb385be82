#!/usr/bin/env python
"""Profiling target: the exact search (+-4, 32 x 1080p, float64 non-integer frames) -- second and first generation."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ivclab_b200 as ivc  # noqa: E402
import bench_configs as BC  # noqa: E402

dev = torch.device("cuda", 0)
s = BC.luma_seq(torch, dev, 33, 1080, 1920, 5000) + 0.25
pc = ivc.PFrameBlockCoder(1.0, 4, me_mode="exact")
a = pc.estimate(s[:-1], s[1:])
os.environ["IVC_ME_EXACT_V1"] = "1"
b = pc.estimate(s[:-1], s[1:])
torch.cuda.synchronize()
assert torch.equal(a, b)
