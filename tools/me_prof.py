#!/usr/bin/env python
"""Profiling target: a few launches of the integer ME kernel (+-4 on 32 x 1080p, +-16 on one 4K pair)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ivclab_b200 as ivc  # noqa: E402
from bench_configs import luma_seq  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "int"
s = luma_seq(33, 1080, 1920, 5000)
pc = ivc.PFrameBlockCoder(1.0, 4, me_mode=mode)
for _ in range(2):
    pc.estimate(s[:-1], s[1:])
del s
s4 = luma_seq(2, 2160, 3840, 4000, shift=12)
pc = ivc.PFrameBlockCoder(1.0, 16, me_mode=mode)
for _ in range(2):
    pc.estimate(s4[:-1], s4[1:])
torch.cuda.synchronize()
