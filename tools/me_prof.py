#!/usr/bin/env python
"""Profiling target: one launch each of the integer ME kernels (+-4 on 32 x 1080p: k_me_int; +-16 on 4 x 4K:
k_me_mma16, or the dp4a kernel with IVC_ME_MMA=0) plus the exact kernel on one 1080p pair.
    ncu --set full --import-source on -k regex:k_me -o gpurun_out/me python tools/me_prof.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ivclab_b200 as ivc  # noqa: E402
import bench_configs as BC  # noqa: E402

dev = torch.device("cuda", 0)
mode = sys.argv[1] if len(sys.argv) > 1 else "int"
s = BC.luma_seq(torch, dev, 33, 1080, 1920, 5000)
ivc.PFrameBlockCoder(1.0, 4, me_mode=mode).estimate(s[:-1], s[1:])
ivc.PFrameBlockCoder(1.0, 4, me_mode="exact").estimate(s[:1], s[1:2])
del s
s4 = BC.luma_seq(torch, dev, 5, 2160, 3840, 4000, shift=12)
ivc.PFrameBlockCoder(1.0, 16, me_mode=mode).estimate(s4[:-1], s4[1:])
torch.cuda.synchronize()
