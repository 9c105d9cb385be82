#!/usr/bin/env python
"""Profiling target: one launch of the fused search + P-frame forward kernel (+-4, 32 x 1080p) and of its two halves.
    ncu --set full --import-source on -k regex:"k_me_int|k_pframe_forward" -o gpurun_out/pf python tools/pf_prof.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ivclab_b200 as ivc  # noqa: E402
import bench_configs as BC  # noqa: E402

dev = torch.device("cuda", 0)
s = BC.luma_seq(torch, dev, 33, 1080, 1920, 5000)
pc = ivc.PFrameBlockCoder(1.0, 4, me_mode="int")
mv, zz = pc.estimate_forward(s[:-1], s[1:])
mv2 = pc.estimate(s[:-1], s[1:])
zz2 = pc.forward(s[1:], s[:-1], mv2)
torch.cuda.synchronize()
assert torch.equal(mv, mv2) and torch.equal(zz, zz2)
