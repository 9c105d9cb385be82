#!/usr/bin/env python
"""Developer aid: cProfile of StreamedCoder.run (host-side overhead per chunk)."""
import cProfile
import os
import pstats
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ivclab_b200 as ivc  # noqa: E402

F = 16
rgb = torch.randint(0, 256, (F, 1080, 1920, 3), dtype=torch.uint8).pin_memory()
cur = torch.randint(0, 256, (F, 1080, 1920), dtype=torch.uint8).pin_memory()
ref = torch.randint(0, 256, (F, 1080, 1920), dtype=torch.uint8).pin_memory()
sc = ivc.StreamedCoder(1.0, 4, chunk_frames=int(sys.argv[1]) if len(sys.argv) > 1 else 1)
for _ in range(2):
    sc.run(rgb, cur, ref)
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    sc.run(rgb, cur, ref)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
