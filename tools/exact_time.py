#!/usr/bin/env python
"""Developer aid: the exact search timed on the three shapes that matter (+-4 on 32 x 1080p, +-16 on one 4K frame, +-4 on
one 1080p frame), second against first generation, vectors compared."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ivclab_b200 as ivc  # noqa: E402
import bench_configs as BC  # noqa: E402

dev = torch.device("cuda", 0)
cx = BC.Ctx(torch, None, dev, 0, 1, 6542.1)
cases = (("+-4, 32 x 1080p", 33, 1080, 1920, 4, 0.25), ("+-4, 32 x 1080p integer", 33, 1080, 1920, 4, 0.0),
         ("+-16, 1 x 2160p", 2, 2160, 3840, 16, 0.25), ("+-4, 1 x 1080p", 2, 1080, 1920, 4, 0.25),
         ("+-4, 32 x 1080p in [0, 1]", 33, 1080, 1920, 4, None))
for name, T, H, W, sr, off in cases:
    s = BC.luma_seq(torch, dev, T, H, W, 5000)
    s = (s + off) if off is not None else (s + 0.25) / 255.0
    pc = ivc.PFrameBlockCoder(1.0, sr, me_mode="exact")
    out = {}
    for gen in ("2", "1"):
        os.environ["IVC_ME_EXACT_V1"] = "1" if gen == "1" else "0"
        out[gen] = pc.estimate(s[:-1], s[1:])
        t = cx.timed(lambda: pc.estimate(s[:-1], s[1:]), 10, warm=3)
        print(f"{name}: generation {gen}: {t:.3f} ms")
    assert torch.equal(out["1"], out["2"]), name
os.environ["IVC_ME_EXACT_V1"] = "0"
