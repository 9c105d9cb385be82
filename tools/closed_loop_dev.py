#!/usr/bin/env python
"""Developer aid: where a closed-loop P-frame's time goes on ONE 1080p / QCIF frame (exact search, K1p, K2p alone and
back to back), against the batched per-frame cost."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ivclab_b200 as ivc  # noqa: E402
import bench_configs as BC  # noqa: E402

dev = torch.device("cuda", 0)
cx = BC.Ctx(torch, None, dev, 0, 1, 6542.1)
for name, (T, H, W) in (("1080p", (9, 1080, 1920)), ("qcif", (9, 144, 176))):
    s = BC.luma_seq(torch, dev, T, H, W, 5000) + 0.25                     # non-integer: what a reconstruction looks like
    pc = ivc.PFrameBlockCoder(1.0, 4, me_mode="exact")
    ref, cur = s[:1].contiguous(), s[1:2].contiguous()
    mv = pc.estimate(ref, cur)
    zz = pc.forward(cur, ref, mv)
    t_me = cx.timed(lambda: pc.estimate(ref, cur), 50, warm=5)
    t_f = cx.timed(lambda: pc.forward(cur, ref, mv), 50, warm=5)
    t_i = cx.timed(lambda: pc.inverse(zz, ref=ref, mv=mv), 50, warm=5)
    t_all = cx.timed(lambda: pc.inverse(pc.forward(cur, ref, pc.estimate(ref, cur)), ref=ref, mv=mv), 50, warm=5)
    t_me8 = cx.timed(lambda: pc.estimate(s[:-1], s[1:]), 20, warm=3) / (T - 1)
    print(f"{name}: exact search {t_me * 1e3:.1f} us, forward {t_f * 1e3:.1f} us, inverse {t_i * 1e3:.1f} us, "
          f"back to back {t_all * 1e3:.1f} us; search per frame in a batch of {T - 1}: {t_me8 * 1e3:.1f} us")
