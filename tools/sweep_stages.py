#!/usr/bin/env python
"""Developer aid: where one rate-distortion point of the sweep spends its time (8 x 1080p chunk, resident)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ivclab_b200 as ivc  # noqa: E402
from ivclab_b200 import _lib  # noqa: E402
import bench_configs as BC  # noqa: E402

dev = torch.device("cuda", 0)
cx = BC.Ctx(torch, None, dev, 0, 1, 6542.1)
n = 8
rgb = BC.rgb_frames(torch, dev, range(3000, 3000 + n), 1080, 1920)
for q in (0.2, 1.0, 4.0):
    coder = ivc.IntraBlockCoder(q)
    zz = coder.forward_rgb(rgb)
    hist = torch.zeros((n, 8192), dtype=torch.int32, device=dev)
    outside = torch.zeros((n,), dtype=torch.int32, device=dev)
    sp = torch.cuda.current_stream(dev).cuda_stream

    def h():
        _lib.check(_lib.lib.ivc_zerorun_symbol_histogram(0, sp, zz.data_ptr(), n, zz.numel() // 64 // n, 4000, -4096, 8192,
                                                         hist.data_ptr(), outside.data_ptr()), "hist")
    t_f = cx.timed(lambda: coder.forward_rgb(rgb), 20, warm=3)
    t_h = cx.timed(h, 20, warm=3)
    t_i = cx.timed(lambda: coder.inverse_with_distortion(zz, rgb, space="rgb"), 20, warm=3)
    print(f"qScale {q}: forward_rgb {t_f / n * 1e3:.1f} us, symbol histogram {t_h / n * 1e3:.1f} us, decode + distortion "
          f"{t_i / n * 1e3:.1f} us per frame; sum {(t_f + t_h + t_i) / n * 1e3:.1f} us")
sw = ivc.RateDistortionSweep([0.07, 0.2, 0.4, 0.8, 1.0, 1.5, 2, 3, 4, 4.5])
t = cx.timed(lambda: sw.code(rgb), 10, warm=2)
print(f"sweep.code, 10 scales: {t / n / 10 * 1e3:.1f} us per frame and RD point")
