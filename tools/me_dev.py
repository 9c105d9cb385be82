#!/usr/bin/env python
"""Developer aid: integer-kernel vs exact-kernel agreement over many shapes / search ranges, then ME timings
(+-4 on 32 x 1080p, +-16 on 4K).   python tools/me_dev.py [--time-only]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ivclab_b200 as ivc  # noqa: E402
import bench_configs as BC  # noqa: E402

DEV = torch.device("cuda", 0)


def luma_seq(T, H, W, seed, shift=3):
    return BC.luma_seq(torch, DEV, T, H, W, seed, shift)


def timed(fn, reps, warm=3):
    return BC.Ctx(torch, None, DEV, 0, 1, 6542.1).timed(fn, reps, warm)


def main():
    if "--time-only" not in sys.argv:
        g = torch.Generator(device="cuda").manual_seed(1)
        bad = 0
        for (n, H, W) in [(1, 8, 8), (2, 16, 24), (3, 40, 72), (1, 144, 176), (2, 136, 264), (1, 264, 1032)]:
            for sr in (1, 2, 3, 4, 5, 6, 7, 8, 11, 16, 20):
                for kind in ("noise", "flat", "smooth"):
                    if kind == "noise":
                        r = torch.randint(0, 256, (n, H, W), generator=g, device="cuda").double()
                        c = torch.randint(0, 256, (n, H, W), generator=g, device="cuda").double()
                    elif kind == "flat":
                        r = torch.full((n, H, W), 255.0, device="cuda", dtype=torch.float64)
                        c = torch.zeros((n, H, W), device="cuda", dtype=torch.float64)
                    else:
                        s = luma_seq(n + 1, H, W, 7 + sr, shift=min(sr, 3))
                        r, c = s[:-1].contiguous(), s[1:].contiguous()
                    a = ivc.PFrameBlockCoder(1.0, sr, me_mode="int").estimate(r, c)
                    b = ivc.PFrameBlockCoder(1.0, sr, me_mode="exact").estimate(r, c)
                    if not torch.equal(a, b):
                        bad += 1
                        print("MISMATCH", n, H, W, sr, kind, int((a != b).sum()), "of", a.numel())
        print("parity sweep done, mismatching cases:", bad)
    s = luma_seq(33, 1080, 1920, 5000)
    for mode in ("int", "auto", "exact"):
        pc = ivc.PFrameBlockCoder(1.0, 4, me_mode=mode)
        t = timed(lambda: pc.estimate(s[:-1], s[1:]), 10)
        print(f"1080p x32 sr=4 {mode}: {t:.3f} ms  {32 * 1080 * 1920 / t / 1e3:.0f} Mpixel/s")
    for mode in ("int", "auto"):                                   # search + P-frame forward: two kernels vs the fused one
        pc = ivc.PFrameBlockCoder(1.0, 4, me_mode=mode)
        t2 = timed(lambda: pc.forward(s[1:], s[:-1], pc.estimate(s[:-1], s[1:])), 10)
        t1 = timed(lambda: pc.estimate_forward(s[:-1], s[1:]), 10)
        t1c = timed(lambda: pc.estimate_forward(s[:-1], s[1:], channels=2), 10)
        print(f"1080p x32 sr=4 {mode}: estimate + forward {t2:.3f} ms, fused {t1:.3f} ms, fused 2-channel {t1c:.3f} ms")
    pc = ivc.PFrameBlockCoder(1.0, 4, me_mode="auto")
    t = timed(lambda: pc.estimate(s[:1], s[1:2]), 20)
    print(f"1080p x1 sr=4 auto: {t * 1e3:.1f} us")
    for sr in (8, 16):
        pc = ivc.PFrameBlockCoder(1.0, sr, me_mode="int")
        t = timed(lambda: pc.estimate(s[:-1], s[1:]), 5)
        print(f"1080p x32 sr={sr} int: {t:.3f} ms  {32 * 1080 * 1920 / t / 1e3:.0f} Mpixel/s")
    del s
    s4 = luma_seq(9, 2160, 3840, 4000, shift=12)
    res = {}
    for mma in ("1", "0"):                                  # tensor-core cross term vs the dp4a kernel (A/B switch)
        os.environ["IVC_ME_MMA"] = mma
        pc = ivc.PFrameBlockCoder(1.0, 16, me_mode="int")
        res[mma] = pc.estimate(s4[:-1], s4[1:])
        t = timed(lambda: pc.estimate(s4[:-1], s4[1:]), 5, warm=2)
        print(f"4K x8 sr=16 int IVC_ME_MMA={mma}: {t / 8:.4f} ms/frame  {8 * 2160 * 3840 / t / 1e3:.0f} Mpixel/s")
        s8 = s4.to(torch.uint8)
        t = timed(lambda: pc.estimate(s8[:-1], s8[1:]), 5, warm=2)
        print(f"4K x8 sr=16 uint8 planes IVC_ME_MMA={mma}: {t / 8:.4f} ms/frame")
    print("mma == dp4a vectors:", bool(torch.equal(res["1"], res["0"])))
    os.environ.pop("IVC_ME_MMA")
    if "--exact" in sys.argv:
        pc = ivc.PFrameBlockCoder(1.0, 16, me_mode="exact")
        t = timed(lambda: pc.estimate(s4[:2], s4[1:3]), 2, warm=1)
        print(f"4K x2 sr=16 exact: {t / 2:.3f} ms/frame; == int: {bool(torch.equal(pc.estimate(s4[:2], s4[1:3]), res['1'][:2]))}")


if __name__ == "__main__":
    main()
