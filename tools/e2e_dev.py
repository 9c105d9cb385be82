#!/usr/bin/env python
"""Developer aid: PCIe bandwidth (pinned, each direction and both at once) and StreamedCoder timings for a
few chunk sizes / frame counts.   python tools/e2e_dev.py"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ivclab_b200 as ivc  # noqa: E402
from bench_configs import luma_seq  # noqa: E402


def bw():
    n = 256 << 20
    h1, h2 = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
    d1, d2 = torch.empty(n, dtype=torch.uint8, device="cuda"), torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(up, down):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(4):
            if up:
                with torch.cuda.stream(s1):
                    d1.copy_(h1, non_blocking=True)
            if down:
                with torch.cuda.stream(s2):
                    h2.copy_(d2, non_blocking=True)
        torch.cuda.synchronize()
        return 4 * n / (time.perf_counter() - t0) / 1e9
    for _ in range(2):
        r = (run(True, False), run(False, True), run(True, True))
    print(f"PCIe pinned: H2D {r[0]:.1f} GB/s, D2H {r[1]:.1f} GB/s, both at once {r[2]:.1f} GB/s each")


def main():
    bw()
    F = 32
    g = torch.Generator(device="cuda").manual_seed(0)
    rgb = (torch.nn.functional.avg_pool2d(torch.rand((F, 3, 1080, 1920), generator=g, device="cuda") * 255, 5, 1, 2)
           .permute(0, 2, 3, 1).contiguous().to(torch.uint8).cpu().pin_memory())
    s = luma_seq(F + 1, 1080, 1920, 5000).to(torch.uint8).cpu()
    ref, cur = s[:-1].contiguous().pin_memory(), s[1:].contiguous().pin_memory()
    for frames in (8, 32):
        for chunk in (1, 2, 4, 8):
            sc = ivc.StreamedCoder(1.0, 4, chunk_frames=chunk, slots=int(os.environ.get("SLOTS", "3")), compute_streams=int(os.environ.get("CS", "2")),
                                   ramp=tuple(int(v) for v in os.environ.get("RAMP", "").split(",") if v))
            for _ in range(2):
                out = sc.run(rgb[:frames], cur[:frames], first_ref=ref[0])
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                out = sc.run(rgb[:frames], cur[:frames], first_ref=ref[0])
            ms = (time.perf_counter() - t0) / 5 * 1e3
            print(f"frames {frames:2d} chunk {chunk}: {ms:7.3f} ms  {frames * 1080 * 1920 / ms / 1e3:8.0f} Mpixel/s  "
                  f"h2d {out['h2d_bytes'] / 1e6:.0f} MB d2h {out['d2h_bytes'] / 1e6:.0f} MB")


if __name__ == "__main__":
    main()
