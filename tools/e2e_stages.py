#!/usr/bin/env python
"""Developer aid: device time of every stage of one StreamedCoder chunk (inputs resident), per frame."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ivclab_b200 as ivc  # noqa: E402
from ivclab_b200.signal.color import rgb2ycbcr  # noqa: E402
from ivclab_b200.utils.metrics import frame_sse  # noqa: E402
import bench_configs as BC  # noqa: E402

DEV = torch.device("cuda", 0)
luma_seq = lambda T, H, W, seed: BC.luma_seq(torch, DEV, T, H, W, seed)
timed = lambda fn, reps: BC.Ctx(torch, None, DEV, 0, 1, 6542.1).timed(fn, reps, 3)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
g = torch.Generator(device="cuda").manual_seed(0)
rgb = (torch.nn.functional.avg_pool2d(torch.rand((n, 3, 1080, 1920), generator=g, device="cuda") * 255, 5, 1, 2)
       .permute(0, 2, 3, 1).contiguous().to(torch.uint8))
s = luma_seq(n + 1, 1080, 1920, 5000).to(torch.uint8)
ref8, cur8 = s[:-1].contiguous(), s[1:].contiguous()
intra, pf, zr = ivc.IntraBlockCoder(1.0), ivc.PFrameBlockCoder(1.0, 4), ivc.ZeroRunCoder()
cur, ref = cur8.double(), ref8.double()
zz = intra.forward_rgb(rgb)
rec = intra.inverse(zz)
ycc = rgb2ycbcr(rgb)
mv = pf.estimate(ref, cur)
zzp = pf.forward(cur, ref, mv, channels=2)
recp = pf.inverse(zzp, ref=ref, mv=mv)
luma8 = torch.empty((n + 1, 1080, 1920), dtype=torch.uint8, device="cuda")
rgb1 = torch.cat([rgb[-1:], rgb]).contiguous()
stages = {                                               # the stages of StreamedCoder._code, in its order
    "luma8_from_rgb8 (n+1 frames)": lambda: ivc.luma8_from_rgb8(rgb1, out=luma8),
    "u8->f64 (n+1 planes)": lambda: luma8.double(),
    "forward_rgb": lambda: intra.forward_rgb(rgb),
    "zr intra (count+scan+write)": lambda: zr.encode(zz),
    "decode + distortion (K2d)": lambda: intra.inverse_with_distortion(zz, rgb, space="ycbcr"),
    "search + pframe fwd, fused, uint8 planes": lambda: pf.estimate_forward(ref8, cur8, channels=2),
    "(ME on uint8 planes alone)": lambda: pf.estimate(ref8, cur8),
    "(pframe fwd alone, 2 channels)": lambda: pf.forward(cur, ref, mv, channels=2),
    "zr inter (count+scan+write)": lambda: zr.encode(zzp),
    "pframe inv": lambda: pf.inverse(zzp, ref=ref, mv=mv),
    "sse inter": lambda: frame_sse(cur, recp),
}
tot = 0.0
for name, fn in stages.items():
    t = timed(fn, 20)
    if not name.startswith("("):
        tot += t
    print(f"{name:32s} {t * 1e3 / n:8.1f} us/frame")
print(f"{'total':32s} {tot * 1e3 / n:8.1f} us/frame  -> {tot / n * 32:.2f} ms per 32 frames;  symbols/frame intra {zr.encode(zz).numel() / n:.0f} inter {zr.encode(zzp).numel() / n:.0f}")
