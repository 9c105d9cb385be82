#!/usr/bin/env python
"""Developer aid: how many candidates survive the 8-bit prefilter of the exact search on closed-loop content (reference =
a reconstruction, current = an integer frame), and what the search costs there."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ivclab_b200 as ivc  # noqa: E402
import bench_configs as BC  # noqa: E402

dev = torch.device("cuda", 0)
cx = BC.Ctx(torch, None, dev, 0, 1, 6542.1)
H, W, sr = 1080, 1920, 4
s = BC.luma_seq(torch, dev, 4, H, W, 5000)
cl = ivc.ClosedLoopLumaCoder(1.0, sr, decode="luma", me_mode="exact", use_graph=False)
out = cl.code_sequence(s)
ref, cur = out["recon"][2:3].contiguous(), s[3:4].contiguous()
pc = ivc.PFrameBlockCoder(1.0, sr, me_mode="exact")
for gen in ("2", "1"):
    os.environ["IVC_ME_EXACT_V1"] = "1" if gen == "1" else "0"
    t = cx.timed(lambda: pc.estimate(ref, cur), 20, warm=3)
    print(f"search alone, one 1080p frame, closed-loop content: generation {gen}: {t * 1e3:.1f} us")
os.environ["IVC_ME_EXACT_V1"] = "0"
# the prefilter replayed in torch: bytes under a map, S~ for the 81 candidates, survivors under eps = 16 d sqrt(S~) + 64 d^2
def replay(name, lo, inv_q):
    xr, xc = (ref[0] - lo) * inv_q, (cur[0] - lo) * inv_q
    rb, cb = torch.round(xr), torch.round(xc)
    bad = int(((rb < 0) | (rb > 255)).sum() + ((cb < 0) | (cb > 255)).sum())
    d = (0.5 if (xr != rb).any() else 0.0) + (0.5 if (xc != cb).any() else 0.0)
    pad = torch.nn.functional.pad(rb, (sr, sr, sr, sr), value=float("nan"))
    cbk = cb.reshape(H // 8, 8, W // 8, 8).permute(0, 2, 1, 3)
    S = []
    for dy in range(2 * sr + 1):
        for dx in range(2 * sr + 1):
            w = pad[dy:dy + H, dx:dx + W].reshape(H // 8, 8, W // 8, 8).permute(0, 2, 1, 3)
            S.append(((cbk - w) ** 2).sum((2, 3)))
    S = torch.stack(S, -1)                                                    # [Hp, Wp, 81], NaN = outside the frame
    eps = 16 * d * S.sqrt() + 64 * d * d
    thr = torch.nan_to_num(S + eps, nan=float("inf")).min(-1, keepdim=True).values
    alive = (S - eps) <= thr
    surv = alive.sum(-1).double()
    rounds = sum(torch.ceil(alive[..., b:b + 32].sum(-1) / 4) for b in range(0, 81, 32))
    print(f"{name}: out of range {bad}, delta {d}; survivors per block: mean {surv.mean():.2f}, 90% "
          f"{surv.flatten().kthvalue(int(0.9 * surv.numel())).values:.0f}, max {surv.max():.0f}; evaluation rounds per block: "
          f"mean {rounds.mean():.2f}; best S~ mean {torch.nan_to_num(S, nan=float('inf')).min(-1).values.mean():.0f}")


replay("identity", 0.0, 1.0)
replay("headroom [-64, 320)", -64.0, 255.0 / 384.0)
lo, hi = float(min(ref.min(), cur.min())), float(max(ref.max(), cur.max()))
replay(f"frame range [{lo:.1f}, {hi:.1f}]", lo, 255.0 / (hi - lo))
# tiles of 16 x 128 pixels that leave [0, 255]
t = ref[0, :1072].reshape(67, 16, 15, 128)
print("share of 2 x 16-block tiles with a value outside [-0.5, 255.5):", float(((t < -0.5) | (t >= 255.5)).any(3).any(1).double().mean()))
