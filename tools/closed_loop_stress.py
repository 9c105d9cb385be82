#!/usr/bin/env python
"""Developer aid: randomised comparison of the fused closed-loop step (one kernel per P-frame) with the three-kernel loop."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ivclab_b200 as ivc  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator(device="cpu").manual_seed(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
N = int(sys.argv[2]) if len(sys.argv) > 2 else 60


def ri(lo, hi):
    return int(torch.randint(lo, hi + 1, (1,), generator=g))


bad = 0
for it in range(N):
    H, W, T, sr = 8 * ri(1, 40), 8 * ri(1, 48), ri(2, 5), (4 if it % 2 == 0 else ri(1, 8))
    q = [0.07, 0.4, 1.0, 2.5, 4.5][ri(0, 4)]
    kind = ri(0, 3)
    base = torch.rand((T, H, W), generator=g, dtype=torch.float64) * 255.0
    if kind == 0:
        x = base.round()
    elif kind == 1:
        yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float64), torch.arange(W, dtype=torch.float64), indexing="ij")
        x = (128 + 120 * torch.sin(xx / 6.0 + torch.arange(T, dtype=torch.float64)[:, None, None]) * torch.cos(yy / 9.0) + 4 * torch.randn((T, H, W), generator=g, dtype=torch.float64)).round().clamp(0, 255)
    elif kind == 2:
        x = base[:1].round().expand(T, H, W).clone()                          # static scene
    else:
        x = base                                                              # non-integer frames
    x = x.to(dev)
    outs = {}
    for fused in ("1", "0"):
        os.environ["IVC_CLOSED_LOOP_FUSED"] = fused
        cl = ivc.ClosedLoopLumaCoder(q, sr, decode="luma", me_mode="exact", use_graph=False)
        outs[fused] = cl.code_sequence(x)
    same = all(torch.equal(outs["1"][k], outs["0"][k]) for k in ("zz", "mv", "recon"))
    if not same:
        bad += 1
        print(f"MISMATCH it={it} H={H} W={W} T={T} sr={sr} q={q} kind={kind}")
os.environ.pop("IVC_CLOSED_LOOP_FUSED", None)
print(f"{N} sequences, {bad} mismatches")
sys.exit(1 if bad else 0)
