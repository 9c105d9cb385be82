#!/usr/bin/env python
"""Developer aid: randomised comparison of the two generations of the exact search (shapes, ranges, content classes)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ivclab_b200 as ivc  # noqa: E402



def content(kind, n, H, W, g, ri, dev):
    r = lambda *s: torch.rand(*s, generator=g, dtype=torch.float64)          # noqa: E731
    base = r(n + 1, H, W) * 255.0
    if kind == 0:                                                             # noise, non-integer
        x = base
    elif kind == 1:                                                           # integers
        x = base.round()
    elif kind == 2:                                                           # smooth + small noise, overshooting [0, 255]
        yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float64), torch.arange(W, dtype=torch.float64), indexing="ij")
        x = 140 + 150 * torch.sin(xx / 7.0 + 0.3 * torch.arange(n + 1, dtype=torch.float64)[:, None, None]) * torch.cos(yy / 5.0) + r(n + 1, H, W)
    elif kind == 3:                                                           # low contrast
        x = 100.0 + r(n + 1, H, W) * ri(1, 40)
    elif kind == 4:                                                           # constant / piecewise constant: ties everywhere
        x = torch.full((n + 1, H, W), 17.25, dtype=torch.float64)
        x[:, : H // 2] = 3.0
    elif kind == 5:                                                           # tiny range on a large offset
        x = 1.0e6 + r(n + 1, H, W) * 0.5
    elif kind == 6:                                                           # negative and large
        x = (base - 128.0) * 10.0 ** ri(-6, 6)
    elif kind == 7:                                                           # periodic: exact ties between candidates
        yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
        x = (((xx % 4) * 37 + (yy % 2) * 11) % 256).double().expand(n + 1, H, W).clone() + 0.5
    elif kind == 8:                                                           # a few wild pixels
        x = base.clone()
        for _ in range(ri(1, 4)):
            x[ri(0, n), ri(0, H - 1), ri(0, W - 1)] = [float("nan"), float("inf"), -float("inf"), 1e30, -1e18][ri(0, 4)]
    else:                                                                     # a static scene with one outlier per frame
        x = base[:1].expand(n + 1, H, W).clone()
        x[:, ri(0, H - 1), ri(0, W - 1)] = 4.0e3
    return x.to(dev)


def run(seed=1, N=300, verbose=True):
    dev = torch.device("cuda", 0)
    g = torch.Generator(device="cpu").manual_seed(seed)

    def ri(lo, hi):
        return int(torch.randint(lo, hi + 1, (1,), generator=g))

    bad = []
    prev = os.environ.get("IVC_ME_EXACT_V1")
    try:
        for it in range(N):
            H, W, n, sr, kind = 8 * ri(1, 30), 8 * ri(1, 40), ri(1, 3), ri(1, 16), ri(0, 9)
            if it % 7 == 0:
                sr = 4
            x = content(kind, n, H, W, g, ri, dev)
            pc = ivc.PFrameBlockCoder(1.0, sr, me_mode="exact")
            os.environ["IVC_ME_EXACT_V1"] = "0"
            a = pc.estimate(x[:-1], x[1:])
            os.environ["IVC_ME_EXACT_V1"] = "1"
            b = pc.estimate(x[:-1], x[1:])
            if not torch.equal(a, b):
                bad.append((it, H, W, n, sr, kind, int((a != b).sum())))
                if verbose:
                    print(f"MISMATCH it={it} H={H} W={W} n={n} sr={sr} kind={kind}: {bad[-1][-1]} of {a.numel()} vectors")
    finally:
        if prev is None:
            os.environ.pop("IVC_ME_EXACT_V1", None)
        else:
            os.environ["IVC_ME_EXACT_V1"] = prev
    return bad


if __name__ == "__main__":
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    bad = run(seed, N)
    print(f"{N} cases, {len(bad)} mismatches")
    sys.exit(1 if bad else 0)
