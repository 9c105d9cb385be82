#!/usr/bin/env python
"""Developer aid: per-method wall time of the six-method numpy chain of IntraCodec on the cfg1 image."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ivclab_b200 as ivc  # noqa: E402
from oracle import ivc_oracle as O  # noqa: E402

img = O.rgb2ycbcr(O.smooth_noise_rgb(0, 512, 768))
D, Q, Z, P = ivc.DiscreteCosineTransform(), ivc.PatchQuant(1.0), ivc.ZigZag(), ivc.Patcher()
steps = [("patch+transform", lambda x: D.transform(P.patch(x))), ("quantize", Q.quantize), ("flatten", Z.flatten),
         ("unflatten", Z.unflatten), ("dequantize", Q.dequantize), ("inverse_transform", D.inverse_transform)]
best = None
for rep in range(10):
    x, ts = img, []
    for name, fn in steps:
        t0 = time.perf_counter()
        x = fn(x)
        ts.append((name, (time.perf_counter() - t0) * 1e3))
    if best is None or sum(t for _, t in ts) < sum(t for _, t in best):
        best = ts
print(" | ".join(f"{n} {t:.3f} ms" for n, t in best), "| total %.3f ms" % sum(t for _, t in best))
