#!/usr/bin/env python
"""Pin the closed-loop oracle against the REAL reference codec classes and freeze a golden sequence.

    python oracle/gen_golden_video.py       # writes tests/golden/g8_closed_loop.npz

Imports the unmodified ``ivclab.image.IntraCodec`` and ``ivclab.video.MotionCompensator`` from
/root/reference with ``matplotlib`` / ``constriction`` stubbed (neither is installed; neither is
touched by image2symbols / symbols2image), drives them in the order of the working exercise codec
(exercises/ch4/E4-1.py:212-306) and checks oracle/closed_loop.py bit for bit.  TEST INFRASTRUCTURE."""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("IVCLAB_REFERENCE", "/root/reference")
sys.path.insert(0, REF)


class _Any:
    def __getattr__(self, k):
        return _Any()

    def __call__(self, *a, **k):
        return _Any()


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def main():
    mp = _stub("matplotlib")
    mp.pyplot = _stub("matplotlib.pyplot", axes=_Any(), Axes=_Any())
    _stub("constriction", symbol=_Any())
    from ivclab.image import IntraCodec          # the real, unmodified reference classes
    from ivclab.video import MotionCompensator
    from oracle import ivc_oracle as O
    from oracle import closed_loop as CL

    T, H, W, sr, q = 5, 48, 64, 4, 0.4
    frames = O.moving_sequence(21, T, H, W)
    intra, resid = IntraCodec(quantization_scale=q), IntraCodec(quantization_scale=q)
    mc = MotionCompensator(search_range=sr)
    zzs, mvs, recs = [], [], []
    recon = None
    with contextlib.redirect_stdout(io.StringIO()):
        for t, y in enumerate(frames):
            if t == 0:
                sym = intra.image2symbols(y, is_source_rgb=False)
                recon = intra.symbols2image(sym, y.shape)
                recon = recon[..., 0] if recon.ndim == 3 else recon
                zzs.append(intra.zerorun.decode(sym, [H // 8, W // 8, 3]))
            else:
                mv = mc.compute_motion_vector(recon, y)
                pred = mc.reconstruct_with_motion_vector(recon[..., None], mv)[..., 0]
                sym = resid.image2symbols(y - pred, is_source_rgb=False)
                rr = resid.symbols2image(sym, y.shape)
                rr = rr[..., 0] if rr.ndim == 3 else rr
                recon = pred + rr
                mvs.append(mv)
                zzs.append(resid.zerorun.decode(sym, [H // 8, W // 8, 3]))
            recs.append(recon)
    zz, mv, rec = np.stack(zzs), np.stack(mvs), np.stack(recs)
    got = CL.code_sequence(frames, q, sr, "faithful")
    ok = np.array_equal(got["zz"], zz) and np.array_equal(got["mv"], mv) and np.array_equal(got["recon"], rec)
    print("closed-loop oracle == reference classes (faithful decode):", ok)
    if not ok:
        raise SystemExit("closed-loop oracle is NOT pinned")
    luma = CL.code_sequence(frames, q, sr, "luma")
    psnr = lambda a, b: 10 * np.log10(255.0 ** 2 / np.mean((a - b) ** 2))
    print(f"PSNR faithful {psnr(frames, rec):.2f} dB (the reference's scrambled luma decode), luma-correct {psnr(frames, luma['recon']):.2f} dB")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "g8_closed_loop.npz"), frames=frames, zz=zz, mv=mv, recon=rec,
                        qscale=q, sr=sr, recon_luma=luma["recon"], zz_luma=luma["zz"], mv_luma=luma["mv"])


if __name__ == "__main__":
    main()
