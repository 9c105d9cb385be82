#!/usr/bin/env python
"""Golden vectors of round 2, produced by the REAL reference classes in this container.

    python oracle/gen_golden_round2.py       # writes tests/golden/g10_intracodec_cases.npz

The unmodified ``ivclab.image.IntraCodec`` (matplotlib / constriction stubbed as in gen_golden_video.py; neither is
touched by image2symbols / symbols2image) on the corner cases ADVICE round 1 named and on the full-size cfg1 image:

* float32 inputs with ``is_source_rgb=False`` (luma [H,W] and [H,W,3]): scipy's DCT stays float32 and the division
  by the float32 table happens in float32 (dct.py:24-26, patchquant.py:59);
* ``symbols2image(symbols, (H, W, 1))``: the `C == 1` branch returns the three-table decode unconverted
  (intracodec.py:130-136);
* cfg1 (S1: seed 0, 512x768 RGB, qScale 1.0): SHA-256 of the symbol stream and of the reconstruction.
TEST INFRASTRUCTURE."""
from __future__ import annotations

import contextlib
import hashlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.gen_golden_video import _Any, _stub  # noqa: E402  (also puts the reference on sys.path)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    mp = _stub("matplotlib")
    mp.pyplot = _stub("matplotlib.pyplot", axes=_Any(), Axes=_Any())
    _stub("constriction", symbol=_Any())
    from ivclab.image import IntraCodec
    from oracle import ivc_oracle as O

    out = {}
    rng = np.random.default_rng(1010)
    with contextlib.redirect_stdout(io.StringIO()):
        luma32 = (O.smooth_noise_luma(11, 40, 56) + rng.normal(0, 0.37, (40, 56))).astype(np.float32)
        ycc32 = (O.rgb2ycbcr(O.smooth_noise_rgb(12, 48, 64)) + rng.normal(0, 0.21, (48, 64, 3))).astype(np.float32)
        for name, img in (("luma32", luma32), ("ycc32", ycc32)):
            for qi, q in enumerate((0.07, 1.0)):
                c = IntraCodec(quantization_scale=q)
                sym = c.image2symbols(img, is_source_rgb=False)
                out[f"{name}_sym{qi}"] = np.asarray(sym, dtype=np.int32)
            out[name] = img
        # (H, W, 1): encode a luma plane, decode with a three-element shape whose C is 1
        c = IntraCodec(quantization_scale=0.4)
        luma = O.smooth_noise_luma(13, 40, 56)
        sym = c.image2symbols(luma, is_source_rgb=False)
        rec = c.symbols2image(sym, (40, 56, 1))
        out["hw1_luma"], out["hw1_sym"], out["hw1_rec"] = luma, np.asarray(sym, dtype=np.int32), rec
        # cfg1, through the real codec
        rgb1 = O.smooth_noise_rgb(0, 512, 768)
        c = IntraCodec(quantization_scale=1.0)
        sym1 = np.asarray(c.image2symbols(rgb1, is_source_rgb=True), dtype=np.int32)
        rec1 = c.symbols2image(sym1, rgb1.shape)
    out["cfg1_sym_sha"], out["cfg1_rec_sha"] = np.array(sha(sym1)), np.array(sha(rec1))
    out["cfg1_sym_len"] = np.array(sym1.size)
    # the oracle restates the same things
    tab = O.quant_table(1.0)
    zz1 = O.intra_forward(O.rgb2ycbcr(rgb1), tab)
    assert np.array_equal(O.zerorun_encode_fast(zz1), sym1), "oracle symbol stream != real IntraCodec"
    assert np.array_equal(O.ycbcr2rgb(O.intra_inverse(zz1, tab)), rec1), "oracle reconstruction != real IntraCodec"
    print("cfg1 through the real IntraCodec: oracle pinned (symbols, reconstruction); hw1 rec shape", out["hw1_rec"].shape, out["hw1_rec"].dtype)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "g10_intracodec_cases.npz"), **out)


if __name__ == "__main__":
    main()
