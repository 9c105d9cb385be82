#!/usr/bin/env python
"""Pin the oracle's zero-run decoder and symbol statistics against the REAL reference and freeze goldens.

    python oracle/gen_golden_entropy.py     # writes tests/golden/g9_entropy.npz

Imports the unmodified ``ivclab.entropy.ZeroRunCoder`` / ``stats_marg`` and ``ivclab.image.IntraCodec`` from
/root/reference (``matplotlib`` / ``constriction`` stubbed: neither is installed, neither is touched here),
runs them on seeded inputs and checks oracle/ivc_oracle.py bit for bit.  TEST INFRASTRUCTURE."""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gen_golden_video import REF, _Any, _stub  # noqa: E402,F401


def main():
    mp = _stub("matplotlib")
    mp.pyplot = _stub("matplotlib.pyplot", axes=_Any(), Axes=_Any())
    _stub("constriction", symbol=_Any())
    from ivclab.entropy import ZeroRunCoder, stats_marg      # the real, unmodified reference code
    from ivclab.image import IntraCodec
    from oracle import ivc_oracle as O

    rng = np.random.default_rng(9)
    ok = {}
    with contextlib.redirect_stdout(io.StringIO()):
        # symbols of a colour image and of a luma-only image (3*Hp*Wp blocks encoded, Hp*Wp decoded: SURVEY A13)
        img = O.smooth_noise_rgb(9, 48, 64)
        codec = IntraCodec(quantization_scale=0.4)
        sym = np.asarray(codec.image2symbols(img, is_source_rgb=True), dtype=np.int32)
        zr = ZeroRunCoder(end_of_block=4000)
        dec_full = zr.decode(sym, [6, 8, 3])
        dec_trunc = zr.decode(sym, [6, 8, 1])                # stops after Hp*Wp blocks, rest ignored
        ok["decode_full"] = np.array_equal(O.zerorun_decode(sym, (6, 8, 3)), dec_full)
        ok["decode_truncated"] = np.array_equal(O.zerorun_decode(sym, (6, 8, 1)), dec_trunc)
        ok["decode_inverts_encode"] = np.array_equal(zr.encode(dec_full), sym)
        # random sparse blocks incl. empty blocks, full blocks, leading / trailing zeros
        blocks = rng.integers(-40, 41, size=(5, 7, 3, 64)).astype(np.int32)
        blocks[rng.random(blocks.shape) < 0.7] = 0
        blocks[0, 0, 0] = 0
        blocks[0, 0, 1] = np.arange(1, 65)
        blocks[0, 0, 2, :63] = 0
        blocks[0, 1, 0, 1:] = 0
        sym2 = zr.encode(blocks)
        ok["encode_random"] = np.array_equal(O.zerorun_encode(blocks), sym2)
        ok["decode_random"] = np.array_equal(O.zerorun_decode(sym2, (5, 7, 3)), zr.decode(sym2, [5, 7, 3])) and \
            np.array_equal(zr.decode(sym2, [5, 7, 3]), blocks)
        # error cases of the reference
        errs = {}
        for name, stream, shape in (("short", sym2[:-1], (5, 7, 3)), ("too_few", sym2, (5, 7, 4)),
                                    ("overflow", np.concatenate([np.arange(1, 66), [4000]]).astype(np.int32), (1, 1, 1))):
            msgs = []
            for fn in (lambda: zr.decode(stream, list(shape)), lambda: O.zerorun_decode(stream, shape)):
                try:
                    fn()
                    msgs.append("no error")
                except ValueError as e:
                    msgs.append(str(e).split(":")[0])
            errs[name] = msgs
        ok["errors_same_kind"] = all(a == b for a, b in errs.values())
        # statistics as train_huffman_from_image computes them (intracodec.py:160-166)
        lo, hi = int(sym.min()) - 20, int(sym.max()) + 21
        pmf = stats_marg(sym, pixel_range=np.arange(lo, hi))
        ok["bounds"] = O.symbol_bounds(sym) == (lo, hi)
        ok["stats_marg"] = np.array_equal(O.stats_marg(sym, np.arange(lo, hi)), pmf)
        # a range that cuts the data: values below lo / above the last edge are dropped, the last bin is closed
        pmf_cut = stats_marg(sym, pixel_range=np.arange(-3, 9))
        ok["stats_marg_cut"] = np.array_equal(O.stats_marg(sym, np.arange(-3, 9)), pmf_cut)
        # IntraCodec-level calls: colour image -> symbols -> image; luma plane (2-D) both ways; ragged size (edge padding)
        rec_rgb = codec.symbols2image(sym, img.shape)
        luma = O.smooth_noise_luma(11, 48, 64)
        sym_luma = np.asarray(codec.image2symbols(luma, is_source_rgb=False), dtype=np.int32)
        rec_luma = codec.symbols2image(sym_luma, luma.shape)                # (H, W, 3): SURVEY A13
        ragged = O.smooth_noise_rgb(12, 45, 61)
        sym_ragged = np.asarray(codec.image2symbols(ragged, is_source_rgb=True), dtype=np.int32)
        codec.train_huffman_from_image = None                                # (needs constriction; statistics are pinned below)
        ok["intracodec_symbols_rgb"] = np.array_equal(
            O.zerorun_encode(O.intra_forward(O.rgb2ycbcr(img), O.quant_table(0.4))), sym)
        ok["intracodec_image_rgb"] = np.array_equal(
            O.ycbcr2rgb(O.intra_inverse(O.zerorun_decode(sym, (6, 8, 3)), O.quant_table(0.4))), rec_rgb)
        ok["intracodec_symbols_luma"] = np.array_equal(
            O.zerorun_encode(O.intra_forward(luma[..., None], O.quant_table(0.4))), sym_luma)
        ok["intracodec_image_luma"] = np.array_equal(
            O.intra_inverse(O.zerorun_decode(sym_luma, (6, 8, 1)), O.quant_table(0.4)), rec_luma)
        padded = np.pad(O.rgb2ycbcr(ragged), ((0, 3), (0, 3), (0, 0)), mode="edge")
        ok["intracodec_symbols_ragged"] = np.array_equal(
            O.zerorun_encode(O.intra_forward(padded, O.quant_table(0.4))), sym_ragged)
        img8 = O.smooth_noise_luma(10, 40, 56).astype(np.uint8)
        pmf8 = stats_marg(img8, pixel_range=np.arange(256))
        ok["stats_marg_u8"] = np.array_equal(O.stats_marg(img8, np.arange(256)), pmf8)
    for k, v in ok.items():
        print(f"{k}: {v}")
    print("error kinds (reference, oracle):", errs)
    if not all(ok.values()):
        raise SystemExit("entropy oracle is NOT pinned")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "g9_entropy.npz"), sym=sym, dec_full=dec_full,
                        dec_trunc=dec_trunc, blocks=blocks, sym2=sym2, lo=lo, hi=hi, pmf=pmf, pmf_cut=pmf_cut,
                        img8=img8, pmf8=pmf8, img=img, rec_rgb=rec_rgb, luma=luma, sym_luma=sym_luma, rec_luma=rec_luma,
                        ragged=ragged, sym_ragged=sym_ragged)


if __name__ == "__main__":
    main()
