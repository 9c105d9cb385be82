#!/usr/bin/env python
"""Chapter-3 style use of ivclab_b200: the intra codec on one image, then a rate-distortion sweep over a batch.

Mirrors what the reference's exercises do with ``ivclab.image.IntraCodec`` (exercises/ch4/ex1.py:421-447: for every
qScale encode, decode, PSNR), with the symbol-level codec and the fused sweep kernels of this package.  The bitrate
column is the first-order entropy of the zero-run symbols (the reference codes them with a Huffman coder from the
third-party ``constriction`` package, which is outside this package).

    python examples/ch3_intra_rd_sweep.py            # synthetic images; needs a CUDA device
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ivclab_b200 as ivc  # noqa: E402

QSCALES = [0.07, 0.2, 0.4, 0.8, 1.0, 1.5, 2, 3, 4, 4.5]          # exercises/ch4/ex1.py:385


def smooth_noise_rgb(seed, H, W):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randint(0, 256, (1, 3, H, W), generator=g, device="cuda").float()
    x = torch.nn.functional.avg_pool2d(x, 5, 1, 2) + torch.randn((1, 3, H, W), generator=g, device="cuda") * 4
    return x.clamp(0, 255)[0].permute(1, 2, 0).contiguous().to(torch.uint8)


def main():
    # 1. one image through the codec-level API, numpy in / numpy out like the reference
    img = smooth_noise_rgb(0, 512, 768).cpu().numpy()
    codec = ivc.IntraCodec(quantization_scale=1.0)
    symbols = codec.image2symbols(img, is_source_rgb=True)
    rec = codec.symbols2image(symbols, img.shape)
    print(f"one 512x768 image: {symbols.size} symbols, PSNR {ivc.calc_psnr(img, rec):.2f} dB")

    # 2. a sweep over a resident batch: two kernels per rate-distortion point (forward, decode + PSNR)
    frames = torch.stack([smooth_noise_rgb(100 + i, 1080, 1920) for i in range(8)])
    npix = frames.shape[1] * frames.shape[2]
    print("qScale   PSNR[dB]   entropy[bit/pixel]")
    for q in QSCALES:
        coder = ivc.IntraBlockCoder(q)
        zz = coder.forward_rgb(frames)                                        # uint8 RGB -> scan indices
        sse = coder.inverse_with_distortion(zz, frames, space="rgb")          # decode + ycbcr2rgb + squared error
        psnr = (20 * torch.log10(255.0 / torch.sqrt(sse / (npix * 3)))).mean().item()
        sym = ivc.ZeroRunCoder().encode(zz)
        lo, hi = ivc.symbol_minmax(sym)
        pmf = ivc.stats_marg(sym, np.arange(lo, hi + 2))
        bits = -(pmf[pmf > 0] * np.log2(pmf[pmf > 0])).sum() * sym.numel()
        print(f"{q:6.2f}   {psnr:8.2f}   {bits / (frames.shape[0] * npix):8.3f}")


if __name__ == "__main__":
    main()
