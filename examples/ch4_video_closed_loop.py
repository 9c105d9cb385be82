#!/usr/bin/env python
"""Chapter-4 style use of ivclab_b200: closed-loop luma coding of a sequence (I-frame, then P-frames predicted
from the decoder's own reconstruction), as the working exercise codec does (exercises/ch4/E4-1.py:212-306), and
the drop-in MotionCompensator on a single frame pair.

    python examples/ch4_video_closed_loop.py         # synthetic QCIF sequence; needs a CUDA device
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ivclab_b200 as ivc  # noqa: E402


def moving_sequence(T, H, W, seed=2):
    g = torch.Generator(device="cuda").manual_seed(seed)
    m = 16
    base = torch.randint(0, 256, (1, 1, H + 2 * m, W + 2 * m), generator=g, device="cuda").double()
    canvas = torch.nn.functional.avg_pool2d(base, 5, 1, 2)[0, 0]
    shifts = torch.randint(-3, 4, (T, 2), generator=g, device="cuda").cpu().tolist()
    return torch.stack([(canvas[m + dy:m + dy + H, m + dx:m + dx + W]).round().clamp(0, 255) for dy, dx in shifts])


def main():
    seq = moving_sequence(21, 144, 176)                                       # QCIF, 21 frames, float64 luma
    # drop-in class, numpy in / numpy out (ivclab/video/motion.py)
    mc = ivc.MotionCompensator(search_range=4)
    mv = mc.compute_motion_vector(seq[0].cpu().numpy(), seq[1].cpu().numpy())
    pred = mc.reconstruct_with_motion_vector(seq[0].cpu().numpy()[..., None], mv)[..., 0]
    print("frame 1 predicted from frame 0: PSNR %.2f dB" % ivc.calc_psnr(seq[1].cpu().numpy(), pred))

    # the whole closed loop on the device
    coder = ivc.ClosedLoopLumaCoder(quantization_scale=1.0, search_range=4, decode="luma", me_mode="auto")
    out = coder.code_sequence(seq)
    mse = ((out["recon"] - seq) ** 2).mean(dim=(1, 2))
    psnr = 10 * torch.log10(255.0 ** 2 / mse)
    sym = ivc.ZeroRunCoder().encode(out["zz"].reshape(-1, 18, 22, 3, 64))
    print(f"21 frames: mean PSNR {psnr.mean().item():.2f} dB, {sym.numel()} zero-run symbols, "
          f"{out['mv'].numel()} motion vectors")


if __name__ == "__main__":
    main()
