"""Closed-loop luma video coding at the symbol level (TEST INFRASTRUCTURE).

Restates what the working ch4 exercise codec does per frame
(/root/reference/exercises/ch4/E4-1.py:212-306; the same steps as
ivclab/video/videocodec.py:41-75) without the entropy coder:

  I-frame : symbols = image2symbols(Y)            (intracodec.py:66-78, C=1 -> 3 tables)
            recon   = symbols2image(symbols, Y.shape)[..., 0]   (intracodec.py:99-138)
  P-frame : mv   = compute_motion_vector(decoder_recon, Y)
            pred = reconstruct_with_motion_vector(decoder_recon[..., None], mv)[..., 0]
            symbols = image2symbols(Y - pred);  recon = pred + symbols2image(...)[..., 0]

``decode='faithful'`` reproduces the reference's luma decode exactly: with a 2-D ``original_shape``
``symbols2image`` zero-run-decodes only the first Hp*Wp of the 3*Hp*Wp encoded blocks and treats
them as the [Hp, Wp, 1] block grid, i.e. decoded block k is
scan-block k of the flat (h w c) list.  ``decode='luma'`` uses channel 0 of every block instead
(what a correct luma codec would do).  Pinned by oracle/gen_golden_video.py against the real
IntraCodec / MotionCompensator classes.
"""
from __future__ import annotations

import numpy as np

from . import ivc_oracle as O


def decode_blocks(zz: np.ndarray, decode: str) -> np.ndarray:
    """[Hp,Wp,3,64] scan indices -> the [Hp,Wp,1,64] blocks the decoder consumes."""
    hp, wp = zz.shape[:2]
    if decode == "faithful":
        return zz.reshape(-1, 64)[:hp * wp].reshape(hp, wp, 1, 64)
    if decode == "luma":
        return zz[:, :, :1]
    raise ValueError(decode)


def code_sequence(frames: np.ndarray, qscale=1.0, sr=4, decode="faithful", me=O.me_full_search):
    """frames [T,H,W] float64 luma -> dict(zz [T,Hp,Wp,3,64], mv [T-1,Hp,Wp,1], recon [T,H,W])."""
    table = O.quant_table(qscale)
    zzs, mvs, recs = [], [], []
    recon = None
    for t, y in enumerate(frames):
        if t == 0:
            zz = O.intra_forward(y[..., None], table)
            recon = O.intra_inverse(decode_blocks(zz, decode), table)[..., 0]
        else:
            mv = me(recon, y, sr)
            pred, zz = O.pframe_forward(y, recon, mv, sr, table)
            recon = pred + O.intra_inverse(decode_blocks(zz, decode), table)[..., 0]
            mvs.append(mv)
        zzs.append(zz)
        recs.append(recon)
    return {"zz": np.stack(zzs), "mv": np.stack(mvs) if mvs else np.zeros((0,) + zzs[0].shape[:2] + (1,), dtype=np.int64),
            "recon": np.stack(recs)}
