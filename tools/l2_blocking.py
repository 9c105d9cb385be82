#!/usr/bin/env python
"""Experiment: the bench step (K1, K2, K3, K1p, K2p over 32 x 1080p frames) run in frame chunks, so that a kernel
finds in L2 what the previous kernel of the same chunk just read or wrote.   python tools/l2_blocking.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ivclab_b200 as ivc  # noqa: E402
from ivclab_b200 import _lib  # noqa: E402
import bench  # noqa: E402

L = _lib.lib
H, W, Fr, SR = 1080, 1920, 32, 4
dev = torch.device("cuda", 0)
ycbcr, luma = bench.make_inputs(torch, dev, Fr, 1234)
ref = torch.roll(luma, 1, dims=0).contiguous()
Hp, Wp = H // 8, W // 8
_, dtab = ivc.PatchQuant(1.0)._table_on(dev)
tcode = _lib.F32 if dtab.dtype == torch.float32 else _lib.F64
zz_i = torch.empty((Fr, Hp, Wp, 3, 64), dtype=torch.int32, device=dev)
rec_i = torch.empty((Fr, H, W, 3), dtype=torch.float64, device=dev)
mv = torch.empty((Fr, Hp, Wp, 1), dtype=torch.int64, device=dev)
zz_p = torch.empty((Fr, Hp, Wp, 3, 64), dtype=torch.int32, device=dev)
rec_p = torch.empty((Fr, H, W), dtype=torch.float64, device=dev)
ws = torch.empty(256, dtype=torch.uint8, device=dev)
sp = torch.cuda.current_stream(dev).cuda_stream
chk = _lib.check


def step(c):
    for f in range(0, Fr, c):
        n = min(c, Fr - f)
        chk(L.ivc_intra_forward(0, sp, ycbcr[f].data_ptr(), _lib.F64, n, H, W, 3, H * W * 3, dtab.data_ptr(), tcode,
                                zz_i[f].data_ptr()), "k1")
        chk(L.ivc_intra_inverse(0, sp, zz_i[f].data_ptr(), n, Hp, Wp, 3, dtab.data_ptr(), tcode, rec_i[f].data_ptr(), _lib.F64), "k2")
    for f in range(0, Fr, c):
        n = min(c, Fr - f)
        chk(L.ivc_me_full_search(0, sp, ref[f].data_ptr(), luma[f].data_ptr(), _lib.F64, n, H, W, H * W, H * W, SR,
                                 _lib.ME_AUTO, mv[f].data_ptr(), ws.data_ptr(), 256), "k3")
        chk(L.ivc_pframe_forward(0, sp, luma[f].data_ptr(), ref[f].data_ptr(), mv[f].data_ptr(), _lib.F64, n, H, W, SR,
                                 dtab.data_ptr(), tcode, None, zz_p[f].data_ptr()), "k1p")
        chk(L.ivc_pframe_inverse(0, sp, zz_p[f].data_ptr(), 3, None, ref[f].data_ptr(), mv[f].data_ptr(), _lib.F64, n, H, W,
                                 SR, dtab.data_ptr(), tcode, rec_p[f].data_ptr()), "k2p")


for c in (32, 16, 8, 4, 2, 1):
    for _ in range(3):
        step(c)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        step(c)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"chunk {c:2d} frames: {ms:.3f} ms per step  {Fr * H * W / ms / 1e3:.0f} Mpixel/s")
