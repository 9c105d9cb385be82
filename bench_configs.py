#!/usr/bin/env python
"""Secondary measurements on the five configurations BASELINE.json names (device-resident inputs,
CUDA events, one GPU).  `bench.py` is the contract benchmark; this script fills the per-config table
of DESIGN.md.    python bench_configs.py [--quick] > profiles/<round>_configs.json"""
from __future__ import annotations

import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import ivclab_b200 as ivc  # noqa: E402

QS = [0.07, 0.2, 0.4, 0.8, 1.0, 1.5, 2, 3, 4, 4.5]     # exercises/ch4/ex1.py:385


def timed(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def luma_seq(T, H, W, seed, shift=3):
    g = torch.Generator(device="cuda").manual_seed(seed)
    m = 4 * shift + 8
    base = torch.randint(0, 256, (1, 1, H + 2 * m, W + 2 * m), generator=g, device="cuda").double()
    canvas = ((torch.nn.functional.avg_pool2d(base, 5, stride=1, padding=2) - 127.5) * 3 + 127.5)[0, 0]
    out = torch.empty((T, H, W), dtype=torch.float64, device="cuda")
    sh = torch.randint(-shift, shift + 1, (T, 2), generator=g, device="cuda").cpu().tolist()
    for t, (dy, dx) in enumerate(sh):
        f = canvas[m + dy:m + dy + H, m + dx:m + dx + W] + torch.randn((H, W), generator=g, device="cuda", dtype=torch.float64)
        out[t] = f.round().clamp(0, 255)
    return out


def main():
    quick = "--quick" in sys.argv
    res = {"gpu": torch.cuda.get_device_name(0)}
    g = torch.Generator(device="cuda").manual_seed(0)

    # cfg1: one 512x768 YCbCr image, qScale 1: device time of the fused pair and of the six per-method calls
    img = torch.rand((512, 768, 3), generator=g, device="cuda", dtype=torch.float64) * 255
    c = ivc.IntraBlockCoder(1.0)
    D, Q, Z, P = ivc.DiscreteCosineTransform(), ivc.PatchQuant(1.0), ivc.ZigZag(), ivc.Patcher()
    t_f = timed(lambda: c.inverse(c.forward(img)), 50)
    t_u = timed(lambda: D.inverse_transform(Q.dequantize(Z.unflatten(Z.flatten(Q.quantize(D.transform(P.patch(img))))))), 50)
    h = img.cpu().numpy()
    t0 = time.perf_counter()
    for _ in range(10):
        c.inverse(c.forward(h))
    t_h = (time.perf_counter() - t0) / 10 * 1e3
    # codec level (ivclab_b200.IntraCodec): uint8 RGB numpy image -> symbols (numpy) -> float64 RGB numpy image, i.e.
    # what IntraCodec.image2symbols + symbols2image do (colour transforms, zero-run coding included), host to host
    rgb_np = (torch.rand((512, 768, 3), generator=g, device="cuda") * 255).to(torch.uint8).cpu().numpy()
    ic = ivc.IntraCodec(1.0)
    ic.symbols2image(ic.image2symbols(rgb_np), rgb_np.shape)
    t0 = time.perf_counter()
    for _ in range(10):
        sym_np = ic.image2symbols(rgb_np)
        ic.symbols2image(sym_np, rgb_np.shape)
    t_c = (time.perf_counter() - t0) / 10 * 1e3
    res["cfg1_512x768_rgb"] = {"fused_fwd_inv_ms": t_f, "six_method_calls_ms": t_u, "numpy_in_numpy_out_ms": t_h,
                               "intracodec_image2symbols_symbols2image_numpy_ms": t_c,
                               "mpixel_s_fused": 512 * 768 / t_f / 1e3}

    # neighbours of the path (next rows): zero-run encode of the cfg1 scan indices, colour front end, SSE
    zz1 = c.forward(img)
    zr = ivc.ZeroRunCoder()
    rgb8 = (torch.rand((8, 1080, 1920, 3), generator=g, device="cuda") * 255).to(torch.uint8)
    ycc = ivc.rgb2ycbcr(rgb8)
    zzb = c.forward(ycc)
    res["next_rows"] = {
        "zerorun_encode_cfg1_ms": timed(lambda: zr.encode(zz1), 20),
        "zerorun_encode_8x1080p_ms": timed(lambda: zr.encode(zzb), 10),
        "zerorun_symbols_per_pixel_8x1080p": zr.encode(zzb).numel() / (8 * 1080 * 1920),
        "forward_rgb8_8x1080p_ms": timed(lambda: c.forward_rgb(rgb8), 20),
        "forward_f64_8x1080p_ms": timed(lambda: c.forward(ycc), 20),
        "rgb2ycbcr_8x1080p_ms": timed(lambda: ivc.rgb2ycbcr(rgb8), 20),
        "frame_sse_8x1080p_rgb_f64_ms": timed(lambda: ivc.frame_sse(ycc, ycc), 20),
    }
    del rgb8, ycc, zzb

    # cfg2: QCIF 21 frames closed loop
    seq = luma_seq(21, 144, 176, 2)
    for graph in (False, True):
        cl = ivc.ClosedLoopLumaCoder(1.0, 4, decode="luma", me_mode="exact", use_graph=graph)
        t = timed(lambda: cl.code_sequence(seq), 20)
        res[f"cfg2_qcif_21f_closed_loop_graph={graph}"] = {"ms_per_sequence": t, "us_per_frame": t / 21 * 1e3,
                                                          "mpixel_s": 21 * 144 * 176 / t / 1e3}

    # cfg3: intra RD sweep, 10 qScales over a resident shard of 1080p frames
    F = 16 if quick else 64
    frames = torch.rand((F, 1080, 1920, 3), generator=g, device="cuda", dtype=torch.float64) * 255
    coders = [ivc.IntraBlockCoder(q) for q in QS]

    def sweep():
        for cd in coders:
            rec = cd.inverse(cd.forward(frames))
            ivc.frame_sse(frames, rec)                          # per-frame squared error for PSNR (row N3)
    t = timed(sweep, 3, warm=1)
    res["cfg3_rd_sweep_1080p"] = {"frames": F, "qscales": len(QS), "ms_per_sweep": t,
                                  "mpixel_s_incl_psnr": F * len(QS) * 1080 * 1920 / t / 1e3}
    t = timed(lambda: [cd.inverse(cd.forward(frames)) for cd in coders], 3, warm=1)
    res["cfg3_rd_sweep_1080p"]["mpixel_s_kernels_only"] = F * len(QS) * 1080 * 1920 / t / 1e3
    # the same sweep from uint8 RGB originals with the distortion measured inside the decoder (one RD point = two
    # kernels: forward_rgb, inverse_with_distortion in RGB space as calc_psnr(img, symbols2image(...)) does)
    rgb8s = (torch.rand((F, 1080, 1920, 3), generator=g, device="cuda") * 255).to(torch.uint8)

    def sweep_fused():
        for cd in coders:
            cd.inverse_with_distortion(cd.forward_rgb(rgb8s), rgb8s, space="rgb")

    def sweep_unfused():
        for cd in coders:
            rec = ivc.ycbcr2rgb(cd.inverse(cd.forward_rgb(rgb8s)))
            ivc.frame_sse(rgb8s, rec)
    t_f, t_u = timed(sweep_fused, 3, warm=1), timed(sweep_unfused, 3, warm=1)
    res["cfg3_rd_sweep_1080p"].update({"rgb8_psnr_fused_ms_per_sweep": t_f, "rgb8_psnr_unfused_ms_per_sweep": t_u,
                                       "mpixel_s_rgb8_psnr_fused": F * len(QS) * 1080 * 1920 / t_f / 1e3,
                                       "mpixel_s_rgb8_psnr_unfused": F * len(QS) * 1080 * 1920 / t_u / 1e3})
    del rgb8s
    del frames
    torch.cuda.empty_cache()

    # cfg4: 4K, +-16 full search, open loop on integer-valued frames (integer kernel) and forced exact kernel
    T4 = 3 if quick else 9
    s4 = luma_seq(T4, 2160, 3840, 4000, shift=12)
    for mode in ("auto", "exact"):
        pc = ivc.PFrameBlockCoder(1.0, 16, me_mode=mode)
        t = timed(lambda: pc.estimate(s4[:-1], s4[1:]), 3 if mode == "auto" else 1, warm=1)
        res[f"cfg4_4k_sr16_me_{mode}"] = {"frame_pairs": T4 - 1, "ms_per_frame": t / (T4 - 1),
                                          "mpixel_s": (T4 - 1) * 2160 * 3840 / t / 1e3}
    del s4

    # cfg5: 1080p closed loop, per-frame latency
    T5 = 30 if quick else 300
    s5 = luma_seq(T5, 1080, 1920, 5000)
    for graph in (False, True):
        cl = ivc.ClosedLoopLumaCoder(1.0, 4, decode="luma", me_mode="exact", use_graph=graph)
        t = timed(lambda: cl.code_sequence(s5), 3, warm=1)
        res[f"cfg5_1080p_{T5}f_closed_loop_graph={graph}"] = {"ms_per_sequence": t, "us_per_frame": t / T5 * 1e3,
                                                               "fps": T5 / t * 1e3, "mpixel_s": T5 * 1080 * 1920 / t / 1e3}
    # the same closed loop over 8 independent 1080p sequences in lockstep (sequences / GOPs are the unit of parallelism)
    S8, T8 = 8, (10 if quick else 30)
    s8 = torch.stack([luma_seq(T8, 1080, 1920, 5100 + i) for i in range(S8)])
    cl = ivc.ClosedLoopLumaCoder(1.0, 4, decode="luma", me_mode="exact", use_graph=False)
    t = timed(lambda: cl.code_sequences(s8), 3, warm=1)
    res[f"cfg5_lockstep_{S8}x{T8}f_1080p_closed_loop"] = {"ms_per_run": t, "us_per_frame": t / (S8 * T8) * 1e3,
                                                          "fps": S8 * T8 / t * 1e3, "mpixel_s": S8 * T8 * 1080 * 1920 / t / 1e3}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
