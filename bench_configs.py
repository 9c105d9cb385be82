#!/usr/bin/env python
"""The five configurations BASELINE.json names, measured as named.

`bench.py` imports :func:`run_all` and puts its result under the key ``configs`` of the JSON line, so the
driver's records hold every configuration.  Sharding follows section 8e:

* cfg1 / cfg2 (one image, one QCIF sequence): rank 0 only, the other ranks idle;
* cfg3: a FIXED pool of frames (1024 by default) split into contiguous ranges over the ranks (strong scaling),
  all 10 qScales on the resident shard, host-fed: every frame uploaded once, PSNR + symbol statistics
  gathered on rank 0 inside the timed region (`RateDistortionSweep`, `shard.gather_rows`);
* cfg4: 8 sequences of 120 4K frames, +-16 full search, sequence s on rank s mod N (strong scaling);
* cfg5: closed loop, sequential dependency: replicas only -- every rank runs the same sequence, the slowest counts.

Inputs are synthetic and generated on the device from per-frame / per-sequence seeds (generation is not timed).
Times are CUDA events on the launching stream for device-resident entries and barrier-bracketed wall clock for
the host-fed / gathered ones; the maximum over ranks is reported.  Standalone:  python bench_configs.py [--quick]"""
from __future__ import annotations

import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

QS = [0.07, 0.2, 0.4, 0.8, 1.0, 1.5, 2, 3, 4, 4.5]          # exercises/ch4/ex1.py:385
IDP_LANES_PER_CLK_SM = 64                                   # measured: tools/ubench_int.cu, profiles/r1h_ubench_int.txt
FP64_LANES_PER_CLK_SM = 64
MMA_U8_MAC_PER_CLK_SM = 1935                                # mma.sync.m16n8k32.u8, measured (same file)


def candidates(H, W, sr):
    """in-bounds (dy, dx) pairs summed over the 8x8 blocks of a frame"""
    hp, wp = H // 8, W // 8
    rows = sum(min(sr, (hp - 1 - i) * 8) + min(sr, i * 8) + 1 for i in range(hp))
    cols = sum(min(sr, (wp - 1 - i) * 8) + min(sr, i * 8) + 1 for i in range(wp))
    return rows * cols


class Ctx:
    def __init__(self, torch, dist, device, rank, world, peak_gbs, sm_mhz=None, sms=148):
        self.torch, self.dist, self.device, self.rank, self.world = torch, dist, device, rank, world
        self.peak = peak_gbs
        self.sm_hz = (sm_mhz or 1965.0) * 1e6
        self.sms = sms

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world == 1:
            return float(v)
        t = self.torch.tensor([float(v)], dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, reps, warm=2):
        torch = self.torch
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


# ---- synthetic inputs (device-side, seeded) -----------------------------------------------------------------
def rgb_frames(torch, device, seeds, H, W):
    """smooth-noise RGB uint8 frames, one generator seed per frame (S3: frame i <- seed 3000 + i)"""
    out = torch.empty((len(seeds), H, W, 3), dtype=torch.uint8, device=device)
    g = torch.Generator(device=device)
    for i, seed in enumerate(seeds):
        g.manual_seed(int(seed))
        base = torch.randint(0, 256, (1, 3, H + 4, W + 4), generator=g, device=device, dtype=torch.int32).float()
        sm = torch.nn.functional.avg_pool2d(base, 5, stride=1)
        sm = (sm - 127.5) * 3.0 + 127.5 + 4.0 * torch.randn(sm.shape, generator=g, device=device)
        out[i] = sm.clamp_(0, 255)[0].permute(1, 2, 0).to(torch.uint8)
    return out


def luma_seq(torch, device, T, H, W, seed, shift=3):
    """integer-valued float64 luma frames: crops of one smooth canvas under per-frame global shifts + noise"""
    g = torch.Generator(device=device).manual_seed(int(seed))
    m = 4 * shift + 8
    base = torch.randint(0, 256, (1, 1, H + 2 * m, W + 2 * m), generator=g, device=device, dtype=torch.int32).float()
    canvas = ((torch.nn.functional.avg_pool2d(base, 5, stride=1, padding=2) - 127.5) * 3 + 127.5)[0, 0]
    out = torch.empty((T, H, W), dtype=torch.float64, device=device)
    sh = torch.randint(-shift, shift + 1, (T, 2), generator=g, device=device).cpu().tolist()
    for t, (dy, dx) in enumerate(sh):
        f = canvas[m + dy:m + dy + H, m + dx:m + dx + W] + torch.randn((H, W), generator=g, device=device)
        out[t] = f.round().clamp(0, 255)
    return out


# ---- cfg1: ch3 IntraCodec path on one 512x768 RGB image ------------------------------------------------------
def cfg1(cx, ivc):
    torch = cx.torch
    H, W = 512, 768
    rgb = rgb_frames(torch, cx.device, [0], H, W)[0]
    img = ivc.rgb2ycbcr(rgb)                                                     # float64 HWC, what the codec transforms
    c = ivc.IntraBlockCoder(1.0)
    D, Q, Z, P = ivc.DiscreteCosineTransform(), ivc.PatchQuant(1.0), ivc.ZigZag(), ivc.Patcher()
    t_f = cx.timed(lambda: c.inverse(c.forward(img)), 50, warm=5)
    chain = lambda x: D.inverse_transform(Q.dequantize(Z.unflatten(Z.flatten(Q.quantize(D.transform(P.patch(x)))))))
    t_dev = cx.timed(lambda: chain(img), 50, warm=5)
    h = img.cpu().numpy()
    chain(h)
    t0 = time.perf_counter()
    for _ in range(10):
        chain(h)
    t_np = (time.perf_counter() - t0) / 10 * 1e3
    rgb_np = rgb.cpu().numpy()
    ic = ivc.IntraCodec(1.0)
    ic.symbols2image(ic.image2symbols(rgb_np), rgb_np.shape)
    t0 = time.perf_counter()
    for _ in range(10):
        ic.symbols2image(ic.image2symbols(rgb_np), rgb_np.shape)
    t_c = (time.perf_counter() - t0) / 10 * 1e3
    px = H * W
    return {"workload": "one synthetic 512x768 RGB image, qScale 1: DCT, quantise, zig-zag and inverse",
            "fused_fwd_inv_ms": round(t_f, 4), "mpixel_s": round(px / t_f / 1e3, 1),
            "hbm_frac": round(px * 72 / (t_f * 1e-3) / 1e9 / cx.peak, 4),
            "bound": "launch latency (14 MB of traffic per call: two launches of a few microseconds each)",
            "six_method_calls_device_ms": round(t_dev, 4),
            "six_method_calls_numpy_in_numpy_out_ms": round(t_np, 3),
            "intracodec_image2symbols_symbols2image_numpy_ms": round(t_c, 3)}


# ---- cfg2: ch4 codec on a QCIF 21-frame luma sequence -----------------------------------------------------------
def cfg2(cx, ivc):
    seq = luma_seq(cx.torch, cx.device, 21, 144, 176, 2)
    out = {"workload": "QCIF 176x144, 21 luma frames, closed loop: +-4 order-exact full search, MC, residual DCT/quant, reconstruction (one fused kernel per frame)"}
    for graph in (False, True):
        cl = ivc.ClosedLoopLumaCoder(1.0, 4, decode="luma", me_mode="exact", use_graph=graph)
        t = cx.timed(lambda: cl.code_sequence(seq), 20, warm=3)
        out["cuda_graph" if graph else "direct"] = {"ms_per_sequence": round(t, 4), "us_per_frame": round(t / 21 * 1e3, 2),
                                                    "mpixel_s": round(21 * 144 * 176 / t / 1e3, 1)}
    out["us_per_frame"] = out["cuda_graph"]["us_per_frame"]
    out["mpixel_s"] = out["cuda_graph"]["mpixel_s"]
    out["bound"] = "launch latency / one wave (396 blocks per frame)"
    return out


# ---- cfg3: intra RD sweep over a fixed pool of 1080p frames, strong-scaled --------------------------------------
def cfg3(cx, ivc, pool=1024, chunk=8):
    torch = cx.torch
    from ivclab_b200.shard import gather_rows, shard_range
    from ivclab_b200.sweep import RateDistortionSweep
    H, W = 1080, 1920
    lo, hi = shard_range(pool, cx.rank, cx.world)
    Fl = hi - lo
    frames = rgb_frames(torch, cx.device, [3000 + i for i in range(lo, hi)], H, W)
    h_rgb = torch.empty(frames.shape, dtype=torch.uint8).pin_memory()
    h_rgb.copy_(frames)
    sweep = RateDistortionSweep(QS, chunk_frames=chunk, device=cx.device)
    Q, NB = len(QS), sweep.nb
    px = H * W

    # (a) device-resident: every RD point of the shard, CUDA events
    def resident(n_frames):
        for a in range(0, n_frames, chunk):
            sweep.code(frames[a:min(a + chunk, n_frames)])
    resident(min(Fl, 2 * chunk))
    torch.cuda.synchronize()
    cx.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    resident(Fl)
    b.record()
    torch.cuda.synchronize()
    ms_res = cx.max_over_ranks(a.elapsed_time(b))

    # (b) host-fed + gathered: upload once, 10 scales; PSNR + symbol histograms of every RD point end up in rank 0's
    #     host memory.  Preferred: every rank's downloads land directly in ONE shared pinned array (shard.SharedPinned:
    #     no collective, no second copy, overlapped with the coding); fallback: NCCL gather of device tensors + one copy.
    from ivclab_b200.shard import SharedPinned
    shared, how = None, None
    try:
        shared = {"sse": SharedPinned((Q, pool), torch.float64, "sse"), "hist": SharedPinned((Q, pool, NB), torch.int32, "hist"),
                  "outside": SharedPinned((Q, pool), torch.int32, "outside")}
        ok = 1
    except Exception as e:                                                     # no /dev/shm, registration refused, ...
        ok, how = 0, f"shared pinned segment unavailable ({type(e).__name__}: {e}); "
    if cx.world > 1:
        t = torch.tensor([ok], device=cx.device)
        cx.dist.all_reduce(t, op=cx.dist.ReduceOp.MIN)
        ok = int(t.item())
    if ok:
        out = {k: v.tensor for k, v in shared.items()}

        def fed():
            sweep.run(h_rgb, out=out, out_at=lo)                              # returns when this rank's downloads are complete
            t0 = time.perf_counter()
            cx.barrier()                                                       # ... and this when everybody's are: the gather
            return time.perf_counter() - t0
        how = "every rank's device-to-host copies write its frame range into one POSIX shared-memory array page-locked by all ranks " \
              "(shard.SharedPinned), overlapped with the coding; the gather itself is a barrier"
    else:
        out_sse = torch.empty((Q, pool), dtype=torch.float64).pin_memory() if cx.rank == 0 else None
        out_hist = torch.empty((Q, pool, NB), dtype=torch.int32).pin_memory() if cx.rank == 0 else None
        out_outs = torch.empty((Q, pool), dtype=torch.int32).pin_memory() if cx.rank == 0 else None
        out = {"sse": out_sse, "hist": out_hist, "outside": out_outs}

        def fed():
            r = sweep.run(h_rgb, to_host=False)
            sweep._s_cmp.synchronize()                       # so that the gather's own cost can be stated separately
            t0 = time.perf_counter()
            gather_rows(r["sse"], pool, axis=1, out=out_sse)
            gather_rows(r["hist"], pool, axis=1, out=out_hist)
            gather_rows(r["outside"], pool, axis=1, out=out_outs)
            torch.cuda.synchronize()
            return time.perf_counter() - t0
        how = (how or "") + "torch.distributed.gather of the per-rank result tensors (NCCL) + one device-to-host copy into rank 0's pinned buffers"
    sweep.run(h_rgb[:min(Fl, 2 * chunk)], to_host=False)                      # warm-up (allocator, kernel attributes)
    fed()
    cx.barrier()
    t0 = time.perf_counter()
    t_gather = fed()
    cx.barrier()
    ms_fed = cx.max_over_ranks((time.perf_counter() - t0) * 1e3)
    ms_gather = cx.max_over_ranks(t_gather * 1e3)
    res = {"workload": f"{pool} synthetic 1920x1080 RGB frames x {Q} qScales {QS}: contiguous frame ranges over "
                       f"{cx.world} rank(s) (strong scaling), every frame uploaded once (uint8 RGB), per RD point the "
                       "PSNR (RGB space, out of the decoder kernel) and the zero-run symbol histogram gathered on rank 0",
           "frames": pool, "qscales": Q, "frames_per_rank": Fl,
           "resident": {"ms": round(ms_res, 3), "mpixel_s": round(pool * Q * px / ms_res / 1e3, 1),
                        "us_per_rd_point": round(ms_res * 1e3 / (Fl * Q), 2),
                        # per RD point and pixel: RGB in 3 / Q + indices out 12 (forward, all scales in one kernel), indices
                        # in 12 (statistics), indices in 12 + RGB in 3 (decode + distortion) = 39 bytes + 3 / Q
                        "hbm_frac": round(Fl * px * (Q * 39 + 3) / (ms_res * 1e-3) / 1e9 / cx.peak, 4),
                        "bound": "FP64 pipe / issue of the decoder (dequantise + IDCT + ycbcr2rgb + error: 27 rounded FP64 ops per sample); "
                                 "colour transform + DCT run once per frame for all scales (ivc_intra_forward_rgb8_multi)"},
           "e2e": {"ms": round(ms_fed, 3), "mpixel_s": round(pool * Q * px / ms_fed / 1e3, 1),
                   "h2d_bytes": pool * px * 3, "d2h_bytes": Q * pool * (8 + 4 * NB + 4),
                   "gather_ms": round(ms_gather, 3), "gather": how + "; inside the timed region"}}
    res["mpixel_s"] = res["e2e"]["mpixel_s"]
    if cx.rank == 0:
        psnr = sweep.psnr(out["sse"].numpy(), px * 3)
        bits = sweep.entropy_bits(out["hist"].numpy())
        res["mean_psnr_db"] = [round(float(v), 3) for v in psnr.mean(axis=1)]
        res["mean_entropy_bpp"] = [round(float(v), 4) for v in (bits / px).mean(axis=1)]
        res["symbols_outside_histogram"] = int(out["outside"].numpy().sum())
        res["symbols_total"] = int(out["hist"].numpy().sum(dtype="int64"))
    del frames, h_rgb, out
    if shared is not None:
        for v in shared.values():
            v.close()
    torch.cuda.empty_cache()
    return res


# ---- cfg4: 8 x 120 x 4K, +-16 full search, one sequence per GPU ----------------------------------------------------
def cfg4(cx, ivc, n_seq=8, T=120, exact_pairs=4):
    torch = cx.torch
    from ivclab_b200.shard import shard_round_robin
    H, W, sr = 2160, 3840, 16
    mine = shard_round_robin(n_seq, cx.rank, cx.world)
    pc = ivc.PFrameBlockCoder(1.0, sr, me_mode="auto")
    pe = ivc.PFrameBlockCoder(1.0, sr, me_mode="exact")
    ms_int, ms_exact, checked = 0.0, None, None
    for s in mine:
        seq = luma_seq(torch, cx.device, T, H, W, 4000 + s, shift=12)
        pc.estimate(seq[:2], seq[1:3])
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        mv = pc.estimate(seq[:-1], seq[1:])                                   # open loop: frame t against the original t-1
        b.record()
        torch.cuda.synchronize()
        ms_int += a.elapsed_time(b)
        if s == mine[0] and exact_pairs:                                      # the order-exact FP64 kernel on a few pairs
            pe.estimate(seq[:1], seq[1:2])
            torch.cuda.synchronize()
            a.record()
            mve = pe.estimate(seq[:exact_pairs], seq[1:exact_pairs + 1])
            b.record()
            torch.cuda.synchronize()
            ms_exact = a.elapsed_time(b) / exact_pairs
            checked = bool(torch.equal(mve, mv[:exact_pairs]))
        del seq, mv
    cx.barrier()
    ms = cx.max_over_ranks(ms_int)
    pairs = n_seq * (T - 1)
    cand = candidates(H, W, sr)
    per_frame = ms / (len(mine) * (T - 1)) if mine else None
    lanes = cand * 16                                                         # dp4a instructions-lanes: 64 pixels / 4 per candidate
    pipe = IDP_LANES_PER_CLK_SM * cx.sms * cx.sm_hz
    res = {"workload": f"{n_seq} synthetic 3840x2160 luma sequences of {T} frames (float64, integer-valued), +-{sr} full search "
                       f"against the previous original frame, sequence s on rank s mod {cx.world} (strong scaling)",
           "ms": round(ms, 3), "frame_pairs": pairs, "mpixel_s": round(pairs * H * W / ms / 1e3, 1),
           "ms_per_frame_per_gpu": round(per_frame, 4) if per_frame else None,
           "kernel": "k_f64_to_u8 (every frame of the sequence converted and validated once) + k_me_mma16<uint8> (cross term on "
                     "mma.sync.m16n8k32.u8; IVC_ME_MMA=0 selects the dp4a kernel k_me_int<uint8,11,160>); both inside the timed region",
           "bound": "tensor pipe + shared-memory wavefronts + staging phases (profiles/README.md)",
           # 70 IMMA.16832 per block (4096 MAC each) against the measured mma.sync u8 rate; useful = 64 MAC per candidate
           "tensor_pipe_frac": round((H // 8) * (W // 8) * 70 * 4096 / (per_frame * 1e-3) / (MMA_U8_MAC_PER_CLK_SM * cx.sms * cx.sm_hz), 4) if per_frame else None,
           "tensor_pipe_peak": f"{MMA_U8_MAC_PER_CLK_SM} MAC/clk/SM (mma.sync u8, measured: profiles/r1h_ubench_int.txt) x {cx.sms} SMs x {cx.sm_hz / 1e6:.0f} MHz",
           "useful_mac_per_s": round(cand * 64 / (per_frame * 1e-3), 0) if per_frame else None,
           "dp4a_equivalent_idp_pipe_frac": round(lanes / (per_frame * 1e-3) / pipe, 4) if per_frame else None,
           "hbm_frac": round((2 * 8 * H * W + 8 * (H // 8) * (W // 8)) / (per_frame * 1e-3) / 1e9 / cx.peak, 4) if per_frame else None,
           "candidates_per_frame": cand}
    if ms_exact is not None:
        res["exact_fp64_kernel"] = {"kernel": "k_me_exact2 (byte prefilter, order-exact FP64 evaluation of the survivors)",
                                    "ms_per_frame": round(ms_exact, 4), "mpixel_s": round(H * W / ms_exact / 1e3, 1),
                                    "replay_equivalent_fp64_pipe_frac": round(cand * 192 / (ms_exact * 1e-3) / (FP64_LANES_PER_CLK_SM * cx.sms * cx.sm_hz), 4),
                                    "note": "the fraction is what evaluating EVERY candidate in FP64 at this speed would need; the kernel evaluates about one per block",
                                    "vectors_equal_integer_kernel": checked}
    torch.cuda.empty_cache()
    return res


# ---- cfg5: closed loop over 300 1080p frames, per-frame latency ------------------------------------------------------
def cfg5(cx, ivc, T=300):
    torch = cx.torch
    H, W = 1080, 1920
    s5 = luma_seq(torch, cx.device, T, H, W, 5000)
    res = {"workload": f"{T} synthetic 1920x1080 luma frames, closed loop (frame t is predicted from the decoder's reconstruction "
                       "of t-1): +-4 order-exact search, MC + residual DCT/quant, reconstruction in ONE kernel per frame (k_me_exact2<STEP>); replicas only across GPUs",
           "replicas": cx.world}
    for graph in (False, True):
        cl = ivc.ClosedLoopLumaCoder(1.0, 4, decode="luma", me_mode="exact", use_graph=graph)
        t = cx.max_over_ranks(cx.timed(lambda: cl.code_sequence(s5), 3, warm=1))
        res["cuda_graph" if graph else "direct"] = {"ms_per_sequence": round(t, 3), "us_per_frame": round(t / T * 1e3, 2),
                                                    "fps": round(T / t * 1e3, 1), "mpixel_s": round(T * H * W / t / 1e3, 1)}
        del cl
    best = min(res["cuda_graph"], res["direct"], key=lambda d: d["us_per_frame"])
    res["us_per_frame"] = best["us_per_frame"]
    res["mpixel_s"] = best["mpixel_s"]
    cand = candidates(H, W, 4)
    res["hbm_frac"] = round((16 + 12 + 8) * H * W / (best["us_per_frame"] * 1e-6) / 1e9 / cx.peak, 4)     # two frames in, scan blocks + reconstruction out
    res["replay_equivalent_fp64_pipe_frac"] = round(cand * 192 / (best["us_per_frame"] * 1e-6) / (FP64_LANES_PER_CLK_SM * cx.sms * cx.sm_hz), 4)
    res["bound"] = ("latency of ONE frame: 1020 tiles of 32 blocks on 444 resident CTAs = 2.3 waves, issue slots about half busy "
                    "(profiles/r2p_ncu_exact2.md); the prefilter leaves about one FP64 evaluation per block")
    del s5
    torch.cuda.empty_cache()
    return res


def run_all(cx, ivc, which=("cfg1", "cfg2", "cfg3", "cfg4", "cfg5"), cfg3_frames=1024, cfg4_frames=120, cfg5_frames=300):
    out = {}
    for name in which:
        cx.barrier()
        t0 = time.perf_counter()
        if name in ("cfg1", "cfg2"):
            r = (cfg1 if name == "cfg1" else cfg2)(cx, ivc) if cx.rank == 0 else None
        elif name == "cfg3":
            r = cfg3(cx, ivc, pool=cfg3_frames)
        elif name == "cfg4":
            r = cfg4(cx, ivc, T=cfg4_frames)
        elif name == "cfg5":
            r = cfg5(cx, ivc, T=cfg5_frames)
        else:
            raise ValueError(name)
        cx.barrier()
        if r is not None:
            r["wall_s_incl_input_generation"] = round(time.perf_counter() - t0, 2)
        out[name] = r
    return out


if __name__ == "__main__":
    import torch
    import ivclab_b200 as ivc
    quick = "--quick" in sys.argv
    peak = 6542.1
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    cx = Ctx(torch, None, torch.device("cuda", 0), 0, 1, peak)
    print(json.dumps(run_all(cx, ivc, cfg3_frames=64 if quick else 1024, cfg4_frames=12 if quick else 120,
                             cfg5_frames=30 if quick else 300), indent=1))
