"""GPU parity tests added in round 2 (``-m gpu``): the parity specification of SURVEY.md section 8(d) at full size
against the C oracle (cfg3: 8 frames x 10 qScales at 1080p; cfg4: 2 frame pairs x 8 sequences at +-16; cfg5:
teacher-forced frames 1..3 at 1080p), the unmodified reference ``IntraCodec`` running on the installed classes,
the rate-distortion sweep, the symbol statistics kernel, the two-channel P-frame stream and the device memo.
Tolerance 0 everywhere except summed squared errors (1e-12 relative: summation order)."""
import contextlib
import hashlib
import io
import os
import sys
import time

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden

pytestmark = pytest.mark.gpu

import ivclab_b200 as ivc  # noqa: E402
from ivclab_b200 import _runtime  # noqa: E402
from oracle import c_oracle as CO  # noqa: E402  (the checker)
from oracle import ivc_oracle as O  # noqa: E402

QS10 = [0.07, 0.2, 0.4, 0.8, 1.0, 1.5, 2, 3, 4, 4.5]                # exercises/ch4/ex1.py:385


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    assert torch.cuda.is_available(), "GPU tests selected but no CUDA device is visible"
    assert CO.available(), "oracle/_build/libivc_oracle.so missing (make -C oracle/c)"


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ---------------------------------------------------------------- section 8(d): cfg3 at full size
def test_cfg3_eight_1080p_frames_ten_qscales_vs_c_oracle():
    """S3 frames 0..7 x the ten qScales of ex1.py:385 plus an np.float64 scale (float64 table): scan
    indices and reconstructions bit-identical to the oracle, from float64 YCbCr and from uint8 RGB."""
    rgb = np.stack([O.smooth_noise_rgb(3000 + i, 1080, 1920) for i in range(8)])
    ycc = np.stack([O.rgb2ycbcr(f) for f in rgb])
    d_y, d_rgb = torch.from_numpy(ycc).cuda(), torch.from_numpy(rgb).cuda()
    for q in QS10 + [np.float64(0.4)]:
        coder = ivc.IntraBlockCoder(quantization_scale=q)
        tab = coder.quant.get_quantization_table()
        assert tab.dtype == (np.float64 if isinstance(q, np.float64) else np.float32)
        zz = coder.forward(d_y)
        assert torch.equal(zz, coder.forward_rgb(d_rgb))                      # colour transform fused in the load
        rec = coder.inverse(zz)
        zz_h, rec_h = zz.cpu().numpy(), rec.cpu().numpy()
        for i in range(8):
            zo = CO.intra_forward(ycc[i], tab, threads=8)
            assert np.array_equal(zz_h[i], zo), (q, i)
            assert np.array_equal(rec_h[i], CO.intra_inverse(zo, tab, threads=8)), (q, i)


# ---------------------------------------------------------------- section 8(d): cfg4 crops
@pytest.mark.parametrize("s", range(8))
def test_cfg4_two_pairs_per_sequence_sr16_int_and_exact(s):
    """S4: sequence s (seed 4000 + s, global shifts up to 12, 128x128 object), 256x256 crops of two frame pairs,
    +-16 search: integer kernel (auto), order-exact FP64 kernel and the oracle agree vector for vector."""
    seq = O.moving_sequence(4000 + s, 3, 512, 640, max_shift=12, obj=128)
    oy, ox = 64 + 8 * s, 128 + 16 * s
    crops = seq[:, oy:oy + 256, ox:ox + 256]
    for t in (1, 2):
        ref, cur = np.ascontiguousarray(crops[t - 1]), np.ascontiguousarray(crops[t])
        want = CO.me_full_search(ref, cur, 16, threads=8)
        assert np.array_equal(ivc.MotionCompensator(16).compute_motion_vector(ref, cur), want), (s, t, "auto")
        assert np.array_equal(ivc.MotionCompensator(16, me_mode="exact").compute_motion_vector(ref, cur), want), (s, t, "exact")
    # the uint8-plane entry of the integer kernel (what the host-fed pipeline searches on)
    pc = ivc.PFrameBlockCoder(1.0, 16)
    got = pc.estimate(crops[:2].astype(np.uint8), crops[1:].astype(np.uint8))
    assert np.array_equal(got[0], CO.me_full_search(crops[0], crops[1], 16, threads=8))


# ---------------------------------------------------------------- section 8(d): cfg5 teacher-forced
def test_cfg5_teacher_forced_frames_1_to_3_at_1080p():
    """S5 (seed 5000, 1080p, +-4, qScale 1): the closed loop of E4-1.py:249-306 with the ORACLE's reconstruction of
    frame t-1 as the reference of frame t (teacher forcing: a mismatch cannot hide behind, or compound through, the
    loop).  The reference frames are non-integer reconstructions, so the search is the order-exact FP64 kernel."""
    frames = O.moving_sequence(5000, 4, 1080, 1920)
    tab = O.quant_table(1.0)
    pc = ivc.PFrameBlockCoder(1.0, 4, me_mode="auto")
    pe = ivc.PFrameBlockCoder(1.0, 4, me_mode="exact")
    zz0 = CO.intra_forward(frames[0], tab, threads=8)
    recon = CO.intra_inverse(zz0[:, :, :1], tab, threads=8)[..., 0]          # luma decode of the I-frame
    assert np.array_equal(ivc.IntraBlockCoder(1.0).forward(frames[0]), zz0)
    for t in (1, 2, 3):
        cur = frames[t]
        mv = CO.me_full_search(recon, cur, 4, threads=8)
        assert np.array_equal(pc.estimate(recon, cur), mv), (t, "auto")
        assert np.array_equal(pe.estimate(recon, cur), mv), (t, "exact")
        pred = CO.mc_reconstruct(recon[..., None], mv, 4)[..., 0]
        zzp = CO.intra_forward(cur - pred, tab, threads=8)
        got_zz, got_pred = pc.forward(cur, recon, mv, return_prediction=True)
        assert np.array_equal(got_pred, pred) and np.array_equal(got_zz, zzp), t
        assert np.array_equal(pc.forward(cur, recon, mv, channels=2), zzp[:, :, :2]), t
        rec = pred + CO.intra_inverse(zzp[:, :, :1], tab, threads=8)[..., 0]
        assert np.array_equal(pc.inverse(got_zz, ref=recon, mv=mv), rec), t
        recon = rec
    assert abs(recon - frames[3]).max() < 64                                 # a reconstruction, not garbage


# ---------------------------------------------------------------- the unmodified reference codec on the installed classes
def _reference_root():
    for cand in (os.environ.get("IVCLAB_REFERENCE"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if cand and os.path.exists(os.path.join(cand, "ivclab", "image", "intracodec.py")):
            return cand
    return None


@pytest.mark.skipif(_reference_root() is None, reason="the reference tree is not on this box (set IVCLAB_REFERENCE, or "
                    "stage it with tools/stage_reference.py into the git-ignored baseline/_ref)")
def test_unmodified_reference_intracodec_on_installed_classes(g1, g6):
    """import the REAL ivclab package (matplotlib / constriction stubbed, as in
    oracle/gen_golden_video.py), `install()` the five classes, then run the unmodified IntraCodec.image2symbols /
    symbols2image and the P-frame branch of the working exercise codec (E4-1.py:249-306) on them.  Results must equal
    the goldens recorded from the reference running on its own classes (g6, g8, g10) bit for bit."""
    from oracle.gen_golden_video import _Any, _stub
    ref_root = _reference_root()
    saved = {k: v for k, v in sys.modules.items() if k == "ivclab" or k.startswith("ivclab.") or k in ("matplotlib", "matplotlib.pyplot", "constriction")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, ref_root)
    try:
        mp = _stub("matplotlib")
        mp.pyplot = _stub("matplotlib.pyplot", axes=_Any(), Axes=_Any())
        _stub("constriction", symbol=_Any())
        import ivclab.quantization, ivclab.signal, ivclab.utils, ivclab.video      # noqa: E401  leaf packages first
        done = ivc.install()
        assert {"ivclab.signal", "ivclab.quantization", "ivclab.utils", "ivclab.video"} <= set(done)
        from ivclab.image import IntraCodec                                         # binds the installed names (intracodec.py:3-5)
        import ivclab.image.intracodec as ic_mod
        assert ic_mod.DiscreteCosineTransform is ivc.DiscreteCosineTransform and ic_mod.PatchQuant is ivc.PatchQuant
        assert ic_mod.ZigZag is ivc.ZigZag and ic_mod.Patcher is ivc.Patcher
        assert ic_mod.__file__.startswith(ref_root) and IntraCodec is not ivc.IntraCodec
        from ivclab.video import MotionCompensator
        assert MotionCompensator is ivc.MotionCompensator
        g8, g10 = load_golden("g8_closed_loop.npz"), load_golden("g10_intracodec_cases.npz")
        with contextlib.redirect_stdout(io.StringIO()):
            # (1) intra, colour: the g1 image at qScale 1.0 -> g6 symbols and RGB reconstruction
            codec = IntraCodec(quantization_scale=1.0)
            assert isinstance(codec.dct, ivc.DiscreteCosineTransform) and isinstance(codec.quant, ivc.PatchQuant)
            sym = codec.image2symbols(g1["rgb"], is_source_rgb=True)
            assert np.array_equal(np.asarray(sym), g6["sym"])
            assert np.array_equal(codec.symbols2image(sym, g1["rgb"].shape), g6["rec_rgb"])
            # (2) cfg1 (S1) through the real codec: hashes recorded from the reference on its own classes
            rgb1 = O.smooth_noise_rgb(0, 512, 768)
            t0 = time.perf_counter()
            sym1 = np.asarray(codec.image2symbols(rgb1, is_source_rgb=True), dtype=np.int32)
            rec1 = codec.symbols2image(sym1, rgb1.shape)
            t_cfg1 = time.perf_counter() - t0
            assert sha(sym1) == str(g10["cfg1_sym_sha"]) and sha(rec1) == str(g10["cfg1_rec_sha"])
            # (3) the float32 / (H, W, 1) corner cases
            for name in ("luma32", "ycc32"):
                for qi, q in enumerate((0.07, 1.0)):
                    s = IntraCodec(quantization_scale=q).image2symbols(g10[name], is_source_rgb=False)
                    assert np.array_equal(np.asarray(s), g10[f"{name}_sym{qi}"]), (name, q)
            c4 = IntraCodec(quantization_scale=0.4)
            assert np.array_equal(c4.symbols2image(g10["hw1_sym"], (40, 56, 1)), g10["hw1_rec"])
            # (4) the closed loop of E4-1.py:212-306, driven as oracle/gen_golden_video.py drives it
            frames, q, sr = g8["frames"], float(g8["qscale"]), int(g8["sr"])
            H, W = frames.shape[1:]
            intra, resid = IntraCodec(quantization_scale=q), IntraCodec(quantization_scale=q)
            mc = MotionCompensator(search_range=sr)
            recon = None
            for t, y in enumerate(frames):
                if t == 0:
                    s = intra.image2symbols(y, is_source_rgb=False)
                    recon = intra.symbols2image(s, y.shape)
                    recon = recon[..., 0] if recon.ndim == 3 else recon
                    zz_t = intra.zerorun.decode(s, [H // 8, W // 8, 3])
                else:
                    mv = mc.compute_motion_vector(recon, y)
                    pred = mc.reconstruct_with_motion_vector(recon[..., None], mv)[..., 0]
                    s = resid.image2symbols(y - pred, is_source_rgb=False)
                    rr = resid.symbols2image(s, y.shape)
                    rr = rr[..., 0] if rr.ndim == 3 else rr
                    recon = pred + rr
                    assert np.array_equal(mv, g8["mv"][t - 1]), t
                    zz_t = resid.zerorun.decode(s, [H // 8, W // 8, 3])
                assert np.array_equal(zz_t, g8["zz"][t]) and np.array_equal(recon, g8["recon"][t]), t
        print(f"unmodified IntraCodec (reference at {ref_root}) on the installed B200 classes: g6 / g8 / g10 reproduced; "
              f"cfg1 image2symbols + symbols2image {t_cfg1 * 1e3:.1f} ms (host zero-run coder included)")
    finally:
        sys.path.remove(ref_root)
        for k in [k for k in sys.modules if k == "ivclab" or k.startswith("ivclab.") or k in ("matplotlib", "matplotlib.pyplot", "constriction")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_repo_intracodec_float32_and_hw1_cases():
    """float32 non-RGB input keeps float32 arithmetic; a (H, W, 1) shape returns the three-table decode
    unconverted -- against goldens recorded from the real IntraCodec (oracle/gen_golden_round2.py)."""
    g10 = load_golden("g10_intracodec_cases.npz")
    for name in ("luma32", "ycc32"):
        for qi, q in enumerate((0.07, 1.0)):
            s = ivc.IntraCodec(quantization_scale=q).image2symbols(g10[name], is_source_rgb=False)
            assert np.array_equal(s, g10[f"{name}_sym{qi}"]), (name, q)
    c = ivc.IntraCodec(quantization_scale=0.4)
    assert np.array_equal(c.image2symbols(g10["hw1_luma"], is_source_rgb=False), g10["hw1_sym"])
    rec = c.symbols2image(g10["hw1_sym"], (40, 56, 1))
    assert rec.shape == (40, 56, 3) and np.array_equal(rec, g10["hw1_rec"])
    rgb1 = O.smooth_noise_rgb(0, 512, 768)                                    # cfg1 through the repo's own codec
    c1 = ivc.IntraCodec(1.0)
    sym1 = c1.image2symbols(rgb1)
    assert sha(sym1.astype(np.int32)) == str(g10["cfg1_sym_sha"]) and sha(c1.symbols2image(sym1, rgb1.shape)) == str(g10["cfg1_rec_sha"])


# ---------------------------------------------------------------- device memo (drop-in chain without re-uploads)
def test_device_memo_chain_views_and_writability(g1):
    D, Q, Z, P = ivc.DiscreteCosineTransform(), ivc.PatchQuant(1.0), ivc.ZigZag(), ivc.Patcher()
    coef = D.transform(P.patch(g1["img"]))
    assert not coef.flags.writeable and np.array_equal(coef, g1["coef"])
    hit = _runtime._memo_get(coef, None)
    assert hit is not None and hit.is_cuda and hit.shape == coef.shape                     # served from the device
    zz = Z.flatten(Q.quantize(coef))
    assert np.array_equal(zz, g1["zz1"])
    rec = D.inverse_transform(Q.dequantize(Z.unflatten(zz)))
    assert np.array_equal(P.unpatch(rec), g1["rec1"])
    # views of a returned array are found too (offset + strides), and give the same bytes as an upload would
    view = rec[1:4, 2:, :, ::2, :]
    dv = _runtime._memo_get(view, None)
    assert dv is not None and np.array_equal(dv.cpu().numpy(), view)
    assert _runtime._memo_get(rec[::-1], None) is None                                     # negative strides: plain upload
    # the returned arrays cannot be written (numpy refuses to make a view of foreign memory writable again), so the
    # device copy can never go stale; a caller that wants to modify a result copies it, and the copy is uploaded
    coef2 = D.transform(P.patch(g1["img"]))
    with pytest.raises(ValueError):
        coef2[0, 0, 0, 0, 0] += 1000.0
    with pytest.raises(ValueError):
        coef2.flags.writeable = True
    mine = coef2.copy()
    mine[0, 0, 0, 0, 0] += 1000.0
    assert _runtime._memo_get(mine, None) is None
    assert Q.quantize(mine)[0, 0, 0, 0, 0] != Q.quantize(g1["coef"])[0, 0, 0, 0, 0]
    # switched off: writable arrays, no memo
    ivc.set_device_memo(False)
    try:
        c3 = D.transform(P.patch(g1["img"]))
        assert c3.flags.writeable and _runtime._memo_get(c3, None) is None and np.array_equal(c3, g1["coef"])
    finally:
        ivc.set_device_memo(True)


# ---------------------------------------------------------------- symbol statistics without the stream
def test_zerorun_symbol_histogram_equals_histogram_of_the_stream():
    rng = np.random.default_rng(5)
    frames = np.stack([O.rgb2ycbcr(O.smooth_noise_rgb(40 + i, 64, 96)) for i in range(3)])
    for q in (0.07, 1.0, 4.5):
        zz = ivc.IntraBlockCoder(q).forward(frames)
        counts, outside = ivc.zerorun_symbol_histogram(torch.from_numpy(zz).cuda(), lo=-4096, n_bins=8192)
        counts, outside = counts.cpu().numpy(), outside.cpu().numpy()
        for i in range(3):
            sym = O.zerorun_encode(zz[i])
            assert np.array_equal(counts[i], np.histogram(sym, bins=np.arange(-4096, 4097))[0]), (q, i)
            assert outside[i] == 0 and counts[i].sum() == sym.size
    # adversarial blocks: empty, full, a lone last coefficient, long runs, values at the range ends; ragged block count
    zz = np.where(rng.random((1, 7, 5, 3, 64)) < 0.2, rng.integers(-3000, 3000, (1, 7, 5, 3, 64)), 0).astype(np.int32)
    zz[0, 0, 0, 0] = 0
    zz[0, 0, 0, 1] = 7
    zz[0, 0, 0, 2] = 0
    zz[0, 0, 0, 2, 63] = -5
    zz[0, 1, 0, 0] = 0
    zz[0, 1, 0, 0, 0] = 1
    sym = O.zerorun_encode(zz[0])
    c, o = ivc.zerorun_symbol_histogram(zz[0], lo=-4096, n_bins=8192)
    assert np.array_equal(c.cpu().numpy(), np.histogram(sym, bins=np.arange(-4096, 4097))[0]) and int(o) == 0
    c, o = ivc.zerorun_symbol_histogram(zz[0], lo=-100, n_bins=150, end_of_block=4000)       # narrow range: the rest is counted outside
    inside = (sym >= -100) & (sym < 50)
    assert np.array_equal(c.cpu().numpy(), np.bincount(sym[inside] + 100, minlength=150)) and int(o) == int((~inside).sum())
    with pytest.raises(ValueError):
        ivc.zerorun_symbol_histogram(zz[0, :, :, :, :32])


# ---------------------------------------------------------------- rate-distortion sweep
def test_rate_distortion_sweep_vs_reference_statement():
    """cfg3's consumer: per (qScale, frame) PSNR of calc_psnr(img, symbols2image(image2symbols(img))) and the histogram
    of the symbols, host-fed in chunks; equal to the oracle's statement of the same chain."""
    F, H, W = 5, 64, 96
    rgb = np.stack([O.smooth_noise_rgb(90 + i, H, W) for i in range(F)])
    qs = (0.07, 1.0, np.float64(0.4))
    sw = ivc.RateDistortionSweep(qs, chunk_frames=2, slots=2)
    out = sw.run(rgb)
    assert out["sse"].shape == (3, F) and out["hist"].shape == (3, F, 8192) and int(out["outside"].sum()) == 0
    for qi, q in enumerate(qs):
        tab = O.quant_table(q)
        for f in range(F):
            zz = O.intra_forward(O.rgb2ycbcr(rgb[f]), tab)
            rec = O.ycbcr2rgb(O.intra_inverse(zz, tab))
            want = float(((rgb[f].astype(np.float64) - rec) ** 2).sum())
            assert abs(out["sse"][qi, f] / want - 1) < 1e-12, (q, f)
            assert abs(sw.psnr(out["sse"][qi, f], H * W * 3) - O.calc_psnr(rgb[f], rec)) < 1e-9
            sym = O.zerorun_encode(zz)
            assert np.array_equal(out["hist"][qi, f], np.histogram(sym, bins=np.arange(-4096, 4097))[0]), (q, f)
            pmf = O.stats_marg(sym, np.arange(sym.min() - 20, sym.max() + 21))      # intracodec.py:160-166
            nzb = np.flatnonzero(out["hist"][qi, f])
            lo_b, hi_b = nzb[0] - 4096 - 20, nzb[-1] - 4096 + 21
            assert (lo_b, hi_b) == (sym.min() - 20, sym.max() + 21)
            assert np.array_equal(out["hist"][qi, f][lo_b + 4096:hi_b + 4096 - 1] / sym.size, pmf)
    dev = sw.run(rgb, to_host=False)                                               # results left on the device
    torch.cuda.synchronize()
    assert np.array_equal(dev["hist"].cpu().numpy(), out["hist"]) and np.array_equal(dev["sse"].cpu().numpy(), out["sse"])
    # results downloaded straight into caller-owned (shared, page-locked) arrays at a column offset: the multi-GPU host gather
    from ivclab_b200.shard import SharedPinned
    keep = {k: out[k].copy() for k in ("sse", "hist", "outside")}
    sh = {"sse": SharedPinned((3, F + 3), torch.float64, "sse"), "hist": SharedPinned((3, F + 3, 8192), torch.int32, "hist"),
          "outside": SharedPinned((3, F + 3), torch.int32, "outside")}
    assert all(v.tensor.is_pinned() for v in sh.values())
    got = sw.run(rgb, out={k: v.tensor for k, v in sh.items()}, out_at=2)
    for k in keep:
        assert np.array_equal(got[k], keep[k]) and np.array_equal(sh[k].tensor[:, 2:2 + F].numpy(), keep[k]), k
    del got
    for v in sh.values():
        v.close()
    out = keep
    bits = sw.entropy_bits(out["hist"])
    assert bits.shape == (3, F) and np.all(bits[0] > bits[1])                      # finer quantisation costs more bits
    with pytest.raises(ValueError):
        sw.run(rgb[:, :, :40])


# ---------------------------------------------------------------- two-channel P-frame stream, int16 vectors
def test_pframe_two_channel_output_all_kernel_generations(g5, monkeypatch):
    pc = ivc.PFrameBlockCoder(0.4, 4)
    cur, ref, mv = g5["cur"], g5["ref"], g5["mv"]
    assert np.array_equal(pc.forward(cur, ref, mv, channels=2), g5["zz"][:, :, :2])
    seq = O.moving_sequence(77, 4, 72, 136)                                     # ragged tiles, a batch
    mvb = pc.estimate(seq[:-1], seq[1:])
    full = pc.forward(seq[1:], seq[:-1], mvb)
    assert np.array_equal(pc.forward(seq[1:], seq[:-1], mvb, channels=2), full[:, :, :, :2])
    # second generation (reference plane not 16-byte aligned) and first generation (IVC_FUSED_V1=1)
    d = torch.empty(seq[:-1].size + 1, dtype=torch.float64, device="cuda")
    r8 = d[1:].view(seq[:-1].shape)
    r8.copy_(torch.from_numpy(seq[:-1]))
    assert r8.data_ptr() % 16 == 8
    dc, dm = torch.from_numpy(seq[1:]).cuda(), torch.from_numpy(mvb).cuda()
    assert np.array_equal(pc.forward(dc, r8, dm, channels=2).cpu().numpy(), full[:, :, :, :2])
    monkeypatch.setenv("IVC_FUSED_V1", "1")
    assert np.array_equal(pc.forward(dc, r8, dm, channels=2).cpu().numpy(), full[:, :, :, :2])
    assert np.array_equal(pc.forward(dc, r8, dm).cpu().numpy(), full)
    monkeypatch.delenv("IVC_FUSED_V1")
    rec = pc.inverse(pc.forward(dc, r8, dm, channels=2), ref=r8, mv=dm)          # the decoder reads channel 0 of either layout
    assert torch.equal(rec, pc.inverse(torch.from_numpy(full).cuda(), ref=r8, mv=dm))
    with pytest.raises(ValueError):
        pc.forward(cur, ref, mv, channels=1)


def test_streamed_coder_two_channel_inter_stream_and_int16_vectors():
    F, H, W = 6, 64, 96
    rgb = np.stack([O.smooth_noise_rgb(120 + i, H, W) for i in range(F + 1)])
    luma = np.clip(np.round(np.stack([O.rgb2ycbcr(f)[..., 0] for f in rgb])), 0, 255)
    sc = ivc.StreamedCoder(0.4, 4, chunk_frames=4)
    assert sc.inter_channels == 2 and sc.mv_dtype == torch.int16
    out = sc.run(rgb[1:], first_ref=rgb[0])
    tab = O.quant_table(0.4)
    sym3, sym2 = [], []
    for i in range(1, F + 1):
        mv = O.me_full_search(luma[i - 1], luma[i], 4)
        assert out["mv"].dtype == torch.int16 and np.array_equal(out["mv"][i - 1].numpy().astype(np.int64), mv)
        _, zzp = O.pframe_forward(luma[i], luma[i - 1], mv, 4, tab)
        sym3.append(O.zerorun_encode(zzp))
        sym2.append(O.zerorun_encode(zzp[:, :, :2]))
    assert np.array_equal(out["sym_inter"].numpy().astype(np.int32), np.concatenate(sym2))
    assert np.array_equal(ivc.StreamedCoder.expand_inter(out["sym_inter"]), np.concatenate(sym3))   # the reference's stream
    full = ivc.StreamedCoder(0.4, 4, chunk_frames=4, inter_channels=3, mv_dtype=torch.int64).run(rgb[1:], first_ref=rgb[0])
    assert full["mv"].dtype == torch.int64 and np.array_equal(full["sym_inter"].numpy().astype(np.int32), np.concatenate(sym3))
    assert full["d2h_bytes"] > out["d2h_bytes"]
    assert ivc.StreamedCoder(0.4, 64).mv_dtype == torch.int16 and ivc.StreamedCoder(0.4, 64, mv_dtype=torch.int32).mv_dtype == torch.int32
    empty = sc.run(rgb[:0], first_ref=rgb[0])                                     # ADVICE: F == 0 returns empty results
    assert empty["sym_intra"].numel() == 0 and empty["len_intra"] == []


# ---------------------------------------------------------------- ADVICE: one graph coder, two shapes
def test_closed_loop_graph_coder_alternating_shapes():
    a = O.moving_sequence(31, 4, 48, 64)
    b = O.moving_sequence(32, 3, 64, 96)
    g = ivc.ClosedLoopLumaCoder(0.4, 4, decode="luma", me_mode="exact", use_graph=True)
    d = ivc.ClosedLoopLumaCoder(0.4, 4, decode="luma", me_mode="exact", use_graph=False)
    want_a, want_b = d.code_sequence(a), d.code_sequence(b)
    for _ in range(2):                                                            # a, b, a, b: each graph keeps its own zero plane
        for seq, want in ((a, want_a), (b, want_b)):
            torch.cuda.empty_cache()
            junk = torch.full((1 << 20,), 7.0, dtype=torch.float64, device="cuda")   # recycle freed blocks with non-zero bytes
            del junk
            got = g.code_sequence(seq)
            for k in ("zz", "mv", "recon"):
                assert np.array_equal(got[k], want[k]), k


def test_quantize_guard_at_the_int16_edge():
    """ADVICE: y in [32767.5, 32768) must round to 32768, not wrap (x = 327679, t = 10)."""
    x = np.zeros((1, 1, 3, 8, 8))
    x[0, 0, :, 0, 0] = [327679.0, 327675.0, -327679.0]
    tab10 = np.full((8, 8), 10.0, dtype=np.float32)
    pq = ivc.PatchQuant(1.0, luminance=tab10, chrominance=tab10)
    got = pq.quantize(x)
    want = np.int32(np.round(x / pq.get_quantization_table()[None, None]))
    assert np.array_equal(got, want) and got[0, 0, 0, 0, 0] == 32768
    img = np.zeros((8, 8, 3))
    img[..., 0] = 40959.875                                                      # DC = 8 * mean = 327679 -> y = 32767.9
    coder = ivc.IntraBlockCoder(1.0, luminance=tab10, chrominance=tab10)
    assert np.array_equal(coder.forward(img), O.intra_forward(img, coder.quant.get_quantization_table()))


def test_abi_call_leaves_current_device_alone():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    torch.cuda.set_device(0)
    x = torch.rand((16, 16, 3), dtype=torch.float64, device="cuda:1") * 255
    ivc.IntraBlockCoder(1.0).forward(x)
    assert torch.cuda.current_device() == 0
    assert torch.zeros(1, device="cuda").device.index == 0


# ---------------------------------------------------------------- fused search + P-frame forward
@pytest.mark.parametrize("mode", ["auto", "int", "exact"])
def test_fused_search_forward_equals_the_two_kernels(mode):
    """ivc_pframe_search_forward: one kernel (integer frames, +-4) or the stand-alone pair -- identical vectors and
    indices, every tile shape (small frames get small tiles), ragged edges, batches, 2- and 3-channel output."""
    for (n, H, W, seed) in ((1, 8, 8, 1), (1, 48, 64, 2), (3, 72, 136, 3), (2, 144, 176, 4), (1, 264, 1032, 5), (5, 200, 328, 6)):
        seq = O.moving_sequence(300 + seed, n + 1, H, W)
        ref, cur = seq[:-1], seq[1:]
        pc = ivc.PFrameBlockCoder(0.4, 4, me_mode=mode)
        mv_want = pc.estimate(ref, cur)
        for ch in (3, 2):
            mv, zz = pc.estimate_forward(ref, cur, channels=ch)
            assert np.array_equal(mv, mv_want), (n, H, W, ch)
            assert np.array_equal(zz, pc.forward(cur, ref, mv_want, channels=ch)), (n, H, W, ch)
        tab = O.quant_table(0.4)
        for i in range(n):                                                     # and against the oracle
            assert np.array_equal(mv[i], O.me_full_search(ref[i], cur[i], 4))
            _, zo = O.pframe_forward(cur[i], ref[i], mv[i], 4, tab)
            assert np.array_equal(zz[i], zo[:, :, :2])
    # non-integer frames: 'auto' falls back on the device (the fused kernel raises the flag, the pair runs)
    if mode != "int":
        rng = np.random.default_rng(9)
        seq = O.moving_sequence(77, 3, 72, 136) + rng.normal(0, 0.3, (3, 72, 136))
        pc = ivc.PFrameBlockCoder(1.0, 4, me_mode=mode)
        mv, zz = pc.estimate_forward(seq[:-1], seq[1:])
        for i in range(2):
            assert np.array_equal(mv[i], O.me_full_search(seq[i], seq[i + 1], 4))
            assert np.array_equal(zz[i], O.pframe_forward(seq[i + 1], seq[i], mv[i], 4, O.quant_table(1.0))[1])
    # other search ranges take the pair; flat frames (all ties) and frame edges
    for sr in (2, 7, 16):
        seq = O.moving_sequence(80 + sr, 2, 64, 96)
        pc = ivc.PFrameBlockCoder(1.0, sr, me_mode=mode)
        mv, zz = pc.estimate_forward(seq[0], seq[1])
        assert np.array_equal(mv, O.me_full_search(seq[0], seq[1], sr))
        assert np.array_equal(zz, O.pframe_forward(seq[1], seq[0], mv, sr, O.quant_table(1.0))[1])
    if mode != "exact":                                                        # uint8 planes: the fused kernel on the bytes as they are
        seq = O.moving_sequence(91, 3, 72, 136)
        s8 = torch.from_numpy(seq.astype(np.uint8)).cuda()
        pc = ivc.PFrameBlockCoder(0.4, 4, me_mode=mode)
        mv, zz = pc.estimate_forward(s8[:-1], s8[1:], channels=2)
        for i in range(2):
            assert np.array_equal(mv[i].cpu().numpy(), O.me_full_search(seq[i], seq[i + 1], 4))
            assert np.array_equal(zz[i].cpu().numpy(), O.pframe_forward(seq[i + 1], seq[i], mv[i].cpu().numpy(), 4, O.quant_table(0.4))[1][:, :, :2])
        odd = torch.empty(s8.numel() + 1, dtype=torch.uint8, device="cuda")[1:].view(s8.shape)     # unaligned planes: scalar staging
        odd.copy_(s8)
        mv2, zz2 = pc.estimate_forward(odd[:-1], odd[1:], channels=2)
        assert torch.equal(mv2, mv) and torch.equal(zz2, zz)
    flat = np.full((2, 40, 56), 255.0)
    mv, zz = ivc.PFrameBlockCoder(1.0, 4, me_mode=mode).estimate_forward(flat[0], flat[1])
    assert np.array_equal(mv, O.me_full_search(flat[0], flat[1], 4)) and not zz.any()


def test_fused_search_forward_1080p_batch_custom_tables():
    """full-size frames through the fused kernel, float64 quantisation table (np.float64 scale) and distinct chrominance
    tables (no channel-2 shortcut)"""
    seq = O.moving_sequence(5000, 3, 1080, 1920)
    d = torch.from_numpy(seq).cuda()
    for kw in ({"quantization_scale": np.float64(0.4)}, {"quantization_scale": 1.0, "chrominance": np.arange(1, 65, dtype=np.float32).reshape(8, 8)}):
        pc = ivc.PFrameBlockCoder(search_range=4, **kw)
        mv, zz = pc.estimate_forward(d[:-1], d[1:])
        mv2 = pc.estimate(d[:-1], d[1:])
        assert torch.equal(mv, mv2) and torch.equal(zz, pc.forward(d[1:], d[:-1], mv2))
    tab = O.quant_table(1.0)
    mv, zz = ivc.PFrameBlockCoder(1.0, 4).estimate_forward(d[:1], d[1:2])
    mvo = CO.me_full_search(seq[0], seq[1], 4, threads=8)
    assert np.array_equal(mv[0].cpu().numpy(), mvo)
    pred = CO.mc_reconstruct(seq[0][..., None], mvo, 4)[..., 0]
    assert np.array_equal(zz[0].cpu().numpy(), CO.intra_forward(seq[1] - pred, tab, threads=8))


def test_forward_kernels_emit_zero_run_counts_and_masks():
    """forward_rgb(zr=True) / estimate_forward(zr=True): the counts and masks of the zero-run coder's count pass, taken in
    the forward kernels -- equal to the stand-alone pass, and the symbol streams built from them equal the oracle's."""
    from ivclab_b200 import _lib as L
    zr = ivc.ZeroRunCoder()

    def count_pass(zz):
        n = zz.numel() // 64
        c = torch.empty(n, dtype=torch.int32, device="cuda")
        m = torch.empty(n, dtype=torch.int64, device="cuda")
        L.check(L.lib.ivc_zerorun_count_masks(0, torch.cuda.current_stream().cuda_stream, zz.data_ptr(), n, c.data_ptr(), m.data_ptr()), "count")
        return c, m
    for (n, H, W) in ((1, 8, 16), (2, 64, 96), (3, 72, 144), (1, 136, 272)):       # ragged tiles incl. a single block row
        rgb = torch.from_numpy(np.stack([O.smooth_noise_rgb(400 + i, H, W) for i in range(n)])).cuda()
        for q in (0.07, 1.0, 4.5):
            coder = ivc.IntraBlockCoder(q)
            zz, c, m = coder.forward_rgb(rgb, zr=True)
            assert torch.equal(zz, coder.forward_rgb(rgb))
            c2, m2 = count_pass(zz)
            assert torch.equal(c, c2) and torch.equal(m, m2), (n, H, W, q)
            sym = zr.encode_finish(zr.encode_begin(zz, counts=c, masks=m))
            want = np.concatenate([O.zerorun_encode_fast(z) for z in zz.cpu().numpy()])
            assert np.array_equal(sym.cpu().numpy(), want)
        seq = O.moving_sequence(500 + H, n + 1, H, W)
        for dt in (torch.float64, torch.uint8):
            d = torch.from_numpy(seq).cuda().to(dt)
            for ch in (2, 3):
                for sr in (4, 2):                                                  # fused kernel / the stand-alone pair
                    if dt == torch.uint8 and sr != 4:
                        continue
                    pc = ivc.PFrameBlockCoder(0.4, sr)
                    mv, zz, c, m = pc.estimate_forward(d[:-1], d[1:], channels=ch, zr=True)
                    mv2, zz2 = pc.estimate_forward(d[:-1], d[1:], channels=ch)
                    assert torch.equal(mv, mv2) and torch.equal(zz, zz2)
                    c2, m2 = count_pass(zz)
                    assert torch.equal(c, c2) and torch.equal(m, m2), (n, H, W, dt, ch, sr)
    # non-integer frames in auto mode: the gated fallback fills counts / masks as well
    seq = torch.from_numpy(O.moving_sequence(9, 3, 72, 136) + 0.25).cuda()
    mv, zz, c, m = ivc.PFrameBlockCoder(1.0, 4).estimate_forward(seq[:-1], seq[1:], zr=True)
    c2, m2 = count_pass(zz)
    assert torch.equal(c, c2) and torch.equal(m, m2)
    with pytest.raises(ValueError):
        ivc.IntraBlockCoder(1.0).forward_rgb(rgb.cpu().numpy(), zr=True)


# ---------------------------------------------------------------- fused closed-loop step (search + encoder + decoder half)
def test_closed_loop_fused_step_equals_three_kernels_and_oracle(monkeypatch):
    """decode='luma': one kernel per P-frame (ivc_pframe_step).  Same scan indices, vectors and reconstructions as the
    three stand-alone kernels and as the oracle's closed loop -- ragged tiles, tiny frames, lockstep batches, +-2 / +-7."""
    from oracle import closed_loop as CL
    g8 = load_golden("g8_closed_loop.npz")
    fused = ivc.ClosedLoopLumaCoder(float(g8["qscale"]), int(g8["sr"]), decode="luma")
    assert fused.fused_step
    got = fused.code_sequence(g8["frames"])
    assert np.array_equal(got["zz"], g8["zz_luma"]) and np.array_equal(got["mv"], g8["mv_luma"]) and np.array_equal(got["recon"], g8["recon_luma"])
    monkeypatch.setenv("IVC_CLOSED_LOOP_FUSED", "0")
    plain = ivc.ClosedLoopLumaCoder(0.4, 4, decode="luma")
    assert not plain.fused_step
    monkeypatch.delenv("IVC_CLOSED_LOOP_FUSED")
    for (T, H, W, sr, q) in ((3, 8, 8, 4, 1.0), (4, 40, 56, 4, 0.4), (3, 72, 136, 2, 0.4), (3, 144, 176, 7, 1.0), (3, 264, 520, 4, 0.07)):
        seq = O.moving_sequence(700 + H, T, H, W)
        a = ivc.ClosedLoopLumaCoder(q, sr, decode="luma").code_sequence(seq)
        monkeypatch.setenv("IVC_CLOSED_LOOP_FUSED", "0")
        b = ivc.ClosedLoopLumaCoder(q, sr, decode="luma").code_sequence(seq)
        monkeypatch.delenv("IVC_CLOSED_LOOP_FUSED")
        for k in ("zz", "mv", "recon"):
            assert np.array_equal(a[k], b[k]), (T, H, W, sr, q, k)
        if H <= 144:
            o = CL.code_sequence(seq, q, sr, "luma")
            for k in ("zz", "mv", "recon"):
                assert np.array_equal(a[k], o[k]), (T, H, W, sr, q, k)
    seqs = np.stack([O.moving_sequence(800 + i, 4, 48, 64) for i in range(3)])          # three sequences in lockstep
    lock = ivc.ClosedLoopLumaCoder(0.4, 4, decode="luma", use_graph=True).code_sequences(seqs)
    for i in range(3):
        one = ivc.ClosedLoopLumaCoder(0.4, 4, decode="luma").code_sequence(seqs[i])
        for k in ("zz", "mv", "recon"):
            assert np.array_equal(lock[k][i], one[k]), (i, k)
    assert not ivc.ClosedLoopLumaCoder(0.4, 4, decode="faithful").fused_step                # the reference's scrambled decode crosses tiles


# ---------------------------------------------------------------- exact search with the float32 prefilter: adversarial inputs
def test_exact_search_prefilter_adversarial_magnitudes_and_ties():
    """k_me_exact2 keeps a candidate unless float32 PROVES it loses; the vectors must equal the oracle's on inputs built to
    sit inside, at and beyond the error bound: candidates that differ by far less than float32 resolves, huge / tiny /
    negative magnitudes, NaN and Inf (no bound: every candidate is evaluated exactly), flat frames (all ties)."""
    rng = np.random.default_rng(42)
    base = O.moving_sequence(900, 2, 72, 136)
    mc = ivc.MotionCompensator(4, me_mode="exact")

    def check(ref, cur, sr=4, tag=""):
        got = ivc.MotionCompensator(sr, me_mode="exact").compute_motion_vector(ref, cur)
        assert np.array_equal(got, CO.me_full_search(ref, cur, sr, threads=8)), tag
    check(base[0] + 0.25, base[1], tag="plain")
    # periodic texture: many candidates with SSDs that agree to ~1e-9 relative (float32 cannot tell them apart)
    yy, xx = np.mgrid[0:72, 0:136]
    per = 128 + 100 * np.sin(2 * np.pi * xx / 4) * np.sin(2 * np.pi * yy / 4)
    check(per + 1e-7 * rng.normal(size=per.shape), per + 1e-7 * rng.normal(size=per.shape), tag="periodic 1e-7")
    check(per + 1e-11 * rng.normal(size=per.shape), per, tag="periodic 1e-11")
    check(per, per, tag="periodic exact ties")
    for scale in (1e-9, 1e-3, 1e3, 1e6, 3e14, 5e15, 1e30, 1e200):                   # 5e15 and beyond: no bound -> exact for all
        check((base[0] + 0.25) * scale, base[1] * scale, tag=f"scale {scale}")
    check(base[0] - 1000.0, base[1] - 1000.0 + 0.5, tag="negative offset")
    check(np.zeros((40, 56)), np.zeros((40, 56)), tag="zeros")
    check(np.full((40, 56), 1e-300), np.full((40, 56), 3e-300), tag="denormal-ish")
    for bad in (np.nan, np.inf, -np.inf):
        r, c = base[0] + 0.25, base[1].copy()
        r[10, 20] = bad
        c[50, 100] = bad
        check(r, c, tag=f"{bad} pixels")
    r = base[0] + 0.25
    r[::7, ::5] = 4e14                                                           # a few huge pixels next to ordinary ones
    check(r, base[1], tag="mixed magnitudes")
    for sr in (1, 2, 7, 16):
        check(base[0] + rng.normal(0, 0.3, base[0].shape), base[1], sr=sr, tag=f"sr {sr}")
    big = O.moving_sequence(901, 2, 264, 520)
    check(big[0] + rng.normal(0, 0.3, big[0].shape), big[1], tag="larger frame")


def test_exact_search_generations_agree_randomised():
    """The prefilter + survivors kernel against the kernel that replays numpy on every candidate: 150 random cases over
    shapes (8..240 x 8..320), ranges 1..16 and ten content classes (noise, integers, overshooting smooth scenes, low
    contrast, piecewise constant, tiny range on a 1e6 offset, scaled negatives, periodic exact ties, NaN / Inf pixels,
    static scene with an outlier) -- tools/exact_stress.py."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("exact_stress", os.path.join(ROOT, "tools", "exact_stress.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.run(seed=11, N=150, verbose=False) == []


def test_dct_other_norms_bit_identical_to_scipy():
    """`DiscreteCosineTransform(norm=...)`: the reference hands `norm` to scipy.fft.dct / idct (dct.py:9-10,24,26,42,44);
    None / 'backward' / 'forward' through ivc_dct8x8_norm against the oracle (pinned to scipy on the CPU suite) and against
    scipy itself where it is importable -- float64, float32, int32 and uint8 inputs, a strided patch view, and scipy's
    ValueError for anything else."""
    rng = np.random.default_rng(21)
    img = rng.uniform(-300, 300, size=(48, 64, 3))
    view = ivc.Patcher().patch(img)                                            # the strided [Hp, Wp, C, 8, 8] view
    cases = [view, rng.standard_normal((7, 5, 3, 8, 8)).astype(np.float32) * 100,
             rng.integers(-500, 500, size=(4, 6, 1, 8, 8)).astype(np.int32), rng.integers(0, 256, size=(3, 8, 8)).astype(np.uint8)]
    try:
        from scipy.fft import dct, idct
    except ImportError:                                                        # the oracle is the checker either way
        dct = idct = None
    for norm in (None, "backward", "forward", "ortho"):
        t = ivc.DiscreteCosineTransform(norm=norm)
        assert t.norm == norm
        for x in cases:
            f, b = t.transform(x), t.inverse_transform(x)
            assert f.dtype == (np.float32 if x.dtype == np.float32 else np.float64) and f.flags.c_contiguous
            assert np.array_equal(f, O.dct8x8_forward(np.ascontiguousarray(x), norm))
            assert np.array_equal(b, O.dct8x8_inverse(np.ascontiguousarray(x), norm))
            if dct is not None:
                assert np.array_equal(f, dct(dct(x, axis=-1, norm=norm), axis=-2, norm=norm))
                assert np.array_equal(b, idct(idct(x, axis=-1, norm=norm), axis=-2, norm=norm))
        d = t.transform(torch.from_numpy(np.ascontiguousarray(view)).cuda())    # CUDA tensor in -> CUDA tensor out
        assert d.is_cuda and np.array_equal(d.cpu().numpy(), O.dct8x8_forward(np.ascontiguousarray(view), norm))
    with pytest.raises(ValueError):
        ivc.DiscreteCosineTransform(norm="bogus").transform(view)


def test_forward_rgb_multi_equals_per_scale_forward():
    """One transform, many quantisations (the sweep's forward half): out[q] == IntraBlockCoder(q).forward_rgb(rgb) bit for
    bit for the ten scales of exercises/ch4/ex1.py:385, float32 and float64 tables mixed, more than 16 scales (grouping),
    with the zero-run counts / masks, on ragged tile counts (W = 16 * 7) and through the sweep itself."""
    rng = np.random.default_rng(31)
    rgb = torch.from_numpy(rng.integers(0, 256, size=(3, 40, 112, 3), dtype=np.uint8)).cuda()
    scales = list(QS10) + [np.float64(0.4), np.float64(2.5)] + [0.1 * k for k in range(1, 8)]            # 19 tables, two dtypes
    coders = [ivc.IntraBlockCoder(q) for q in scales]
    zz, counts, masks = ivc.forward_rgb_multi(coders, rgb, zr=True)
    assert zz.shape == (len(scales), 3, 5, 14, 3, 64)
    for qi, c in enumerate(coders):
        one, c1, m1 = c.forward_rgb(rgb, zr=True)
        assert torch.equal(zz[qi], one) and torch.equal(counts[qi], c1) and torch.equal(masks[qi], m1)
    assert torch.equal(ivc.forward_rgb_multi(coders[:3], rgb), zz[:3])
    o = O.intra_forward(O.rgb2ycbcr(rgb[1].cpu().numpy()), coders[4].quant.get_quantization_table())    # the oracle on one frame
    assert np.array_equal(zz[4, 1].cpu().numpy(), o)
    with pytest.raises(ValueError):
        ivc.forward_rgb_multi(coders, rgb[..., :100, :])
    with pytest.raises(ValueError):
        ivc.forward_rgb_multi(coders, rgb.cpu())


def test_wide_search_through_uint8_planes():
    """ivc_me_full_search with room for planes in the workspace (float64 frames, range >= 8): the frames are converted to
    uint8 once and the search reads the planes -- same vectors as the search that converts while staging (minimal
    workspace, through the raw ABI) and as the oracle; two views of one sequence and two separate tensors; AUTO mode with a
    non-integer frame still ends in the exact kernel."""
    from ivclab_b200 import _lib as L
    from ivclab_b200._runtime import code, dev_index, stream_ptr
    seq = torch.from_numpy(O.moving_sequence(123, 4, 72, 136)).cuda()              # integer-valued float64, ragged tiles
    ref, cur = seq[:-1], seq[1:]                                                   # views of one sequence
    N, H, W = ref.shape

    def raw(r, c, sr, mode, ws_bytes):
        mv = torch.empty((N, H // 8, W // 8, 1), dtype=torch.int64, device=r.device)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=r.device)
        st = L.lib.ivc_me_full_search(dev_index(r), stream_ptr(r.device), r.data_ptr(), c.data_ptr(), code(torch.float64), N, H, W,
                                      H * W, H * W, sr, mode, mv.data_ptr(), ws.data_ptr(), ws_bytes)
        L.check(st, "ivc_me_full_search")
        return mv
    small, big = L.lib.ivc_me_workspace_bytes(N, H, W), L.lib.ivc_me_workspace_bytes_planes(N, H, W)
    assert big == 256 + 2 * N * H * W
    for sr in (8, 16):
        want = np.stack([O.me_full_search(ref[i].cpu().numpy(), cur[i].cpu().numpy(), sr) for i in range(N)])
        for mode in (L.ME_AUTO, L.ME_INT):
            a = raw(ref, cur, sr, mode, small)                                      # converts while staging
            b = raw(ref, cur, sr, mode, big)                                        # one sequence: N + 1 planes
            c = raw(ref.clone(), cur.clone(), sr, mode, big)                        # two tensors: 2 N planes
            assert torch.equal(a, b) and torch.equal(a, c)
            assert np.array_equal(a.cpu().numpy().reshape(want.shape), want)
        pc = ivc.PFrameBlockCoder(1.0, sr)                                          # the class asks for the planes itself
        assert torch.equal(pc.estimate(ref, cur), a)
    noisy = seq + 0.25 * torch.rand_like(seq)                                       # not integer-valued: AUTO falls back on the device
    a = raw(noisy[:-1], noisy[1:], 8, L.ME_AUTO, big)
    b = raw(noisy[:-1], noisy[1:], 8, L.ME_EXACT, small)
    assert torch.equal(a, b)
