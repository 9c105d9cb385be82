#!/usr/bin/env python
"""Pin the oracle against the REAL reference and freeze golden vectors.

Run in the build container (where /root/reference exists):

    python oracle/gen_golden.py            # writes tests/golden/*.npz + PINNING.json

It loads the reference's four hot-path modules *by file path* (plain
``import ivclab`` needs matplotlib/constriction, which are not installed),
executes them on seeded inputs, asserts that every oracle function in
``oracle/ivc_oracle.py`` is bit-identical to them, and stores the
input/output vectors as small fixtures so that the tests can run where the
reference is absent (the GPU box).  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import contextlib
import hashlib
import importlib.util
import io
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ivc_oracle as O  # noqa: E402

REF = os.environ.get("IVCLAB_REFERENCE", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")


def load_ref(rel, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def sha(a: np.ndarray) -> str:
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest()


def main():
    dct_m = load_ref("ivclab/signal/dct.py", "ref_dct")
    pq_m = load_ref("ivclab/quantization/patchquant.py", "ref_pq")
    sh_m = load_ref("ivclab/utils/shape.py", "ref_shape")
    mo_m = load_ref("ivclab/video/motion.py", "ref_motion")
    co_m = load_ref("ivclab/signal/color.py", "ref_color")
    zr_m = load_ref("ivclab/entropy/zerorun.py", "ref_zerorun")
    me_m = load_ref("ivclab/utils/metrics.py", "ref_metrics")
    zz_m = load_ref("ivclab/signal/zigzag.py", "ref_zigzag_scan")

    DCT = dct_m.DiscreteCosineTransform()
    ZZ = sh_m.ZigZag()
    PATCH = sh_m.Patcher()
    report = {"reference": REF, "numpy": np.__version__, "checks": {}}
    import scipy
    report["scipy"] = scipy.__version__

    def check(name, ok, detail=""):
        report["checks"][name] = {"ok": bool(ok), "detail": detail}
        print(("PASS " if ok else "FAIL ") + name + (" " + detail if detail else ""))
        if not ok:
            raise SystemExit(f"oracle is NOT pinned: {name}")

    # ---- constants ------------------------------------------------------
    tw, wa = O.derive_ducc_constants()
    check("ducc_twiddles_rederived", np.array_equal(tw, O.DUCC_TW) and np.array_equal(wa, O.DUCC_WA))
    check("zigzag_table_equals_reference", np.array_equal(O.ZIGZAG_ORDER, ZZ.zigzag_order))
    blk = np.arange(64).reshape(8, 8)
    check("zigzag_equals_signal_zigzag_scan",
          np.array_equal(np.asarray(zz_m.zigzag_scan(blk)), O.ZIGZAG_SCAN))
    pq = pq_m.PatchQuant()
    check("quant_base_tables", np.array_equal(pq.luminance, O.LUMINANCE)
          and np.array_equal(pq.chrominance, O.CHROMINANCE))

    qscales = [0.07, 1.0, 4.5, np.float64(0.4)]

    # ---- G1: small colour image, all six transform ops -------------------
    rgb = O.smooth_noise_rgb(0, 48, 64)
    img = co_m.rgb2ycbcr(rgb)                       # f64 HWC, as IntraCodec feeds it
    check("rgb2ycbcr", np.array_equal(img, O.rgb2ycbcr(rgb)))
    patches = PATCH.patch(img)
    check("patch_view", np.array_equal(patches, O.patch(img)))
    coef = DCT.transform(patches)
    check("dct_fwd_f64_bitexact", np.array_equal(coef, O.dct8x8_forward(patches)), sha(coef)[:12])
    g1 = {"rgb": rgb, "img": img, "coef": coef}
    for qi, q in enumerate(qscales):
        pqq = pq_m.PatchQuant(quantization_scale=q)
        tab = pqq.get_quantization_table()
        check(f"quant_table[{q!r}]", np.array_equal(tab, O.quant_table(q)) and tab.dtype == O.quant_table(q).dtype,
              str(tab.dtype))
        qz = pqq.quantize(coef)
        check(f"quantize[{q!r}]", np.array_equal(qz, O.quantize(coef, tab)))
        with quiet():
            zz = ZZ.flatten(qz)
            un = ZZ.unflatten(zz)
        check(f"zigzag_flatten[{q!r}]", np.array_equal(zz, O.zigzag_flatten(qz)))
        check(f"zigzag_unflatten[{q!r}]", np.array_equal(un, O.zigzag_unflatten(zz)) and np.array_equal(un, qz))
        dq = pqq.dequantize(un)
        check(f"dequantize[{q!r}]", np.array_equal(dq, O.dequantize(un, tab)) and dq.dtype == np.int32)
        rec = DCT.inverse_transform(dq)
        check(f"dct_inv_bitexact[{q!r}]", np.array_equal(rec, O.dct8x8_inverse(dq)))
        check(f"intra_forward_fused[{q!r}]", np.array_equal(zz, O.intra_forward(img, tab)))
        check(f"intra_inverse_fused[{q!r}]", np.array_equal(PATCH.unpatch(rec), O.intra_inverse(zz, tab)))
        g1[f"table{qi}"] = tab
        g1[f"zz{qi}"] = zz
        g1[f"dq{qi}"] = dq
        g1[f"rec{qi}"] = PATCH.unpatch(rec)
    np.savez_compressed(os.path.join(GOLD, "g1_intra_color_48x64.npz"), **g1)

    # ---- G2: luma, integer valued, C=1 -> 3 broadcast, rounding ties -----
    luma = O.smooth_noise_luma(7, 40, 56)
    # force exact DC ties: block sum == 64 (mod 128) makes DC/16 = k + 0.5 at qScale 1
    luma[0:8, 0:8] = 100.0
    luma[0, 0] = 100.0 + 64.0               # sum = 6400+64 -> DC = 808 -> /16 = 50.5
    luma[8:16, 8:16] = 3.0                  # sum = 192 -> DC = 24 -> /16 = 1.5, tie again
    p1 = PATCH.patch(luma[..., None])
    coef1 = DCT.transform(p1)
    check("dct_fwd_luma_int_bitexact", np.array_equal(coef1, O.dct8x8_forward(p1)))
    g2 = {"luma": luma, "coef": coef1}
    for qi, q in enumerate(qscales):
        pqq = pq_m.PatchQuant(quantization_scale=q)
        tab = pqq.get_quantization_table()
        qz = pqq.quantize(coef1)
        check(f"quantize_bcast_C1[{q!r}]", qz.shape == (5, 7, 3, 8, 8) and np.array_equal(qz, O.quantize(coef1, tab)))
        with quiet():
            zz = ZZ.flatten(qz)
        dq = pqq.dequantize(qz[:, :, :1])
        check(f"dequantize_bcast_C1[{q!r}]", dq.shape == (5, 7, 3, 8, 8)
              and np.array_equal(dq, O.dequantize(qz[:, :, :1], tab)))
        rec = DCT.inverse_transform(dq)
        check(f"dct_inv_luma[{q!r}]", np.array_equal(rec, O.dct8x8_inverse(dq)))
        g2[f"zz{qi}"] = zz
        g2[f"rec{qi}"] = PATCH.unpatch(rec)
        g2[f"table{qi}"] = tab
    tie = np.abs(np.abs(coef1[..., 0, 0] / 16.0 % 1.0) - 0.5) < 1e-9
    check("ties_present_in_G2", tie.sum() >= 2, f"{int(tie.sum())} exact DC ties")
    np.savez_compressed(os.path.join(GOLD, "g2_intra_luma_ties_40x56.npz"), **g2)

    # ---- G3: other dtypes through each op (f32, uint8, int32, small shapes)
    rng = np.random.default_rng(11)
    x32 = rng.uniform(-300, 300, size=(3, 4, 3, 8, 8)).astype(np.float32)
    c32 = DCT.transform(x32)
    check("dct_fwd_f32_bitexact", c32.dtype == np.float32 and np.array_equal(c32, O.dct8x8_forward(x32)))
    i32 = DCT.inverse_transform(x32)
    check("dct_inv_f32_bitexact", i32.dtype == np.float32 and np.array_equal(i32, O.dct8x8_inverse(x32)))
    xu8 = rng.integers(0, 256, size=(2, 3, 1, 8, 8)).astype(np.uint8)
    cu8 = DCT.transform(xu8)
    check("dct_fwd_u8", cu8.dtype == np.float64 and np.array_equal(cu8, O.dct8x8_forward(xu8)))
    one = rng.uniform(0, 255, size=(8, 8))
    c_one = DCT.transform(one)
    check("dct_fwd_plain_8x8", np.array_equal(c_one, O.dct8x8_forward(one)))
    tab1 = pq_m.PatchQuant().get_quantization_table()
    q_u8 = pq_m.PatchQuant().quantize(xu8)          # tests/ch3.py:37-40 quantises raw pixels
    check("quantize_u8_pixels", np.array_equal(q_u8, O.quantize(xu8, tab1)))
    q_388 = pq_m.PatchQuant().quantize(c32[0, 0])   # (3,8,8) input -> (1,1,3,8,8)
    check("quantize_3x8x8_input", q_388.shape == (1, 1, 3, 8, 8) and np.array_equal(q_388, O.quantize(c32[0, 0], tab1)))
    q_f32 = pq_m.PatchQuant(0.07).quantize(c32)
    np.savez_compressed(os.path.join(GOLD, "g3_dtypes.npz"), x32=x32, c32=c32, i32=i32, xu8=xu8, cu8=cu8,
                        one=one, c_one=c_one, q_u8=q_u8, q_f32=q_f32,
                        tab007=pq_m.PatchQuant(0.07).get_quantization_table(), tab1=tab1)

    # ---- G4: motion estimation / compensation ----------------------------
    seq = O.moving_sequence(2, 4, 48, 64)
    g4 = {"seq": seq}
    cases = []
    rngm = np.random.default_rng(5)
    ref_f = seq[0] + rngm.normal(0, 0.37, size=seq[0].shape)      # non-integer decoder-like reference
    cases.append(("int_sr4", seq[0], seq[1], 4))
    cases.append(("int_sr2", seq[1], seq[2], 2))
    cases.append(("int_sr16", seq[2], seq[3], 16))
    cases.append(("float_sr4", ref_f, seq[1], 4))
    cases.append(("float_sr7", ref_f, seq[2] + 0.25, 7))
    cases.append(("flat_sr4", np.zeros((32, 40)), np.zeros((32, 40)), 4))
    cases.append(("flat255_sr3", np.full((24, 32), 255.0), np.full((24, 32), 255.0), 3))
    cases.append(("f32_sr4", ref_f.astype(np.float32), seq[1].astype(np.float32), 4))
    cases.append(("mixed_f32ref_f64cur_sr3", ref_f.astype(np.float32), seq[1], 3))
    cases.append(("tiny_8x8_sr4", seq[0][:8, :8].copy(), seq[1][:8, :8].copy(), 4))
    cases.append(("row_8x64_sr5", seq[0][:8].copy(), seq[1][:8].copy(), 5))
    t_loop = 0.0
    for name, r, c, sr in cases:
        mc = mo_m.MotionCompensator(search_range=sr)
        t0 = time.perf_counter()
        mv = mc.compute_motion_vector(r, c)
        t_loop += time.perf_counter() - t0
        ok = (np.array_equal(mv, O.me_full_search(r, c, sr)) and mv.dtype == np.int64
              and np.array_equal(mv, O.me_full_search_loops(r, c, sr)))
        check(f"me[{name}]", ok, f"shape {mv.shape}")
        pred = mc.reconstruct_with_motion_vector(r[..., None], mv)
        check(f"mc[{name}]", np.array_equal(pred, O.mc_reconstruct(r[..., None], mv, sr)) and pred.dtype == r.dtype)
        g4[f"{name}__ref"] = r
        g4[f"{name}__cur"] = c
        g4[f"{name}__mv"] = mv
        g4[f"{name}__pred"] = pred
    # MC with arbitrary (incl. out-of-frame) vectors and C=3
    mc = mo_m.MotionCompensator(search_range=4)
    ref3 = rngm.uniform(0, 255, size=(32, 40, 3))
    mv_rand = rngm.integers(0, 81, size=(4, 5, 1))
    pred3 = mc.reconstruct_with_motion_vector(ref3, mv_rand)
    check("mc_oob_C3", np.array_equal(pred3, O.mc_reconstruct(ref3, mv_rand, 4)))
    g4.update(mc_ref3=ref3, mc_mv_rand=mv_rand, mc_pred3=pred3)
    np.savez_compressed(os.path.join(GOLD, "g4_motion.npz"), **g4)

    # ---- G5: P-frame step (MC + residual + transform path + recon), E4-1.py:257-306 style
    sr = 4
    mc = mo_m.MotionCompensator(search_range=sr)
    pqq = pq_m.PatchQuant(quantization_scale=0.4)
    tab = pqq.get_quantization_table()
    cur = seq[1]
    mv = mc.compute_motion_vector(ref_f, cur)
    pred = mc.reconstruct_with_motion_vector(ref_f[..., None], mv)[..., 0]
    resid = cur - pred
    with quiet():
        zz = ZZ.flatten(pqq.quantize(DCT.transform(PATCH.patch(resid[..., None]))))
        rec3 = PATCH.unpatch(DCT.inverse_transform(pqq.dequantize(ZZ.unflatten(zz[:, :, :1]))))
    recon = pred + rec3[..., 0]
    p2, z2 = O.pframe_forward(cur, ref_f, mv, sr, tab)
    check("pframe_forward", np.array_equal(p2, pred) and np.array_equal(z2, zz))
    check("pframe_inverse", np.array_equal(O.pframe_inverse(zz[:, :, :1], pred, tab), recon))
    np.savez_compressed(os.path.join(GOLD, "g5_pframe.npz"), ref=ref_f, cur=cur, mv=mv, pred=pred,
                        zz=zz, recon=recon, table=tab)

    # ---- G6: zero-run coder + metrics (neighbours of the path) -----------
    zz1 = g1["zz1"]
    with quiet():
        sym = zr_m.ZeroRunCoder().encode(zz1)
        dec = zr_m.ZeroRunCoder().decode(sym, list(zz1.shape[:3]))
        dec_trunc = zr_m.ZeroRunCoder().decode(sym, [zz1.shape[0], zz1.shape[1], 1])
    check("zerorun_encode", np.array_equal(sym, O.zerorun_encode(zz1)))
    check("zerorun_decode", np.array_equal(dec, O.zerorun_decode(sym, zz1.shape[:3])) and np.array_equal(dec, zz1))
    check("zerorun_decode_truncating", np.array_equal(dec_trunc, O.zerorun_decode(sym, (zz1.shape[0], zz1.shape[1], 1))))
    rec_rgb = co_m.ycbcr2rgb(g1["rec1"])
    check("ycbcr2rgb", np.array_equal(rec_rgb, O.ycbcr2rgb(g1["rec1"])))
    mse = me_m.calc_mse(rgb, rec_rgb)
    psnr = me_m.calc_psnr(rgb, rec_rgb)
    check("mse_psnr", mse == O.calc_mse(rgb, rec_rgb) and psnr == O.calc_psnr(rgb, rec_rgb), f"psnr={psnr:.4f}")
    np.savez_compressed(os.path.join(GOLD, "g6_neighbours.npz"), sym=sym, dec_trunc=dec_trunc,
                        rec_rgb=rec_rgb, mse=mse, psnr=psnr)

    # ---- G7: full-size configs, stored as hashes (inputs are re-generated from seeds)
    big = {}
    rgb1 = O.smooth_noise_rgb(0, 512, 768)                      # cfg1
    img1 = co_m.rgb2ycbcr(rgb1)
    big["cfg1_img_sha"] = sha(img1)
    p = PATCH.patch(img1)
    t0 = time.perf_counter()
    coef = DCT.transform(p)
    t_dct = time.perf_counter() - t0
    check("cfg1_dct_fwd_bitexact", np.array_equal(coef, O.dct8x8_forward(p)))
    for q in (0.07, 1.0, 4.5):
        pqq = pq_m.PatchQuant(quantization_scale=q)
        tab = pqq.get_quantization_table()
        with quiet():
            zz = ZZ.flatten(pqq.quantize(coef))
            rec = PATCH.unpatch(DCT.inverse_transform(pqq.dequantize(ZZ.unflatten(zz))))
        check(f"cfg1_fwd[{q}]", np.array_equal(zz, O.intra_forward(img1, tab)))
        check(f"cfg1_inv[{q}]", np.array_equal(rec, O.intra_inverse(zz, tab)))
        big[f"cfg1_zz_sha[{q}]"] = sha(zz)
        big[f"cfg1_rec_sha[{q}]"] = sha(rec)
        big[f"cfg1_psnr_ycbcr[{q}]"] = me_m.calc_psnr(img1, rec)
    # cfg2: QCIF sequence, ME of frame t against frame t-1 (open loop) with the real loops
    seq2 = O.moving_sequence(2, 6, 144, 176)
    big["cfg2_seq_sha"] = sha(seq2)
    mc = mo_m.MotionCompensator(search_range=4)
    mvs = []
    t0 = time.perf_counter()
    for t in range(1, 6):
        mvs.append(mc.compute_motion_vector(seq2[t - 1], seq2[t]))
    t_me = (time.perf_counter() - t0) / 5
    mvs = np.stack(mvs)
    ok = all(np.array_equal(mvs[t - 1], O.me_full_search(seq2[t - 1], seq2[t], 4)) for t in range(1, 6))
    check("cfg2_me_qcif_open_loop", ok, f"{t_me:.3f} s/frame reference loops")
    # non-integer reference (decoder-like), one QCIF frame
    ref_q = seq2[0] + np.random.default_rng(9).normal(0, 0.5, size=seq2[0].shape)
    mv_q = mc.compute_motion_vector(ref_q, seq2[1])
    check("cfg2_me_qcif_float_ref", np.array_equal(mv_q, O.me_full_search(ref_q, seq2[1], 4)))
    np.savez_compressed(os.path.join(GOLD, "g7_qcif_mv.npz"), mvs=mvs.astype(np.int16), mv_float=mv_q.astype(np.int16))
    report["hashes"] = big
    report["reference_timings_here"] = {"dct_fwd_cfg1_s": t_dct, "me_qcif_s_per_frame": t_me}
    with open(os.path.join(GOLD, "PINNING.json"), "w") as f:
        json.dump(report, f, indent=1, default=str)
    n = len(report["checks"])
    print(f"\noracle pinned: {n}/{n} checks bit-identical to the reference at {REF}")


if __name__ == "__main__":
    main()
