"""Developer aid: P-frame kernels (K1p / K2p) against the oracle on small and ragged frames, then timing at 1080p.
    IVC_PFRAME=2|3 python tools/pframe_dev.py [--time] [--case H W]    (MV=zero|oob|idxN fixes the vectors)"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ivclab_b200 as ivc
from oracle import ivc_oracle as O

def check(H, W, sr, q, seed, nframes=1):
    rng = np.random.default_rng(seed)
    coder = ivc.PFrameBlockCoder(quantization_scale=q, search_range=sr)
    tab = coder.quant.get_quantization_table()
    ok = True
    refs = rng.uniform(0, 255, size=(nframes, H, W)); curs = rng.uniform(0, 255, size=(nframes, H, W))
    mvs = rng.integers(0, (2 * sr + 1) ** 2, size=(nframes, H // 8, W // 8, 1))
    if os.environ.get("MV") == "zero":
        mvs[:] = (2 * sr + 1) * sr + sr
    if os.environ.get("MV", "").startswith("idx"):
        mvs[:] = int(os.environ["MV"][3:])
    if os.environ.get("MV") == "oob":
        mvs[:] = 0
        mvs[:, 1:, 1:] = (2 * sr + 1) ** 2 - 1
    zz, pred = coder.forward(torch.from_numpy(curs).cuda(), torch.from_numpy(refs).cuda(), torch.from_numpy(mvs).cuda(), return_prediction=True)
    rec = coder.inverse(zz, ref=torch.from_numpy(refs).cuda(), mv=torch.from_numpy(mvs).cuda())
    torch.cuda.synchronize()
    zz, pred, rec = zz.cpu().numpy(), pred.cpu().numpy(), rec.cpu().numpy()
    for f in range(nframes):
        pred_o, zz_o = O.pframe_forward(curs[f], refs[f], mvs[f], sr, tab)
        rec_o = O.pframe_inverse(zz_o[:, :, :1], pred_o, tab)
        a, b, c = np.array_equal(pred[f], pred_o), np.array_equal(zz[f], zz_o), np.array_equal(rec[f], rec_o)
        if not (a and b and c):
            ok = False
            print(f"  MISMATCH H={H} W={W} frame {f}: pred {a} zz {b} rec {c}")
            if not a:
                bad = np.argwhere(pred[f] != pred_o)
                print("   first bad pred px", bad[:4].tolist(), "of", len(bad))
    print(f"H={H} W={W} sr={sr} q={q} frames={nframes}: {'ok' if ok else 'FAIL'}", flush=True)
    return ok

if __name__ == "__main__":
    print("IVC_PFRAME =", os.environ.get("IVC_PFRAME"))
    allok = True
    if "--case" in sys.argv:
        i = sys.argv.index("--case")
        H, W = int(sys.argv[i + 1]), int(sys.argv[i + 2])
        check(H, W, 4, 1.0, 3, 1)
        sys.exit(0)
    for (H, W, sr, q, n) in [(8, 8, 4, 1.0, 1), (16, 64, 4, 1.0, 1), (32, 208, 4, 0.2, 1), (144, 176, 4, 1.0, 3), (48, 72, 16, 0.07, 2), (1080, 1920, 4, 1.0, 1)]:
        allok &= check(H, W, sr, q, 7 + H + W, n)
    print("ALL OK" if allok else "SOME FAILED")
    if "--time" in sys.argv:
        F, H, W = 32, 1080, 1920
        g = torch.Generator(device="cuda").manual_seed(1)
        cur = torch.randint(0, 256, (F, H, W), device="cuda", generator=g).double()
        ref = torch.randint(0, 256, (F, H, W), device="cuda", generator=g).double()
        mv = torch.randint(0, 81, (F, H // 8, W // 8, 1), device="cuda", generator=g)
        coder = ivc.PFrameBlockCoder(quantization_scale=1.0, search_range=4)
        for name, fn in (("K1p", lambda: coder.forward(cur, ref, mv)),):
            zz = fn()
        inv = lambda: coder.inverse(zz, ref=ref, mv=mv)
        for name, fn, bpp in (("K1p", lambda: coder.forward(cur, ref, mv), 28), ("K2p", inv, 20)):
            for _ in range(5): fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): fn()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            print(f"{name}: {ms:.4f} ms / 32 frames, {F*H*W*bpp/ms/1e6:.0f} GB/s = {F*H*W*bpp/ms/1e6/6542.1*100:.1f} % of HBM peak")
