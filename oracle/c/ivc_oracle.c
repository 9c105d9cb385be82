/*
 * ivc_oracle.c -- plain C restatement of the ivclab per-block coding loop.
 * TEST INFRASTRUCTURE: used only by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
 * legs, never by the product package.  Parity status: pinned transitively -- tests/
 * test_oracle_cpu.py checks every function here bit-for-bit against oracle/ivc_oracle.py, which
 * oracle/gen_golden.py pins against the real reference modules.
 *
 * Build: see Makefile (gcc -O2 -ffp-contract=off: no FMA contraction, SSE2 double arithmetic).
 *
 * Reference behaviour restated (file:line under /root/reference):
 *   ivclab/signal/dct.py:24,26,42,44   scipy.fft.dct/idct(norm='ortho') on length 8 == ducc0's
 *                                      T_dcst23 over a radix-2/radix-4 real FFT, op for op
 *   ivclab/quantization/patchquant.py:59-60,77-78   round-half-even divide / truncating multiply
 *   ivclab/utils/shape.py:10-19,26,32  zig-zag scatter / gather
 *   ivclab/video/motion.py:28-57       full search, (dy,dx) raster order, strict <, np.sum order
 *   ivclab/video/motion.py:76-95       block copy, out-of-frame source -> zeros
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

static const double TW[7] = {0x1.f6297cff75cb0p-1, 0x1.d906bcf328d46p-1, 0x1.a9b66290ea1a3p-1,
                             0x1.6a09e667f3bccp-1, 0x1.1c73b39ae68c8p-1, 0x1.87de2a6aea963p-2,
                             0x1.8f8b83c69a60ap-3};
static const double WA0 = 0x1.6a09e667f3bccp-1, WA1 = 0x1.6a09e667f3bcdp-1;
static const double SQRT2 = 0x1.6a09e667f3bcdp+0;

static const unsigned char ZZ[64] = {
     0,  1,  5,  6, 14, 15, 27, 28,   2,  4,  7, 13, 16, 26, 29, 42,
     3,  8, 12, 17, 25, 30, 41, 43,   9, 11, 18, 24, 31, 40, 44, 53,
    10, 19, 23, 32, 39, 45, 52, 54,  20, 22, 33, 38, 46, 51, 55, 60,
    21, 34, 37, 47, 50, 56, 59, 61,  35, 36, 48, 49, 57, 58, 62, 63};

/* DCT-II, length 8, stride s, ducc0 order (un-simplified: every x2 / x0.25 / x0.5 is spelled out) */
void ivc_o_dct2_8(double *x, int s) {
    double c[8], ch[8], o[8];
    for (int i = 0; i < 8; ++i) c[i] = x[i * s];
    c[0] *= 2; c[7] *= 2;
    for (int k = 1; k < 7; k += 2) { double a = c[k + 1], b = c[k]; c[k + 1] = a - b; c[k] = b + a; }
    ch[0] = c[0] + c[7]; ch[4] = c[0] - c[7];
    ch[3] = 2 * c[3];    ch[7] = -2 * c[4];
    ch[1] = c[1] + c[5]; double tr2 = c[1] - c[5];
    double ti2 = c[2] + c[6]; ch[2] = c[2] - c[6];
    ch[6] = WA0 * ti2 + WA1 * tr2;
    ch[5] = WA0 * tr2 - WA1 * ti2;
    for (int k = 0; k < 2; ++k) {
        double t2 = ch[4 * k] + ch[4 * k + 3], t1 = ch[4 * k] - ch[4 * k + 3];
        double t3 = 2 * ch[4 * k + 1], t4 = 2 * ch[4 * k + 2];
        o[k] = t2 + t3; o[k + 4] = t2 - t3; o[k + 6] = t1 + t4; o[k + 2] = t1 - t4;
    }
    for (int i = 0; i < 8; ++i) o[i] *= 0.25;
    x[0] = o[0] * (SQRT2 * 0.5);
    for (int k = 1; k < 4; ++k) {
        int kc = 8 - k;
        double t1 = TW[k - 1] * o[kc] + TW[kc - 1] * o[k];
        double t2 = TW[k - 1] * o[k] - TW[kc - 1] * o[kc];
        x[k * s] = 0.5 * (t1 + t2); x[kc * s] = 0.5 * (t1 - t2);
    }
    x[4 * s] = o[4] * TW[3];
}

/* DCT-III (inverse), length 8, stride s */
void ivc_o_dct3_8(double *x, int s) {
    double c[8], ch[8], o[8];
    for (int i = 0; i < 8; ++i) c[i] = x[i * s];
    c[0] *= SQRT2;
    for (int k = 1; k < 4; ++k) {
        int kc = 8 - k;
        double t1 = c[k] + c[kc], t2 = c[k] - c[kc];
        c[k] = TW[k - 1] * t2 + TW[kc - 1] * t1;
        c[kc] = TW[k - 1] * t1 - TW[kc - 1] * t2;
    }
    c[4] *= 2 * TW[3];
    for (int k = 0; k < 2; ++k) {
        double t1 = c[k + 6] + c[k + 2]; ch[2 + 4 * k] = c[k + 6] - c[k + 2];
        double t2 = c[k] + c[k + 4];     ch[1 + 4 * k] = c[k] - c[k + 4];
        ch[4 * k] = t2 + t1; ch[3 + 4 * k] = t2 - t1;
    }
    o[0] = ch[0] + ch[4]; o[7] = ch[0] - ch[4];
    o[4] = -ch[7]; o[3] = ch[3];
    double tr2 = WA0 * ch[5] + WA1 * ch[6], ti2 = WA0 * ch[6] - WA1 * ch[5];
    o[1] = ch[1] + tr2; o[5] = ch[1] - tr2;
    o[2] = ti2 + ch[2]; o[6] = ti2 - ch[2];
    for (int i = 0; i < 8; ++i) o[i] *= 0.25;
    for (int k = 1; k < 7; k += 2) { double a = o[k], b = o[k + 1]; o[k] = a - b; o[k + 1] = b + a; }
    for (int i = 0; i < 8; ++i) x[i * s] = o[i];
}

static int32_t cast_i32(double r) {     /* numpy's float64 -> int32 cast on x86 (cvttsd2si) */
    return (r >= -2147483648.0 && r < 2147483648.0) ? (int32_t)r : INT32_MIN;
}

/* patch -> DCT -> quantize -> zig-zag.  img [H][W][C] (C in {1,3}) -> out [Hp][Wp][3][64] */
void ivc_o_intra_forward(const double *img, int64_t H, int64_t W, int C, const double *table, int32_t *out,
                         int64_t by0, int64_t by1) {
    const int64_t Wp = W / 8;
    for (int64_t by = by0; by < by1; ++by)
        for (int64_t bx = 0; bx < Wp; ++bx)
            for (int c = 0; c < C; ++c) {
                double b[64];
                for (int i = 0; i < 8; ++i)
                    for (int j = 0; j < 8; ++j) b[i * 8 + j] = img[((by * 8 + i) * W + bx * 8 + j) * C + c];
                for (int i = 0; i < 8; ++i) ivc_o_dct2_8(b + i * 8, 1);      /* axis -1 first */
                for (int j = 0; j < 8; ++j) ivc_o_dct2_8(b + j, 8);          /* then axis -2  */
                for (int ch = (C == 1 ? 0 : c); ch < (C == 1 ? 3 : c + 1); ++ch) {
                    int32_t *o = out + ((by * Wp + bx) * 3 + ch) * 64;
                    for (int k = 0; k < 64; ++k) o[ZZ[k]] = cast_i32(nearbyint(b[k] / table[ch * 64 + k]));
                }
            }
}

/* un-zig-zag -> dequantize -> IDCT -> un-patch.  zz [Hp][Wp][C][64] -> out [H][W][3] */
void ivc_o_intra_inverse(const int32_t *zz, int64_t Hp, int64_t Wp, int C, const double *table, double *out,
                         int64_t by0, int64_t by1) {
    const int64_t W = Wp * 8;
    (void)Hp;
    for (int64_t by = by0; by < by1; ++by)
        for (int64_t bx = 0; bx < Wp; ++bx)
            for (int ch = 0; ch < 3; ++ch) {
                const int32_t *q = zz + ((by * Wp + bx) * C + (C == 1 ? 0 : ch)) * 64;
                double b[64];
                for (int k = 0; k < 64; ++k) b[k] = (double)cast_i32((double)q[ZZ[k]] * table[ch * 64 + k]);
                for (int i = 0; i < 8; ++i) ivc_o_dct3_8(b + i * 8, 1);
                for (int j = 0; j < 8; ++j) ivc_o_dct3_8(b + j, 8);
                for (int i = 0; i < 8; ++i)
                    for (int j = 0; j < 8; ++j) out[((by * 8 + i) * W + bx * 8 + j) * 3 + ch] = b[i * 8 + j];
            }
}

#define ME_BODY(T)                                                                                     \
    const int64_t Wp = W / 8;                                                                          \
    const int span = 2 * sr + 1;                                                                       \
    for (int64_t by = by0; by < by1; ++by)                                                                \
        for (int64_t bx = 0; bx < Wp; ++bx) {                                                          \
            const int64_t y = by * 8, x = bx * 8;                                                      \
            int have = 0, bdy = 0, bdx = 0;                                                            \
            T best = 0;                                                                                \
            for (int dy = -sr; dy <= sr; ++dy)                                                         \
                for (int dx = -sr; dx <= sr; ++dx) {                                                   \
                    const int64_t yy = y + dy, xx = x + dx;                                            \
                    if (yy < 0 || yy + 8 > H || xx < 0 || xx + 8 > W) continue;                        \
                    T r[8];                                                                            \
                    for (int j = 0; j < 8; ++j) {                                                      \
                        T d = cur[y * W + x + j] - ref[yy * W + xx + j];                               \
                        r[j] = d * d;                                                                  \
                    }                                                                                  \
                    for (int i = 1; i < 8; ++i)                                                        \
                        for (int j = 0; j < 8; ++j) {                                                  \
                            T d = cur[(y + i) * W + x + j] - ref[(yy + i) * W + xx + j];               \
                            r[j] = r[j] + d * d;                                                       \
                        }                                                                              \
                    const T s = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));     \
                    if (!have ? (s < (T)INFINITY) : (s < best)) { have = 1; best = s; bdy = dy; bdx = dx; } \
                }                                                                                      \
            mv[by * Wp + bx] = (int64_t)(bdy + sr) * span + (bdx + sr);                                \
        }

/* full-search ME, numpy summation order (8 column accumulators, then the pairwise tree) */
void ivc_o_me_f64(const double *ref, const double *cur, int64_t H, int64_t W, int sr, int64_t *mv,
                  int64_t by0, int64_t by1) { ME_BODY(double) }
void ivc_o_me_f32(const float *ref, const float *cur, int64_t H, int64_t W, int sr, int64_t *mv,
                  int64_t by0, int64_t by1) { ME_BODY(float) }

void ivc_o_mc_f64(const double *ref, int64_t H, int64_t W, int64_t C, const int64_t *mv, int sr, double *out) {
    const int64_t Hp = H / 8, Wp = W / 8, span = 2 * (int64_t)sr + 1;
    memset(out, 0, (size_t)(H * W * C) * sizeof(double));
    for (int64_t by = 0; by < Hp; ++by)
        for (int64_t bx = 0; bx < Wp; ++bx) {
            int64_t idx = mv[by * Wp + bx], q = idx / span, r = idx % span;
            if (r < 0) { r += span; q -= 1; }
            const int64_t yy = by * 8 + q - sr, xx = bx * 8 + r - sr;
            if (yy < 0 || yy + 8 > H || xx < 0 || xx + 8 > W) continue;
            for (int i = 0; i < 8; ++i)
                memcpy(out + ((by * 8 + i) * W + bx * 8) * C, ref + ((yy + i) * W + xx) * C, (size_t)(8 * C) * sizeof(double));
        }
}

int ivc_o_version(void) { return 1; }
