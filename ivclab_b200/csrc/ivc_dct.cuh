// ivc_dct.cuh -- order-exact 8-point DCT-II / DCT-III and exact quantiser arithmetic.
//
// The reference computes its transforms with scipy.fft.dct/idct (ivclab/signal/dct.py:24,26,42,44),
// i.e. ducc0's T_dcst23 on a length-8 real FFT.  The functions below evaluate the SAME rounded
// operations in the SAME order (individually rounded add/sub/mul, never contracted to FMA), so the
// results are bit-identical to scipy.  Two exact simplifications are applied:
//   * every multiplication by 2 / 0.5 / 0.25 in ducc0 is a power-of-two scaling, which commutes
//     with rounding; all of them are folded into the final twiddle constants (64 -> 56 operations);
//   * "a + (-b)" is emitted as "a - b".
// oracle/ivc_oracle.py::dct2_8 / dct3_8 is the un-simplified numpy statement of the same sequence
// (pinned bit-for-bit against scipy by oracle/gen_golden.py); tests compare the two on the GPU.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ivc {

// ---- individually rounded arithmetic (immune to -fmad contraction) ----------------------------
template <typename T> struct Rn;
template <> struct Rn<double> {
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
};
template <> struct Rn<float> {
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
};

// ---- ducc0's length-8 constants ---------------------------------------------------------------
// TW[i] = Re(UnityRoots<double>(32)[i+1]), evaluated by ducc0 in double from a rounded angle, hence
// up to 3 ulp away from the correctly rounded cosine.  WA = UnityRoots(8)[1].  Same literals as
// oracle/ivc_oracle.py::DUCC_TW / DUCC_WA.  For float, ducc0 casts the double values.
template <typename T> struct Ducc {
    static constexpr double TW0 = 0x1.f6297cff75cb0p-1;
    static constexpr double TW1 = 0x1.d906bcf328d46p-1;
    static constexpr double TW2 = 0x1.a9b66290ea1a3p-1;
    static constexpr double TW3 = 0x1.6a09e667f3bccp-1;
    static constexpr double TW4 = 0x1.1c73b39ae68c8p-1;
    static constexpr double TW5 = 0x1.87de2a6aea963p-2;
    static constexpr double TW6 = 0x1.8f8b83c69a60ap-3;
    static constexpr double WA0 = 0x1.6a09e667f3bccp-1;
    static constexpr double WA1 = 0x1.6a09e667f3bcdp-1;
    static constexpr double SQRT2 = 0x1.6a09e667f3bcdp+0;   // double(1.41421356237309504880L); its float cast equals float(long double)
    // constants in T, with the exact power-of-two scalings folded in (see header comment)
    static __device__ __forceinline__ T tw(int i, T scale) {
        const double v = (i == 0) ? TW0 : (i == 1) ? TW1 : (i == 2) ? TW2 : (i == 3) ? TW3
                       : (i == 4) ? TW4 : (i == 5) ? TW5 : TW6;
        return (T)v * scale;       // (T)v rounds once (float case); * 2^k is exact
    }
    static __device__ __forceinline__ T wa0() { return (T)WA0; }
    static __device__ __forceinline__ T wa1() { return (T)WA1; }
    static __device__ __forceinline__ T sqrt2(T scale) { return (T)SQRT2 * scale; }
};

// scipy's `norm` (dct.py:24,26,42,44 forward it): ducc0 multiplies the FFT output by fct and, for "ortho", the DC term by
// sqrt2 / 2 (type 2) or sqrt2 (type 3).  For length 8 fct is a power of two in all three modes -- forward transform: ortho
// 1/4, backward 1, forward 1/16; the inverse takes the complementary factor -- so it folds into the constants exactly.
enum { kNormOrtho = 0, kNormBackward = 1, kNormForward = 2 };
template <typename T, int NORM, bool INVERSE>
__device__ __forceinline__ T dct_fct() {
    return NORM == kNormOrtho ? (T)0.25 : ((NORM == kNormForward) != INVERSE) ? (T)0.0625 : (T)1.0;
}

// ---- forward: DCT-II of x[0..7], in place (orthonormal unless NORM says otherwise) --------------
template <typename T, int NORM = kNormOrtho>
__device__ __forceinline__ void dct2_8(T (&x)[8]) {
    using R = Rn<T>;
    using K = Ducc<T>;
    // T_dcst23::exec type 2 pre-step: MPINPLACE(c[k+1], c[k]) for k = 1,3,5
    const T a1 = R::add(x[1], x[2]), a2 = R::sub(x[2], x[1]);
    const T a3 = R::add(x[3], x[4]), a4 = R::sub(x[4], x[3]);
    const T a5 = R::add(x[5], x[6]), a6 = R::sub(x[6], x[5]);
    // radb2 (ido=4, l1=1); the reference's 2*c0, 2*c7, 2*c3, -2*c4 doublings are folded out
    const T u = R::add(x[0], x[7]), v = R::sub(x[0], x[7]);
    const T h1 = R::add(a1, a5), tr2 = R::sub(a1, a5);
    const T ti2 = R::add(a2, a6), h2 = R::sub(a2, a6);
    const T h6 = R::add(R::mul(K::wa0(), ti2), R::mul(K::wa1(), tr2));
    const T h5 = R::sub(R::mul(K::wa0(), tr2), R::mul(K::wa1(), ti2));
    // radb4 (ido=1, l1=2)
    T o[8];
    {
        const T p = R::add(u, a3), q = R::sub(u, a3);
        o[0] = R::add(p, h1); o[4] = R::sub(p, h1);
        o[6] = R::add(q, h2); o[2] = R::sub(q, h2);
    }
    {
        const T p = R::sub(v, a4), q = R::add(v, a4);
        o[1] = R::add(p, h5); o[5] = R::sub(p, h5);
        o[7] = R::add(q, h6); o[3] = R::sub(q, h6);
    }
    // o == (reference's FFT output) / 2.  ortho: fct = 0.25, final 0.5*(t1 +- t2): constants carry 0.25.
    const T s = dct_fct<T, NORM, false>();
    x[0] = R::mul(o[0], NORM == kNormOrtho ? K::sqrt2(s) : (T)2 * s);   // * fct * (sqrt2*0.5) * 2; without ortho an exact scaling
#pragma unroll
    for (int k = 1; k <= 3; ++k) {
        const int kc = 8 - k;
        const T twk = K::tw(k - 1, s), twc = K::tw(kc - 1, s);
        const T t1 = R::add(R::mul(twk, o[kc]), R::mul(twc, o[k]));
        const T t2 = R::sub(R::mul(twk, o[k]), R::mul(twc, o[kc]));
        x[k] = R::add(t1, t2);
        x[kc] = R::sub(t1, t2);
    }
    x[4] = R::mul(o[4], K::tw(3, (T)2 * s));
}

// ---- inverse: DCT-III of X[0..7], in place ----------------------------------------------------
template <typename T, int NORM = kNormOrtho>
__device__ __forceinline__ void dct3_8(T (&X)[8]) {
    using R = Rn<T>;
    using K = Ducc<T>;
    const T s = dct_fct<T, NORM, true>();               // fct folded into the pre-step constants
    T c[8];
    c[0] = R::mul(X[0], NORM == kNormOrtho ? K::sqrt2(s) : s);
#pragma unroll
    for (int k = 1; k <= 3; ++k) {
        const int kc = 8 - k;
        const T twk = K::tw(k - 1, s), twc = K::tw(kc - 1, s);
        const T t1 = R::add(X[k], X[kc]), t2 = R::sub(X[k], X[kc]);
        c[k] = R::add(R::mul(twk, t2), R::mul(twc, t1));
        c[kc] = R::sub(R::mul(twk, t1), R::mul(twc, t2));
    }
    c[4] = R::mul(X[4], K::tw(3, (T)2 * s));            // * (2*tw3) * fct
    // radf4 (ido=1, l1=2): k=0 works on c0,c2,c4,c6; k=1 on c1,c3,c5,c7
    T h[8];
    {
        const T tr1 = R::add(c[6], c[2]); h[2] = R::sub(c[6], c[2]);
        const T tr2 = R::add(c[0], c[4]); h[1] = R::sub(c[0], c[4]);
        h[0] = R::add(tr2, tr1); h[3] = R::sub(tr2, tr1);
    }
    {
        const T tr1 = R::add(c[7], c[3]); h[6] = R::sub(c[7], c[3]);
        const T tr2 = R::add(c[1], c[5]); h[5] = R::sub(c[1], c[5]);
        h[4] = R::add(tr2, tr1); h[7] = R::sub(tr2, tr1);
    }
    // radf2 (ido=4, l1=1)
    const T o0 = R::add(h[0], h[4]), o7 = R::sub(h[0], h[4]);
    const T o4 = -h[7], o3 = h[3];
    const T tr2 = R::add(R::mul(K::wa0(), h[5]), R::mul(K::wa1(), h[6]));
    const T ti2 = R::sub(R::mul(K::wa0(), h[6]), R::mul(K::wa1(), h[5]));
    const T o1 = R::add(h[1], tr2), o5 = R::sub(h[1], tr2);
    const T o2 = R::add(ti2, h[2]), o6 = R::sub(ti2, h[2]);
    // post-step MPINPLACE(c[k], c[k+1]) for k = 1,3,5
    X[0] = o0;
    X[1] = R::sub(o1, o2); X[2] = R::add(o2, o1);
    X[3] = R::sub(o3, o4); X[4] = R::add(o4, o3);
    X[5] = R::sub(o5, o6); X[6] = R::add(o6, o5);
    X[7] = o7;
}

// ---- PatchQuant.quantize arithmetic: int32(rint(x / t)) (patchquant.py:59-60) ------------------
// x86 semantics for the float->int32 cast of out-of-range / NaN values (numpy uses cvttsd2si):
__device__ __forceinline__ int cast_i32_x86(double r) {
    return (r >= -2147483648.0 && r < 2147483648.0) ? __double2int_rz(r) : (int)0x80000000;
}

// Slow, always-exact path.
static __device__ __noinline__ int quantize_exact_f64(double x, double t) {
    return cast_i32_x86(rint(__ddiv_rn(x, t)));
}

// Fast path: y = x * fl(1/t) differs from fl(x/t) by < 2 ulp, so rint() can only disagree when y
// lies within 2^-17 of a half-integer.  y + 1.5*2^36 exposes y as a fixed-point number with 16
// fractional bits in the low mantissa word; the exact IEEE division is taken only when those 16
// bits read exactly one half (probability 2^-16 on generic data, always on true ties) or when
// |y| >= 2^15 - 1 (or NaN/Inf; for y in [32767.5, 32768) the rounding add below would wrap).
__device__ __forceinline__ int quantize_f64(double x, double t, double rt) {
    const double y = __dmul_rn(x, rt);
    const double sft = __dadd_rn(y, 103079215104.0);          // 1.5 * 2^36
    const int lo = __double2loint(sft);
    const bool ok = (fabs(y) < 32767.0) & ((lo & 0xFFFF) != 0x8000);
    if (__builtin_expect(ok, 1)) return (lo + 0x8000) >> 16;
    return quantize_exact_f64(x, t);
}

__device__ __forceinline__ int quantize_f32(float x, float t) {           // float32 numpy path
    const float r = rintf(__fdiv_rn(x, t));
    return (r >= -2147483648.0f && r < 2147483648.0f) ? __float2int_rz(r) : (int)0x80000000;
}

// ---- PatchQuant.dequantize arithmetic: int32(trunc(q * t)) (patchquant.py:77-78) ---------------
// int32 -> double without the conversion pipe: 2^52+2^31 bias trick (exact).
__device__ __forceinline__ double i32_to_f64(int q) {
    return __dsub_rn(__hiloint2double(0x43300000, (int)((unsigned)q ^ 0x80000000u)), 4503601774854144.0);
}
// trunc toward zero as a double: add +-2^52 in round-toward-zero mode, subtract it again (exact).
__device__ __forceinline__ double trunc_f64_small(double p) {             // |p| < 2^51
    const double m = __hiloint2double((__double2hiint(p) & 0x80000000) | 0x43300000, 0);   // copysign(2^52, p)
    return __dsub_rn(__dadd_rz(p, m), m);
}
__device__ __forceinline__ double dequantize_f64(int q, double t) {       // returns trunc(q*t) as double
    const double p = __dmul_rn(i32_to_f64(q), t);
    if (__builtin_expect(fabs(p) < 2147483648.0, 1)) {
        return trunc_f64_small(p);                     // x - x == +0 in RN, so never -0 (int32 has none)
    }
    return (double)(int)0x80000000;                    // x86 "integer indefinite", like numpy's cast
}

// ---- zig-zag table (ivclab/utils/shape.py:10-19): ZZ_ORDER[raster k] = scan position ----------
static __device__ __constant__ unsigned char ZZ_ORDER[64] = {
     0,  1,  5,  6, 14, 15, 27, 28,   2,  4,  7, 13, 16, 26, 29, 42,
     3,  8, 12, 17, 25, 30, 41, 43,   9, 11, 18, 24, 31, 40, 44, 53,
    10, 19, 23, 32, 39, 45, 52, 54,  20, 22, 33, 38, 46, 51, 55, 60,
    21, 34, 37, 47, 50, 56, 59, 61,  35, 36, 48, 49, 57, 58, 62, 63};

}  // namespace ivc
