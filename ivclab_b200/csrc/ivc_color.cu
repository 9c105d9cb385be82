// ivc_color.cu -- BT.601 colour transforms (ivclab/signal/color.py:15-63).
//
// rgb2ycbcr is `image @ M.T + offset` in numpy: the image is cast to float64 and multiplied by the 3x3
// matrix through BLAS, whose micro-kernel accumulates each output as ONE FMA CHAIN over the three
// input channels starting from zero -- fma(b, m2, fma(g, m1, r*m0)) -- followed by a separately rounded
// `+ offset` (probed against numpy/OpenBLAS 0.3.30: bit-identical on every sample; a mul/add chain
// with separate roundings matches only 85 %).  ycbcr2rgb is elementwise numpy: every operation
// individually rounded, then clip to [0, 255].
#include "ivc_color.cuh"
#include "ivc_common.cuh"

namespace ivc {

struct ColorArgs {
    const void *in;
    int in_dtype;
    double *out;
    int64_t npix;
};

__global__ void __launch_bounds__(256) k_rgb2ycbcr(const ColorArgs a) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.npix; i += (int64_t)gridDim.x * blockDim.x) {
        double r, g, b;
        switch (a.in_dtype) {
            case IVC_U8: { const unsigned char *p = (const unsigned char *)a.in + 3 * i; r = p[0]; g = p[1]; b = p[2]; break; }
            case IVC_F32: { const float *p = (const float *)a.in + 3 * i; r = p[0]; g = p[1]; b = p[2]; break; }
            case IVC_I32: { const int *p = (const int *)a.in + 3 * i; r = p[0]; g = p[1]; b = p[2]; break; }
            default: { const double *p = (const double *)a.in + 3 * i; r = p[0]; g = p[1]; b = p[2]; break; }
        }
        double y, cb, cr;
        rgb2ycbcr_px(r, g, b, y, cb, cr);
        double *o = a.out + 3 * i;
        o[0] = y; o[1] = cb; o[2] = cr;
    }
}

__global__ void __launch_bounds__(256) k_ycbcr2rgb(const ColorArgs a) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.npix; i += (int64_t)gridDim.x * blockDim.x) {
        const double *p = (const double *)a.in + 3 * i;
        double r, g, b;
        ycbcr2rgb_px(p[0], p[1], p[2], r, g, b);
        double *o = a.out + 3 * i;
        o[0] = r; o[1] = g; o[2] = b;
    }
}

// uint8 RGB -> uint8 luma plane: clip(rint(Y), 0, 255) with Y exactly as rgb2ycbcr computes it (rint = round half to
// even = np.round).  What the video codecs code is Y = rgb2ycbcr(frame)[..., 0] (videocodec.py:38); deriving the
// plane on the device means a host-fed pipeline uploads 3 bytes per pixel instead of 4.  Four pixels per thread:
// three 32-bit loads, one 32-bit store; out64 (optional) receives the same plane as float64, the dtype the transform
// kernels read (saves the separate conversion pass).
__global__ void __launch_bounds__(256) k_rgb8_luma8(const unsigned char *__restrict__ rgb, unsigned char *__restrict__ out,
                                                    double *__restrict__ out64, int64_t npix, int vec) {
    const auto luma = [](unsigned r, unsigned g, unsigned b) {
        double y, cb, cr;
        rgb2ycbcr_px((double)r, (double)g, (double)b, y, cb, cr);
        const int v = __double2int_rn(y);
        return (unsigned)min(max(v, 0), 255);
    };
    const int64_t ngrp = vec ? npix / 4 : 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ngrp; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned *p = reinterpret_cast<const unsigned *>(rgb) + 3 * i;
        const unsigned w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);       // r0 g0 b0 r1 | g1 b1 r2 g2 | b2 r3 g3 b3
        const unsigned y0 = luma(w0 & 255u, (w0 >> 8) & 255u, (w0 >> 16) & 255u);
        const unsigned y1 = luma(w0 >> 24, w1 & 255u, (w1 >> 8) & 255u);
        const unsigned y2 = luma((w1 >> 16) & 255u, w1 >> 24, w2 & 255u);
        const unsigned y3 = luma((w2 >> 8) & 255u, (w2 >> 16) & 255u, w2 >> 24);
        reinterpret_cast<unsigned *>(out)[i] = y0 | (y1 << 8) | (y2 << 16) | (y3 << 24);
        if (out64) {
            double2 *o = reinterpret_cast<double2 *>(out64 + 4 * i);
            o[0] = make_double2((double)y0, (double)y1);
            o[1] = make_double2((double)y2, (double)y3);
        }
    }
    for (int64_t i = 4 * ngrp + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned y = luma(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]);
        out[i] = (unsigned char)y;
        if (out64) out64[i] = (double)y;
    }
}

cudaError_t launch_rgb8_luma8(int device, cudaStream_t st, const void *rgb, int64_t npix, void *out, void *out64) {
    if (npix == 0) return cudaSuccess;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    int64_t grid = (npix / 4 + 255) / 256 + 1;
    if (grid > (int64_t)sms * 16) grid = (int64_t)sms * 16;
    const int vec = (((uintptr_t)rgb | (uintptr_t)out) & 3) == 0 && ((uintptr_t)out64 & 15) == 0;
    k_rgb8_luma8<<<(unsigned)grid, 256, 0, st>>>((const unsigned char *)rgb, (unsigned char *)out, (double *)out64, npix, vec);
    return cudaGetLastError();
}

cudaError_t launch_color(int device, cudaStream_t st, bool to_rgb, const void *in, int in_dtype, int64_t npix, double *out) {
    if (npix == 0) return cudaSuccess;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    int64_t grid = (npix + 256 * 4 - 1) / (256 * 4);
    if (grid > (int64_t)sms * 16) grid = (int64_t)sms * 16;
    ColorArgs a{in, in_dtype, out, npix};
    if (to_rgb) k_ycbcr2rgb<<<(unsigned)grid, 256, 0, st>>>(a);
    else k_rgb2ycbcr<<<(unsigned)grid, 256, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace ivc
