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
