// ivc_color.cuh -- per-pixel BT.601 arithmetic shared by the standalone colour kernels and the fused
// RGB front end of K1 (see ivc_color.cu for the derivation of the rounding order).
#pragma once
#include <cuda_runtime.h>

namespace ivc {

// image @ M.T + offset (color.py:27-36): one FMA chain per output, then a rounded add of the offset
__device__ __forceinline__ void rgb2ycbcr_px(double r, double g, double b, double &y, double &cb, double &cr) {
    y = __dadd_rn(__fma_rn(b, 0.114, __fma_rn(g, 0.587, __dmul_rn(r, 0.299))), 0.0);
    cb = __dadd_rn(__fma_rn(b, 0.5, __fma_rn(g, -0.331264, __dmul_rn(r, -0.168736))), 128.0);
    cr = __dadd_rn(__fma_rn(b, -0.081312, __fma_rn(g, -0.418688, __dmul_rn(r, 0.5))), 128.0);
}

// color.py:51-62: Cb -= 128, Cr -= 128; R = Y + 1.402 Cr; G = Y - 0.344136 Cb - 0.714136 Cr; B = Y + 1.772 Cb; clip
__device__ __forceinline__ double clip255(double v) { return v < 0.0 ? 0.0 : (v > 255.0 ? 255.0 : v); }   // NaN passes through like np.clip
__device__ __forceinline__ void ycbcr2rgb_px(double y, double cb0, double cr0, double &r, double &g, double &b) {
    const double cb = __dsub_rn(cb0, 128.0), cr = __dsub_rn(cr0, 128.0);
    r = clip255(__dadd_rn(y, __dmul_rn(1.402, cr)));
    g = clip255(__dsub_rn(__dsub_rn(y, __dmul_rn(0.344136, cb)), __dmul_rn(0.714136, cr)));
    b = clip255(__dadd_rn(y, __dmul_rn(1.772, cb)));
}

}  // namespace ivc
