// ivc_abi.cu -- the extern "C" surface declared in include/ivclab_b200.h.
// Argument validation lives here; kernels live in ivc_transform.cu / ivc_motion.cu.
#include "ivc_common.cuh"

#ifndef IVC_VERSION
#define IVC_VERSION "0.1.0"
#endif
#define IVC_STR2(x) #x
#define IVC_STR(x) IVC_STR2(x)

namespace {

thread_local cudaError_t g_last_cuda = cudaSuccess;

int cuda_fail(cudaError_t e) {
    g_last_cuda = e;
    return IVC_ERR_CUDA;
}

// Every entry point runs on `device` and leaves the calling thread's current device as it found it
// (a caller such as PyTorch keeps its own notion of the current device per thread).
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    cudaError_t err;
    explicit DeviceGuard(int device) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != device) {
            err = cudaSetDevice(device);
            switched = err == cudaSuccess;
        }
    }
    ~DeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
};
#define IVC_ENTER(device)               \
    DeviceGuard ivc_device_guard_(device); \
    if (ivc_device_guard_.err != cudaSuccess) return cuda_fail(ivc_device_guard_.err)

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline bool is_float(int dt) { return dt == IVC_F32 || dt == IVC_F64; }
inline int elem_size(int dt) {
    switch (dt) {
        case IVC_U8: return 1;
        case IVC_I32: case IVC_F32: return 4;
        case IVC_F64: case IVC_I64: return 8;
        default: return 0;
    }
}

}  // namespace

extern "C" {

int ivc_abi_version(void) { return IVC_ABI_VERSION; }

const char *ivc_build_info(void) {
    return "ivclab_b200 " IVC_VERSION " sm_100a nvcc " IVC_STR(__CUDACC_VER_MAJOR__) "." IVC_STR(__CUDACC_VER_MINOR__);
}

const char *ivc_error_string(int status) {
    switch (status) {
        case IVC_OK: return "ok";
        case IVC_ERR_ARG: return "invalid argument (null/misaligned pointer, negative size or bad flag)";
        case IVC_ERR_DTYPE: return "unsupported dtype combination";
        case IVC_ERR_SHAPE: return "unsupported shape (frame sides must be multiples of 8; C must broadcast against 3)";
        case IVC_ERR_CUDA: return "CUDA error (see ivc_last_cuda_error_string)";
        case IVC_ERR_WORKSPACE: return "workspace missing or too small";
        default: return "unknown status";
    }
}

int ivc_last_cuda_error(void) { return (int)g_last_cuda; }
const char *ivc_last_cuda_error_string(void) { return cudaGetErrorString(g_last_cuda); }

int ivc_dct8x8(int device, void *stream, int inverse, const void *x, int x_dtype, int64_t n0, int64_t n1, int64_t C,
               const int64_t strides[5], void *out, int out_dtype) {
    return ivc_dct8x8_norm(device, stream, inverse, IVC_NORM_ORTHO, x, x_dtype, n0, n1, C, strides, out, out_dtype);
}

int ivc_dct8x8_norm(int device, void *stream, int inverse, int norm, const void *x, int x_dtype, int64_t n0, int64_t n1,
                    int64_t C, const int64_t strides[5], void *out, int out_dtype) {
    if (norm != IVC_NORM_ORTHO && norm != IVC_NORM_BACKWARD && norm != IVC_NORM_FORWARD) return IVC_ERR_ARG;
    if (n0 < 0 || n1 < 0 || C < 0 || !strides) return IVC_ERR_ARG;
    if (n0 * n1 * C == 0) return IVC_OK;
    if (!x || !out) return IVC_ERR_ARG;
    if (x_dtype != IVC_U8 && x_dtype != IVC_I32 && x_dtype != IVC_F32 && x_dtype != IVC_F64) return IVC_ERR_DTYPE;
    const int want = (x_dtype == IVC_F32) ? IVC_F32 : IVC_F64;          // scipy keeps f32, promotes the rest to f64
    if (out_dtype != want) return IVC_ERR_DTYPE;
    IVC_ENTER(device);
    cudaError_t e = ivc::launch_dct(device, (cudaStream_t)stream, inverse != 0, x, x_dtype, n0, n1, C, strides, out,
                                    want == IVC_F32, norm);
    return e == cudaSuccess ? IVC_OK : cuda_fail(e);
}

static int quant_common(int device, void *stream, bool dequant, const void *x, int x_dtype, int64_t n0, int64_t n1,
                        int64_t C, const int64_t strides[5], const void *table, int table_dtype, int compute_dtype,
                        int32_t *out) {
    if (n0 < 0 || n1 < 0 || !strides) return IVC_ERR_ARG;
    if (C != 1 && C != 3) return IVC_ERR_SHAPE;                         // numpy broadcasting against [3,8,8]
    if (!table || !is_float(table_dtype) || !is_float(compute_dtype)) return IVC_ERR_DTYPE;
    if (n0 * n1 == 0) return IVC_OK;
    if (!x || !out) return IVC_ERR_ARG;
    if (compute_dtype == IVC_F32) {
        if (table_dtype != IVC_F32) return IVC_ERR_DTYPE;
        if (x_dtype != IVC_U8 && x_dtype != IVC_F32 && x_dtype != IVC_I32) return IVC_ERR_DTYPE;
    } else {
        if (elem_size(x_dtype) == 0) return IVC_ERR_DTYPE;
    }
    IVC_ENTER(device);
    cudaError_t e = ivc::launch_quant(device, (cudaStream_t)stream, dequant, x, x_dtype, n0, n1, C, strides, table,
                                      table_dtype, compute_dtype == IVC_F32, out);
    return e == cudaSuccess ? IVC_OK : cuda_fail(e);
}

int ivc_quantize(int device, void *stream, const void *x, int x_dtype, int64_t n0, int64_t n1, int64_t C,
                 const int64_t strides[5], const void *table, int table_dtype, int compute_dtype, int32_t *out) {
    return quant_common(device, stream, false, x, x_dtype, n0, n1, C, strides, table, table_dtype, compute_dtype, out);
}

int ivc_dequantize(int device, void *stream, const void *q, int q_dtype, int64_t n0, int64_t n1, int64_t C,
                   const int64_t strides[5], const void *table, int table_dtype, int compute_dtype, int32_t *out) {
    return quant_common(device, stream, true, q, q_dtype, n0, n1, C, strides, table, table_dtype, compute_dtype, out);
}

int ivc_zigzag(int device, void *stream, int inverse, const void *x, int elem_sz, int64_t nblocks, void *out) {
    if (nblocks < 0) return IVC_ERR_ARG;
    if (elem_sz != 1 && elem_sz != 2 && elem_sz != 4 && elem_sz != 8) return IVC_ERR_DTYPE;
    if (nblocks == 0) return IVC_OK;
    if (!x || !out || x == out) return IVC_ERR_ARG;
    IVC_ENTER(device);
    cudaError_t e = ivc::launch_zigzag(device, (cudaStream_t)stream, inverse != 0, x, elem_sz, nblocks, out);
    return e == cudaSuccess ? IVC_OK : cuda_fail(e);
}

int ivc_intra_forward(int device, void *stream, const void *img, int dtype, int64_t n_frames, int64_t H, int64_t W,
                      int64_t C, int64_t frame_stride, const void *table, int table_dtype, int32_t *out) {
    if (n_frames < 0 || H < 0 || W < 0 || frame_stride < 0) return IVC_ERR_ARG;
    if (dtype != IVC_F64 || !is_float(table_dtype)) return IVC_ERR_DTYPE;
    if ((H & 7) || (W & 7) || (C != 1 && C != 3)) return IVC_ERR_SHAPE;
    if (n_frames * H * W == 0) return IVC_OK;
    if (!img || !table || !out) return IVC_ERR_ARG;
    if (!aligned16(img) || !aligned16(out) || (frame_stride & 1)) return IVC_ERR_ARG;
    IVC_ENTER(device);
    cudaError_t e = ivc::launch_forward(device, (cudaStream_t)stream, img, n_frames, H, W, (int)C, frame_stride, table,
                                        table_dtype, out, nullptr, nullptr, 0, nullptr, false);
    return e == cudaSuccess ? IVC_OK : cuda_fail(e);
}

int ivc_intra_inverse(int device, void *stream, const int32_t *zz, int64_t n_frames, int64_t Hp, int64_t Wp, int64_t C,
                      const void *table, int table_dtype, void *out, int out_dtype) {
    if (n_frames < 0 || Hp < 0 || Wp < 0) return IVC_ERR_ARG;
    if (out_dtype != IVC_F64 || !is_float(table_dtype)) return IVC_ERR_DTYPE;
    if (C != 1 && C != 3) return IVC_ERR_SHAPE;
    if (n_frames * Hp * Wp == 0) return IVC_OK;
    if (!zz || !table || !out) return IVC_ERR_ARG;
    if (!aligned16(zz) || !aligned16(out)) return IVC_ERR_ARG;
    IVC_ENTER(device);
    cudaError_t e = ivc::launch_inverse(device, (cudaStream_t)stream, zz, n_frames, Hp, Wp, (int)C, table, table_dtype,
                                        out, C == 3 ? 0 : 1, nullptr, nullptr, nullptr, 0);
    return e == cudaSuccess ? IVC_OK : cuda_fail(e);
}

int ivc_intra_inverse_rgb(int device, void *stream, const int32_t *zz, int64_t n_frames, int64_t Hp, int64_t Wp,
                          const void *table, int table_dtype, void *rgb_out) {
    if (n_frames < 0 || Hp < 0 || Wp < 0) return IVC_ERR_ARG;
    if (!is_float(table_dtype)) return IVC_ERR_DTYPE;
    if (n_frames * Hp * Wp == 0) return IVC_OK;
    if (!zz || !table || !rgb_out) return IVC_ERR_ARG;
    if (!aligned16(zz) || !aligned16(rgb_out)) return IVC_ERR_ARG;
    IVC_ENTER(device);
    cudaError_t e = ivc::launch_inverse(device, (cudaStream_t)stream, zz, n_frames, Hp, Wp, 3, table, table_dtype, rgb_out, 3,
                                        nullptr, nullptr, nullptr, 0);
    return e == cudaSuccess ? IVC_OK : cuda_fail(e);
}

int64_t ivc_intra_inverse_sse_workspace_bytes(int64_t n_frames, int64_t Hp, int64_t Wp) {
    if (n_frames < 0 || Hp < 0 || Wp < 0) return -1;
    return (ivc::inverse_sse_tiles(n_frames, Hp, Wp) + 1) * (int64_t)sizeof(double);
}

int ivc_intra_inverse_sse(int device, void *stream, const int32_t *zz, int64_t n_frames, int64_t Hp, int64_t Wp,
                          const void *table, int table_dtype, void *out, const void *orig_rgb8,
                          int64_t orig_frame_stride_bytes, int mode, void *workspace, int64_t workspace_bytes,
                          double *sse_out) {
    if (n_frames < 0 || Hp < 0 || Wp < 0 || orig_frame_stride_bytes < 0) return IVC_ERR_ARG;
    if (mode != IVC_DIST_RGB && mode != IVC_DIST_YCBCR) return IVC_ERR_ARG;
    if (!is_float(table_dtype)) return IVC_ERR_DTYPE;
    if (Wp & 1) return IVC_ERR_SHAPE;                                       // 16-byte rows of packed RGB
    if (n_frames == 0) return IVC_OK;
    if (!sse_out) return IVC_ERR_ARG;
    if (Hp * Wp > 0) {
        if (!zz || !table || !orig_rgb8) return IVC_ERR_ARG;
        if (!aligned16(zz) || !aligned16(orig_rgb8) || (out && !aligned16(out)) || (orig_frame_stride_bytes & 15)) return IVC_ERR_ARG;
    }
    if (!workspace || workspace_bytes < ivc_intra_inverse_sse_workspace_bytes(n_frames, Hp, Wp)) return IVC_ERR_WORKSPACE;
    IVC_ENTER(device);
    cudaError_t e = ivc::launch_inverse_sse(device, (cudaStream_t)stream, zz, n_frames, Hp, Wp, table, table_dtype, out,
                                            orig_rgb8, orig_frame_stride_bytes, mode, (double *)workspace, sse_out);
    return e == cudaSuccess ? IVC_OK : cuda_fail(e);
}

int64_t ivc_me_workspace_bytes(int64_t n_frames, int64_t H, int64_t W) {
    if (n_frames < 0 || H < 0 || W < 0) return -1;
    return 256;                                // one device flag (kept 256-byte sized/aligned)
}

int64_t ivc_me_workspace_bytes_planes(int64_t n_frames, int64_t H, int64_t W) {
    if (n_frames < 0 || H < 0 || W < 0) return -1;
    return 256 + 2 * n_frames * H * W;         // the flag + a uint8 plane per reference and per current frame
}

// search ranges from which float64 frames are converted to uint8 planes first when the workspace has room for them
static const int kMePlanesMinRange = 8;

int ivc_me_full_search(int device, void *stream, const void *ref, const void *cur, int dtype, int64_t n_frames,
                       int64_t H, int64_t W, int64_t ref_frame_stride, int64_t cur_frame_stride, int search_range,
                       int mode, int64_t *mv_out, void *workspace, int64_t workspace_bytes) {
    if (n_frames < 0 || H < 0 || W < 0 || search_range < 0 || search_range > 64) return IVC_ERR_ARG;
    if (dtype != IVC_F32 && dtype != IVC_F64 && dtype != IVC_U8) return IVC_ERR_DTYPE;
    if ((H & 7) || (W & 7)) return IVC_ERR_SHAPE;                         // the reference raises on ragged frames
    if (mode != IVC_ME_AUTO && mode != IVC_ME_EXACT && mode != IVC_ME_INT) return IVC_ERR_ARG;
    if (n_frames * H * W == 0) return IVC_OK;
    if (!ref || !cur || !mv_out) return IVC_ERR_ARG;
    if (dtype == IVC_U8) {
        // uint8 PLANES holding the frames' values (float semantics, no wrap-around: that is ivc_me_full_search_intdtype).
        // Always integer-valued, so every mode is served by the packed-integer kernel, whose vectors are exact.
        IVC_ENTER(device);
        cudaError_t e8 = ivc::launch_me_int(device, (cudaStream_t)stream, ref, cur, IVC_U8, n_frames, H, W, ref_frame_stride,
                                            cur_frame_stride, search_range, mv_out, nullptr, 0);
        return e8 == cudaSuccess ? IVC_OK : cuda_fail(e8);
    }
    if (mode == IVC_ME_AUTO && (!workspace || workspace_bytes < ivc_me_workspace_bytes(n_frames, H, W)))
        return IVC_ERR_WORKSPACE;
    IVC_ENTER(device);
    cudaStream_t st = (cudaStream_t)stream;
    const bool f32 = dtype == IVC_F32;
    cudaError_t e;
    // A wide search reads every frame many times over (as a window with its halo and as blocks): with room in the workspace
    // float64 frames are converted -- and, in AUTO mode, validated -- ONCE, and the search runs on the uint8 planes.
    // Frames handed over as two views of one sequence (cur = ref + one frame) are converted once, not twice.
    const int64_t plane = H * W;
    const bool seq = cur_frame_stride == ref_frame_stride && ref_frame_stride == plane &&
                     (const char *)cur == (const char *)ref + plane * 8;
    if (dtype == IVC_F64 && mode != IVC_ME_EXACT && search_range >= kMePlanesMinRange && workspace &&
        workspace_bytes >= 256 + (seq ? n_frames + 1 : 2 * n_frames) * plane && aligned16(workspace) && aligned16(ref) && aligned16(cur) &&
        !(ref_frame_stride & 1) && !(cur_frame_stride & 1)) {
        int *flag = mode == IVC_ME_AUTO ? (int *)workspace : nullptr;
        unsigned char *planes = (unsigned char *)workspace + 256, *cur8 = planes + (seq ? plane : n_frames * plane);
        if (flag && (e = cudaMemsetAsync(flag, 0, sizeof(int), st)) != cudaSuccess) return cuda_fail(e);
        e = ivc::launch_f64_to_u8(device, st, ref, ref_frame_stride, seq ? n_frames + 1 : n_frames, plane, planes, flag);
        if (e == cudaSuccess && !seq) e = ivc::launch_f64_to_u8(device, st, cur, cur_frame_stride, n_frames, plane, cur8, flag);
        if (e == cudaSuccess)
            e = ivc::launch_me_int(device, st, planes, cur8, IVC_U8, n_frames, H, W, plane, plane, search_range, mv_out, nullptr, 0);
        if (e != cudaSuccess) return cuda_fail(e);
        if (mode == IVC_ME_AUTO) {             // the exact kernel overwrites the vectors if a frame was not integer-valued
            e = ivc::launch_me_exact(device, st, ref, cur, false, n_frames, H, W, ref_frame_stride, cur_frame_stride, search_range,
                                     mv_out, flag, 1);
            if (e != cudaSuccess) return cuda_fail(e);
        }
        return IVC_OK;
    }
    if (mode != IVC_ME_EXACT) {
        // integer kernel: converts the frames to packed u8 while staging; in AUTO mode it validates them
        // and raises the device flag instead of producing vectors from a non-integer frame
        e = ivc::launch_me_int(device, st, ref, cur, dtype, n_frames, H, W, ref_frame_stride, cur_frame_stride,
                               search_range, mv_out, (int *)workspace, mode == IVC_ME_AUTO ? 1 : 0);
        if (e != cudaSuccess) return cuda_fail(e);
    }
    if (mode != IVC_ME_INT) {
        // exact kernel: always in EXACT mode; in AUTO mode it exits at once unless the flag was raised
        e = ivc::launch_me_exact(device, st, ref, cur, f32, n_frames, H, W, ref_frame_stride, cur_frame_stride,
                                 search_range, mv_out, mode == IVC_ME_AUTO ? (int *)workspace : nullptr, 1);
        if (e != cudaSuccess) return cuda_fail(e);
    }
    return IVC_OK;
}

int ivc_me_full_search_intdtype(int device, void *stream, const void *ref, const void *cur, int dtype, int64_t n_frames,
                                int64_t H, int64_t W, int64_t ref_frame_stride, int64_t cur_frame_stride,
                                int search_range, int64_t *mv_out) {
    if (n_frames < 0 || H < 0 || W < 0 || search_range < 0 || search_range > 64) return IVC_ERR_ARG;
    if (dtype != IVC_U8 && dtype != IVC_I16 && dtype != IVC_I32) return IVC_ERR_DTYPE;
    if ((H & 7) || (W & 7)) return IVC_ERR_SHAPE;                         // the reference raises on ragged frames
    if (n_frames * H * W == 0) return IVC_OK;
    if (!ref || !cur || !mv_out || H > 2147483647LL || W > 2147483647LL) return IVC_ERR_ARG;
    IVC_ENTER(device);
    cudaError_t e = ivc::launch_me_wrap(device, (cudaStream_t)stream, ref, cur, dtype, n_frames, H, W, ref_frame_stride,
                                        cur_frame_stride, search_range, mv_out);
    return e == cudaSuccess ? IVC_OK : cuda_fail(e);
}

int ivc_mc_reconstruct(int device, void *stream, const void *ref, int elem_sz, int64_t n_frames, int64_t H, int64_t W,
                       int64_t C, const int64_t *mv, int search_range, void *out) {
    if (n_frames < 0 || H < 0 || W < 0 || C < 0 || search_range < 0) return IVC_ERR_ARG;
    if (elem_sz != 1 && elem_sz != 2 && elem_sz != 4 && elem_sz != 8) return IVC_ERR_DTYPE;
    if ((H & 7) || (W & 7)) return IVC_ERR_SHAPE;
    if (n_frames * H * W * C == 0) return IVC_OK;
    if (!ref || !mv || !out || ref == out) return IVC_ERR_ARG;
    IVC_ENTER(device);
    cudaError_t e = ivc::launch_mc(device, (cudaStream_t)stream, ref, elem_sz, n_frames, H, W, C, mv, search_range, out);
    return e == cudaSuccess ? IVC_OK : cuda_fail(e);
}

int ivc_pframe_forward(int device, void *stream, const void *cur, const void *ref, const int64_t *mv, int dtype,
                       int64_t n_frames, int64_t H, int64_t W, int search_range, const void *table, int table_dtype,
                       void *pred_out, int32_t *zz_out) {
    return ivc_pframe_forward_ch(device, stream, cur, ref, mv, dtype, n_frames, H, W, search_range, table, table_dtype,
                                 pred_out, 3, zz_out);
}

int ivc_pframe_forward_ch(int device, void *stream, const void *cur, const void *ref, const int64_t *mv, int dtype,
                          int64_t n_frames, int64_t H, int64_t W, int search_range, const void *table, int table_dtype,
                          void *pred_out, int out_channels, int32_t *zz_out) {
    if (n_frames < 0 || H < 0 || W < 0 || search_range < 0 || (out_channels != 2 && out_channels != 3)) return IVC_ERR_ARG;
    if (dtype != IVC_F64 || !is_float(table_dtype)) return IVC_ERR_DTYPE;
    if ((H & 7) || (W & 7)) return IVC_ERR_SHAPE;
    if (n_frames * H * W == 0) return IVC_OK;
    if (!cur || !ref || !mv || !table || !zz_out) return IVC_ERR_ARG;
    if (!aligned16(cur) || !aligned16(zz_out) || (pred_out && !aligned16(pred_out))) return IVC_ERR_ARG;
    IVC_ENTER(device);
    cudaError_t e = ivc::launch_forward(device, (cudaStream_t)stream, cur, n_frames, H, W, 1, H * W, table, table_dtype,
                                        zz_out, ref, mv, search_range, pred_out, true, out_channels);
    return e == cudaSuccess ? IVC_OK : cuda_fail(e);
}

static int search_forward_impl(int device, void *stream, const void *cur, const void *ref, int dtype, int64_t n_frames,
                               int64_t H, int64_t W, int search_range, int mode, const void *table, int table_dtype,
                               int out_channels, int64_t *mv_out, int32_t *zz_out, void *workspace, int64_t workspace_bytes,
                               int32_t *zr_counts, uint64_t *zr_masks);

int ivc_pframe_search_forward(int device, void *stream, const void *cur, const void *ref, int dtype, int64_t n_frames,
                              int64_t H, int64_t W, int search_range, int mode, const void *table, int table_dtype,
                              int out_channels, int64_t *mv_out, int32_t *zz_out, void *workspace, int64_t workspace_bytes) {
    return search_forward_impl(device, stream, cur, ref, dtype, n_frames, H, W, search_range, mode, table, table_dtype,
                               out_channels, mv_out, zz_out, workspace, workspace_bytes, nullptr, nullptr);
}

int ivc_pframe_search_forward_zr(int device, void *stream, const void *cur, const void *ref, int dtype, int64_t n_frames,
                                 int64_t H, int64_t W, int search_range, int mode, const void *table, int table_dtype,
                                 int out_channels, int64_t *mv_out, int32_t *zz_out, void *workspace, int64_t workspace_bytes,
                                 int32_t *counts_out, uint64_t *masks_out) {
    if (!counts_out || !masks_out) return IVC_ERR_ARG;
    return search_forward_impl(device, stream, cur, ref, dtype, n_frames, H, W, search_range, mode, table, table_dtype,
                               out_channels, mv_out, zz_out, workspace, workspace_bytes, counts_out, masks_out);
}

static int search_forward_impl(int device, void *stream, const void *cur, const void *ref, int dtype, int64_t n_frames,
                               int64_t H, int64_t W, int search_range, int mode, const void *table, int table_dtype,
                               int out_channels, int64_t *mv_out, int32_t *zz_out, void *workspace, int64_t workspace_bytes,
                               int32_t *zr_counts, uint64_t *zr_masks) {
    if (n_frames < 0 || H < 0 || W < 0 || search_range < 0 || search_range > 64) return IVC_ERR_ARG;
    if (out_channels != 2 && out_channels != 3) return IVC_ERR_ARG;
    if (mode != IVC_ME_AUTO && mode != IVC_ME_EXACT && mode != IVC_ME_INT) return IVC_ERR_ARG;
    if ((dtype != IVC_F64 && dtype != IVC_U8) || !is_float(table_dtype)) return IVC_ERR_DTYPE;
    // uint8 PLANES (the frames' values, as in ivc_me_full_search): always integer-valued, so there is nothing to fall
    // back to -- and nothing to fall back ON: only the fused kernel reads them
    if (dtype == IVC_U8 && !ivc::me_pf_fusable(dtype, search_range)) return IVC_ERR_DTYPE;
    if ((H & 7) || (W & 7)) return IVC_ERR_SHAPE;
    if (n_frames * H * W == 0) return IVC_OK;
    if (!cur || !ref || !table || !mv_out || !zz_out) return IVC_ERR_ARG;
    if ((dtype == IVC_F64 && !aligned16(cur)) || !aligned16(zz_out)) return IVC_ERR_ARG;
    if (dtype == IVC_F64 && mode == IVC_ME_AUTO && (!workspace || workspace_bytes < ivc_me_workspace_bytes(n_frames, H, W)))
        return IVC_ERR_WORKSPACE;
    IVC_ENTER(device);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
    const int64_t n_scan = n_frames * (H / 8) * (W / 8) * out_channels;           // scan blocks of zz_out
    if (dtype == IVC_U8) {
        e = ivc::launch_me_int(device, st, ref, cur, dtype, n_frames, H, W, H * W, H * W, search_range, mv_out, nullptr, 0,
                               table, table_dtype, zz_out, out_channels, zr_counts, zr_masks);
        return e == cudaSuccess ? IVC_OK : cuda_fail(e);
    }
    if (mode != IVC_ME_EXACT && ivc::me_pf_fusable(dtype, search_range)) {
        // one kernel searches and codes; in AUTO mode it leaves non-integer frames to the two stand-alone kernels, which
        // run only if it raised the flag (no host round trip)
        int *flag = mode == IVC_ME_AUTO ? (int *)workspace : nullptr;
        e = ivc::launch_me_int(device, st, ref, cur, dtype, n_frames, H, W, H * W, H * W, search_range, mv_out, flag,
                               mode == IVC_ME_AUTO ? 1 : 0, table, table_dtype, zz_out, out_channels, zr_counts, zr_masks);
        if (e != cudaSuccess) return cuda_fail(e);
        if (mode == IVC_ME_INT) return IVC_OK;
        e = ivc::launch_me_exact(device, st, ref, cur, false, n_frames, H, W, H * W, H * W, search_range, mv_out, flag, 1);
        if (e != cudaSuccess) return cuda_fail(e);
        bool zr_done = false;
        e = ivc::launch_forward(device, st, cur, n_frames, H, W, 1, H * W, table, table_dtype, zz_out, ref, mv_out,
                                search_range, nullptr, true, out_channels, flag, zr_counts, zr_masks, &zr_done);
        if (e == cudaSuccess && zr_counts && !zr_done)                           // the kernel variant that ran cannot emit them
            e = ivc::launch_zr_count(device, st, zz_out, n_scan, zr_counts, zr_masks, flag);
        return e == cudaSuccess ? IVC_OK : cuda_fail(e);
    }
    int rc = ivc_me_full_search(device, stream, ref, cur, dtype, n_frames, H, W, H * W, H * W, search_range, mode, mv_out,
                                workspace, workspace_bytes);
    if (rc != IVC_OK) return rc;
    if (!aligned16(cur)) return IVC_ERR_ARG;
    bool zr_done = false;
    e = ivc::launch_forward(device, st, cur, n_frames, H, W, 1, H * W, table, table_dtype, zz_out, ref, mv_out, search_range,
                            nullptr, true, out_channels, nullptr, zr_counts, zr_masks, &zr_done);
    if (e == cudaSuccess && zr_counts && !zr_done) e = ivc::launch_zr_count(device, st, zz_out, n_scan, zr_counts, zr_masks);
    return e == cudaSuccess ? IVC_OK : cuda_fail(e);
}

int ivc_pframe_step(int device, void *stream, const void *cur, const void *ref, int dtype, int64_t n_frames, int64_t H,
                    int64_t W, int search_range, const void *table, int table_dtype, int out_channels, int64_t *mv_out,
                    int32_t *zz_out, void *recon_out) {
    if (n_frames < 0 || n_frames > 65535 || H < 0 || W < 0 || search_range < 0 || search_range > 64) return IVC_ERR_ARG;
    if (out_channels != 2 && out_channels != 3) return IVC_ERR_ARG;
    if (dtype != IVC_F64 || !is_float(table_dtype)) return IVC_ERR_DTYPE;
    if ((H & 7) || (W & 7)) return IVC_ERR_SHAPE;
    if (n_frames * H * W == 0) return IVC_OK;
    if (!cur || !ref || !table || !mv_out || !zz_out || !recon_out || !aligned16(zz_out)) return IVC_ERR_ARG;
    IVC_ENTER(device);
    cudaError_t e = ivc::launch_pframe_step(device, (cudaStream_t)stream, ref, cur, n_frames, H, W, search_range, table,
                                            table_dtype, out_channels, mv_out, zz_out, (double *)recon_out);
    return e == cudaSuccess ? IVC_OK : cuda_fail(e);
}

int ivc_pframe_inverse(int device, void *stream, const int32_t *zz, int64_t Czz, const void *pred, const void *ref,
                       const int64_t *mv, int dtype, int64_t n_frames, int64_t H, int64_t W, int search_range,
                       const void *table, int table_dtype, void *recon_out) {
    if (n_frames < 0 || H < 0 || W < 0 || search_range < 0 || Czz < 1) return IVC_ERR_ARG;
    if (dtype != IVC_F64 || !is_float(table_dtype)) return IVC_ERR_DTYPE;
    if ((H & 7) || (W & 7)) return IVC_ERR_SHAPE;
    if (n_frames * H * W == 0) return IVC_OK;
    if (!zz || !table || !recon_out) return IVC_ERR_ARG;
    if (!pred && (!ref || !mv)) return IVC_ERR_ARG;
    if (!aligned16(zz) || !aligned16(recon_out) || (pred && !aligned16(pred))) return IVC_ERR_ARG;
    IVC_ENTER(device);
    cudaError_t e = ivc::launch_inverse(device, (cudaStream_t)stream, zz, n_frames, H / 8, W / 8, (int)Czz, table,
                                        table_dtype, recon_out, 2, pred, ref, mv, search_range);
    return e == cudaSuccess ? IVC_OK : cuda_fail(e);
}

int64_t ivc_sse_workspace_bytes(int64_t n_units, int64_t unit_elems) {
    if (n_units < 0 || unit_elems < 0) return -1;
    return (n_units * (int64_t)ivc::sse_chunks(n_units, unit_elems, 148) + 1) * (int64_t)sizeof(double);
}

int ivc_sum_squared_error(int device, void *stream, const void *a, int a_dtype, const void *b, int b_dtype,
                          int64_t n_units, int64_t unit_elems, int a_broadcast, void *workspace, int64_t workspace_bytes,
                          double *sse_out) {
    if (n_units < 0 || unit_elems < 0 || (a_broadcast != 1 && a_broadcast != 3 && a_broadcast != IVC_SSE_RGB8_AS_YCBCR))
        return IVC_ERR_ARG;
    if (a_broadcast == IVC_SSE_RGB8_AS_YCBCR && (a_dtype != IVC_U8 || b_dtype != IVC_F64)) return IVC_ERR_DTYPE;
    if (elem_size(a_dtype) == 0 || elem_size(b_dtype) == 0) return IVC_ERR_DTYPE;
    if (a_broadcast != 1 && unit_elems % 3) return IVC_ERR_SHAPE;
    if (n_units == 0) return IVC_OK;
    if (!sse_out) return IVC_ERR_ARG;
    if (unit_elems > 0 && (!a || !b)) return IVC_ERR_ARG;
    if (!workspace || workspace_bytes < ivc_sse_workspace_bytes(n_units, unit_elems)) return IVC_ERR_WORKSPACE;
    IVC_ENTER(device);
    cudaError_t e = ivc::launch_sse(device, (cudaStream_t)stream, a, a_dtype, b, b_dtype, n_units, unit_elems, a_broadcast,
                                    (double *)workspace, sse_out);
    return e == cudaSuccess ? IVC_OK : cuda_fail(e);
}

int ivc_zerorun_count(int device, void *stream, const int32_t *zz, int64_t nblocks, int32_t *counts_out) {
    if (nblocks < 0) return IVC_ERR_ARG;
    if (nblocks == 0) return IVC_OK;
    if (!zz || !counts_out || !aligned16(zz)) return IVC_ERR_ARG;
    IVC_ENTER(device);
    cudaError_t e = ivc::launch_zr_count(device, (cudaStream_t)stream, zz, nblocks, counts_out, nullptr);
    return e == cudaSuccess ? IVC_OK : cuda_fail(e);
}

int ivc_zerorun_count_masks(int device, void *stream, const int32_t *zz, int64_t nblocks, int32_t *counts_out,
                            uint64_t *masks_out) {
    if (nblocks < 0) return IVC_ERR_ARG;
    if (nblocks == 0) return IVC_OK;
    if (!zz || !counts_out || !masks_out || !aligned16(zz)) return IVC_ERR_ARG;
    IVC_ENTER(device);
    cudaError_t e = ivc::launch_zr_count(device, (cudaStream_t)stream, zz, nblocks, counts_out, masks_out);
    return e == cudaSuccess ? IVC_OK : cuda_fail(e);
}

int ivc_zerorun_symbol_histogram(int device, void *stream, const int32_t *zz, int64_t n_units, int64_t blocks_per_unit,
                                 int32_t end_of_block, int64_t lo, int64_t n_bins, uint32_t *counts_out,
                                 uint32_t *outside_out) {
    if (n_units < 0 || blocks_per_unit < 0 || n_bins < 1 || n_bins > 49152) return IVC_ERR_ARG;      // 192 KB of shared bins
    if (lo < -2147483647LL || lo > 2147483647LL) return IVC_ERR_ARG;
    if (n_units == 0) return IVC_OK;
    if (!counts_out || !outside_out || (blocks_per_unit > 0 && (!zz || !aligned16(zz)))) return IVC_ERR_ARG;
    IVC_ENTER(device);
    cudaError_t e = ivc::launch_zr_hist(device, (cudaStream_t)stream, zz, n_units, blocks_per_unit, end_of_block, lo, n_bins,
                                        counts_out, outside_out);
    return e == cudaSuccess ? IVC_OK : cuda_fail(e);
}

int ivc_zerorun_write_masks(int device, void *stream, const int32_t *zz, int64_t nblocks, int32_t end_of_block,
                            const int64_t *offsets, const uint64_t *masks, int32_t *symbols_out, int64_t total_symbols) {
    if (nblocks < 0) return IVC_ERR_ARG;
    if (nblocks == 0) return IVC_OK;
    if (!zz || !offsets || !masks || !symbols_out || !aligned16(zz)) return IVC_ERR_ARG;
    IVC_ENTER(device);
    cudaError_t e = ivc::launch_zr_write(device, (cudaStream_t)stream, zz, nblocks, end_of_block, offsets, masks, symbols_out, 4, total_symbols);
    return e == cudaSuccess ? IVC_OK : cuda_fail(e);
}

int ivc_zerorun_write(int device, void *stream, const int32_t *zz, int64_t nblocks, int32_t end_of_block,
                      const int64_t *offsets, int32_t *symbols_out) {
    if (nblocks < 0) return IVC_ERR_ARG;
    if (nblocks == 0) return IVC_OK;
    if (!zz || !offsets || !symbols_out || !aligned16(zz)) return IVC_ERR_ARG;
    IVC_ENTER(device);
    cudaError_t e = ivc::launch_zr_write(device, (cudaStream_t)stream, zz, nblocks, end_of_block, offsets, nullptr, symbols_out, 4, -1);
    return e == cudaSuccess ? IVC_OK : cuda_fail(e);
}

int ivc_zerorun_write_masks_i16(int device, void *stream, const int32_t *zz, int64_t nblocks, int32_t end_of_block,
                                const int64_t *offsets, const uint64_t *masks, int16_t *symbols_out, int64_t total_symbols) {
    if (nblocks < 0 || end_of_block > 32767 || end_of_block < -32768) return IVC_ERR_ARG;
    if (nblocks == 0) return IVC_OK;
    if (!zz || !offsets || !masks || !symbols_out || !aligned16(zz)) return IVC_ERR_ARG;
    IVC_ENTER(device);
    cudaError_t e = ivc::launch_zr_write(device, (cudaStream_t)stream, zz, nblocks, end_of_block, offsets, masks, symbols_out, 2, total_symbols);
    return e == cudaSuccess ? IVC_OK : cuda_fail(e);
}

int64_t ivc_zerorun_offsets_workspace_bytes(int64_t nblocks) {
    if (nblocks < 0) return -1;
    return ivc::zr_offsets_workspace_bytes(nblocks);
}

int ivc_zerorun_offsets(int device, void *stream, const int32_t *counts, int64_t nblocks, int64_t *offsets_out,
                        void *workspace, int64_t workspace_bytes, int64_t *total_mapped_out, int64_t *total_dev_out) {
    if (nblocks < 0) return IVC_ERR_ARG;
    if (nblocks == 0) return IVC_OK;
    if (!counts || !offsets_out || !aligned16(counts)) return IVC_ERR_ARG;
    if (!workspace || workspace_bytes < ivc::zr_offsets_workspace_bytes(nblocks)) return IVC_ERR_WORKSPACE;
    IVC_ENTER(device);
    cudaError_t e = ivc::launch_zr_offsets((cudaStream_t)stream, counts, nblocks, offsets_out, workspace, total_mapped_out,
                                           total_dev_out);
    return e == cudaSuccess ? IVC_OK : cuda_fail(e);
}

int ivc_post_words_to_host(int device, void *stream, const int64_t *src, int64_t *dst_mapped, int n) {
    if (n < 0 || n > 32) return IVC_ERR_ARG;
    if (n == 0) return IVC_OK;
    if (!src || !dst_mapped) return IVC_ERR_ARG;
    IVC_ENTER(device);
    cudaError_t e = ivc::launch_post_words((cudaStream_t)stream, src, dst_mapped, n);
    return e == cudaSuccess ? IVC_OK : cuda_fail(e);
}

int ivc_zerorun_decode_mark(int device, void *stream, const int32_t *symbols, int64_t n_symbols, int32_t end_of_block,
                            int32_t *is_eob_out) {
    if (n_symbols < 0) return IVC_ERR_ARG;
    if (n_symbols == 0) return IVC_OK;
    if (!symbols || !is_eob_out) return IVC_ERR_ARG;
    IVC_ENTER(device);
    cudaError_t e = ivc::launch_zrd_mark(device, (cudaStream_t)stream, symbols, n_symbols, end_of_block, is_eob_out);
    return e == cudaSuccess ? IVC_OK : cuda_fail(e);
}

int ivc_zerorun_decode_ends(int device, void *stream, const int32_t *is_eob, const int64_t *rank, int64_t n_symbols,
                            int64_t n_blocks, int64_t *ends_out) {
    if (n_symbols < 0 || n_blocks < 0) return IVC_ERR_ARG;
    if (n_symbols == 0 || n_blocks == 0) return IVC_OK;
    if (!is_eob || !rank || !ends_out) return IVC_ERR_ARG;
    IVC_ENTER(device);
    cudaError_t e = ivc::launch_zrd_ends(device, (cudaStream_t)stream, is_eob, rank, n_symbols, n_blocks, ends_out);
    return e == cudaSuccess ? IVC_OK : cuda_fail(e);
}

int ivc_zerorun_decode_write(int device, void *stream, const int32_t *symbols, const int64_t *ends, int64_t n_blocks,
                             int32_t *blocks_out, int32_t *err_out) {
    if (n_blocks < 0) return IVC_ERR_ARG;
    if (n_blocks == 0) return IVC_OK;
    if (!symbols || !ends || !blocks_out || !err_out) return IVC_ERR_ARG;
    IVC_ENTER(device);
    cudaError_t e = ivc::launch_zrd_write(device, (cudaStream_t)stream, symbols, ends, n_blocks, blocks_out, (int *)err_out);
    return e == cudaSuccess ? IVC_OK : cuda_fail(e);
}

int ivc_symbol_minmax(int device, void *stream, const void *x, int dtype, int64_t n, int64_t *minmax_out) {
    if (n < 0) return IVC_ERR_ARG;
    if (dtype != IVC_U8 && dtype != IVC_I32 && dtype != IVC_I64) return IVC_ERR_DTYPE;
    if (!minmax_out || (n > 0 && !x)) return IVC_ERR_ARG;
    IVC_ENTER(device);
    cudaError_t e = ivc::launch_minmax(device, (cudaStream_t)stream, x, dtype, n, minmax_out);
    return e == cudaSuccess ? IVC_OK : cuda_fail(e);
}

int ivc_symbol_histogram(int device, void *stream, const void *x, int dtype, int64_t n, int64_t lo, int64_t n_bins,
                         int64_t hot, uint64_t *counts_out) {
    if (n < 0 || n_bins < 0 || n_bins > 2147483647LL) return IVC_ERR_ARG;
    if (dtype != IVC_U8 && dtype != IVC_I32 && dtype != IVC_I64) return IVC_ERR_DTYPE;
    if (n_bins == 0) return IVC_OK;
    if (!counts_out || (n > 0 && !x)) return IVC_ERR_ARG;
    IVC_ENTER(device);
    cudaError_t e = ivc::launch_hist(device, (cudaStream_t)stream, x, dtype, n, lo, n_bins, hot, counts_out);
    return e == cudaSuccess ? IVC_OK : cuda_fail(e);
}

int ivc_rgb2ycbcr(int device, void *stream, const void *rgb, int dtype, int64_t npixels, void *ycbcr_out) {
    if (npixels < 0) return IVC_ERR_ARG;
    if (dtype != IVC_U8 && dtype != IVC_I32 && dtype != IVC_F32 && dtype != IVC_F64) return IVC_ERR_DTYPE;
    if (npixels == 0) return IVC_OK;
    if (!rgb || !ycbcr_out) return IVC_ERR_ARG;
    IVC_ENTER(device);
    cudaError_t e = ivc::launch_color(device, (cudaStream_t)stream, false, rgb, dtype, npixels, (double *)ycbcr_out);
    return e == cudaSuccess ? IVC_OK : cuda_fail(e);
}

int ivc_rgb8_to_luma8(int device, void *stream, const void *rgb, int64_t npixels, void *luma_out, void *luma_f64_out) {
    if (npixels < 0) return IVC_ERR_ARG;
    if (npixels == 0) return IVC_OK;
    if (!rgb || !luma_out) return IVC_ERR_ARG;
    IVC_ENTER(device);
    cudaError_t e = ivc::launch_rgb8_luma8(device, (cudaStream_t)stream, rgb, npixels, luma_out, luma_f64_out);
    return e == cudaSuccess ? IVC_OK : cuda_fail(e);
}

int ivc_ycbcr2rgb(int device, void *stream, const void *ycbcr, int64_t npixels, void *rgb_out) {
    if (npixels < 0) return IVC_ERR_ARG;
    if (npixels == 0) return IVC_OK;
    if (!ycbcr || !rgb_out) return IVC_ERR_ARG;
    IVC_ENTER(device);
    cudaError_t e = ivc::launch_color(device, (cudaStream_t)stream, true, ycbcr, IVC_F64, npixels, (double *)rgb_out);
    return e == cudaSuccess ? IVC_OK : cuda_fail(e);
}

int ivc_intra_forward_rgb8(int device, void *stream, const void *rgb, int64_t n_frames, int64_t H, int64_t W,
                           int64_t frame_stride_bytes, const void *table, int table_dtype, int32_t *out) {
    if (n_frames < 0 || H < 0 || W < 0 || frame_stride_bytes < 0) return IVC_ERR_ARG;
    if (!is_float(table_dtype)) return IVC_ERR_DTYPE;
    if ((H & 7) || (W & 15)) return IVC_ERR_SHAPE;          // 16-byte rows for the bulk copies: W % 16 == 0
    if (n_frames * H * W == 0) return IVC_OK;
    if (!rgb || !table || !out) return IVC_ERR_ARG;
    if (!aligned16(rgb) || !aligned16(out) || (frame_stride_bytes & 15)) return IVC_ERR_ARG;
    IVC_ENTER(device);
    cudaError_t e = ivc::launch_forward_rgb8(device, (cudaStream_t)stream, rgb, n_frames, H, W, frame_stride_bytes, table,
                                             table_dtype, out);
    return e == cudaSuccess ? IVC_OK : cuda_fail(e);
}

int ivc_intra_forward_rgb8_zr(int device, void *stream, const void *rgb, int64_t n_frames, int64_t H, int64_t W,
                              int64_t frame_stride_bytes, const void *table, int table_dtype, int32_t *out,
                              int32_t *counts_out, uint64_t *masks_out) {
    if (n_frames < 0 || H < 0 || W < 0 || frame_stride_bytes < 0) return IVC_ERR_ARG;
    if (!is_float(table_dtype)) return IVC_ERR_DTYPE;
    if ((H & 7) || (W & 15)) return IVC_ERR_SHAPE;
    if (n_frames * H * W == 0) return IVC_OK;
    if (!rgb || !table || !out || !counts_out || !masks_out) return IVC_ERR_ARG;
    if (!aligned16(rgb) || !aligned16(out) || (frame_stride_bytes & 15)) return IVC_ERR_ARG;
    IVC_ENTER(device);
    cudaError_t e = ivc::launch_forward_rgb8(device, (cudaStream_t)stream, rgb, n_frames, H, W, frame_stride_bytes, table,
                                             table_dtype, out, counts_out, masks_out);
    return e == cudaSuccess ? IVC_OK : cuda_fail(e);
}

int ivc_intra_forward_rgb8_multi(int device, void *stream, const void *rgb, int64_t n_frames, int64_t H, int64_t W,
                                 int64_t frame_stride_bytes, const void *tables, int table_dtype, int n_tables, int32_t *out,
                                 int32_t *counts_out, uint64_t *masks_out) {
    if (n_frames < 0 || H < 0 || W < 0 || frame_stride_bytes < 0) return IVC_ERR_ARG;
    if (n_tables < 1 || n_tables > IVC_MAX_FORWARD_TABLES) return IVC_ERR_ARG;
    if (!is_float(table_dtype)) return IVC_ERR_DTYPE;
    if ((H & 7) || (W & 15)) return IVC_ERR_SHAPE;
    if (n_frames * H * W == 0) return IVC_OK;
    if (!rgb || !tables || !out || (!counts_out != !masks_out)) return IVC_ERR_ARG;
    if (!aligned16(rgb) || !aligned16(out) || (frame_stride_bytes & 15)) return IVC_ERR_ARG;
    IVC_ENTER(device);
    cudaError_t e = ivc::launch_forward_rgb8(device, (cudaStream_t)stream, rgb, n_frames, H, W, frame_stride_bytes, tables,
                                             table_dtype, out, counts_out, masks_out, n_tables);
    return e == cudaSuccess ? IVC_OK : cuda_fail(e);
}

}  // extern "C"
