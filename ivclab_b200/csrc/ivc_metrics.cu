// ivc_metrics.cu -- sum of squared differences per unit (frame), the kernel behind calc_mse / calc_psnr
// (ivclab/utils/metrics.py:3-40).
//
// Deterministic two-stage reduction (no floating-point atomics): stage 1 gives every (unit, chunk) pair
// one CTA that accumulates (double)a - (double)b squared with a fixed thread stride and a fixed
// shuffle/shared tree; stage 2 lets one warp add the chunk partials of a unit in a fixed order.  The
// result differs from numpy's pairwise mean only by summation order (relative ~1e-15); the parity
// tests use a 1e-12 relative tolerance, far inside the north star's 0.01 dB PSNR bar.
#include "ivc_color.cuh"
#include "ivc_common.cuh"

namespace ivc {

constexpr int kSseThreads = 256;

struct SseArgs {
    const void *a, *b;
    int a_dtype, b_dtype;
    int64_t n_units, unit_elems;     // elements of b per unit
    int a_div;                       // 1, or 3 when a gray `a` is compared with an RGB `b` (metrics.py:16-19)
    int chunks;
    double *partial;                 // [n_units][chunks]
    double *out;                     // [n_units]
};

__device__ __forceinline__ double ld_f64(const void *p, int dtype, int64_t i) {
    switch (dtype) {
        case IVC_U8: return (double)((const unsigned char *)p)[i];
        case IVC_I32: return (double)((const int *)p)[i];
        case IVC_F32: return (double)((const float *)p)[i];
        case IVC_I64: return (double)((const long long *)p)[i];
        default: return ((const double *)p)[i];
    }
}

__global__ void __launch_bounds__(kSseThreads) k_sse_stage1(const SseArgs s) {
    __shared__ double sh[kSseThreads / 32];
    const int64_t unit = blockIdx.x / s.chunks;
    const int chunk = blockIdx.x % s.chunks;
    const int64_t per = (s.unit_elems + s.chunks - 1) / s.chunks;
    const int64_t lo = chunk * per, hi = min(s.unit_elems, lo + per);
    const int64_t a_unit = s.unit_elems / (s.a_div == 3 ? 3 : 1);
    double acc = 0.0;
    if (s.a_dtype == IVC_F64 && s.b_dtype == IVC_F64 && s.a_div == 1) {
        const double *pa = (const double *)s.a + unit * s.unit_elems, *pb = (const double *)s.b + unit * s.unit_elems;
        for (int64_t i = lo + threadIdx.x; i < hi; i += kSseThreads) {
            const double d = pa[i] - pb[i];
            acc += d * d;
        }
    } else if (s.a_div == IVC_SSE_RGB8_AS_YCBCR) {
        // a = uint8 RGB, compared in YCbCr space (rgb2ycbcr fused in, color.py:27-36): element i is channel i % 3 of
        // pixel i / 3, visited in exactly the order of the float64 path above, so both give the same bits
        const unsigned char *pa = (const unsigned char *)s.a + unit * s.unit_elems;
        const double *pb = (const double *)s.b + unit * s.unit_elems;
        int64_t i = lo + threadIdx.x;
        int64_t px = i / 3;
        int c = (int)(i - 3 * px);
        for (; i < hi; i += kSseThreads) {
            double y, cb, cr;
            rgb2ycbcr_px((double)pa[3 * px], (double)pa[3 * px + 1], (double)pa[3 * px + 2], y, cb, cr);
            const double d = (c == 0 ? y : (c == 1 ? cb : cr)) - pb[i];
            acc += d * d;
            px += kSseThreads / 3;                                   // 256 = 3 * 85 + 1
            if (++c == 3) { c = 0; ++px; }
        }
    } else {
        for (int64_t i = lo + threadIdx.x; i < hi; i += kSseThreads) {
            const double d = ld_f64(s.a, s.a_dtype, unit * a_unit + i / s.a_div) - ld_f64(s.b, s.b_dtype, unit * s.unit_elems + i);
            acc += d * d;
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < kSseThreads / 32; ++w) t += sh[w];
        s.partial[unit * s.chunks + chunk] = t;
    }
}

__global__ void __launch_bounds__(32) k_sse_stage2(const SseArgs s) {
    const int64_t unit = blockIdx.x;
    double acc = 0.0;
    for (int c = threadIdx.x; c < s.chunks; c += 32) acc += s.partial[unit * s.chunks + c];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (threadIdx.x == 0) s.out[unit] = acc;
}

int sse_chunks(int64_t n_units, int64_t unit_elems, int sms) {
    int64_t c = (unit_elems + 16383) / 16384;                 // about 16K elements per CTA
    (void)n_units; (void)sms;
    if (c < 1) c = 1;
    if (c > 4096) c = 4096;
    return (int)c;
}

cudaError_t launch_sse(int device, cudaStream_t st, const void *a, int a_dtype, const void *b, int b_dtype,
                       int64_t n_units, int64_t unit_elems, int a_div, double *partial, double *out) {
    if (n_units == 0) return cudaSuccess;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    SseArgs s;
    s.a = a; s.b = b; s.a_dtype = a_dtype; s.b_dtype = b_dtype; s.n_units = n_units; s.unit_elems = unit_elems;
    s.a_div = a_div; s.chunks = sse_chunks(n_units, unit_elems, sms); s.partial = partial; s.out = out;
    const int64_t ctas = n_units * s.chunks;
    if (ctas > 2147483647LL) return cudaErrorInvalidValue;
    k_sse_stage1<<<(unsigned)ctas, kSseThreads, 0, st>>>(s);
    k_sse_stage2<<<(unsigned)n_units, 32, 0, st>>>(s);
    return cudaGetLastError();
}

// ---- symbol statistics: min/max and unit-width histogram (stats_marg, ivclab/entropy/entropy.py:6-29, as
//      IntraCodec.train_huffman_from_image uses it: intracodec.py:160-166) ---------------------------
// np.histogram(x, bins=np.arange(lo, hi)) has hi-lo-1 unit bins [lo+k, lo+k+1), the last one closed:
// x == hi-1 lands in bin hi-lo-2.  Zero-run symbol streams are dominated by two values (the zero marker and
// EOB): those are counted with warp ballots (one shared atomic per warp and value), the rest with
// shared-memory atomics on a per-CTA histogram, flushed with one global atomic per non-empty bin.
constexpr int kHistThreads = 256;

__device__ __forceinline__ long long ld_i64(const void *p, int dtype, int64_t i) {
    switch (dtype) {
        case IVC_U8: return ((const unsigned char *)p)[i];
        case IVC_I32: return ((const int *)p)[i];
        default: return ((const long long *)p)[i];
    }
}

__global__ void __launch_bounds__(kHistThreads) k_hist(const void *x, int dtype, int64_t n, long long lo, int nbins,
                                                       long long hot, unsigned long long *counts, int use_smem) {
    extern __shared__ unsigned sh[];
    if (use_smem) {
        for (int i = threadIdx.x; i < nbins; i += kHistThreads) sh[i] = 0;
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    auto add = [&](long long v, bool in) {                                             // whole warp calls this together
        long long b = v - lo;
        if (b == nbins) b = nbins - 1;                                                  // closed last bin
        bool todo = in && b >= 0 && b < nbins;
        if (use_smem) {
            // the two values that dominate a zero-run stream (the zero marker and EOB) are counted per warp
            const long long hots[2] = {0, hot};
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const bool is = todo && v == hots[h] && (h == 0 || hot != 0);
                const unsigned mask = __ballot_sync(0xffffffffu, is);
                if (mask && lane == __ffs(mask) - 1) atomicAdd(&sh[(int)b], (unsigned)__popc(mask));
                todo = todo && !is;
            }
            if (todo) atomicAdd(&sh[(int)b], 1u);
        } else if (todo) {
            atomicAdd(&counts[b], 1ull);
        }
    };
    const int64_t stride = (int64_t)gridDim.x * kHistThreads;
    if (dtype == IVC_I32 && ((uintptr_t)x & 15) == 0) {                                  // four symbols per thread and round
        const int64_t n4 = n >> 2;
        for (int64_t i0 = (int64_t)blockIdx.x * kHistThreads; i0 < n4; i0 += stride) {   // warp-uniform trip count
            const int64_t i = i0 + threadIdx.x;
            const bool in = i < n4;
            const int4 v = in ? __ldg(reinterpret_cast<const int4 *>(x) + i) : make_int4(0, 0, 0, 0);
            add(v.x, in); add(v.y, in); add(v.z, in); add(v.w, in);
        }
        if (blockIdx.x == 0) {                                                           // the last n % 4 symbols
            const int64_t i = 4 * n4 + threadIdx.x;
            const bool in = threadIdx.x < 32 && i < n;
            if (threadIdx.x < 32) add(in ? ((const int *)x)[i] : 0, in);
        }
    } else {
        for (int64_t i0 = (int64_t)blockIdx.x * kHistThreads; i0 < n; i0 += stride) {
            const int64_t i = i0 + threadIdx.x;
            const bool in = i < n;
            add(in ? ld_i64(x, dtype, i) : 0, in);
        }
    }
    if (use_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < nbins; i += kHistThreads)
            if (sh[i]) atomicAdd(&counts[i], (unsigned long long)sh[i]);
    }
}

__global__ void k_minmax_init(long long *out) {
    out[0] = 0x7fffffffffffffffLL;
    out[1] = -0x7fffffffffffffffLL - 1;
}

// out[0] = min, out[1] = max (int64; k_minmax_init sets them to INT64_MAX / INT64_MIN first)
__global__ void __launch_bounds__(kHistThreads) k_minmax(const void *x, int dtype, int64_t n, long long *out) {
    long long mn = 0x7fffffffffffffffLL, mx = -0x7fffffffffffffffLL - 1;
    for (int64_t i = (int64_t)blockIdx.x * kHistThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kHistThreads) {
        const long long v = ld_i64(x, dtype, i);
        mn = min(mn, v);
        mx = max(mx, v);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, off));
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(out, mn);
        atomicMax(out + 1, mx);
    }
}

cudaError_t launch_hist(int device, cudaStream_t st, const void *x, int dtype, int64_t n, int64_t lo, int64_t nbins,
                        int64_t hot, uint64_t *counts) {
    cudaError_t e = cudaMemsetAsync(counts, 0, sizeof(uint64_t) * (size_t)nbins, st);
    if (e != cudaSuccess || n == 0) return e;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    int64_t grid = (n + kHistThreads * 16 - 1) / (kHistThreads * 16);
    if (grid > (int64_t)sms * 8) grid = (int64_t)sms * 8;
    const int use_smem = nbins <= 12 * 1024;                                   // 48 KB of 32-bit bins per CTA
    k_hist<<<(unsigned)grid, kHistThreads, use_smem ? (size_t)nbins * 4 : 0, st>>>(x, dtype, n, lo, (int)nbins, hot,
                                                                                   (unsigned long long *)counts, use_smem);
    return cudaGetLastError();
}

cudaError_t launch_minmax(int device, cudaStream_t st, const void *x, int dtype, int64_t n, int64_t *out) {
    k_minmax_init<<<1, 1, 0, st>>>((long long *)out);      // a kernel, not a copy from host memory: capturable in a CUDA graph
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess || n == 0) return e;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    int64_t grid = (n + kHistThreads * 16 - 1) / (kHistThreads * 16);
    if (grid > (int64_t)sms * 8) grid = (int64_t)sms * 8;
    k_minmax<<<(unsigned)grid, kHistThreads, 0, st>>>(x, dtype, n, (long long *)out);
    return cudaGetLastError();
}

}  // namespace ivc
