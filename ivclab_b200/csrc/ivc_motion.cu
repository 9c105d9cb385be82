// ivc_motion.cu -- full-search SSD block matching and motion compensation for sm_100a.
//
// Reference: ivclab/video/motion.py:8-58 (compute_motion_vector), :60-97 (reconstruct_with_motion_vector).
//
// K3 (exact): one CTA stages the reference search window (+halo, zero outside the frame) and the
// current strip of up to 16 blocks in shared memory.  One WARP owns one 8x8 block at a time; its
// lanes enumerate (dx, dy-group) tasks, a task being G=3 vertically adjacent candidates that share
// their reference rows in registers.  Each SSD is accumulated in numpy's own summation order
// (8 column accumulators filled row by row, then a fixed pairwise tree; motion.py:46 == np.sum of
// a contiguous 64-element array) with individually rounded sub/mul/add, so motion vectors are
// bit-exact for arbitrary float frames.  The argmin is lexicographic on (ssd, index), which equals
// the reference's "first strict minimum in (dy, dx) raster order" (motion.py:35-51).
//
// K3 (integer): the same decomposition on packed uint8 planes with __vabsdiffu4 + __dp4a (4 pixels
// per instruction, exact integers), valid -- and bit-identical -- whenever both frames are
// integer-valued in [0,255]; ivc_me_full_search(IVC_ME_AUTO) checks that on the device and runs
// exactly one of the two kernels without a host round trip.
#include "ivc_dct.cuh"
#include "ivc_common.cuh"

namespace ivc {

constexpr int kMeWarps = 8;
constexpr int kMeG = 3;            // candidates per task (vertically adjacent)

struct MeArgs {
    const void *ref, *cur;
    int64_t n, H, W, ref_fs, cur_fs;
    int Hp, Wp, sr, span, ngrp, ntask;
    int nbx;                       // blocks per strip
    int strips_per_row;
    int R, P;                      // window rows / pitch (elements)
    int64_t *mv;
    const int *flag;               // optional device flag; kernel runs only if *flag == run_if
    int run_if;
};

template <typename T> struct Inf;
template <> struct Inf<double> { static __device__ __forceinline__ double v() { return __longlong_as_double(0x7ff0000000000000LL); } };
template <> struct Inf<float> { static __device__ __forceinline__ float v() { return __int_as_float(0x7f800000); } };

template <typename T>
__global__ void __launch_bounds__(kMeWarps * 32, 2) k_me_exact(const MeArgs a) {
    if (a.flag && *a.flag != a.run_if) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *s_win = reinterpret_cast<T *>(smem_raw);                 // [R][P]
    T *s_cur = s_win + (size_t)a.R * a.P;                        // [nbx][64]
    using R_ = Rn<T>;

    int64_t cta = blockIdx.x;
    const int strip = (int)(cta % a.strips_per_row);
    cta /= a.strips_per_row;
    const int by = (int)(cta % a.Hp);
    const int64_t frame = cta / a.Hp;
    const int bx0 = strip * a.nbx;
    const int nb = min(a.nbx, a.Wp - bx0);
    const T *ref = (const T *)a.ref + frame * a.ref_fs;
    const T *cur = (const T *)a.cur + frame * a.cur_fs;
    const int sr = a.sr, span = a.span;

    // ---- stage window and current strip (coalesced along x) ----
    const int Wc = 8 * a.nbx + 2 * sr;
    for (int idx = threadIdx.x; idx < a.R * Wc; idx += blockDim.x) {
        const int row = idx / Wc, col = idx - row * Wc;
        const int64_t gy = (int64_t)8 * by - sr + row, gx = (int64_t)8 * bx0 - sr + col;
        T v = (T)0;
        if (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) v = ref[gy * a.W + gx];
        s_win[row * a.P + col] = v;
    }
    for (int idx = threadIdx.x; idx < 64 * nb; idx += blockDim.x) {
        const int row = idx / (8 * nb), col = idx - row * 8 * nb;
        s_cur[(col >> 3) * 64 + row * 8 + (col & 7)] = cur[((int64_t)8 * by + row) * a.W + 8 * bx0 + col];
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int center = sr * span + sr;
    for (int b = warp; b < nb; b += kMeWarps) {
        const T *cb = s_cur + b * 64;
        T best = Inf<T>::v();
        int bidx = center;
        const int gx0 = 8 * (bx0 + b);
        for (int task = lane; task < a.ntask; task += 32) {
            const int g = task / span, dxi = task - g * span;
            const int dy0 = g * kMeG - sr;
            const int gx = gx0 + dxi - sr;
            if (gx < 0 || gx + 8 > a.W) continue;                            // motion.py:41-43 (x bound)
            T acc[kMeG][8];
#pragma unroll
            for (int gg = 0; gg < kMeG; ++gg)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[gg][j] = (T)0;
            const T *wp = s_win + (dy0 + sr) * a.P + 8 * b + dxi;
#pragma unroll
            for (int rr = 0; rr < kMeG + 7; ++rr) {
                T rv[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) rv[j] = wp[rr * a.P + j];
#pragma unroll
                for (int gg = 0; gg < kMeG; ++gg) {
                    const int i = rr - gg;                                   // row of the block for candidate gg
                    if (i >= 0 && i < 8) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const T d = R_::sub(cb[i * 8 + j], rv[j]);       // block - ref_block
                            acc[gg][j] = R_::add(acc[gg][j], R_::mul(d, d)); // r[j] += d**2, rows in order
                        }
                    }
                }
            }
#pragma unroll
            for (int gg = 0; gg < kMeG; ++gg) {
                const int dy = dy0 + gg;
                const int64_t gy = (int64_t)8 * by + dy;
                if (dy <= sr && gy >= 0 && gy + 8 <= a.H) {                  // motion.py:41-43 (y bound)
                    const T s = R_::add(R_::add(R_::add(acc[gg][0], acc[gg][1]), R_::add(acc[gg][2], acc[gg][3])),
                                        R_::add(R_::add(acc[gg][4], acc[gg][5]), R_::add(acc[gg][6], acc[gg][7])));
                    const int idx = (dy + sr) * span + dxi;                  // motion.py:55
                    if (s < best || (s == best && idx < bidx && s != Inf<T>::v())) { best = s; bidx = idx; }
                }
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const T os = __shfl_xor_sync(0xffffffffu, best, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bidx, off);
            if (os < best || (os == best && oi < bidx && os != Inf<T>::v())) { best = os; bidx = oi; }
        }
        if (lane == 0) a.mv[(frame * a.Hp + by) * (int64_t)a.Wp + bx0 + b] = bidx;
    }
}

// ---- integer fast path -------------------------------------------------------------------------
struct PackArgs {
    const void *ref, *cur;
    int64_t n, HW, ref_fs, cur_fs;
    unsigned char *ref8, *cur8;
    int *flag;
};

template <typename T>
__global__ void __launch_bounds__(256) k_me_pack_u8(const PackArgs a) {
    const int64_t total = a.n * a.HW;
    bool bad = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t f = i / a.HW, p = i - f * a.HW;
        const T r = ((const T *)a.ref)[f * a.ref_fs + p], c = ((const T *)a.cur)[f * a.cur_fs + p];
        const int ri = (int)r, ci = (int)c;                                  // saturating; NaN -> 0
        bad |= !((T)ri == r && ri >= 0 && ri <= 255 && (T)ci == c && ci >= 0 && ci <= 255);
        a.ref8[i] = (unsigned char)ri;
        a.cur8[i] = (unsigned char)ci;
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(a.flag, 1);
}

struct MeIntArgs {
    const unsigned char *ref8, *cur8;
    int64_t n, H, W;
    int Hp, Wp, sr, span, ngrp, ntask, nbx, strips_per_row, R, P;           // P in bytes, multiple of 4
    int64_t *mv;
    const int *flag;
};

__global__ void __launch_bounds__(kMeWarps * 32, 4) k_me_int(const MeIntArgs a) {
    if (a.flag && *a.flag != 0) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char *s_win = smem_raw;                                         // [R][P] bytes
    unsigned int *s_cur = reinterpret_cast<unsigned int *>(smem_raw + (((size_t)a.R * a.P + 15) & ~(size_t)15));   // [nbx][8 rows][2 words]

    int64_t cta = blockIdx.x;
    const int strip = (int)(cta % a.strips_per_row);
    cta /= a.strips_per_row;
    const int by = (int)(cta % a.Hp);
    const int64_t frame = cta / a.Hp;
    const int bx0 = strip * a.nbx;
    const int nb = min(a.nbx, a.Wp - bx0);
    const unsigned char *ref = a.ref8 + frame * a.H * a.W;
    const unsigned char *cur = a.cur8 + frame * a.H * a.W;
    const int sr = a.sr, span = a.span;

    const int Wc = 8 * a.nbx + 2 * sr;
    for (int idx = threadIdx.x; idx < a.R * a.P; idx += blockDim.x) {
        const int row = idx / a.P, col = idx - row * a.P;
        const int64_t gy = (int64_t)8 * by - sr + row, gx = (int64_t)8 * bx0 - sr + col;
        unsigned char v = 0;
        if (col < Wc && gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) v = ref[gy * a.W + gx];
        s_win[idx] = v;
    }
    unsigned char *s_cur_b = reinterpret_cast<unsigned char *>(s_cur);
    for (int idx = threadIdx.x; idx < 64 * nb; idx += blockDim.x) {
        const int row = idx / (8 * nb), col = idx - row * 8 * nb;
        s_cur_b[(col >> 3) * 64 + row * 8 + (col & 7)] = cur[((int64_t)8 * by + row) * a.W + 8 * bx0 + col];
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long kNone = ~0ull;
    for (int b = warp; b < nb; b += kMeWarps) {
        const uint2 *cb = reinterpret_cast<const uint2 *>(s_cur + b * 16);
        unsigned long long best = kNone;
        const int gx0 = 8 * (bx0 + b);
        for (int task = lane; task < a.ntask; task += 32) {
            const int g = task / span, dxi = task - g * span;
            const int dy0 = g * kMeG - sr;
            const int gx = gx0 + dxi - sr;
            if (gx < 0 || gx + 8 > a.W) continue;
            unsigned int acc[kMeG];
#pragma unroll
            for (int gg = 0; gg < kMeG; ++gg) acc[gg] = 0u;
            const int colb = 8 * b + dxi;                                    // byte column in the window
            const unsigned int *wrow = reinterpret_cast<const unsigned int *>(s_win + (dy0 + sr) * a.P) + (colb >> 2);
            const int sh = (colb & 3) * 8;
#pragma unroll
            for (int rr = 0; rr < kMeG + 7; ++rr) {
                const unsigned int *w = wrow + rr * (a.P >> 2);
                const unsigned int w0 = w[0], w1 = w[1], w2 = w[2];
                const unsigned int r0 = __funnelshift_r(w0, w1, sh), r1 = __funnelshift_r(w1, w2, sh);
#pragma unroll
                for (int gg = 0; gg < kMeG; ++gg) {
                    const int i = rr - gg;
                    if (i >= 0 && i < 8) {
                        const uint2 c = cb[i];
                        const unsigned int d0 = __vabsdiffu4(c.x, r0), d1 = __vabsdiffu4(c.y, r1);
                        acc[gg] = __dp4a(d0, d0, acc[gg]);
                        acc[gg] = __dp4a(d1, d1, acc[gg]);
                    }
                }
            }
#pragma unroll
            for (int gg = 0; gg < kMeG; ++gg) {
                const int dy = dy0 + gg;
                const int64_t gy = (int64_t)8 * by + dy;
                if (dy <= sr && gy >= 0 && gy + 8 <= a.H) {
                    const unsigned long long key = ((unsigned long long)acc[gg] << 32) | (unsigned int)((dy + sr) * span + dxi);
                    best = key < best ? key : best;
                }
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, off);
            best = o < best ? o : best;
        }
        if (lane == 0) a.mv[(frame * a.Hp + by) * (int64_t)a.Wp + bx0 + b] = (int64_t)(best & 0xffffffffull);
    }
}

// ---- K4: motion compensation (motion.py:60-97) --------------------------------------------------
struct McArgs {
    const void *ref;
    void *out;
    const int64_t *mv;
    int64_t n, H, W, C;
    int Hp, Wp, sr;
};

template <typename E>
__global__ void __launch_bounds__(256) k_mc(const McArgs a) {
    const int64_t row_e = a.W * a.C, total = a.n * a.H * row_e;
    const int64_t span = 2 * (int64_t)a.sr + 1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t xe = i % row_e;
        int64_t q = i / row_e;
        const int64_t y = q % a.H, f = q / a.H;
        const int64_t x = xe / a.C, c = xe - x * a.C;
        const int64_t idx = a.mv[(f * a.Hp + (y >> 3)) * a.Wp + (x >> 3)];
        int64_t qd = idx / span, rm = idx % span;                            // python floor div / mod
        if (rm < 0) { rm += span; qd -= 1; }
        const int64_t sy = (y & ~7LL) + qd - a.sr, sx = (x & ~7LL) + rm - a.sr;
        E v = (E)0;
        if (sy >= 0 && sy + 8 <= a.H && sx >= 0 && sx + 8 <= a.W)
            v = ((const E *)a.ref)[(f * a.H + sy + (y & 7)) * row_e + (sx + (x & 7)) * a.C + c];
        ((E *)a.out)[i] = v;
    }
}

// ---- launchers ----------------------------------------------------------------------------------
static int sm_count(int device) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    return sms;
}

template <typename A>
static void me_geometry(A &a, int64_t H, int64_t W, int sr, int elem, size_t &smem, int pitch_quantum, int pitch_skew) {
    a.Hp = (int)(H / 8); a.Wp = (int)(W / 8); a.sr = sr; a.span = 2 * sr + 1;
    a.ngrp = (a.span + kMeG - 1) / kMeG;
    a.ntask = a.ngrp * a.span;
    a.R = a.ngrp * kMeG + 7;
    int nbx = 16;
    for (;;) {
        const int Wc = 8 * nbx + 2 * sr;
        a.P = ((Wc + pitch_quantum - 1) / pitch_quantum) * pitch_quantum + pitch_skew;
        smem = (size_t)a.R * a.P * elem + (size_t)nbx * 64 * elem;
        if (smem <= 100 * 1024 || nbx == 1) break;
        nbx >>= 1;
    }
    if (nbx > a.Wp) nbx = a.Wp > 0 ? a.Wp : 1;
    a.nbx = nbx;
    a.strips_per_row = (a.Wp + nbx - 1) / nbx;
}

cudaError_t launch_me_exact(int device, cudaStream_t st, const void *ref, const void *cur, bool f32, int64_t n,
                            int64_t H, int64_t W, int64_t ref_fs, int64_t cur_fs, int sr, int64_t *mv,
                            const int *flag, int run_if) {
    MeArgs a;
    a.ref = ref; a.cur = cur; a.n = n; a.H = H; a.W = W; a.ref_fs = ref_fs; a.cur_fs = cur_fs; a.mv = mv;
    a.flag = flag; a.run_if = run_if;
    size_t smem = 0;
    me_geometry(a, H, W, sr, f32 ? 4 : 8, smem, 32, 3);
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    const int64_t ctas = n * a.Hp * (int64_t)a.strips_per_row;
    if (ctas == 0) return cudaSuccess;
    if (ctas > 2147483647LL) return cudaErrorInvalidValue;
    cudaError_t e;
    if (f32) {
        if ((e = cudaFuncSetAttribute(k_me_exact<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        k_me_exact<float><<<(unsigned)ctas, kMeWarps * 32, smem, st>>>(a);
    } else {
        if ((e = cudaFuncSetAttribute(k_me_exact<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        k_me_exact<double><<<(unsigned)ctas, kMeWarps * 32, smem, st>>>(a);
    }
    (void)device;
    return cudaGetLastError();
}

cudaError_t launch_me_pack_u8(int device, cudaStream_t st, const void *ref, const void *cur, bool f32, int64_t n,
                              int64_t H, int64_t W, int64_t ref_fs, int64_t cur_fs, unsigned char *ref8,
                              unsigned char *cur8, int *flag) {
    PackArgs a;
    a.ref = ref; a.cur = cur; a.n = n; a.HW = H * W; a.ref_fs = ref_fs; a.cur_fs = cur_fs; a.ref8 = ref8; a.cur8 = cur8;
    a.flag = flag;
    cudaError_t e = cudaMemsetAsync(flag, 0, sizeof(int), st);
    if (e != cudaSuccess) return e;
    const int64_t total = n * H * W;
    if (total == 0) return cudaSuccess;
    int64_t grid = (total + 256 * 8 - 1) / (256 * 8);
    const int64_t cap = (int64_t)sm_count(device) * 16;
    if (grid > cap) grid = cap;
    if (f32) k_me_pack_u8<float><<<(unsigned)grid, 256, 0, st>>>(a);
    else k_me_pack_u8<double><<<(unsigned)grid, 256, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_me_int(int device, cudaStream_t st, const unsigned char *ref8, const unsigned char *cur8, int64_t n,
                          int64_t H, int64_t W, int sr, int64_t *mv, const int *flag) {
    MeIntArgs a;
    a.ref8 = ref8; a.cur8 = cur8; a.n = n; a.H = H; a.W = W; a.mv = mv; a.flag = flag;
    size_t smem = 0;
    me_geometry(a, H, W, sr, 1, smem, 4, 4);      // +4 bytes: the funnel shift reads one word past the last column
    smem = (((size_t)a.R * a.P + 15) & ~(size_t)15) + (size_t)a.nbx * 64;
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    const int64_t ctas = n * a.Hp * (int64_t)a.strips_per_row;
    if (ctas == 0) return cudaSuccess;
    if (ctas > 2147483647LL) return cudaErrorInvalidValue;
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(k_me_int, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    k_me_int<<<(unsigned)ctas, kMeWarps * 32, smem, st>>>(a);
    (void)device;
    return cudaGetLastError();
}

cudaError_t launch_mc(int device, cudaStream_t st, const void *ref, int elem_size, int64_t n, int64_t H, int64_t W,
                      int64_t C, const int64_t *mv, int sr, void *out) {
    McArgs a;
    a.ref = ref; a.out = out; a.mv = mv; a.n = n; a.H = H; a.W = W; a.C = C; a.Hp = (int)(H / 8); a.Wp = (int)(W / 8); a.sr = sr;
    const int64_t total = n * H * W * C;
    if (total == 0) return cudaSuccess;
    int64_t grid = (total + 256 * 4 - 1) / (256 * 4);
    const int64_t cap = (int64_t)sm_count(device) * 16;
    if (grid > cap) grid = cap;
    switch (elem_size) {
        case 1: k_mc<unsigned char><<<(unsigned)grid, 256, 0, st>>>(a); break;
        case 2: k_mc<unsigned short><<<(unsigned)grid, 256, 0, st>>>(a); break;
        case 4: k_mc<unsigned int><<<(unsigned)grid, 256, 0, st>>>(a); break;
        case 8: k_mc<unsigned long long><<<(unsigned)grid, 256, 0, st>>>(a); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace ivc
