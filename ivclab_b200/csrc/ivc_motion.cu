// ivc_motion.cu -- full-search SSD block matching and motion compensation for sm_100a.
//
// Reference: ivclab/video/motion.py:8-58 (compute_motion_vector), :60-97 (reconstruct_with_motion_vector).
//
// Both search kernels work on a 2-D CTA tile of TBY x TBX blocks: the CTA stages the reference search
// window of the whole tile (+-sr halo, zero outside the frame) and the current blocks in shared memory
// once, so halo re-reads are ~1.3x instead of 2-5x and every CTA has thousands of candidates to chew on.
// A TASK is G=3 vertically adjacent candidates (dy0..dy0+2, dx) of one block: the three candidates
// share their reference rows in registers.
//
// K3 exact (k_me_exact<T>): one WARP owns one block at a time, lanes enumerate its tasks.  Each SSD
// is accumulated in numpy's own summation order (8 column accumulators filled row by row, then a
// fixed pairwise tree; motion.py:46 == np.sum of a contiguous 64-element array) with individually
// rounded sub/mul/add, so motion vectors are bit-exact for arbitrary float frames.  The argmin is
// lexicographic on (ssd, index) == the reference's "first strict minimum in (dy, dx) raster order"
// (motion.py:35-51).
//
// K3 integer (k_me_int<T>): reads the float frames directly, converts them to packed uint8 while
// staging (raising a device flag if any value is not an integer in [0,255]) and evaluates SSDs with
// __vabsdiffu4 + __dp4a (4 pixels per instruction, exact integers).  Tasks of all blocks of the tile
// are flattened over the CTA's threads; the per-block argmin is a shared-memory atomicMin on the
// packed key (ssd << k | index).  ivc_me_full_search(IVC_ME_AUTO) launches k_me_int and then
// k_me_exact, which exits immediately unless the flag was raised -- no host round trip, no workspace
// beyond the 4-byte flag.
#include "ivc_dct.cuh"
#include "ivc_common.cuh"

namespace ivc {

constexpr int kMeWarps = 8;
constexpr int kMeThreads = kMeWarps * 32;
constexpr int kMeG = 3;            // candidates per task (vertically adjacent)

struct MeArgs {
    const void *ref, *cur;
    int64_t n, H, W, ref_fs, cur_fs;
    int Hp, Wp, sr, span, ngrp, ntpb;      // ntpb = tasks per block
    int tby, tbx;                          // CTA tile in blocks
    int tiles_y, tiles_x;
    int R, P, Wc;                          // window rows / pitch / used columns (elements or bytes)
    int cur_off;                           // byte offset of the current-blocks area in dynamic smem
    int64_t *mv;
    int *flag;                             // device flag (may be null)
    int run_if;                            // exact kernel: run only if *flag == run_if (when flag != null)
    int check;                             // int kernel: 1 = validate integer-valuedness and raise the flag
};

template <typename T> struct Inf;
template <> struct Inf<double> { static __device__ __forceinline__ double v() { return __longlong_as_double(0x7ff0000000000000LL); } };
template <> struct Inf<float> { static __device__ __forceinline__ float v() { return __int_as_float(0x7f800000); } };

template <int BYTES>
__device__ __forceinline__ void cp_async_zfill(uint32_t dst, const void *src, bool valid) {
    const uint32_t n = valid ? BYTES : 0;
    if (BYTES == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}

struct MeTile {
    int64_t frame;
    int by0, bx0, nby, nbx;
};
__device__ __forceinline__ MeTile me_tile(const MeArgs &a) {
    MeTile t;
    int64_t cta = blockIdx.x;
    const int tx = (int)(cta % a.tiles_x);
    cta /= a.tiles_x;
    const int ty = (int)(cta % a.tiles_y);
    t.frame = cta / a.tiles_y;
    t.by0 = ty * a.tby;
    t.bx0 = tx * a.tbx;
    t.nby = min(a.tby, a.Hp - t.by0);
    t.nbx = min(a.tbx, a.Wp - t.bx0);
    return t;
}

// ================================================================================================
// exact float kernel
// ================================================================================================
template <typename T>
__global__ void __launch_bounds__(kMeThreads, 2) k_me_exact(const MeArgs a) {
    if (a.flag && *a.flag != a.run_if) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *s_win = reinterpret_cast<T *>(smem_raw);                         // [R][P]
    T *s_cur = reinterpret_cast<T *>(smem_raw + a.cur_off);             // [tby*tbx][64]
    using R_ = Rn<T>;
    const MeTile tl = me_tile(a);
    const T *ref = (const T *)a.ref + tl.frame * a.ref_fs;
    const T *cur = (const T *)a.cur + tl.frame * a.cur_fs;
    const int sr = a.sr, span = a.span;

    // ---- stage window and current blocks with cp.async (element-wise, zero-fill outside the frame):
    //      every copy of the tile is in flight at once, so the HBM/L2 latency is paid once per CTA ----
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rows_used = 8 * tl.nby + 2 * sr;
    const uint32_t win_s = (uint32_t)__cvta_generic_to_shared(s_win), cur_s = (uint32_t)__cvta_generic_to_shared(s_cur);
    for (int row = warp; row < a.R; row += kMeWarps) {
        const int64_t gy = (int64_t)8 * tl.by0 - sr + row;
        const bool row_ok = row < rows_used && gy >= 0 && gy < a.H;
        const T *rp = ref + (row_ok ? gy : 0) * a.W;
        for (int col = lane; col < a.Wc; col += 32) {
            const int64_t gx = (int64_t)8 * tl.bx0 - sr + col;
            const bool ok = row_ok && gx >= 0 && gx < a.W;
            cp_async_zfill<sizeof(T)>(win_s + (uint32_t)(row * a.P + col) * (uint32_t)sizeof(T), ok ? rp + gx : ref, ok);
        }
    }
    const int cw = 8 * tl.nbx;
    for (int row = warp; row < 8 * tl.nby; row += kMeWarps) {
        const T *cp = cur + ((int64_t)8 * tl.by0 + row) * a.W + 8 * tl.bx0;
        for (int col = lane; col < cw; col += 32)
            cp_async_zfill<sizeof(T)>(cur_s + (uint32_t)(((row >> 3) * a.tbx + (col >> 3)) * 64 + (row & 7) * 8 + (col & 7)) *
                                                  (uint32_t)sizeof(T), cp + col, true);
    }
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    const int center = sr * span + sr;
    const int nblk = tl.nby * tl.nbx;
    for (int blk = warp; blk < nblk; blk += kMeWarps) {
        const int brow = blk / tl.nbx, b = blk - brow * tl.nbx;
        const T *cb = s_cur + (brow * a.tbx + b) * 64;
        T best = Inf<T>::v();
        int bidx = center;
        const int gx0 = 8 * (tl.bx0 + b);
        const int64_t gy0 = (int64_t)8 * (tl.by0 + brow);
        for (int task = lane; task < a.ntpb; task += 32) {
            const int g = task / span, dxi = task - g * span;
            const int dy0 = g * kMeG - sr;
            const int gx = gx0 + dxi - sr;
            if (gx < 0 || gx + 8 > a.W) continue;                            // motion.py:41-43 (x bound)
            T acc[kMeG][8];
#pragma unroll
            for (int gg = 0; gg < kMeG; ++gg)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[gg][j] = (T)0;
            const T *wp = s_win + (8 * brow + dy0 + sr) * a.P + 8 * b + dxi;
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
                const int64_t gy = gy0 + dy;
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
        if (lane == 0) a.mv[(tl.frame * a.Hp + tl.by0 + brow) * (int64_t)a.Wp + tl.bx0 + b] = bidx;
    }
}

// ================================================================================================
// integer kernel (packed uint8, converted from the float frames while staging)
// ================================================================================================
// value -> uint8 plus "is an integer in [0,255]" without the (quarter-rate) FP64 conversion instructions:
// v + 2^52 leaves the integer in the low mantissa word; (v + 2^52) - 2^52 == v proves v had no fraction.
__device__ __forceinline__ unsigned to_u8_checked(double v, bool &bad) {
    const double s = __dadd_rn(v, 4503599627370496.0);
    const unsigned iv = (unsigned)__double2loint(s);
    bad |= !((__double2hiint(s) == 0x43300000) & (iv <= 255u) & (__dsub_rn(s, 4503599627370496.0) == v));
    return iv & 255u;
}
__device__ __forceinline__ unsigned to_u8_checked(float v, bool &bad) {
    const int iv = (int)v;                                                    // saturating; NaN -> 0
    bad |= !((float)iv == v && iv >= 0 && iv <= 255);
    return (unsigned)iv & 255u;
}

template <typename T>
__device__ __forceinline__ unsigned pack4(const T *p, int64_t base, int64_t gx, int64_t W, bool row_ok, bool &bad) {
    unsigned w = 0;
    if (row_ok && gx >= 0 && gx + 3 < W) {                                    // interior word: four independent loads
        T v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = p[base + gx + k];
#pragma unroll
        for (int k = 0; k < 4; ++k) w |= to_u8_checked(v[k], bad) << (8 * k);
    } else if (row_ok) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (gx + k >= 0 && gx + k < W) w |= to_u8_checked(p[base + gx + k], bad) << (8 * k);
    }
    return w;
}

// floor(x / d) for the small operands of the task decomposition: one multiply-high
struct FastDiv {
    unsigned magic, d;
    __device__ __forceinline__ explicit FastDiv(unsigned dd) : magic(dd > 1 ? 0xFFFFFFFFu / dd + 1u : 0u), d(dd) {}
    __device__ __forceinline__ int div(int x) const { return d > 1 ? (int)__umulhi((unsigned)x, magic) : x; }   // x*d < 2^32
};

template <typename T>
__global__ void __launch_bounds__(kMeThreads, 3) k_me_int(const MeArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned *s_win = reinterpret_cast<unsigned *>(smem_raw);                 // [R][P/4] words
    unsigned *s_cur = reinterpret_cast<unsigned *>(smem_raw + a.cur_off);     // [tby*tbx][8 rows][2 words]
    __shared__ unsigned long long s_best[64];
    __shared__ int s_bad;
    const MeTile tl = me_tile(a);
    const T *ref = (const T *)a.ref + tl.frame * a.ref_fs;
    const T *cur = (const T *)a.cur + tl.frame * a.cur_fs;
    const int sr = a.sr, span = a.span;
    const int nblk = tl.nby * tl.nbx;

    if (threadIdx.x < 64) s_best[threadIdx.x] = ~0ull;
    if (threadIdx.x == 0) s_bad = 0;
    bool bad = false;
    const int pw = a.P >> 2, rows_used = 8 * tl.nby + 2 * sr;
    for (int idx = threadIdx.x; idx < a.R * pw; idx += blockDim.x) {
        const int row = idx / pw, w = idx - row * pw;
        const int64_t gy = (int64_t)8 * tl.by0 - sr + row, gx = (int64_t)8 * tl.bx0 - sr + 4 * w;
        s_win[idx] = pack4(ref, gy * a.W, gx, a.W, row < rows_used && gy >= 0 && gy < a.H && 4 * w < a.Wc, bad);
    }
    const int cww = 2 * tl.nbx;                                               // words per current row
    for (int idx = threadIdx.x; idx < 8 * tl.nby * cww; idx += blockDim.x) {
        const int row = idx / cww, w = idx - row * cww;
        const int64_t gy = (int64_t)8 * tl.by0 + row;
        s_cur[((row >> 3) * a.tbx + (w >> 1)) * 16 + (row & 7) * 2 + (w & 1)] =
            pack4(cur, gy * a.W, (int64_t)8 * tl.bx0 + 4 * w, a.W, true, bad);
    }
    __syncthreads();
    if (a.check) {
        if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) s_bad = 1;
        __syncthreads();
        if (s_bad) {                                                          // not an integer frame: leave it to k_me_exact
            if (threadIdx.x == 0) atomicOr(a.flag, 1);
            return;
        }
    }

    const bool small = span * span <= 512;                                    // ssd < 2^22, index < 2^9: key fits 32 bits
    unsigned *s_best32 = reinterpret_cast<unsigned *>(s_best);
    const int total = nblk * a.ntpb;
    const FastDiv d_ntpb((unsigned)a.ntpb), d_span((unsigned)span), d_nbx((unsigned)tl.nbx);
    for (int task = threadIdx.x; task < total; task += blockDim.x) {
        const int blk = d_ntpb.div(task), rem = task - blk * a.ntpb;
        const int g = d_span.div(rem), dxi = rem - g * span;
        const int brow = d_nbx.div(blk), b = blk - brow * tl.nbx;
        const int dy0 = g * kMeG - sr;
        const int gx = 8 * (tl.bx0 + b) + dxi - sr;
        if (gx < 0 || gx + 8 > a.W) continue;
        const uint2 *cb = reinterpret_cast<const uint2 *>(s_cur + (brow * a.tbx + b) * 16);
        uint2 c[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) c[i] = cb[i];
        unsigned acc[kMeG];
#pragma unroll
        for (int gg = 0; gg < kMeG; ++gg) acc[gg] = 0u;
        const int colb = 8 * b + dxi;                                         // byte column in the window
        const unsigned *wrow = s_win + (8 * brow + dy0 + sr) * pw + (colb >> 2);
        const int sh = (colb & 3) * 8;
#pragma unroll
        for (int rr = 0; rr < kMeG + 7; ++rr) {
            const unsigned *w = wrow + rr * pw;
            const unsigned w0 = w[0], w1 = w[1], w2 = w[2];
            const unsigned r0 = __funnelshift_r(w0, w1, sh), r1 = __funnelshift_r(w1, w2, sh);
#pragma unroll
            for (int gg = 0; gg < kMeG; ++gg) {
                const int i = rr - gg;
                if (i >= 0 && i < 8) {
                    const unsigned d0 = __vabsdiffu4(c[i].x, r0), d1 = __vabsdiffu4(c[i].y, r1);
                    acc[gg] = __dp4a(d0, d0, acc[gg]);
                    acc[gg] = __dp4a(d1, d1, acc[gg]);
                }
            }
        }
        unsigned long long key = ~0ull;
        const int64_t gy0 = (int64_t)8 * (tl.by0 + brow);
#pragma unroll
        for (int gg = 0; gg < kMeG; ++gg) {
            const int dy = dy0 + gg;
            const int64_t gy = gy0 + dy;
            if (dy <= sr && gy >= 0 && gy + 8 <= a.H) {
                const unsigned long long k = ((unsigned long long)acc[gg] << 32) | (unsigned)((dy + sr) * span + dxi);
                key = k < key ? k : key;
            }
        }
        if (key != ~0ull) {
            if (small) atomicMin(s_best32 + blk, ((unsigned)(key >> 32) << 9) | (unsigned)(key & 511u));
            else atomicMin(s_best + blk, key);
        }
    }
    __syncthreads();
    if ((int)threadIdx.x < nblk) {
        const int blk = threadIdx.x, brow = blk / tl.nbx, b = blk - brow * tl.nbx;
        const int64_t idx = small ? (int64_t)(s_best32[blk] & 511u) : (int64_t)(s_best[blk] & 0xffffffffull);
        a.mv[(tl.frame * a.Hp + tl.by0 + brow) * (int64_t)a.Wp + tl.bx0 + b] = idx;
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

// choose the CTA tile (at most 64 blocks) and the shared-memory layout
static size_t me_geometry(MeArgs &a, int64_t n_frames, int64_t H, int64_t W, int sr, int elem, int pitch_quantum,
                          int pitch_skew, size_t budget, int64_t min_ctas) {
    a.Hp = (int)(H / 8); a.Wp = (int)(W / 8); a.sr = sr; a.span = 2 * sr + 1;
    a.ngrp = (a.span + kMeG - 1) / kMeG;
    a.ntpb = a.ngrp * a.span;
    // largest tile that fits the shared-memory budget AND still yields enough CTAs to fill the GPU a few
    // times over (single small frames -- QCIF, one 1080p frame of a closed loop -- get smaller tiles)
    static const int shapes[][2] = {{4, 16}, {2, 16}, {2, 8}, {1, 8}, {1, 4}, {1, 2}, {1, 1}};
    size_t smem = 0;
    for (auto &s : shapes) {
        a.tby = s[0]; a.tbx = s[1];
        const int64_t ctas = n_frames * ((a.Hp + a.tby - 1) / a.tby) * (int64_t)((a.Wp + a.tbx - 1) / a.tbx);
        if (ctas < min_ctas && !(s[0] == 1 && s[1] == 1) && s[0] * s[1] > 8) continue;
        a.R = 8 * (a.tby - 1) + a.ngrp * kMeG + 7;                 // >= 8*tby + 2*sr, covers the last dy-group
        a.Wc = 8 * a.tbx + 2 * sr;
        a.P = ((a.Wc + pitch_quantum - 1) / pitch_quantum) * pitch_quantum + pitch_skew;
        a.cur_off = (int)((((size_t)a.R * a.P * elem) + 15) & ~(size_t)15);
        smem = (size_t)a.cur_off + (size_t)a.tby * a.tbx * 64 * elem;
        if (smem <= budget) break;
    }
    a.tiles_y = (a.Hp + a.tby - 1) / a.tby;
    a.tiles_x = (a.Wp + a.tbx - 1) / a.tbx;
    return smem;
}

cudaError_t launch_me_exact(int device, cudaStream_t st, const void *ref, const void *cur, bool f32, int64_t n,
                            int64_t H, int64_t W, int64_t ref_fs, int64_t cur_fs, int sr, int64_t *mv,
                            int *flag, int run_if) {
    MeArgs a;
    a.ref = ref; a.cur = cur; a.n = n; a.H = H; a.W = W; a.ref_fs = ref_fs; a.cur_fs = cur_fs; a.mv = mv;
    a.flag = flag; a.run_if = run_if; a.check = 0;
    const size_t smem = me_geometry(a, n, H, W, sr, f32 ? 4 : 8, 32, 3, 100 * 1024, 4 * 2 * (int64_t)sm_count(device));
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    const int64_t ctas = n * a.tiles_y * (int64_t)a.tiles_x;
    if (ctas == 0) return cudaSuccess;
    if (ctas > 2147483647LL) return cudaErrorInvalidValue;
    cudaError_t e;
    if (f32) {
        if ((e = cudaFuncSetAttribute(k_me_exact<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        k_me_exact<float><<<(unsigned)ctas, kMeThreads, smem, st>>>(a);
    } else {
        if ((e = cudaFuncSetAttribute(k_me_exact<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        k_me_exact<double><<<(unsigned)ctas, kMeThreads, smem, st>>>(a);
    }
    (void)device;
    return cudaGetLastError();
}

cudaError_t launch_me_int(int device, cudaStream_t st, const void *ref, const void *cur, bool f32, int64_t n,
                          int64_t H, int64_t W, int64_t ref_fs, int64_t cur_fs, int sr, int64_t *mv, int *flag,
                          int check) {
    MeArgs a;
    a.ref = ref; a.cur = cur; a.n = n; a.H = H; a.W = W; a.ref_fs = ref_fs; a.cur_fs = cur_fs; a.mv = mv;
    a.flag = flag; a.run_if = 0; a.check = check;
    // pitch in bytes: multiple of 4 plus 4 (the funnel shift reads one word past the last column)
    const size_t smem = me_geometry(a, n, H, W, sr, 1, 4, 4, 64 * 1024, 4 * 3 * (int64_t)sm_count(device));
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    const int64_t ctas = n * a.tiles_y * (int64_t)a.tiles_x;
    if (ctas == 0) return cudaSuccess;
    if (ctas > 2147483647LL) return cudaErrorInvalidValue;
    cudaError_t e;
    if (check && (e = cudaMemsetAsync(flag, 0, sizeof(int), st)) != cudaSuccess) return e;
    if (f32) {
        if ((e = cudaFuncSetAttribute(k_me_int<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        k_me_int<float><<<(unsigned)ctas, kMeThreads, smem, st>>>(a);
    } else {
        if ((e = cudaFuncSetAttribute(k_me_int<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        k_me_int<double><<<(unsigned)ctas, kMeThreads, smem, st>>>(a);
    }
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
