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
// (motion.py:35-51).  It serves float32 frames (and IVC_ME_EXACT_V1=1).
//
// K3 exact, second generation (k_me_exact2, float64 frames): the same answer with about one such evaluation per block --
// a byte prefilter (vabsdiff4 + dp4a on the tile quantised with its own affine map) and a rounding bound prove every
// other candidate away; k_me_exact2<STEP> goes on to code and reconstruct the warp's blocks (the closed-loop P-frame in
// one kernel, ivc_pframe_step).
//
// K3 integer (k_me_int<T,G,PC>): reads the float frames directly (or uint8 planes as they are), converts them to
// packed uint8 while staging (raising a device flag if any value is not an integer in [0,255]) and evaluates
// SSD = sum(c^2) + sum(r^2) - 2 sum(c r): the cross term is one __dp4a per four candidate-pixels, sum(r^2) comes
// from a per-window-position table built once per CTA, sum(c^2) once per block -- all exact integers.  Tasks of
// all blocks of the tile are flattened over the CTA's threads; the per-block argmin is a shared-memory atomicMin
// on the packed key (ssd, index).  ivc_me_full_search(IVC_ME_AUTO) launches k_me_int and then k_me_exact, which
// exits immediately unless the flag was raised -- no host round trip, no workspace beyond the 4-byte flag.
#include <cstdio>
#include <algorithm>
#include <cstdlib>
#include <type_traits>
#include "ivc_dct.cuh"
#include "ivc_common.cuh"
#include "ivc_tile.cuh"

namespace ivc {

constexpr int kMeWarps = 8;
constexpr int kMeThreads = kMeWarps * 32;
constexpr int kMeG = 3;            // candidates per task (vertically adjacent)
constexpr int kExactCurPitch = 66; // elements per current block in the exact kernel's shared memory (64 + 2: a warp
                                   // that straddles two blocks reads them from different banks)

struct MeArgs {
    const void *ref, *cur;
    int64_t n, H, W, ref_fs, cur_fs;
    int Hp, Wp, sr, span, ngrp, ntpb;      // ntpb = tasks per block
    int tby, tbx;                          // CTA tile in blocks
    int tiles_y, tiles_x;
    int R, P, Wc;                          // window rows / pitch / used columns (elements or bytes)
    int cur_off;                           // byte offset of the current-blocks area in dynamic smem
    int part_off, npieces;                 // exact kernel: partial (ssd, index) results per block and lane
    int pw, hs_off, b_off;                 // integer kernel: words per packed row, offsets of HS and B
    int pwl, nseg;                         // ... words per row actually staged, 8-row segments of S
    int vec;                               // ... 16-byte staging loads are legal (alignment of base, strides, sr)
    unsigned m_pwl, m_q4, m_p4, m_ntpb, m_span;     // multiply-high reciprocals (host-computed)
    unsigned m_nbx[2], m_cww[2];                    // ... of nbx and 2*nbx for full / last-column tiles
    int ipr[2];                                     // int kernels, float64 staging: 8-pixel items per window row (full / last-column tiles)
    unsigned m_ipr[2];
    int64_t *mv;
    int *flag;                             // device flag (may be null)
    int run_if;                            // exact kernel: run only if *flag == run_if (when flag != null)
    int check;                             // int kernel: 1 = validate integer-valuedness and raise the flag
    // fused search + P-frame forward (k_me_int<.., PF = true>): quantiser table, scan-index output, channels stored per block
    const void *table;
    int table_dtype;
    int32_t *zz;
    int och;
    int win32_off, cur32_off, acand_off, acand_pitch;   // k_me_exact2: unaligned-word view (packed rows at b_off), quantised blocks, per-warp candidate scores
    double *recon;                         // fused closed-loop step (k_me_exact<double, STEP>): the decoder's reconstruction of the frame
    int work_off;                          // ... byte offset of the warps' WORK buffers in dynamic shared memory
    int32_t *zr_counts;                    // optional: zero-run symbol count and non-zero mask per scan block (see ivc_tile.cuh)
    unsigned long long *zr_masks;
};

template <typename T> struct Inf;
template <> struct Inf<double> { static __device__ __forceinline__ double v() { return __longlong_as_double(0x7ff0000000000000LL); } };
template <> struct Inf<float> { static __device__ __forceinline__ float v() { return __int_as_float(0x7f800000); } };

template <int BYTES>
__device__ __forceinline__ void cp_async_zfill(uint32_t dst, const void *src, bool valid) {
    const uint32_t n = valid ? BYTES : 0;
    if (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
    else if (BYTES == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}

// floor(x / d) for the small operands of the index decompositions: one multiply-high (x*d < 2^32)
struct FastDiv {
    unsigned magic;                                                           // 0 = divide by one
    __device__ __forceinline__ explicit FastDiv(unsigned m) : magic(m) {}
    __device__ __forceinline__ int div(int x) const { return magic ? (int)__umulhi((unsigned)x, magic) : x; }
};
static unsigned fastdiv_magic(unsigned d) { return d > 1 ? 0xFFFFFFFFu / d + 1u : 0u; }

struct MeTile {
    int64_t frame;
    int by0, bx0, nby, nbx;
};
// grid: x = tile (row-major over tiles_y x tiles_x), y = frame (launchers chunk batches above 65535 frames)
__device__ __forceinline__ MeTile me_tile(const MeArgs &a) {
    MeTile t;
    const int ty = (int)blockIdx.x / a.tiles_x, tx = (int)blockIdx.x - ty * a.tiles_x;
    t.frame = blockIdx.y;
    t.by0 = ty * a.tby;
    t.bx0 = tx * a.tbx;
    t.nby = min(a.tby, a.Hp - t.by0);
    t.nbx = min(a.tbx, a.Wp - t.bx0);
    return t;
}

// ================================================================================================
// fused closed-loop step: what follows the exact search for the warp's own blocks
// ================================================================================================
// One P-frame of the closed loop is search -> MC + residual + DCT + quantise + zig-zag -> dequantise + IDCT +
// prediction add (exercises/ch4/E4-1.py:257-306; videocodec.py:52-75), and per block it needs nothing but the block, its
// search window and its vector -- all of which the exact search kernel holds in shared memory.  With decode = "luma"
// (channel 0 of every block) nothing crosses a tile, so the warp that found a block's vector codes and reconstructs it on
// the spot: the frame pair is read once, one launch per frame instead of three, and no kernel ramps up or drains between
// the phases.  Arithmetic and operation order are those of k_pframe_forward_tm / k_pframe_inverse_tm (ivc_transform.cu):
// lane (r, u) owns pixel row r of blocks u and u + 4 of the group in the row passes and column r in the column passes.
constexpr int kStepWork = 4 * kP3TU * 8;                                      // 4352 B per warp: transposition buffer / scan staging
constexpr int kStepCurPitch = 66;                                             // == kExactCurPitch (defined below)

// mvs[i * mv_stride] = vector of block blk_first + i; TUB = bytes per u-plane of the transposition buffer (1088 for groups
// of up to eight blocks, 576 when the caller never passes more than four: WORK is then 3264 bytes)
template <int TUB>
__device__ __forceinline__ void pstep_group(const MeArgs &a, const MeTile &tl, const double *s_win, const double *s_cur,
                                            const int *mvs, int mv_stride, int blk_first, int nb, FastDiv d_nbx, unsigned char *work_b,
                                            const double *s_rt, const double *s_t, const double *s_tT, bool chroma_twice) {
    const int lane = threadIdx.x & 31, r = lane & 7, u = lane >> 3;
    const uint32_t work_s = smem_u32(work_b);
    const FastDiv d_span(a.m_span);
    const int mc = nb > 4 ? 2 : 1;                      // a group of at most four blocks leaves the m = 1 half empty: skip its arithmetic (warp-uniform)
    double x[2][8];
    int wrow[2], wcol[2], gby[2], gbx[2];                                     // window position of the prediction, block coordinates
    bool valid[2];
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        const int bi = u + 4 * m;
        valid[m] = bi < nb;
        if (m >= mc) continue;
        const int blk = blk_first + (valid[m] ? bi : 0);
        const int brow = d_nbx.div(blk), b = blk - brow * tl.nbx;
        const int idx = mvs[(valid[m] ? bi : 0) * mv_stride];                 // the block's vector
        const int dyi = d_span.div(idx), dxi = idx - dyi * a.span;
        wrow[m] = 8 * brow + dyi;
        wcol[m] = 8 * b + dxi;
        gby[m] = tl.by0 + brow;
        gbx[m] = tl.bx0 + b;
        const double *cb = s_cur + (brow * a.tbx + b) * kStepCurPitch + r * 8;
        const double *wp = s_win + (wrow[m] + r) * a.P + wcol[m];
#pragma unroll
        for (int k = 0; k < 8; ++k) x[m][k] = valid[m] ? __dsub_rn(cb[k], wp[k]) : 0.0;   // residual = cur - prediction (videocodec.py:71)
    }
    // ---- forward: rows, transpose, columns ----
#pragma unroll
    for (int m = 0; m < 2; ++m)
        if (m < mc) dct2_8(x[m]);
    bulk_wait_read0();                                  // an earlier group's stores have drained WORK
    __syncwarp();
    unsigned char *t_base = work_b + u * TUB;
#pragma unroll
    for (int m = 0; m < 2; ++m)
        if (m < mc) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                *reinterpret_cast<double *>(t_base + ((((r >> 1) ^ (j >> 1)) << 1) + (r & 1)) * 8 + (m * 8 + j) * 64) = x[m][j];
        }
    __syncwarp();
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        if (m >= mc) continue;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double2 v = *reinterpret_cast<const double2 *>(t_base + r * 64 + ((k ^ (r >> 1)) << 4) + m * 512);
            x[m][2 * k] = v.x;
            x[m][2 * k + 1] = v.y;
        }
        dct2_8(x[m]);
    }
    __syncwarp();
    // ---- quantise against [lum, chrom, chrom] (patchquant.py:59), zig-zag, store; keep channel 0 for the decoder ----
    const int nch = (chroma_twice || a.och == 2) ? 2 : 3;
    const double *rt_l = s_rt + r, *t_l = s_t + r;
    unsigned char *stage = work_b + u * (kStageUF * 4);
    int q0[2][8];                                       // channel 0 of row r (columns j): the decoder's input
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        if (m >= mc) continue;
        if (m == 1) { bulk_wait_read0(); __syncwarp(); }     // round 0's stores have drained the staging area
        for (int ch = 0; ch < nch; ++ch) {
            QuantGuard qg;
            int qv[8];
#pragma unroll
            for (int v = 0; v < 8; ++v) qv[v] = qg.q(x[m][v], rt_l[ch * 64 + v * 8]);
            if (__builtin_expect(qg.risky(), 0)) {
#pragma unroll
                for (int v = 0; v < 8; ++v) qv[v] = quantize_exact_f64(x[m][v], t_l[ch * 64 + v * 8]);
            }
#pragma unroll
            for (int v = 0; v < 8; ++v) {
                const int pos = ZZ_ORDER[v * 8 + r];
                *reinterpret_cast<int *>(stage + pos * 4 + ch * 256) = qv[v];
                if (ch == 1 && nch == 2) *reinterpret_cast<int *>(stage + pos * 4 + 512) = qv[v];   // channel 2 repeats channel 1
            }
        }
        fence_proxy_async();
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) q0[m][j] = *reinterpret_cast<const int *>(stage + ZZ_ORDER[r * 8 + j] * 4);   // coefficient (r, j)
        {
            const int bi = lane + 4 * m;                  // lane u (< 4) stores the och scan blocks of the group's block u + 4m
            if (lane < 4 && bi < nb) {
                const int blk = blk_first + bi, brow = d_nbx.div(blk), b = blk - brow * tl.nbx;
                bulk_s2g(a.zz + ((tl.frame * a.Hp + tl.by0 + brow) * (int64_t)a.Wp + tl.bx0 + b) * (64 * a.och),
                         work_s + lane * (kStageUF * 4), 256u * (uint32_t)a.och);
            }
        }
        bulk_commit();                                    // every lane commits (possibly empty) groups: counts stay in step
    }
    // ---- decoder: dequantise with the luminance table, IDCT rows then columns, + prediction (videocodec.py:74) ----
    {
        int mx = 0;
#pragma unroll
        for (int m = 0; m < 2; ++m)
            if (m < mc) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const double pr = __dmul_rn(i32_to_f64(q0[m][j]), s_tT[j * 8 + r]);     // s_tT[j*8 + r] = lum[r][j]
                    mx = max(mx, __double2hiint(pr) & 0x7fffffff);
                    x[m][j] = trunc_f64_small(pr);
                }
            }
        if (__builtin_expect(mx >= 0x41E00000, 0)) {
#pragma unroll
            for (int m = 0; m < 2; ++m)
                if (m < mc) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) x[m][j] = dequantize_f64(q0[m][j], s_tT[j * 8 + r]);
                }
        }
    }
#pragma unroll
    for (int m = 0; m < 2; ++m)
        if (m < mc) dct3_8(x[m]);
    bulk_wait_read0();                                  // the scan stores have read the staging area: WORK is free again
    __syncwarp();
#pragma unroll
    for (int m = 0; m < 2; ++m)
        if (m < mc) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                *reinterpret_cast<double *>(t_base + ((((r >> 1) ^ (j >> 1)) << 1) + (r & 1)) * 8 + (m * 8 + j) * 64) = x[m][j];
        }
    __syncwarp();
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        if (m >= mc) continue;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double2 v = *reinterpret_cast<const double2 *>(t_base + r * 64 + ((k ^ (r >> 1)) << 4) + m * 512);
            x[m][2 * k] = v.x;
            x[m][2 * k + 1] = v.y;
        }
        dct3_8(x[m]);
    }
    __syncwarp();                                       // WORK may be rewritten by the next group from here on
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        if (!valid[m]) continue;
        const double *wp = s_win + wrow[m] * a.P + wcol[m] + r;                // column r of the prediction
        double *out = a.recon + tl.frame * (a.H * a.W) + ((int64_t)gby[m] * 8) * a.W + gbx[m] * 8 + r;
#pragma unroll
        for (int i = 0; i < 8; ++i) out[i * a.W] = __dadd_rn(wp[i * a.P], x[m][i]);   // recon = prediction + recon_residual
    }
}

// ================================================================================================
// exact float kernel
// ================================================================================================
template <typename T, bool STEP = false>
__global__ void __launch_bounds__(kMeThreads, 2) k_me_exact(const MeArgs a) {
    if (a.flag && *a.flag != a.run_if) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double s_qt[STEP ? 448 : 1];              // STEP: fl(1/t) [192], t [192], luminance table transposed [64]
    bool chroma_twice = false;
    if (STEP) {
        const int t_ = threadIdx.x;
        if (t_ < 192) {
            const double t = load_table_elem(a.table, a.table_dtype, t_);
            s_qt[192 + t_] = t;
            s_qt[t_] = __drcp_rn(t);
            if (t_ < 64) s_qt[384 + (t_ & 7) * 8 + (t_ >> 3)] = t;
        }
        bool same = true;
        if (t_ < 64)
            same = __double_as_longlong(load_table_elem(a.table, a.table_dtype, 64 + t_)) ==
                   __double_as_longlong(load_table_elem(a.table, a.table_dtype, 128 + t_));
        chroma_twice = __syncthreads_and(same) != 0;
    }
    T *s_win = reinterpret_cast<T *>(smem_raw);                         // [R][P]
    T *s_cur = reinterpret_cast<T *>(smem_raw + a.cur_off);             // [tby*tbx][kExactCurPitch]
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
            cp_async_zfill<sizeof(T)>(cur_s + (uint32_t)(((row >> 3) * a.tbx + (col >> 3)) * kExactCurPitch + (row & 7) * 8 + (col & 7)) *
                                                  (uint32_t)sizeof(T), cp + col, true);
    }
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    // A warp owns a contiguous range of the tile's blocks and walks the tasks of ALL its blocks as one list, 32
    // at a time (a block has ntpb tasks, rarely a multiple of 32: one block per round would idle 5 of 32 lanes
    // at +-4).  Every lane folds its result into a private shared-memory slot (block, lane); one butterfly per
    // block at the end turns the 32 slots into the block's vector.  No warp ever waits on another warp.
    const int center = sr * span + sr;
    const int nblk = tl.nby * tl.nbx;
    const int per_w = (nblk + kMeWarps - 1) / kMeWarps, blk0 = warp * per_w;        // blocks blk0 .. blk0+nb_w-1
    const int nb_w = max(0, min(per_w, nblk - blk0));
    const int total = nb_w * a.ntpb;
    T *s_pssd = reinterpret_cast<T *>(smem_raw + a.part_off);                // [64][32]
    int *s_pidx = reinterpret_cast<int *>(s_pssd + 64 * 32);                  // [64][32]
    for (int k = 0; k < nb_w; ++k) {
        s_pssd[(blk0 + k) * 32 + lane] = Inf<T>::v();
        s_pidx[(blk0 + k) * 32 + lane] = center;
    }
    __syncwarp();
    const FastDiv d_ntpb(a.m_ntpb), d_span(a.m_span), d_nbx(a.m_nbx[tl.nbx != a.tbx]);
    for (int tbase = 0; tbase < total; tbase += 32) {
        const int task = tbase + lane;
        const bool have = task < total;
        const int tk = have ? task : total - 1;
        const int kb = d_ntpb.div(tk), tt = tk - kb * a.ntpb;
        const int blk = blk0 + kb;
        const int brow = d_nbx.div(blk), b = blk - brow * tl.nbx;
        const int g = d_span.div(tt), dxi = tt - g * span;
        const T *cb = s_cur + (brow * a.tbx + b) * kExactCurPitch;
        T best = Inf<T>::v();
        int bidx = center;
        const int gx = 8 * (tl.bx0 + b) + dxi - sr;
        const int64_t gy0 = (int64_t)8 * (tl.by0 + brow);
        const int dy0 = g * kMeG - sr;
        if (have && gx >= 0 && gx + 8 <= a.W) {                                  // motion.py:41-43 (x bound)
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
        if (have) {                                                          // fold into this lane's slot of the block
            const T ps = s_pssd[blk * 32 + lane];
            const int pi = s_pidx[blk * 32 + lane];
            if (best < ps || (best == ps && bidx < pi && best != Inf<T>::v())) {
                s_pssd[blk * 32 + lane] = best;
                s_pidx[blk * 32 + lane] = bidx;
            }
        }
    }
    __syncwarp();
    for (int k = 0; k < nb_w; ++k) {
        const int blk = blk0 + k, brow = d_nbx.div(blk), b = blk - brow * tl.nbx;
        T best = s_pssd[blk * 32 + lane];
        int bidx = s_pidx[blk * 32 + lane];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const T os = __shfl_xor_sync(0xffffffffu, best, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bidx, off);
            if (os < best || (os == best && oi < bidx && os != Inf<T>::v())) { best = os; bidx = oi; }
        }
        if (lane == 0) {
            a.mv[(tl.frame * a.Hp + tl.by0 + brow) * (int64_t)a.Wp + tl.bx0 + b] = bidx;
            if (STEP) s_pidx[blk * 32] = bidx;
        }
    }
    if constexpr (STEP && sizeof(T) == 8) {
        __syncwarp();
        unsigned char *work_b = smem_raw + a.work_off + warp * kStepWork;
        for (int k0 = 0; k0 < nb_w; k0 += 8)
            pstep_group<kP3TU * 8>(a, tl, reinterpret_cast<const double *>(s_win), reinterpret_cast<const double *>(s_cur),
                                   s_pidx + (blk0 + k0) * 32, 32, blk0 + k0, min(8, nb_w - k0), d_nbx, work_b, s_qt, s_qt + 192,
                                   s_qt + 384, chroma_twice);
        bulk_wait_all0();
    }
}

// ================================================================================================
// exact search, second generation (float64 frames): 8-bit prefilter, exact evaluation of the survivors
// ================================================================================================
// The vector of a block is the FIRST strict minimum of the exactly rounded SSDs (motion.py:35-51) -- but almost every
// candidate loses by a margin no rounding can bridge.  The tile is quantised to bytes with an affine map
// v' = (v - lo) / q -- a fixed first guess that holds an 8-bit frame and its reconstruction's overshoot; if a value leaves
// it, or the tile spans too few levels, the tile's OWN range (1 / q an integer whenever it spans 1 .. 255, so integer-valued
// frames are quantised exactly) -- delta = 1/2 for each side that was rounded at all; the packed-byte
// machinery of the integer search (unaligned-word view, vabsdiff4 + dp4a: two instructions per four pixels) scores
// every candidate with S~ = sum of squared byte differences, and in quantised units
//     |S' - S~| <= 2 delta sum|d~| + 64 delta^2 <= 16 delta sqrt(S~) + 64 delta^2 =: eps(S~)       (Cauchy-Schwarz)
// where S' = S / q^2 is the real SSD.  A candidate can only win or tie if S~ - eps(S~) <= min (S~ + eps(S~)); typically
// one or two survive (on integer-valued frames delta = 0 and only exact ties do).  The survivors -- all in-frame
// candidates if the tile holds NaN / Inf or is flat -- are evaluated exactly, in numpy's own order, eight lanes per
// candidate: lane j is numpy's column accumulator j (rows in order), three shuffles are its pairwise tree.  Ascending
// candidate order and a strict "<" reproduce the reference loop.  With STEP the warp then codes and reconstructs its
// blocks (see pstep_group).
constexpr int kCurPitch = 18;      // words per current block in shared memory (16 + 2: blocks on distinct banks)
constexpr double kX2GuessLo = -64.0, kX2GuessInvQ = 255.0 / 384.0;                 // 85 / 128: exact
constexpr int kX2Work = 4 * kStageUF * 4;                                     // 3264 B per warp: groups of four blocks

__device__ __forceinline__ float sqrt_approx(float x) {          // MUFU.SQRT: relative error ~2^-22, sqrt(0) = 0
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// four values -> one word of bytes under the map x = (v - lo) * inv_q.  Everything the tile needs to know about the values
// is OR-ed / min-maxed into the accumulators and looked at once per tile:
//   bad     != 0  <=> some x left [0, 255] (NaN and Inf do)
//   nonint  != 0  <=> some x is not an integer (-0.0 counts as one)
//   kmin, kmax    (KEYS) order-preserving integer keys of the high words of the VALUES: the tile's range
struct X2Acc {
    unsigned bad = 0u, nonint = 0u;
    int kmin = 0x7fffffff, kmax = (int)0x80000000;
};
template <bool KEYS>
__device__ __forceinline__ unsigned x2_quant4(const double *p, int cnt, double lo, double inv_q, X2Acc &acc) {
    unsigned w = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k < cnt) {
            const double v = p[k], x = __dmul_rn(__dsub_rn(v, lo), inv_q);
            const double sft = __dadd_rn(x, 4503599627370496.0);              // 2^52: rint(x) in the low mantissa word
            const double back = __dadd_rn(sft, -4503599627370496.0);
            const unsigned n = (unsigned)__double2loint(sft);
            acc.bad |= ((unsigned)__double2hiint(sft) ^ 0x43300000u) | (n >> 8);
            acc.nonint |= ((unsigned)__double2hiint(back) ^ (unsigned)__double2hiint(x)) | ((unsigned)__double2loint(back) ^ (unsigned)__double2loint(x));
            if (KEYS) {
                const int hi = __double2hiint(v), key = hi ^ ((hi >> 31) & 0x7fffffff);
                acc.kmin = min(acc.kmin, key); acc.kmax = max(acc.kmax, key);
            }
            w |= (n & 255u) << (8 * k);
        }
    }
    return w;
}

// CSPAN / CP: compile-time search span and window pitch (0 = take them from the arguments): with constants every row
// offset of the unrolled loops is an immediate and the index divisions are multiply-shifts
template <bool STEP, int CSPAN = 0, int CP = 0>
__global__ void __launch_bounds__(kMeThreads, CSPAN ? 3 : 2) k_me_exact2(const MeArgs a) {
    if (a.flag && *a.flag != a.run_if) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double s_qt[STEP ? 448 : 1];              // STEP: fl(1/t) [192], t [192], luminance table transposed [64]
    __shared__ unsigned s_redi[3][kMeWarps];
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ int s_mvw[kMeWarps][8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    bool chroma_twice = false;
    double *s_win = reinterpret_cast<double *>(smem_raw);                     // [R][P]
    double *s_cur = reinterpret_cast<double *>(smem_raw + a.cur_off);         // [tby*tbx][kExactCurPitch]
    unsigned *s_u = reinterpret_cast<unsigned *>(smem_raw + a.win32_off);     // [R][PU]  unaligned-word view of the quantised window
    unsigned *s_b8 = reinterpret_cast<unsigned *>(smem_raw + a.b_off);        // [R][PW]  packed quantised window
    unsigned *s_c8 = reinterpret_cast<unsigned *>(smem_raw + a.cur32_off);    // [tby*tbx][kCurPitch] quantised blocks
    unsigned *s_A = reinterpret_cast<unsigned *>(smem_raw + a.acand_off) + warp * a.acand_pitch;   // this warp's candidate scores
    using R_ = Rn<double>;
    const MeTile tl = me_tile(a);
    const double *ref = (const double *)a.ref + tl.frame * a.ref_fs;
    const double *cur = (const double *)a.cur + tl.frame * a.cur_fs;
    const int sr = a.sr, span = CSPAN ? CSPAN : a.span, P = CP ? CP : a.P, ncand = span * span;
    const int ntpb = CSPAN ? ((CSPAN + kMeG - 1) / kMeG) * CSPAN : a.ntpb;
    const auto div_span = [&](int x) { return CSPAN ? x / (CSPAN ? CSPAN : 1) : FastDiv(a.m_span).div(x); };
    const FastDiv d_nbx(a.m_nbx[tl.nbx != a.tbx]);

    // ---- stage window and current blocks (float64, zero outside the frame), all copies in flight at once ----
    // A tile whose window lies inside the frame takes one bulk copy per window row (lanes of warp 0) and 16-byte copies for
    // the blocks; at the frame's edge (or with an odd range / width) every element is copied -- or zero-filled -- alone.
    const int rows_used = 8 * tl.nby + 2 * sr;
    const uint32_t win_s = (uint32_t)__cvta_generic_to_shared(s_win), cur_s = (uint32_t)__cvta_generic_to_shared(s_cur);
    const int64_t wy0 = (int64_t)8 * tl.by0 - sr, wx0 = (int64_t)8 * tl.bx0 - sr;
    const int cw = 8 * tl.nbx;
    const bool pairs = (((sr | a.W) & 1) == 0) && (((reinterpret_cast<uintptr_t>(ref) | reinterpret_cast<uintptr_t>(cur)) & 15) == 0);
    const bool interior = pairs && wy0 >= 0 && wy0 + a.R <= a.H && wx0 >= 0 && wx0 + a.Wc <= a.W;
    const uint32_t bar = smem_u32(&s_bar);
    if (interior) {
        if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
        __syncthreads();
        if (warp == 0) {
            if (lane == 0) mbar_expect_tx(bar, (uint32_t)(a.R * a.Wc * 8));
            __syncwarp();
            for (int row = lane; row < a.R; row += 32)
                bulk_g2s(win_s + (uint32_t)(row * P) * 8u, ref + (wy0 + row) * a.W + wx0, (uint32_t)(a.Wc * 8), bar);
        }
    } else {
        for (int row = warp; row < a.R; row += kMeWarps) {
            const int64_t gy = wy0 + row;
            const bool row_ok = row < rows_used && gy >= 0 && gy < a.H;
            const double *rp = ref + (row_ok ? gy : 0) * a.W;
            for (int col = lane; col < a.Wc; col += 32) {
                const int64_t gx = wx0 + col;
                const bool ok = row_ok && gx >= 0 && gx < a.W;
                cp_async_zfill<8>(win_s + (uint32_t)(row * P + col) * 8u, ok ? rp + gx : ref, ok);
            }
        }
    }
    if (pairs) {
        const int hw = cw >> 1;                                               // 16-byte pairs per row of blocks
        for (int idx = tid; idx < 8 * tl.nby * hw; idx += kMeThreads) {
            const int row = d_nbx.div(idx >> 2), col = 2 * (idx - row * hw);          // hw = 4 nbx
            cp_async_zfill<16>(cur_s + (uint32_t)(((row >> 3) * a.tbx + (col >> 3)) * kExactCurPitch + (row & 7) * 8 + (col & 7)) * 8u,
                               cur + ((int64_t)8 * tl.by0 + row) * a.W + 8 * tl.bx0 + col, true);
        }
    } else {
        for (int row = warp; row < 8 * tl.nby; row += kMeWarps) {
            const double *cp = cur + ((int64_t)8 * tl.by0 + row) * a.W + 8 * tl.bx0;
            for (int col = lane; col < cw; col += 32)
                cp_async_zfill<8>(cur_s + (uint32_t)(((row >> 3) * a.tbx + (col >> 3)) * kExactCurPitch + (row & 7) * 8 + (col & 7)) * 8u, cp + col, true);
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (STEP) {                                          // the tables, while the tile's copies are in flight
        if (tid < 192) {
            const double t = load_table_elem(a.table, a.table_dtype, tid);
            s_qt[192 + tid] = t;
            s_qt[tid] = __drcp_rn(t);
            if (tid < 64) s_qt[384 + (tid & 7) * 8 + (tid >> 3)] = t;
        }
        bool same = true;
        if (tid < 64)
            same = __double_as_longlong(load_table_elem(a.table, a.table_dtype, 64 + tid)) ==
                   __double_as_longlong(load_table_elem(a.table, a.table_dtype, 128 + tid));
        chroma_twice = __syncthreads_and(same) != 0;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (interior) mbar_wait(bar, 0);
    __syncthreads();

    // ---- quantise.  First guess: [-64, 320) on 255 levels -- an 8-bit frame or its reconstruction, which overshoots [0, 255]
    // wherever the content is saturated (on the bench sequences 97 % of the tiles hold such a value) ----
    const int PW = a.pw, PU = a.pwl;                                          // words per packed row / per row of the unaligned view
    const int wpr = (a.Wc + 3) >> 2, cwpr = 2 * tl.nbx;                       // words per window row / per row of blocks
    const FastDiv d_wpr(a.m_pwl);
    // flags after a pass: 1 = a value left [0, 255] (or is NaN / Inf), 2 / 4 = the window / the blocks hold a non-integer
    const auto quant_tile = [&](auto keys_tag, double lo, double inv_q, unsigned &flags, int &kmin, int &kmax) {
        constexpr bool KEYS = decltype(keys_tag)::value;
        X2Acc aw, ac;
        for (int idx = tid; idx < a.R * wpr; idx += kMeThreads) {
            const int row = d_wpr.div(idx), w = idx - row * wpr;
            s_b8[row * PW + w] = x2_quant4<KEYS>(s_win + row * P + 4 * w, a.Wc - 4 * w, lo, inv_q, aw);
        }
        for (int idx = tid; idx < 8 * tl.nby * cwpr; idx += kMeThreads) {
            const int row = d_nbx.div(idx >> 1), w = idx - row * cwpr, o = (row >> 3) * a.tbx + (w >> 1);   // cwpr = 2 nbx
            s_c8[o * kCurPitch + (row & 7) * 2 + (w & 1)] =
                x2_quant4<KEYS>(s_cur + o * kExactCurPitch + (row & 7) * 8 + (w & 1) * 4, 4, lo, inv_q, ac);
        }
        for (int idx = tid; idx < a.R * 2; idx += kMeThreads) s_b8[(idx >> 1) * PW + wpr + (idx & 1)] = 0u;   // the spare words
        flags = ((aw.bad | ac.bad) ? 1u : 0u) | (aw.nonint ? 2u : 0u) | (ac.nonint ? 4u : 0u);
        flags = __reduce_or_sync(0xffffffffu, flags);
        unsigned umin = (unsigned)min(aw.kmin, ac.kmin) ^ 0x80000000u, umax = (unsigned)max(aw.kmax, ac.kmax) ^ 0x80000000u;
        if (KEYS) { umin = __reduce_min_sync(0xffffffffu, umin); umax = __reduce_max_sync(0xffffffffu, umax); }
        if (lane == 0) { s_redi[0][warp] = flags; s_redi[1][warp] = umin; s_redi[2][warp] = umax; }
        __syncthreads();
        flags = 0u; umin = 0xffffffffu; umax = 0u;
#pragma unroll
        for (int w = 0; w < kMeWarps; ++w) { flags |= s_redi[0][w]; umin = min(umin, s_redi[1][w]); umax = max(umax, s_redi[2][w]); }
        kmin = (int)(umin ^ 0x80000000u); kmax = (int)(umax ^ 0x80000000u);
    };
    unsigned flags;
    int kmin, kmax;
    quant_tile(std::true_type{}, kX2GuessLo, kX2GuessInvQ, flags, kmin, kmax);
    // the tile's range from the keys, one step outwards: vmin <= every value <= vmax
    const auto decode = [](int key0, int step) {
        const long long k = (long long)key0 + step;
        const int key = (int)max(min(k, 0x7fffffffLL), -0x7fffffffLL - 1);
        return __hiloint2double(key ^ ((key >> 31) & 0x7fffffff), 0);
    };
    const double vmin = decode(kmin, -1), vmax = decode(kmax, 1);
    bool fallback = false;
    // the guess is kept unless it failed or resolves the tile poorly (non-integer content on fewer than 48 levels)
    if ((flags & 1u) || ((flags & 6u) && vmax - vmin < 48.0)) {
        const bool wild = !(fabs(vmin) < 1e300) || !(fabs(vmax) < 1e300);     // NaN / Inf (or next to the end of the range)
        // an INTEGER scale while the tile spans 1 .. 255: integer-valued frames are then quantised exactly
        const double lo = (vmax - floor(vmin) >= 1.0) ? floor(vmin) : vmin, range = vmax - lo;
        const double inv_q = (range >= 1.0 && range <= 255.0) ? floor(255.0 / range) : 255.0 / (range * (1.0 + 1e-12));
        fallback = wild || !(range > 0.0) || !(range < 1e290);                // no usable bound: every in-frame candidate survives
        __syncthreads();                                                      // s_redi is written again
        if (!fallback) {
            quant_tile(std::false_type{}, lo, inv_q, flags, kmin, kmax);
            fallback = (flags & 1u) != 0u;                                    // cannot happen; if it does the bound is void
        }
    }
    // the unaligned-word view: U[y][x] = bytes x .. x+3 of window row y
    if (!fallback) {
        const int q4 = PU >> 2;
        const FastDiv d_q4(a.m_q4);
        for (int idx = tid; idx < a.R * q4; idx += kMeThreads) {
            const int row = d_q4.div(idx), w = idx - row * q4;
            const unsigned *bp = s_b8 + row * PW + w;                         // PW >= q4 + 2: no guards
            const unsigned w0 = bp[0], w1 = bp[1];
            uint4 u;
            u.x = w0;
            u.y = __funnelshift_r(w0, w1, 8);
            u.z = __funnelshift_r(w0, w1, 16);
            u.w = __funnelshift_r(w0, w1, 24);
            *reinterpret_cast<uint4 *>(s_u + row * PU + 4 * w) = u;
        }
    }
    __syncthreads();
    // delta: 0 if both sides are integers after the map, else 1/2 per rounded side (plus the rounding of the map itself)
    const float delta = ((flags & 2u) ? 0.5f : 0.0f) + ((flags & 4u) ? 0.5f : 0.0f) + ((flags & 6u) ? 1e-6f : 0.0f);
    // eps(S~) = 16 delta sqrt(S~) + 64 delta^2, a little upwards (the float evaluation and numpy's own rounding of S)
    const float e16 = 16.0f * delta * 1.001f, e64 = 64.0f * delta * delta * 1.001f + 1e-3f;

    const int center = sr * span + sr;
    const int nblk = tl.nby * tl.nbx;
    const int per_w = (nblk + kMeWarps - 1) / kMeWarps, blk0 = warp * per_w;        // blocks blk0 .. blk0+nb_w-1 (per_w <= 8)
    const int nb_w = max(0, min(per_w, nblk - blk0));
    const int g8 = lane >> 3, j8 = lane & 7;
    const double DINF = Inf<double>::v();

    // byte scores of one block into this warp's s_A; returns the largest score that can still win or tie
    const auto score_block = [&](int brow, int b, int slot, int gy0, int gx0) {
        unsigned smin = 0xffffffffu;
        {
            // ---- byte scores: a task = three vertically adjacent candidates sharing their window rows in registers ----
            const uint2 *cb8 = reinterpret_cast<const uint2 *>(s_c8 + slot * kCurPitch);
            for (int tt = lane; tt < ntpb; tt += 32) {
                const int g = div_span(tt), dxi = tt - g * span, dy0 = g * kMeG;
                const int gx = gx0 + dxi;
                const bool x_ok = gx >= 0 && gx + 8 <= a.W;
                uint2 c[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) c[i] = cb8[i];
                unsigned acc[kMeG];
#pragma unroll
                for (int gg = 0; gg < kMeG; ++gg) acc[gg] = 0u;
                const unsigned *up = s_u + (8 * brow + dy0) * PU + 8 * b + dxi;
#pragma unroll
                for (int rr = 0; rr < kMeG + 7; ++rr) {
                    const unsigned r0 = up[rr * PU], r1 = up[rr * PU + 4];
#pragma unroll
                    for (int gg = 0; gg < kMeG; ++gg) {
                        const int i = rr - gg;                                 // row of the block for candidate gg
                        if (i >= 0 && i < 8) {
                            const unsigned d0 = __vabsdiffu4(c[i].x, r0), d1 = __vabsdiffu4(c[i].y, r1);
                            acc[gg] = __dp4a(d0, d0, acc[gg]);
                            acc[gg] = __dp4a(d1, d1, acc[gg]);
                        }
                    }
                }
#pragma unroll
                for (int gg = 0; gg < kMeG; ++gg) {
                    const int dyi = dy0 + gg, gy = gy0 + dyi;
                    if (dyi < span) {
                        const bool ok = x_ok && gy >= 0 && gy + 8 <= a.H;             // motion.py:41-43
                        const unsigned sc = ok ? acc[gg] : 0xffffffffu;               // all ones marks "outside the frame"
                        s_A[dyi * span + dxi] = sc;
                        smin = min(smin, sc);
                    }
                }
            }
            smin = __reduce_min_sync(0xffffffffu, smin);
            __syncwarp();
        }
        // S~ - eps(S~) <= min (S~ + eps(S~)) = smin + eps(smin) (eps grows with S~)  <=>  S~ <= lim, the root of a quadratic
        // in sqrt(S~), taken a little upwards; below 64 delta^2, where S~ - eps(S~) falls, every candidate passes anyway
        const float fmin_ = (float)smin, thr = fmin_ + __fmaf_rn(e16, sqrt_approx(fmin_), e64);
        const float rt = 0.5f * (e16 + sqrt_approx(__fmaf_rn(e16, e16, 4.0f * (thr + e64)))) * 1.00001f;
        return (unsigned)fminf(__fmaf_rn(rt, rt, 1.0f), 4.0e9f);
    };
    if constexpr (CSPAN == 9) {
        // ---- 81 candidates: four blocks at a time.  Their survivors are three ballot words each; then the warp's four groups
        // of eight lanes walk one block's list each -- typically a single exact evaluation for all four blocks together ----
        for (int k0 = 0; k0 < nb_w; k0 += 4) {
            unsigned m0 = 0u, m1 = 0u, m2 = 0u;                                   // this lane group's block: candidates still to evaluate
            for (int q = 0; q < 4 && k0 + q < nb_w; ++q) {
                const int blk = blk0 + k0 + q, brow = d_nbx.div(blk), b = blk - brow * tl.nbx;
                const int gy0 = 8 * (tl.by0 + brow) - sr, gx0 = 8 * (tl.bx0 + b) - sr;
                const unsigned lim = fallback ? 0u : score_block(brow, b, brow * a.tbx + b, gy0, gx0);
                unsigned w[3];
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const int c = 32 * j + lane;
                    bool surv = false;
                    if (c < 81) {
                        if (fallback) {
                            const int dyi = c / 9, dxi = c - dyi * 9;
                            surv = gy0 + dyi >= 0 && gy0 + dyi + 8 <= a.H && gx0 + dxi >= 0 && gx0 + dxi + 8 <= a.W;
                        } else {
                            surv = s_A[c] <= lim;
                        }
                    }
                    w[j] = __ballot_sync(0xffffffffu, surv);
                }
                if (g8 == q) { m0 = w[0]; m1 = w[1]; m2 = w[2]; }
                __syncwarp();                                                     // s_A is rewritten for the next block
            }
            const int kq = k0 + g8;
            const bool have = kq < nb_w;
            const int blk = blk0 + (have ? kq : 0), brow = d_nbx.div(blk), b = blk - brow * tl.nbx;
            const double *cb = s_cur + (brow * a.tbx + b) * kExactCurPitch + j8;  // column j8 of the block
            const double *wb = s_win + (8 * brow) * P + 8 * b + j8;
            double best = DINF;
            int bidx = center;
            while (__any_sync(0xffffffffu, (m0 | m1 | m2) != 0u)) {
                int mine = -1;                                                    // ascending order: lowest candidate first
                if (m0) { mine = __ffs(m0) - 1; m0 &= m0 - 1; }
                else if (m1) { mine = 31 + __ffs(m1); m1 &= m1 - 1; }
                else if (m2) { mine = 63 + __ffs(m2); m2 &= m2 - 1; }
                double acc = DINF;
                if (mine >= 0) {
                    const int dyi = mine / 9, dxi = mine - dyi * 9;
                    const double *wp = wb + dyi * P + dxi;
                    acc = 0.0;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {                                 // r[j] += d**2, rows in order
                        const double d = R_::sub(cb[i * 8], wp[i * P]);
                        acc = R_::add(acc, R_::mul(d, d));
                    }
                }
                acc = R_::add(acc, __shfl_xor_sync(0xffffffffu, acc, 1));        // ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7))
                acc = R_::add(acc, __shfl_xor_sync(0xffffffffu, acc, 2));
                acc = R_::add(acc, __shfl_xor_sync(0xffffffffu, acc, 4));
                if (mine >= 0 && acc < best) { best = acc; bidx = mine; }        // strict < (motion.py:48)
            }
            if (have && j8 == 0) {
                a.mv[(tl.frame * a.Hp + tl.by0 + brow) * (int64_t)a.Wp + tl.bx0 + b] = bidx;
                s_mvw[warp][kq] = bidx;
            }
            __syncwarp();
        }
    } else {
        for (int k = 0; k < nb_w; ++k) {
            const int blk = blk0 + k, brow = d_nbx.div(blk), b = blk - brow * tl.nbx;
            const int slot = brow * a.tbx + b;
            const int gy0 = 8 * (tl.by0 + brow) - sr, gx0 = 8 * (tl.bx0 + b) - sr;       // frame position of candidate (0, 0)
            const unsigned lim = fallback ? 0u : score_block(brow, b, slot, gy0, gx0);
            // ---- survivors, in ascending candidate order, four at a time: exact SSD in numpy's order ----
            double best = DINF;
            int bidx = center;
            const double *cb = s_cur + slot * kExactCurPitch + j8;                        // column j8 of the block
            for (int base = 0; base < ncand; base += 32) {
                const int c = base + lane;
                bool surv = false;
                if (c < ncand) {
                    if (fallback) {
                        const int dyi = div_span(c), dxi = c - dyi * span;
                        surv = gy0 + dyi >= 0 && gy0 + dyi + 8 <= a.H && gx0 + dxi >= 0 && gx0 + dxi + 8 <= a.W;
                    } else {
                        surv = s_A[c] <= lim;
                    }
                }
                unsigned mask = __ballot_sync(0xffffffffu, surv);
                while (mask) {
                    int mine = -1;                                                    // the candidate of this lane's group of eight
    #pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        if (mask) {
                            const int bit = __ffs(mask) - 1;
                            mask &= mask - 1;
                            if (g == g8) mine = base + bit;
                        }
                    }
                    double acc = DINF;
                    if (mine >= 0) {
                        const int dyi = div_span(mine), dxi = mine - dyi * span;
                        const double *wp = s_win + (8 * brow + dyi) * P + 8 * b + dxi + j8;
                        acc = 0.0;
    #pragma unroll
                        for (int i = 0; i < 8; ++i) {                                 // r[j] += d**2, rows in order
                            const double d = R_::sub(cb[i * 8], wp[i * P]);
                            acc = R_::add(acc, R_::mul(d, d));
                        }
                    }
                    acc = R_::add(acc, __shfl_xor_sync(0xffffffffu, acc, 1));        // ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7))
                    acc = R_::add(acc, __shfl_xor_sync(0xffffffffu, acc, 2));
                    acc = R_::add(acc, __shfl_xor_sync(0xffffffffu, acc, 4));
    #pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const double sg = __shfl_sync(0xffffffffu, acc, 8 * g);
                        const int ig = __shfl_sync(0xffffffffu, mine, 8 * g);
                        if (ig >= 0 && sg < best) { best = sg; bidx = ig; }           // ascending order: strict < (motion.py:48)
                    }
                }
            }
            if (lane == 0) {
                a.mv[(tl.frame * a.Hp + tl.by0 + brow) * (int64_t)a.Wp + tl.bx0 + b] = bidx;
                s_mvw[warp][k] = bidx;
            }
            __syncwarp();
        }
    }
    if constexpr (STEP) {
        __syncthreads();                                  // the WORK buffers lie over the search's tables (U view, bytes, scores)
        unsigned char *work_b = smem_raw + a.work_off + warp * kX2Work;
        for (int k0 = 0; k0 < nb_w; k0 += 4)
            pstep_group<576>(a, tl, s_win, s_cur, &s_mvw[warp][k0], 1, blk0 + k0, min(4, nb_w - k0), d_nbx, work_b, s_qt, s_qt + 192,
                             s_qt + 384, chroma_twice);
        bulk_wait_all0();
    }
}

// ================================================================================================
// integer kernel (packed uint8, converted from the float frames while staging)
// ================================================================================================
// SSD = sum(c^2) + sum(r^2) - 2 sum(c r): the cross term is ONE dp4a per four candidate-pixels (the direct
// form needs an absolute difference and a dp4a); sum(r^2) over an 8x8 window depends on the window
// position only and is computed once per position of the CTA's search window (a horizontal dp4a pass and
// a vertical running sum), sum(c^2) once per block.  All three terms are exact integers, so the key
// (ssd, index) and therefore the vector is exactly the reference's.
//
// Shared-memory layout per CTA:
//   B  [R][pw]  packed bytes of the search window (4 pixels per word), zero outside the frame
//   U  [R][P4]  "unaligned word" view: U[y][x] = bytes x..x+3 of row y, so a candidate's 8-byte row is
//               U[y][x], U[y][x+4] for ANY x and 32 consecutive candidates read 32 consecutive banks
//   HS [R][P4]  H[y][x] = sum_{j<8} r[y][x+j]^2, then in place S[y][x] = sum_{i<8} H[y+i][x]
//   cur         current blocks, 8 rows x 2 words each (+2 words padding per block)
// A TASK is G vertically adjacent candidates (dy0..dy0+G-1, dx) of one block sharing their G+7 window
// rows in registers; tasks of all blocks are flattened over the CTA, dx fastest.

// value -> uint8 plus "is an integer in [0,255]" without the (quarter-rate) FP64 conversion instructions:
// s = v + 2^52 leaves rint(v) in the low mantissa word, s - 2^52 == v proves v had no fraction, and the
// word pair of s must read (0x43300000, 0..255).  err accumulates violated bits, ne "not exact".
struct U8Check {
    unsigned err = 0;
    bool ne = false;
    __device__ __forceinline__ bool bad() const { return (err != 0) | ne; }
};
__device__ __forceinline__ unsigned to_u8(double v, U8Check &c) {
    const double s = __dadd_rn(v, 4503599627370496.0);
    const unsigned lo = (unsigned)__double2loint(s);
    c.err |= ((unsigned)__double2hiint(s) ^ 0x43300000u) | (lo & 0xffffff00u);
    c.ne |= __dsub_rn(s, 4503599627370496.0) != v;
    return lo;
}
__device__ __forceinline__ unsigned to_u8(float v, U8Check &c) {
    const float s = __fadd_rn(v, 8388608.0f);                                 // 2^23
    const unsigned n = __float_as_uint(s) ^ 0x4b000000u;
    c.err |= n & 0xffffff00u;
    c.ne |= __fsub_rn(s, 8388608.0f) != v;
    return n;
}
__device__ __forceinline__ unsigned pack_bytes(unsigned b0, unsigned b1, unsigned b2, unsigned b3) {
    return __byte_perm(__byte_perm(b0, b1, 0x0040), __byte_perm(b2, b3, 0x0040), 0x5410);
}

constexpr int kMeIntMaxThreads = 512;   // integer kernel, constant-pitch variants: 256..512 threads per CTA (chosen by the launcher), <= 64 registers
constexpr int kStageUnroll = 4;    // packed words (16 pixel loads) in flight per thread while staging

// 16-byte read-only loads (two doubles / four floats)
__device__ __forceinline__ void ldg16(const double *p, double (&v)[2]) {
    asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "l"(p));
}
__device__ __forceinline__ void ldg16(const float *p, float (&v)[4]) {
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p));
}

// ---- staging: frames -> packed bytes in shared memory ------------------------------------------------------------
// The words of a tile form ONE list: the search window of the reference frame first (word idx < n_ref covers pixels
// (ry0 + idx / pwl, rx0 + 4 (idx % pwl) .. +3), zero where that leaves the frame, stored at s_b[row * pw + w]), then the
// current blocks (stored block by block in s_cur).  A thread takes UNR words per batch and issues all their loads
// before the first conversion, so the window AND the blocks are in flight together and a tile waits for memory once
// or twice, not once per array and remainder (what is left after the full batches goes into ONE guarded batch of the
// smallest sufficient size).  VEC: origins, W and the frame bases are multiples of 16 bytes (uint8 planes: 4), so
// 16-byte groups are loaded whole and never straddle a frame edge.
struct StageGeom {
    int H, W;
    int ry0, rx0, n_ref, pwl, pw;          // reference window: origin, words, words per staged row, words per smem row
    int cy0, cx0, cww, tbx;                // current blocks: origin, words per row, blocks per tile row (smem slots)
    unsigned m_pwl, m_cww;
    int total;
};

struct StageItem {
    int gy, gx, out;                       // pixel position in the frame, shared-memory word index (< 0: nothing to do)
    bool is_cur;
};
__device__ __forceinline__ StageItem stage_item(const StageGeom &g, int idx, int cur_off_words) {
    StageItem it;
    it.is_cur = idx >= g.n_ref;
    const int j = it.is_cur ? idx - g.n_ref : idx;
    const int wpr = it.is_cur ? g.cww : g.pwl;
    const int row = FastDiv(it.is_cur ? g.m_cww : g.m_pwl).div(j), w = j - row * wpr;
    it.gy = (it.is_cur ? g.cy0 : g.ry0) + row;
    it.gx = (it.is_cur ? g.cx0 : g.rx0) + 4 * w;
    const int o_ref = row * g.pw + w;
    const int o_cur = cur_off_words + ((row >> 3) * g.tbx + (w >> 1)) * kCurPitch + (row & 7) * 2 + (w & 1);
    it.out = idx < g.total ? (it.is_cur ? o_cur : o_ref) : -1;
    return it;
}

// s_b and s_cur are addressed through one base (s_b) and the word offset of s_cur from it
template <bool VEC, int UNR, typename T>
__device__ __forceinline__ void stage_batch(const T *ref, const T *cur, const StageGeom &g, int base, U8Check &chk,
                                            unsigned *s_b, int cur_off_words) {
    constexpr int V = 16 / (int)sizeof(T);                                    // elements per 16-byte group
    const int nthr = blockDim.x;
    T v[UNR][4];
    int out[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
        const StageItem it = stage_item(g, base + u * nthr, cur_off_words);
        out[u] = it.out;
        const bool rok = it.out >= 0 && (unsigned)it.gy < (unsigned)g.H;
        const T *p = (it.is_cur ? cur : ref) + (it.gy * g.W + it.gx);         // H * W < 2^31 (checked by the launcher)
        if (VEC) {
#pragma unroll
            for (int k = 0; k < 4; k += V) {
                T t[V];
#pragma unroll
                for (int j = 0; j < V; ++j) t[j] = (T)0;
                if (rok && (unsigned)(it.gx + k) < (unsigned)g.W) ldg16(p + k, t);
#pragma unroll
                for (int j = 0; j < V; ++j) v[u][k + j] = t[j];
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) v[u][k] = (rok && (unsigned)(it.gx + k) < (unsigned)g.W) ? __ldg(p + k) : (T)0;
        }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
        const unsigned b0 = to_u8(v[u][0], chk), b1 = to_u8(v[u][1], chk), b2 = to_u8(v[u][2], chk), b3 = to_u8(v[u][3], chk);
        if (out[u] >= 0) s_b[out[u]] = pack_bytes(b0, b1, b2, b3);
    }
}
// uint8 planes: a packed word IS four consecutive pixels -- one aligned 32-bit load (VEC) or four byte loads;
// nothing to convert, nothing to check.
template <bool VEC, int UNR>
__device__ __forceinline__ void stage_batch(const unsigned char *ref, const unsigned char *cur, const StageGeom &g, int base,
                                            U8Check &, unsigned *s_b, int cur_off_words) {
    const int nthr = blockDim.x;
    unsigned v[UNR];
    int out[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
        const StageItem it = stage_item(g, base + u * nthr, cur_off_words);
        out[u] = it.out;
        const bool rok = it.out >= 0 && (unsigned)it.gy < (unsigned)g.H;
        const unsigned char *p = (it.is_cur ? cur : ref) + (it.gy * g.W + it.gx);
        v[u] = 0u;
        if (VEC) {
            if (rok && (unsigned)it.gx < (unsigned)g.W) v[u] = __ldg(reinterpret_cast<const unsigned *>(p));
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (rok && (unsigned)(it.gx + k) < (unsigned)g.W) v[u] |= (unsigned)__ldg(p + k) << (8 * k);
        }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u)
        if (out[u] >= 0) s_b[out[u]] = v[u];
}

template <bool VEC, int UNR, typename T>
__device__ __forceinline__ void stage_tile(const T *ref, const T *cur, const StageGeom &g, U8Check &chk, unsigned *s_b,
                                           int cur_off_words) {
    const int nthr = blockDim.x, tid = threadIdx.x;
    int done = 0;
    for (; done + UNR * nthr <= g.total; done += UNR * nthr) stage_batch<VEC, UNR>(ref, cur, g, done + tid, chk, s_b, cur_off_words);
    const int rem = g.total - done;                                           // < UNR * nthr: one guarded batch
    if (rem <= 0) return;
    if (rem <= nthr) stage_batch<VEC, 1>(ref, cur, g, done + tid, chk, s_b, cur_off_words);
    else if (UNR >= 2 && rem <= 2 * nthr) stage_batch<VEC, (UNR >= 2 ? 2 : 1)>(ref, cur, g, done + tid, chk, s_b, cur_off_words);
    else if (UNR >= 4 && rem <= 4 * nthr) stage_batch<VEC, (UNR >= 4 ? 4 : 1)>(ref, cur, g, done + tid, chk, s_b, cur_off_words);
    else stage_batch<VEC, UNR>(ref, cur, g, done + tid, chk, s_b, cur_off_words);
}

// float64 frames with 16-byte aligned rows (the codecs' case): the lean staging path.  An ITEM is eight consecutive
// pixels of one frame row -- four 16-byte loads from one address, two packed words, one 8-byte store: a row of a
// current block, or an 8-pixel column group of a window row.  Everything about an item but its pixels is computed
// once per eight pixels; the validity tests ride on the FP64 pipe, which the search itself leaves idle (five FP64
// operations per pixel: v + 2^52, the exactness probe (v + 2^52) - 2^52 != v, 0 <= v, v <= 255), so the integer pipe
// sees about two instructions per pixel.  EDGE: the window leaves the frame somewhere -- every 16-byte group is
// tested (groups never straddle an edge: origins and W are even) and zero-filled; CHECK: validate the values.
template <bool CHECK>
__device__ __forceinline__ unsigned to_u8_fp(double v, bool &bad) {
    const double s = __dadd_rn(v, 4503599627370496.0);                        // 2^52: rint(v) lands in the low mantissa word
    if (CHECK) bad |= (__dsub_rn(s, 4503599627370496.0) != v) | !(v >= 0.0) | !(v <= 255.0);
    return (unsigned)__double2loint(s);
}

// One list of items (rows x per_row, eight pixels each) of one frame: thread t takes items t, t + nthr, ... and walks
// (row, c) incrementally -- no division per item; the loads of item k + 1 are issued before item k is converted.
// dst32(row, c) gives the 32-bit shared-memory address of the item's two packed words.
template <bool EDGE, bool CHECK, typename Dst>
__device__ __forceinline__ void stage_item_list(const double *frame, int H, int W, int y0, int x0, int rows, int per_row,
                                                FastDiv d_per_row, bool &bad, Dst dst32) {
    const int nthr = blockDim.x;
    const int dr = d_per_row.div(nthr), dc = nthr - dr * per_row;             // what advancing by nthr items does to (row, c)
    int row = d_per_row.div((int)threadIdx.x), c = (int)threadIdx.x - row * per_row;
    const auto load = [&](int r, int cc, double (&v)[8]) {
        const int gy = y0 + r, gx = x0 + 8 * cc;
        const double *p = frame + (gy * W + gx);                               // H * W < 2^31 (checked by the launcher)
        const bool rok = !EDGE || (unsigned)gy < (unsigned)H;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double t[2] = {0.0, 0.0};
            if (rok && (!EDGE || (unsigned)(gx + 2 * k) < (unsigned)W)) ldg16(p + 2 * k, t);
            v[2 * k] = t[0];
            v[2 * k + 1] = t[1];
        }
    };
    const auto convert_store = [&](int r, int cc, const double (&v)[8]) {
        const unsigned w0 = pack_bytes(to_u8_fp<CHECK>(v[0], bad), to_u8_fp<CHECK>(v[1], bad), to_u8_fp<CHECK>(v[2], bad), to_u8_fp<CHECK>(v[3], bad));
        const unsigned w1 = pack_bytes(to_u8_fp<CHECK>(v[4], bad), to_u8_fp<CHECK>(v[5], bad), to_u8_fp<CHECK>(v[6], bad), to_u8_fp<CHECK>(v[7], bad));
        asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(dst32(r, cc)), "r"(w0), "r"(w1) : "memory");
    };
    const auto advance = [&](int &r, int &cc) {
        r += dr;
        cc += dc;
        if (cc >= per_row) { cc -= per_row; ++r; }
    };
    double va[8], vb[8];
    if (row >= rows) return;
    load(row, c, va);
    for (;;) {
        int r2 = row, c2 = c;
        advance(r2, c2);
        const bool have_b = r2 < rows;
        if (have_b) load(r2, c2, vb);
        convert_store(row, c, va);
        if (!have_b) break;
        row = r2; c = c2;
        advance(row, c);
        const bool have_a = row < rows;
        if (have_a) load(row, c, va);
        convert_store(r2, c2, vb);
        if (!have_a) break;
    }
}

template <bool EDGE, bool CHECK>
__device__ __forceinline__ bool stage_items_f64(const double *ref, const double *cur, const MeArgs &a, const MeTile &tl, int R,
                                                int pw, unsigned *s_b, unsigned *s_cur) {
    const int H = (int)a.H, W = (int)a.W, sr = a.sr;
    const int last_col = tl.nbx != a.tbx;
    bool bad = false;
    const uint32_t b32 = (uint32_t)__cvta_generic_to_shared(s_b), c32 = (uint32_t)__cvta_generic_to_shared(s_cur);
    const int tbx = a.tbx;
    // search window: rows x ipr items, stored row by row in s_b
    stage_item_list<EDGE, CHECK>(ref, H, W, 8 * tl.by0 - sr, 8 * tl.bx0 - sr, min(R, 8 * tl.nby + 2 * sr), a.ipr[last_col],
                                 FastDiv(a.m_ipr[last_col]), bad,
                                 [&](int r, int c) { return b32 + (uint32_t)(r * pw + 2 * c) * 4u; });
    // current blocks: 8 nby rows x nbx items (an item is one row of one block), stored block by block; never outside the frame
    stage_item_list<false, CHECK>(cur, H, W, 8 * tl.by0, 8 * tl.bx0, 8 * tl.nby, tl.nbx, FastDiv(a.m_nbx[last_col]), bad,
                                  [&](int r, int c) { return c32 + (uint32_t)(((r >> 3) * tbx + c) * kCurPitch + (r & 7) * 2) * 4u; });
    return bad;
}

// Phases A-C of the integer kernels: stage window + current blocks as packed bytes, build the unaligned-word view U,
// the per-position sums of squares S (x16) and the per-block sum(c^2) (x16).  PC: compile-time U / HS pitch (0 = take
// it from the arguments).  Returns false when the frames are not integer-valued (the device flag has been raised and
// the CTA must leave the vectors to k_me_exact).  Ends with a CTA-wide barrier.
template <typename T, int PC, int UNR, int KEYSHIFT = 4, bool BAKEX = false>
__device__ __forceinline__ bool me_int_prepare(const MeArgs &a, const MeTile &tl, unsigned *s_u, unsigned *s_hs, unsigned *s_b,
                                               unsigned *s_cur, unsigned *s_c2) {
    const T *ref = (const T *)a.ref + tl.frame * a.ref_fs;
    const T *cur = (const T *)a.cur + tl.frame * a.cur_fs;
    const int sr = a.sr, P4 = PC ? PC : a.P, pw = PC ? PC / 4 + 2 : a.pw, R = a.R;
    const int H = (int)a.H, W = (int)a.W;
    const int nblk = tl.nby * tl.nbx;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int last_col = tl.nbx != a.tbx;
    const FastDiv d_nbx(a.m_nbx[last_col]);

    // ---- phase A: frames -> packed bytes (window with halo, current blocks).  Rows / words past the part
    //      of the window this tile can use stay unwritten: only discarded candidates see them ----
    int bad = 0;
    {
        U8Check chk;
        StageGeom g;
        g.H = H; g.W = W;
        g.ry0 = 8 * tl.by0 - sr; g.rx0 = 8 * tl.bx0 - sr;
        g.pwl = a.pwl; g.pw = pw; g.m_pwl = a.m_pwl;
        g.n_ref = min(R, 8 * tl.nby + 2 * sr) * a.pwl;
        g.cy0 = 8 * tl.by0; g.cx0 = 8 * tl.bx0;
        g.cww = 2 * tl.nbx; g.tbx = a.tbx; g.m_cww = a.m_cww[last_col];
        g.total = g.n_ref + 8 * tl.nby * g.cww;
        const int cur_off_words = (int)(s_cur - s_b);
        bool fast = false;
        if (sizeof(T) == 8) {                                                 // float64 frames, 16-byte aligned rows: the lean path
            fast = a.vec && (sr & 1) == 0 && (pw & 1) == 0;                   // 8-byte stores into s_b rows
            if (fast) {
                const bool edge = g.ry0 < 0 || g.rx0 < 0 || g.ry0 + R > H || g.rx0 + 8 * a.ipr[last_col] > W;
                const double *r64 = (const double *)ref, *c64 = (const double *)cur;
                bool b;
                if (edge) b = a.check ? stage_items_f64<true, true>(r64, c64, a, tl, R, pw, s_b, s_cur) : stage_items_f64<true, false>(r64, c64, a, tl, R, pw, s_b, s_cur);
                else b = a.check ? stage_items_f64<false, true>(r64, c64, a, tl, R, pw, s_b, s_cur) : stage_items_f64<false, false>(r64, c64, a, tl, R, pw, s_b, s_cur);
                bad = b;
            }
        }
        if (!fast) {
            if (a.vec) stage_tile<true, UNR>(ref, cur, g, chk, s_b, cur_off_words);
            else stage_tile<false, UNR>(ref, cur, g, chk, s_b, cur_off_words);
        }
        bad = bad | (a.check && chk.bad());
    }
    if (__syncthreads_or(bad)) {                                              // not an integer frame: leave it to k_me_exact
        if (tid == 0) atomicOr(a.flag, 1);
        return false;
    }

    // ---- phase B: unaligned-word view U and horizontal sums of squares H; sum(c^2) per block ----
    {
        const int q4 = P4 >> 2;
        const FastDiv d_q4(a.m_q4);
        for (int idx = tid; idx < R * q4; idx += nthr) {
            const int row = d_q4.div(idx), w = idx - row * q4;
            const unsigned *bp = s_b + row * pw + w;                          // pw >= q4 + 2: no guards
            const unsigned w0 = bp[0], w1 = bp[1], w2 = bp[2];
            uint4 u, h;
            u.x = w0;
            u.y = __funnelshift_r(w0, w1, 8);
            u.z = __funnelshift_r(w0, w1, 16);
            u.w = __funnelshift_r(w0, w1, 24);
            const unsigned v1 = __funnelshift_r(w1, w2, 8), v2 = __funnelshift_r(w1, w2, 16), v3 = __funnelshift_r(w1, w2, 24);
            h.x = __dp4a(u.x, u.x, __dp4a(w1, w1, 0u));
            h.y = __dp4a(u.y, u.y, __dp4a(v1, v1, 0u));
            h.z = __dp4a(u.z, u.z, __dp4a(v2, v2, 0u));
            h.w = __dp4a(u.w, u.w, __dp4a(v3, v3, 0u));
            *reinterpret_cast<uint4 *>(s_u + idx * 4) = u;                    // == row * P4 + 4 * w
            *reinterpret_cast<uint4 *>(s_hs + idx * 4) = h;
        }
        if (tid < nblk) {
            const int brow = d_nbx.div(tid), b = tid - brow * tl.nbx;
            const unsigned *cb = s_cur + (brow * a.tbx + b) * kCurPitch;
            unsigned c2 = 0;
#pragma unroll
            for (int i = 0; i < 16; ++i) c2 = __dp4a(cb[i], cb[i], c2);
            s_c2[tid] = c2 << KEYSHIFT;
        }
    }
    __syncthreads();

    // ---- phase C: S[y][x] = sum_{i<8} H[y+i][x], in place.  An item = 8 output rows of one column;
    //      items are ordered by rows, so what a round overwrites is never read by a later round.  HS has
    //      8*nseg+7 rows: the tail rows hold garbage and only feed S rows no candidate uses ----
    {
        const int total = a.nseg * P4;
        for (int base = 0; base < total; base += nthr) {
            const int item = base + tid;
            const bool active = item < total;
            const int seg = active ? FastDiv(a.m_p4).div(item) : 0;
            unsigned *hp = s_hs + item + 7 * seg * P4;                        // == (8 * seg) * P4 + x
            unsigned sv[8];
            if (active) {
                unsigned h[15];
#pragma unroll
                for (int i = 0; i < 15; ++i) h[i] = hp[i * P4];
                sv[0] = ((h[0] + h[1]) + (h[2] + h[3])) + ((h[4] + h[5]) + (h[6] + h[7]));
#pragma unroll
                for (int j = 1; j < 8; ++j) sv[j] = sv[j - 1] + h[j + 7] - h[j - 1];
            }
            __syncthreads();
            if (active) {
                const unsigned xcol = BAKEX ? (unsigned)(item - seg * P4) : 0u;     // BAKEX: the column rides in the free low bits
#pragma unroll
                for (int j = 0; j < 8; ++j) hp[j * P4] = (sv[j] << KEYSHIFT) + xcol; // S * 2^KEYSHIFT: room for a candidate code
            }
        }
    }
    __syncthreads();
    return true;
}

// ---- fused P-frame forward for one group of up to eight horizontally adjacent blocks of an integer-search tile ----
// What k_pframe_forward_tm does per tile (ivc_transform.cu: residual -> DCT rows -> transpose -> DCT columns ->
// quantise against [lum, chrom, chrom] -> zig-zag -> 768-byte bulk stores), fed from the tile's BYTES instead of
// float64 planes: on integer-valued frames the staged bytes ARE the pixel values, so residual = (double)(c - p) is
// the reference's `y_channel - prediction` bit for bit (videocodec.py:71) and everything after it is the same rounded
// arithmetic as the stand-alone kernel.  Lane (r, u) owns pixel row r of blocks u and u + 4 of the group in the
// row pass and column r in the column pass.  work_b: 4352 bytes of warp-private shared memory (the search's U / S
// tables are dead by now and lend their space).
constexpr int kMePfThreads = 288, kMePfCtas = 3;                              // the fused kernel's launch shape (+-4: 576 tasks = two rounds of 288)
constexpr int kPfWork = 4 * kP3TU * 8;                                        // 4352: transposition buffer >= scan staging (3264)
template <int PC>
__device__ __forceinline__ void pf_forward_group(const MeArgs &a, const MeTile &tl, int brow, int b0, int nb, const unsigned *s_b,
                                                 const unsigned *s_cur, const unsigned *s_best32, unsigned char *work_b,
                                                 const double *s_rt, const double *s_t, bool chroma_twice) {
    constexpr int pw = PC / 4 + 2;
    const int lane = threadIdx.x & 31, r = lane & 7, u = lane >> 3;
    const uint32_t work_s = smem_u32(work_b);
    double x[2][8];
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        const int bi = u + 4 * m, b = b0 + bi;
        if (bi < nb) {
            const unsigned *cb = s_cur + (brow * a.tbx + b) * kCurPitch + 2 * r;
            const unsigned c0 = cb[0], c1 = cb[1];
            const int idx = (int)(s_best32[brow * tl.nbx + b] & 511u);         // the block's vector (motion.py:55)
            const int dyi = FastDiv(a.m_span).div(idx), dxi = idx - dyi * a.span;
            const int xx = 8 * b + dxi;
            const unsigned *wp = s_b + (8 * brow + dyi + r) * pw + (xx >> 2);
            const unsigned w0 = wp[0], w1 = wp[1], w2 = wp[2];
            const int sh = (xx & 3) * 8;
            const unsigned p0 = __funnelshift_r(w0, w1, sh), p1 = __funnelshift_r(w1, w2, sh);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                x[m][k] = i32_to_f64((int)((c0 >> (8 * k)) & 255u) - (int)((p0 >> (8 * k)) & 255u));
                x[m][4 + k] = i32_to_f64((int)((c1 >> (8 * k)) & 255u) - (int)((p1 >> (8 * k)) & 255u));
            }
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) x[m][k] = 0.0;
        }
    }
#pragma unroll
    for (int m = 0; m < 2; ++m) dct2_8(x[m]);
    bulk_wait_read0();                                  // an earlier group's stores have drained WORK
    __syncwarp();
    {
        unsigned char *t_wr = work_b + u * (kP3TU * 8) + (r & 1) * 8;
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int j = 0; j < 8; ++j)
                *reinterpret_cast<double *>(t_wr + ((((r >> 1) ^ (j >> 1)) << 1) * 8) + (m * 8 + j) * 64) = x[m][j];
    }
    __syncwarp();
    {
        const unsigned char *t_rd = work_b + u * (kP3TU * 8) + r * 64;
#pragma unroll
        for (int m = 0; m < 2; ++m) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double2 v = *reinterpret_cast<const double2 *>(t_rd + ((k ^ (r >> 1)) << 4) + m * 512);
                x[m][2 * k] = v.x;
                x[m][2 * k + 1] = v.y;
            }
            dct2_8(x[m]);
        }
    }
    __syncwarp();
    // numpy broadcasting: the single luma channel is quantised with all three tables (patchquant.py:59)
    const int nch = (chroma_twice || a.och == 2) ? 2 : 3;
    const double *rt_l = s_rt + r, *t_l = s_t + r;
    unsigned zzo[2];                                    // scan positions of raster (v, r), v = 0..7, one byte each
    zzo[0] = ZZ_ORDER[r] | (ZZ_ORDER[8 + r] << 8) | (ZZ_ORDER[16 + r] << 16) | (ZZ_ORDER[24 + r] << 24);
    zzo[1] = ZZ_ORDER[32 + r] | (ZZ_ORDER[40 + r] << 8) | (ZZ_ORDER[48 + r] << 16) | (ZZ_ORDER[56 + r] << 24);
    unsigned char *stage = work_b + u * (kStageUF * 4);
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        if (m == 1) { bulk_wait_read0(); __syncwarp(); }     // round 0's stores have drained the staging area
        for (int ch = 0; ch < nch; ++ch) {
            QuantGuard qg;
            int qv[8];
#pragma unroll
            for (int v = 0; v < 8; ++v) qv[v] = qg.q(x[m][v], rt_l[ch * 64 + v * 8]);
            if (__builtin_expect(qg.risky(), 0)) {
#pragma unroll
                for (int v = 0; v < 8; ++v) qv[v] = quantize_exact_f64(x[m][v], t_l[ch * 64 + v * 8]);
            }
#pragma unroll
            for (int v = 0; v < 8; ++v) {
                const unsigned pos = (zzo[v >> 2] >> (8 * (v & 3))) & 255u;
                *reinterpret_cast<int *>(stage + pos * 4 + ch * 256) = qv[v];
                if (ch == 1 && nch == 2) *reinterpret_cast<int *>(stage + pos * 4 + 512) = qv[v];   // channel 2 repeats channel 1
            }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane < 4 && lane + 4 * m < nb)                // lane u stores the och scan blocks of image block u + 4m
            bulk_s2g(a.zz + ((tl.frame * a.Hp + tl.by0 + brow) * (int64_t)a.Wp + tl.bx0 + b0 + lane + 4 * m) * (64 * a.och),
                     work_s + lane * (kStageUF * 4), 256u * (uint32_t)a.och);
        bulk_commit();                                    // every lane commits (possibly empty) groups: counts stay in step
        if (a.zr_masks) {                                 // scan block j = 3 u + ch of this round (ch < och are stored)
            unsigned long long mk;
            zr_masks_from_staging<12>(work_b, [](int j) { return (j / 3) * (kStageUF * 4) + (j % 3) * 256; }, lane, mk);
            const int u_ = lane / 3, ch_ = lane - 3 * u_;
            if (lane < 12 && u_ + 4 * m < nb && ch_ < a.och) {
                const int64_t sblk = ((tl.frame * a.Hp + tl.by0 + brow) * (int64_t)a.Wp + tl.bx0 + b0 + u_ + 4 * m) * a.och + ch_;
                a.zr_masks[sblk] = mk;
                a.zr_counts[sblk] = zr_block_count(mk);
            }
        }
    }
}

// PF: after the search, code the tile's blocks (MC + residual + DCT + quantise + zig-zag) from the staged bytes
template <typename T, int G, int PC, bool PF = false>
__global__ void __launch_bounds__(PF ? kMePfThreads : (PC ? kMeIntMaxThreads : kMeThreads), PF ? kMePfCtas : (PC ? 2 : 3)) k_me_int(const MeArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned *s_u = reinterpret_cast<unsigned *>(smem_raw);                   // [R][P4]
    unsigned *s_hs = reinterpret_cast<unsigned *>(smem_raw + a.hs_off);       // [8*nseg+7][P4]
    unsigned *s_b = reinterpret_cast<unsigned *>(smem_raw + a.b_off);         // [R][pw]
    unsigned *s_cur = reinterpret_cast<unsigned *>(smem_raw + a.cur_off);     // [tby*tbx][kCurPitch]
    __shared__ unsigned long long s_best[64];
    __shared__ unsigned s_c2[64];
    const MeTile tl = me_tile(a);
    const int sr = a.sr, span = a.span, P4 = PC ? PC : a.P;
    const int H = (int)a.H, W = (int)a.W;
    const int nblk = tl.nby * tl.nbx;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int last_col = tl.nbx != a.tbx;
    const FastDiv d_nbx(a.m_nbx[last_col]);

    if (tid < 64) s_best[tid] = ~0ull;
    __shared__ double s_qt[PF ? 384 : 1];                                     // PF: fl(1/t) [192] and t [192] of the three tables
    bool chroma_twice = false;
    if (PF) {
        if (tid < 192) {
            const double t = load_table_elem(a.table, a.table_dtype, tid);
            s_qt[192 + tid] = t;
            s_qt[tid] = __drcp_rn(t);
        }
        bool same = true;                                                     // [lum, chrom, chrom]: the third copy IS the second
        if (tid < 64)
            same = __double_as_longlong(load_table_elem(a.table, a.table_dtype, 64 + tid)) ==
                   __double_as_longlong(load_table_elem(a.table, a.table_dtype, 128 + tid));
        chroma_twice = __syncthreads_and(same) != 0;
    }
    if (!me_int_prepare<T, PC, kStageUnroll>(a, tl, s_u, s_hs, s_b, s_cur, s_c2)) return;

    // ---- main loop ------------------------------------------------------------------------------
    const bool small = span * span <= 512;                                    // ssd < 2^22, index < 2^9: key fits 32 bits
    unsigned *s_best32 = reinterpret_cast<unsigned *>(s_best);
    const int total = nblk * a.ntpb;
    const FastDiv d_ntpb(a.m_ntpb), d_span(a.m_span);
    // every warp runs the same number of rounds: lanes without a task (or whose dx leaves the frame) carry
    // an empty result into the warp-level argmin, so that one lane per block issues the shared atomic
    for (int tbase = tid & ~31; tbase < total; tbase += nthr) {
        const int task = tbase + (tid & 31);
        const int blk = d_ntpb.div(min(task, total - 1)), rem = task - blk * a.ntpb;
        const int g = d_span.div(rem), dxi = rem - g * span;
        const int brow = d_nbx.div(blk), b = blk - brow * tl.nbx;
        const int gx = 8 * (tl.bx0 + b) + dxi - sr;
        unsigned best_ssd = 0xffffffffu, best_idx = 0xffffffffu;
        if (task < total && gx >= 0 && gx + 8 <= W) {                         // motion.py:41-43 (x bound)
            const uint2 *cb = reinterpret_cast<const uint2 *>(s_cur + (brow * a.tbx + b) * kCurPitch);
            uint2 c[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) c[i] = cb[i];
            unsigned acc[G];
#pragma unroll
            for (int gg = 0; gg < G; ++gg) acc[gg] = 0u;
            const int off = (8 * brow + g * G) * P4 + 8 * b + dxi;            // window row of dy0, byte column of dx
            const unsigned *up = s_u + off;
#pragma unroll
            for (int rr = 0; rr < G + 7; ++rr) {
                const unsigned r0 = up[rr * P4], r1 = up[rr * P4 + 4];
#pragma unroll
                for (int gg = 0; gg < G; ++gg) {
                    const int i = rr - gg;                                    // row of the block for candidate gg
                    if (i >= 0 && i < 8) {
                        acc[gg] = __dp4a(c[i].x, r0, acc[gg]);
                        acc[gg] = __dp4a(c[i].y, r1, acc[gg]);
                    }
                }
            }
            // key = ssd * 16 + gg (ssd < 2^22, gg < 16) = (16 c2 + gg) + 16 S - 32 acc; its minimum is the first
            // best candidate of the task.  Candidates gg in [lo, hi) are inside the search range and the frame
            // (motion.py:41-43, y bound); almost every task has all G of them.
            const int gyb = 8 * (tl.by0 + brow) - sr + g * G;
            const int lo = max(0, -gyb), hi = min(min(G, span - g * G), H - 7 - gyb);
            const unsigned c2 = s_c2[blk];
            const unsigned *sp = s_hs + off;
            unsigned key = 0xffffffffu;
            if (lo == 0 && hi == G) {
#pragma unroll
                for (int gg = 0; gg < G; ++gg) key = min(key, (c2 + gg) + sp[gg * P4] - (acc[gg] << 5));
            } else {
#pragma unroll
                for (int gg = 0; gg < G; ++gg)
                    if (gg >= lo && gg < hi) key = min(key, (c2 + gg) + sp[gg * P4] - (acc[gg] << 5));
            }
            if (key != 0xffffffffu) {
                best_ssd = key >> 4;
                best_idx = (unsigned)((g * G + (int)(key & 15u)) * span + dxi);   // motion.py:55
            }
        }
        // lexicographic (ssd, index) argmin: when the whole warp works on one block (the usual case for wide
        // searches) two warp reductions and ONE shared atomic; otherwise one atomic per lane
        if (__all_sync(0xffffffffu, blk == __shfl_sync(0xffffffffu, blk, 0))) {
            const unsigned m = __reduce_min_sync(0xffffffffu, best_ssd);
            const unsigned mi = __reduce_min_sync(0xffffffffu, best_ssd == m ? best_idx : 0xffffffffu);
            if ((tid & 31) == 0) { best_ssd = m; best_idx = mi; } else best_ssd = 0xffffffffu;
        }
        if (best_ssd != 0xffffffffu) {
            if (small) atomicMin(s_best32 + blk, (best_ssd << 9) | best_idx);
            else atomicMin(s_best + blk, ((unsigned long long)best_ssd << 32) | best_idx);
        }
    }
    __syncthreads();
    if (tid < nblk) {
        const int blk = tid, brow = d_nbx.div(blk), b = blk - brow * tl.nbx;
        const int64_t idx = small ? (int64_t)(s_best32[blk] & 511u) : (int64_t)(s_best[blk] & 0xffffffffull);
        a.mv[(tl.frame * a.Hp + tl.by0 + brow) * (int64_t)a.Wp + tl.bx0 + b] = idx;
    }
    if (PF) {
        // U and S are dead (every warp is past the barrier above): their space holds one WORK buffer per warp
        const int warp = tid >> 5, nwarps = nthr >> 5;
        const int gpr = (tl.nbx + 7) >> 3, ngroups = tl.nby * gpr;            // groups of eight blocks per block row
        unsigned char *work_b = smem_raw + warp * kPfWork;
        for (int grp = warp; grp < ngroups; grp += nwarps) {
            const int brow = grp / gpr, b0 = 8 * (grp - brow * gpr);
            pf_forward_group<PC>(a, tl, brow, b0, min(8, tl.nbx - b0), s_b, s_cur, s_best32, work_b, s_qt, s_qt + 192, chroma_twice);
        }
        bulk_wait_all0();
    }
}

// ================================================================================================
// integer kernel, +-16, cross term on the tensor cores (mma.sync.m16n8k32.u8)
// ================================================================================================
// The cross term X[dy][dx] = sum_{i,j} c[i][j] r[y+dy+i][x+dx+j] of ONE block is a small GEMM once the vertical
// shift is folded into a Toeplitz operand:  with t a window row and j a block column,
//     A[m][(t, j)] = c[t - m][j]   (zero unless 0 <= t - m < 8)        m = 16 vertical shifts of an M-tile
//     B[(t, j)][n] = r[y0 + t][x0 + n + j]                             n = 8 horizontal shifts of an N-tile
//     C[m][n]      = sum_t sum_j c[t - m][j] r[y0 + t][x0 + n + j] = X at (dy, dx) = (y0 + m, x0 + n).
// K runs over 24 window rows x 8 columns = six k-steps of 32; a k-step's B fragment is exactly two words of the
// unaligned-word view U (bytes x..x+3 of a window row, any x), and its A fragment four words of the block stored
// with 15 zero rows above and below -- no operand is ever materialised.  A fragments do not depend on the M-tile,
// and the B fragment of (M-tile mt, k-step ks) is that of 4-row group 4 mt + ks, so a warp walks the ten row
// groups of its block's 40-row window once: 24 + 100 shared-memory loads and 70 mma per block (M-tiles dy -16..-1,
// 0..15 and the single row +16, which needs two k-steps only; N-tiles dx -16..23, the last one for +16 alone).
// 35 % of K and 66 % of M x N are useful, which still is 2.6 times the useful multiply-adds per clock of dp4a
// (mma.sync u8: 1935 MAC/clk/SM measured, dp4a 256).  sum(r^2), sum(c^2) and the staging are the phases the dp4a
// kernel uses (me_int_prepare); a warp owns its block's whole search, so the argmin needs no atomics:
// per-thread strict minimum in ascending index order, then two warp reductions (ssd, then index among equals)
// = the reference's first strict minimum in (dy, dx) raster order (motion.py:35-51).
constexpr int kMmaSr = 16, kMmaSpan = 33;
constexpr int kMmaPitch = 112;       // U / HS pitch in words for the 4 x 8-block tile: 96 used; == 16 (mod 32), so the two
                                     // window rows a B fragment touches fall on disjoint banks
constexpr int kMmaTby = 4, kMmaTbx = 8;
constexpr int kCpPitch = 80;         // words per zero-padded current block: 39 rows (i = -15 .. 23) x 2 words + 2

__device__ __forceinline__ void mma_u8(int (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <typename T>
__global__ void __launch_bounds__(kMeThreads, 2) k_me_mma16(const MeArgs a) {
    constexpr int P4 = kMmaPitch;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned *s_u = reinterpret_cast<unsigned *>(smem_raw);                   // [64][P4]
    unsigned *s_hs = reinterpret_cast<unsigned *>(smem_raw + a.hs_off);       // [71][P4]
    unsigned *s_b = reinterpret_cast<unsigned *>(smem_raw + a.b_off);         // [64][pw]
    unsigned *s_cur = reinterpret_cast<unsigned *>(smem_raw + a.cur_off);     // [32][kCurPitch]
    unsigned *s_cp = reinterpret_cast<unsigned *>(smem_raw + a.part_off);     // [32][kCpPitch]
    __shared__ unsigned s_c2[64];
    const MeTile tl = me_tile(a);
    const int tid = threadIdx.x;
    // S is stored as 128 S + x (x = window column < 112): a candidate's key 128 ssd + x orders the candidates of one
    // window row by (ssd, dx) with nothing but a multiply-add and a minimum per candidate
    if (!me_int_prepare<T, P4, 4, 7, true>(a, tl, s_u, s_hs, s_b, s_cur, s_c2)) return;

    // current blocks with 15 zero rows above and 16 below: row r of a slot holds block row r - 15
    for (int idx = tid; idx < kMmaTby * kMmaTbx * kCpPitch; idx += kMeThreads) {
        const int slot = idx / kCpPitch, w = idx - slot * kCpPitch;
        const int i = (w >> 1) - 15;
        s_cp[idx] = ((unsigned)i < 8u) ? s_cur[slot * kCurPitch + i * 2 + (w & 1)] : 0u;
    }
    __syncthreads();

    const int lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tq = lane & 3, tlo = tq >> 1, h = tq & 1;
    const int H = (int)a.H, W = (int)a.W;
    const int nblk = tl.nby * tl.nbx;
    const FastDiv d_nbx(a.m_nbx[tl.nbx != a.tbx]);
    for (int blk = warp; blk < nblk; blk += kMeWarps) {
        const int brow = d_nbx.div(blk), b = blk - brow * tl.nbx;
        const int slot = brow * kMmaTbx + b;
        // ---- A fragments: rows g / g + 8 of the Toeplitz operand, k-steps 0..5 ----
        const unsigned *cp = s_cp + slot * kCpPitch + ((tlo - g + 15) * 2 + h);
        unsigned af[6][4];
#pragma unroll
        for (int ks = 0; ks < 6; ++ks) {
            af[ks][0] = cp[(4 * ks) * 2];
            af[ks][1] = cp[(4 * ks - 8) * 2];
            af[ks][2] = cp[(4 * ks + 2) * 2];
            af[ks][3] = cp[(4 * ks + 2 - 8) * 2];
        }
        int acc[3][5][4];
#pragma unroll
        for (int mt = 0; mt < 3; ++mt)
#pragma unroll
            for (int nt = 0; nt < 5; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0;
        const unsigned *ub = s_u + (8 * brow + tlo) * P4 + 8 * b + g + 4 * h;
#pragma unroll
        for (int rg = 0; rg < 10; ++rg) {
            unsigned b0[5], b1[5];
#pragma unroll
            for (int nt = 0; nt < 5; ++nt) {
                b0[nt] = ub[(4 * rg) * P4 + 8 * nt];
                b1[nt] = ub[(4 * rg + 2) * P4 + 8 * nt];
            }
#pragma unroll
            for (int mt = 0; mt < 3; ++mt) {
                const int ks = rg - 4 * mt;
                if (ks >= 0 && ks < (mt == 2 ? 2 : 6)) {
#pragma unroll
                    for (int nt = 0; nt < 5; ++nt) mma_u8(acc[mt][nt], af[ks], b0[nt], b1[nt]);
                }
            }
        }
        // ---- argmin.  key = 128 (S - 2 X) + x per candidate (c2 is common to the block and joins at the end);
        //      rows ascend within a thread, so a later row replaces the best only on a strictly smaller ssd ----
        const int gy0 = 8 * (tl.by0 + brow) - kMmaSr, gx0 = 8 * (tl.bx0 + b) - kMmaSr;     // frame position of candidate (0, 0)
        const bool interior = gy0 >= 0 && gy0 + 2 * kMmaSr + 8 <= H && gx0 >= 0 && gx0 + 2 * kMmaSr + 8 <= W;   // warp-uniform
        unsigned best = 0xffffffffu, bidx = 0xffffffffu;                  // best = 128 (S - 2X) of the winner, bidx its index
#pragma unroll
        for (int mt = 0; mt < 3; ++mt)
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                if (mt == 2 && hf == 1) continue;
                const int dyi = 16 * mt + 8 * hf + g;
                const unsigned *sp = s_hs + (8 * brow + dyi) * P4 + 8 * b + 2 * tq;
                int rk = 0x7fffffff;                                      // min over the row's candidates of 128 (S - 2X) + x (signed: S < 2X happens)
                if (interior) {
                    if (mt < 2 || g == 0) {
#pragma unroll
                        for (int nt = 0; nt < 4; ++nt) {
                            const uint2 sv = *reinterpret_cast<const uint2 *>(sp + 8 * nt);
                            rk = min(rk, (int)sv.x - (acc[mt][nt][2 * hf] << 8));
                            rk = min(rk, (int)sv.y - (acc[mt][nt][2 * hf + 1] << 8));
                        }
                        const int k32 = (int)sp[32] - (acc[mt][4][2 * hf] << 8);                 // dx = +16: lanes tq == 0 only
                        rk = min(rk, tq == 0 ? k32 : 0x7fffffff);
                    }
                } else {
                    const bool row_ok = dyi < kMmaSpan && gy0 + dyi >= 0 && gy0 + dyi + 8 <= H;
#pragma unroll
                    for (int nt = 0; nt < 5; ++nt) {
                        uint2 sv = make_uint2(0u, 0u);
                        if (row_ok) sv = *reinterpret_cast<const uint2 *>(sp + 8 * nt);
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int sx = 8 * nt + 2 * tq + e;
                            const bool ok = row_ok && sx < kMmaSpan && gx0 + sx >= 0 && gx0 + sx + 8 <= W;
                            const int key = (int)(e ? sv.y : sv.x) - (acc[mt][nt][2 * hf + e] << 8);
                            if (ok) rk = min(rk, key);
                        }
                    }
                }
                // 128 c2 joins here: ssd = c2 + S - 2X >= 0, so the full key is a plain unsigned number again
                if (rk != 0x7fffffff) {
                    const unsigned full = (unsigned)rk + s_c2[blk];
                    const unsigned ssd7 = full & ~127u;
                    if (ssd7 < best) { best = ssd7; bidx = (unsigned)(dyi * kMmaSpan) + ((full & 127u) - (unsigned)(8 * b)); }
                }
            }
        const unsigned m = __reduce_min_sync(0xffffffffu, best);
        const unsigned mi = __reduce_min_sync(0xffffffffu, best == m ? bidx : 0xffffffffu);
        if (lane == 0) a.mv[(tl.frame * a.Hp + tl.by0 + brow) * (int64_t)a.Wp + tl.bx0 + b] = (int64_t)mi;
    }
}

static int sm_count(int device) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    return sms;
}

// ================================================================================================
// integer-DTYPE frames: numpy's own wrap-around arithmetic
// ================================================================================================
// For uint8 / int16 / int32 frames `(block - ref_block) ** 2` is evaluated IN THAT DTYPE (a uint8 difference
// wraps mod 256 and so does its square; 255**2 is -511 in int16) and np.sum then accumulates in uint64 / int64
// (motion.py:46).  The reference's vectors on such inputs are whatever this arithmetic yields, so it is replayed
// literally: one warp per block, lanes over candidates, T-typed difference and square, 64-bit accumulation,
// lexicographic (ssd, index) argmin in the accumulator's signedness.  Not a fast path: uint8 / int16 frames are a
// corner of the reference's behaviour, the codecs hand over floats.
template <typename T> struct WrapAcc { using type = long long; };
template <> struct WrapAcc<unsigned char> { using type = unsigned long long; };

template <typename T>
__global__ void __launch_bounds__(256) k_me_wrap(const T *__restrict__ ref, const T *__restrict__ cur, int64_t n, int H, int W,
                                                 int64_t ref_fs, int64_t cur_fs, int sr, int64_t *__restrict__ mv) {
    using A = typename WrapAcc<T>::type;
    const int lane = threadIdx.x & 31;
    const int Hp = H / 8, Wp = W / 8, span = 2 * sr + 1;
    const int64_t nblocks = n * Hp * (int64_t)Wp;
    for (int64_t blk = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); blk < nblocks; blk += (int64_t)gridDim.x * 8) {
        const int bx = (int)(blk % Wp), by = (int)((blk / Wp) % Hp);
        const int64_t f = blk / ((int64_t)Wp * Hp);
        const T *c = cur + f * cur_fs + (int64_t)(8 * by) * W + 8 * bx, *r0 = ref + f * ref_fs;
        bool have = false;
        A best = 0;
        int bidx = sr * span + sr;                                              // default (0, 0): motion.py:31-33
        for (int cand = lane; cand < span * span; cand += 32) {
            const int dyi = cand / span, dxi = cand - dyi * span;
            const int yy = 8 * by + dyi - sr, xx = 8 * bx + dxi - sr;
            if (yy < 0 || yy + 8 > H || xx < 0 || xx + 8 > W) continue;        // motion.py:41-43
            const T *r = r0 + (int64_t)yy * W + xx;
            A ssd = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const T d = (T)((unsigned)c[i * W + j] - (unsigned)r[i * W + j]);   // wraps like numpy's T - T
                    ssd += (A)(T)((unsigned)d * (unsigned)d);                           // T ** 2 wraps, the sum does not
                }
            if (!have || ssd < best) { best = ssd; bidx = cand; have = true; }  // candidates ascend per lane
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const A os = __shfl_xor_sync(0xffffffffu, best, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bidx, off);
            const bool oh = __shfl_xor_sync(0xffffffffu, (int)have, off) != 0;
            if (oh && (!have || os < best || (os == best && oi < bidx))) { best = os; bidx = oi; have = true; }
        }
        if (lane == 0) mv[blk] = bidx;
    }
}

cudaError_t launch_me_wrap(int device, cudaStream_t st, const void *ref, const void *cur, int dtype, int64_t n, int64_t H,
                           int64_t W, int64_t ref_fs, int64_t cur_fs, int sr, int64_t *mv) {
    const int64_t nblocks = n * (H / 8) * (W / 8);
    if (nblocks == 0) return cudaSuccess;
    int64_t grid = (nblocks + 7) / 8;
    const int64_t cap = (int64_t)sm_count(device) * 16;
    if (grid > cap) grid = cap;
    switch (dtype) {
        case IVC_U8: k_me_wrap<unsigned char><<<(unsigned)grid, 256, 0, st>>>((const unsigned char *)ref, (const unsigned char *)cur, n, (int)H, (int)W, ref_fs, cur_fs, sr, mv); break;
        case IVC_I16: k_me_wrap<short><<<(unsigned)grid, 256, 0, st>>>((const short *)ref, (const short *)cur, n, (int)H, (int)W, ref_fs, cur_fs, sr, mv); break;
        case IVC_I32: k_me_wrap<int><<<(unsigned)grid, 256, 0, st>>>((const int *)ref, (const int *)cur, n, (int)H, (int)W, ref_fs, cur_fs, sr, mv); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
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

// choose the CTA tile (at most 64 blocks) and the shared-memory layout
static size_t me_geometry(MeArgs &a, int64_t n_frames, int64_t H, int64_t W, int sr, int elem, int pitch_quantum,
                          int pitch_skew, size_t budget, int64_t min_ctas, int extra_per_warp = 0) {
    a.Hp = (int)(H / 8); a.Wp = (int)(W / 8); a.sr = sr; a.span = 2 * sr + 1;
    a.ngrp = (a.span + kMeG - 1) / kMeG;
    a.ntpb = a.ngrp * a.span;
    // largest tile that fits the shared-memory budget AND still yields enough CTAs to fill the GPU a few
    // times over (single small frames -- QCIF, one 1080p frame of a closed loop -- get smaller tiles)
    static const int shapes[][2] = {{4, 16}, {2, 16}, {2, 8}, {1, 8}, {1, 4}, {1, 2}, {1, 1}};
    size_t smem = 0;
    if (const char *env = getenv("IVC_ME_MIN_CTAS")) min_ctas = atoll(env);          // developer override
    for (auto &s : shapes) {
        a.tby = s[0]; a.tbx = s[1];
        const int64_t ctas = n_frames * ((a.Hp + a.tby - 1) / a.tby) * (int64_t)((a.Wp + a.tbx - 1) / a.tbx);
        if (ctas < min_ctas && !(s[0] == 1 && s[1] == 1) && s[0] * s[1] > 8) continue;
        a.R = 8 * (a.tby - 1) + a.ngrp * kMeG + 7;                 // >= 8*tby + 2*sr, covers the last dy-group
        a.Wc = 8 * a.tbx + 2 * sr;
        a.P = ((a.Wc + pitch_quantum - 1) / pitch_quantum) * pitch_quantum + pitch_skew;
        a.cur_off = (int)((((size_t)a.R * a.P * elem) + 15) & ~(size_t)15);
        a.npieces = 32;                                            // one (ssd, index) slot per block and lane
        a.part_off = (int)(((size_t)a.cur_off + (size_t)a.tby * a.tbx * kExactCurPitch * elem + 15) & ~(size_t)15);
        smem = (size_t)a.part_off + (size_t)64 * 32 * (elem + 4);
        a.work_off = (int)((smem + 127) & ~(size_t)127);
        if (extra_per_warp) smem = (size_t)a.work_off + (size_t)kMeWarps * extra_per_warp;
        if (smem <= budget) break;
    }
    a.tiles_y = (a.Hp + a.tby - 1) / a.tby;
    a.tiles_x = (a.Wp + a.tbx - 1) / a.tbx;
    a.m_ntpb = fastdiv_magic(a.ntpb); a.m_span = fastdiv_magic(a.span);
    a.m_nbx[0] = fastdiv_magic(a.tbx); a.m_nbx[1] = fastdiv_magic(a.Wp - (a.tiles_x - 1) * a.tbx);
    return smem;
}

// launch over a 2-D grid (x = tile, y = frame), in chunks of at most 65535 frames
template <typename K>
static cudaError_t me_launch_chunks(K kernel, MeArgs a, int elem, size_t smem, cudaStream_t st, int threads = kMeThreads) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int64_t tiles = (int64_t)a.tiles_y * a.tiles_x;
    if (tiles == 0 || a.n == 0) return cudaSuccess;
    if (tiles > 2147483647LL) return cudaErrorInvalidValue;
    const int64_t n = a.n;
    for (int64_t f0 = 0; f0 < n; f0 += 65535) {
        const int64_t nf = n - f0 < 65535 ? n - f0 : 65535;
        kernel<<<dim3((unsigned)tiles, (unsigned)nf), threads, smem, st>>>(a);
        a.ref = (const char *)a.ref + 65535 * a.ref_fs * elem;
        a.cur = (const char *)a.cur + 65535 * a.cur_fs * elem;
        a.mv += 65535 * (int64_t)a.Hp * a.Wp;
    }
    return cudaGetLastError();
}

// shared-memory layout of k_me_exact2 on top of me_geometry's (window + blocks in float64): float copies, per-warp
// candidate scores, and -- for the fused step -- one WORK buffer per warp
static size_t me_geometry2(MeArgs &a, int64_t n_frames, int64_t H, int64_t W, int sr, size_t budget, int64_t min_ctas, bool step) {
    a.Hp = (int)(H / 8); a.Wp = (int)(W / 8); a.sr = sr; a.span = 2 * sr + 1;
    a.ngrp = (a.span + kMeG - 1) / kMeG;
    a.ntpb = a.ngrp * a.span;
    static const int shapes[][2] = {{4, 16}, {2, 16}, {2, 8}, {1, 8}, {1, 4}, {1, 2}, {1, 1}};
    size_t smem = 0;
    if (const char *env = getenv("IVC_ME_MIN_CTAS")) min_ctas = atoll(env);          // developer override
    for (auto &s : shapes) {
        a.tby = s[0]; a.tbx = s[1];
        const int64_t ctas = n_frames * ((a.Hp + a.tby - 1) / a.tby) * (int64_t)((a.Wp + a.tbx - 1) / a.tbx);
        if (ctas < min_ctas && !(s[0] == 1 && s[1] == 1) && s[0] * s[1] > 8) continue;
        a.R = 8 * (a.tby - 1) + a.ngrp * kMeG + 7;                 // >= 8*tby + 2*sr, covers the last dy-group
        a.Wc = 8 * a.tbx + 2 * sr;
        a.P = a.Wc + 2;                                            // even: rows start 16-byte aligned (bulk copies); 138 for 16 blocks at +-4
        const size_t win = (size_t)a.R * a.P, blocks = (size_t)a.tby * a.tbx;
        a.cur_off = (int)((win * 8 + 15) & ~(size_t)15);
        a.pwl = ((a.Wc + 3 - 4 + 31) / 32) * 32 + 4;               // PU: words per row of the unaligned view, == 4 (mod 32): the dy-groups of a warp on disjoint banks
        a.pw = a.pwl / 4 + 2;                                      // PW: packed words per row, two spare words at the end
        a.m_q4 = fastdiv_magic(a.pwl / 4); a.m_pwl = fastdiv_magic((a.Wc + 3) / 4); a.m_p4 = fastdiv_magic(a.Wc);
        a.win32_off = (int)((a.cur_off + blocks * kExactCurPitch * 8 + 15) & ~(size_t)15);       // U view
        a.b_off = (int)((a.win32_off + (size_t)a.R * a.pwl * 4 + 15) & ~(size_t)15);             // packed window
        a.cur32_off = (int)((a.b_off + (size_t)a.R * a.pw * 4 + 15) & ~(size_t)15);              // quantised blocks
        a.acand_off = (int)((a.cur32_off + blocks * kCurPitch * 4 + 15) & ~(size_t)15);
        a.acand_pitch = (a.span * a.span + 3) & ~3;
        smem = (size_t)a.acand_off + (size_t)kMeWarps * a.acand_pitch * 4;
        a.work_off = (int)(((size_t)a.win32_off + 127) & ~(size_t)127);          // over the search's tables: dead when the step starts
        if (step) smem = std::max(smem, (size_t)a.work_off + (size_t)kMeWarps * kX2Work);
        if (smem <= budget) break;
    }
    a.tiles_y = (a.Hp + a.tby - 1) / a.tby;
    a.tiles_x = (a.Wp + a.tbx - 1) / a.tbx;
    a.m_ntpb = fastdiv_magic(a.ntpb); a.m_span = fastdiv_magic(a.span);
    a.m_nbx[0] = fastdiv_magic(a.tbx); a.m_nbx[1] = fastdiv_magic(a.Wp - (a.tiles_x - 1) * a.tbx);
    return smem;
}

static bool me_exact_v1() {          // A/B switch: IVC_ME_EXACT_V1=1 selects the first-generation exact kernel (every candidate in FP64)
    const char *e = getenv("IVC_ME_EXACT_V1");
    return e && e[0] == '1';
}

// The fused closed-loop step (decode = "luma"): exact search, then the warp codes and reconstructs its own blocks.
cudaError_t launch_pframe_step(int device, cudaStream_t st, const void *ref, const void *cur, int64_t n, int64_t H, int64_t W,
                               int sr, const void *table, int table_dtype, int out_channels, int64_t *mv, int32_t *zz,
                               double *recon) {
    MeArgs a;
    a.ref = ref; a.cur = cur; a.n = n; a.H = H; a.W = W; a.ref_fs = H * W; a.cur_fs = H * W; a.mv = mv;
    a.flag = nullptr; a.run_if = 0; a.check = 0;
    a.table = table; a.table_dtype = table_dtype; a.zz = zz; a.och = out_channels; a.recon = recon;
    a.zr_counts = nullptr; a.zr_masks = nullptr;
    if (n > 65535) return cudaErrorInvalidValue;                               // one launch: y = frame
    if (2 * sr + 1 <= 33 && !me_exact_v1()) {
        const size_t smem2 = me_geometry2(a, n, H, W, sr, (sr == 4 ? 73 : 113) * 1024, 2 * 2 * (int64_t)sm_count(device), true);     // +-4: three CTAs per SM
        if (smem2 <= 227 * 1024)
            return (a.span == 9 && a.P == 138) ? me_launch_chunks(k_me_exact2<true, 9, 138>, a, 8, smem2, st)
                                              : me_launch_chunks(k_me_exact2<true>, a, 8, smem2, st);
    }
    const size_t smem = me_geometry(a, n, H, W, sr, 8, 32, 3, 113 * 1024, 2 * 2 * (int64_t)sm_count(device), kStepWork);
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    return me_launch_chunks(k_me_exact<double, true>, a, 8, smem, st);
}

cudaError_t launch_me_exact(int device, cudaStream_t st, const void *ref, const void *cur, bool f32, int64_t n,
                            int64_t H, int64_t W, int64_t ref_fs, int64_t cur_fs, int sr, int64_t *mv,
                            int *flag, int run_if) {
    MeArgs a;
    a.ref = ref; a.cur = cur; a.n = n; a.H = H; a.W = W; a.ref_fs = ref_fs; a.cur_fs = cur_fs; a.mv = mv;
    a.flag = flag; a.run_if = run_if; a.check = 0;
    if (!f32 && 2 * sr + 1 <= 33 && !me_exact_v1()) {                       // float64 frames: prefilter + exact survivors
        const size_t smem2 = me_geometry2(a, n, H, W, sr, (sr == 4 ? 73 : 113) * 1024, 2 * 2 * (int64_t)sm_count(device), false);
        if (smem2 <= 227 * 1024)
            return (a.span == 9 && a.P == 138) ? me_launch_chunks(k_me_exact2<false, 9, 138>, a, 8, smem2, st)
                                              : me_launch_chunks(k_me_exact2<false>, a, 8, smem2, st);
    }
    const size_t smem = me_geometry(a, n, H, W, sr, f32 ? 4 : 8, 32, 3, 113 * 1024, 2 * 2 * (int64_t)sm_count(device));
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    (void)device;
    if (f32) return me_launch_chunks(k_me_exact<float>, a, 4, smem, st);
    return me_launch_chunks(k_me_exact<double>, a, 8, smem, st);
}

static bool me_mma_enabled() {       // A/B switch for profiling: IVC_ME_MMA=0 selects the dp4a kernel at +-16
    const char *e = getenv("IVC_ME_MMA");
    return !(e && e[0] == '0');
}

// integer kernel: candidates per task for a search span (least padding of the last dy-group, then larger)
static int me_int_group(int span) {
    static const int gs[] = {11, 9, 5, 3};
    int best = 3, waste = 1 << 30;
    for (int g : gs) {
        const int w = (span + g - 1) / g * g - span;
        if (w < waste) { waste = w; best = g; }
    }
    return best;
}

static size_t me_int_geometry(MeArgs &a, int G, int pitch, int64_t n_frames, int64_t H, int64_t W, int sr, size_t budget,
                              int64_t min_ctas, int force_tby = 0, int force_tbx = 0) {
    a.Hp = (int)(H / 8); a.Wp = (int)(W / 8); a.sr = sr; a.span = 2 * sr + 1;
    a.ngrp = (a.span + G - 1) / G;
    a.ntpb = a.ngrp * a.span;
    static const int shapes[][2] = {{4, 16}, {2, 16}, {2, 8}, {1, 8}, {1, 4}, {1, 2}, {1, 1}};
    size_t smem = 0;
    for (auto &s : shapes) {
        a.tby = s[0]; a.tbx = s[1];
        if (force_tby) { a.tby = force_tby; a.tbx = force_tbx; }
        const int64_t ctas = n_frames * ((a.Hp + a.tby - 1) / a.tby) * (int64_t)((a.Wp + a.tbx - 1) / a.tbx);
        if (!force_tby && ctas < min_ctas && !(s[0] == 1 && s[1] == 1) && s[0] * s[1] > 8) continue;
        a.R = 8 * (a.tby - 1) + a.ngrp * G + 7;                    // covers the last (padded) dy-group
        a.Wc = 8 * a.tbx + 2 * sr;
        a.P = pitch ? pitch : (a.Wc + 3) & ~3;                     // U / HS pitch in words (one per byte column)
        a.pw = a.P / 4 + 2;                                        // packed words per row, two zero words at the end
        a.pwl = (a.Wc + 3) / 4;
        a.nseg = (a.R - 7 + 7) / 8;                                // S has R - 7 rows
        a.hs_off = a.R * a.P * 4;
        a.b_off = a.hs_off + (8 * a.nseg + 7) * a.P * 4;
        a.cur_off = (a.b_off + a.R * a.pw * 4 + 15) & ~15;
        smem = (size_t)a.cur_off + (size_t)a.tby * a.tbx * kCurPitch * 4;
        if (smem <= budget || force_tby) break;
    }
    a.tiles_y = (a.Hp + a.tby - 1) / a.tby;
    a.tiles_x = (a.Wp + a.tbx - 1) / a.tbx;
    const int last_nbx = a.Wp - (a.tiles_x - 1) * a.tbx;
    a.m_pwl = fastdiv_magic(a.pwl); a.m_q4 = fastdiv_magic(a.P / 4); a.m_p4 = fastdiv_magic(a.P);
    a.m_ntpb = fastdiv_magic(a.ntpb); a.m_span = fastdiv_magic(a.span);
    a.m_nbx[0] = fastdiv_magic(a.tbx); a.m_nbx[1] = fastdiv_magic(last_nbx);
    a.m_cww[0] = fastdiv_magic(2 * a.tbx); a.m_cww[1] = fastdiv_magic(2 * last_nbx);
    a.ipr[0] = (8 * a.tbx + 2 * sr + 7) / 8; a.ipr[1] = (8 * last_nbx + 2 * sr + 7) / 8;
    a.m_ipr[0] = fastdiv_magic(a.ipr[0]); a.m_ipr[1] = fastdiv_magic(a.ipr[1]);
    return smem;
}

bool me_pf_fusable(int dtype, int sr) { return sr == 4 && (dtype == IVC_F64 || dtype == IVC_U8); }

cudaError_t launch_me_int(int device, cudaStream_t st, const void *ref, const void *cur, int dtype, int64_t n,
                          int64_t H, int64_t W, int64_t ref_fs, int64_t cur_fs, int sr, int64_t *mv, int *flag,
                          int check, const void *pf_table, int pf_table_dtype, int32_t *pf_zz, int pf_och,
                          int32_t *zr_counts, uint64_t *zr_masks) {
    const bool f32 = dtype == IVC_F32, u8 = dtype == IVC_U8;
    if (u8) check = 0;                                                        // uint8 planes are integer-valued by construction
    MeArgs a;
    a.table = pf_table; a.table_dtype = pf_table_dtype; a.zz = pf_zz; a.och = pf_och;
    a.zr_counts = zr_counts; a.zr_masks = (unsigned long long *)zr_masks;
    a.ref = ref; a.cur = cur; a.n = n; a.H = H; a.W = W; a.ref_fs = ref_fs; a.cur_fs = cur_fs; a.mv = mv;
    a.flag = flag; a.run_if = 0; a.check = check;
    if (sr == kMmaSr && me_mma_enabled()) {
        // +-16: the cross term on the tensor cores (k_me_mma16); same staging / sum-of-squares phases
        const size_t sm = me_int_geometry(a, 11, kMmaPitch, n, H, W, sr, 227 * 1024, 0, kMmaTby, kMmaTbx);
        a.part_off = (int)((sm + 15) & ~(size_t)15);
        const size_t smem_mma = (size_t)a.part_off + (size_t)kMmaTby * kMmaTbx * kCpPitch * 4;
        if (H * W >= 2147483647LL || smem_mma > 113 * 1024) return cudaErrorInvalidValue;
        const int el = u8 ? 1 : f32 ? 4 : 8;
        if (u8)
            a.vec = sr % 4 == 0 && ((uintptr_t)ref & 3) == 0 && ((uintptr_t)cur & 3) == 0 && ref_fs % 4 == 0 && cur_fs % 4 == 0;
        else
            a.vec = sr % (16 / el) == 0 && ((uintptr_t)ref & 15) == 0 && ((uintptr_t)cur & 15) == 0 &&
                    (ref_fs * el) % 16 == 0 && (cur_fs * el) % 16 == 0;
        cudaError_t e0;
        if (check && (e0 = cudaMemsetAsync(flag, 0, sizeof(int), st)) != cudaSuccess) return e0;
        return u8 ? me_launch_chunks(k_me_mma16<unsigned char>, a, 1, smem_mma, st, kMeThreads)
             : f32 ? me_launch_chunks(k_me_mma16<float>, a, 4, smem_mma, st, kMeThreads)
                   : me_launch_chunks(k_me_mma16<double>, a, 8, smem_mma, st, kMeThreads);
    }
    const int G = me_int_group(2 * sr + 1);
    // common cases get a compile-time pitch: 16-block-wide tiles at +-4 (136) and up to +-16 (160)
    const int pitch = (G == 9 && sr <= 4) ? 136 : (G == 9 && sr <= 8) ? 144 : (G == 11 && sr <= 16) ? 160 : 0;
    int f_tby = 0, f_tbx = 0, f_thr = 0;                                       // developer overrides: IVC_ME_TILE="tby,tbx", IVC_ME_THREADS
    if (const char *e = getenv("IVC_ME_TILE")) sscanf(e, "%d,%d", &f_tby, &f_tbx);
    if (const char *e = getenv("IVC_ME_THREADS")) f_thr = atoi(e);
    const size_t smem = me_int_geometry(a, G, pitch, n, H, W, sr, 200 * 1024, 4 * 3 * (int64_t)sm_count(device), f_tby, f_tbx);
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    if (H * W >= 2147483647LL) return cudaErrorInvalidValue;                  // 32-bit pixel coordinates inside a frame
    const int elem = u8 ? 1 : f32 ? 4 : 8;
    if (u8)                                                                    // 4-byte loads: window origin and rows 4-aligned
        a.vec = sr % 4 == 0 && ((uintptr_t)ref & 3) == 0 && ((uintptr_t)cur & 3) == 0 && ref_fs % 4 == 0 && cur_fs % 4 == 0;
    else
        a.vec = sr % (16 / elem) == 0 && ((uintptr_t)ref & 15) == 0 && ((uintptr_t)cur & 15) == 0 &&
                (ref_fs * elem) % 16 == 0 && (cur_fs * elem) % 16 == 0;        // W is a multiple of 8
    cudaError_t e;
    if (check && (e = cudaMemsetAsync(flag, 0, sizeof(int), st)) != cudaSuccess) return e;
    // threads per CTA: when three CTAs fit an SM by shared memory, the count in 256..352 that wastes the
    // fewest lanes in the last round of tasks; otherwise two CTAs of 512
    int threads = kMeIntMaxThreads;
    if (pitch == 0) threads = kMeThreads;
    else if (smem * 3 + 3 * 1024 <= 227 * 1024) {
        const int tasks = a.tby * a.tbx * a.ntpb;
        double best = 1e30;
        for (int t = 256; t <= 352; t += 32) {
            const double waste = (double)((tasks + t - 1) / t) * t / tasks;
            if (waste < best - 1e-9) { best = waste; threads = t; }
        }
    }
    if (f_thr) threads = f_thr;
    if (pf_zz) {                                                               // fused search + P-frame forward: +-4, tile 4 x 16
        if (!me_pf_fusable(dtype, sr) || G != 9 || pitch != 136) return cudaErrorInvalidValue;
        const int groups = a.tby * ((a.tbx + 7) / 8);                          // one warp-private WORK buffer per group, in U + S
        if (groups > (f_thr ? f_thr : kMePfThreads) / 32 || (size_t)groups * kPfWork > (size_t)a.b_off) return cudaErrorInvalidValue;
        const int pf_thr = f_thr ? f_thr : kMePfThreads;
        return u8 ? me_launch_chunks(k_me_int<unsigned char, 9, 136, true>, a, 1, smem, st, pf_thr)
                  : me_launch_chunks(k_me_int<double, 9, 136, true>, a, 8, smem, st, pf_thr);
    }
#define IVC_ME_INT_CASE(GG, PP)                                                                  \
    if (G == GG && pitch == PP)                                                                  \
        return u8 ? me_launch_chunks(k_me_int<unsigned char, GG, PP>, a, 1, smem, st, threads)   \
             : f32 ? me_launch_chunks(k_me_int<float, GG, PP>, a, 4, smem, st, threads)          \
                   : me_launch_chunks(k_me_int<double, GG, PP>, a, 8, smem, st, threads);
    IVC_ME_INT_CASE(9, 136)
    IVC_ME_INT_CASE(9, 144)
    IVC_ME_INT_CASE(11, 160)
    IVC_ME_INT_CASE(11, 0)
    IVC_ME_INT_CASE(9, 0)
    IVC_ME_INT_CASE(5, 0)
    IVC_ME_INT_CASE(3, 0)
#undef IVC_ME_INT_CASE
    return cudaErrorInvalidValue;
}

// ---- float64 frames -> uint8 planes, once per frame (the +-16 search reads every frame as a window AND as blocks, and the
// conversion is a fifth of its time when done while staging); CHECK raises the flag on a value that is not an integer in
// [0, 255], exactly like the staging of the integer kernels.  One thread = 8 pixels: four 16-byte loads, one 8-byte store.
template <bool CHECK>
__global__ void __launch_bounds__(256) k_f64_to_u8(const double *__restrict__ src, int64_t frame_stride, int64_t plane,
                                                   unsigned char *__restrict__ dst, int *flag) {
    const int64_t items = plane >> 3, f = blockIdx.y;
    const double *sp = src + f * frame_stride;
    unsigned char *dp = dst + f * plane;
    bool bad = false;
    for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += (int64_t)gridDim.x * blockDim.x) {
        double v[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double t[2];
            ldg16(sp + 8 * it + 2 * k, t);
            v[2 * k] = t[0]; v[2 * k + 1] = t[1];
        }
        uint2 w;
        w.x = pack_bytes(to_u8_fp<CHECK>(v[0], bad), to_u8_fp<CHECK>(v[1], bad), to_u8_fp<CHECK>(v[2], bad), to_u8_fp<CHECK>(v[3], bad));
        w.y = pack_bytes(to_u8_fp<CHECK>(v[4], bad), to_u8_fp<CHECK>(v[5], bad), to_u8_fp<CHECK>(v[6], bad), to_u8_fp<CHECK>(v[7], bad));
        *reinterpret_cast<uint2 *>(dp + 8 * it) = w;
    }
    if (CHECK && __syncthreads_or(bad) && threadIdx.x == 0) atomicOr(flag, 1);
}

// src: n frames of H*W float64 (16-byte aligned, frame_stride even, H*W a multiple of 8) -> dst [n][H*W] uint8
cudaError_t launch_f64_to_u8(int device, cudaStream_t st, const void *src, int64_t frame_stride, int64_t n, int64_t plane,
                             void *dst, int *flag) {
    if (n == 0 || plane == 0) return cudaSuccess;
    int64_t gx = (plane / 8 + 255) / 256;
    const int64_t cap = (int64_t)sm_count(device) * 8;
    if (gx > cap) gx = cap;
    for (int64_t f0 = 0; f0 < n; f0 += 65535) {
        const unsigned nf = (unsigned)(n - f0 < 65535 ? n - f0 : 65535);
        const double *sp = (const double *)src + f0 * frame_stride;
        unsigned char *dp = (unsigned char *)dst + f0 * plane;
        if (flag) k_f64_to_u8<true><<<dim3((unsigned)gx, nf), 256, 0, st>>>(sp, frame_stride, plane, dp, flag);
        else k_f64_to_u8<false><<<dim3((unsigned)gx, nf), 256, 0, st>>>(sp, frame_stride, plane, dp, nullptr);
    }
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
