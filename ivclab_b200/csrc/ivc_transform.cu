// ivc_transform.cu -- DCT / quantiser / zig-zag kernels for sm_100a.
//
// Work decomposition of the fused kernels (K1 forward, K2 inverse, and their P-frame variants):
//   * a TILE is 8 pixel rows x 12 "block-channels" (12 independent 8x8 transforms):
//       C == 3 : 4 horizontally adjacent blocks x 3 interleaved channels
//       C == 1 : 12 horizontally adjacent luma blocks
//     in both cases one tile row is 96 consecutive doubles (768 B) of the HWC image.
//   * ONE WARP owns a tile end to end and never synchronises with other warps.  Lane = r + 8u:
//     in the row pass lane (r,u) transforms pixel row r of the three block-channels (u,0..2);
//     the 8x8 transposition goes through a warp-private shared-memory buffer; in the column pass
//     lane (j,u) transforms column j of the same three block-channels.
//   * global traffic is always 16-byte vector loads/stores in which consecutive lanes touch
//     consecutive addresses; the de-interleave / zig-zag scatter happens in shared memory.
//   * the per-warp 8 KB buffer is reused by the phases of a tile (input rows -> transposition ->
//     output staging); phases are separated by __syncwarp().
//   * persistent grid: every warp strides over the tile list of the whole batch.
// HBM traffic is exactly the algorithmic bytes (each input element read once, each output written
// once); arithmetic is FP64 (14 rounded operations per sample for the 2-D transform).
#include <cstdlib>
#include <type_traits>
#include <cuda.h>          // CUtensorMap types only; the encoder is fetched through cudaGetDriverEntryPoint
#include "ivc_dct.cuh"
#include "ivc_color.cuh"
#include "ivc_common.cuh"
#include "ivc_tile.cuh"

namespace ivc {

// ------------------------------------------------------------------------------------------------
// per-warp scratch geometry (in doubles unless stated)
// ------------------------------------------------------------------------------------------------
constexpr int kWarpsPerCta = 8;
constexpr int kWarpBufBytes = 8192;
constexpr int kRowPitch = 98;        // 96 doubles of tile row + 2 pad: 784 B == 16 (mod 128)
constexpr int kTJ = 10;              // transposition buffer: T[u][m*8+j][r], 80-byte rows
constexpr int kTU = 248;             // 24 rows * 10 + 8 pad: u-planes land 64 B apart (mod 128)
constexpr int kStageU = 200;         // int32 staging: 3 chunks * 64 ints + 8 pad per u

static_assert(8 * kRowPitch * 8 <= kWarpBufBytes, "row tile must fit");
static_assert(4 * kTU * 8 <= kWarpBufBytes, "transposition buffer must fit");
static_assert(4 * kStageU * 4 <= kWarpBufBytes, "staging must fit");
static_assert((kStageUF * 4) % 16 == 0, "bulk-store sources are 16-byte aligned");

struct TileGeom {
    int64_t n_frames, H, W;          // pixels
    int Hp, Wp;                      // blocks
    int C;                           // channels of the image (1 or 3)
    int tiles_per_row;               // ceil(Wp / blocks_per_tile)
    int64_t total_tiles;
};

__device__ __forceinline__ void tile_coords(const TileGeom &g, int64_t t, int64_t &frame, int &by, int &tx) {
    tx = (int)(t % g.tiles_per_row);
    const int64_t q = t / g.tiles_per_row;
    by = (int)(q % g.Hp);
    frame = q / g.Hp;
}

__device__ __forceinline__ double2 ldg_stream(const double *p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ int4 ldg_stream(const int4 *p) {
    int4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_stream(double *p, double2 v) {
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void stg_stream(int4 *p, int4 v) {
    asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}

// decode a motion-vector index (motion.py:83-84) and test the source window (motion.py:90-92)
__device__ __forceinline__ void mv_decode(int64_t idx, int sr, int &dy, int &dx) {
    const unsigned span = 2u * (unsigned)sr + 1u;
    if (__builtin_expect((unsigned long long)idx < 0x7fffffffull, 1)) {     // every vector ME produces: 32-bit divide
        const unsigned q = (unsigned)idx / span;
        dy = (int)q - sr;
        dx = (int)((unsigned)idx - q * span) - sr;
        return;
    }
    // python floor division / modulo semantics for arbitrary (negative, huge) indices
    const int64_t sp = (int64_t)span;
    int64_t qd = idx / sp, rm = idx % sp;
    if (rm < 0) { rm += sp; qd -= 1; }
    qd -= sr;
    dy = qd > 0x3fffffff ? 0x3fffffff : (qd < -0x3fffffff ? -0x3fffffff : (int)qd);   // clamped: window is out of frame anyway
    dx = (int)(rm - sr);
}

// ================================================================================================
// K1: fused forward  (patch -> DCT -> quantize -> zig-zag), optionally with MC + residual in front
// ================================================================================================
struct FwdArgs {
    TileGeom g;
    const double *img;               // intra: HWC image(s); pframe: current luma plane(s)
    int64_t frame_stride;            // elements
    const void *table;
    int table_dtype;
    int32_t *out;                    // [n, Hp, Wp, 3, 64]
    // P-frame extras
    const double *ref;
    const int64_t *mv;
    int sr;
    double *pred_out;                // may be null
    int och;                         // P-frame: scan channels stored per block (3 = the reference's broadcast; 2 = tables 0 and 1 only)
    const int *run_flag;             // P-frame: when not null the kernel runs only if *run_flag != 0 (fallback after the fused search)
    int32_t *zr_counts;              // optional (TMA forward kernels): per scan block, the symbols ZeroRunCoder.encode emits for it ...
    unsigned long long *zr_masks;    // ... and its 64-bit non-zero mask: the zero-run coder's count pass, done while the block is in smem
    // k_forward_rgb8_tma only: nq quantisation tables [nq][3][64] applied to ONE transform of the frames (a rate-distortion sweep
    // codes the same frame at every scale: colour transform and DCT do not depend on the scale)
    int nq = 1;
    int64_t out_q_stride = 0;        // int32 elements between the outputs of successive tables
    int64_t zr_q_stride = 0;         // scan blocks between the counts / masks of successive tables
};

template <int C, bool PFRAME>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 2) k_forward(const FwdArgs a) {
    if (PFRAME && a.run_flag && *a.run_flag == 0) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_rt = reinterpret_cast<double *>(smem_raw);            // [3*64]
    double *s_t = s_rt + 192;                                        // [3*64]
    unsigned char *s_warp = smem_raw + 2 * 192 * sizeof(double);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *buf = reinterpret_cast<double *>(s_warp + warp * kWarpBufBytes);
    int *ibuf = reinterpret_cast<int *>(buf);

    for (int i = threadIdx.x; i < 192; i += blockDim.x) {
        const double t = load_table_elem(a.table, a.table_dtype, i);
        s_t[i] = t;
        s_rt[i] = __drcp_rn(t);
    }
    __syncthreads();

    const int r = lane & 7, u = lane >> 3;       // row pass: pixel row r; column pass: column j == r
    int zz[8];                                   // scan position of raster (v, j=r)
#pragma unroll
    for (int v = 0; v < 8; ++v) zz[v] = ZZ_ORDER[v * 8 + r];

    const TileGeom &g = a.g;
    constexpr int kBlocksPerTile = (C == 3) ? 4 : 12;
    const int64_t row_elems = g.W * C;           // doubles per image row
    const int64_t gw = (int64_t)blockIdx.x * kWarpsPerCta + warp;
    const int64_t nw = (int64_t)gridDim.x * kWarpsPerCta;

    for (int64_t t = gw; t < g.total_tiles; t += nw) {
        int64_t frame; int by, tx;
        tile_coords(g, t, frame, by, tx);
        const int b0 = tx * kBlocksPerTile;
        const int nb = min(kBlocksPerTile, g.Wp - b0);               // valid blocks in this tile
        const int nchunk = nb * (C == 3 ? 12 : 4);                   // valid 16-byte chunks per tile row
        const double *src = a.img + frame * a.frame_stride + (int64_t)by * 8 * row_elems + (int64_t)tx * 96;

        // ---- phase 1: coalesced 16-byte loads of the 8 x 768 B tile into padded smem rows ----
        if (!PFRAME) {
            double2 v[12];
#pragma unroll
            for (int k = 0; k < 12; ++k) {
                const int id = lane + 32 * k, row = id / 48, c16 = id % 48;
                v[k] = (c16 < nchunk) ? ldg_stream(src + row * row_elems + 2 * c16) : make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int k = 0; k < 12; ++k) {
                const int id = lane + 32 * k, row = id / 48, c16 = id % 48;
                *reinterpret_cast<double2 *>(buf + row * kRowPitch + 2 * c16) = v[k];
            }
        } else {
            // residual = cur - MC(ref, mv) (videocodec.py:68-71); 12 luma blocks per tile.
            // lanes 0..11 decode the block vectors once, the rest fetch them by shuffle.
            int dy = 0, dx = 0;
            if (lane < nb) mv_decode(a.mv[(frame * g.Hp + by) * g.Wp + b0 + lane], a.sr, dy, dx);
            const double *refp = a.ref + frame * a.frame_stride;
            double *predp = a.pred_out ? a.pred_out + frame * a.frame_stride + (int64_t)by * 8 * row_elems + (int64_t)tx * 96
                                       : nullptr;
#pragma unroll
            for (int k = 0; k < 12; ++k) {
                const int id = lane + 32 * k, row = id / 48, c16 = id % 48;
                const int blk = c16 >> 2;
                const int bdy = __shfl_sync(0xffffffffu, dy, blk), bdx = __shfl_sync(0xffffffffu, dx, blk);
                double2 res = make_double2(0.0, 0.0);
                if (c16 < nchunk) {
                    const double2 cur = ldg_stream(src + row * row_elems + 2 * c16);
                    const int sy = (by + 0) * 8 + bdy, sx = (b0 + blk) * 8 + bdx;       // source window origin
                    double2 pr = make_double2(0.0, 0.0);
                    if (sy >= 0 && sy + 8 <= g.H && sx >= 0 && sx + 8 <= g.W) {
                        const double *rp = refp + (int64_t)(sy + row) * row_elems + sx + ((2 * c16) & 7);
                        pr.x = __ldg(rp);
                        pr.y = __ldg(rp + 1);
                    }
                    if (predp) stg_stream(predp + row * row_elems + 2 * c16, pr);
                    res.x = __dsub_rn(cur.x, pr.x);
                    res.y = __dsub_rn(cur.y, pr.y);
                }
                *reinterpret_cast<double2 *>(buf + row * kRowPitch + 2 * c16) = res;
            }
        }
        __syncwarp();

        // ---- phase 2: row pass.  lane (r,u) reads its 24 consecutive doubles (conflict-free) ----
        double x[3][8];
        {
            double raw[24];
#pragma unroll
            for (int k = 0; k < 12; ++k) {
                const double2 v = *reinterpret_cast<const double2 *>(buf + r * kRowPitch + 24 * u + 2 * k);
                raw[2 * k] = v.x;
                raw[2 * k + 1] = v.y;
            }
#pragma unroll
            for (int m = 0; m < 3; ++m)
#pragma unroll
                for (int p = 0; p < 8; ++p) x[m][p] = (C == 3) ? raw[3 * p + m] : raw[8 * m + p];
        }
#pragma unroll
        for (int m = 0; m < 3; ++m) dct2_8(x[m]);
        __syncwarp();                                   // everyone has consumed the input rows
#pragma unroll
        for (int m = 0; m < 3; ++m)
#pragma unroll
            for (int j = 0; j < 8; ++j) buf[u * kTU + (m * 8 + j) * kTJ + r] = x[m][j];
        __syncwarp();

        // ---- phase 3: column pass.  lane (j,u), j == r ----
#pragma unroll
        for (int m = 0; m < 3; ++m) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double2 v = *reinterpret_cast<const double2 *>(buf + u * kTU + (m * 8 + r) * kTJ + 2 * k);
                x[m][2 * k] = v.x;
                x[m][2 * k + 1] = v.y;
            }
            dct2_8(x[m]);                               // x[m][v] = coefficient (v, j)
        }
        __syncwarp();                                   // transposition buffer is dead

        // ---- phase 4: quantise, zig-zag scatter into staging, coalesced 16-byte stores ----
        const int och = PFRAME ? a.och : 3;
        int32_t *outf = a.out + ((frame * g.Hp + by) * (int64_t)g.Wp) * (64 * och);
        if (C == 3) {
#pragma unroll
            for (int m = 0; m < 3; ++m)
#pragma unroll
                for (int v = 0; v < 8; ++v) {
                    const int k = m * 64 + v * 8 + r;
                    ibuf[u * kStageU + m * 64 + zz[v]] = quantize_f64(x[m][v], s_t[k], s_rt[k]);
                }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                const int id = lane + 32 * k, chunk = id >> 4, part = id & 15;   // chunk = u*3 + m
                const int cu = chunk / 3;
                if (cu < nb) {
                    const int4 v = *reinterpret_cast<const int4 *>(ibuf + cu * kStageU + (chunk - cu * 3) * 64 + part * 4);
                    stg_stream(reinterpret_cast<int4 *>(outf + (int64_t)(b0 * 3 + chunk) * 64 + part * 4), v);
                }
            }
            __syncwarp();
        } else {
            // C == 1: numpy broadcasting quantises the single channel with all three tables
            // (patchquant.py:59).  One round per sub-block m: 4 blocks x 3 tables = 12 chunks.
#pragma unroll
            for (int m = 0; m < 3; ++m) {
#pragma unroll
                for (int ch = 0; ch < 3; ++ch)
#pragma unroll
                    for (int v = 0; v < 8; ++v) {
                        const int k = ch * 64 + v * 8 + r;
                        ibuf[u * kStageU + ch * 64 + zz[v]] = quantize_f64(x[m][v], s_t[k], s_rt[k]);
                    }
                __syncwarp();
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    const int id = lane + 32 * k, chunk = id >> 4, part = id & 15;
                    const int cu = chunk / 3, ch = chunk - cu * 3;
                    const int blk = 3 * cu + m;
                    if (blk < nb && ch < och) {
                        const int4 v = *reinterpret_cast<const int4 *>(ibuf + cu * kStageU + ch * 64 + part * 4);
                        stg_stream(reinterpret_cast<int4 *>(outf + ((int64_t)(b0 + blk) * och + ch) * 64 + part * 4), v);
                    }
                }
                __syncwarp();
            }
        }
    }
}

// ================================================================================================
// K2: fused inverse (un-zig-zag -> dequantize -> IDCT -> un-patch), optionally + prediction
// ================================================================================================
struct InvArgs {
    TileGeom g;                      // g.C = channels of the scan input (1 or 3); PFRAME: luma geometry
    const int32_t *zz;               // [n, Hp, Wp, Czz, 64]
    int Czz;
    const void *table;
    int table_dtype;
    double *out;                     // intra: [n, H, W, 3]; pframe: [n, H, W]
    // P-frame extras
    const double *pred;              // may be null -> gather from (ref, mv)
    const double *ref;
    const int64_t *mv;
    int sr;
    // intra + distortion (k_inverse_c3_tma<SSE>): uint8 RGB originals and one partial sum per tile
    const unsigned char *orig;       // [n, H, W, 3] uint8, frames orig_frame_stride bytes apart
    int64_t orig_frame_stride;
    double *sse_partial;             // [total_tiles]
};

// MODE 0: intra, 3 scan channels -> 3 image channels      (tile = 4 blocks x 3 channels)
// MODE 1: intra, 1 scan channel broadcast against 3 tables (tile = 4 blocks, m = table index)
// MODE 2: P-frame luma: channel 0 of the scan, luminance table, + prediction (tile = 12 blocks)
template <int MODE>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 2) k_inverse(const InvArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_tT = reinterpret_cast<double *>(smem_raw);            // [3][j*8 + r]
    unsigned char *s_warp = smem_raw + 192 * sizeof(double);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *buf = reinterpret_cast<double *>(s_warp + warp * kWarpBufBytes);
    int *ibuf = reinterpret_cast<int *>(buf);

    for (int i = threadIdx.x; i < 192; i += blockDim.x) {
        const int ch = i >> 6, k = i & 63, rr = k >> 3, jj = k & 7;
        s_tT[ch * 64 + jj * 8 + rr] = load_table_elem(a.table, a.table_dtype, i);
    }
    __syncthreads();

    const int r = lane & 7, u = lane >> 3;
    int zr[8];                                   // scan position of raster (r, j)
#pragma unroll
    for (int j = 0; j < 8; ++j) zr[j] = ZZ_ORDER[r * 8 + j];

    const TileGeom &g = a.g;
    constexpr int kBlocksPerTile = (MODE == 2) ? 12 : 4;
    constexpr int kOutC = (MODE == 2) ? 1 : 3;
    const int64_t row_elems = g.W * kOutC;
    const int64_t out_frame = g.H * row_elems;
    const int64_t gw = (int64_t)blockIdx.x * kWarpsPerCta + warp;
    const int64_t nw = (int64_t)gridDim.x * kWarpsPerCta;

    for (int64_t t = gw; t < g.total_tiles; t += nw) {
        int64_t frame; int by, tx;
        tile_coords(g, t, frame, by, tx);
        const int b0 = tx * kBlocksPerTile;
        const int nb = min(kBlocksPerTile, g.Wp - b0);
        const int32_t *zsrc = a.zz + ((frame * g.Hp + by) * (int64_t)g.Wp + b0) * a.Czz * 64;

        // ---- phase 1: stage the scan blocks (256-byte chunks) ----
        // staging slot s (0..11) holds block-channel (u = s/3, m = s%3)
        if (MODE == 0) {
            int4 v[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                const int id = lane + 32 * k, chunk = id >> 4, part = id & 15;
                v[k] = (chunk / 3 < nb) ? ldg_stream(reinterpret_cast<const int4 *>(zsrc + chunk * 64 + part * 4))
                                        : make_int4(0, 0, 0, 0);
            }
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                const int id = lane + 32 * k, chunk = id >> 4, part = id & 15;
                const int cu = chunk / 3;
                *reinterpret_cast<int4 *>(ibuf + cu * kStageU + (chunk - cu * 3) * 64 + part * 4) = v[k];
            }
        } else if (MODE == 1) {
            // 4 blocks, one scan channel each (Czz == 1): slot (u, 0)
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int id = lane + 32 * k, cu = id >> 4, part = id & 15;
                const int4 v = (cu < nb) ? ldg_stream(reinterpret_cast<const int4 *>(zsrc + (int64_t)cu * 64 + part * 4))
                                         : make_int4(0, 0, 0, 0);
                *reinterpret_cast<int4 *>(ibuf + cu * kStageU + part * 4) = v;
            }
        } else {
            // 12 luma blocks: slot (u, m) <- block 3u+m, scan channel 0 of Czz
            int4 v[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                const int id = lane + 32 * k, blk = id >> 4, part = id & 15;
                v[k] = (blk < nb) ? ldg_stream(reinterpret_cast<const int4 *>(zsrc + (int64_t)blk * a.Czz * 64 + part * 4))
                                  : make_int4(0, 0, 0, 0);
            }
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                const int id = lane + 32 * k, blk = id >> 4, part = id & 15;
                const int cu = blk / 3;
                *reinterpret_cast<int4 *>(ibuf + cu * kStageU + (blk - cu * 3) * 64 + part * 4) = v[k];
            }
        }
        __syncwarp();

        // ---- phase 2: gather raster row r, dequantise, row IDCT (axis -1 first, dct.py:42) ----
        double x[3][8];
#pragma unroll
        for (int m = 0; m < 3; ++m) {
            const int slot = (MODE == 1) ? 0 : m;
            const int tch = (MODE == 2) ? 0 : m;         // P-frame: luminance table for every block
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int q = ibuf[u * kStageU + slot * 64 + zr[j]];
                x[m][j] = dequantize_f64(q, s_tT[tch * 64 + j * 8 + r]);
            }
            dct3_8(x[m]);
        }
        __syncwarp();
#pragma unroll
        for (int m = 0; m < 3; ++m)
#pragma unroll
            for (int j = 0; j < 8; ++j) buf[u * kTU + (m * 8 + j) * kTJ + r] = x[m][j];
        __syncwarp();

        // ---- phase 3: column IDCT.  lane (j,u), j == r ----
#pragma unroll
        for (int m = 0; m < 3; ++m) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double2 v = *reinterpret_cast<const double2 *>(buf + u * kTU + (m * 8 + r) * kTJ + 2 * k);
                x[m][2 * k] = v.x;
                x[m][2 * k + 1] = v.y;
            }
            dct3_8(x[m]);                               // x[m][i] = pixel (row i, column j)
        }
        __syncwarp();

        // ---- phase 4: un-patch through a padded row tile, then coalesced 16-byte stores ----
#pragma unroll
        for (int m = 0; m < 3; ++m)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int col = (MODE == 2) ? (3 * u + m) * 8 + r : (8 * u + r) * 3 + m;
                buf[i * kRowPitch + col] = x[m][i];
            }
        __syncwarp();
        const int nchunk = nb * (MODE == 2 ? 4 : 12);
        double *dst = a.out + frame * out_frame + (int64_t)by * 8 * row_elems + (int64_t)tx * 96;
        if (MODE != 2) {
#pragma unroll
            for (int k = 0; k < 12; ++k) {
                const int id = lane + 32 * k, row = id / 48, c16 = id % 48;
                if (c16 < nchunk)
                    stg_stream(dst + row * row_elems + 2 * c16,
                               *reinterpret_cast<const double2 *>(buf + row * kRowPitch + 2 * c16));
            }
        } else {
            int dy = 0, dx = 0;
            if (!a.pred && lane < nb) mv_decode(a.mv[(frame * g.Hp + by) * g.Wp + b0 + lane], a.sr, dy, dx);
            const double *predp = a.pred ? a.pred + frame * out_frame + (int64_t)by * 8 * row_elems + (int64_t)tx * 96 : nullptr;
            const double *refp = a.ref ? a.ref + frame * out_frame : nullptr;
#pragma unroll
            for (int k = 0; k < 12; ++k) {
                const int id = lane + 32 * k, row = id / 48, c16 = id % 48;
                const int blk = c16 >> 2;
                const int bdy = __shfl_sync(0xffffffffu, dy, blk), bdx = __shfl_sync(0xffffffffu, dx, blk);
                if (c16 < nchunk) {
                    double2 pr = make_double2(0.0, 0.0);
                    if (predp) {
                        pr = ldg_stream(predp + row * row_elems + 2 * c16);
                    } else {
                        const int sy = by * 8 + bdy, sx = (b0 + blk) * 8 + bdx;
                        if (sy >= 0 && sy + 8 <= g.H && sx >= 0 && sx + 8 <= g.W) {
                            const double *rp = refp + (int64_t)(sy + row) * row_elems + sx + ((2 * c16) & 7);
                            pr.x = __ldg(rp);
                            pr.y = __ldg(rp + 1);
                        }
                    }
                    const double2 rec = *reinterpret_cast<const double2 *>(buf + row * kRowPitch + 2 * c16);
                    // recon = prediction + recon_residual (videocodec.py:74)
                    stg_stream(dst + row * row_elems + 2 * c16,
                               make_double2(__dadd_rn(pr.x, rec.x), __dadd_rn(pr.y, rec.y)));
                }
            }
        }
        __syncwarp();
    }
}

// ================================================================================================
// v2 of K1 / K2 for the 3-channel intra path: same arithmetic and lane mapping, but the tile moves
// through the TMA unit.  Each warp owns an mbarrier and two buffers:
//   IN   (6272 B) filled by cp.async.bulk (one 1-D bulk copy per pixel row / per scan block) --
//        the copy of tile i+1 is issued as soon as the row pass has pulled tile i into registers,
//        so HBM latency overlaps the column pass, the quantiser and the store of tile i;
//   WORK (6400 B) transposition buffer (dense, XOR-swizzled 16-byte chunks instead of padding),
//        then output staging that a cp.async.bulk store drains asynchronously.
// No warp ever waits on another warp; no registers are spent on staging.
// ================================================================================================
constexpr int kInBytes = 8 * kRowPitch * 8;      // 6272
constexpr int kWorkBytes = 6400;
constexpr int kWarpBuf2 = kInBytes + kWorkBytes; // 12672 = 99 * 128
constexpr int kTU2 = 200;                        // doubles per u-plane: 24 rows * 8 + 8 skew (64 B)
static_assert(kWarpBuf2 % 128 == 0, "per-warp buffers stay 128-byte aligned");
static_assert(8 * kRowPitch * 8 <= kWorkBytes, "the output row tile reuses WORK");

// tile iterator: (frame, block row, tile column) advanced by a fixed stride without divisions
struct TileIter {
    int tx, by, dtx, dby;
    int64_t frame, dframe;
    __device__ __forceinline__ void init(const TileGeom &g, int64_t t0, int64_t step) {
        tx = (int)(t0 % g.tiles_per_row);
        int64_t q = t0 / g.tiles_per_row;
        by = (int)(q % g.Hp);
        frame = q / g.Hp;
        dtx = (int)(step % g.tiles_per_row);
        q = step / g.tiles_per_row;
        dby = (int)(q % g.Hp);
        dframe = q / g.Hp;
    }
    __device__ __forceinline__ void advance(const TileGeom &g) {
        tx += dtx;
        int c = tx >= g.tiles_per_row;
        tx -= c ? g.tiles_per_row : 0;
        by += dby + c;
        c = by >= g.Hp;
        by -= c ? g.Hp : 0;
        frame += dframe + c;
    }
};

__global__ void __launch_bounds__(kWarpsPerCta * 32, 2) k_forward_c3_tma(const FwdArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *s_rt = reinterpret_cast<double *>(smem_raw);                       // [192]
    double *s_t = s_rt + 192;                                                   // [192]
    unsigned long long *s_bar = reinterpret_cast<unsigned long long *>(smem_raw + 3072);   // [8]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *in_b = smem_raw + 3200 + warp * kWarpBuf2;
    unsigned char *work_b = in_b + kInBytes;
    const uint32_t bar = smem_u32(s_bar + warp), in_s = smem_u32(in_b), work_s = smem_u32(work_b);

    for (int i = threadIdx.x; i < 192; i += blockDim.x) {
        const double t = load_table_elem(a.table, a.table_dtype, i);
        s_t[i] = t;
        s_rt[i] = __drcp_rn(t);
    }
    if (lane == 0) mbar_init(bar, 1);
    fence_mbar_init();
    __syncthreads();

    // lane-constant addresses (everything below indexes them with compile-time offsets)
    const int r = lane & 7, u = lane >> 3;
    const unsigned char *rd_in = in_b + r * (kRowPitch * 8) + u * 192;
    unsigned char *t_wr[4], *zz_wr[8];
    const unsigned char *t_rd[4];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        t_wr[h] = work_b + u * (kTU2 * 8) + ((((r >> 1) ^ h) << 1) + (r & 1)) * 8;    // + row*64
        t_rd[h] = work_b + u * (kTU2 * 8) + r * 64 + ((h ^ (r >> 1)) << 4);           // + m*512
    }
#pragma unroll
    for (int v = 0; v < 8; ++v) zz_wr[v] = work_b + (u * kStageUF + ZZ_ORDER[v * 8 + r]) * 4;   // + m*256
    const double *rt_l = s_rt + r, *t_l = s_t + r;

    const TileGeom &g = a.g;
    const int64_t row_elems = g.W * 3;
    const int64_t gw = (int64_t)blockIdx.x * kWarpsPerCta + warp;
    const int64_t nw = (int64_t)gridDim.x * kWarpsPerCta;
    if (gw >= g.total_tiles) return;
    const int64_t my_tiles = (g.total_tiles - gw + nw - 1) / nw;
    TileIter cur, nxt;
    cur.init(g, gw, nw);
    nxt = cur;

    auto issue = [&](const TileIter &ti) {                  // whole warp: lanes 0..7 each copy one row
        const int nb = min(4, g.Wp - ti.tx * 4);
        const uint32_t row_bytes = (uint32_t)nb * 192u;
        fence_proxy_async();
        if (lane == 0) mbar_expect_tx(bar, 8u * row_bytes);
        __syncwarp();
        if (lane < 8) {                                     // lane r copies pixel row r
            const double *src = a.img + ti.frame * a.frame_stride + ((int64_t)ti.by * 8 + lane) * row_elems + (int64_t)ti.tx * 96;
            bulk_g2s(in_s + lane * (kRowPitch * 8), src, row_bytes, bar);
        }
    };

    issue(cur);
    uint32_t parity = 0;
    for (int64_t it = 0; it < my_tiles; ++it, parity ^= 1u) {
        nxt.advance(g);
        mbar_wait(bar, parity);
        double x[3][8];
        {
            double raw[24];
#pragma unroll
            for (int k = 0; k < 12; ++k) {
                const double2 v = *reinterpret_cast<const double2 *>(rd_in + 16 * k);
                raw[2 * k] = v.x;
                raw[2 * k + 1] = v.y;
            }
#pragma unroll
            for (int m = 0; m < 3; ++m)
#pragma unroll
                for (int p = 0; p < 8; ++p) x[m][p] = raw[3 * p + m];
        }
        __syncwarp();                                   // IN is consumed: prefetch the next tile into it
        if (it + 1 < my_tiles) issue(nxt);
        bulk_wait_read0();                              // previous tile's stores (each lane its own group) have drained WORK
#pragma unroll
        for (int m = 0; m < 3; ++m) dct2_8(x[m]);
        __syncwarp();
#pragma unroll
        for (int m = 0; m < 3; ++m)
#pragma unroll
            for (int j = 0; j < 8; ++j) *reinterpret_cast<double *>(t_wr[j >> 1] + (m * 8 + j) * 64) = x[m][j];
        __syncwarp();
#pragma unroll
        for (int m = 0; m < 3; ++m) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double2 v = *reinterpret_cast<const double2 *>(t_rd[k] + m * 512);
                x[m][2 * k] = v.x;
                x[m][2 * k + 1] = v.y;
            }
            dct2_8(x[m]);
        }
        __syncwarp();
        QuantGuard qg;
        {
            double rtv[3][8];                           // loaded as one batch: the staging stores below would otherwise
#pragma unroll                                          // serialise each (possibly aliasing) shared-memory load
            for (int m = 0; m < 3; ++m)
#pragma unroll
                for (int v = 0; v < 8; ++v) rtv[m][v] = rt_l[m * 64 + v * 8];
            int qv[3][8];
#pragma unroll
            for (int m = 0; m < 3; ++m)
#pragma unroll
                for (int v = 0; v < 8; ++v) qv[m][v] = qg.q(x[m][v], rtv[m][v]);
#pragma unroll
            for (int m = 0; m < 3; ++m)
#pragma unroll
                for (int v = 0; v < 8; ++v) *reinterpret_cast<int *>(zz_wr[v] + m * 256) = qv[m][v];
        }
        if (__builtin_expect(qg.risky(), 0)) {          // rare: redo this lane's 24 samples with the IEEE division
#pragma unroll
            for (int m = 0; m < 3; ++m)
#pragma unroll
                for (int v = 0; v < 8; ++v)
                    *reinterpret_cast<int *>(zz_wr[v] + m * 256) = quantize_exact_f64(x[m][v], t_l[m * 64 + v * 8]);
        }
        fence_proxy_async();                            // make the staging visible to the bulk-copy unit
        __syncwarp();
        {
            const int b0 = cur.tx * 4, nb = min(4, g.Wp - b0);
            if (lane < nb) {                              // lane u stores the 3 scan blocks of image block u
                int32_t *outf = a.out + ((cur.frame * g.Hp + cur.by) * (int64_t)g.Wp + b0 + lane) * 192;
                bulk_s2g(outf, work_s + lane * (kStageUF * 4), 768u);
                bulk_commit();
            }
            if (a.zr_masks) {                             // scan block j = 3 u + m of the tile, in (h w c) order
                unsigned long long mk;
                zr_masks_from_staging<12>(work_b, [](int j) { return (j / 3) * (kStageUF * 4) + (j % 3) * 256; }, lane, mk);
                if (lane < 3 * nb) {
                    const int64_t sblk = ((cur.frame * g.Hp + cur.by) * (int64_t)g.Wp + b0) * 3 + lane;
                    a.zr_masks[sblk] = mk;
                    a.zr_counts[sblk] = zr_block_count(mk);
                }
            }
        }
        cur = nxt;
    }
    bulk_wait_all0();
}

// K1 with the colour transform in front (row N1): uint8 RGB HWC in, same scan indices out.
constexpr int kRgbPitch = 112;                       // 96 B of pixels + 16 B pad (bulk copies need 16-byte rows)
constexpr int kRgbIn = 8 * kRgbPitch;                // 896
constexpr int kRgbBuf = 7424;                        // kRgbIn + kWorkBytes rounded up to 128
constexpr int kMaxForwardScales = 16;                // tables of one launch: 16 x 3 KB next to the warps' buffers keeps two CTAs per SM
static_assert(kRgbIn + kWorkBytes <= kRgbBuf && kRgbBuf % 128 == 0, "layout");

template <bool MULTI>                                // MULTI: a.nq tables (a runtime loop around the quantiser); else exactly one
__global__ void __launch_bounds__(kWarpsPerCta * 32, 2) k_forward_rgb8_tma(const FwdArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *s_rt = reinterpret_cast<double *>(smem_raw);                       // [nq][fl(1/t) 192 | t 192]
    const int nq = MULTI ? a.nq : 1;
    const int tab_bytes = nq * 3072;
    unsigned long long *s_bar = reinterpret_cast<unsigned long long *>(smem_raw + tab_bytes);   // [8]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *in_b = smem_raw + tab_bytes + 128 + warp * kRgbBuf;   // 8 rows x 96 B of packed RGB, pitch 112 B
    unsigned char *work_b = in_b + kRgbIn;
    const uint32_t bar = smem_u32(s_bar + warp), in_s = smem_u32(in_b), work_s = smem_u32(work_b);

    for (int i = threadIdx.x; i < 192 * nq; i += blockDim.x) {
        const int q = i / 192, e = i - q * 192;
        const double t = load_table_elem(a.table, a.table_dtype, i);
        s_rt[q * 384 + 192 + e] = t;
        s_rt[q * 384 + e] = __drcp_rn(t);
    }
    if (lane == 0) mbar_init(bar, 1);
    fence_mbar_init();
    __syncthreads();

    // lane-constant addresses (everything below indexes them with compile-time offsets)
    const int r = lane & 7, u = lane >> 3;
    const unsigned char *rd_in = in_b + r * kRgbPitch + u * 24;
    unsigned char *t_wr[4], *zz_wr[8];
    const unsigned char *t_rd[4];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        t_wr[h] = work_b + u * (kTU2 * 8) + ((((r >> 1) ^ h) << 1) + (r & 1)) * 8;    // + row*64
        t_rd[h] = work_b + u * (kTU2 * 8) + r * 64 + ((h ^ (r >> 1)) << 4);           // + m*512
    }
#pragma unroll
    for (int v = 0; v < 8; ++v) zz_wr[v] = work_b + (u * kStageUF + ZZ_ORDER[v * 8 + r]) * 4;   // + m*256
    const double *rt_l0 = s_rt + r;

    const TileGeom &g = a.g;
    const int64_t row_elems = g.W * 3;                                   // BYTES per image row (uint8 RGB)
    const int64_t gw = (int64_t)blockIdx.x * kWarpsPerCta + warp;
    const int64_t nw = (int64_t)gridDim.x * kWarpsPerCta;
    if (gw >= g.total_tiles) return;
    const int64_t my_tiles = (g.total_tiles - gw + nw - 1) / nw;
    TileIter cur, nxt;
    cur.init(g, gw, nw);
    nxt = cur;

    auto issue = [&](const TileIter &ti) {                  // whole warp: lanes 0..7 each copy one row
        const int nb = min(4, g.Wp - ti.tx * 4);
        const uint32_t row_bytes = (uint32_t)nb * 24u;                      // nb is even (W % 16 == 0): multiple of 16
        fence_proxy_async();
        if (lane == 0) mbar_expect_tx(bar, 8u * row_bytes);
        __syncwarp();
        if (lane < 8) {
            const unsigned char *src = reinterpret_cast<const unsigned char *>(a.img) + ti.frame * a.frame_stride +
                                       ((int64_t)ti.by * 8 + lane) * row_elems + (int64_t)ti.tx * 96;
            bulk_g2s(in_s + lane * kRgbPitch, src, row_bytes, bar);
        }
    };

    issue(cur);
    uint32_t parity = 0;
    for (int64_t it = 0; it < my_tiles; ++it, parity ^= 1u) {
        nxt.advance(g);
        mbar_wait(bar, parity);
        double x[3][8];
        {
            // 8 pixels x RGB = 24 bytes; rgb2ycbcr (color.py:15-37) on the fly, then the same path as K1
            const uint2 w0 = *reinterpret_cast<const uint2 *>(rd_in), w1 = *reinterpret_cast<const uint2 *>(rd_in + 8),
                        w2 = *reinterpret_cast<const uint2 *>(rd_in + 16);
            const unsigned w[6] = {w0.x, w0.y, w1.x, w1.y, w2.x, w2.y};
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                double c[3];
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    const int bi = 3 * p + ch;
                    const unsigned byte = (w[bi >> 2] >> (8 * (bi & 3))) & 255u;
                    c[ch] = __dsub_rn(__hiloint2double(0x43300000, (int)byte), 4503599627370496.0);   // exact u8 -> f64
                }
                rgb2ycbcr_px(c[0], c[1], c[2], x[0][p], x[1][p], x[2][p]);
            }
        }
        __syncwarp();                                   // IN is consumed: prefetch the next tile into it
        if (it + 1 < my_tiles) issue(nxt);
        bulk_wait_read0();                              // previous tile's stores (each lane its own group) have drained WORK
#pragma unroll
        for (int m = 0; m < 3; ++m) dct2_8(x[m]);
        __syncwarp();
#pragma unroll
        for (int m = 0; m < 3; ++m)
#pragma unroll
            for (int j = 0; j < 8; ++j) *reinterpret_cast<double *>(t_wr[j >> 1] + (m * 8 + j) * 64) = x[m][j];
        __syncwarp();
#pragma unroll
        for (int m = 0; m < 3; ++m) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double2 v = *reinterpret_cast<const double2 *>(t_rd[k] + m * 512);
                x[m][2 * k] = v.x;
                x[m][2 * k + 1] = v.y;
            }
            dct2_8(x[m]);
        }
        __syncwarp();
        for (int q = 0; q < nq; ++q) {                  // one transform, nq quantisations (nq = 1: the plain forward)
            if (q) { bulk_wait_read0(); __syncwarp(); } // the previous scale's stores have drained the staging area
            const double *rt_l = rt_l0 + q * 384, *t_l = rt_l + 192;
            QuantGuard qg;
            {
                double rtv[3][8];                           // loaded as one batch: the staging stores below would otherwise
    #pragma unroll                                          // serialise each (possibly aliasing) shared-memory load
                for (int m = 0; m < 3; ++m)
    #pragma unroll
                    for (int v = 0; v < 8; ++v) rtv[m][v] = rt_l[m * 64 + v * 8];
                int qv[3][8];
    #pragma unroll
                for (int m = 0; m < 3; ++m)
    #pragma unroll
                    for (int v = 0; v < 8; ++v) qv[m][v] = qg.q(x[m][v], rtv[m][v]);
    #pragma unroll
                for (int m = 0; m < 3; ++m)
    #pragma unroll
                    for (int v = 0; v < 8; ++v) *reinterpret_cast<int *>(zz_wr[v] + m * 256) = qv[m][v];
            }
            if (__builtin_expect(qg.risky(), 0)) {          // rare: redo this lane's 24 samples with the IEEE division
    #pragma unroll
                for (int m = 0; m < 3; ++m)
    #pragma unroll
                    for (int v = 0; v < 8; ++v)
                        *reinterpret_cast<int *>(zz_wr[v] + m * 256) = quantize_exact_f64(x[m][v], t_l[m * 64 + v * 8]);
            }
            fence_proxy_async();                            // make the staging visible to the bulk-copy unit
            __syncwarp();
            {
                const int b0 = cur.tx * 4, nb = min(4, g.Wp - b0);
                if (lane < nb) {                              // lane u stores the 3 scan blocks of image block u
                    int32_t *outf = a.out + q * a.out_q_stride + ((cur.frame * g.Hp + cur.by) * (int64_t)g.Wp + b0 + lane) * 192;
                    bulk_s2g(outf, work_s + lane * (kStageUF * 4), 768u);
                    bulk_commit();
                }
                if (a.zr_masks) {                             // scan block j = 3 u + m of the tile, in (h w c) order
                    unsigned long long mk;
                    zr_masks_from_staging<12>(work_b, [](int j) { return (j / 3) * (kStageUF * 4) + (j % 3) * 256; }, lane, mk);
                    if (lane < 3 * nb) {
                        const int64_t sblk = q * a.zr_q_stride + ((cur.frame * g.Hp + cur.by) * (int64_t)g.Wp + b0) * 3 + lane;
                        a.zr_masks[sblk] = mk;
                        a.zr_counts[sblk] = zr_block_count(mk);
                    }
                }
            }
        }
        cur = nxt;
    }
    bulk_wait_all0();
}

// SSE != 0 fuses the distortion measurement of a rate-distortion point into the decoder: the uint8 RGB original of
// the tile rides along on the same mbarrier (8 bulk rows, double-buffered: it is needed at the END of a tile while
// the next tile's copy is issued at its start), and each warp leaves one partial sum per tile --
//   SSE 1: sum (orig_rgb - ycbcr2rgb(rec))^2, what calc_psnr(img, symbols2image(...)) measures (metrics.py:3-40,
//          intracodec.py:139-141, color.py:39-63);   SSE 2: sum (rgb2ycbcr(orig_rgb) - rec)^2.
// A second kernel adds the partials of a frame in a fixed order (deterministic).  a.out == nullptr skips the
// reconstruction store: decode + PSNR then moves 15 bytes per pixel instead of 36 + 24 + 48 + 27.
constexpr int kInvIn = 4 * kStageU * 4;                                          // 3200 B of scan blocks
constexpr int kInvBuf = kInvIn + kWorkBytes;                                     // 9600 B
constexpr int kInvBufSse = kInvBuf + 2 * kRgbIn + 128;                           // + two 896-byte RGB tiles -> 11520 B
static_assert(kInvBufSse % 128 == 0, "alignment");

// RGB: the store applies ycbcr2rgb + clip (color.py:39-63) to every pixel first -- symbols2image's last step
// (intracodec.py:139-141) without a second pass over the image (row N1).
template <int SSE, bool RGB = false>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 2) k_inverse_c3_tma(const InvArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *s_tT = reinterpret_cast<double *>(smem_raw);                        // [192]
    unsigned long long *s_bar = reinterpret_cast<unsigned long long *>(smem_raw + 1536);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int kIn2 = kInvIn;
    constexpr int kBuf = SSE ? kInvBufSse : kInvBuf;
    unsigned char *in_b = smem_raw + 1664 + warp * kBuf;
    unsigned char *work_b = in_b + kIn2;
    const uint32_t bar = smem_u32(s_bar + warp), in_s = smem_u32(in_b), work_s = smem_u32(work_b);

    for (int i = threadIdx.x; i < 192; i += blockDim.x) {
        const int ch = i >> 6, k = i & 63, rr = k >> 3, jj = k & 7;
        s_tT[ch * 64 + jj * 8 + rr] = load_table_elem(a.table, a.table_dtype, i);
    }
    if (lane == 0) mbar_init(bar, 1);
    fence_mbar_init();
    __syncthreads();

    const int r = lane & 7, u = lane >> 3;
    const unsigned char *q_rd[8];
    unsigned char *t_wr[4], *o_wr;
    const unsigned char *t_rd[4];
#pragma unroll
    for (int j = 0; j < 8; ++j) q_rd[j] = in_b + (u * kStageU + ZZ_ORDER[r * 8 + j]) * 4;   // + m*256
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        t_wr[h] = work_b + u * (kTU2 * 8) + ((((r >> 1) ^ h) << 1) + (r & 1)) * 8;
        t_rd[h] = work_b + u * (kTU2 * 8) + r * 64 + ((h ^ (r >> 1)) << 4);
    }
    o_wr = work_b + ((8 * u + r) * 3) * 8;                                     // + i*784 + m*8
    const double *tq_l = s_tT + r;                      // table entry of raster (r, j), channel m: tq_l[m*64 + j*8]

    const TileGeom &g = a.g;
    const int64_t row_elems = g.W * 3, out_frame = g.H * row_elems;
    const int64_t gw = (int64_t)blockIdx.x * kWarpsPerCta + warp;
    const int64_t nw = (int64_t)gridDim.x * kWarpsPerCta;
    if (gw >= g.total_tiles) return;
    const int64_t my_tiles = (g.total_tiles - gw + nw - 1) / nw;
    TileIter cur, nxt;
    cur.init(g, gw, nw);
    nxt = cur;

    auto issue = [&](const TileIter &ti, uint32_t slot) {
        const int b0 = ti.tx * 4, nb = min(4, g.Wp - b0);
        const int32_t *zsrc = a.zz + ((ti.frame * g.Hp + ti.by) * (int64_t)g.Wp + b0) * 192;
        mbar_expect_tx(bar, (uint32_t)nb * (SSE ? 768u + 192u : 768u));         // + 8 rows x nb x 24 bytes of RGB
        for (int cu = 0; cu < nb; ++cu) bulk_g2s(in_s + cu * (kStageU * 4), zsrc + cu * 192, 768u, bar);
        if (SSE) {
            const unsigned char *osrc = a.orig + ti.frame * a.orig_frame_stride + (int64_t)ti.by * 8 * (g.W * 3) + (int64_t)ti.tx * 96;
            for (int row = 0; row < 8; ++row)                                    // nb is even (W % 16 == 0): 16-byte multiples
                bulk_g2s(work_s + kWorkBytes + slot * kRgbIn + row * kRgbPitch, osrc + row * (g.W * 3), (uint32_t)nb * 24u, bar);
        }
    };

    if (lane == 0) issue(cur, 0u);
    uint32_t parity = 0;
    for (int64_t it = 0; it < my_tiles; ++it, parity ^= 1u) {
        nxt.advance(g);
        mbar_wait(bar, parity);
        int q[3][8];
#pragma unroll
        for (int m = 0; m < 3; ++m)
#pragma unroll
            for (int j = 0; j < 8; ++j) q[m][j] = *reinterpret_cast<const int *>(q_rd[j] + m * 256);
        __syncwarp();
        if (lane == 0) {
            if (it + 1 < my_tiles) { fence_proxy_async(); issue(nxt, parity ^ 1u); }
            bulk_wait_read0();
        }
        double x[3][8];
        int mx = 0;
#pragma unroll
        for (int m = 0; m < 3; ++m)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const double p = __dmul_rn(i32_to_f64(q[m][j]), tq_l[m * 64 + j * 8]);
                mx = max(mx, __double2hiint(p) & 0x7fffffff);
                x[m][j] = trunc_f64_small(p);
            }
        if (__builtin_expect(mx >= 0x41E00000, 0)) {    // |q*t| >= 2^31: numpy's int32 cast saturates to INT_MIN
#pragma unroll
            for (int m = 0; m < 3; ++m)
#pragma unroll
                for (int j = 0; j < 8; ++j) x[m][j] = dequantize_f64(q[m][j], tq_l[m * 64 + j * 8]);
        }
#pragma unroll
        for (int m = 0; m < 3; ++m) dct3_8(x[m]);
        __syncwarp();
#pragma unroll
        for (int m = 0; m < 3; ++m)
#pragma unroll
            for (int j = 0; j < 8; ++j) *reinterpret_cast<double *>(t_wr[j >> 1] + (m * 8 + j) * 64) = x[m][j];
        __syncwarp();
#pragma unroll
        for (int m = 0; m < 3; ++m) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double2 v = *reinterpret_cast<const double2 *>(t_rd[k] + m * 512);
                x[m][2 * k] = v.x;
                x[m][2 * k + 1] = v.y;
            }
            dct3_8(x[m]);
        }
        __syncwarp();
        if (SSE) {
            // this lane holds pixel column 8u + r of the tile, rows 0..7, all three channels
            const unsigned char *ob = work_b + kWorkBytes + parity * kRgbIn + (8 * u + r) * 3;     // [2][8 rows][kRgbPitch]
            double acc = 0.0;
            if (u < min(4, g.Wp - cur.tx * 4)) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const double o0 = (double)ob[i * kRgbPitch], o1 = (double)ob[i * kRgbPitch + 1], o2 = (double)ob[i * kRgbPitch + 2];
                    double d0, d1, d2;
                    if (SSE == 1) {
                        double rr, gg, bb;
                        ycbcr2rgb_px(x[0][i], x[1][i], x[2][i], rr, gg, bb);
                        d0 = __dsub_rn(o0, rr); d1 = __dsub_rn(o1, gg); d2 = __dsub_rn(o2, bb);
                    } else {
                        double yy, cb, cr;
                        rgb2ycbcr_px(o0, o1, o2, yy, cb, cr);
                        d0 = __dsub_rn(yy, x[0][i]); d1 = __dsub_rn(cb, x[1][i]); d2 = __dsub_rn(cr, x[2][i]);
                    }
                    acc = __dadd_rn(acc, __dmul_rn(d0, d0));
                    acc = __dadd_rn(acc, __dmul_rn(d1, d1));
                    acc = __dadd_rn(acc, __dmul_rn(d2, d2));
                }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, off));
            if (lane == 0) a.sse_partial[(cur.frame * g.Hp + cur.by) * (int64_t)g.tiles_per_row + cur.tx] = acc;
        }
        if (!SSE || a.out) {
            if (RGB) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    double rr, gg, bb;
                    ycbcr2rgb_px(x[0][i], x[1][i], x[2][i], rr, gg, bb);
                    x[0][i] = rr; x[1][i] = gg; x[2][i] = bb;
                }
            }
#pragma unroll
            for (int m = 0; m < 3; ++m)
#pragma unroll
                for (int i = 0; i < 8; ++i) *reinterpret_cast<double *>(o_wr + i * (kRowPitch * 8) + m * 8) = x[m][i];
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                const int nb = min(4, g.Wp - cur.tx * 4);
                double *dst = a.out + cur.frame * out_frame + (int64_t)cur.by * 8 * row_elems + (int64_t)cur.tx * 96;
                const uint32_t row_bytes = (uint32_t)nb * 192u;
#pragma unroll
                for (int row = 0; row < 8; ++row) bulk_s2g(dst + row * row_elems, work_s + row * (kRowPitch * 8), row_bytes);
                bulk_commit();
            }
        }
        cur = nxt;
    }
    if (lane == 0) bulk_wait_all0();
}

// sum the tile partials of each frame in a fixed order: one CTA per frame, thread t adds partials t, t + 256, ... in
// four interleaved chains, then a fixed shuffle / shared-memory tree (the order does not depend on the batch)
__global__ void __launch_bounds__(256) k_sum_tile_partials(const double *__restrict__ partial, int64_t tiles_per_frame,
                                                           double *__restrict__ out) {
    __shared__ double s_w[8];
    const double *p = partial + (int64_t)blockIdx.x * tiles_per_frame;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int64_t c = threadIdx.x;
    for (; c + 768 < tiles_per_frame; c += 1024) {
        a0 = __dadd_rn(a0, p[c]); a1 = __dadd_rn(a1, p[c + 256]); a2 = __dadd_rn(a2, p[c + 512]); a3 = __dadd_rn(a3, p[c + 768]);
    }
    for (; c < tiles_per_frame; c += 256) a0 = __dadd_rn(a0, p[c]);
    double acc = __dadd_rn(__dadd_rn(a0, a1), __dadd_rn(a2, a3));
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, off));
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = s_w[0];
#pragma unroll
        for (int w = 1; w < 8; ++w) t = __dadd_rn(t, s_w[w]);
        out[blockIdx.x] = t;
    }
}

// ================================================================================================
// v2 of the P-frame kernels (K1p / K2p): the same per-warp pipeline, with the motion-compensated
// prediction gathered asynchronously.  Per warp: IN (current rows / scan blocks, TMA bulk copies),
// PRED (8 x 96 doubles: cp.async 8-byte copies from ref at the block's motion vector -- sources are
// only 8-byte aligned; zero-fill when the source window leaves the frame, motion.py:90-92) and WORK.
// Both land on one mbarrier (1 expect_tx arrival + 32 deferred cp.async arrivals).  The vectors of
// tile i+2 are fetched while tile i is computed, so no global load is ever waited on.
// One CTA of 12 warps per SM (3200 + 12 * 18944 bytes of shared memory).
// ================================================================================================
constexpr int kPfWarps = 12;
constexpr int kPfBuf = 2 * kInBytes + kWorkBytes;        // 18944
static_assert(kPfBuf % 128 == 0, "alignment");

__device__ __forceinline__ void cp_async8_zfill(uint32_t dst, const void *src, uint32_t src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

// Issue the gather of one tile's prediction into PRED.  Copy k of a lane (k = 0..23) is pixel row
// k/3, tile column 32*(k%3) + lane, so a lane touches three blocks per tile.
struct PredGather {
    const double *ref;       // frame base is added per tile
    int64_t H, W;
    int sr;
    __device__ __forceinline__ void issue(uint32_t pred_s, uint32_t bar, int lane, int64_t mvidx, int nb,
                                          const double *ref_frame, int by, int b0) const {
        int dy = 0, dx = 0;
        if (lane < nb) mv_decode(mvidx, sr, dy, dx);
        const int64_t pitch = W * 8;                                          // bytes per frame row
        const int Hi = (int)H, Wi = (int)W;
#pragma unroll
        for (int c3 = 0; c3 < 3; ++c3) {
            const int col = 32 * c3 + lane, blk = col >> 3;
            const int bdy = __shfl_sync(0xffffffffu, dy, blk), bdx = __shfl_sync(0xffffffffu, dx, blk);
            if (blk < nb) {
                const int sy = by * 8 + bdy, sx = (b0 + blk) * 8 + bdx;
                const bool ok = sy >= 0 && sy <= Hi - 8 && sx >= 0 && sx <= Wi - 8;
                const char *src = reinterpret_cast<const char *>(ref_frame) + (ok ? (int64_t)sy * pitch + (int64_t)(sx + (col & 7)) * 8 : 0);
                const int64_t step = ok ? pitch : 0;
                const uint32_t nbytes = ok ? 8u : 0u, dst = pred_s + col * 8;
#pragma unroll
                for (int row = 0; row < 8; ++row) {
                    cp_async8_zfill(dst + row * (kRowPitch * 8), src, nbytes);
                    src += step;
                }
            }
        }
        cp_async_mbar_arrive(bar);
    }
};

__global__ void __launch_bounds__(kPfWarps * 32, 1) k_pframe_forward_tma(const FwdArgs a) {
    if (a.run_flag && *a.run_flag == 0) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *s_rt = reinterpret_cast<double *>(smem_raw);                       // [192]
    double *s_t = s_rt + 192;                                                   // [192]
    unsigned long long *s_bar = reinterpret_cast<unsigned long long *>(smem_raw + 3072);   // [12]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *in_b = smem_raw + 3200 + warp * kPfBuf;
    unsigned char *pred_b = in_b + kInBytes;
    unsigned char *work_b = pred_b + kInBytes;
    const uint32_t bar = smem_u32(s_bar + warp), in_s = smem_u32(in_b), pred_s = smem_u32(pred_b), work_s = smem_u32(work_b);

    for (int i = threadIdx.x; i < 192; i += blockDim.x) {
        const double t = load_table_elem(a.table, a.table_dtype, i);
        s_t[i] = t;
        s_rt[i] = __drcp_rn(t);
    }
    if (lane == 0) mbar_init(bar, 33);
    fence_mbar_init();
    __syncthreads();

    const int r = lane & 7, u = lane >> 3;
    const unsigned char *rd_in = in_b + r * (kRowPitch * 8) + u * 192;
    const unsigned char *rd_pr = pred_b + r * (kRowPitch * 8) + u * 192;
    unsigned char *t_wr[4], *zz_wr[8];
    const unsigned char *t_rd[4];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        t_wr[h] = work_b + u * (kTU2 * 8) + ((((r >> 1) ^ h) << 1) + (r & 1)) * 8;
        t_rd[h] = work_b + u * (kTU2 * 8) + r * 64 + ((h ^ (r >> 1)) << 4);
    }
#pragma unroll
    for (int v = 0; v < 8; ++v) zz_wr[v] = work_b + (u * kStageU + ZZ_ORDER[v * 8 + r]) * 4;   // + ch*256 (+3200 for region B)
    const double *rt_l = s_rt + r, *t_l = s_t + r;

    const TileGeom &g = a.g;
    const int64_t row_elems = g.W, frame_elems = g.H * g.W;
    const int64_t gw = (int64_t)blockIdx.x * kPfWarps + warp;
    const int64_t nw = (int64_t)gridDim.x * kPfWarps;
    if (gw >= g.total_tiles) return;
    const int64_t my_tiles = (g.total_tiles - gw + nw - 1) / nw;
    TileIter cur, nxt, nx2;
    cur.init(g, gw, nw);
    nxt = cur;
    PredGather pg{a.ref, g.H, g.W, a.sr};

    auto load_mv = [&](const TileIter &ti) -> int64_t {          // lanes 0..11: vector of block (b0 + lane)
        const int b0 = ti.tx * 12;
        return (lane < min(12, g.Wp - b0)) ? a.mv[(ti.frame * g.Hp + ti.by) * (int64_t)g.Wp + b0 + lane] : 0;
    };
    auto issue = [&](const TileIter &ti, int64_t mvidx) {        // whole warp
        const int b0 = ti.tx * 12, nb = min(12, g.Wp - b0);
        const uint32_t row_bytes = (uint32_t)nb * 64u;
        fence_proxy_async();
        if (lane == 0) mbar_expect_tx(bar, 8u * row_bytes);
        __syncwarp();
        if (lane < 8) {                                     // lane r copies current-frame row r
            const double *src = a.img + ti.frame * a.frame_stride + ((int64_t)ti.by * 8 + lane) * row_elems + (int64_t)b0 * 8;
            bulk_g2s(in_s + lane * (kRowPitch * 8), src, row_bytes, bar);
        }
        pg.issue(pred_s, bar, lane, mvidx, nb, a.ref + ti.frame * frame_elems, ti.by, b0);
    };

    int64_t mv_nxt = load_mv(cur);
    issue(cur, mv_nxt);
    nxt.advance(g);
    mv_nxt = (my_tiles > 1) ? load_mv(nxt) : 0;
    nx2 = nxt;
    uint32_t parity = 0;
    for (int64_t it = 0; it < my_tiles; ++it, parity ^= 1u) {
        nx2.advance(g);
        const int b0 = cur.tx * 12, nb = min(12, g.Wp - b0);
        mbar_wait(bar, parity);
        double x[3][8];
        {
#pragma unroll
            for (int k = 0; k < 12; ++k) {
                const double2 c = *reinterpret_cast<const double2 *>(rd_in + 16 * k);
                const double2 p = *reinterpret_cast<const double2 *>(rd_pr + 16 * k);
                if (a.pred_out && (3 * u + (k >> 2)) < nb)
                    stg_stream(a.pred_out + cur.frame * frame_elems + ((int64_t)cur.by * 8 + r) * row_elems +
                                   (int64_t)b0 * 8 + 24 * u + 2 * k, p);
                x[k >> 2][(2 * k) & 7] = __dsub_rn(c.x, p.x);              // residual = cur - prediction
                x[k >> 2][(2 * k + 1) & 7] = __dsub_rn(c.y, p.y);
            }
        }
        __syncwarp();                                   // IN / PRED are consumed
        const int64_t mv_cur_next = mv_nxt;
        if (it + 1 < my_tiles) {
            issue(nxt, mv_cur_next);
            mv_nxt = (it + 2 < my_tiles) ? load_mv(nx2) : 0;      // consumed one iteration later
        }
#pragma unroll
        for (int m = 0; m < 3; ++m) dct2_8(x[m]);
        bulk_wait_read0();                              // previous tile's stores have drained WORK
        __syncwarp();
#pragma unroll
        for (int m = 0; m < 3; ++m)
#pragma unroll
            for (int j = 0; j < 8; ++j) *reinterpret_cast<double *>(t_wr[j >> 1] + (m * 8 + j) * 64) = x[m][j];
        __syncwarp();
#pragma unroll
        for (int m = 0; m < 3; ++m) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double2 v = *reinterpret_cast<const double2 *>(t_rd[k] + m * 512);
                x[m][2 * k] = v.x;
                x[m][2 * k + 1] = v.y;
            }
            dct2_8(x[m]);
        }
        __syncwarp();
        // numpy broadcasting: the single luma channel is quantised with all three tables
        // (patchquant.py:59).  One round per sub-block m; staging regions A/B/A.
        const int och64 = 64 * a.och;
        int32_t *outf = a.out + ((cur.frame * g.Hp + cur.by) * (int64_t)g.Wp + b0) * och64;
#pragma unroll
        for (int m = 0; m < 3; ++m) {
            const int reg = (m & 1) * 3200;
            if (m == 2) { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); __syncwarp(); }
            QuantGuard qg;
            {
                double rtv[3][8];
#pragma unroll
                for (int ch = 0; ch < 3; ++ch)
#pragma unroll
                    for (int v = 0; v < 8; ++v) rtv[ch][v] = rt_l[ch * 64 + v * 8];
                int qv[3][8];
#pragma unroll
                for (int ch = 0; ch < 3; ++ch)
#pragma unroll
                    for (int v = 0; v < 8; ++v) qv[ch][v] = qg.q(x[m][v], rtv[ch][v]);
#pragma unroll
                for (int ch = 0; ch < 3; ++ch)
#pragma unroll
                    for (int v = 0; v < 8; ++v) *reinterpret_cast<int *>(zz_wr[v] + reg + ch * 256) = qv[ch][v];
            }
            if (__builtin_expect(qg.risky(), 0)) {
#pragma unroll
                for (int ch = 0; ch < 3; ++ch)
#pragma unroll
                    for (int v = 0; v < 8; ++v)
                        *reinterpret_cast<int *>(zz_wr[v] + reg + ch * 256) = quantize_exact_f64(x[m][v], t_l[ch * 64 + v * 8]);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane < 4 && 3 * lane + m < nb) bulk_s2g(outf + (3 * lane + m) * och64, work_s + reg + lane * (kStageU * 4), 256u * (uint32_t)a.och);
            bulk_commit();                                // every lane commits (possibly empty) groups: counts stay in step
        }
        cur = nxt;
        nxt = nx2;
    }
    bulk_wait_all0();
}

// K2p: per warp IN (scan blocks, 3200 B) + PRED x2 (double-buffered so that the prediction can be read
// at the very end of a tile while the next tile's gather is already in flight) + WORK; 10 warps per SM.
constexpr int kPiWarps = 10;
constexpr int kPiIn = 4 * kStageU * 4;                         // 3200
constexpr int kPiBuf = kPiIn + 2 * kInBytes + kWorkBytes;      // 22144
static_assert(kPiBuf % 128 == 0, "alignment");

__global__ void __launch_bounds__(kPiWarps * 32, 1) k_pframe_inverse_tma(const InvArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *s_tT = reinterpret_cast<double *>(smem_raw);                        // [64] luminance table, transposed
    unsigned long long *s_bar = reinterpret_cast<unsigned long long *>(smem_raw + 3072);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *in_b = smem_raw + 3200 + warp * kPiBuf;
    unsigned char *pred_b = in_b + kPiIn;                                       // two buffers of kInBytes
    unsigned char *work_b = pred_b + 2 * kInBytes;
    const uint32_t bar = smem_u32(s_bar + warp), in_s = smem_u32(in_b), pred_s = smem_u32(pred_b), work_s = smem_u32(work_b);

    for (int i = threadIdx.x; i < 64; i += blockDim.x) s_tT[(i & 7) * 8 + (i >> 3)] = load_table_elem(a.table, a.table_dtype, i);
    const bool gather = a.pred == nullptr;
    if (lane == 0) mbar_init(bar, gather ? 33 : 1);
    fence_mbar_init();
    __syncthreads();

    const int r = lane & 7, u = lane >> 3;
    const unsigned char *q_rd[8];
    unsigned char *t_wr[4], *o_wr;
    const unsigned char *t_rd[4], *p_rd;
#pragma unroll
    for (int j = 0; j < 8; ++j) q_rd[j] = in_b + (u * kStageU + ZZ_ORDER[r * 8 + j]) * 4;   // + m*256
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        t_wr[h] = work_b + u * (kTU2 * 8) + ((((r >> 1) ^ h) << 1) + (r & 1)) * 8;
        t_rd[h] = work_b + u * (kTU2 * 8) + r * 64 + ((h ^ (r >> 1)) << 4);
    }
    o_wr = work_b + (24 * u + r) * 8;                                          // + i*784 + m*64
    p_rd = pred_b + (24 * u + r) * 8;                                          // + buffer*kInBytes
    const double *tq_l = s_tT + r;                                             // tq_l[j*8] = lum[r][j]

    const TileGeom &g = a.g;
    const int64_t row_elems = g.W, frame_elems = g.H * g.W;
    const int64_t gw = (int64_t)blockIdx.x * kPiWarps + warp;
    const int64_t nw = (int64_t)gridDim.x * kPiWarps;
    if (gw >= g.total_tiles) return;
    const int64_t my_tiles = (g.total_tiles - gw + nw - 1) / nw;
    TileIter cur, nxt, nx2;
    cur.init(g, gw, nw);
    nxt = cur;
    PredGather pg{a.ref, g.H, g.W, a.sr};

    auto load_mv = [&](const TileIter &ti) -> int64_t {
        const int b0 = ti.tx * 12;
        return (gather && lane < min(12, g.Wp - b0)) ? a.mv[(ti.frame * g.Hp + ti.by) * (int64_t)g.Wp + b0 + lane] : 0;
    };
    auto issue = [&](const TileIter &ti, int64_t mvidx, uint32_t pbuf_s) {
        const int b0 = ti.tx * 12, nb = min(12, g.Wp - b0);
        // the bulk copies of a tile are issued by different lanes in parallel (lane b: scan block b,
        // lanes 16..23: prediction rows) after lane 0 has armed the barrier
        const uint32_t row_bytes = (uint32_t)nb * 64u;
        fence_proxy_async();
        if (lane == 0) mbar_expect_tx(bar, (uint32_t)nb * 256u + (gather ? 0u : 8u * row_bytes));
        __syncwarp();
        if (lane < nb) {                                  // scan channel 0 of block `lane` (stride Czz*64 ints)
            const int32_t *zsrc = a.zz + ((ti.frame * g.Hp + ti.by) * (int64_t)g.Wp + b0 + lane) * a.Czz * 64;
            const int cu = lane / 3;
            bulk_g2s(in_s + (cu * kStageU + (lane - cu * 3) * 64) * 4, zsrc, 256u, bar);
        }
        if (!gather && lane >= 16 && lane < 24) {
            const int row = lane - 16;
            const double *psrc = a.pred + ti.frame * frame_elems + ((int64_t)ti.by * 8 + row) * row_elems + (int64_t)b0 * 8;
            bulk_g2s(pbuf_s + row * (kRowPitch * 8), psrc, row_bytes, bar);
        }
        if (gather) pg.issue(pbuf_s, bar, lane, mvidx, nb, a.ref + ti.frame * frame_elems, ti.by, b0);
    };

    int64_t mv_nxt = load_mv(cur);
    issue(cur, mv_nxt, pred_s);
    nxt.advance(g);
    mv_nxt = (my_tiles > 1) ? load_mv(nxt) : 0;
    nx2 = nxt;
    uint32_t parity = 0;
    for (int64_t it = 0; it < my_tiles; ++it, parity ^= 1u) {
        nx2.advance(g);
        const int b0 = cur.tx * 12, nb = min(12, g.Wp - b0);
        const int pb = (int)(it & 1);
        mbar_wait(bar, parity);
        int q[3][8];
#pragma unroll
        for (int m = 0; m < 3; ++m)
#pragma unroll
            for (int j = 0; j < 8; ++j) q[m][j] = *reinterpret_cast<const int *>(q_rd[j] + m * 256);
        __syncwarp();
        const int64_t mv_cur_next = mv_nxt;
        if (it + 1 < my_tiles) {
            issue(nxt, mv_cur_next, pred_s + (uint32_t)((pb ^ 1) * kInBytes));
            mv_nxt = (it + 2 < my_tiles) ? load_mv(nx2) : 0;
        }
        double x[3][8];
        int mx = 0;
#pragma unroll
        for (int m = 0; m < 3; ++m)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const double p = __dmul_rn(i32_to_f64(q[m][j]), tq_l[j * 8]);       // luminance table for every block
                mx = max(mx, __double2hiint(p) & 0x7fffffff);
                x[m][j] = trunc_f64_small(p);
            }
        if (__builtin_expect(mx >= 0x41E00000, 0)) {
#pragma unroll
            for (int m = 0; m < 3; ++m)
#pragma unroll
                for (int j = 0; j < 8; ++j) x[m][j] = dequantize_f64(q[m][j], tq_l[j * 8]);
        }
#pragma unroll
        for (int m = 0; m < 3; ++m) dct3_8(x[m]);
        bulk_wait_read0();                              // every lane drains its own bulk-store group
        __syncwarp();
#pragma unroll
        for (int m = 0; m < 3; ++m)
#pragma unroll
            for (int j = 0; j < 8; ++j) *reinterpret_cast<double *>(t_wr[j >> 1] + (m * 8 + j) * 64) = x[m][j];
        __syncwarp();
#pragma unroll
        for (int m = 0; m < 3; ++m) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double2 v = *reinterpret_cast<const double2 *>(t_rd[k] + m * 512);
                x[m][2 * k] = v.x;
                x[m][2 * k + 1] = v.y;
            }
            dct3_8(x[m]);
        }
        __syncwarp();
        {
            double pr[3][8];                             // prediction of column r of the three blocks (loaded as a batch)
            const unsigned char *pp = p_rd + pb * kInBytes;
#pragma unroll
            for (int m = 0; m < 3; ++m)
#pragma unroll
                for (int i = 0; i < 8; ++i) pr[m][i] = *reinterpret_cast<const double *>(pp + i * (kRowPitch * 8) + m * 64);
#pragma unroll
            for (int m = 0; m < 3; ++m)
#pragma unroll
                for (int i = 0; i < 8; ++i)              // recon = prediction + recon_residual (videocodec.py:74)
                    *reinterpret_cast<double *>(o_wr + i * (kRowPitch * 8) + m * 64) = __dadd_rn(pr[m][i], x[m][i]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane < 8) {                                   // lane r stores output row r
            double *dst = a.out + cur.frame * frame_elems + ((int64_t)cur.by * 8 + lane) * row_elems + (int64_t)b0 * 8;
            bulk_s2g(dst, work_s + lane * (kRowPitch * 8), (uint32_t)nb * 64u);
            bulk_commit();
        }
        cur = nxt;
        nxt = nx2;
    }
    bulk_wait_all0();
}

// ================================================================================================
// v3 of the P-frame kernels (K1p / K2p): the motion-compensated prediction is gathered by the TMA
// unit itself.  The reference planes are described by ONE tensor map (W x H x frames, float64); a
// block's prediction is a single box copy landing on the tile's mbarrier.  The unit faults ("illegal
// instruction") when a box starts at an address that is not a multiple of 16 bytes -- measured here:
// every odd dx -- so the box is 10 x 8 and starts at the even column sx & ~1; the 8 wanted columns
// begin at element sx & 1 of each 80-byte row (80 == 16 * 5 mod 128: the row reads are conflict-free
// without any swizzle).  A window that leaves the frame must give an all-zero block
// (motion.py:90-92), not a partially filled one: such a block asks for a box that lies entirely
// outside the tensor and the unit zero-fills it.
// Against v2 this removes 24 cp.async + their address arithmetic per lane and tile, and the tile
// shrinks to 8 blocks (lane (r,u): blocks u and u + 4): 13.4 / 11.5 KB of shared memory per warp and
// <= 128 registers, i.e. 16 warps per SM instead of 12 / 10.
// The eight boxes of a tile are issued by ONE elected lane from warp-uniform operands (the block
// coordinates are broadcast with shuffles; blocks past the right edge fetch a zero box so that the
// sequence has no branches): eight back-to-back UTMALDG on distinct uniform registers.  Issued from
// eight lanes, the compiler serialises them in a loop whose every turn waits for the previous copy
// to release its operands (K1p 0.365 -> 0.345 ms per 32 frames).  The other transfers of a tile
// stay 1-D bulk copies (current rows, scan blocks, stores): moving them to tensor boxes as well
// (66 x 8 row box, {36,6,4,1} / {66,1,8,1} clipped stores: 11 instead of 24 bulk operations per
// tile) measured SLOWER (K1p 0.355, K2p 0.245 against 0.345 / 0.228 ms) and was dropped.
// ================================================================================================
constexpr int kP3Blocks = 8;
constexpr int kP3Pitch = 528;                         // bytes per IN row: 8 blocks * 64 + 16 (== 16 mod 128)
constexpr int kP3Box = 640;                           // one 10 x 8 box of doubles (80-byte rows)
constexpr int kP3Pred = kP3Blocks * kP3Box;           // 5120: block-major boxes
constexpr int kP3In = 8 * kP3Pitch;                   // 4224
constexpr int kP3Trans = 4 * kP3TU * 8;               // 4352
constexpr int kP3Region = 4 * kStageUF * 4;           // 3264: one round of scan staging (4 blocks x 3 tables)
constexpr int kP3Header = 3584;                       // tables + barriers (a multiple of 128)
constexpr int kP3ZzBlk = 288;                         // K2p IN: one scan block (64 ints) + 8 pad (blocks land 32 B apart mod 128)
constexpr int kP3ZzIn = kP3Blocks * kP3ZzBlk;         // 2304
constexpr int kP3FwdBuf = (kP3Pred + kP3In + kP3Trans + 127) / 128 * 128;      // 13696
static_assert(kP3Region <= kP3Trans, "scan staging reuses the transposition buffer");
constexpr int kP3InvBuf = (kP3Pred + kP3ZzIn + kP3Trans + 127) / 128 * 128;    // 11776
static_assert(kP3Box % 128 == 0 && kP3FwdBuf % 128 == 0 && kP3InvBuf % 128 == 0 && kP3Header % 128 == 0, "tensor copies need 128-byte aligned shared addresses");
static_assert(kP3In <= kP3Trans, "the output row tile reuses the transposition buffer");

__device__ __forceinline__ void tma_box_g2s(uint32_t dst, const CUtensorMap *tm, int x, int y, int z, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(tm), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
}
// one lane of the (converged) warp, chosen by the hardware: tells the compiler that what follows runs once per warp,
// so bulk copies with warp-uniform operands are issued straight from uniform registers
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{ .reg .pred p; elect.sync _|p, 0xffffffff; selp.u32 %0, 1, 0, p; }" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ int uniform_warp_idx() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }

// Lanes 8..15: box coordinates of block (lane - 8) of a tile from its vector (motion.py:83-92).  bx is the even
// column the box starts at, par the column parity sx & 1 of the block's window inside its box.
__device__ __forceinline__ void p3_box(int lane, int nb, int64_t mvidx, int sr, int Hi, int Wi, int by, int b0,
                                       int &bx, int &byy, int &par) {
    const int b = lane - 8;
    bx = -16; byy = -8; par = 0;                          // a box entirely outside the tensor: zero-filled
    if (b >= 0 && b < nb) {
        int dy, dx;
        mv_decode(mvidx, sr, dy, dx);
        const int sy = by * 8 + dy, sx = (b0 + b) * 8 + dx;
        if (sy >= 0 && sy <= Hi - 8 && sx >= 0 && sx <= Wi - 8) { par = sx & 1; bx = sx - par; byy = sy; }
    }
}
// Whole warp: broadcast the eight boxes' coordinates; one elected lane issues the copies from warp-uniform operands.
// Always eight copies (blocks past the right edge fetch a zero box): one basic block, so the copies sit on distinct
// uniform registers and do not wait on one another.
__device__ __forceinline__ void p3_gather(const CUtensorMap *tm, uint32_t pred_s, uint32_t bar, int bx, int byy, int frame) {
    int ux[kP3Blocks], uy[kP3Blocks];
#pragma unroll
    for (int b = 0; b < kP3Blocks; ++b) {
        ux[b] = __shfl_sync(0xffffffffu, bx, 8 + b);
        uy[b] = __shfl_sync(0xffffffffu, byy, 8 + b);
    }
    if (elect_one()) {
#pragma unroll
        for (int b = 0; b < kP3Blocks; ++b) tma_box_g2s(pred_s + b * kP3Box, tm, ux[b], uy[b], frame, bar);
    }
}

template <int WARPS, int CTAS>
__global__ void __launch_bounds__(WARPS * 32, CTAS) k_pframe_forward_tm(const FwdArgs a, const __grid_constant__ CUtensorMap tm_ref) {
    if (a.run_flag && *a.run_flag == 0) return;
    constexpr int kBuf = kP3FwdBuf;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *s_rt = reinterpret_cast<double *>(smem_raw);                       // [192]
    double *s_t = s_rt + 192;                                                   // [192]
    unsigned long long *s_bar = reinterpret_cast<unsigned long long *>(smem_raw + 3072);   // [WARPS]
    const int lane = threadIdx.x & 31, warp = uniform_warp_idx();
    unsigned char *pred_b = smem_raw + kP3Header + warp * kBuf;
    unsigned char *in_b = pred_b + kP3Pred;
    unsigned char *work_b = in_b + kP3In;
    const uint32_t bar = smem_u32(s_bar + warp), in_s = smem_u32(in_b), pred_s = smem_u32(pred_b), work_s = smem_u32(work_b);

    for (int i = threadIdx.x; i < 192; i += blockDim.x) {
        const double t = load_table_elem(a.table, a.table_dtype, i);
        s_t[i] = t;
        s_rt[i] = __drcp_rn(t);
    }
    if (lane == 0) mbar_init(bar, 1);
    fence_mbar_init();
    // PatchQuant's table is [lum, chrom, chrom] (patchquant.py:40): when tables 1 and 2 hold the same bits, the third
    // quantised copy of a block IS the second -- computed once, stored twice
    bool same = true;
    if (threadIdx.x < 64) {
        const double t1 = load_table_elem(a.table, a.table_dtype, 64 + threadIdx.x), t2 = load_table_elem(a.table, a.table_dtype, 128 + threadIdx.x);
        same = __double_as_longlong(t1) == __double_as_longlong(t2);
    }
    const bool chroma_twice = __syncthreads_and(same) != 0;

    const int r = lane & 7, u = lane >> 3;
    const unsigned char *rd_in = in_b + r * kP3Pitch + u * 64;                  // + m*256 + 16k
    const unsigned char *rd_pr = pred_b + u * kP3Box + r * 80;                  // + m*4*kP3Box + parity*8 + 8e
    unsigned char *t_wr[4], *zz_wr[8];
    const unsigned char *t_rd[4];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        t_wr[h] = work_b + u * (kP3TU * 8) + ((((r >> 1) ^ h) << 1) + (r & 1)) * 8;    // + row*64
        t_rd[h] = work_b + u * (kP3TU * 8) + r * 64 + ((h ^ (r >> 1)) << 4);           // + m*512
    }
#pragma unroll
    for (int v = 0; v < 8; ++v) zz_wr[v] = work_b + (u * kStageUF + ZZ_ORDER[v * 8 + r]) * 4;   // + ch*256
    const double *rt_l = s_rt + r, *t_l = s_t + r;

    const TileGeom &g = a.g;
    const int64_t row_elems = g.W, frame_elems = g.H * g.W;
    const int Hi = (int)g.H, Wi = (int)g.W;
    const int64_t gw = (int64_t)blockIdx.x * WARPS + warp;
    const int64_t nw = (int64_t)gridDim.x * WARPS;
    if (gw >= g.total_tiles) return;
    const int64_t my_tiles = (g.total_tiles - gw + nw - 1) / nw;
    TileIter cur, nxt, nx2;
    cur.init(g, gw, nw);
    nxt = cur;

    auto load_mv = [&](const TileIter &ti) -> int64_t {          // lanes 8..15: vector of block (b0 + lane - 8)
        const int b0 = ti.tx * kP3Blocks, b = lane - 8;
        return (b >= 0 && b < min(kP3Blocks, g.Wp - b0)) ? a.mv[(ti.frame * g.Hp + ti.by) * (int64_t)g.Wp + b0 + b] : 0;
    };
    auto issue = [&](const TileIter &ti, int64_t mvidx) -> int {  // whole warp; returns the parity (lanes 8..15)
        const int b0 = ti.tx * kP3Blocks, nb = min(kP3Blocks, g.Wp - b0);
        int bx, byy, par;
        p3_box(lane, nb, mvidx, a.sr, Hi, Wi, ti.by, b0, bx, byy, par);
        fence_proxy_async();
        if (lane == 0) mbar_expect_tx(bar, (uint32_t)nb * 512u + (uint32_t)kP3Pred);   // 8 rows of nb*64 bytes + 8 boxes
        __syncwarp();
        if (lane < 8)                                           // lane r copies current-frame row r
            bulk_g2s(in_s + lane * kP3Pitch, a.img + ti.frame * a.frame_stride + ((int64_t)ti.by * 8 + lane) * row_elems + (int64_t)b0 * 8,
                     (uint32_t)nb * 64u, bar);
        p3_gather(&tm_ref, pred_s, bar, bx, byy, (int)ti.frame);
        return par;
    };

    int64_t mv_nxt = load_mv(cur);
    int par_l = issue(cur, mv_nxt);                  // lanes 8..15: column parity of the tile in flight
    nxt.advance(g);
    mv_nxt = (my_tiles > 1) ? load_mv(nxt) : 0;
    nx2 = nxt;
    uint32_t parity = 0;
    for (int64_t it = 0; it < my_tiles; ++it, parity ^= 1u) {
        nx2.advance(g);
        const int b0 = cur.tx * kP3Blocks, nb = min(kP3Blocks, g.Wp - b0);
        mbar_wait(bar, parity);
        double x[2][8];
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            const unsigned char *pp = rd_pr + m * (4 * kP3Box) + 8 * __shfl_sync(0xffffffffu, par_l, 8 + u + 4 * m);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double2 c = *reinterpret_cast<const double2 *>(rd_in + m * 256 + 16 * k);
                double2 p;
                p.x = *reinterpret_cast<const double *>(pp + 16 * k);
                p.y = *reinterpret_cast<const double *>(pp + 16 * k + 8);
                if (a.pred_out && (u + 4 * m) < nb)
                    stg_stream(a.pred_out + cur.frame * frame_elems + ((int64_t)cur.by * 8 + r) * row_elems +
                                   (int64_t)(b0 + u + 4 * m) * 8 + 2 * k, p);
                x[m][2 * k] = __dsub_rn(c.x, p.x);                       // residual = cur - prediction
                x[m][2 * k + 1] = __dsub_rn(c.y, p.y);
            }
        }
        __syncwarp();                                   // IN / PRED are consumed
        const int64_t mv_cur_next = mv_nxt;
        if (it + 1 < my_tiles) {
            par_l = issue(nxt, mv_cur_next);
            mv_nxt = (it + 2 < my_tiles) ? load_mv(nx2) : 0;      // consumed one iteration later
        }
#pragma unroll
        for (int m = 0; m < 2; ++m) dct2_8(x[m]);
        bulk_wait_read0();                              // previous tile's stores have drained WORK
        __syncwarp();
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int j = 0; j < 8; ++j) *reinterpret_cast<double *>(t_wr[j >> 1] + (m * 8 + j) * 64) = x[m][j];
        __syncwarp();
#pragma unroll
        for (int m = 0; m < 2; ++m) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double2 v = *reinterpret_cast<const double2 *>(t_rd[k] + m * 512);
                x[m][2 * k] = v.x;
                x[m][2 * k + 1] = v.y;
            }
            dct2_8(x[m]);
        }
        __syncwarp();
        // numpy broadcasting: the single luma channel is quantised with all three tables
        // (patchquant.py:59).  One round per sub-block m.
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            if (m == 1) { bulk_wait_read0(); __syncwarp(); }     // round 0's stores have drained the staging area
            const auto quantise = [&](auto nch_c) {       // NCH tables computed; the last one also stored as channel 2
                constexpr int NCH = decltype(nch_c)::value;
                QuantGuard qg;
                double rtv[NCH][8];
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
                    for (int v = 0; v < 8; ++v) rtv[ch][v] = rt_l[ch * 64 + v * 8];
                int qv[NCH][8];
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
                    for (int v = 0; v < 8; ++v) qv[ch][v] = qg.q(x[m][v], rtv[ch][v]);
                if (__builtin_expect(qg.risky(), 0)) {
#pragma unroll
                    for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
                        for (int v = 0; v < 8; ++v) qv[ch][v] = quantize_exact_f64(x[m][v], t_l[ch * 64 + v * 8]);
                }
#pragma unroll
                for (int ch = 0; ch < 3; ++ch)
#pragma unroll
                    for (int v = 0; v < 8; ++v) *reinterpret_cast<int *>(zz_wr[v] + ch * 256) = qv[ch < NCH ? ch : NCH - 1][v];
            };
            if (chroma_twice || a.och == 2) quantise(std::integral_constant<int, 2>{});
            else quantise(std::integral_constant<int, 3>{});
            fence_proxy_async();
            __syncwarp();
            if (lane < 4 && lane + 4 * m < nb)            // lane u stores the och (3, or 2) scan blocks of image block u + 4m
                bulk_s2g(a.out + ((cur.frame * g.Hp + cur.by) * (int64_t)g.Wp + b0 + lane + 4 * m) * (64 * a.och),
                         work_s + lane * (kStageUF * 4), 256u * (uint32_t)a.och);
            bulk_commit();                                // every lane commits (possibly empty) groups: counts stay in step
            if (a.zr_masks) {                             // scan block j = 3 u + ch of this round (ch < och are stored)
                unsigned long long mk;
                zr_masks_from_staging<12>(work_b, [](int j) { return (j / 3) * (kStageUF * 4) + (j % 3) * 256; }, lane, mk);
                const int u_ = lane / 3, ch_ = lane - 3 * u_;
                if (lane < 12 && u_ + 4 * m < nb && ch_ < a.och) {
                    const int64_t sblk = ((cur.frame * g.Hp + cur.by) * (int64_t)g.Wp + b0 + u_ + 4 * m) * a.och + ch_;
                    a.zr_masks[sblk] = mk;
                    a.zr_counts[sblk] = zr_block_count(mk);
                }
            }
        }
        cur = nxt;
        nxt = nx2;
    }
    bulk_wait_all0();
}

// K2p v3: the prediction is needed only at the very end of a tile, so its gather is issued at the top
// of the tile's own iteration on a second mbarrier (a single PRED buffer); the scan blocks of the next
// tile are prefetched as before.
template <int WARPS, int CTAS>
__global__ void __launch_bounds__(WARPS * 32, CTAS) k_pframe_inverse_tm(const InvArgs a, const __grid_constant__ CUtensorMap tm_ref) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *s_tT = reinterpret_cast<double *>(smem_raw);                        // [64] luminance table, transposed
    unsigned long long *s_bar = reinterpret_cast<unsigned long long *>(smem_raw + 3072);   // [2 * WARPS]
    const int lane = threadIdx.x & 31, warp = uniform_warp_idx();
    unsigned char *pred_b = smem_raw + kP3Header + warp * kP3InvBuf;
    unsigned char *in_b = pred_b + kP3Pred;
    unsigned char *work_b = in_b + kP3ZzIn;
    const uint32_t bar_in = smem_u32(s_bar + 2 * warp), bar_p = bar_in + 8;
    const uint32_t in_s = smem_u32(in_b), pred_s = smem_u32(pred_b), work_s = smem_u32(work_b);

    for (int i = threadIdx.x; i < 64; i += blockDim.x) s_tT[(i & 7) * 8 + (i >> 3)] = load_table_elem(a.table, a.table_dtype, i);
    if (lane == 0) { mbar_init(bar_in, 1); mbar_init(bar_p, 1); }
    fence_mbar_init();
    __syncthreads();

    const int r = lane & 7, u = lane >> 3;
    const unsigned char *q_rd[8];
    unsigned char *t_wr[4], *o_wr;
    const unsigned char *t_rd[4];
    const unsigned char *p_rd = pred_b + u * kP3Box + r * 8;                    // + m*4*kP3Box + parity*8 + i*80
#pragma unroll
    for (int j = 0; j < 8; ++j) q_rd[j] = in_b + u * kP3ZzBlk + ZZ_ORDER[r * 8 + j] * 4;     // + m * 4 * kP3ZzBlk
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        t_wr[h] = work_b + u * (kP3TU * 8) + ((((r >> 1) ^ h) << 1) + (r & 1)) * 8;
        t_rd[h] = work_b + u * (kP3TU * 8) + r * 64 + ((h ^ (r >> 1)) << 4);
    }
    o_wr = work_b + (8 * u + r) * 8;                                            // + i*528 + m*256
    const double *tq_l = s_tT + r;                                              // tq_l[j*8] = lum[r][j]

    const TileGeom &g = a.g;
    const int64_t row_elems = g.W, frame_elems = g.H * g.W;
    const int Hi = (int)g.H, Wi = (int)g.W;
    const int64_t gw = (int64_t)blockIdx.x * WARPS + warp;
    const int64_t nw = (int64_t)gridDim.x * WARPS;
    if (gw >= g.total_tiles) return;
    const int64_t my_tiles = (g.total_tiles - gw + nw - 1) / nw;
    TileIter cur, nxt, nx2;
    cur.init(g, gw, nw);
    nxt = cur;

    auto load_mv = [&](const TileIter &ti) -> int64_t {
        const int b0 = ti.tx * kP3Blocks, b = lane - 8;
        return (b >= 0 && b < min(kP3Blocks, g.Wp - b0)) ? a.mv[(ti.frame * g.Hp + ti.by) * (int64_t)g.Wp + b0 + b] : 0;
    };
    auto issue_in = [&](const TileIter &ti) {               // lane b: scan channel 0 of block b (stride Czz*64 ints)
        const int b0 = ti.tx * kP3Blocks, nb = min(kP3Blocks, g.Wp - b0);
        fence_proxy_async();
        if (lane == 0) mbar_expect_tx(bar_in, (uint32_t)nb * 256u);
        __syncwarp();
        if (lane < nb)
            bulk_g2s(in_s + lane * kP3ZzBlk, a.zz + ((ti.frame * g.Hp + ti.by) * (int64_t)g.Wp + b0 + lane) * a.Czz * 64, 256u, bar_in);
    };
    auto issue_pred = [&](const TileIter &ti, int64_t mvidx) -> int {
        const int b0 = ti.tx * kP3Blocks, nb = min(kP3Blocks, g.Wp - b0);
        int bx, byy, par;
        p3_box(lane, nb, mvidx, a.sr, Hi, Wi, ti.by, b0, bx, byy, par);
        fence_proxy_async();
        if (elect_one()) mbar_expect_tx(bar_p, (uint32_t)kP3Pred);
        p3_gather(&tm_ref, pred_s, bar_p, bx, byy, (int)ti.frame);
        return par;
    };

    int64_t mv_cur = load_mv(cur);
    issue_in(cur);
    nxt.advance(g);
    int64_t mv_nxt = (my_tiles > 1) ? load_mv(nxt) : 0;
    nx2 = nxt;
    uint32_t parity = 0;
    for (int64_t it = 0; it < my_tiles; ++it, parity ^= 1u) {
        nx2.advance(g);
        const int b0 = cur.tx * kP3Blocks, nb = min(kP3Blocks, g.Wp - b0);
        const int par_l = issue_pred(cur, mv_cur);      // PRED was consumed at the end of the previous iteration
        mbar_wait(bar_in, parity);
        int q[2][8];
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int j = 0; j < 8; ++j) q[m][j] = *reinterpret_cast<const int *>(q_rd[j] + m * (4 * kP3ZzBlk));
        __syncwarp();
        mv_cur = mv_nxt;
        if (it + 1 < my_tiles) {
            issue_in(nxt);
            mv_nxt = (it + 2 < my_tiles) ? load_mv(nx2) : 0;
        }
        double x[2][8];
        int mx = 0;
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const double p = __dmul_rn(i32_to_f64(q[m][j]), tq_l[j * 8]);       // luminance table for every block
                mx = max(mx, __double2hiint(p) & 0x7fffffff);
                x[m][j] = trunc_f64_small(p);
            }
        if (__builtin_expect(mx >= 0x41E00000, 0)) {
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int j = 0; j < 8; ++j) x[m][j] = dequantize_f64(q[m][j], tq_l[j * 8]);
        }
#pragma unroll
        for (int m = 0; m < 2; ++m) dct3_8(x[m]);
        bulk_wait_read0();                              // every lane drains its own bulk-store group
        __syncwarp();
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int j = 0; j < 8; ++j) *reinterpret_cast<double *>(t_wr[j >> 1] + (m * 8 + j) * 64) = x[m][j];
        __syncwarp();
#pragma unroll
        for (int m = 0; m < 2; ++m) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double2 v = *reinterpret_cast<const double2 *>(t_rd[k] + m * 512);
                x[m][2 * k] = v.x;
                x[m][2 * k + 1] = v.y;
            }
            dct3_8(x[m]);
        }
        __syncwarp();
        mbar_wait(bar_p, parity);
        {
            double pr[2][8];                             // prediction of column r of the two blocks (loaded as a batch)
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                const unsigned char *pp = p_rd + m * (4 * kP3Box) + 8 * __shfl_sync(0xffffffffu, par_l, 8 + u + 4 * m);
#pragma unroll
                for (int i = 0; i < 8; ++i) pr[m][i] = *reinterpret_cast<const double *>(pp + i * 80);
            }
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int i = 0; i < 8; ++i)              // recon = prediction + recon_residual (videocodec.py:74)
                    *reinterpret_cast<double *>(o_wr + i * kP3Pitch + m * 256) = __dadd_rn(pr[m][i], x[m][i]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane < 8) {                                   // lane r stores output row r
            double *dst = a.out + cur.frame * frame_elems + ((int64_t)cur.by * 8 + lane) * row_elems + (int64_t)b0 * 8;
            bulk_s2g(dst, work_s + lane * kP3Pitch, (uint32_t)nb * 64u);
            bulk_commit();
        }
        cur = nxt;
        nxt = nx2;
    }
    bulk_wait_all0();
}

// ================================================================================================
// unfused per-method kernels (each class method alone; simple, still coalesced where it matters)
// ================================================================================================
template <typename TI>
__device__ __forceinline__ double load_as_f64(const void *p, int64_t i) { return (double)((const TI *)p)[i]; }

struct DctArgs {
    const void *x;
    int x_dtype;
    int64_t nblocks;                 // n0*n1*C
    int64_t n1, C;
    int64_t s[5];
    void *out;
};

// one warp = 4 blocks; lane (r,u) row pass, lane (j,u) column pass, T buffer [u][j][r] pitch 9
template <typename T, bool INVERSE, int NORM = kNormOrtho>
__global__ void __launch_bounds__(256) k_dct8x8(const DctArgs a) {
    __shared__ T s_T[8][4 * 8 * 9];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r = lane & 7, u = lane >> 3;
    T *tb = s_T[warp] + u * 72;
    const int64_t gw = (int64_t)blockIdx.x * 8 + warp, nw = (int64_t)gridDim.x * 8;
    const int64_t ngroups = (a.nblocks + 3) / 4;
    for (int64_t grp = gw; grp < ngroups; grp += nw) {
        const int64_t blk = grp * 4 + u;
        const bool valid = blk < a.nblocks;
        T x[8];
        if (valid) {
            const int64_t c = blk % a.C, q = blk / a.C, i1 = q % a.n1, i0 = q / a.n1;
            const int64_t base = i0 * a.s[0] + i1 * a.s[1] + c * a.s[2] + r * a.s[3];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int64_t off = base + j * a.s[4];
                switch (a.x_dtype) {
                    case IVC_U8: x[j] = (T)((const unsigned char *)a.x)[off]; break;
                    case IVC_I32: x[j] = (T)((const int *)a.x)[off]; break;
                    case IVC_F32: x[j] = (T)((const float *)a.x)[off]; break;
                    default: x[j] = (T)((const double *)a.x)[off]; break;
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = (T)0;
        }
        if (INVERSE) dct3_8<T, NORM>(x); else dct2_8<T, NORM>(x);
#pragma unroll
        for (int j = 0; j < 8; ++j) tb[j * 9 + r] = x[j];
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = tb[r * 9 + i];
        __syncwarp();
        if (INVERSE) dct3_8<T, NORM>(x); else dct2_8<T, NORM>(x);
        if (valid) {
            T *o = (T *)a.out + blk * 64;
#pragma unroll
            for (int v = 0; v < 8; ++v) o[v * 8 + r] = x[v];
        }
    }
}

struct QuantArgs {
    const void *x;
    int x_dtype;
    int64_t n0, n1, C;               // input channels (1 or 3)
    int64_t s[5];
    const void *table;
    int table_dtype;
    int32_t *out;                    // [n0,n1,3,8,8]
};

template <typename TI>
__device__ __forceinline__ TI ld_elem(const void *p, int64_t i) { return ((const TI *)p)[i]; }

// one thread per OUTPUT element; COMPUTE_F32 selects numpy's float32 division path
template <bool COMPUTE_F32, bool DEQUANT>
__global__ void __launch_bounds__(256) k_quant(const QuantArgs a) {
    const int64_t total = a.n0 * a.n1 * 3 * 64;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(i & 63);
        int64_t q = i >> 6;
        const int ch = (int)(q % 3);
        q /= 3;
        const int64_t i1 = q % a.n1, i0 = q / a.n1;
        const int64_t cin = (a.C == 1) ? 0 : ch;
        const int64_t off = i0 * a.s[0] + i1 * a.s[1] + cin * a.s[2] + (k >> 3) * a.s[3] + (k & 7) * a.s[4];
        const int ti = ch * 64 + k;
        if (COMPUTE_F32) {
            float x;
            switch (a.x_dtype) {
                case IVC_U8: x = (float)ld_elem<unsigned char>(a.x, off); break;
                case IVC_I32: x = (float)ld_elem<int>(a.x, off); break;
                default: x = ld_elem<float>(a.x, off); break;
            }
            const float t = ((const float *)a.table)[ti];
            if (!DEQUANT) {
                a.out[i] = quantize_f32(x, t);
            } else {
                const float p = __fmul_rn(x, t);
                a.out[i] = (p >= -2147483648.0f && p < 2147483648.0f) ? __float2int_rz(p) : (int)0x80000000;
            }
        } else {
            double x;
            switch (a.x_dtype) {
                case IVC_U8: x = (double)ld_elem<unsigned char>(a.x, off); break;
                case IVC_I32: x = (double)ld_elem<int>(a.x, off); break;
                case IVC_F32: x = (double)ld_elem<float>(a.x, off); break;
                case IVC_I64: x = (double)ld_elem<long long>(a.x, off); break;
                default: x = ld_elem<double>(a.x, off); break;
            }
            const double t = load_table_elem(a.table, a.table_dtype, ti);
            if (!DEQUANT) a.out[i] = quantize_exact_f64(x, t);
            else a.out[i] = cast_i32_x86(__dmul_rn(x, t));
        }
    }
}

template <typename E>
__global__ void __launch_bounds__(256) k_zigzag(const E *__restrict__ x, E *__restrict__ out, int64_t nelem, int inverse) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nelem; i += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(i & 63);
        const int64_t base = i - k;
        if (!inverse) out[base + ZZ_ORDER[k]] = x[i];           // scatter (shape.py:26)
        else out[i] = x[base + ZZ_ORDER[k]];                    // gather  (shape.py:32)
    }
}

// ================================================================================================
// host-side launchers (called from ivc_abi.cu)
// ================================================================================================
static bool use_v1() {            // A/B switch for profiling: IVC_FUSED_V1=1 selects the first-generation kernels
    static const bool v = [] { const char *e = getenv("IVC_FUSED_V1"); return e && e[0] == '1'; }();
    return v;
}

// P-frame kernel generation: IVC_PFRAME=2 keeps the cp.async gather (v2, also the fallback when the reference planes
// cannot be described by a tensor map); default: v3, tensor-map gather, 16 warps per SM.
static bool pframe_v3() {
    static const bool v = [] { const char *e = getenv("IVC_PFRAME"); return !(e && e[0] == '2'); }();
    return v;
}

// Tensor maps for the v3 P-frame kernels (no interleave, no swizzle, zero fill).  make_map returns false when the
// driver entry point is missing or the view does not meet the unit's rules (base and strides multiples of 16 bytes,
// extents below 2^32, strides below 2^40) -- the caller then uses the v2 kernels.
static bool make_map(CUtensorMap *tm, CUtensorMapDataType dt, int elem, int rank, const void *base, const uint64_t *dims,
                     const uint64_t *strides /* bytes, rank - 1 */, const uint32_t *box) {
    typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static const EncodeTiled encode = [] {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            fn = nullptr;
        return (EncodeTiled)fn;
    }();
    if (!encode || !base || ((uintptr_t)base & 15)) return false;
    cuuint64_t d[5], st[4];
    cuuint32_t b[5], es[5];
    for (int i = 0; i < rank; ++i) {
        if (dims[i] < 1 || dims[i] > 0xffffffffull || box[i] < 1 || box[i] > 256) return false;
        d[i] = dims[i]; b[i] = box[i]; es[i] = 1;
    }
    for (int i = 0; i + 1 < rank; ++i) {
        if ((strides[i] & 15) || strides[i] >= (1ull << 40)) return false;
        st[i] = strides[i];
    }
    if (((uint64_t)box[0] * elem) & 15) return false;
    return encode(tm, dt, (cuuint32_t)rank, const_cast<void *>(base), d, st, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// float64 planes [n, H, W] (frame stride in elements), boxes of bw x 8 elements
static bool make_plane_map(CUtensorMap *tm, const void *base, int64_t n, int64_t H, int64_t W, int64_t frame_stride, int bw) {
    if (n > 0x7fffffff || H > 0x7fffffff || W > 0x7fffffff) return false;       // coordinates are int32
    const uint64_t dims[3] = {(uint64_t)W, (uint64_t)H, (uint64_t)n};
    const uint64_t strides[2] = {(uint64_t)W * 8, (uint64_t)frame_stride * 8};
    const uint32_t box[3] = {(uint32_t)bw, 8, 1};
    return make_map(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, 3, base, dims, strides, box);
}

static int grid_for(int64_t work_items, int per_cta, int device, int ctas_per_sm) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int64_t need = (work_items + per_cta - 1) / per_cta;
    const int64_t cap = (int64_t)sms * ctas_per_sm;
    return (int)(need < 1 ? 1 : (need < cap ? need : cap));
}

static TileGeom make_geom(int64_t n, int64_t H, int64_t W, int C, int blocks_per_tile) {
    TileGeom g;
    g.n_frames = n; g.H = H; g.W = W; g.Hp = (int)(H / 8); g.Wp = (int)(W / 8); g.C = C;
    g.tiles_per_row = (g.Wp + blocks_per_tile - 1) / blocks_per_tile;
    g.total_tiles = n * g.Hp * (int64_t)g.tiles_per_row;
    return g;
}

template <typename K>
static cudaError_t set_smem(K kernel, size_t bytes) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

cudaError_t launch_forward(int device, cudaStream_t st, const void *img, int64_t n, int64_t H, int64_t W, int C,
                           int64_t frame_stride, const void *table, int table_dtype, int32_t *out,
                           const void *ref, const int64_t *mv, int sr, void *pred_out, bool pframe, int out_channels,
                           const int *run_flag, int32_t *zr_counts, uint64_t *zr_masks, bool *zr_done) {
    FwdArgs a;
    a.och = pframe ? out_channels : 3;
    a.run_flag = pframe ? run_flag : nullptr;
    a.zr_counts = nullptr; a.zr_masks = nullptr;                  // only the tensor-map P-frame kernel emits them
    if (zr_done) *zr_done = false;
    a.g = make_geom(n, H, W, C, C == 3 ? 4 : 12);
    a.img = (const double *)img; a.frame_stride = frame_stride; a.table = table; a.table_dtype = table_dtype;
    a.out = out; a.ref = (const double *)ref; a.mv = mv; a.sr = sr; a.pred_out = (double *)pred_out;
    if (a.g.total_tiles == 0) return cudaSuccess;
    const size_t smem = 2 * 192 * sizeof(double) + kWarpsPerCta * kWarpBufBytes;
    const int grid = grid_for(a.g.total_tiles, kWarpsPerCta, device, 2);
    cudaError_t e;
    CUtensorMap tm_ref;
    if (pframe && !use_v1() && pframe_v3() && make_plane_map(&tm_ref, ref, n, H, W, H * W, 10)) {
        a.g = make_geom(n, H, W, 1, kP3Blocks);
        if (zr_counts && zr_masks) {
            a.zr_counts = zr_counts; a.zr_masks = (unsigned long long *)zr_masks;
            if (zr_done) *zr_done = true;
        }
        const size_t smem3 = kP3Header + (size_t)8 * kP3FwdBuf;
        if ((e = set_smem(k_pframe_forward_tm<8, 2>, smem3)) != cudaSuccess) return e;
        k_pframe_forward_tm<8, 2><<<grid_for(a.g.total_tiles, 8, device, 2), 8 * 32, smem3, st>>>(a, tm_ref);
    } else if (pframe && !use_v1()) {
        const size_t smem2 = 3200 + (size_t)kPfWarps * kPfBuf;
        if ((e = set_smem(k_pframe_forward_tma, smem2)) != cudaSuccess) return e;
        k_pframe_forward_tma<<<grid_for(a.g.total_tiles, kPfWarps, device, 1), kPfWarps * 32, smem2, st>>>(a);
    } else if (pframe) {
        if ((e = set_smem(k_forward<1, true>, smem)) != cudaSuccess) return e;
        k_forward<1, true><<<grid, kWarpsPerCta * 32, smem, st>>>(a);
    } else if (C == 3 && !use_v1()) {
        const size_t smem2 = 3200 + (size_t)kWarpsPerCta * kWarpBuf2;
        if ((e = set_smem(k_forward_c3_tma, smem2)) != cudaSuccess) return e;
        k_forward_c3_tma<<<grid, kWarpsPerCta * 32, smem2, st>>>(a);
    } else if (C == 3) {
        if ((e = set_smem(k_forward<3, false>, smem)) != cudaSuccess) return e;
        k_forward<3, false><<<grid, kWarpsPerCta * 32, smem, st>>>(a);
    } else {
        if ((e = set_smem(k_forward<1, false>, smem)) != cudaSuccess) return e;
        k_forward<1, false><<<grid, kWarpsPerCta * 32, smem, st>>>(a);
    }
    return cudaGetLastError();
}

cudaError_t launch_forward_rgb8(int device, cudaStream_t st, const void *rgb, int64_t n, int64_t H, int64_t W,
                                int64_t frame_stride_bytes, const void *table, int table_dtype, int32_t *out,
                                int32_t *zr_counts, uint64_t *zr_masks, int nq) {
    if (nq < 1 || nq > kMaxForwardScales) return cudaErrorInvalidValue;
    FwdArgs a;
    a.nq = nq; a.out_q_stride = n * (H / 8) * (W / 8) * 192; a.zr_q_stride = n * (H / 8) * (W / 8) * 3;
    a.zr_counts = zr_counts; a.zr_masks = (unsigned long long *)zr_masks;
    a.g = make_geom(n, H, W, 3, 4);
    a.img = (const double *)rgb; a.frame_stride = frame_stride_bytes; a.table = table; a.table_dtype = table_dtype;
    a.out = out; a.ref = nullptr; a.mv = nullptr; a.sr = 0; a.pred_out = nullptr; a.och = 3; a.run_flag = nullptr;
    if (a.g.total_tiles == 0) return cudaSuccess;
    const size_t smem = (size_t)nq * 3072 + 128 + (size_t)kWarpsPerCta * kRgbBuf;      // <= 113 KB: two CTAs per SM
    cudaError_t e;
    const int grid = grid_for(a.g.total_tiles, kWarpsPerCta, device, 2);
    if (nq == 1) {
        if ((e = set_smem(k_forward_rgb8_tma<false>, smem)) != cudaSuccess) return e;
        k_forward_rgb8_tma<false><<<grid, kWarpsPerCta * 32, smem, st>>>(a);
    } else {
        if ((e = set_smem(k_forward_rgb8_tma<true>, (size_t)kMaxForwardScales * 3072 + 128 + (size_t)kWarpsPerCta * kRgbBuf)) != cudaSuccess) return e;   // a cap, not the launch size
        k_forward_rgb8_tma<true><<<grid, kWarpsPerCta * 32, smem, st>>>(a);
    }
    return cudaGetLastError();
}

cudaError_t launch_inverse(int device, cudaStream_t st, const int32_t *zz, int64_t n, int64_t Hp, int64_t Wp, int Czz,
                           const void *table, int table_dtype, void *out, int mode,
                           const void *pred, const void *ref, const int64_t *mv, int sr) {
    InvArgs a;
    a.g = make_geom(n, Hp * 8, Wp * 8, Czz, mode == 2 ? 12 : 4);      // mode 3: like 0, RGB store
    a.zz = zz; a.Czz = Czz; a.table = table; a.table_dtype = table_dtype; a.out = (double *)out;
    a.pred = (const double *)pred; a.ref = (const double *)ref; a.mv = mv; a.sr = sr;
    a.orig = nullptr; a.orig_frame_stride = 0; a.sse_partial = nullptr;
    if (a.g.total_tiles == 0) return cudaSuccess;
    const size_t smem = 192 * sizeof(double) + kWarpsPerCta * kWarpBufBytes;
    const int grid = grid_for(a.g.total_tiles, kWarpsPerCta, device, 2);
    cudaError_t e;
    CUtensorMap tm_ref;
    if (mode == 3) {                                         // C = 3 with the colour transform in the store
        const size_t smem2 = 1664 + (size_t)kWarpsPerCta * kInvBuf;
        if ((e = set_smem(k_inverse_c3_tma<0, true>, smem2)) != cudaSuccess) return e;
        k_inverse_c3_tma<0, true><<<grid, kWarpsPerCta * 32, smem2, st>>>(a);
    } else if (mode == 0 && !use_v1()) {
        const size_t smem2 = 1664 + (size_t)kWarpsPerCta * kInvBuf;
        if ((e = set_smem(k_inverse_c3_tma<0>, smem2)) != cudaSuccess) return e;
        k_inverse_c3_tma<0><<<grid, kWarpsPerCta * 32, smem2, st>>>(a);
    } else if (mode == 0) {
        if ((e = set_smem(k_inverse<0>, smem)) != cudaSuccess) return e;
        k_inverse<0><<<grid, kWarpsPerCta * 32, smem, st>>>(a);
    } else if (mode == 1) {
        if ((e = set_smem(k_inverse<1>, smem)) != cudaSuccess) return e;
        k_inverse<1><<<grid, kWarpsPerCta * 32, smem, st>>>(a);
    } else if (!use_v1() && !pred && pframe_v3() && make_plane_map(&tm_ref, ref, n, Hp * 8, Wp * 8, Hp * Wp * 64, 10)) {
        a.g = make_geom(n, Hp * 8, Wp * 8, Czz, kP3Blocks);
        const size_t smem3 = kP3Header + (size_t)8 * kP3InvBuf;
        if ((e = set_smem(k_pframe_inverse_tm<8, 2>, smem3)) != cudaSuccess) return e;
        k_pframe_inverse_tm<8, 2><<<grid_for(a.g.total_tiles, 8, device, 2), 8 * 32, smem3, st>>>(a, tm_ref);
    } else if (!use_v1()) {
        const size_t smem2 = 3200 + (size_t)kPiWarps * kPiBuf;
        if ((e = set_smem(k_pframe_inverse_tma, smem2)) != cudaSuccess) return e;
        k_pframe_inverse_tma<<<grid_for(a.g.total_tiles, kPiWarps, device, 1), kPiWarps * 32, smem2, st>>>(a);
    } else {
        if ((e = set_smem(k_inverse<2>, smem)) != cudaSuccess) return e;
        k_inverse<2><<<grid, kWarpsPerCta * 32, smem, st>>>(a);
    }
    return cudaGetLastError();
}

int64_t inverse_sse_tiles(int64_t n, int64_t Hp, int64_t Wp) { return n * Hp * ((Wp + 3) / 4); }

// decode (3 scan channels) and measure the distortion against uint8 RGB originals in the same kernel
cudaError_t launch_inverse_sse(int device, cudaStream_t st, const int32_t *zz, int64_t n, int64_t Hp, int64_t Wp,
                               const void *table, int table_dtype, void *out, const void *orig_rgb8,
                               int64_t orig_frame_stride, int sse_mode, double *partial, double *sse_out) {
    InvArgs a;
    a.g = make_geom(n, Hp * 8, Wp * 8, 3, 4);
    a.zz = zz; a.Czz = 3; a.table = table; a.table_dtype = table_dtype; a.out = (double *)out;
    a.pred = nullptr; a.ref = nullptr; a.mv = nullptr; a.sr = 0;
    a.orig = (const unsigned char *)orig_rgb8; a.orig_frame_stride = orig_frame_stride; a.sse_partial = partial;
    if (a.g.total_tiles == 0) return cudaSuccess;
    const int grid = grid_for(a.g.total_tiles, kWarpsPerCta, device, 2);
    const size_t smem = 1664 + (size_t)kWarpsPerCta * kInvBufSse;
    cudaError_t e;
    if (sse_mode == 1) {
        if ((e = set_smem(k_inverse_c3_tma<1>, smem)) != cudaSuccess) return e;
        k_inverse_c3_tma<1><<<grid, kWarpsPerCta * 32, smem, st>>>(a);
    } else {
        if ((e = set_smem(k_inverse_c3_tma<2>, smem)) != cudaSuccess) return e;
        k_inverse_c3_tma<2><<<grid, kWarpsPerCta * 32, smem, st>>>(a);
    }
    k_sum_tile_partials<<<(unsigned)n, 256, 0, st>>>(partial, Hp * (int64_t)a.g.tiles_per_row, sse_out);
    return cudaGetLastError();
}

template <int NORM>
static void launch_dct_norm(int grid, cudaStream_t st, bool inverse, bool f32, const DctArgs &a) {
    if (f32) {
        if (inverse) k_dct8x8<float, true, NORM><<<grid, 256, 0, st>>>(a);
        else k_dct8x8<float, false, NORM><<<grid, 256, 0, st>>>(a);
    } else {
        if (inverse) k_dct8x8<double, true, NORM><<<grid, 256, 0, st>>>(a);
        else k_dct8x8<double, false, NORM><<<grid, 256, 0, st>>>(a);
    }
}

cudaError_t launch_dct(int device, cudaStream_t st, bool inverse, const void *x, int x_dtype, int64_t n0, int64_t n1,
                       int64_t C, const int64_t s[5], void *out, bool f32, int norm) {
    DctArgs a;
    a.x = x; a.x_dtype = x_dtype; a.nblocks = n0 * n1 * C; a.n1 = n1; a.C = C; a.out = out;
    for (int i = 0; i < 5; ++i) a.s[i] = s[i];
    if (a.nblocks == 0) return cudaSuccess;
    const int grid = grid_for((a.nblocks + 3) / 4, 8, device, 8);
    switch (norm) {
        case kNormOrtho: launch_dct_norm<kNormOrtho>(grid, st, inverse, f32, a); break;
        case kNormBackward: launch_dct_norm<kNormBackward>(grid, st, inverse, f32, a); break;
        case kNormForward: launch_dct_norm<kNormForward>(grid, st, inverse, f32, a); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_quant(int device, cudaStream_t st, bool dequant, const void *x, int x_dtype, int64_t n0, int64_t n1,
                         int64_t C, const int64_t s[5], const void *table, int table_dtype, bool f32, int32_t *out) {
    QuantArgs a;
    a.x = x; a.x_dtype = x_dtype; a.n0 = n0; a.n1 = n1; a.C = C; a.table = table; a.table_dtype = table_dtype; a.out = out;
    for (int i = 0; i < 5; ++i) a.s[i] = s[i];
    const int64_t total = n0 * n1 * 3 * 64;
    if (total == 0) return cudaSuccess;
    const int grid = grid_for(total, 256 * 4, device, 16);
    if (f32) {
        if (dequant) k_quant<true, true><<<grid, 256, 0, st>>>(a);
        else k_quant<true, false><<<grid, 256, 0, st>>>(a);
    } else {
        if (dequant) k_quant<false, true><<<grid, 256, 0, st>>>(a);
        else k_quant<false, false><<<grid, 256, 0, st>>>(a);
    }
    return cudaGetLastError();
}

cudaError_t launch_zigzag(int device, cudaStream_t st, bool inverse, const void *x, int elem_size, int64_t nblocks, void *out) {
    const int64_t nelem = nblocks * 64;
    if (nelem == 0) return cudaSuccess;
    const int grid = grid_for(nelem, 256 * 4, device, 16);
    switch (elem_size) {
        case 1: k_zigzag<unsigned char><<<grid, 256, 0, st>>>((const unsigned char *)x, (unsigned char *)out, nelem, inverse); break;
        case 2: k_zigzag<unsigned short><<<grid, 256, 0, st>>>((const unsigned short *)x, (unsigned short *)out, nelem, inverse); break;
        case 4: k_zigzag<unsigned int><<<grid, 256, 0, st>>>((const unsigned int *)x, (unsigned int *)out, nelem, inverse); break;
        case 8: k_zigzag<unsigned long long><<<grid, 256, 0, st>>>((const unsigned long long *)x, (unsigned long long *)out, nelem, inverse); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace ivc
