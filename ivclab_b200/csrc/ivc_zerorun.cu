// ivc_zerorun.cu -- zero-run coder over scan blocks, both directions (ivclab/entropy/zerorun.py:10-87;
// the first host-side consumer of K1's output and the real bottleneck of
// IntraCodec.image2symbols / symbols2image).
//
// ENCODE (zerorun.py:10-43).  Per 64-coefficient block the reference emits: every non-zero value as
// itself; every run of zeros that is followed by a non-zero value as the pair (0, run_length); then EOB.
// With m = the 64-bit "non-zero" mask of a block, L = its highest set bit and S = the zero positions below
// L that start a run (S = ~m & ((m << 1) | 1)), the block contributes popc(m) + 2*popc(S) + 1 symbols, and
// position p writes at offset popc(m & below(p)) + 2*popc(S & below(p)); a run starting at p has length
// ctz(m >> p).  Two passes: count (-> exclusive scan on the caller's side) and write.  One WARP owns one
// block at a time (lane l holds coefficients l and l+32: two coalesced 128-byte loads, the mask is two
// ballots, every lane places its own symbols), eight blocks' loads in flight per warp.
//
// DECODE (zerorun.py:44-87).  A symbol slot holds a value, the zero marker or EOB; the slot after a zero
// marker holds a run length.  With run lengths >= 1 (all the encoder emits) a slot is a run length iff its
// predecessor reads 0, so EOBs can be marked independently (mark), numbered by the caller's inclusive scan,
// turned into block boundaries (ends), and every block is then expanded by one warp (write): per-slot
// output counts, a warp scan, a scatter into a zeroed 64-entry row.  Streams the reference would reject
// (or a zero run length) raise a flag instead of producing blocks.
#include "ivc_common.cuh"
#include "ivc_tile.cuh"

namespace ivc {

constexpr int kZrWarps = 8;
constexpr int kZrBatch = 8;        // blocks whose loads are in flight per warp


// counts[b] = symbols of block b; masks[b] (optional) = its 64-bit non-zero mask, which lets the write pass skip
// the parts of a block that hold no symbol (most of it, for typical quantised blocks)
__global__ void __launch_bounds__(kZrWarps * 32) k_zr_count(const int32_t *__restrict__ zz, int64_t nblocks,
                                                            int32_t *__restrict__ counts, unsigned long long *__restrict__ masks,
                                                            const int *run_flag) {
    if (run_flag && *run_flag == 0) return;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * kZrWarps + (threadIdx.x >> 5), nw = (int64_t)gridDim.x * kZrWarps;
    for (int64_t blk0 = warp * 32; blk0 < nblocks; blk0 += nw * 32) {                 // 32 blocks per warp and round
        unsigned long long mine_m = 0;                                                // lane j keeps the mask of block j
#pragma unroll 1
        for (int sub = 0; sub < 32; sub += kZrBatch) {
            int a[kZrBatch], b[kZrBatch];
#pragma unroll
            for (int j = 0; j < kZrBatch; ++j) {
                const int64_t blk = blk0 + sub + j;
                a[j] = b[j] = 0;
                if (blk < nblocks) {
                    a[j] = __ldg(zz + blk * 64 + lane);
                    b[j] = __ldg(zz + blk * 64 + 32 + lane);
                }
            }
#pragma unroll
            for (int j = 0; j < kZrBatch; ++j) {
                const unsigned lo = __ballot_sync(0xffffffffu, a[j] != 0), hi = __ballot_sync(0xffffffffu, b[j] != 0);
                if (lane == sub + j) mine_m = (unsigned long long)lo | ((unsigned long long)hi << 32);
            }
        }
        if (blk0 + lane < nblocks) {                                                  // the symbol arithmetic once per block
            int c = 1;                                                                // EOB
            if (mine_m) c += __popcll(mine_m) + 2 * __popcll(zr_run_starts(mine_m));
            counts[blk0 + lane] = c;
            if (masks) masks[blk0 + lane] = mine_m;
        }
    }
}

// masks == nullptr: the masks are recomputed from the blocks (every block is read whole).  With masks, a lane
// loads its coefficient only if it is non-zero, so only the 32-byte sectors that hold symbols are fetched.
// OutT = int32_t (the reference's dtype) or int16_t (a lossless transfer format when every symbol is known to fit).
template <bool MASKS, typename OutT>
__global__ void __launch_bounds__(kZrWarps * 32) k_zr_write(const int32_t *__restrict__ zz, int64_t nblocks, int32_t eob,
                                                            const int64_t *__restrict__ offsets,
                                                            const unsigned long long *__restrict__ masks, OutT *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * kZrWarps + (threadIdx.x >> 5), nw = (int64_t)gridDim.x * kZrWarps;
    const unsigned long long low0 = (1ull << lane) - 1ull, low1 = (1ull << (lane + 32)) - 1ull;
    for (int64_t blk0 = warp * 32; blk0 < nblocks; blk0 += nw * 32) {
        const int64_t my_off = blk0 + lane < nblocks ? offsets[blk0 + lane] : 0;
        const unsigned long long my_m = (MASKS && blk0 + lane < nblocks) ? masks[blk0 + lane] : 0ull;
#pragma unroll 1
        for (int sub = 0; sub < 32; sub += kZrBatch) {
            int a[kZrBatch], b[kZrBatch];
            unsigned long long mm[kZrBatch];
#pragma unroll
            for (int j = 0; j < kZrBatch; ++j) {
                const int64_t blk = blk0 + sub + j;
                a[j] = b[j] = 0;
                if (MASKS) {
                    mm[j] = __shfl_sync(0xffffffffu, my_m, sub + j);               // 0 past the last block
                    if ((mm[j] >> lane) & 1ull) a[j] = __ldg(zz + blk * 64 + lane);
                    if ((mm[j] >> (lane + 32)) & 1ull) b[j] = __ldg(zz + blk * 64 + 32 + lane);
                } else if (blk < nblocks) {
                    a[j] = __ldg(zz + blk * 64 + lane);
                    b[j] = __ldg(zz + blk * 64 + 32 + lane);
                }
            }
#pragma unroll
            for (int j = 0; j < kZrBatch; ++j) {
                const int64_t off = __shfl_sync(0xffffffffu, my_off, sub + j);
                if (blk0 + sub + j >= nblocks) break;                                 // warp-uniform
                const unsigned long long m = MASKS ? mm[j]
                                                   : (unsigned long long)__ballot_sync(0xffffffffu, a[j] != 0) |
                                                         ((unsigned long long)__ballot_sync(0xffffffffu, b[j] != 0) << 32);
                OutT *o = out + off;
                if (m == 0) {
                    if (lane == 0) o[0] = (OutT)eob;
                    continue;
                }
                const unsigned long long S = zr_run_starts(m);
                const int p0 = __popcll(m & low0) + 2 * __popcll(S & low0);
                const int p1 = __popcll(m & low1) + 2 * __popcll(S & low1);
                if ((m >> lane) & 1ull) o[p0] = (OutT)a[j];
                else if ((S >> lane) & 1ull) { o[p0] = 0; o[p0 + 1] = (OutT)(__ffsll((long long)(m >> lane)) - 1); }
                if ((m >> (lane + 32)) & 1ull) o[p1] = (OutT)b[j];
                else if ((S >> (lane + 32)) & 1ull) { o[p1] = 0; o[p1 + 1] = (OutT)(__ffsll((long long)(m >> (lane + 32))) - 1); }
                if (lane == 0) o[__popcll(m) + 2 * __popcll(S)] = (OutT)eob;
            }
        }
    }
}

// Sparse variant of the write pass: ONE THREAD per block walks the set bits of the block's mask.  Typical quantised
// blocks hold a handful of non-zero coefficients (chroma blocks often none), so a thread issues a few scattered
// 4-byte loads and stores where the cooperative kernel spends ~60 warp instructions per block regardless; the
// launcher picks this kernel when the stream averages fewer than kZrSparseLimit symbols per block.
constexpr int kZrSparseLimit = 16;

template <typename OutT>
__global__ void __launch_bounds__(256) k_zr_write_sparse(const int32_t *__restrict__ zz, int64_t nblocks, int32_t eob,
                                                         const int64_t *__restrict__ offsets,
                                                         const unsigned long long *__restrict__ masks, OutT *__restrict__ out) {
    for (int64_t blk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; blk < nblocks; blk += (int64_t)gridDim.x * blockDim.x) {
        unsigned long long m = masks[blk];
        OutT *o = out + offsets[blk];
        const int32_t *z = zz + blk * 64;
        int p = 0;
        while (m) {
            const int pos = __ffsll((long long)m) - 1;
            const int32_t v = __ldg(z + pos);
            if (pos > p) { o[0] = 0; o[1] = (OutT)(pos - p); o += 2; }            // the zeros since the last value
            *o++ = (OutT)v;
            p = pos + 1;
            m &= m - 1;
        }
        *o = (OutT)eob;
    }
}

// ---- symbol statistics of the stream, without the stream -------------------------------------------
// What IntraCodec.train_huffman_from_image needs from image2symbols is the marginal histogram of the symbols
// (intracodec.py:160-166), not their order: per block the non-zero coefficients, one zero marker and one run length
// per zero run below the last non-zero coefficient, and one EOB.  k_zr_hist counts exactly that from the scan
// blocks, per UNIT (frame): a rate-distortion sweep gets its rate statistics with one read of the indices and
// neither a count / scan / write pass nor a symbol buffer.  counts[u][k] = symbols of unit u equal to lo + k
// (uint32), outside[u] = symbols outside [lo, lo + nbins).  Grid: x = CTAs sharing a unit, y = unit; per-CTA
// histogram in shared memory; zero markers and EOBs (a third of a typical stream) are counted in registers.
__global__ void __launch_bounds__(kZrWarps * 32) k_zr_hist(const int32_t *__restrict__ zz, int64_t blocks_per_unit, int32_t eob,
                                                           int lo, int nbins, uint32_t *__restrict__ counts,
                                                           uint32_t *__restrict__ outside) {
    extern __shared__ unsigned s_hist[];
    const int lane = threadIdx.x & 31;
    for (int k = threadIdx.x; k < nbins; k += kZrWarps * 32) s_hist[k] = 0u;
    __syncthreads();
    const int64_t unit = blockIdx.y;
    const int32_t *zu = zz + unit * blocks_per_unit * 64;
    unsigned n_zero = 0, n_eob = 0, n_out = 0;
    const auto add = [&](int v) {
        const unsigned k = (unsigned)(v - lo);
        if (k < (unsigned)nbins) atomicAdd(&s_hist[k], 1u);
        else ++n_out;
    };
    const int64_t warp = (int64_t)blockIdx.x * kZrWarps + (threadIdx.x >> 5), nw = (int64_t)gridDim.x * kZrWarps;
    for (int64_t blk0 = warp * 32; blk0 < blocks_per_unit; blk0 += nw * 32) {
        unsigned long long mine_m = 0;
#pragma unroll 1
        for (int sub = 0; sub < 32; sub += kZrBatch) {
            int a[kZrBatch], b[kZrBatch];
#pragma unroll
            for (int j = 0; j < kZrBatch; ++j) {
                const int64_t blk = blk0 + sub + j;
                a[j] = b[j] = 0;
                if (blk < blocks_per_unit) {
                    a[j] = __ldg(zu + blk * 64 + lane);
                    b[j] = __ldg(zu + blk * 64 + 32 + lane);
                }
            }
#pragma unroll
            for (int j = 0; j < kZrBatch; ++j) {
                const unsigned ml = __ballot_sync(0xffffffffu, a[j] != 0), mh = __ballot_sync(0xffffffffu, b[j] != 0);
                if (lane == sub + j) mine_m = (unsigned long long)ml | ((unsigned long long)mh << 32);
                if (a[j] != 0) add(a[j]);                                             // every non-zero coefficient is a symbol
                if (b[j] != 0) add(b[j]);
            }
        }
        if (blk0 + lane < blocks_per_unit) {                                          // lane j: the runs of block blk0 + j
            ++n_eob;
            if (mine_m) {
                unsigned long long S = zr_run_starts(mine_m);
                n_zero += (unsigned)__popcll(S);
                while (S) {
                    const int p = __ffsll((long long)S) - 1;
                    add(__ffsll((long long)(mine_m >> p)) - 1);                       // the run length that follows the marker
                    S &= S - 1;
                }
            }
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        n_zero += __shfl_xor_sync(0xffffffffu, n_zero, off);
        n_eob += __shfl_xor_sync(0xffffffffu, n_eob, off);
        n_out += __shfl_xor_sync(0xffffffffu, n_out, off);
    }
    if (lane == 0) {
        const unsigned kz = (unsigned)(0 - lo), ke = (unsigned)(eob - lo);
        if (kz < (unsigned)nbins) atomicAdd(&s_hist[kz], n_zero); else n_out += n_zero;
        if (ke < (unsigned)nbins) atomicAdd(&s_hist[ke], n_eob); else n_out += n_eob;
        if (n_out) atomicAdd(outside + unit, n_out);
    }
    __syncthreads();
    uint32_t *cu = counts + unit * (int64_t)nbins;
    for (int k = threadIdx.x; k < nbins; k += kZrWarps * 32) {
        const unsigned c = s_hist[k];
        if (c) atomicAdd(cu + k, c);
    }
}

// ---- decode --------------------------------------------------------------------------------------
// is_eob[i] = 1 iff symbol i is an EOB in a symbol slot (i.e. not the run length after a zero marker)
__global__ void __launch_bounds__(256) k_zrd_mark(const int32_t *__restrict__ sym, int64_t n, int32_t eob,
                                                  int32_t *__restrict__ is_eob) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        is_eob[i] = (sym[i] == eob && !(i > 0 && sym[i - 1] == 0)) ? 1 : 0;
}

// rank[i] = number of marked EOBs in [0, i] (the caller's inclusive scan): EOB number k ends block k-1
__global__ void __launch_bounds__(256) k_zrd_ends(const int32_t *__restrict__ is_eob, const int64_t *__restrict__ rank,
                                                  int64_t n, int64_t want, int64_t *__restrict__ ends) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (is_eob[i] && rank[i] <= want) ends[rank[i] - 1] = i;
}

// err bits: 1 = a block expands to more than 64 coefficients (zerorun.py:76-77), 2 = zero run length
__global__ void __launch_bounds__(kZrWarps * 32) k_zrd_write(const int32_t *__restrict__ sym, const int64_t *__restrict__ ends,
                                                             int64_t nblocks, int32_t *__restrict__ out, int *err) {
    __shared__ int32_t rows[kZrWarps][64];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t warp = (int64_t)blockIdx.x * kZrWarps + w, nw = (int64_t)gridDim.x * kZrWarps;
    int32_t *row = rows[w];
    for (int64_t blk = warp; blk < nblocks; blk += nw) {
        const int64_t start = blk ? ends[blk - 1] + 1 : 0, end = ends[blk];           // [start, end) = the block's symbols
        const int64_t len = end - start;
        row[lane] = 0;
        row[lane + 32] = 0;
        __syncwarp();
        int bad = 0;
        int base = 0;                                                                 // coefficients emitted by earlier slots
        for (int64_t c0 = 0; c0 < len && base <= 64; c0 += 128) {                     // a legal block has <= 127 symbols
            int s[5], e[4];
#pragma unroll
            for (int k = 0; k < 5; ++k) {                                             // slots 4*lane-1 .. 4*lane+3 of this chunk
                const int64_t i = start + c0 + 4 * lane + k - 1;
                s[k] = (i >= start && i < end) ? sym[i] : 1;                           // "1" = harmless predecessor / filler
            }
            int tot = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const bool in = c0 + 4 * lane + k < len, is_run = in && s[k] == 0 && (c0 + 4 * lane + k > 0);
                const int v = s[k + 1];
                e[k] = !in ? 0 : is_run ? v : (v != 0 ? 1 : 0);
                if (is_run && v <= 0) bad |= 2;
                if (is_run && v > 64) e[k] = 65;                                      // clamp: only "more than 64" matters
                tot += e[k];
            }
            int incl = tot;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            int pos = base + incl - tot;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const bool in = c0 + 4 * lane + k < len, is_run = in && s[k] == 0 && (c0 + 4 * lane + k > 0);
                if (in && !is_run && s[k + 1] != 0) {
                    if (pos < 64) row[pos] = s[k + 1];
                    else bad |= 1;
                }
                pos += e[k];
            }
            base += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (base > 64 || len > 128) bad |= 1;
        __syncwarp();
        out[blk * 64 + lane] = row[lane];
        out[blk * 64 + 32 + lane] = row[lane + 32];
        if (bad) atomicOr(err, bad);
        __syncwarp();
    }
}

// A few 64-bit words from device memory into MAPPED pinned host memory, written by a kernel: a pipeline's
// stream lengths reach the host without queueing behind bulk transfers on the copy engines.
__global__ void k_post_words(const int64_t *__restrict__ src, volatile int64_t *dst, int n) {
    if ((int)threadIdx.x < n) dst[threadIdx.x] = src[threadIdx.x];
    __threadfence_system();
}

cudaError_t launch_post_words(cudaStream_t st, const int64_t *src, int64_t *dst_mapped, int n) {
    if (n == 0) return cudaSuccess;
    k_post_words<<<1, 32, 0, st>>>(src, dst_mapped, n);
    return cudaGetLastError();
}

// Exclusive prefix sum of the per-block symbol counts (int32 -> int64 offsets) in ONE kernel: tiles of 4096
// counts, decoupled look-back between tiles (state word = flag << 62 | value; 1 = tile aggregate, 2 = inclusive
// prefix), and the grand total posted to mapped pinned host memory by the last tile.  Replaces a library scan,
// a subtraction, a copy and a separate post kernel on the compute chain of the host-fed pipeline.
constexpr int kScanThreads = 1024, kScanTile = 4 * kScanThreads;

__global__ void __launch_bounds__(kScanThreads) k_zr_offsets(const int32_t *__restrict__ counts, int64_t n,
                                                             int64_t *__restrict__ offsets, unsigned long long *state,
                                                             volatile int64_t *total_mapped, int64_t *total_dev) {
    __shared__ long long s_warp[32];
    __shared__ long long s_prefix;
    const int tile = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t base = (int64_t)tile * kScanTile + 4 * threadIdx.x;
    int c[4] = {0, 0, 0, 0};
    if (base + 3 < n) {
        const int4 v = *reinterpret_cast<const int4 *>(counts + base);
        c[0] = v.x; c[1] = v.y; c[2] = v.z; c[3] = v.w;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (base + k < n) c[k] = counts[base + k];
    }
    const long long mine = (long long)c[0] + c[1] + c[2] + c[3];
    long long incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const long long t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        long long w = s_warp[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const long long t = __shfl_up_sync(0xffffffffu, w, d);
            if (lane >= d) w += t;
        }
        s_warp[lane] = w;                                                     // inclusive over warps
    }
    __syncthreads();
    const long long tile_sum = s_warp[31];
    if (warp == 0) {                                                          // warp 0 looks back 32 tiles at a time
        long long prefix = 0;
        if (tile > 0) {
            if (lane == 0) atomicExch(&state[tile], (1ull << 62) | (unsigned long long)tile_sum);
            for (int p0 = tile - 1; p0 >= 0; p0 -= 32) {
                const int p = p0 - lane;
                unsigned long long st = 2ull << 62;                           // lanes before tile 0: "inclusive, value 0"
                if (p >= 0)
                    do { st = *reinterpret_cast<volatile unsigned long long *>(&state[p]); } while ((st >> 62) == 0);
                const unsigned incl_mask = __ballot_sync(0xffffffffu, (st >> 62) == 2);
                const int first = incl_mask ? __ffs(incl_mask) - 1 : 32;      // nearest predecessor with an inclusive prefix
                long long v = lane <= first ? (long long)(st & ((1ull << 62) - 1)) : 0;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
                prefix += v;
                if (incl_mask) break;
            }
        }
        if (lane == 0) {
            atomicExch(&state[tile], (2ull << 62) | (unsigned long long)(prefix + tile_sum));
            s_prefix = prefix;
            if (tile == (int)gridDim.x - 1) {
                if (total_dev) *total_dev = prefix + tile_sum;
                if (total_mapped) { *total_mapped = prefix + tile_sum; __threadfence_system(); }
            }
        }
    }
    __syncthreads();
    long long off = s_prefix + (warp ? s_warp[warp - 1] : 0) + (incl - mine);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (base + k < n) offsets[base + k] = off;
        off += c[k];
    }
}

int64_t zr_offsets_workspace_bytes(int64_t n) { return ((n + kScanTile - 1) / kScanTile + 1) * (int64_t)sizeof(unsigned long long); }

cudaError_t launch_zr_offsets(cudaStream_t st, const int32_t *counts, int64_t n, int64_t *offsets, void *workspace,
                              int64_t *total_mapped, int64_t *total_dev) {
    if (n == 0) return cudaSuccess;
    const int64_t tiles = (n + kScanTile - 1) / kScanTile;
    if (tiles > 2147483647LL) return cudaErrorInvalidValue;
    cudaError_t e = cudaMemsetAsync(workspace, 0, (size_t)tiles * sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    k_zr_offsets<<<(unsigned)tiles, kScanThreads, 0, st>>>(counts, n, offsets, (unsigned long long *)workspace, total_mapped, total_dev);
    return cudaGetLastError();
}

static int zr_grid(int device, int64_t units, int per_cta) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    int64_t grid = (units + per_cta - 1) / per_cta;
    if (grid > (int64_t)sms * 8) grid = (int64_t)sms * 8;
    return (int)(grid < 1 ? 1 : grid);
}

cudaError_t launch_zr_count(int device, cudaStream_t st, const int32_t *zz, int64_t nblocks, int32_t *counts,
                            uint64_t *masks, const int *run_flag) {
    if (nblocks == 0) return cudaSuccess;
    k_zr_count<<<zr_grid(device, nblocks, 32 * kZrWarps), kZrWarps * 32, 0, st>>>(zz, nblocks, counts,
                                                                                  (unsigned long long *)masks, run_flag);
    return cudaGetLastError();
}

cudaError_t launch_zr_write(int device, cudaStream_t st, const int32_t *zz, int64_t nblocks, int32_t eob,
                            const int64_t *offsets, const uint64_t *masks, void *out, int out_elem_size,
                            int64_t total_symbols) {
    if (nblocks == 0) return cudaSuccess;
    const int grid = zr_grid(device, nblocks, 32 * kZrWarps);
    const unsigned long long *mk = (const unsigned long long *)masks;
    if (masks && total_symbols >= 0 && total_symbols < (int64_t)kZrSparseLimit * nblocks) {      // sparse stream
        const int g2 = zr_grid(device, nblocks, 256);
        if (out_elem_size == 2) k_zr_write_sparse<int16_t><<<g2, 256, 0, st>>>(zz, nblocks, eob, offsets, mk, (int16_t *)out);
        else k_zr_write_sparse<int32_t><<<g2, 256, 0, st>>>(zz, nblocks, eob, offsets, mk, (int32_t *)out);
        return cudaGetLastError();
    }
    if (out_elem_size == 2) {
        if (!masks) return cudaErrorInvalidValue;
        k_zr_write<true, int16_t><<<grid, kZrWarps * 32, 0, st>>>(zz, nblocks, eob, offsets, mk, (int16_t *)out);
    } else if (masks) {
        k_zr_write<true, int32_t><<<grid, kZrWarps * 32, 0, st>>>(zz, nblocks, eob, offsets, mk, (int32_t *)out);
    } else {
        k_zr_write<false, int32_t><<<grid, kZrWarps * 32, 0, st>>>(zz, nblocks, eob, offsets, nullptr, (int32_t *)out);
    }
    return cudaGetLastError();
}

cudaError_t launch_zr_hist(int device, cudaStream_t st, const int32_t *zz, int64_t n_units, int64_t blocks_per_unit, int32_t eob,
                           int64_t lo, int64_t nbins, uint32_t *counts, uint32_t *outside) {
    if (n_units == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(counts, 0, sizeof(uint32_t) * (size_t)(n_units * nbins), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(outside, 0, sizeof(uint32_t) * (size_t)n_units, st);
    if (e != cudaSuccess || blocks_per_unit == 0) return e;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const size_t smem = (size_t)nbins * sizeof(unsigned);
    if ((e = cudaFuncSetAttribute(k_zr_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    // CTAs per unit: enough to fill the GPU about four times over across all units, each with at least a few
    // rounds of 256 blocks so that zeroing and flushing the shared histogram stays a small part of its work
    int64_t per_unit = ((int64_t)sms * 4 + n_units - 1) / n_units;
    const int64_t most = (blocks_per_unit + 4 * 32 * kZrWarps - 1) / (4 * 32 * kZrWarps);
    if (per_unit > most) per_unit = most;
    if (per_unit < 1) per_unit = 1;
    for (int64_t u0 = 0; u0 < n_units; u0 += 65535) {
        const int64_t nu = n_units - u0 < 65535 ? n_units - u0 : 65535;
        k_zr_hist<<<dim3((unsigned)per_unit, (unsigned)nu), kZrWarps * 32, smem, st>>>(
            zz + u0 * blocks_per_unit * 64, blocks_per_unit, eob, (int)lo, (int)nbins, counts + u0 * nbins, outside + u0);
    }
    return cudaGetLastError();
}

cudaError_t launch_zrd_mark(int device, cudaStream_t st, const int32_t *sym, int64_t n, int32_t eob, int32_t *is_eob) {
    if (n == 0) return cudaSuccess;
    k_zrd_mark<<<zr_grid(device, n, 256 * 8), 256, 0, st>>>(sym, n, eob, is_eob);
    return cudaGetLastError();
}

cudaError_t launch_zrd_ends(int device, cudaStream_t st, const int32_t *is_eob, const int64_t *rank, int64_t n,
                            int64_t want, int64_t *ends) {
    if (n == 0 || want == 0) return cudaSuccess;
    k_zrd_ends<<<zr_grid(device, n, 256 * 8), 256, 0, st>>>(is_eob, rank, n, want, ends);
    return cudaGetLastError();
}

cudaError_t launch_zrd_write(int device, cudaStream_t st, const int32_t *sym, const int64_t *ends, int64_t nblocks,
                             int32_t *out, int *err) {
    if (nblocks == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(err, 0, sizeof(int), st);
    if (e != cudaSuccess) return e;
    k_zrd_write<<<zr_grid(device, nblocks, kZrWarps), kZrWarps * 32, 0, st>>>(sym, ends, nblocks, out, err);
    return cudaGetLastError();
}

}  // namespace ivc
