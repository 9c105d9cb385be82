// ivc_zerorun.cu -- zero-run encoder over scan blocks (ivclab/entropy/zerorun.py:10-43; SURVEY.md
// section 8f row N2: the first host-side consumer of K1's output and the real bottleneck of
// IntraCodec.image2symbols).
//
// Per 64-coefficient block the reference emits: every non-zero value as itself; every run of zeros
// that is followed by a non-zero value as the pair (0, run_length); then EOB.  With m = the 64-bit
// "non-zero" mask of a block, L = its highest set bit and S = the zero positions below L that start a
// run (S = ~m & ((m << 1) | 1)), the block contributes popc(m) + 2*popc(S) + 1 symbols, and position
// p writes at offset popc(m & below(p)) + 2*popc(S & below(p)); a run starting at p has length
// ctz(m >> p).  Two passes: count (-> exclusive scan on the caller's side) and write.
// One warp stages 32 blocks (8 KB, coalesced 16-byte loads, row pitch 65 words so that the per-thread
// scans are bank-conflict free); one thread then owns one block.
#include "ivc_common.cuh"

namespace ivc {

constexpr int kZrWarps = 4;
constexpr int kZrPitch = 65;

__device__ __forceinline__ unsigned long long stage_and_mask(const int32_t *zz, int64_t nblocks, int64_t blk0, int *sm, int lane) {
    // coalesced: the 32 blocks of this warp are 8 KB contiguous = 512 int4
#pragma unroll 4
    for (int k = 0; k < 16; ++k) {
        const int id = lane + 32 * k, b = id >> 4, q = id & 15;
        int4 v = make_int4(0, 0, 0, 0);
        if (blk0 + b < nblocks) v = *reinterpret_cast<const int4 *>(zz + (blk0 + b) * 64 + q * 4);
        int *d = sm + b * kZrPitch + q * 4;
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
    __syncwarp();
    unsigned long long m = 0;
    const int *mine = sm + lane * kZrPitch;
#pragma unroll 16
    for (int p = 0; p < 64; ++p) m |= (unsigned long long)(mine[p] != 0) << p;
    return m;
}

__global__ void __launch_bounds__(kZrWarps * 32) k_zr_count(const int32_t *zz, int64_t nblocks, int32_t *counts) {
    __shared__ int sm_all[kZrWarps][32 * kZrPitch];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t nw = (int64_t)gridDim.x * kZrWarps;
    for (int64_t grp = (int64_t)blockIdx.x * kZrWarps + warp; grp * 32 < nblocks; grp += nw) {
        const unsigned long long m = stage_and_mask(zz, nblocks, grp * 32, sm_all[warp], lane);
        const int64_t blk = grp * 32 + lane;
        if (blk < nblocks) {
            int c = 1;                                                       // EOB
            if (m) {
                const unsigned long long below_top = (m == 0) ? 0 : ((2ull << (63 - __clzll((long long)m))) - 1ull);
                const unsigned long long S = ~m & ((m << 1) | 1ull) & below_top;
                c += __popcll(m) + 2 * __popcll(S);
            }
            counts[blk] = c;
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(kZrWarps * 32) k_zr_write(const int32_t *zz, int64_t nblocks, int32_t eob,
                                                            const int64_t *offsets, int32_t *out) {
    __shared__ int sm_all[kZrWarps][32 * kZrPitch];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t nw = (int64_t)gridDim.x * kZrWarps;
    for (int64_t grp = (int64_t)blockIdx.x * kZrWarps + warp; grp * 32 < nblocks; grp += nw) {
        unsigned long long m = stage_and_mask(zz, nblocks, grp * 32, sm_all[warp], lane);
        const int64_t blk = grp * 32 + lane;
        if (blk < nblocks) {
            const int *mine = sm_all[warp] + lane * kZrPitch;
            int32_t *o = out + offsets[blk];
            int p = 0;
            while (m >> p) {                                                 // there is a non-zero at or above p
                const int run = __ffsll((long long)(m >> p)) - 1;            // zeros before the next non-zero
                if (run > 0) { *o++ = 0; *o++ = run; p += run; }
                *o++ = mine[p];
                ++p;
                if (p >= 64) break;
            }
            *o = eob;
        }
        __syncwarp();
    }
}

cudaError_t launch_zr_count(int device, cudaStream_t st, const int32_t *zz, int64_t nblocks, int32_t *counts) {
    if (nblocks == 0) return cudaSuccess;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    int64_t grid = (nblocks + 32 * kZrWarps - 1) / (32 * kZrWarps);
    if (grid > (int64_t)sms * 8) grid = (int64_t)sms * 8;
    k_zr_count<<<(unsigned)grid, kZrWarps * 32, 0, st>>>(zz, nblocks, counts);
    return cudaGetLastError();
}

cudaError_t launch_zr_write(int device, cudaStream_t st, const int32_t *zz, int64_t nblocks, int32_t eob,
                            const int64_t *offsets, int32_t *out) {
    if (nblocks == 0) return cudaSuccess;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    int64_t grid = (nblocks + 32 * kZrWarps - 1) / (32 * kZrWarps);
    if (grid > (int64_t)sms * 8) grid = (int64_t)sms * 8;
    k_zr_write<<<(unsigned)grid, kZrWarps * 32, 0, st>>>(zz, nblocks, eob, offsets, out);
    return cudaGetLastError();
}

}  // namespace ivc
