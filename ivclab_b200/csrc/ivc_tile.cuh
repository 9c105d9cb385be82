// ivc_tile.cuh -- pieces shared by the fused transform kernels (ivc_transform.cu) and the fused search + P-frame
// forward kernel (ivc_motion.cu): shared-memory / bulk-copy primitives, the branch-free quantiser, staging geometry.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "ivc_dct.cuh"
#include "../../include/ivclab_b200.h"

namespace ivc {

constexpr int kStageUF = 204;        // int32 staging of the TMA forward kernels: 3 chunks * 64 ints + 12 pad per block, which puts the
                                     // zig-zag SCATTER of a warp on 16 instead of 20 bank wavefronts per 8 stores
constexpr int kP3TU = 136;           // transposition buffer of the 8-block tiles: doubles per u-plane, 16 rows * 8 + 8 skew (64 B)

// quantiser tables in shared memory, one copy per CTA:
//   fwd: rt[ch*64+k] = fl(1/t), t[ch*64+k]                    (raster k = 8v + j)
//   inv: tT[ch*64 + j*8 + r] = t[ch][8r + j]                  (transposed so lanes r are contiguous)
__device__ __forceinline__ double load_table_elem(const void *table, int table_dtype, int i) {
    return table_dtype == IVC_F32 ? (double)((const float *)table)[i] : ((const double *)table)[i];
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (spin > (1u << 24)) __trap();          // a lost copy must not hang the GPU
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// branch-free quantiser: always produces the fast-path integer and folds "this sample needs the
// exact division" into two running values (see quantize_f64 in ivc_dct.cuh for the argument).
struct QuantGuard {
    int mx = 0;                      // max of |y| high words
    unsigned nz = 0xffffffffu;       // min of (frac16 ^ 0x8000): 0 <=> some sample sits exactly on a half
    __device__ __forceinline__ int q(double x, double rt) {
        const double y = __dmul_rn(x, rt);
        const int lo = __double2loint(__dadd_rn(y, 103079215104.0));         // 1.5 * 2^36
        mx = max(mx, __double2hiint(y) & 0x7fffffff);
        nz = min(nz, (unsigned)((lo & 0xFFFF) ^ 0x8000));
        return (lo + 0x8000) >> 16;
    }
    __device__ __forceinline__ bool risky() const { return (mx >= 0x40DFFFC0) | (nz == 0u); }   // |y| >= 2^15 - 1 (the rounding add would wrap at 32767.5), NaN, tie
};

// ---- zero-run coder arithmetic on a block's 64-bit "non-zero" mask (ivclab/entropy/zerorun.py:10-43) ----
// S = the zero positions below the highest set bit that start a run; a block emits popc(m) + 2 popc(S) + 1 symbols.
__device__ __forceinline__ unsigned long long zr_run_starts(unsigned long long m) {
    const unsigned long long below_top = (2ull << (63 - __clzll((long long)m))) - 1ull;        // m != 0
    return ~m & ((m << 1) | 1ull) & below_top;
}
__device__ __forceinline__ int zr_block_count(unsigned long long m) {
    return m ? 1 + __popcll(m) + 2 * __popcll(zr_run_starts(m)) : 1;
}
// The masks and symbol counts of NB scan blocks that sit, 64 int32 each, in a warp's staging area (block j at
// stage + off(j) bytes): two ballots per block, lane j keeps block j.  What the zero-run coder's count pass would
// compute by re-reading the indices from HBM -- here they are still in shared memory.
template <int NB, typename Off>
__device__ __forceinline__ void zr_masks_from_staging(const unsigned char *stage, Off off, int lane, unsigned long long &mask_out) {
    mask_out = 0ull;
#pragma unroll
    for (int j = 0; j < NB; ++j) {
        const int *sb = reinterpret_cast<const int *>(stage + off(j));
        const unsigned lo = __ballot_sync(0xffffffffu, sb[lane] != 0), hi = __ballot_sync(0xffffffffu, sb[32 + lane] != 0);
        if (lane == j) mask_out = (unsigned long long)lo | ((unsigned long long)hi << 32);
    }
}

}  // namespace ivc
