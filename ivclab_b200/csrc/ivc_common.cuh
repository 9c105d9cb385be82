// ivc_common.cuh -- shared declarations between the kernel translation units and the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/ivclab_b200.h"

namespace ivc {

// launchers implemented in ivc_transform.cu
cudaError_t launch_forward(int device, cudaStream_t st, const void *img, int64_t n, int64_t H, int64_t W, int C,
                           int64_t frame_stride, const void *table, int table_dtype, int32_t *out,
                           const void *ref, const int64_t *mv, int sr, void *pred_out, bool pframe, int out_channels = 3,
                           const int *run_flag = nullptr,        // P-frame kernels: run only if *run_flag != 0 (null: always)
                           int32_t *zr_counts = nullptr, uint64_t *zr_masks = nullptr, bool *zr_done = nullptr);
                           // zr_*: per scan block the zero-run symbol count and non-zero mask, when the kernel variant that runs
                           // can emit them (*zr_done says whether it did)
cudaError_t launch_inverse(int device, cudaStream_t st, const int32_t *zz, int64_t n, int64_t Hp, int64_t Wp, int Czz,
                           const void *table, int table_dtype, void *out, int mode,
                           const void *pred, const void *ref, const int64_t *mv, int sr);
int64_t inverse_sse_tiles(int64_t n, int64_t Hp, int64_t Wp);
cudaError_t launch_inverse_sse(int device, cudaStream_t st, const int32_t *zz, int64_t n, int64_t Hp, int64_t Wp,
                               const void *table, int table_dtype, void *out, const void *orig_rgb8,
                               int64_t orig_frame_stride, int sse_mode, double *partial, double *sse_out);
cudaError_t launch_dct(int device, cudaStream_t st, bool inverse, const void *x, int x_dtype, int64_t n0, int64_t n1,
                       int64_t C, const int64_t s[5], void *out, bool f32, int norm = 0);
cudaError_t launch_quant(int device, cudaStream_t st, bool dequant, const void *x, int x_dtype, int64_t n0, int64_t n1,
                         int64_t C, const int64_t s[5], const void *table, int table_dtype, bool f32, int32_t *out);
cudaError_t launch_zigzag(int device, cudaStream_t st, bool inverse, const void *x, int elem_size, int64_t nblocks,
                          void *out);

// launchers implemented in ivc_motion.cu
cudaError_t launch_me_exact(int device, cudaStream_t st, const void *ref, const void *cur, bool f32, int64_t n,
                            int64_t H, int64_t W, int64_t ref_fs, int64_t cur_fs, int sr, int64_t *mv,
                            int *flag, int run_if);
// pf_zz != nullptr: the fused kernel -- after the search every tile codes its blocks (MC + residual + DCT + quantise +
// zig-zag, pf_och scan channels per block) from the staged bytes; only where me_pf_fusable(dtype, sr)
cudaError_t launch_f64_to_u8(int device, cudaStream_t st, const void *src, int64_t frame_stride, int64_t n, int64_t plane,
                             void *dst, int *flag);
cudaError_t launch_me_int(int device, cudaStream_t st, const void *ref, const void *cur, int dtype, int64_t n,
                          int64_t H, int64_t W, int64_t ref_fs, int64_t cur_fs, int sr, int64_t *mv, int *flag,
                          int check, const void *pf_table = nullptr, int pf_table_dtype = 0, int32_t *pf_zz = nullptr,
                          int pf_och = 3, int32_t *zr_counts = nullptr, uint64_t *zr_masks = nullptr);
bool me_pf_fusable(int dtype, int sr);
cudaError_t launch_pframe_step(int device, cudaStream_t st, const void *ref, const void *cur, int64_t n, int64_t H, int64_t W,
                               int sr, const void *table, int table_dtype, int out_channels, int64_t *mv, int32_t *zz,
                               double *recon);
cudaError_t launch_me_wrap(int device, cudaStream_t st, const void *ref, const void *cur, int dtype, int64_t n, int64_t H,
                           int64_t W, int64_t ref_fs, int64_t cur_fs, int sr, int64_t *mv);
cudaError_t launch_mc(int device, cudaStream_t st, const void *ref, int elem_size, int64_t n, int64_t H, int64_t W,
                      int64_t C, const int64_t *mv, int sr, void *out);

// implemented in ivc_metrics.cu
int sse_chunks(int64_t n_units, int64_t unit_elems, int sms);
cudaError_t launch_sse(int device, cudaStream_t st, const void *a, int a_dtype, const void *b, int b_dtype,
                       int64_t n_units, int64_t unit_elems, int a_div, double *partial, double *out);

// implemented in ivc_zerorun.cu
cudaError_t launch_zr_count(int device, cudaStream_t st, const int32_t *zz, int64_t nblocks, int32_t *counts,
                            uint64_t *masks, const int *run_flag = nullptr);     // run_flag: run only if *run_flag != 0
cudaError_t launch_zr_write(int device, cudaStream_t st, const int32_t *zz, int64_t nblocks, int32_t eob,
                            const int64_t *offsets, const uint64_t *masks, void *out, int out_elem_size,
                            int64_t total_symbols);

cudaError_t launch_zr_hist(int device, cudaStream_t st, const int32_t *zz, int64_t n_units, int64_t blocks_per_unit, int32_t eob,
                           int64_t lo, int64_t nbins, uint32_t *counts, uint32_t *outside);

cudaError_t launch_zrd_mark(int device, cudaStream_t st, const int32_t *sym, int64_t n, int32_t eob, int32_t *is_eob);
cudaError_t launch_zrd_ends(int device, cudaStream_t st, const int32_t *is_eob, const int64_t *rank, int64_t n,
                            int64_t want, int64_t *ends);
cudaError_t launch_zrd_write(int device, cudaStream_t st, const int32_t *sym, const int64_t *ends, int64_t nblocks,
                             int32_t *out, int *err);

int64_t zr_offsets_workspace_bytes(int64_t n);
cudaError_t launch_zr_offsets(cudaStream_t st, const int32_t *counts, int64_t n, int64_t *offsets, void *workspace,
                              int64_t *total_mapped, int64_t *total_dev);
cudaError_t launch_post_words(cudaStream_t st, const int64_t *src, int64_t *dst_mapped, int n);

// symbol statistics (ivc_metrics.cu)
cudaError_t launch_hist(int device, cudaStream_t st, const void *x, int dtype, int64_t n, int64_t lo, int64_t nbins,
                        int64_t hot, uint64_t *counts);
cudaError_t launch_minmax(int device, cudaStream_t st, const void *x, int dtype, int64_t n, int64_t *out);

// colour (ivc_color.cu) and the RGB front end of K1 (ivc_transform.cu)
cudaError_t launch_color(int device, cudaStream_t st, bool to_rgb, const void *in, int in_dtype, int64_t npix, double *out);
cudaError_t launch_rgb8_luma8(int device, cudaStream_t st, const void *rgb, int64_t npix, void *out, void *out64);
cudaError_t launch_forward_rgb8(int device, cudaStream_t st, const void *rgb, int64_t n, int64_t H, int64_t W,
                                int64_t frame_stride_bytes, const void *table, int table_dtype, int32_t *out,
                                int32_t *zr_counts = nullptr, uint64_t *zr_masks = nullptr, int nq = 1);

}  // namespace ivc
