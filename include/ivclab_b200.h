/*
 * ivclab_b200.h -- C ABI of the B200-native ivclab per-block coding loop.
 *
 * This is the drop-in boundary: a reference-side binding (ctypes from Python,
 * see INTEGRATION.md) needs nothing but this header and libivcb200.so.
 * Every entry point
 *   - takes plain device pointers + sizes (no torch / numpy types),
 *   - enqueues its kernels on the caller's stream (cudaStream_t passed as
 *     void*; NULL = legacy default stream) on `device`, and returns without
 *     synchronising,
 *   - allocates nothing and keeps no mutable global state (the caller owns
 *     inputs, outputs and workspaces),
 *   - returns IVC_OK (0) or a negative IVC_ERR_* code; it never throws.
 *
 * Reference interfaces replaced (file:line under n2oblife/ivclab):
 *   ivclab/signal/dct.py:12-46            DiscreteCosineTransform.transform / inverse_transform
 *   ivclab/quantization/patchquant.py:39-78  PatchQuant.get_quantization_table / quantize / dequantize
 *   ivclab/utils/shape.py:21-36           ZigZag.flatten / unflatten
 *   ivclab/video/motion.py:8-97           MotionCompensator.compute_motion_vector / reconstruct_with_motion_vector
 *   ivclab/image/intracodec.py:66-75,115-124   the chained calls the fused entry points replace
 *   ivclab/video/videocodec.py:68-74      P-frame residual / reconstruction adds
 *
 * Layout vocabulary: a "patched" array is [n0, n1, C, 8, 8] (Patcher.patch,
 * shape.py:45-54); a "scan" array is [n0, n1, C, 64] in zig-zag order; an
 * image is HWC, a luma plane is HW.  Strides are in ELEMENTS.
 */
#ifndef IVCLAB_B200_H
#define IVCLAB_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IVC_ABI_VERSION 10

/* element types */
#define IVC_U8   0
#define IVC_I32  1
#define IVC_F32  2
#define IVC_F64  3
#define IVC_I64  4
#define IVC_I16  5

/* status codes */
#define IVC_OK            0
#define IVC_ERR_ARG      -1   /* null pointer / negative size / bad flag            */
#define IVC_ERR_DTYPE    -2   /* dtype combination not supported by this entry      */
#define IVC_ERR_SHAPE    -3   /* shape the reference would reject (e.g. H%8 != 0)   */
#define IVC_ERR_CUDA     -4   /* a CUDA call failed; see ivc_last_cuda_error()      */
#define IVC_ERR_WORKSPACE -5  /* workspace too small                                */

/* ME kernel selection */
#define IVC_ME_AUTO   0   /* integer kernel if both frames are integer-valued in [0,255], else exact-float */
#define IVC_ME_EXACT  1   /* IEEE order-exact float kernel (numpy summation order, no FMA)                 */
#define IVC_ME_INT    2   /* packed-u8 integer kernel; caller guarantees integer-valued [0,255] frames     */

int         ivc_abi_version(void);
const char *ivc_build_info(void);            /* "ivclab_b200 <ver> sm_100a nvcc <ver>" */
const char *ivc_error_string(int status);
int         ivc_last_cuda_error(void);       /* cudaError_t of the last failing CUDA call on this thread */
const char *ivc_last_cuda_error_string(void);

/* ---- a2/a3: 2-D 8x8 DCT-II / DCT-III, orthonormal (dct.py:12-46) -----------------------------
 * x: patched [n0,n1,C,8,8] with arbitrary element strides (the reference feeds a strided view of
 * the HWC image, intracodec.py:66).  out: C-contiguous [n0,n1,C,8,8].
 * in f32 -> out f32 (float arithmetic); in u8/i32/f64 -> out f64.  Arithmetic is op-for-op the
 * sequence scipy's ducc0 back-end executes, so results are bit-identical to scipy.fft.dct/idct. */
int ivc_dct8x8(int device, void *stream, int inverse,
               const void *x, int x_dtype, int64_t n0, int64_t n1, int64_t C,
               const int64_t strides[5], void *out, int out_dtype);

/* The same with scipy's other normalisations -- the reference forwards `self.norm` to scipy.fft.dct / idct
 * (dct.py:9-10,24,26,42,44); nothing in ivclab uses one, the class accepts them.  norm: IVC_NORM_ORTHO, IVC_NORM_BACKWARD
 * (scipy's None / "backward": unscaled forward transform, inverse / 16 per axis), IVC_NORM_FORWARD (forward / 16 per axis,
 * unscaled inverse).  For length 8 ducc0's factor is a power of two in every mode, so the results stay bit-identical. */
#define IVC_NORM_ORTHO 0
#define IVC_NORM_BACKWARD 1
#define IVC_NORM_FORWARD 2
int ivc_dct8x8_norm(int device, void *stream, int inverse, int norm,
                    const void *x, int x_dtype, int64_t n0, int64_t n1, int64_t C,
                    const int64_t strides[5], void *out, int out_dtype);

/* ---- a6: PatchQuant.quantize (patchquant.py:56-60) ------------------------------------------
 * out[n0,n1,Cout,8,8] = int32(rint(x / table[c])) with numpy broadcasting of C against 3
 * (C==1 -> Cout=3, C==3 -> Cout=3).  table: device [3,8,8] in table_dtype (F32 or F64).
 * compute_dtype: F32 only when numpy would divide in float32 (x u8/f32 and table f32), else F64. */
int ivc_quantize(int device, void *stream,
                 const void *x, int x_dtype, int64_t n0, int64_t n1, int64_t C,
                 const int64_t strides[5],
                 const void *table, int table_dtype, int compute_dtype,
                 int32_t *out);

/* ---- a7: PatchQuant.dequantize (patchquant.py:74-78) ----------------------------------------
 * out[n0,n1,3,8,8] = int32(trunc(q * table[c])), product in compute_dtype (F64 for int32 q). */
int ivc_dequantize(int device, void *stream,
                   const void *q, int q_dtype, int64_t n0, int64_t n1, int64_t C,
                   const int64_t strides[5],
                   const void *table, int table_dtype, int compute_dtype,
                   int32_t *out);

/* ---- a9/a10: ZigZag.flatten / unflatten (shape.py:21-36) -------------------------------------
 * nblocks contiguous blocks of 64 elements of elem_size bytes (1,2,4,8).
 * inverse=0: out[b][order[k]] = x[b][k];  inverse=1: out[b][k] = x[b][order[k]]. */
int ivc_zigzag(int device, void *stream, int inverse,
               const void *x, int elem_size, int64_t nblocks, void *out);

/* ---- fused intra forward: patch -> DCT -> quantize -> flatten (intracodec.py:66-75) ----------
 * img: n_frames HWC images, C in {1,3}, dtype F64, pixel stride C, row stride W*C,
 * frame stride frame_stride elements.  H, W multiples of 8.
 * out: [n_frames, H/8, W/8, 3, 64] int32. */
int ivc_intra_forward(int device, void *stream,
                      const void *img, int dtype, int64_t n_frames, int64_t H, int64_t W, int64_t C,
                      int64_t frame_stride,
                      const void *table, int table_dtype,
                      int32_t *out);

/* ---- fused intra inverse: unflatten -> dequantize -> IDCT -> un-patch (intracodec.py:115-124)
 * zz: [n_frames, Hp, Wp, C, 64] int32, C in {1,3}.  out: [n_frames, 8*Hp, 8*Wp, 3] F64. */
int ivc_intra_inverse(int device, void *stream,
                      const int32_t *zz, int64_t n_frames, int64_t Hp, int64_t Wp, int64_t C,
                      const void *table, int table_dtype,
                      void *out, int out_dtype);

/* ivc_intra_inverse (C = 3) with ycbcr2rgb + clip applied in the store (ivclab/signal/color.py:39-63): the last
 * step of symbols2image for colour images (image/intracodec.py:139-141) without a second pass.  out: float64 RGB in
 * [0, 255], bit-identical to ivc_ycbcr2rgb(ivc_intra_inverse(...)). */
int ivc_intra_inverse_rgb(int device, void *stream,
                          const int32_t *zz, int64_t n_frames, int64_t Hp, int64_t Wp,
                          const void *table, int table_dtype, void *rgb_out);

/* ---- fused intra inverse + distortion: one rate-distortion point of a sweep -----------------------
 * Decodes like ivc_intra_inverse (C = 3) and, in the same kernel, measures the squared error of every frame
 * against its uint8 RGB original (orig_rgb8: [n_frames, 8*Hp, 8*Wp, 3], W % 16 == 0, frames
 * orig_frame_stride_bytes apart, 16-byte aligned):
 *   mode IVC_DIST_RGB   : sse[f] = sum (orig - ycbcr2rgb(rec))^2 -- what calc_psnr(img, symbols2image(...))
 *                         measures (utils/metrics.py:3-40, image/intracodec.py:139-141, signal/color.py:39-63);
 *   mode IVC_DIST_YCBCR : sse[f] = sum (rgb2ycbcr(orig) - rec)^2.
 * out may be NULL: the reconstruction is then not stored (decode + PSNR moves 15 bytes per pixel).
 * Deterministic (per-tile partial sums added in a fixed order).  workspace: ivc_intra_inverse_sse_workspace_bytes(). */
#define IVC_DIST_RGB   1
#define IVC_DIST_YCBCR 2
int64_t ivc_intra_inverse_sse_workspace_bytes(int64_t n_frames, int64_t Hp, int64_t Wp);
int ivc_intra_inverse_sse(int device, void *stream,
                          const int32_t *zz, int64_t n_frames, int64_t Hp, int64_t Wp,
                          const void *table, int table_dtype,
                          void *out, const void *orig_rgb8, int64_t orig_frame_stride_bytes, int mode,
                          void *workspace, int64_t workspace_bytes, double *sse_out);

/* ---- a13: MotionCompensator.compute_motion_vector (motion.py:8-58) ---------------------------
 * ref, cur: n_frames luma planes [H,W] (dtype F32 or F64, both the same), contiguous rows; or U8: uint8 planes
 * holding the frames' VALUES (float semantics -- what the reference computes after casting them to float; numpy's
 * wrap-around on uint8-dtype frames is ivc_me_full_search_intdtype) -- always served by the packed-integer kernel.
 * mv_out: [n_frames, H/8, W/8, 1] int64, index = (dy+sr)*(2sr+1) + (dx+sr); first minimum in
 * (dy asc, dx asc) order over in-bounds candidates.
 * workspace: needed for IVC_ME_AUTO only (holds the device-side "not an integer frame" flag),
 * ivc_me_workspace_bytes() bytes; NULL otherwise.  Optional: with ivc_me_workspace_bytes_planes() bytes (AUTO and INT modes,
 * F64 frames, search_range >= 8) the frames are converted to uint8 planes -- and validated -- once, and the search reads the
 * planes; two views of one contiguous sequence (cur = ref + one frame) are converted as n_frames + 1 planes.  Same vectors. */
int64_t ivc_me_workspace_bytes(int64_t n_frames, int64_t H, int64_t W);
int64_t ivc_me_workspace_bytes_planes(int64_t n_frames, int64_t H, int64_t W);
int ivc_me_full_search(int device, void *stream,
                       const void *ref, const void *cur, int dtype,
                       int64_t n_frames, int64_t H, int64_t W,
                       int64_t ref_frame_stride, int64_t cur_frame_stride,
                       int search_range, int mode,
                       int64_t *mv_out, void *workspace, int64_t workspace_bytes);

/* a13 on INTEGER-dtype frames (U8, I16 or I32, both frames the same): numpy evaluates (block - ref_block)**2 in the
 * frames' dtype -- a uint8 difference and its square wrap mod 256, 255**2 is -511 in int16 -- and sums in
 * uint64 / int64 (motion.py:46).  This entry replays exactly that arithmetic, so the
 * vectors equal the reference's on such inputs. */
int ivc_me_full_search_intdtype(int device, void *stream,
                                const void *ref, const void *cur, int dtype, int64_t n_frames,
                                int64_t H, int64_t W, int64_t ref_frame_stride, int64_t cur_frame_stride,
                                int search_range, int64_t *mv_out);

/* ---- a14: MotionCompensator.reconstruct_with_motion_vector (motion.py:60-97) -----------------
 * ref: [n_frames, H, W, C] of elem_size-byte elements; mv: [n_frames, H/8, W/8, 1] int64.
 * Blocks whose source window leaves the frame are zero. */
int ivc_mc_reconstruct(int device, void *stream,
                       const void *ref, int elem_size, int64_t n_frames, int64_t H, int64_t W, int64_t C,
                       const int64_t *mv, int search_range, void *out);

/* ---- a15 fused: P-frame encoder half (videocodec.py:68-71 + intracodec.py:66-75) -------------
 * pred = MC(ref, mv); residual = cur - pred; zz = flatten(quantize(dct(patch(residual)))).
 * cur/ref: [n_frames,H,W] F64.  pred_out may be NULL.  zz_out: [n_frames,Hp,Wp,3,64] int32.
 * cur, zz_out, pred_out must be 16-byte aligned; ref needs 8 bytes only (a 16-byte aligned ref takes the faster
 * tensor-map gather, any other the cp.async gather: same results). */
int ivc_pframe_forward(int device, void *stream,
                       const void *cur, const void *ref, const int64_t *mv, int dtype,
                       int64_t n_frames, int64_t H, int64_t W, int search_range,
                       const void *table, int table_dtype,
                       void *pred_out, int32_t *zz_out);
/* The same with out_channels scan blocks stored per image block: 3 = the reference's layout (above); 2 = channels 0
 * and 1 only, zz_out [n_frames,Hp,Wp,2,64].  PatchQuant's table is [lum, chrom, chrom] (patchquant.py:40), so channel
 * 2 of the broadcast luma residual repeats channel 1 bit for bit: a pipeline that ships symbols to a host need not
 * compute, store or send it (ivc_pframe_inverse reads channel 0 at stride Czz = 2). */
int ivc_pframe_forward_ch(int device, void *stream,
                          const void *cur, const void *ref, const int64_t *mv, int dtype,
                          int64_t n_frames, int64_t H, int64_t W, int search_range,
                          const void *table, int table_dtype,
                          void *pred_out, int out_channels, int32_t *zz_out);

/* ---- a13 + a15 fused: search AND P-frame encoder half in one call (videocodec.py:52 + :68-71 + intracodec.py:66-75) --
 * mv = compute_motion_vector(ref, cur); zz = flatten(quantize(dct(patch(cur - MC(ref, mv))))) with out_channels (3 or 2,
 * see ivc_pframe_forward_ch) scan blocks per image block.  cur/ref [n_frames,H,W] F64; mode as ivc_me_full_search.
 * For +-4 searches of integer-valued frames ONE kernel does both: its tiles code their blocks from the bytes the search
 * staged, so the frames are read once (16 B/pixel instead of 16 + 16).  In IVC_ME_AUTO mode that kernel validates the
 * frames and otherwise leaves everything to the exact search and the stand-alone forward kernel, which are enqueued
 * behind it and run only if it raised the device flag (workspace as for ivc_me_full_search).  Other search ranges,
 * IVC_ME_EXACT: the two stand-alone kernels.  Results are identical in every case.
 * dtype IVC_U8: cur/ref are uint8 PLANES holding the frames' values (no alignment requirement, no workspace); served by
 * the fused kernel only, i.e. for search_range 4 -- IVC_ERR_DTYPE otherwise. */
int ivc_pframe_search_forward(int device, void *stream,
                              const void *cur, const void *ref, int dtype,
                              int64_t n_frames, int64_t H, int64_t W, int search_range, int mode,
                              const void *table, int table_dtype, int out_channels,
                              int64_t *mv_out, int32_t *zz_out, void *workspace, int64_t workspace_bytes);

/* The same, and the zero-run coder's count pass with it: counts_out[b] (int32) = the symbols ZeroRunCoder.encode emits for
 * scan block b of zz_out (in its (h w c) order), masks_out[b] (uint64) = the block's non-zero mask -- exactly what
 * ivc_zerorun_count_masks would compute from zz_out, taken while the block is still in shared memory, so a pipeline goes
 * straight on to ivc_zerorun_offsets / ivc_zerorun_write_masks without re-reading the indices. */
int ivc_pframe_search_forward_zr(int device, void *stream,
                                 const void *cur, const void *ref, int dtype,
                                 int64_t n_frames, int64_t H, int64_t W, int search_range, int mode,
                                 const void *table, int table_dtype, int out_channels,
                                 int64_t *mv_out, int32_t *zz_out, void *workspace, int64_t workspace_bytes,
                                 int32_t *counts_out, uint64_t *masks_out);

/* ---- a13 + a15 + a15: one whole closed-loop P-frame step in ONE kernel (exercises/ch4/E4-1.py:257-306) --------------
 * mv = compute_motion_vector(ref, cur) with the order-exact search (ref is a reconstruction: not integer-valued);
 * zz = flatten(quantize(dct(patch(cur - MC(ref, mv))))) (out_channels scan blocks per image block);
 * recon = MC(ref, mv) + idct(dequantize(unflatten(zz[..., 0, :]))) with the luminance table, i.e. the decoder's next
 * reference for decode = "luma" (channel 0 of every block -- the reference's own `symbols2image` with a 2-D shape reads
 * other blocks, SURVEY.md section 0 item 10, and needs the three stand-alone kernels).  Per block the three stages need
 * only the block, its window and its vector, so the warp that finds the vector codes and reconstructs the block from
 * shared memory: the frame pair is read once.  cur/ref/recon_out [n_frames,H,W] F64, n_frames <= 65535. */
int ivc_pframe_step(int device, void *stream,
                    const void *cur, const void *ref, int dtype,
                    int64_t n_frames, int64_t H, int64_t W, int search_range,
                    const void *table, int table_dtype, int out_channels,
                    int64_t *mv_out, int32_t *zz_out, void *recon_out);

/* ---- a15 fused: P-frame decoder half (intracodec.py:115-124 + videocodec.py:74) --------------
 * recon = pred + idct(dequantize(unflatten(zz[..., 0, :])))[channel 0]  (luminance table).
 * zz: [n_frames,Hp,Wp,Czz,64] int32 (channel 0 is used).  Either pred != NULL, or pred == NULL
 * and (ref, mv) are given and the prediction is re-gathered.  recon_out: [n_frames,H,W] F64. */
int ivc_pframe_inverse(int device, void *stream,
                       const int32_t *zz, int64_t Czz,
                       const void *pred, const void *ref, const int64_t *mv, int dtype,
                       int64_t n_frames, int64_t H, int64_t W, int search_range,
                       const void *table, int table_dtype,
                       void *recon_out);

/* ---- N3 (next row): calc_mse / calc_psnr (ivclab/utils/metrics.py:3-40) ------------------------
 * sse_out[u] = sum_i (double(a[u][i / a_broadcast]) - double(b[u][i]))^2 over unit_elems elements of b,
 * for n_units units (frames).  a_broadcast = 3 pairs a gray `a` with an RGB `b` (metrics.py:16-19),
 * else 1; a_broadcast = IVC_SSE_RGB8_AS_YCBCR takes a = uint8 RGB and b = float64 YCbCr of the same shape and
 * compares rgb2ycbcr(a) with b without materialising it (same bits as the two-step form).
 * Deterministic two-stage reduction; mse = sse / unit_elems, psnr = 20*log10(max/sqrt(mse))
 * are left to the caller.  workspace: ivc_sse_workspace_bytes() bytes. */
#define IVC_SSE_RGB8_AS_YCBCR 103
int64_t ivc_sse_workspace_bytes(int64_t n_units, int64_t unit_elems);
int ivc_sum_squared_error(int device, void *stream,
                          const void *a, int a_dtype, const void *b, int b_dtype,
                          int64_t n_units, int64_t unit_elems, int a_broadcast,
                          void *workspace, int64_t workspace_bytes, double *sse_out);

/* ---- N2 (next row): ZeroRunCoder.encode (ivclab/entropy/zerorun.py:10-43) ---------------------
 * zz: nblocks contiguous scan blocks of 64 int32.  Two passes around an exclusive scan that the
 * caller performs: counts_out[b] = number of symbols block b emits (non-zeros + 2 per zero run that
 * precedes a non-zero + 1 EOB); then symbols are written at offsets[b] (exclusive prefix sum of the
 * counts, int64).  The symbol stream equals the reference's list bit for bit. */
int ivc_zerorun_count(int device, void *stream, const int32_t *zz, int64_t nblocks, int32_t *counts_out);
int ivc_zerorun_write(int device, void *stream, const int32_t *zz, int64_t nblocks, int32_t end_of_block,
                      const int64_t *offsets, int32_t *symbols_out);
/* The same two passes with the per-block 64-bit non-zero masks handed from the first to the second
 * (masks: nblocks uint64): the write pass then fetches only the parts of each block that hold symbols.
 * total_symbols: the stream length if the caller knows it (it sized symbols_out with it), else -1 -- sparse
 * streams (< 16 symbols per block on average) take a one-thread-per-block kernel. */
int ivc_zerorun_count_masks(int device, void *stream, const int32_t *zz, int64_t nblocks, int32_t *counts_out,
                            uint64_t *masks_out);
int ivc_zerorun_write_masks(int device, void *stream, const int32_t *zz, int64_t nblocks, int32_t end_of_block,
                            const int64_t *offsets, const uint64_t *masks, int32_t *symbols_out, int64_t total_symbols);

/* The write pass with 16-bit symbols: a lossless TRANSFER format for callers that know every symbol fits (e.g.
 * |coefficient| <= 2040 / min(table) for 8-bit images; end_of_block must fit too).  Values are truncated to int16
 * without a check -- the reference's dtype is int32 (ivc_zerorun_write). */
int ivc_zerorun_write_masks_i16(int device, void *stream, const int32_t *zz, int64_t nblocks, int32_t end_of_block,
                                const int64_t *offsets, const uint64_t *masks, int16_t *symbols_out, int64_t total_symbols);

/* The marginal histogram of the symbol stream WITHOUT the stream: what IntraCodec.train_huffman_from_image
 * (intracodec.py:160-166: image2symbols -> min/max -> stats_marg) needs of ZeroRunCoder.encode's output.
 * zz: n_units x blocks_per_unit contiguous scan blocks (a unit = one frame's blocks).  counts_out
 * [n_units, n_bins] uint32 (zeroed by the call): counts_out[u][k] = number of symbols of unit u's stream equal to
 * lo + k, i.e. np.histogram(ZeroRunCoder.encode(unit u), bins=np.arange(lo, lo + n_bins + 1))[0] whenever every
 * symbol lies in [lo, lo + n_bins) -- outside_out[u] (uint32) counts the symbols that do not.  n_bins <= 49152. */
int ivc_zerorun_symbol_histogram(int device, void *stream, const int32_t *zz, int64_t n_units, int64_t blocks_per_unit,
                                 int32_t end_of_block, int64_t lo, int64_t n_bins, uint32_t *counts_out,
                                 uint32_t *outside_out);

/* The scan between the two passes, for callers that do not want to bring their own: offsets_out[b] = sum of
 * counts[0..b) (int64), in one kernel (decoupled look-back over 4096-count tiles).  The grand total is written
 * to total_dev_out (device int64, may be NULL) and, by the kernel itself, to total_mapped_out (MAPPED pinned host
 * memory, may be NULL) -- a pipeline learns the stream length without a copy-engine operation.
 * workspace: ivc_zerorun_offsets_workspace_bytes(nblocks) bytes of device memory. */
int64_t ivc_zerorun_offsets_workspace_bytes(int64_t nblocks);
int ivc_zerorun_offsets(int device, void *stream, const int32_t *counts, int64_t nblocks, int64_t *offsets_out,
                        void *workspace, int64_t workspace_bytes, int64_t *total_mapped_out, int64_t *total_dev_out);

/* Post n (<= 32) int64 words from device memory to MAPPED pinned host memory with a kernel (no copy engine):
 * how a pipeline learns a symbol-stream length (the last element of the caller's prefix sum) without the
 * small copy queueing behind bulk transfers.  dst_mapped: a cudaHostAlloc'd / pinned pointer valid on the device. */
int ivc_post_words_to_host(int device, void *stream, const int64_t *src, int64_t *dst_mapped, int n);

/* ---- N2 (next row): ZeroRunCoder.decode (ivclab/entropy/zerorun.py:44-87) ---------------------
 * Three passes around an inclusive scan that the caller performs.  mark: is_eob[i] = 1 iff symbols[i] is
 * an end-of-block in a symbol slot (not the run length that follows a zero marker).  ends: with
 * rank = inclusive prefix sum of is_eob (int64), ends_out[k] = position of the (k+1)-th EOB for the
 * first n_blocks blocks -- later symbols are ignored, as the reference stops after h*w*c blocks
 * (zerorun.py:62).  write: block k = symbols (ends[k-1], ends[k]) expanded to 64 int32 coefficients.
 * *err_out (device int) = 0, or bit 0: a block expands to more than 64 coefficients (the reference's
 * "Block size exceeded"), bit 1: a run length <= 0 (never emitted by the encoder; the reference accepts
 * it, this decoder does not).  The caller checks rank[n-1] >= n_blocks (else the reference raises
 * "Unexpected end of encoded symbols" / "Expected N blocks"). */
int ivc_zerorun_decode_mark(int device, void *stream, const int32_t *symbols, int64_t n_symbols, int32_t end_of_block,
                            int32_t *is_eob_out);
int ivc_zerorun_decode_ends(int device, void *stream, const int32_t *is_eob, const int64_t *rank, int64_t n_symbols,
                            int64_t n_blocks, int64_t *ends_out);
int ivc_zerorun_decode_write(int device, void *stream, const int32_t *symbols, const int64_t *ends, int64_t n_blocks,
                             int32_t *blocks_out, int32_t *err_out);

/* ---- N3 (next row): symbol statistics for stats_marg (ivclab/entropy/entropy.py:6-29) as
 * IntraCodec.train_huffman_from_image uses it (intracodec.py:160-166) ----------------------------
 * minmax_out[0..1] = min, max of x (U8 / I32 / I64), int64.
 * histogram: counts_out[k] (uint64, n_bins entries, zeroed by the call) = what
 * np.histogram(x, bins=np.arange(lo, lo + n_bins + 1)) returns: unit bins, the last one closed on the
 * right.  `hot` names one frequent value besides 0, +-1, +-2 (the EOB marker) that is counted per warp. */
int ivc_symbol_minmax(int device, void *stream, const void *x, int dtype, int64_t n, int64_t *minmax_out);
int ivc_symbol_histogram(int device, void *stream, const void *x, int dtype, int64_t n, int64_t lo, int64_t n_bins,
                         int64_t hot, uint64_t *counts_out);

/* ---- N1 (next row): colour transforms (ivclab/signal/color.py:15-63) --------------------------
 * rgb: npixels x 3 (U8/I32/F32/F64) -> ycbcr float64, bit-identical to numpy's `image @ M.T + offset`
 * (one FMA chain per output, as BLAS evaluates it).  ycbcr2rgb: float64 in/out, clipped to [0,255]. */
int ivc_rgb2ycbcr(int device, void *stream, const void *rgb, int dtype, int64_t npixels, void *ycbcr_out);
int ivc_ycbcr2rgb(int device, void *stream, const void *ycbcr, int64_t npixels, void *rgb_out);
/* uint8 RGB (npixels x 3) -> uint8 luma plane clip(rint(rgb2ycbcr(rgb)[..., 0]), 0, 255): the plane the video codecs
 * code (videocodec.py:38: Y = rgb2ycbcr(frame)[..., 0]) in the 8-bit form the host-fed pipeline searches on, derived
 * on the device so that only the RGB frames cross PCIe.  luma_f64_out (may be NULL) receives the same plane as float64,
 * the dtype the P-frame entry points read.  ABI version 3. */
int ivc_rgb8_to_luma8(int device, void *stream, const void *rgb, int64_t npixels, void *luma_out, void *luma_f64_out);

/* K1 with rgb2ycbcr fused in front: uint8 RGB HWC images (W % 16 == 0, frames frame_stride_bytes apart)
 * -> [n_frames, H/8, W/8, 3, 64] int32, identical to ivc_intra_forward(rgb2ycbcr(rgb)); reads 3 bytes per
 * pixel instead of 24. */
int ivc_intra_forward_rgb8(int device, void *stream,
                           const void *rgb, int64_t n_frames, int64_t H, int64_t W, int64_t frame_stride_bytes,
                           const void *table, int table_dtype, int32_t *out);
/* ... with the zero-run coder's count pass folded in (see ivc_pframe_search_forward_zr): counts_out / masks_out have
 * n_frames * (H/8) * (W/8) * 3 entries. */
int ivc_intra_forward_rgb8_zr(int device, void *stream, const void *rgb, int64_t n_frames, int64_t H, int64_t W,
                              int64_t frame_stride_bytes, const void *table, int table_dtype, int32_t *out,
                              int32_t *counts_out, uint64_t *masks_out);

/* The same transform quantised with n_tables tables at once: a rate-distortion sweep codes every frame at each of its
 * scales (exercises/ch4/ex1.py:385-405: `for q in scales: IntraCodec(quantization_scale=q) ...` over the same images),
 * and rgb2ycbcr + DCT do not depend on the scale.  tables: device [n_tables,3,8,8]; out: [n_tables, n_frames, Hp, Wp, 3, 64],
 * out[t] identical to ivc_intra_forward_rgb8 with table t; counts_out / masks_out: both NULL, or [n_tables, n_frames*Hp*Wp*3]
 * as in the _zr variant.  n_tables <= IVC_MAX_FORWARD_TABLES. */
#define IVC_MAX_FORWARD_TABLES 16
int ivc_intra_forward_rgb8_multi(int device, void *stream,
                                 const void *rgb, int64_t n_frames, int64_t H, int64_t W, int64_t frame_stride_bytes,
                                 const void *tables, int table_dtype, int n_tables, int32_t *out,
                                 int32_t *counts_out, uint64_t *masks_out);

#ifdef __cplusplus
}
#endif
#endif /* IVCLAB_B200_H */
