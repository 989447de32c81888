/*
 * svtb200.h -- C ABI of libsvtvp9_b200.so: the B200 (sm_100a) implementation of SVT-VP9's
 * per-block encode hot path (motion estimation, forward/inverse transform + quantisation,
 * in-loop deblocking).
 *
 * Plain C: pointers, sizes and POD structs only.  Every entry point returns 0 on success and a
 * negative SVTB200_E* code on failure (the reference-side shim then falls back to the reference's
 * own CPU kernel for that unit of work: results are identical by construction).
 *
 * Reference interfaces replaced (all paths relative to the reference tree):
 *   S1  motion_estimate_sb()            Source/Lib/Codec/EbMotionEstimation.c:4524
 *         -> svtb200_picture_upload() + svtb200_me_picture()       (one call per picture, all SBs)
 *   S2  eb_vp9_loop_filter_frame()      Source/Lib/VPX/vp9_loopfilter.c:1521
 *         -> svtb200_deblock_picture()                             (one call per picture)
 *   S3  RTCD kernels bound in setup_rtcd_internal()  Source/Lib/VPX/vpx_dsp_rtcd.h:1123,
 *       setup_rtcd_internal_vp9() Source/Lib/VPX/vp9_rtcd.h:138, used by perform_coding_loop()
 *       Source/Lib/Codec/EbEncDecProcess.c:365 and perform_inv_trans_add() :589
 *         -> svtb200_fwd_txfm_batch / svtb200_quantize_batch / svtb200_inv_txfm_add_batch
 *            (batched twins of eb_vp9_fdct* / fht*, eb_vp9_quantize_b[_32x32], eb_vp9_idct*_add/iht*_add)
 *         -> svtb200_txq_batch (fused residual -> transform -> quantise -> [inverse + recon])
 *
 *   around the seams (same library): svtb200_picture_upload_luma (padding + 1/4, 1/16 decimation of the picture-analysis
 *       front end), svtb200_zz_sad_picture (compute_zz_sad), svtb200_block_stats_picture
 *       (compute_block_mean_compute_variance), svtb200_build_masks (eb_vp9_build_mask_frame), svtb200_pick_filter_level
 *       (eb_vp9_pick_filter_level), svtb200_pad_plane_dev (eb_vp9_generate_padding), svtb200_inter_predict_batch (inter_predictor), svtb200_intra_predict_batch (the intra predictors),
 *       svtb200_full_distortion_batch_dev (full_distortion_kernel32bit); each cites its reference lines below.
 *
 * "Host" entry points take HOST pointers (pinned or pageable) and perform the host<->device copies
 * themselves; the "_dev" twins take device pointers already resident in HBM.
 */
#ifndef SVTB200_H
#define SVTB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVTB200_OK 0
#define SVTB200_ENODEV (-1)   /* no CUDA device / CUDA runtime failure at create */
#define SVTB200_EINVAL (-2)   /* bad parameter */
#define SVTB200_ENOMEM (-3)   /* device or pinned allocation failed */
#define SVTB200_ECUDA (-4)    /* a launch or copy failed; see svtb200_last_error() */
#define SVTB200_ENOENT (-5)   /* unknown picture id */
#define SVTB200_EPARTIAL (-6) /* a host-pointer call failed AFTER it had begun to overwrite its in/out buffer (deblocking write-back):
                                 the caller's CPU fallback would not start from the original data any more */

typedef struct SvtB200Ctx SvtB200Ctx;

/* ---- context ------------------------------------------------------------------------------ */
int         svtb200_create(int device, SvtB200Ctx **out);
void        svtb200_destroy(SvtB200Ctx *ctx);
const char *svtb200_last_error(const SvtB200Ctx *ctx);
int         svtb200_sync(SvtB200Ctx *ctx);
/* number of kernel launches issued by this context since creation (bench "gpu_launches") */
uint64_t    svtb200_launch_count(const SvtB200Ctx *ctx);
/* raw cudaStream_t the context launches on (for CUDA-event timing by the caller) */
void       *svtb200_stream(SvtB200Ctx *ctx);
/* Page-lock a buffer the CALLER owns (e.g. the reference's picture buffers, allocated once per pool entry) so that copies
 * from / to it are asynchronous DMA transfers instead of staged pageable copies.  Returns SVTB200_OK also when the range is
 * already registered. */
int   svtb200_host_register(void *ptr, size_t bytes);
int   svtb200_host_unregister(void *ptr);

/* ---- pinned host memory (the "pinned async copies" of the staging path) ------------------- */
void *svtb200_host_alloc(size_t bytes);
void  svtb200_host_free(void *p);
/* device memory helpers for the *_dev entry points (tests / bench) */
void *svtb200_dev_alloc(SvtB200Ctx *ctx, size_t bytes);
void  svtb200_dev_free(SvtB200Ctx *ctx, void *p);
int   svtb200_memcpy_h2d(SvtB200Ctx *ctx, void *dst_dev, const void *src_host, size_t bytes);
int   svtb200_memcpy_d2h(SvtB200Ctx *ctx, void *dst_host, const void *src_dev, size_t bytes);

/* A context owns five in-order queues: COMPUTE, COMPUTE2 and COMPUTE3 (kernels, and the synchronous copies above), UPLOAD (picture
 * uploads and svtb200_memcpy_h2d_async) and DOWNLOAD (svtb200_memcpy_d2h_async), so that the PCIe traffic of the
 * previous / next picture overlaps the kernels of the current one -- the role the reference's picture-level pipeline
 * (EbEncHandle.c: one process per stage, several pictures in flight) plays on the CPU.  Picture uploads are ordered
 * against the kernels that read them by the library; everything else is ordered by the caller with markers:
 *   svtb200_marker(q, &t)        t = "everything queued on q so far"
 *   svtb200_stream_wait(q, t)    queue q does not start later work before t is reached
 *   svtb200_fence(from, to)      = marker(from) + stream_wait(to)
 *   svtb200_wait(t)              the calling host thread blocks until t is reached
 * Host buffers of the async copies must be page-locked (svtb200_host_alloc). */
#define SVTB200_STREAM_COMPUTE  0
#define SVTB200_STREAM_UPLOAD   1
#define SVTB200_STREAM_DOWNLOAD 2
#define SVTB200_STREAM_COMPUTE2 3 /* second compute queue: work of a DIFFERENT picture (e.g. the deblocking of picture n beside the
                                    motion estimation of picture n+1) fills the tail of the other queue's kernels */
#define SVTB200_STREAM_COMPUTE3 4 /* third compute queue, e.g. the motion estimation of every other picture */
/* kernels of the following *_dev / host-pointer calls go to compute queue 'which' (COMPUTE, COMPUTE2 or COMPUTE3); not thread safe */
int svtb200_select_compute(SvtB200Ctx *ctx, int which);
int svtb200_memcpy_h2d_async(SvtB200Ctx *ctx, void *dst_dev, const void *src_host, size_t bytes);
/* rows x width_bytes out of a pitched host array into a pitched device array (a chroma plane into the picture area of its padded
 * device plane; svtb200_pad_plane_dev then makes the padding), on the UPLOAD queue */
int svtb200_memcpy_h2d_2d_async(SvtB200Ctx *ctx, void *dst_dev, size_t dst_pitch, const void *src_host, size_t src_pitch, size_t width_bytes,
                                size_t rows);
int svtb200_memcpy_d2h_async(SvtB200Ctx *ctx, void *dst_host, const void *src_dev, size_t bytes);
int svtb200_marker(SvtB200Ctx *ctx, int stream, uint64_t *ticket);
int svtb200_stream_wait(SvtB200Ctx *ctx, int stream, uint64_t ticket);
int svtb200_fence(SvtB200Ctx *ctx, int from_stream, int to_stream);
int svtb200_wait(SvtB200Ctx *ctx, uint64_t ticket);

/* =============================================================================================
 * S1  Motion estimation
 * ============================================================================================= */

/* One 8-bit plane with its padding, as in EbPictureBufferDesc (Source/Lib/Codec/EbPictureBufferDesc.h:27):
 * buf points at the first byte of the PADDED buffer; the picture's (0,0) sample is at
 * buf[origin_y * stride + origin_x]. */
typedef struct SvtB200Plane {
    const uint8_t *buf;
    int32_t        stride;
    int32_t        origin_x, origin_y; /* padding on the left / top */
    int32_t        width, height;      /* without padding */
    int32_t        padded_height;      /* rows in buf (height + top + bottom padding) */
} SvtB200Plane;

/* The three planes of an EbPaReferenceObject (Source/Lib/Codec/EbReferenceObject.h:37): padded
 * source luma, its 1/4 (2:1 x 2:1) and 1/16 (4:1 x 4:1) point decimations.  quarter may have
 * buf == NULL when HME level 1 is disabled. */
int svtb200_picture_upload(SvtB200Ctx *ctx, uint64_t pic_id, const SvtB200Plane *full,
                           const SvtB200Plane *quarter, const SvtB200Plane *sixteenth);
/* Upload only the full-resolution padded luma; the 1/4 and 1/16 decimations (+ their padding)
 * are produced on the device exactly as eb_vp9_decimation_2d + eb_vp9_generate_padding would
 * (Source/Lib/Codec/EbPictureAnalysisProcess.c:102, :5025). quarter_pad/sixteenth_pad are the
 * origin_x(=origin_y) of the decimated planes (reference: 32+... see DESIGN.md).
 * Only the picture area of 'full' is read from the host.  The device copy of the full-resolution plane has its own geometry
 * (svtb200_picture_plane_dev reports it): pitch a multiple of 16 and the picture 80 columns into the row, so that the picture's
 * 16-sample groups are 16-byte aligned vectors and rows are legal TMA rows; the columns beyond the host row's padding hold the
 * bytes a read past the end of the (contiguous) host row finds -- the reference's search windows do that, and the results depend
 * on it. */
int svtb200_picture_upload_luma(SvtB200Ctx *ctx, uint64_t pic_id, const SvtB200Plane *full,
                                int32_t quarter_pad, int32_t sixteenth_pad, int make_quarter);
/* svtb200_picture_upload_luma from PAGEABLE host memory (the reference's picture pool, EbPaReferenceObject::input_padded_picture_ptr, as the
 * encoder allocated it): the calling thread copies the picture area into one of a small ring of page-locked slots the context owns and the
 * transfer runs from there, instead of page-locking every pool object on first sight.  Thread safe; calls for different pictures stage in
 * parallel. */
int svtb200_picture_upload_luma_staged(SvtB200Ctx *ctx, uint64_t pic_id, const SvtB200Plane *full,
                                       int32_t quarter_pad, int32_t sixteenth_pad, int make_quarter);
int svtb200_picture_release(SvtB200Ctx *ctx, uint64_t pic_id);
/* Device-side descriptor of an uploaded plane (which: 0 full, 1 quarter, 2 sixteenth): buf is a DEVICE pointer with the
 * same stride / origin as the uploaded host plane -- for chaining svtb200_txq_batch_dev / svtb200_deblock_picture_dev
 * on planes that are already resident. */
int svtb200_picture_plane_dev(SvtB200Ctx *ctx, uint64_t pic_id, int which, SvtB200Plane *out);
/* eb_vp9_generate_padding() (Source/Lib/Codec/EbMcp.c:17; pad_ref_and_set_flags, EbEncDecProcess.c:4822) on a device-resident
 * plane: buf points at the first byte of the PADDED plane (pitch 'stride'), the width x height picture area starts at
 * (pad_x, pad_y); every sample of the (width + 2 pad_x) x (height + 2 pad_y) rectangle outside the picture area becomes a
 * copy of the nearest picture sample.  stride, buf and width + 2 pad_x must be multiples of 4.  Runs on the current compute queue. */
int svtb200_pad_plane_dev(SvtB200Ctx *ctx, uint8_t *buf_dev, int32_t stride, int32_t width, int32_t height, int32_t pad_x, int32_t pad_y);
/* Tell the library that kernels the CALLER queued on the compute queue (through the *_dev entry points, on planes obtained
 * from svtb200_picture_plane_dev) read picture pic_id: a later re-upload of that id waits for them. */
int svtb200_picture_touch(SvtB200Ctx *ctx, uint64_t pic_id);

/* Everything motion_estimate_sb() reads from MeContext / PictureParentControlSet /
 * SequenceControlSet (EbMotionEstimationContext.h:275, EbPictureControlSet.h:229). */
typedef struct SvtB200MeParams {
    /* picture */
    int32_t luma_width, luma_height;        /* SequenceControlSet::luma_width/height */
    int32_t num_lists;                      /* 1: P picture (list 0), 2: B picture (lists 0,1) */
    int32_t temporal_layer_index;
    int32_t hierarchical_levels;
    uint64_t ref_poc[2];                    /* pcs->ref_pic_poc_array */
    /* HME */
    int32_t enable_hme_flag, enable_hme_level0, enable_hme_level1, enable_hme_level2;
    int32_t single_hme_quadrant;
    int32_t number_hme_search_region_in_width, number_hme_search_region_in_height;
    int32_t hme_level0_total_search_area_width, hme_level0_total_search_area_height;
    int32_t hme_level0_search_area_in_width[2], hme_level0_search_area_in_height[2];
    int32_t hme_level1_search_area_in_width[2], hme_level1_search_area_in_height[2];
    int32_t hme_level2_search_area_in_width[2], hme_level2_search_area_in_height[2];
    int32_t hme_l0_multiplier_x, hme_l0_multiplier_y; /* hme_level_0_search_area_multiplier_x/y[hier][tl], percent */
    /* full-pel */
    int32_t search_area_width, search_area_height;
    /* sub-pel */
    int32_t fractional_search_method; /* 0 SUB_SAD_SEARCH, 1 FULL_SAD_SEARCH, 2 SSD_SEARCH (EbDefinitions.h:664-666) */
    int32_t fractional_search_model;  /* 0 all on, 1 su_pel_enable(), 2 all off */
    int32_t fractional_search64x64;
    int32_t cu8x8_mode, cu16x16_mode; /* pcs->cu8x8_mode / cu16x16_mode */
    int32_t rate_control_mode;        /* != 0 -> rcme_distortion is produced */
} SvtB200MeParams;

/* Same 40-byte layout as MeCuResults (Source/Lib/Codec/EbMotionEstimationLcuResults.h:36-56). */
typedef struct SvtB200MeCuResult {
    int16_t x_mv_l0, y_mv_l0, x_mv_l1, y_mv_l1; /* quarter-pel units */
    struct {
        uint32_t distortion;
        uint32_t direction; /* 0 uni L0, 1 uni L1, 2 bi */
    } cand[3];
    uint8_t total_me_candidate_index;
    uint8_t pad_[7];
} SvtB200MeCuResult;

#define SVTB200_ME_PUS_PER_SB 85

/* Motion-estimate every 64x64 super-block of picture cur_id against ref_id[0..num_lists-1].
 * out_results: HOST buffer of n_sb * 85 entries (raster SB order, reference PU order);
 * out_rcme_distortion: HOST buffer of n_sb uint32 or NULL.  Blocks until the results are in
 * host memory. */
int svtb200_me_picture(SvtB200Ctx *ctx, const SvtB200MeParams *params, uint64_t cur_id,
                       const uint64_t ref_id[2], SvtB200MeCuResult *out_results,
                       uint32_t *out_rcme_distortion);
/* Device-resident variant: results stay in HBM (out_results_dev from svtb200_dev_alloc);
 * asynchronous on the context stream. */
int svtb200_me_picture_dev(SvtB200Ctx *ctx, const SvtB200MeParams *params, uint64_t cur_id,
                           const uint64_t ref_id[2], SvtB200MeCuResult *out_results_dev,
                           uint32_t *out_rcme_distortion_dev);

/* =============================================================================================
 * S3  Transform / quantisation (batched)
 * ============================================================================================= */

enum { SVTB200_TX_4X4 = 0, SVTB200_TX_8X8 = 1, SVTB200_TX_16X16 = 2, SVTB200_TX_32X32 = 3 };
enum { SVTB200_DCT_DCT = 0, SVTB200_ADST_DCT = 1, SVTB200_DCT_ADST = 2, SVTB200_ADST_ADST = 3 };

/* compute_block_mean_compute_variance() (Source/Lib/Codec/EbPictureAnalysisProcess.c:2115-3356, per SB from
 * compute_picture_spatial_statistics :4784): for every 64x64 super-block the means and variances of its 85 square blocks in the
 * reference's ME_TIER_ZERO_PU order (64x64, 4 32x32, 16 16x16, 64 8x8; raster inside a size).  8x8 level: rows 0,2,4,6 only,
 * mean = sum << 3 (8.8 fixed point), mean of squares = sum of squares << 11 (16.16); each larger size = (sum of its four
 * quadrants) >> 2; y_mean = mean >> 8, variance = (uint16)((mean of squares - mean * mean) >> 16).  Incomplete super-blocks read
 * the plane's padding exactly as the reference does.  y_mean: n_sb x 85 bytes, variance: n_sb x 85 uint16. */
int svtb200_block_stats_picture(SvtB200Ctx *ctx, uint64_t pic_id, int32_t luma_width, int32_t luma_height, uint8_t *y_mean_host,
                                uint16_t *variance_host);
int svtb200_block_stats_picture_dev(SvtB200Ctx *ctx, uint64_t pic_id, int32_t luma_width, int32_t luma_height, uint8_t *y_mean_dev,
                                    uint16_t *variance_dev);

/* compute_chroma_block_mean() / zero_out_chroma_block_mean() (Source/Lib/Codec/EbPictureAnalysisProcess.c:1828, :1760; per SB from
 * compute_picture_spatial_statistics :4810-4822): for every COMPLETE 64x64 super-block the means of its 16 chroma 8x8 blocks (the
 * luma 16x16 positions; rows 0, 2, 4, 6 of the block only, 8.8 fixed point), of the four 16x16 and of the 32x32 chroma block, each
 * larger size = (sum of four) >> 2 -- the 32x32 one as the reference computes it, i.e. with the last quadrant added twice and the
 * third never (:2011-2016) --, stored >> 8 in ME_TIER_ZERO_PU order (64x64, 32x32 x 4, 16x16 x 16: 21 bytes per super-block and
 * plane); incomplete super-blocks get zeros.  cb / cr: DEVICE pointers to sample (0,0) of the chroma planes (8-byte aligned, stride a
 * multiple of 8); outputs n_sb x 21 bytes each, device memory.  Runs on the current compute queue. */
int svtb200_chroma_block_mean_dev(SvtB200Ctx *ctx, const uint8_t *cb_dev, const uint8_t *cr_dev, int32_t chroma_stride, int32_t luma_width,
                                  int32_t luma_height, uint8_t *cb_mean_dev, uint8_t *cr_mean_dev);
/* sub_sample_chroma_generate_pixel_intensity_histogram_bins() + calculate_histogram() (EbPictureAnalysisProcess.c:4313, :127): per
 * picture region (regions_w x regions_h, sized in luma samples like the luma histograms) the 256-bin histograms of every 4th sample
 * of every 4th row of the Cb and Cr planes; bin = (1 + count) << 4.  hist: [regions_w][regions_h][2][256] uint32 (the reference's
 * picture_histogram[w][h][1 + plane][bin]), avg: [regions_w][regions_h][2] bytes = ((sum << 4) + (area >> 3)) / (area >> 2) with the
 * region's luma area, sum16[2]: the sums over the regions of (region sum << 4).  All DEVICE pointers. */
int svtb200_chroma_histogram_dev(SvtB200Ctx *ctx, const uint8_t *cb_dev, const uint8_t *cr_dev, int32_t chroma_stride, int32_t luma_width,
                                 int32_t luma_height, int32_t regions_w, int32_t regions_h, uint32_t *hist_dev, uint8_t *avg_dev, uint64_t *sum16_dev);

/* sub_sample_luma_generate_pixel_intensity_histogram_bins() + calculate_histogram() (EbPictureAnalysisProcess.c:4237, :128;
 * from gathering_picture_statistics :4893): the 256-bin intensity histograms of the picture's 1/16 plane over a regions_w x
 * regions_h grid of regions (region = floor(size / regions); the last region of a row / column takes the remainder).  Each
 * bin is (1 + count) << 4, exactly as the reference initialises and scales them.  hist: [regions_w][regions_h][256] uint32
 * (the reference's picture_histogram[w][h][0][bin]); avg: [regions_w][regions_h] bytes = (sum + area / 2) / area
 * (average_intensity_per_region[w][h][0]); sum16: the sum over regions of (region sum << 4)
 * (sum_average_intensity_total_regions_luma).  The picture must be resident with its 1/16 plane. */
int svtb200_luma_histogram_picture(SvtB200Ctx *ctx, uint64_t pic_id, int32_t regions_w, int32_t regions_h, uint32_t *hist_host,
                                   uint8_t *avg_host, uint64_t *sum16_host);

/* compute_zz_sad() (Source/Lib/Codec/EbMotionEstimationProcess.c:431; called per picture with look-ahead, :1079-1096):
 * for every COMPLETE 64x64 super-block the 16x16 SAD between the current picture's 1/16 plane and a 4:1 point decimation
 * (eb_vp9_decimation_2d, step 4) of the PREVIOUS picture's full-resolution plane at the same position, and the non-moving
 * score it maps to: 0 / 10 / 20 / 30 for sad < 512 >> s, < 1024 >> s, < 2048 >> s, else, with
 * s = non_moving_th_shift[input_resolution] = {4, 2, 0, 0} (:353).  Incomplete (edge) super-blocks: sad 0xffffffff,
 * score 30.  The score array is what the reference stores in previous_pcs->non_moving_index_array.
 * scores: n_sb bytes; sads: n_sb uint32 or NULL.  Both pictures must be resident (svtb200_picture_upload*). */
int svtb200_zz_sad_picture(SvtB200Ctx *ctx, uint64_t cur_id, uint64_t prev_id, int32_t luma_width, int32_t luma_height,
                           int32_t th_shift, uint8_t *scores_host, uint32_t *sads_host);
int svtb200_zz_sad_picture_dev(SvtB200Ctx *ctx, uint64_t cur_id, uint64_t prev_id, int32_t luma_width, int32_t luma_height,
                               int32_t th_shift, uint8_t *scores_dev, uint32_t *sads_dev);

/* ---- rest of the picture-analysis front end (noise filters / detection, edge maps) ----------------------------------------------
 * The noise filters of picture_pre_processing_operations (Source/Lib/Codec/EbPictureAnalysisProcess.c:4191): eb_vp9_noise_extract_luma_weak
 * (:1653), _luma_strong (:1450), _chroma_weak (:1575), _chroma_strong (:1498) as detect_input_picture_noise / denoise_input_picture apply
 * them to every super-block row of a plane: get_filtered_types (:1403) with the cross (/8), gaussian (/16), gaussian (/16) and box (/9)
 * kernels on every sample that is not on the plane's border; border samples are copied.  in / den: DEVICE pointers to the sample at (0,0)
 * of a w x h plane (for the chroma kinds the chroma plane's own size); noise (luma weak only, may be NULL, geometry of den) receives
 * clip(in - den, 0, 255). */
#define SVTB200_NOISE_LUMA_WEAK 0
#define SVTB200_NOISE_LUMA_STRONG 1
#define SVTB200_NOISE_CHROMA_WEAK 2
#define SVTB200_NOISE_CHROMA_STRONG 3
int svtb200_noise_filter_dev(SvtB200Ctx *ctx, int kind, const uint8_t *in_dev, int32_t in_stride, int32_t w, int32_t h, uint8_t *den_dev,
                             int32_t den_stride, uint8_t *noise_dev);
/* detect_input_picture_noise() (:3512-3575) up to its picture-level decision: the weak luma filter of the whole picture into den, and per
 * COMPLETE 64x64 super-block compute_variance64x64 (:408) of the noise samples (noise_var, raw) and of the denoised samples (den_var,
 * >> 16), the flat-noise flag (den_var < den_th && noise_var > noise_th: FLAT_MAX_VAR and NOISE_MIN_LEVEL_0 / _1 of the reference,
 * sb_flat_noise_array), and totals[0] = sum of noise_var >> 16, totals[1] = number of complete super-blocks (pic_noise_variance =
 * totals[0] / totals[1]; the class thresholds on it stay with the host).  Incomplete super-blocks: zeros.  All pointers DEVICE memory. */
int svtb200_noise_detect_dev(SvtB200Ctx *ctx, const uint8_t *in_dev, int32_t in_stride, int32_t w, int32_t h, uint8_t *den_dev, int32_t den_stride,
                             uint64_t noise_th, uint64_t den_th, uint64_t *noise_var_dev, uint64_t *den_var_dev, uint8_t *flat_dev, uint64_t *totals_dev);
/* The decimated twins of the same stage, which the reference selects by resolution / tune / enc-mode (eb_vp9_signal_derivation_pre_analysis_oq / _sq,
 * Source/Lib/Codec/EbResourceCoordinationProcess.c:290-390; tune 1 at enc-mode > 3: 1080p -> quarter, 4K -> half precision): quarter_sample_detect_noise
 * (:3809) on the 1/4 plane with block = 32 and sub_sample_detect_noise (:3930) on the 1/16 plane with block = 16 -- one block x block area is one
 * full-resolution super-block, written at (by / block) * picture_width_in_sb + bx / block.  in_dev / w / h: the decimated plane's sample (0,0) and size
 * (svtb200_picture_plane_dev, which 1 / 2).  The statistics follow the reference's routines to the letter, including where they differ from the
 * full-resolution one (whole 64x64 areas of the plane only -- at 3840x2160 the complete super-block row 32 is never visited --, 8x8 means over all
 * eight rows, compute_variance32x32's block pairing, the noise rows of a 64-row band's first block row for every block of the band): the flags
 * are pinned against sb_flat_noise_array inside real encodes (oracle/ref_seam_check.c).  den_dev (w x h, pitch den_stride) is scratch: the areas walked
 * receive the weak-filtered samples.  Unvisited super-blocks: zeros.  noise_th / den_th: NOISE_MIN_LEVEL_DECIM_0 / _1 and FLAT_MAX_VAR_DECIM. */
int svtb200_noise_detect_decimated_dev(SvtB200Ctx *ctx, const uint8_t *in_dev, int32_t in_stride, int32_t w, int32_t h, int32_t block,
                                       int32_t picture_width_in_sb, int32_t n_sb, uint8_t *den_dev, int32_t den_stride, uint64_t noise_th, uint64_t den_th,
                                       uint64_t *noise_var_dev, uint64_t *den_var_dev, uint8_t *flat_dev, uint64_t *totals_dev);
/* edge_detection() (:4545) on the block statistics svtb200_block_stats_picture_dev leaves ([n_sb][85]): per super-block edge_block_num
 * (64x64 variance above 70 % of the picture's average), isolated_high_intensity_sb (the 9x9 neighbourhoods of bright super-blocks beside
 * dark ones, with the reference's raster-order clear-then-mark semantics) and sharp_edge_sb_flag; pic_avg_variance
 * (compute_picture_spatial_statistics, :4827) is produced on the way. */
int svtb200_edge_detection_dev(SvtB200Ctx *ctx, const uint8_t *y_mean_dev, const uint16_t *variance_dev, int32_t luma_width, int32_t luma_height,
                               uint8_t *edge_block_num_dev, uint8_t *isolated_dev, uint8_t *sharp_edge_dev, uint16_t *pic_avg_variance_dev);
/* determine_homogeneous_region_in_picture() (:4660) on the same block statistics: var_of_var_32x32_based_sb_array ([n_sb][4]: the variance of
 * the sixteen 8x8 variances of each 32x32 quadrant; all ones for incomplete super-blocks), sb_homogeneous_area_array (64x64 variance of
 * variances <= detail_th = VAR_BASED_DETAIL_PRESERVATION_SELECTOR_THRSLHD), and counts[3] = {complete super-blocks, those with 64x64 variance
 * < low_var_th = SB_LOW_VAR_TH, those with variance 0}: very_low_var_pic_flag / logo_pic_flag are counts[1] * 100 / counts[0] > 60 / > 80. */
int svtb200_homogeneous_regions_dev(SvtB200Ctx *ctx, const uint16_t *variance_dev, int32_t luma_width, int32_t luma_height, uint32_t low_var_th,
                                    uint64_t detail_th, uint64_t *var_of_var32_dev, uint8_t *homogeneous_dev, uint64_t *counts_dev);
/* edge_detection_mean_luma_chroma16x16() (:4434): edge_cu of the sixteen 16x16 blocks of every super-block ([n_sb][16], raster order) from
 * the 16x16 means of luma ([n_sb][85]) and chroma ([n_sb][21], svtb200_chroma_block_mean_dev); logo_flags[sb] = SbParams::potential_logo_sb
 * (a function of the picture size, EbSequenceControlSet.c:332-400; completeness is derived here); zeros unless picture_number is a multiple
 * of 4.  max_grad_dev: one uint32 of device scratch (holds the picture's largest gradient afterwards). */
int svtb200_edge_cu_dev(SvtB200Ctx *ctx, const uint8_t *y_mean_dev, const uint8_t *cb_mean_dev, const uint8_t *cr_mean_dev, const uint8_t *logo_flags_dev,
                        int32_t luma_width, int32_t luma_height, uint64_t picture_number, uint8_t *edge_cu_dev, uint32_t *max_grad_dev);

/* The five 2-entry (DC, AC) rows quantize_b reads (Source/Lib/VPX/quantize.c:112):
 * zbin, round, quant, quant_shift, dequant -- produced by eb_vp9_init_quantizer
 * (Source/Lib/VPX/vp9_quantize.c:206) on the reference host. */
typedef struct SvtB200QuantRow {
    int16_t zbin[2], round[2], quant[2], quant_shift[2], dequant[2];
} SvtB200QuantRow;

/* Scan order tables (VP9 bitstream constants; Source/Lib/VPX/vp9_scan.c): scan[i] = raster
 * position of the i-th coefficient, iscan = inverse.  One pair per (tx_size, tx_type);
 * 32x32 uses tx_type 0 only.  Must be registered before any quantize / txq call. */
int svtb200_set_scan(SvtB200Ctx *ctx, int tx_size, int tx_type, const int16_t *scan,
                     const int16_t *iscan);

/* Batched eb_vp9_fdct{4x4,8x8,16x16,32x32} / eb_vp9_fht{4x4,8x8,16x16} /
 * eb_vpx_partial_fdct32x32 (partial32 != 0).  in: n blocks of N*N int16 residuals, contiguous
 * (stride N); out: n blocks of N*N int16 coefficients.  tx_type applies to the whole batch.
 * For partial32 only the top-left 16x16 of each output block is written; the rest is zeroed
 * (the reference caller memsets: EbEncDecProcess.c:403-406). */
int svtb200_fwd_txfm_batch(SvtB200Ctx *ctx, int tx_size, int tx_type, int partial32, int n,
                           const int16_t *in, int16_t *out);
int svtb200_fwd_txfm_batch_dev(SvtB200Ctx *ctx, int tx_size, int tx_type, int partial32, int n,
                               const int16_t *in_dev, int16_t *out_dev);

/* Batched eb_vp9_quantize_b (4..16) / eb_vp9_quantize_b_32x32: coeff -> qcoeff, dqcoeff, eob.
 * qrow_index[i] selects the quantiser row of block i (NULL -> row 0 for all). */
int svtb200_quantize_batch(SvtB200Ctx *ctx, int tx_size, int tx_type, int n, const int16_t *coeff,
                           const SvtB200QuantRow *qrows, int n_qrows, const uint16_t *qrow_index,
                           int16_t *qcoeff, int16_t *dqcoeff, uint16_t *eob);
int svtb200_quantize_batch_dev(SvtB200Ctx *ctx, int tx_size, int tx_type, int n,
                               const int16_t *coeff_dev, const SvtB200QuantRow *qrows_dev,
                               int n_qrows, const uint16_t *qrow_index_dev, int16_t *qcoeff_dev,
                               int16_t *dqcoeff_dev, uint16_t *eob_dev);

/* Batched eb_vp9_idct{4x4,8x8,16x16,32x32}_add / eb_vp9_iht*_add (Source/Lib/VPX/vp9_idct.c:111-190):
 * dest (n blocks of N*N uint8, contiguous, stride N) += inverse(dqcoeff), clipped. eob[i]
 * selects the reference's partial variant (results are identical to the full transform). */
int svtb200_inv_txfm_add_batch(SvtB200Ctx *ctx, int tx_size, int tx_type, int n,
                               const int16_t *dqcoeff, const uint16_t *eob, uint8_t *dest);
int svtb200_inv_txfm_add_batch_dev(SvtB200Ctx *ctx, int tx_size, int tx_type, int n,
                                   const int16_t *dqcoeff_dev, const uint16_t *eob_dev,
                                   uint8_t *dest_dev);

/* Fused perform_coding_loop (EbEncDecProcess.c:365-587) for n transform blocks of one size:
 * residual = src - pred; forward transform (tx_type per job); quantise; and, when do_recon,
 * recon = clip(pred + inverse(dqcoeff)).  src/pred/recon are planes addressed by per-job byte
 * offsets; coefficient outputs are contiguous per job (N*N each). */
typedef struct SvtB200TxJob {
    uint32_t src_offset;   /* byte offset of the block's top-left sample in src   */
    uint32_t pred_offset;  /* ... in pred  */
    uint32_t recon_offset; /* ... in recon (used when flags & SVTB200_TXJ_RECON) */
    uint16_t qrow;         /* index into qrows */
    uint8_t  tx_type;      /* SVTB200_DCT_DCT .. ADST_ADST (32x32: DCT only) */
    uint8_t  flags;        /* SVTB200_TXJ_* */
} SvtB200TxJob;
#define SVTB200_TXJ_RECON 1u     /* do inverse transform + reconstruction */
#define SVTB200_TXJ_PARTIAL32 2u /* 32x32 only: eb_vpx_partial_fdct32x32 (pf_md_level path) */

typedef struct SvtB200TxPlanes {
    const uint8_t *src;   int32_t src_stride;   size_t src_bytes;
    const uint8_t *pred;  int32_t pred_stride;  size_t pred_bytes;
    uint8_t       *recon; int32_t recon_stride; size_t recon_bytes;
} SvtB200TxPlanes;

int svtb200_txq_batch(SvtB200Ctx *ctx, int tx_size, int n, const SvtB200TxJob *jobs,
                      const SvtB200TxPlanes *planes, const SvtB200QuantRow *qrows, int n_qrows,
                      int16_t *coeff /* may be NULL */, int16_t *qcoeff, int16_t *dqcoeff,
                      uint16_t *eob);
int svtb200_txq_batch_dev(SvtB200Ctx *ctx, int tx_size, int n, const SvtB200TxJob *jobs_dev,
                          const SvtB200TxPlanes *planes_dev, const SvtB200QuantRow *qrows_dev,
                          int n_qrows, int16_t *coeff_dev, int16_t *qcoeff_dev,
                          int16_t *dqcoeff_dev, uint16_t *eob_dev);
/* Source / prediction samples may sit at any byte address (rows are fetched as 16-byte vectors where the address allows, as
 * aligned words otherwise).  Reconstruction rows are stored as 32-bit words / 16-byte vectors: the recon plane's base, its
 * stride and every recon_offset must be multiples of 4 (block positions are multiples of 4 samples; svtb200_txq_batch checks
 * the offsets, the _dev entry points check base and stride). */

/* Several fused batches of one picture -- different transform sizes, different planes (luma 32x32, chroma 16x16, the smaller
 * blocks of the picture's last rows ...) -- in ONE launch (up to 8).  The array lives in HOST memory, the pointers inside are
 * device pointers; one quantiser table serves all batches. */
typedef struct SvtB200TxBatch {
    int32_t              tx_size, n;
    const SvtB200TxJob  *jobs;
    SvtB200TxPlanes      planes;
    int16_t             *coeff;   /* may be NULL */
    int16_t             *qcoeff, *dqcoeff;
    uint16_t            *eob;
} SvtB200TxBatch;
int svtb200_txq_multi_dev(SvtB200Ctx *ctx, const SvtB200TxBatch *batches, int n_batches,
                          const SvtB200QuantRow *qrows_dev, int n_qrows);

/* The quantised coefficients of the batches' blocks as the entropy coder consumes them (tokenize_b, Source/Lib/VPX/vp9_tokenize.c:275:
 * qcoeff[scan[c]] for c < eob): for every block -- batch after batch, in job order -- its first eob coefficients in scan order, padded
 * with a zero to an even count, back to back in stream[]; offsets[t] = first coefficient of block t, offsets[n_total] = length of the
 * stream in coefficients (when it exceeds `capacity` the blocks that did not fit are missing).  Only jobs / qcoeff / eob / tx_size / n of
 * the batches are read.  stream (4-byte aligned) and offsets (n_total + 1 entries) may be device memory or page-locked host memory
 * (svtb200_host_alloc / svtb200_host_register: mapped into the device's address space): with host memory the kernel's stores are the
 * device-to-host transfer -- a 4K picture's dense coefficients are 24.9 MB, the prefixes a few MB. */
int svtb200_coeff_pack_dev(SvtB200Ctx *ctx, const SvtB200TxBatch *batches, int n_batches, int16_t *stream, uint32_t capacity,
                           uint32_t *offsets);

/* ---- coefficient rate estimation ---------------------------------------------------------------------
 * coeff_rate_estimate() (Source/Lib/Codec/EbRateDistortionCost.c:55; the libvpx cost_coeffs, Source/Lib/VPX/vp9_rdopt.c:334) with
 * use_fast_coef_costing == 0, as perform_dist_rate_calc calls it after every transform (Source/Lib/Codec/EbEncDecProcess.c:736):
 * the cost, in the 1/256-bit units of the tables, of the tokens of one block's quantised coefficients -- batched over the
 * qcoeff / eob arrays svtb200_txq_batch_dev leaves on the device.  Control values arrive as data:
 *   svtb200_set_scan_neighbors  scan_order::neighbors of every (tx_size, tx_type) (Source/Lib/VPX/vp9_scan.c; MAX_NEIGHBORS * (N*N + 1) entries)
 *   svtb200_set_rate_tables     eb_vp9_dct_cat_lt_10_value_tokens[v].token and eb_vp9_dct_cat_lt_10_value_cost[v] for v = -66 .. 66
 *                               (index v + 66), eb_vp9_cat6_low_cost[256], eb_vp9_cat6_high_cost[64] (Source/Lib/VPX/vp9_tokenize.c)
 *   token_costs                 MACROBLOCK::token_costs[tx_size] of the picture (Source/Lib/VPX/vp9_block.h:141, filled by
 *                               fill_token_costs, vp9_rd.c:93): [plane type 2][ref 2][band 6][2][ctx 6][token 12] unsigned
 * Per block: the plane type (0 luma, 1 chroma), is_inter_block(mi), the transform type (scan selection, 32x32: DCT only) and the
 * entropy context of the block = combine_entropy_contexts(left, above) (EbEncDecProcess.c:734). */
typedef struct SvtB200RateBlock {
    uint8_t plane_type, is_inter, tx_type, coeff_ctx;
} SvtB200RateBlock;
int svtb200_set_scan_neighbors(SvtB200Ctx *ctx, int tx_size, int tx_type, const int16_t *neighbors);
int svtb200_set_rate_tables(SvtB200Ctx *ctx, const int16_t value_token[133], const int32_t value_cost[133],
                            const int16_t cat6_low_cost[256], const uint16_t cat6_high_cost[64]);
int svtb200_coeff_rate_batch(SvtB200Ctx *ctx, int tx_size, int n, const int16_t *qcoeff, const uint16_t *eob,
                             const SvtB200RateBlock *blocks, const uint32_t *token_costs, int32_t *rate);
int svtb200_coeff_rate_batch_dev(SvtB200Ctx *ctx, int tx_size, int n, const int16_t *qcoeff_dev, const uint16_t *eob_dev,
                                 const SvtB200RateBlock *blocks_dev, const uint32_t *token_costs_dev, int32_t *rate_dev);
/* several batches (different transform sizes / planes of one picture) in ONE launch, up to 8; HOST array of device pointers */
typedef struct SvtB200RateBatch {
    int32_t                 tx_size, n;
    const int16_t          *qcoeff;
    const uint16_t         *eob;
    const SvtB200RateBlock *blocks;
    const uint32_t         *token_costs; /* of this transform size */
    int32_t                *rate;
} SvtB200RateBatch;
int svtb200_coeff_rate_multi_dev(SvtB200Ctx *ctx, const SvtB200RateBatch *batches, int n_batches);

/* =============================================================================================
 * S2  Deblocking loop filter
 * ============================================================================================= */

/* Same layout as LOOP_FILTER_MASK (Source/Lib/VPX/vp9_loopfilter.h:62-70), one per 64x64 SB in
 * raster order, as built by eb_vp9_build_mask_frame (vp9_loopfilter.c:1548). */
typedef struct SvtB200LfMask {
    uint64_t left_y[4];  /* per TX_SIZES */
    uint64_t above_y[4];
    uint64_t int_4x4_y;
    uint16_t left_uv[4];
    uint16_t above_uv[4];
    uint16_t int_4x4_uv;
    uint8_t  lfl_y[64];
} SvtB200LfMask;

/* loop_filter_thresh (vp9_loopfilter.h:42-46): one per filter level 0..63; only the first byte
 * of each 16-byte SIMD lane array is meaningful. */
typedef struct SvtB200LfThresh {
    uint8_t mblim, lim, hev_thr;
} SvtB200LfThresh;

typedef struct SvtB200LfParams {
    int32_t width, height;         /* luma, in samples (multiple of 8) */
    int32_t mi_rows, mi_cols;      /* cm->mi_rows / mi_cols (8x8 units) */
    int32_t y_only;                /* filter luma only */
    int32_t filter_level;          /* lf->filter_level; 0 -> no-op */
} SvtB200LfParams;

/* eb_vp9_pick_filter_level(cpi, LPF_PICK_FROM_Q) (Source/Lib/VPX/vp9_picklpf.c:37-85, 8-bit): the frame's filter level from
 * ac_q = eb_vp9_ac_quant(base_qindex, 0, VPX_BITS_8): clamp(ROUND_POWER_OF_TWO(ac_q * 20723 + 1015158, 18) - (key frame ? 4 : 0),
 * 0, 63); sharpness stays 0.  A scalar per picture: plain host arithmetic, no device work. */
int svtb200_pick_filter_level(int32_t ac_q, int32_t is_key_frame);

/* What eb_vp9_setup_mask() (Source/Lib/VPX/vp9_loopfilter.c:901) reads of one ModeInfo (vp9_blockd.h:70-95), per 8x8
 * mode-info unit: the block size enum (BLOCK_4X4 = 0 ... BLOCK_64X64 = 12), the luma transform size (TX_4X4 = 0 ... TX_32X32 = 3),
 * mi->skip && is_inter_block(mi), and get_filter_level(&cm->lf_info, mi) (:242).  Every unit of a block carries the block's values
 * (as mi_grid_visible points every unit at its block). */
typedef struct SvtB200LfModeInfo {
    uint8_t sb_type, tx_size, skip_inter, filter_level;
} SvtB200LfModeInfo;

/* eb_vp9_build_mask_frame() (vp9_loopfilter.c:1548) for a whole picture: one SvtB200LfMask (= LOOP_FILTER_MASK) per 64x64
 * super-block at lfm[r * lfm_stride + c], BEFORE eb_vp9_adjust_mask (svtb200_deblock_picture applies that itself).
 * mi: mi_rows x mi_stride units.  The _dev variant leaves the masks on the device for svtb200_deblock_picture_dev. */
int svtb200_build_masks(SvtB200Ctx *ctx, int32_t mi_rows, int32_t mi_cols, const SvtB200LfModeInfo *mi, int32_t mi_stride,
                        SvtB200LfMask *lfm, int32_t lfm_stride);
int svtb200_build_masks_dev(SvtB200Ctx *ctx, int32_t mi_rows, int32_t mi_cols, const SvtB200LfModeInfo *mi_dev, int32_t mi_stride,
                            SvtB200LfMask *lfm_dev, int32_t lfm_stride);

/* In-place eb_vp9_loop_filter_frame on HOST planes y/u/v (pointers at picture origin).  Errors up to and including the kernel launch
 * leave the planes untouched (the caller may filter on the CPU instead); a failure while the filtered planes are being copied back
 * returns SVTB200_EPARTIAL: the planes hold a mixture and must not be filtered again.
 * The wavefront scratch of the _dev variant exists once per compute queue: pictures on DIFFERENT queues of one context may be in
 * flight together, launches on one queue are ordered by the queue. */
int svtb200_deblock_picture(SvtB200Ctx *ctx, const SvtB200LfParams *p, uint8_t *y, int32_t y_stride,
                            uint8_t *u, uint8_t *v, int32_t uv_stride, const SvtB200LfMask *lfm,
                            int32_t lfm_stride /* SBs per row */, const SvtB200LfThresh thr[64]);
int svtb200_deblock_picture_dev(SvtB200Ctx *ctx, const SvtB200LfParams *p, uint8_t *y_dev,
                                int32_t y_stride, uint8_t *u_dev, uint8_t *v_dev, int32_t uv_stride,
                                const SvtB200LfMask *lfm_dev, int32_t lfm_stride,
                                const SvtB200LfThresh *thr_dev);

/* =============================================================================================
 * Next to S3: prediction of the encode pass and the full-loop distortion (the calls either side of every transform call)
 * ============================================================================================= */

/* One inter_predictor() call (Source/Lib/VPX/vp9_reconinter.h:23) as build_inter_predictors() makes it
 * (Source/Lib/VPX/vp9_reconinter.c:102-250): the regular 8-tap kernel eb_vp9_filter_kernels[0] (:105, switchable filters are
 * off in the reference), unscaled (xs = ys = 16), dispatched on (subpel_x != 0, subpel_y != 0) like scale_factors::predict
 * (Source/Lib/VPX/vp9_scale.c:78-128): copy / eb_vp9_convolve8_horiz / _vert / eb_vp9_convolve8 (Source/Lib/VPX/vpx_convolve.c:101-180);
 * a compound block (two references) is the second predictor averaged into the first (eb_vp9_convolve8_avg*, :170; eb_vp9_convolve_avg_c :200).
 * All pointers are DEVICE pointers.  ref[i] points at sample (0,0) of the reference plane inside its padding; (x[i], y[i]) is
 * the integer position of the block's top-left sample in that plane = block origin + (mv >> 4), which may lie in the padding
 * (the 8-tap filter reads 3 samples before and 4 after, exactly the samples the reference functions read). */
typedef struct SvtB200PredBlock {
    const uint8_t *ref[2];
    uint8_t       *dst;
    int32_t ref_stride[2];
    int32_t x[2], y[2];
    int32_t dst_stride;
    uint8_t subpel_x[2], subpel_y[2]; /* mv & 15, in 1/16 sample */
    uint8_t w, h;                     /* 4 .. 64 (the reference's SIMD kernels take no narrower blocks; its encoder asks for none) */
    uint8_t compound;                 /* 0: ref[0] only; 1: ref[1] averaged in */
    uint8_t reserved;
} SvtB200PredBlock;
/* blocks: HOST array of n descriptors (copied to the device by the call); destinations of one batch must not overlap. */
int svtb200_inter_predict_batch(SvtB200Ctx *ctx, const SvtB200PredBlock *blocks, int n);
/* the same with the descriptors already on the device (unchecked) */
int svtb200_inter_predict_batch_dev(SvtB200Ctx *ctx, const SvtB200PredBlock *blocks_dev, int n);

/* The block walk of the encode pass's inter prediction on the device: what inter_prediction() / build_inter_predictors() read of a
 * block (Source/Lib/Codec/EbIntraPrediction.c:49-73, Source/Lib/VPX/vp9_reconinter.c:102-250) -> the three inter_predictor() calls
 * (one per plane, 4:2:0) as SvtB200PredBlock descriptors for svtb200_inter_predict_batch_dev:
 *   - the motion vector of each reference scaled to 1/16 sample of the plane and clamped so that the block reaches at most
 *     VP9_INTERP_EXTEND + block size samples out of the picture (eb_vp9_clamp_mv_to_umv_border_sb, vp9_reconinter.c:68-87, with
 *     mb_to_*_edge as the encode pass sets them, EbEncDecProcess.c:2052-2066);
 *   - use_subpel_flag == 0: the vector rounded to whole luma samples (vp9_reconinter.c:168-180);
 *   - reference plane of the vector's list, chroma origin (block_origin & ~7) >> 1 (ROUND_UV, vp9_reconinter.c:146-160).
 * Blocks of 8x8 luma samples and larger (sb_type >= BLOCK_8X8: one vector per reference for the whole block). */
typedef struct SvtB200InterBlock {
    uint16_t x, y;       /* block_origin_x / y: luma position of the block in the picture */
    uint8_t  bw, bh;     /* luma size: 8, 16, 32 or 64 each */
    uint8_t  list[2];    /* reference list (0 / 1) of mv[0] and mv[1]; list[1] == 0xff: one reference only */
    int16_t  mv[2][2];   /* ModeInfo::mv[ref].as_mv {row, col}: 1/8 luma sample */
} SvtB200InterBlock;
typedef struct SvtB200PredPicture {
    int32_t        width, height;      /* luma size of the picture (mi_cols / mi_rows = size rounded up to 8, as vp9_set_mi_size) */
    const uint8_t *ref[2][3];          /* [list][plane]: DEVICE pointer to sample (0,0) inside the padding */
    int32_t        ref_stride[2][3];
    uint8_t       *dst[3];             /* prediction planes (position (0,0)) */
    int32_t        dst_stride[3];
    int32_t        use_subpel_flag;
} SvtB200PredPicture;
/* blocks_dev: n blocks; out_dev: 3 * n descriptors (block i: out[3 * i + plane]); pic: HOST struct holding device pointers. */
int svtb200_inter_blocks_expand_dev(SvtB200Ctx *ctx, const SvtB200InterBlock *blocks_dev, int n, const SvtB200PredPicture *pic,
                                    SvtB200PredBlock *out_dev);
/* Open-loop stand-in for mode decision (bench and tests; in the encoder ModeInfo comes from the host's mode decision): the picture
 * tiled with 32x32 blocks (16x16 where only 16 rows / columns remain; width and height multiples of 16), every block inter coded
 * with the best candidate motion estimation found for that partition (cand[0].direction: list 0, list 1 or both; vectors
 * me << 1 like EbModeDecision.c:222-227).  me_dev: svtb200_me_picture_dev's output.  Blocks are numbered: the 32x32 blocks in
 * raster order, then the 16x16 blocks of the bottom strip, then those of the right strip; returns their number (or < 0). */
int svtb200_me_best_blocks_dev(SvtB200Ctx *ctx, const SvtB200MeCuResult *me_dev, int32_t width, int32_t height, SvtB200InterBlock *blocks_dev);

/* The inter chain of one picture in ONE call (one lock, seven launches on the current compute queue, no host round trip between the
 * stages): [open-loop partition from the motion-estimation results ->] block walk -> inter prediction of the three planes ->
 * fused transform batches (prediction = the planes just written) -> [scan-order packing of the coefficients] -> [rate of every block].
 * The same work as the single entry points in this order: svtb200_me_best_blocks_dev, svtb200_inter_blocks_expand_dev,
 * svtb200_inter_predict_batch_dev, svtb200_txq_multi_dev, svtb200_coeff_pack_dev, svtb200_coeff_rate_multi_dev. */
typedef struct SvtB200InterChain {
    int32_t                  after_queue;   /* SVTB200_STREAM_*: start behind everything queued there so far (the picture's motion estimation); -1: none */
    int32_t                  on_queue;      /* SVTB200_STREAM_COMPUTE / _COMPUTE2 / _COMPUTE3: run the chain there (the selected queue is left as it was); -1: the selected queue */
    const SvtB200MeCuResult *me;            /* non-NULL: blocks := svtb200_me_best_blocks_dev(me) (n_blocks must be the tiling's count) */
    SvtB200InterBlock       *blocks;        /* DEVICE, n_blocks */
    int32_t                  n_blocks;
    SvtB200PredPicture       pic;
    SvtB200PredBlock        *desc;          /* DEVICE scratch: 3 * n_blocks descriptors */
    const SvtB200TxBatch    *tx;            /* HOST array of n_tx batches (device pointers inside) */
    int32_t                  n_tx;
    const SvtB200QuantRow   *qrows;         /* DEVICE */
    int32_t                  n_qrows;
    int16_t                 *stream;        /* NULL: no packing */
    uint32_t                 capacity;
    uint32_t                *offsets;
    const SvtB200RateBatch  *rate;          /* HOST array of n_rate batches; n_rate 0: no rate */
    int32_t                  n_rate;
    uint64_t                 reads[3];      /* picture ids (svtb200_picture_upload*) whose planes the chain reads: svtb200_picture_touch */
    int32_t                  n_reads;
} SvtB200InterChain;
int svtb200_inter_chain_dev(SvtB200Ctx *ctx, const SvtB200InterChain *chain);

/* One intra prediction call of the encode pass: pred[mode][tx_size](dst, stride, above, left), or for DC_PRED
 * dc_pred[have_left][have_top][tx_size] (build_intra_predictors, Source/Lib/VPX/vp9_reconintra.c:397-407; the predictors:
 * Source/Lib/VPX/intrapred.c:22-218 and the 4x4 forms :253-417), on the reference sample arrays the reference's host code
 * prepares (generate_intra_reference_samples: const_above_row / left_col of EncDecContext).  Device pointers: above[-1] ..
 * above[2 * bs - 1] and left[0 .. bs - 1] must be readable (bs = 4 << tx_size).  mode is a PREDICTION_MODE value:
 * 0 DC, 1 V, 2 H, 3 D45, 4 D135, 5 D117, 6 D153, 7 D207, 8 D63, 9 TM. */
typedef struct SvtB200IntraBlock {
    const uint8_t *above;
    const uint8_t *left;
    uint8_t       *dst;
    int32_t        dst_stride;
    uint8_t        mode, tx_size;
    uint8_t        have_left, have_top; /* DC_PRED only */
} SvtB200IntraBlock;
/* blocks: HOST array (copied to the device by the call) */
int svtb200_intra_predict_batch(SvtB200Ctx *ctx, const SvtB200IntraBlock *blocks, int n);

/* generate_intra_reference_samples() (Source/Lib/Codec/EbEncDecProcess.c:1128-1315) for n transform blocks of one plane, on the device:
 * the above row (above[-1] .. above[2 bs - 1]) and the left column (bs samples) the predictors read, taken from the plane the blocks'
 * neighbours live in (the reconstruction, or the source for the open-loop intra search: the reference's neighbour arrays hold exactly
 * these samples) with the reference's availability and frame-edge rules: 127 / 129 fills without a top / left neighbour, the above-right
 * half replicated from above[bs - 1] except for the left 4x4 of an 8x8 block (have_right), rows / columns past the frame edge replicated
 * from the last sample inside (extend_right / extend_bottom: the prediction block crosses the frame edge, xd->mb_to_right_edge < 0 /
 * mb_to_bottom_edge < 0).  A job is filled from the block the way the reference derives its locals (:1131-1172):
 *   x0, y0        position of the transform block in the plane: (-mb_to_left_edge >> (3 + ss_x)) + 4 * (bmi_index & 1), ... + 4 * (bmi_index >> 1)
 *   have_top/left block origin (chroma: ROUND_UV(origin) >> 1) > 0;   have_right  (bmi_index & 1) + (1 << tx_size) < plane block width in 4-sample units
 * edges_dev: 128 bytes per job: [15] = above[-1], [16 ..) = above, [80 ..) = left -- SvtB200IntraBlock::above = edges + 128 j + 16,
 * ::left = edges + 128 j + 80 feed svtb200_intra_predict_batch without a host round trip.  plane_dev: the sample at (0,0). */
typedef struct SvtB200IntraEdgeJob {
    int32_t x0, y0;
    uint8_t tx_size, have_left, have_top, have_right;
    uint8_t extend_right, extend_bottom, pad[2];
} SvtB200IntraEdgeJob;
int svtb200_intra_edges_dev(SvtB200Ctx *ctx, const uint8_t *plane_dev, int32_t stride, int32_t frame_width, int32_t frame_height,
                            const SvtB200IntraEdgeJob *jobs_dev, int n, uint8_t *edges_dev);

/* full_distortion_intrinsic_func_ptr_array[..][eob != 0][0][..] (Source/Lib/Codec/EbPictureOperators.h:240, call site
 * Source/Lib/Codec/EbEncDecProcess.c:723): per transform block k of n (N x N coefficients each, block-major as svtb200_txq_batch
 * writes them) dist[2k] = sum (coeff - recon_coeff)^2, dist[2k+1] = sum coeff^2; with eob[k] == 0 both are sum coeff^2
 * (full_distortion_kernel32bit / _eob_zero32bit, Source/Lib/C_DEFAULT/EbPictureOperators_C.c:289,314; sums kept in 32 bits like there).
 * Device pointers. */
int svtb200_full_distortion_batch_dev(SvtB200Ctx *ctx, int tx_size, int n, const int16_t *coeff_dev, const int16_t *recon_coeff_dev,
                                      const uint16_t *eob_dev, uint64_t *dist_dev);

#ifdef __cplusplus
}
#endif
#endif /* SVTB200_H */
