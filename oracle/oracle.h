/* TEST INFRASTRUCTURE ONLY -- declarations of the CPU oracle ("port").  See the individual .c
 * files for the reference functions each restatement follows.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use these. */
#ifndef SVTB200_ORACLE_H
#define SVTB200_ORACLE_H
#include <stdint.h>
#include "../include/svtb200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* intermediate values of one super-block, for stage-by-stage comparison with the CUDA path */
typedef struct SvtB200OracleSbTrace {
    int16_t  hme_center_x[2], hme_center_y[2]; /* search centre after HME, before the (0,0) test */
    int16_t  sa_ox[2], sa_oy[2], sa_w[2], sa_h[2];
    uint32_t fullpel_sad[2][85], fullpel_mv[2][85]; /* internal (z-order) PU indexing */
    uint8_t  subpel_enable[2];                      /* bit0 32x32, bit1 16x16, bit2 8x8, bit3 quarter */
} SvtB200OracleSbTrace;

/* cur/ref0/ref1: arrays of 3 planes {full, quarter, sixteenth}; ref1 may be NULL for P pictures */
void svtb200_oracle_me_sb(const SvtB200MeParams *P, const SvtB200Plane cur[3], const SvtB200Plane ref0[3],
                          const SvtB200Plane ref1[3], int sb_x, int sb_y, SvtB200MeCuResult out[85],
                          uint32_t *rcme_distortion, SvtB200OracleSbTrace *trace);
void svtb200_oracle_me_picture(const SvtB200MeParams *P, const SvtB200Plane cur[3], const SvtB200Plane ref0[3],
                               const SvtB200Plane ref1[3], SvtB200MeCuResult *out, uint32_t *rcme, int sb_begin,
                               int sb_end);

void svtb200_oracle_zz_sad_picture(const SvtB200Plane *cur16, const SvtB200Plane *prev_full, int w, int h, int th_shift,
                                   uint8_t *scores, uint32_t *sads);

void svtb200_oracle_build_masks(int mi_rows, int mi_cols, const SvtB200LfModeInfo *mi, int stride, SvtB200LfMask *lfm, int lfm_stride);

void svtb200_oracle_block_stats_picture(const SvtB200Plane *full, int w, int h, uint8_t *y_mean, uint16_t *variance);

void svtb200_oracle_luma_histogram(const SvtB200Plane *p16, int rw, int rh, uint32_t *hist, uint8_t *avg, uint64_t *sum16);

/* prediction of the encode pass (host pointers in the descriptors) and full-loop distortion: pred_port.c */
void svtb200_oracle_inter_predict(const SvtB200PredBlock *blocks, int n);
void svtb200_oracle_intra_predict(const SvtB200IntraBlock *blocks, int n); /* host pointers */
void svtb200_oracle_intra_edges(const uint8_t *plane, int stride, int frame_width, int frame_height, const SvtB200IntraEdgeJob *jobs, int n, uint8_t *out);
void svtb200_oracle_full_distortion(int tx_size, int n, const int16_t *coeff, const int16_t *recon_coeff, const uint16_t *eob, uint64_t *dist);

/* coefficient rate estimation (coeff_rate_estimate, use_fast_coef_costing == 0): rate_port.c */
void svtb200_oracle_coeff_rate(int tx_size, int n, const int16_t *qcoeff, const uint16_t *eob, const SvtB200RateBlock *blocks,
                               const uint32_t *token_costs, const int16_t *const scans[4], const int16_t *const neighbors[4],
                               const int16_t *vtok, const int32_t *vcost, const int16_t *c6lo, const uint16_t *c6hi, int32_t *rate);

/* chroma statistics of the picture-analysis stage: pa_port.c (cb / cr at sample (0,0) of the planes) */
void svtb200_oracle_chroma_block_mean(const uint8_t *cb, const uint8_t *cr, int stride, int luma_width, int luma_height, uint8_t *cb_mean, uint8_t *cr_mean);
void svtb200_oracle_chroma_histogram(const uint8_t *cb, const uint8_t *cr, int stride, int luma_width, int luma_height, int rw, int rh,
                                     uint32_t *hist, uint8_t *avg, uint64_t sum16[2]);

/* rest of the picture-analysis front end: pa_noise_port.c */
void svtb200_oracle_noise_filter(int kind, const uint8_t *in, int in_stride, int w, int h, uint8_t *den, int den_stride, uint8_t *noise);
void svtb200_oracle_noise_detect(const uint8_t *in, int in_stride, int w, int h, uint8_t *den, int den_stride, uint64_t noise_th, uint64_t den_th,
                                 uint64_t *noise_var, uint64_t *den_var, uint8_t *flat, uint64_t totals[2]);
void svtb200_oracle_noise_detect_decimated(const uint8_t *in, int in_stride, int w, int h, int block, int picture_width_in_sb, int n_sb, uint8_t *den, int den_stride,
                                           uint64_t noise_th, uint64_t den_th, uint64_t *noise_var, uint64_t *den_var, uint8_t *flat, uint64_t totals[2]);
void svtb200_oracle_edge_detection(const uint8_t *y_mean, const uint16_t *variance, int w, int h, uint64_t pic_avg_variance, uint8_t *edge_block_num,
                                   uint8_t *isolated, uint8_t *sharp_edge);
void svtb200_oracle_edge_cu(const uint8_t *y_mean, const uint8_t *cb_mean, const uint8_t *cr_mean, const uint8_t *sb_flags, int n_sb, uint64_t picture_number,
                            uint8_t *edge_cu);

void svtb200_oracle_homogeneous_regions(const uint16_t *variance, int w, int h, uint32_t low_var_th, uint64_t detail_th, uint64_t *var_of_var32, uint8_t *homogeneous,
                                        uint64_t counts[3]);

/* block walk in front of the inter predictors, open-loop partition of the bench: blocks_port.c */
void svtb200_oracle_inter_expand(const SvtB200InterBlock *blocks, int n, const SvtB200PredPicture *pic, SvtB200PredBlock *out);
int  svtb200_oracle_me_best_blocks(const SvtB200MeCuResult *me, int width, int height, SvtB200InterBlock *out);

#ifdef __cplusplus
}
#endif
#endif
