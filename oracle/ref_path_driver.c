/*
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY.  Drives the UNMODIFIED reference's own hot-path functions from plain
 * data, so that (a) bench.py's reference arm times the reference's real implementation (C or AVX2 dispatch, as the
 * encoder itself would select) on the host cores, and (b) the tests can compare the CUDA path with the reference
 * on arbitrary inputs, not only on dumped encodes.  Compiled by oracle/build_ref.sh against the reference's headers
 * and linked with the reference objects into oracle/_ref/libSvtVp9Enc_refdrv.so.  Nothing here is product code.
 *
 *   refdrv_init(asm)            eb_vp9_ASM_TYPES + setup_rtcd_internal / setup_rtcd_internal_vp9 (EbEncHandle.c:800-807)
 *   refdrv_me_range(...)        motion_estimate_sb (EbMotionEstimation.c:4524) for a range of SBs, with the per-SB
 *                               buffer set-up of eb_vp9_motion_estimation_kernel (EbMotionEstimationProcess.c:978-1035)
 *   refdrv_txq(...)             the kernel sequence of perform_coding_loop (EbEncDecProcess.c:365-587)
 *   refdrv_deblock(...)         eb_vp9_loop_filter_frame (vp9_loopfilter.c:1521)
 *   refdrv_pool_run(...)        a mini-GOP of pictures through the three calls above on a native, pinned thread pool
 *                               (the CPU arm of bench.py: one call per step, no interpreter in the timed region)
 */
#define _GNU_SOURCE
#define RTCD_C /* the two headers then provide the static setup_rtcd_internal*; built with -fcommon so that the
                  function-pointer variables merge with the definitions in the reference objects */
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "svtb200_ref_glue.h"
#include "vpx_dsp_rtcd.h"
#include "vp9_rtcd.h"
#include "vp9_idct.h"
#include "vp9_scan.h"
#include "EbMotionEstimation.h"
#include "EbPictureOperators.h"

uint32_t get_cpu_asm_type(void);
EbErrorType eb_vp9_me_context_ctor(MeContext **object_dbl_ptr);

static int g_init;

int refdrv_init(int use_simd) {
    if (g_init) return 0;
    /* the constructors register their allocations in the encoder's global memory map */
    static uint32_t idx;
    static uint64_t total;
    memory_map       = (EbMemoryMapEntry *)calloc(MAX_NUM_PTR, sizeof(EbMemoryMapEntry));
    memory_map_index = &idx;
    total_lib_memory = &total;
    eb_vp9_ASM_TYPES = use_simd ? get_cpu_asm_type() : 0;
    setup_rtcd_internal(eb_vp9_ASM_TYPES);
    setup_rtcd_internal_vp9(eb_vp9_ASM_TYPES);
    g_init = 1;
    return (int)eb_vp9_ASM_TYPES;
}

/* ------------------------------------------------------------------------------------------------ ME */
typedef struct {
    MeContext               *me;
    PictureParentControlSet *pcs;
    SequenceControlSet      *scs;
    EbObjectWrapper          scs_w, ref_w[2];
    EbPaReferenceObject      pa_cur, pa_ref[2];
    EbPictureBufferDesc      d_cur[3], d_ref[2][3];
    MeCuResults             *res; /* one SB */
    MeCuResults             *res_ptr[1];
} RefMe;

void *refdrv_me_new(void) {
    RefMe *r = (RefMe *)calloc(1, sizeof(RefMe));
    if (eb_vp9_me_context_ctor(&r->me) != EB_ErrorNone) return NULL;
    r->pcs = (PictureParentControlSet *)calloc(1, sizeof(PictureParentControlSet));
    r->scs = (SequenceControlSet *)calloc(1, sizeof(SequenceControlSet));
    r->res = (MeCuResults *)calloc(85, sizeof(MeCuResults));
    return r;
}

static void desc_of(const SvtB200Plane *p, EbPictureBufferDesc *d) {
    memset(d, 0, sizeof(*d));
    d->buffer_y = (EbByte)p->buf; d->stride_y = (uint16_t)p->stride;
    d->origin_x = (uint16_t)p->origin_x; d->origin_y = (uint16_t)p->origin_y;
    d->width = (uint16_t)p->width; d->height = (uint16_t)p->height;
    d->luma_size = (uint32_t)p->stride * (uint32_t)p->padded_height;
}

int refdrv_me_range(void *h, const SvtB200MeParams *P, const SvtB200Plane cur[3], const SvtB200Plane ref0[3],
                    const SvtB200Plane ref1[3], int sb_begin, int sb_end, SvtB200MeCuResult *out) {
    RefMe *r = (RefMe *)h;
    MeContext *me = r->me;
    PictureParentControlSet *pcs = r->pcs;
    for (int i = 0; i < 3; i++) { desc_of(&cur[i], &r->d_cur[i]); desc_of(&ref0[i], &r->d_ref[0][i]); desc_of(ref1 ? &ref1[i] : &ref0[i], &r->d_ref[1][i]); }
    for (int l = 0; l < 2; l++) {
        r->pa_ref[l].input_padded_picture_ptr = &r->d_ref[l][0];
        r->pa_ref[l].quarter_decimated_picture_ptr = &r->d_ref[l][1];
        r->pa_ref[l].sixteenth_decimated_picture_ptr = &r->d_ref[l][2];
        r->ref_w[l].object_ptr = &r->pa_ref[l];
        pcs->ref_pa_pic_ptr_array[l] = &r->ref_w[l];
        pcs->ref_pic_poc_array[l] = P->ref_poc[l];
    }
    r->scs->luma_width = (uint16_t)P->luma_width; r->scs->luma_height = (uint16_t)P->luma_height;
    r->scs->static_config.rate_control_mode = 0;
    r->scs_w.object_ptr = r->scs;
    pcs->sequence_control_set_wrapper_ptr = &r->scs_w;
    pcs->slice_type = P->num_lists == 1 ? P_SLICE : B_SLICE;
    pcs->temporal_layer_index = (uint8_t)P->temporal_layer_index;
    pcs->hierarchical_levels = (uint8_t)P->hierarchical_levels;
    pcs->enable_hme_flag = (EB_BOOL)P->enable_hme_flag;
    pcs->enable_hme_level_0_flag = (EB_BOOL)P->enable_hme_level0;
    pcs->enable_hme_level_1_flag = (EB_BOOL)P->enable_hme_level1;
    pcs->enable_hme_level_2_flag = (EB_BOOL)P->enable_hme_level2;
    pcs->cu8x8_mode = P->cu8x8_mode; pcs->cu16x16_mode = P->cu16x16_mode;
    pcs->max_number_of_pus_per_sb = 85;
    pcs->me_results = r->res_ptr; r->res_ptr[0] = r->res;
    me->single_hme_quadrant = (EB_BOOL)P->single_hme_quadrant;
    me->number_hme_search_region_in_width = (uint16_t)P->number_hme_search_region_in_width;
    me->number_hme_search_region_in_height = (uint16_t)P->number_hme_search_region_in_height;
    me->hme_level0_total_search_area_width = (uint16_t)P->hme_level0_total_search_area_width;
    me->hme_level0_total_search_area_height = (uint16_t)P->hme_level0_total_search_area_height;
    for (int i = 0; i < 2; i++) {
        me->hme_level0_search_area_in_width_array[i] = (uint16_t)P->hme_level0_search_area_in_width[i];
        me->hme_level0_search_area_in_height_array[i] = (uint16_t)P->hme_level0_search_area_in_height[i];
        me->hme_level1_search_area_in_width_array[i] = (uint16_t)P->hme_level1_search_area_in_width[i];
        me->hme_level1_search_area_in_height_array[i] = (uint16_t)P->hme_level1_search_area_in_height[i];
        me->hme_level2_search_area_in_width_array[i] = (uint16_t)P->hme_level2_search_area_in_width[i];
        me->hme_level2_search_area_in_height_array[i] = (uint16_t)P->hme_level2_search_area_in_height[i];
    }
    me->search_area_width = (uint8_t)P->search_area_width; me->search_area_height = (uint8_t)P->search_area_height;
    me->fractional_search_method = (uint8_t)P->fractional_search_method;
    me->fractional_search_model = (uint8_t)P->fractional_search_model;
    me->fractional_search64x64 = (EB_BOOL)P->fractional_search64x64;

    const EbPictureBufferDesc *in = &r->d_cur[0], *q = &r->d_cur[1], *s = &r->d_cur[2];
    const int sbs_w = (P->luma_width + 63) / 64;
    for (int sb = sb_begin; sb < sb_end; sb++) {
        const uint32_t ox = (uint32_t)(sb % sbs_w) * 64, oy = (uint32_t)(sb / sbs_w) * 64;
        const uint32_t sbw = (uint32_t)(P->luma_width - (int)ox < 64 ? P->luma_width - (int)ox : 64);
        const uint32_t sbh = (uint32_t)(P->luma_height - (int)oy < 64 ? P->luma_height - (int)oy : 64);
        uint32_t bi = (in->origin_y + oy) * in->stride_y + in->origin_x + ox;
        me->hme_search_type = HME_RECTANGULAR;
        for (uint32_t row = 0; row < 64; row++) memcpy(&me->sb_buffer[row * 64], &in->buffer_y[bi + row * in->stride_y], 64);
        me->sb_src_ptr = &in->buffer_y[bi]; me->sb_src_stride = in->stride_y;
        if (P->enable_hme_level1) {
            bi = (q->origin_y + (oy >> 1)) * q->stride_y + q->origin_x + (ox >> 1);
            for (uint32_t row = 0; row < (sbh >> 1); row++)
                memcpy(&me->quarter_sb_buffer[row * me->quarter_sb_buffer_stride], &q->buffer_y[bi + row * q->stride_y], sbw >> 1);
        }
        if (P->enable_hme_level0) {
            bi = (s->origin_y + (oy >> 2)) * s->stride_y + s->origin_x + (ox >> 2);
            uint8_t *fp = &s->buffer_y[bi], *lp = me->sixteenth_sb_buffer;
            for (uint32_t row = 0; row < (sbh >> 2); row += 2) { memcpy(lp, fp, sbw >> 2); lp += 16; fp += s->stride_y << 1; }
        }
        motion_estimate_sb(pcs, 0, ox, oy, me, (EbPictureBufferDesc *)in);
        for (int i = 0; i < 85; i++) {
            SvtB200MeCuResult *o = &out[(size_t)sb * 85 + i];
            const MeCuResults *m = &r->res[i];
            memset(o, 0, sizeof(*o));
            o->x_mv_l0 = m->x_mv_l0; o->y_mv_l0 = m->y_mv_l0;
            if (P->num_lists == 2) { o->x_mv_l1 = m->x_mv_l1; o->y_mv_l1 = m->y_mv_l1; }
            o->total_me_candidate_index = m->total_me_candidate_index;
            for (int k = 0; k < m->total_me_candidate_index && k < 3; k++) { o->cand[k].distortion = m->distortion_direction[k].distortion; o->cand[k].direction = m->distortion_direction[k].direction; }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------ transform / quant */
int refdrv_txq(int tx_size, int n_begin, int n_end, const SvtB200TxJob *jobs, const SvtB200TxPlanes *pl, const SvtB200QuantRow *qrows,
               int16_t *coeff, int16_t *qcoeff, int16_t *dqcoeff, uint16_t *eob) {
    const int N = 4 << tx_size, nn = N * N;
    const int asm_i = (eb_vp9_ASM_TYPES & PREAVX2_MASK) && 1;
    int16_t resid[32 * 32] __attribute__((aligned(64)));
    int16_t tmp[32 * 32] __attribute__((aligned(64)));
    for (int i = n_begin; i < n_end; i++) {
        const SvtB200TxJob *j = &jobs[i];
        /* the SIMD quantisers load 8-entry, 16-byte aligned rows (QUANTS in vp9_quantize.h: [q][8], entries 1..7 = AC) */
        struct { int16_t zbin[8], round[8], quant[8], quant_shift[8], dequant[8]; } qa __attribute__((aligned(32)));
        for (int k = 0; k < 8; k++) {
            qa.zbin[k] = qrows[j->qrow].zbin[k != 0]; qa.round[k] = qrows[j->qrow].round[k != 0]; qa.quant[k] = qrows[j->qrow].quant[k != 0];
            qa.quant_shift[k] = qrows[j->qrow].quant_shift[k != 0]; qa.dequant[k] = qrows[j->qrow].dequant[k != 0];
        }
        const typeof(qa) *q = &qa;
        const uint8_t *s = pl->src + j->src_offset, *p = pl->pred + j->pred_offset;
        int16_t *c = coeff ? coeff + (size_t)i * nn : tmp;
        const int tt = tx_size == 3 ? 0 : j->tx_type;
        const scan_order *so = tx_size == 3 ? &eb_vp9_default_scan_orders[TX_32X32] : &eb_vp9_scan_orders[tx_size][tt];
        eb_vp9_residual_kernel_func_ptr_array[asm_i][N >> 3]((uint8_t *)s, pl->src_stride, (uint8_t *)p, pl->pred_stride, resid, N, N, N);
        switch (tx_size) {
        case 3:
            if (j->flags & SVTB200_TXJ_PARTIAL32) { memset(c, 0, 1024 * sizeof(int16_t)); eb_vpx_partial_fdct32x32(resid, c, N); }
            else eb_vp9_fdct32x32(resid, c, N);
            eb_vp9_quantize_b_32x32(c, 1024, 0, q->zbin, q->round, q->quant, q->quant_shift, qcoeff + (size_t)i * nn, dqcoeff + (size_t)i * nn,
                                    q->dequant, &eob[i], so->scan, so->iscan);
            break;
        case 2: eb_vp9_fht16x16(resid, c, N, tt); break;
        case 1: eb_vp9_fht8x8(resid, c, N, tt); break;
        default: eb_vp9_fht4x4(resid, c, N, tt); break;
        }
        if (tx_size < 3)
            eb_vp9_quantize_b(c, nn, 0, q->zbin, q->round, q->quant, q->quant_shift, qcoeff + (size_t)i * nn, dqcoeff + (size_t)i * nn,
                              q->dequant, &eob[i], so->scan, so->iscan);
        if ((j->flags & SVTB200_TXJ_RECON) && pl->recon) {
            uint8_t *r = pl->recon + j->recon_offset;
            for (int y = 0; y < N; y++) memcpy(r + (size_t)y * pl->recon_stride, p + (size_t)y * pl->pred_stride, N);
            if (eob[i]) {
                const int16_t *d = dqcoeff + (size_t)i * nn;
                switch (tx_size) {
                case 3: eb_vp9_idct32x32_add(d, r, pl->recon_stride, eob[i]); break;
                case 2: eb_vp9_iht16x16_add(tt, d, r, pl->recon_stride, eob[i]); break;
                case 1: eb_vp9_iht8x8_add(tt, d, r, pl->recon_stride, eob[i]); break;
                default: eb_vp9_iht4x4_add(tt, d, r, pl->recon_stride, eob[i]); break;
                }
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------ zz sad */
#include "EbMotionEstimationProcess.h"
EbErrorType compute_zz_sad(MotionEstimationContext *context_ptr, SequenceControlSet *sequence_control_set_ptr,
                           PictureParentControlSet *picture_control_set_ptr, EbPictureBufferDesc *sixteenth_decimated_picture_ptr,
                           uint32_t x_sb_start_index, uint32_t x_sb_end_index, uint32_t y_sb_start_index, uint32_t y_sb_end_index);

/* the reference's compute_zz_sad over a whole picture; input_resolution 0..3 selects non_moving_th_shift */
int refdrv_zz_sad(void *h, const SvtB200Plane *cur16, const SvtB200Plane *prev_full, int w, int hgt, int input_resolution, uint8_t *scores) {
    RefMe *r = (RefMe *)h;
    const int sbs_w = (w + 63) / 64, sbs_h = (hgt + 63) / 64, n_sb = sbs_w * sbs_h;
    MotionEstimationContext mec;
    memset(&mec, 0, sizeof(mec));
    mec.me_context_ptr = r->me;
    SequenceControlSet *scs = (SequenceControlSet *)calloc(1, sizeof(SequenceControlSet));
    scs->picture_width_in_sb = (uint8_t)sbs_w; scs->input_resolution = (uint8_t)input_resolution;
    scs->luma_width = (uint16_t)w; scs->luma_height = (uint16_t)hgt;
    scs->sb_params_array = (SbParams *)calloc((size_t)n_sb, sizeof(SbParams));
    for (int i = 0; i < n_sb; i++) {
        SbParams *p = &scs->sb_params_array[i];
        p->origin_x = (uint16_t)((i % sbs_w) * 64); p->origin_y = (uint16_t)((i / sbs_w) * 64);
        p->width = (uint8_t)(w - p->origin_x < 64 ? w - p->origin_x : 64); p->height = (uint8_t)(hgt - p->origin_y < 64 ? hgt - p->origin_y : 64);
        p->is_complete_sb = (uint8_t)(p->width == 64 && p->height == 64);
    }
    PictureParentControlSet *cur = (PictureParentControlSet *)calloc(1, sizeof(PictureParentControlSet));
    PictureParentControlSet *prev = (PictureParentControlSet *)calloc(1, sizeof(PictureParentControlSet));
    EbObjectWrapper prev_w;
    memset(&prev_w, 0, sizeof(prev_w));
    EbPictureBufferDesc d_prev, d_cur16;
    desc_of(prev_full, &d_prev); desc_of(cur16, &d_cur16);
    prev->enhanced_picture_ptr = &d_prev;
    prev->non_moving_index_array = (uint8_t *)calloc((size_t)n_sb, 1);
    prev_w.object_ptr = prev;
    cur->previous_picture_control_set_wrapper_ptr = &prev_w;
    compute_zz_sad(&mec, scs, cur, &d_cur16, 0, (uint32_t)sbs_w, 0, (uint32_t)sbs_h);
    memcpy(scores, prev->non_moving_index_array, (size_t)n_sb);
    free(prev->non_moving_index_array); free(prev); free(cur); free(scs->sb_params_array); free(scs);
    return 0;
}

/* ------------------------------------------------------------------------------------------------ edge masks */
/* the reference's eb_vp9_setup_mask over a picture described by plain per-unit data: sb_type, tx_size, skip, ref_frame[0]
 * (0 = intra) and which of the two mode classes of mode_lf_lut the block's mode falls in; levels[ref * 2 + cls] plays
 * lf_info.lvl[0][ref][cls] (what eb_vp9_loop_filter_frame_init derives from the deltas) */
typedef struct { uint8_t sb_type, tx_size, skip, ref, cls; } RefdrvMi;
void eb_vp9_setup_mask(VP9_COMMON *const cm, const int mi_row, const int mi_col, ModeInfo **mi, const int mode_info_stride, LOOP_FILTER_MASK *lfm);

int refdrv_build_masks(int mi_rows, int mi_cols, const RefdrvMi *g, int stride, const uint8_t levels[8], SvtB200LfMask *out, int lfm_stride) {
    VP9_COMMON *cm = (VP9_COMMON *)calloc(1, sizeof(VP9_COMMON));
    ModeInfo *mis = (ModeInfo *)calloc((size_t)mi_rows * stride, sizeof(ModeInfo));
    ModeInfo **grid = (ModeInfo **)calloc((size_t)(mi_rows + 8) * stride + 8, sizeof(ModeInfo *));
    cm->mi_rows = mi_rows; cm->mi_cols = mi_cols; cm->mi_stride = stride;
    for (int ref = 0; ref < 4; ref++) for (int c = 0; c < 2; c++) cm->lf_info.lvl[0][ref][c] = levels[ref * 2 + c];
    for (int i = 0; i < mi_rows * stride; i++) {
        mis[i].sb_type = (BLOCK_SIZE)g[i].sb_type; mis[i].tx_size = (TX_SIZE)g[i].tx_size; mis[i].skip = (int8_t)g[i].skip;
        mis[i].ref_frame[0] = (MV_REFERENCE_FRAME)g[i].ref; mis[i].ref_frame[1] = (MV_REFERENCE_FRAME)-1;
        mis[i].mode = g[i].ref == 0 ? DC_PRED : (g[i].cls ? NEARESTMV : ZEROMV);
        grid[i] = &mis[i];
    }
    for (int r = 0; r < mi_rows; r += 8)
        for (int c = 0; c < mi_cols; c += 8)
            eb_vp9_setup_mask(cm, r, c, grid + (size_t)r * stride + c, stride, (LOOP_FILTER_MASK *)&out[(r >> 3) * lfm_stride + (c >> 3)]);
    free(grid); free(mis); free(cm);
    return 0;
}

/* ------------------------------------------------------------------------------------------------ filter level */
#include "vp9_encoder.h"
#include "vp9_picklpf.h"
int refdrv_pick_filter_level(int base_qindex, int key_frame) {
    VP9_COMP *cpi = (VP9_COMP *)calloc(1, sizeof(VP9_COMP));
    cpi->common.base_qindex = base_qindex; cpi->common.bit_depth = VPX_BITS_8;
    cpi->common.frame_type = key_frame ? KEY_FRAME : INTER_FRAME;
    cpi->common.lf.filter_level = 1;
    eb_vp9_pick_filter_level(cpi, LPF_PICK_FROM_Q);
    const int lvl = cpi->common.lf.filter_level * 256 + cpi->common.lf.sharpness_level;
    free(cpi);
    return lvl;
}
int refdrv_ac_quant(int qindex) { return eb_vp9_ac_quant(qindex, 0, VPX_BITS_8); }

/* ------------------------------------------------------------------------------------------------ deblocking */
int refdrv_deblock(const SvtB200LfParams *P, uint8_t *y, int ys, uint8_t *u, uint8_t *v, int cs, const SvtB200LfMask *lfm,
                   int lfm_stride, const SvtB200LfThresh *thr) {
    VP9_COMMON  *cm = (VP9_COMMON *)calloc(1, sizeof(VP9_COMMON));
    MACROBLOCKD *xd = (MACROBLOCKD *)calloc(1, sizeof(MACROBLOCKD));
    const int sbs_h = (P->mi_rows + 7) >> 3;
    cm->mi_rows = P->mi_rows; cm->mi_cols = P->mi_cols;
    cm->lf.lfm_stride = lfm_stride;
    cm->lf.lfm = (LOOP_FILTER_MASK *)malloc(sizeof(LOOP_FILTER_MASK) * (size_t)lfm_stride * sbs_h);
    memcpy(cm->lf.lfm, lfm, sizeof(LOOP_FILTER_MASK) * (size_t)lfm_stride * sbs_h);
    cm->lf.filter_level = P->filter_level;
    for (int l = 0; l < 64; l++) {
        memset(cm->lf_info.lfthr[l].mblim, thr[l].mblim, SIMD_WIDTH);
        memset(cm->lf_info.lfthr[l].lim, thr[l].lim, SIMD_WIDTH);
        memset(cm->lf_info.lfthr[l].hev_thr, thr[l].hev_thr, SIMD_WIDTH);
    }
    xd->plane[0].dst.buf = y; xd->plane[0].dst.stride = ys;
    xd->plane[1].dst.buf = u; xd->plane[1].dst.stride = cs; xd->plane[1].subsampling_x = xd->plane[1].subsampling_y = 1;
    xd->plane[2].dst.buf = v; xd->plane[2].dst.stride = cs; xd->plane[2].subsampling_x = xd->plane[2].subsampling_y = 1;
    eb_vp9_loop_filter_frame(cm, xd, P->filter_level, P->y_only, 0);
    free(cm->lf.lfm); free(cm); free(xd);
    return 0;
}

/* ------------------------------------------------------------------------------------------------ prediction, distortion */
#include "vp9_scale.h"
#include "vp9_filter.h"
#include "vp9_reconinter.h"
/* inter_predictor() exactly as build_inter_predictors calls it (vp9_reconinter.c:246): unscaled scale factors, kernel 0.
 * 'ref' = 1 averages into dst (second reference of a compound block).  flavour 0: the plain-C functions in the positions
 * eb_vp9_setup_scale_factors_for_frame gives them (vp9_scale.c:78-83, 127-128); 1: what the RTCD dispatch selected.
 * The SIMD copy / average kernels store aligned: dst must be aligned to the block width, as the encoder's buffers are. */
int refdrv_inter_predict(const uint8_t *pre, int ref_stride, uint8_t *dst, int dst_stride, int subpel_x, int subpel_y, int w, int h, int ref,
                         int flavour) {
    struct scale_factors sf;
    if (!g_init) return -1;
    eb_vp9_setup_scale_factors_for_frame(&sf, 1920, 1080, 1920, 1080);
    if (!flavour) {
        sf.predict[0][0][0] = eb_vp9_convolve_copy_c;   sf.predict[0][0][1] = eb_vp9_convolve_avg_c;
        sf.predict[0][1][0] = eb_vp9_convolve8_vert_c;  sf.predict[0][1][1] = eb_vp9_convolve8_avg_vert_c;
        sf.predict[1][0][0] = eb_vp9_convolve8_horiz_c; sf.predict[1][0][1] = eb_vp9_convolve8_avg_horiz_c;
        sf.predict[1][1][0] = eb_vp9_convolve8_c;       sf.predict[1][1][1] = eb_vp9_convolve8_avg_c;
    }
    inter_predictor(pre, ref_stride, dst, dst_stride, subpel_x, subpel_y, &sf, w, h, ref, eb_vp9_filter_kernels[0], 16, 16);
    return 0;
}
/* the table lookup of EbEncDecProcess.c:723; flavour 0: the C_DEFAULT row of the table */
int refdrv_full_distortion(int tx_size, int16_t *coeff, int16_t *recon_coeff, int eob, uint64_t out[2], int flavour) {
    const uint32_t n = 4u << tx_size;
    if (!g_init) return -1;
    full_distortion_intrinsic_func_ptr_array[flavour ? ((eb_vp9_ASM_TYPES & PREAVX2_MASK) && 1) : 0][eob != 0][0][n >> 3](coeff, n, recon_coeff, n, out, n, n);
    return 0;
}

/* pred[mode][tx_size] / dc_pred[have_left][have_top][tx_size] as eb_vp9_init_intra_predictors_internal fills them
 * (vp9_reconintra.c:64-87); flavour 0: the _c functions, 1: the RTCD selection. */
typedef void (*refdrv_intra_fn)(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
#define RD_SIZES(type, suffix) {eb_vp9_##type##_predictor_4x4##suffix, eb_vp9_##type##_predictor_8x8##suffix, eb_vp9_##type##_predictor_16x16##suffix, eb_vp9_##type##_predictor_32x32##suffix}
int refdrv_intra_predict(int mode, int tx_size, int have_left, int have_top, uint8_t *dst, int stride, const uint8_t *above, const uint8_t *left, int flavour) {
    if (!g_init || mode < 0 || mode > 9 || tx_size < 0 || tx_size > 3) return -1;
    const refdrv_intra_fn c_tab[10][4] = {RD_SIZES(dc, _c), RD_SIZES(v, _c), RD_SIZES(h, _c), RD_SIZES(d45, _c), RD_SIZES(d135, _c), RD_SIZES(d117, _c),
                                          RD_SIZES(d153, _c), RD_SIZES(d207, _c), RD_SIZES(d63, _c), RD_SIZES(tm, _c)};
    const refdrv_intra_fn s_tab[10][4] = {RD_SIZES(dc, ), RD_SIZES(v, ), RD_SIZES(h, ), RD_SIZES(d45, ), RD_SIZES(d135, ), RD_SIZES(d117, ),
                                          RD_SIZES(d153, ), RD_SIZES(d207, ), RD_SIZES(d63, ), RD_SIZES(tm, )};
    const refdrv_intra_fn c_dc[2][2][4] = {{RD_SIZES(dc_128, _c), RD_SIZES(dc_top, _c)}, {RD_SIZES(dc_left, _c), RD_SIZES(dc, _c)}};
    const refdrv_intra_fn s_dc[2][2][4] = {{RD_SIZES(dc_128, ), RD_SIZES(dc_top, )}, {RD_SIZES(dc_left, ), RD_SIZES(dc, )}};
    refdrv_intra_fn f = mode == 0 ? (flavour ? s_dc : c_dc)[!!have_left][!!have_top][tx_size] : (flavour ? s_tab : c_tab)[mode][tx_size];
    f(dst, stride, above, left);
    return 0;
}

/* ------------------------------------------------------------------------------------------------ rate estimation */
#include "EbEncDecProcess.h"
#include "vp9_block.h"
#include "vp9_entropy.h"
#include "vp9_tokenize.h"
#include "vp9_cost.h"
#include "vp9_onyxc_int.h"
int coeff_rate_estimate(struct EncDecContext *context_ptr, MACROBLOCK *x, int16_t *trans_coeff_buffer, uint16_t eob, int plane, int block, TX_SIZE tx_size,
                        int pt, int use_fast_coef_costing);

/* MACROBLOCK::token_costs as the reference fills them (fill_token_costs, vp9_rd.c:93-107 -- static there, so its loop is driven
 * from here with the reference's own non-static helpers) from its default coefficient probabilities (vp9_entropy.c):
 * out[tx_size 4][plane type 2][ref 2][band 6][2][ctx 6][token 12] */
int refdrv_token_costs(uint32_t *out) {
    VP9_COMMON *cm = (VP9_COMMON *)calloc(1, sizeof(VP9_COMMON));
    FRAME_CONTEXT *fc = (FRAME_CONTEXT *)calloc(1, sizeof(FRAME_CONTEXT));
    vp9_coeff_cost *c = (vp9_coeff_cost *)out;
    cm->fc = fc;
    eb_vp9_default_coef_probs(cm); /* the static default tables of vp9_entropy.c:258-919 */
    memset(out, 0, sizeof(vp9_coeff_cost) * TX_SIZES);
    for (int t = TX_4X4; t <= TX_32X32; ++t)
        for (int i = 0; i < PLANE_TYPES; ++i)
            for (int j = 0; j < REF_TYPES; ++j)
                for (int k = 0; k < COEF_BANDS; ++k)
                    for (int l = 0; l < BAND_COEFF_CONTEXTS(k); ++l) {
                        vpx_prob probs[ENTROPY_NODES];
                        eb_vp9_model_to_full_probs(fc->coef_probs[t][i][j][k][l], probs);
                        eb_vp9_cost_tokens((int *)c[t][i][j][k][0][l], probs, eb_vp9_coef_tree);
                        eb_vp9_cost_tokens_skip((int *)c[t][i][j][k][1][l], probs, eb_vp9_coef_tree);
                    }
    free(fc); free(cm);
    return (int)(sizeof(vp9_coeff_cost) * TX_SIZES);
}
/* the value -> token / cost tables and the cat6 cost tables of vp9_tokenize.c, as plain arrays */
int refdrv_value_tables(int16_t tok[133], int32_t cost[133], int16_t cat6_low[256], uint16_t cat6_high[64]) {
    for (int v = -66; v <= 66; v++) { tok[v + 66] = eb_vp9_dct_cat_lt_10_value_tokens[v].token; cost[v + 66] = eb_vp9_dct_cat_lt_10_value_cost[v]; }
    memcpy(cat6_low, eb_vp9_cat6_low_cost, 256 * sizeof(int16_t));
    memcpy(cat6_high, eb_vp9_cat6_high_cost, 64 * sizeof(uint16_t));
    return 0;
}
/* scan_order::neighbors of (tx_size, tx_type): MAX_NEIGHBORS * (N * N + 1) entries */
int refdrv_scan_neighbors(int tx_size, int tx_type, int16_t *out) {
    const int nn = (4 << tx_size) * (4 << tx_size);
    const scan_order *so = tx_size == 3 ? &eb_vp9_default_scan_orders[TX_32X32] : &eb_vp9_scan_orders[tx_size][tx_type];
    memcpy(out, so->neighbors, (size_t)MAX_NEIGHBORS * (nn + 1) * sizeof(int16_t));
    return 0;
}
/* the reference's coeff_rate_estimate on every block of a batch; token_costs as refdrv_token_costs lays them out (all sizes) */
int refdrv_coeff_rate(int tx_size, int n, const int16_t *qcoeff, const uint16_t *eob, const SvtB200RateBlock *blocks, const uint32_t *token_costs_all,
                      int32_t *rate) {
    static const PREDICTION_MODE mode_of_tx_type[4] = {DC_PRED, V_PRED, H_PRED, TM_PRED}; /* eb_vp9_intra_mode_to_tx_type_lookup backwards */
    const int nn = (4 << tx_size) * (4 << tx_size);
    EncDecContext *ctx = (EncDecContext *)calloc(1, sizeof(EncDecContext));
    MACROBLOCK *x = (MACROBLOCK *)calloc(1, sizeof(MACROBLOCK));
    MACROBLOCKD *xd = (MACROBLOCKD *)calloc(1, sizeof(MACROBLOCKD));
    ModeInfo mi, *mip = &mi;
    memcpy(x->token_costs, token_costs_all, sizeof(x->token_costs));
    ctx->e_mbd = xd;
    xd->mi = &mip;
    for (int b = 0; b < n; b++) {
        memset(&mi, 0, sizeof(mi));
        mi.sb_type = BLOCK_8X8;
        mi.ref_frame[0] = blocks[b].is_inter ? LAST_FRAME : INTRA_FRAME;
        mi.ref_frame[1] = NONE;
        mi.mode = blocks[b].is_inter ? ZEROMV : mode_of_tx_type[blocks[b].tx_type & 3];
        for (int k = 0; k < 4; k++) mi.bmi[k].as_mode = mi.mode;
        rate[b] = coeff_rate_estimate(ctx, x, (int16_t *)(qcoeff + (size_t)b * nn), eob[b], blocks[b].plane_type ? 1 : 0, 0, (TX_SIZE)tx_size, blocks[b].coeff_ctx, 0);
    }
    free(xd); free(x); free(ctx);
    return 0;
}

/* ------------------------------------------------------------------------------------------------ block walk of the encode pass
 * The reference's own inter_prediction() (EbIntraPrediction.c:49) -> build_inter_predictors() (vp9_reconinter.c:102) for blocks
 * [begin, end) and all three planes, on an EncDecContext / MACROBLOCKD filled the way the encode pass fills them
 * (EbEncDecProcess.c:2052-2066, :5431, :5507).  pic holds HOST pointers. */
#include "EbIntraPrediction.h"
#include "EbUtility.h"
#include "vp9_common_data.h"
void inter_prediction(struct EncDecContext *context_ptr, EbByte pred_buffer, uint16_t pred_stride, int plane);

typedef struct {
    EncDecContext       ctx;
    MACROBLOCKD         xd;
    ModeInfo            mi, *mip;
    EpBlockStats        bs;
    struct scale_factors sf;
    EbPictureBufferDesc ref[2];
} RefPredCtx;

void *refdrv_pred_new(void) {
    RefPredCtx *c = (RefPredCtx *)calloc(1, sizeof(RefPredCtx));
    c->ctx.e_mbd = &c->xd;
    c->mip = &c->mi;
    c->xd.mi = &c->mip;
    c->ctx.ep_block_stats_ptr = &c->bs;
    c->ctx.sf = &c->sf;
    c->ctx.ref_pic_list[0] = &c->ref[0];
    c->ctx.ref_pic_list[1] = &c->ref[1];
    c->xd.plane[1].subsampling_x = c->xd.plane[1].subsampling_y = 1;
    c->xd.plane[2].subsampling_x = c->xd.plane[2].subsampling_y = 1;
    return c;
}
void refdrv_pred_free(void *h) { free(h); }

static BLOCK_SIZE refdrv_bsize(int bw, int bh) {
    for (int b = BLOCK_8X8; b < BLOCK_SIZES; b++)
        if (4 * eb_vp9_num_4x4_blocks_wide_lookup[b] == bw && 4 * eb_vp9_num_4x4_blocks_high_lookup[b] == bh) return (BLOCK_SIZE)b;
    return BLOCK_INVALID;
}

int refdrv_inter_prediction(void *h, const SvtB200InterBlock *blocks, int begin, int end, const SvtB200PredPicture *pic) {
    RefPredCtx *c = (RefPredCtx *)h;
    if (!g_init || !c) return -1;
    const int mi_cols = (pic->width + 7) >> 3, mi_rows = (pic->height + 7) >> 3;
    eb_vp9_setup_scale_factors_for_frame(&c->sf, pic->width, pic->height, pic->width, pic->height);
    for (int l = 0; l < 2; l++) {
        c->ref[l].buffer_y = (EbByte)pic->ref[l][0]; c->ref[l].buffer_cb = (EbByte)pic->ref[l][1]; c->ref[l].buffer_cr = (EbByte)pic->ref[l][2];
        c->ref[l].stride_y = (uint16_t)pic->ref_stride[l][0]; c->ref[l].stride_cb = (uint16_t)pic->ref_stride[l][1]; c->ref[l].stride_cr = (uint16_t)pic->ref_stride[l][2];
        c->ref[l].origin_x = c->ref[l].origin_y = 0; /* the pointers already stand at sample (0,0) */
    }
    c->ctx.use_subpel_flag = (uint8_t)pic->use_subpel_flag;
    for (int i = begin; i < end; i++) {
        const SvtB200InterBlock *b = &blocks[i];
        const BLOCK_SIZE bsize = refdrv_bsize(b->bw, b->bh);
        if (bsize == BLOCK_INVALID) return -2;
        const int mi_col = b->x >> 3, mi_row = b->y >> 3;
        c->bs.bsize = bsize;
        c->bs.bsize_uv = eb_vp9_ss_size_lookup[bsize][1][1];
        c->ctx.block_origin_x = b->x; c->ctx.block_origin_y = b->y;
        c->xd.mb_to_top_edge = -((mi_row * MI_SIZE) * 8);
        c->xd.mb_to_bottom_edge = ((mi_rows - eb_vp9_num_8x8_blocks_high_lookup[bsize] - mi_row) * MI_SIZE) * 8;
        c->xd.mb_to_left_edge = -((mi_col * MI_SIZE) * 8);
        c->xd.mb_to_right_edge = ((mi_cols - eb_vp9_num_8x8_blocks_wide_lookup[bsize] - mi_col) * MI_SIZE) * 8;
        memset(&c->mi, 0, sizeof(c->mi));
        c->mi.sb_type = bsize;
        c->mi.mode = NEWMV;
        c->mi.ref_frame[0] = b->list[0] ? ALTREF_FRAME : LAST_FRAME;
        c->mi.ref_frame[1] = b->list[1] == 0xff ? NONE : (b->list[1] ? ALTREF_FRAME : LAST_FRAME);
        for (int r = 0; r < 2; r++) { c->mi.mv[r].as_mv.row = b->mv[r][0]; c->mi.mv[r].as_mv.col = b->mv[r][1]; }
        for (int plane = 0; plane < 3; plane++) {
            const int ox = plane ? ROUND_UV(b->x) >> 1 : b->x, oy = plane ? ROUND_UV(b->y) >> 1 : b->y;
            inter_prediction(&c->ctx, pic->dst[plane] + (ptrdiff_t)oy * pic->dst_stride[plane] + ox, (uint16_t)pic->dst_stride[plane], plane);
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------ CPU arm of bench.py
 * A native thread pool that runs the hot path of a whole mini-GOP with the reference's own functions: the unit of work is
 * a small range of super-blocks (motion estimation) or transform blocks (perform_coding_loop chain); one picture's
 * deblocking is one task (eb_vp9_loop_filter_frame is a single-threaded call in the reference as well) that becomes
 * runnable when the transform tasks of its picture have completed; idle threads keep pulling the tasks of the other
 * pictures meanwhile, as the reference's picture-level pipeline would.  One call per mini-GOP: no interpreter, no
 * per-stage barrier across pictures.  Threads are pinned round-robin to the CPUs the process may run on. */
#include <pthread.h>
#include <sched.h>
#include <time.h>
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
/* test aid: a backtrace on SIGSEGV (faults inside the reference's functions are otherwise silent under pytest) */
static void refdrv_segv(int sig) { void *bt[40]; int n = backtrace(bt, 40); (void)sig; backtrace_symbols_fd(bt, n, 2); _exit(139); }
void refdrv_debug_segv(void) { signal(SIGSEGV, refdrv_segv); }

typedef struct {
    int32_t              tx_size, n;
    const SvtB200TxJob  *jobs;
    SvtB200TxPlanes      pl;
    int16_t             *coeff, *qcoeff, *dqcoeff;
    uint16_t            *eob;
    const SvtB200RateBlock *rblk;  /* with rate: coeff_rate_estimate of every block after its transform chain */
    int32_t             *rate;
} RefdrvTuPass;

typedef struct {
    SvtB200MeParams        P;
    const SvtB200Plane    *cur, *ref0, *ref1; /* three planes each */
    SvtB200MeCuResult     *me_out;            /* n_sb * 85 */
    int32_t                n_pass;
    RefdrvTuPass           pass[8];
    const SvtB200QuantRow *qrows;
    int32_t                deblock;
    SvtB200LfParams        lf;
    uint8_t               *y, *u, *v;
    int32_t                ys, cs;
    const SvtB200LfMask   *lfm;
    int32_t                lfm_stride;
    const SvtB200LfThresh *thr;
    /* n_blocks > 0: inter prediction of the open-loop partition (32x32 blocks, 16x16 in the last strip; best motion-estimation
     * candidate each) into pred before the transform passes, which then take their prediction from there */
    int32_t                n_blocks;
    SvtB200PredPicture     pred;            /* HOST pointers */
    const uint32_t        *token_costs_all; /* [tx_size 4][...]: MACROBLOCK::token_costs, for the rate of the transform passes */
} RefdrvPicture;

typedef struct { int32_t kind, pic, pass, begin, end; } PoolTask; /* kind 0 ME, 1 TU, 2 LF, 3 inter prediction */

typedef struct RefPool {
    int                  n_threads, stop;
    pthread_t           *th;
    void               **me;       /* one RefMe per worker */
    pthread_mutex_t      mu;
    pthread_cond_t       cv_go, cv_done;
    unsigned             generation;
    int                  running;
    const RefdrvPicture *pics;
    int                  n_pics;
    PoolTask            *tasks;
    int                  n_tasks, cap_tasks;
    int                  next;     /* atomic */
    int                 *tu_left;  /* per picture: transform tasks not yet finished (atomic) */
    int                 *me_left, *pred_left; /* per picture: motion-estimation / prediction tasks not yet finished */
    void               **pred;     /* one RefPredCtx per worker */
    void               **rate;     /* one RefRateCtx per worker */
    double               cpu_s[4]; /* per kind, summed over threads */
} RefPool;

typedef struct { RefPool *p; int idx; } PoolArg;

typedef struct { EncDecContext ctx; MACROBLOCK x; MACROBLOCKD xd; ModeInfo mi, *mip; const uint32_t *loaded; } RefRateCtx;
static void *rate_ctx_new(void) {
    RefRateCtx *r = (RefRateCtx *)calloc(1, sizeof(RefRateCtx));
    r->ctx.e_mbd = &r->xd; r->mip = &r->mi; r->xd.mi = &r->mip;
    r->mi.sb_type = BLOCK_32X32; r->mi.ref_frame[0] = LAST_FRAME; r->mi.ref_frame[1] = NONE; r->mi.mode = NEWMV;
    return r;
}
/* coeff_rate_estimate over blocks [begin, end) of a transform pass (inter blocks: DCT_DCT scan) */
static void rate_range(RefRateCtx *r, const RefdrvTuPass *s, const uint32_t *token_costs_all, int begin, int end) {
    const int nn = (4 << s->tx_size) * (4 << s->tx_size);
    if (r->loaded != token_costs_all) { memcpy(r->x.token_costs, token_costs_all, sizeof(r->x.token_costs)); r->loaded = token_costs_all; }
    for (int b = begin; b < end; b++) {
        r->mi.ref_frame[0] = s->rblk[b].is_inter ? LAST_FRAME : INTRA_FRAME;
        s->rate[b] = coeff_rate_estimate(&r->ctx, &r->x, s->qcoeff + (size_t)b * nn, s->eob[b], s->rblk[b].plane_type ? 1 : 0, 0, (TX_SIZE)s->tx_size,
                                         s->rblk[b].coeff_ctx, 0);
    }
}
/* block i of the open-loop partition (the bench's stand-in for mode decision; the same rule as svtb200_me_best_blocks_dev) */
static SvtB200InterBlock open_loop_block(const SvtB200MeCuResult *me, int width, int height, int i) {
    const int w32 = width / 32, h32 = height / 32, w16 = width / 16, n32 = w32 * h32, nbot = (height / 16 - 2 * h32) * w16;
    int x, y, size = 16;
    if (i < n32) { y = (i / w32) * 32; x = (i % w32) * 32; size = 32; }
    else if (i < n32 + nbot) { y = (2 * h32 + (i - n32) / w16) * 16; x = ((i - n32) % w16) * 16; }
    else { const int j = i - n32 - nbot, wr = w16 - 2 * w32; y = (j / wr) * 16; x = (2 * w32 + j % wr) * 16; }
    const SvtB200MeCuResult *sb = me + (size_t)((y / 64) * ((width + 63) / 64) + x / 64) * 85;
    const SvtB200MeCuResult *r = size == 32 ? &sb[1 + ((y % 64) / 32) * 2 + (x % 64) / 32] : &sb[5 + ((y % 64) / 16) * 4 + (x % 64) / 16];
    const unsigned d = r->cand[0].direction;
    SvtB200InterBlock b;
    b.x = (uint16_t)x; b.y = (uint16_t)y; b.bw = b.bh = (uint8_t)size;
    b.list[0] = d == 1; b.list[1] = d == 2 ? 1 : 0xff;
    b.mv[0][0] = (int16_t)((d == 1 ? r->y_mv_l1 : r->y_mv_l0) << 1); b.mv[0][1] = (int16_t)((d == 1 ? r->x_mv_l1 : r->x_mv_l0) << 1);
    b.mv[1][0] = (int16_t)(d == 2 ? r->y_mv_l1 << 1 : 0); b.mv[1][1] = (int16_t)(d == 2 ? r->x_mv_l1 << 1 : 0);
    return b;
}

static double pool_now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

static void pool_do(RefPool *p, int w, const PoolTask *t, double acc[4]) {
    const RefdrvPicture *pic = &p->pics[t->pic];
    double t0 = pool_now();
    if (t->kind == 0) {
        refdrv_me_range(p->me[w], &pic->P, pic->cur, pic->ref0, pic->ref1, t->begin, t->end, pic->me_out);
        __atomic_sub_fetch(&p->me_left[t->pic], 1, __ATOMIC_ACQ_REL);
    } else if (t->kind == 3) {
        /* the tasks this one depends on were queued earlier, i.e. are already claimed by some thread: the wait is short */
        while (__atomic_load_n(&p->me_left[t->pic], __ATOMIC_ACQUIRE) > 0) sched_yield();
        t0 = pool_now();
        SvtB200InterBlock blk[64];
        for (int i = t->begin; i < t->end; i += 64) {
            const int n = t->end - i < 64 ? t->end - i : 64;
            for (int k = 0; k < n; k++) blk[k] = open_loop_block(pic->me_out, pic->pred.width, pic->pred.height, i + k);
            refdrv_inter_prediction(p->pred[w], blk, 0, n, &pic->pred);
        }
        __atomic_sub_fetch(&p->pred_left[t->pic], 1, __ATOMIC_ACQ_REL);
    } else if (t->kind == 1) {
        const RefdrvTuPass *s = &pic->pass[t->pass];
        while (__atomic_load_n(&p->pred_left[t->pic], __ATOMIC_ACQUIRE) > 0) sched_yield();
        t0 = pool_now();
        refdrv_txq(s->tx_size, t->begin, t->end, s->jobs, &s->pl, pic->qrows, s->coeff, s->qcoeff, s->dqcoeff, s->eob);
        if (s->rate && pic->token_costs_all) rate_range((RefRateCtx *)p->rate[w], s, pic->token_costs_all, t->begin, t->end);
        __atomic_sub_fetch(&p->tu_left[t->pic], 1, __ATOMIC_ACQ_REL);
    } else {
        while (__atomic_load_n(&p->tu_left[t->pic], __ATOMIC_ACQUIRE) > 0) sched_yield();
        t0 = pool_now();
        refdrv_deblock(&pic->lf, pic->y, pic->ys, pic->u, pic->v, pic->cs, pic->lfm, pic->lfm_stride, pic->thr);
    }
    acc[t->kind] += pool_now() - t0;
}

static void *pool_worker(void *a) {
    PoolArg *pa = (PoolArg *)a;
    RefPool *p = pa->p;
    const int w = pa->idx;
    unsigned seen = 0;
    free(pa);
    for (;;) {
        pthread_mutex_lock(&p->mu);
        while (!p->stop && p->generation == seen) pthread_cond_wait(&p->cv_go, &p->mu);
        if (p->stop) { pthread_mutex_unlock(&p->mu); return NULL; }
        seen = p->generation;
        pthread_mutex_unlock(&p->mu);
        double acc[4] = {0, 0, 0, 0};
        for (;;) {
            const int i = __atomic_fetch_add(&p->next, 1, __ATOMIC_RELAXED);
            if (i >= p->n_tasks) break;
            pool_do(p, w, &p->tasks[i], acc);
        }
        pthread_mutex_lock(&p->mu);
        for (int k = 0; k < 4; k++) p->cpu_s[k] += acc[k];
        if (--p->running == 0) pthread_cond_signal(&p->cv_done);
        pthread_mutex_unlock(&p->mu);
    }
}

void *refdrv_pool_new(int threads, int pin) {
    if (!g_init || threads < 1) return NULL;
    RefPool *p = (RefPool *)calloc(1, sizeof(RefPool));
    p->n_threads = threads;
    p->th = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    p->me = (void **)calloc((size_t)threads, sizeof(void *));
    p->pred = (void **)calloc((size_t)threads, sizeof(void *));
    p->rate = (void **)calloc((size_t)threads, sizeof(void *));
    for (int i = 0; i < threads; i++) { p->me[i] = refdrv_me_new(); p->pred[i] = refdrv_pred_new(); p->rate[i] = rate_ctx_new(); } /* the constructors share a global allocation map: single-threaded here */
    pthread_mutex_init(&p->mu, NULL);
    pthread_cond_init(&p->cv_go, NULL);
    pthread_cond_init(&p->cv_done, NULL);
    cpu_set_t allowed;
    int cpus[1024], n_cpus = 0;
    if (pin && sched_getaffinity(0, sizeof(allowed), &allowed) == 0)
        for (int c = 0; c < 1024 && c < CPU_SETSIZE; c++) if (CPU_ISSET(c, &allowed)) cpus[n_cpus++] = c;
    for (int i = 0; i < threads; i++) {
        PoolArg *a = (PoolArg *)malloc(sizeof(PoolArg));
        a->p = p; a->idx = i;
        pthread_create(&p->th[i], NULL, pool_worker, a);
        if (n_cpus) {
            cpu_set_t one;
            CPU_ZERO(&one); CPU_SET(cpus[i % n_cpus], &one);
            pthread_setaffinity_np(p->th[i], sizeof(one), &one);
        }
    }
    return p;
}

static void pool_push(RefPool *p, int kind, int pic, int pass, int begin, int end) {
    if (p->n_tasks == p->cap_tasks) {
        p->cap_tasks = p->cap_tasks ? 2 * p->cap_tasks : 4096;
        p->tasks = (PoolTask *)realloc(p->tasks, (size_t)p->cap_tasks * sizeof(PoolTask));
    }
    PoolTask t = {kind, pic, pass, begin, end};
    p->tasks[p->n_tasks++] = t;
}

/* Runs the pictures (in the order given: encode order) and returns when all are done; stage_cpu_s = thread-seconds spent in
 * motion estimation / transform chain (+ rate) / deblocking / inter prediction.  me_chunk / tu_chunk = super-blocks / transform
 * blocks per task.  Tasks are claimed in queue order; a picture's prediction, transform and deblocking tasks are queued behind the
 * NEXT picture's motion estimation, so that the tasks they wait for have long been claimed when their turn comes. */
static void pool_push_rest(RefPool *p, const RefdrvPicture *pic, int i, int tu_chunk) {
    int n_pred = 0, n_tu = 0;
    for (int s = 0; s < pic->n_blocks; s += 256, n_pred++) pool_push(p, 3, i, 0, s, s + 256 < pic->n_blocks ? s + 256 : pic->n_blocks);
    p->pred_left[i] = n_pred;
    for (int k = 0; k < pic->n_pass; k++)
        for (int s = 0; s < pic->pass[k].n; s += tu_chunk, n_tu++) pool_push(p, 1, i, k, s, s + tu_chunk < pic->pass[k].n ? s + tu_chunk : pic->pass[k].n);
    p->tu_left[i] = n_tu;
    if (pic->deblock) pool_push(p, 2, i, 0, 0, 0);
}
int refdrv_pool_run(void *h, const RefdrvPicture *pics, int n_pics, int me_chunk, int tu_chunk, double stage_cpu_s[4]) {
    RefPool *p = (RefPool *)h;
    if (!p || n_pics < 1) return -1;
    p->pics = pics; p->n_pics = n_pics; p->n_tasks = 0;
    p->tu_left = (int *)realloc(p->tu_left, (size_t)n_pics * sizeof(int));
    p->me_left = (int *)realloc(p->me_left, (size_t)n_pics * sizeof(int));
    p->pred_left = (int *)realloc(p->pred_left, (size_t)n_pics * sizeof(int));
    for (int i = 0; i < n_pics; i++) {
        const RefdrvPicture *pic = &pics[i];
        const int n_sb = ((pic->P.luma_width + 63) / 64) * ((pic->P.luma_height + 63) / 64);
        int n_me = 0;
        if (pic->me_out) for (int s = 0; s < n_sb; s += me_chunk, n_me++) pool_push(p, 0, i, 0, s, s + me_chunk < n_sb ? s + me_chunk : n_sb);
        p->me_left[i] = n_me;
        if (i > 0) pool_push_rest(p, &pics[i - 1], i - 1, tu_chunk);
    }
    pool_push_rest(p, &pics[n_pics - 1], n_pics - 1, tu_chunk);
    pthread_mutex_lock(&p->mu);
    p->next = 0; p->running = p->n_threads;
    for (int k = 0; k < 4; k++) p->cpu_s[k] = 0;
    p->generation++;
    pthread_cond_broadcast(&p->cv_go);
    while (p->running) pthread_cond_wait(&p->cv_done, &p->mu);
    if (stage_cpu_s) for (int k = 0; k < 4; k++) stage_cpu_s[k] = p->cpu_s[k];
    pthread_mutex_unlock(&p->mu);
    return 0;
}

void refdrv_pool_free(void *h) {
    RefPool *p = (RefPool *)h;
    if (!p) return;
    pthread_mutex_lock(&p->mu);
    p->stop = 1;
    pthread_cond_broadcast(&p->cv_go);
    pthread_mutex_unlock(&p->mu);
    for (int i = 0; i < p->n_threads; i++) pthread_join(p->th[i], NULL);
    for (int i = 0; i < p->n_threads; i++) { refdrv_pred_free(p->pred[i]); free(p->rate[i]); }
    free(p->tasks); free(p->tu_left); free(p->me_left); free(p->pred_left); free(p->th); free(p->me); free(p->pred); free(p->rate); free(p);
}

/* ------------------------------------------------------------------------------------------------ picture analysis: noise filters, edge maps */
#include "EbPictureAnalysisProcess.h" /* the reference's filter dispatch tables (static, one copy per includer) */

/* The noise filters of the picture-analysis stage exactly as detect_input_picture_noise / denoise_input_picture call them
 * (EbPictureAnalysisProcess.c:3538-3546, :3384-3440): per super-block row the dispatch-table entry at the row's first super-block, then the
 * plain-C function on the last super-block column when that one is partial.  kind: 0 luma weak (+ noise), 1 luma strong, 2 chroma weak,
 * 3 chroma strong; flavour: 0 = plain C entries, 1 = AVX2 entries.  plane / den: the sample at (0,0); w x h = the PLANE's size.
 * noise (kind 0): h rows of den_stride bytes -- the reference's noise buffer is one super-block row tall, each row is copied out after it was made. */
int refdrv_noise_filter(int kind, int flavour, uint8_t *plane, int stride, int w, int h, uint8_t *den, int den_stride, uint8_t *noise) {
    EbPictureBufferDesc in, dn, nz;
    memset(&in, 0, sizeof(in)); memset(&dn, 0, sizeof(dn)); memset(&nz, 0, sizeof(nz));
    const int chroma = kind >= 2, lw = chroma ? 2 * w : w, lh = chroma ? 2 * h : h;
    in.width = dn.width = nz.width = (uint16_t)lw; in.height = dn.height = nz.height = (uint16_t)lh;
    in.buffer_y = in.buffer_cb = in.buffer_cr = plane; in.stride_y = in.stride_cb = in.stride_cr = (uint16_t)stride;
    dn.buffer_y = dn.buffer_cb = dn.buffer_cr = den;   dn.stride_y = dn.stride_cb = dn.stride_cr = (uint16_t)den_stride;
    uint8_t *nbuf = NULL;
    if (kind == 0) { nbuf = (uint8_t *)calloc((size_t)den_stride * 66 + 64, 1); nz.buffer_y = nbuf; nz.stride_y = (uint16_t)den_stride; }
    for (int sy = 0; sy < lh; sy += 64)
        for (int sx = 0; sx < lw; sx += 64) {
            const uint32_t oy = (uint32_t)(chroma ? sy / 2 : sy), ox = (uint32_t)(chroma ? sx / 2 : sx);
            const int last_partial = sx + 64 > lw;
            switch (kind) {
            case 0:
                if (sx == 0) weak_luma_filter_func_ptr_array[flavour](&in, &dn, &nz, oy, ox);
                if (last_partial) eb_vp9_noise_extract_luma_weak(&in, &dn, &nz, oy, ox);
                break;
            case 1:
                if (sx == 0) strong_luma_filter_func_ptr_array[flavour](&in, &dn, oy, ox);
                if (last_partial) eb_vp9_noise_extract_luma_strong(&in, &dn, oy, ox);
                break;
            case 2:
                if (sx == 0) weak_chroma_filter_func_ptr_array[flavour](&in, &dn, oy, ox);
                if (last_partial) eb_vp9_noise_extract_chroma_weak(&in, &dn, oy, ox);
                break;
            default:
                if (sx == 0) strong_chroma_filter_func_ptr_array[flavour](&in, &dn, oy, ox);
                if (last_partial) eb_vp9_noise_extract_chroma_strong(&in, &dn, oy, ox);
            }
            if (kind == 0 && noise && sx + 64 >= lw) /* the row is complete */
                for (int y = 0; y < 64 && sy + y < lh; y++) memcpy(noise + (size_t)(sy + y) * den_stride, nbuf + (size_t)y * den_stride, (size_t)lw);
        }
    free(nbuf);
    return 0;
}

static SequenceControlSet *pa_scs(int w, int h, const uint8_t *logo_flags) {
    SequenceControlSet *scs = (SequenceControlSet *)calloc(1, sizeof(SequenceControlSet));
    const int sbs_w = (w + 63) / 64, sbs_h = (h + 63) / 64;
    scs->luma_width = (uint16_t)w; scs->luma_height = (uint16_t)h; scs->sb_total_count = (uint16_t)(sbs_w * sbs_h);
    scs->sb_params_array = (SbParams *)calloc((size_t)sbs_w * sbs_h, sizeof(SbParams));
    for (int i = 0; i < sbs_w * sbs_h; i++) {
        SbParams *p = &scs->sb_params_array[i];
        p->horizontal_index = (uint8_t)(i % sbs_w); p->vertical_index = (uint8_t)(i / sbs_w);
        p->origin_x = (uint16_t)((i % sbs_w) * 64); p->origin_y = (uint16_t)((i / sbs_w) * 64);
        p->is_complete_sb = p->origin_x + 64 <= w && p->origin_y + 64 <= h;
        p->potential_logo_sb = logo_flags ? logo_flags[i] : 0;
    }
    return scs;
}

/* edge_detection (EbPictureAnalysisProcess.c:4545) on given block statistics */
int refdrv_edge_detection(const uint8_t *y_mean, const uint16_t *variance, int w, int h, int pic_avg_variance, uint8_t *edge_block_num, uint8_t *isolated,
                          uint8_t *sharp_edge) {
    SequenceControlSet *scs = pa_scs(w, h, NULL);
    PictureParentControlSet *pcs = (PictureParentControlSet *)calloc(1, sizeof(PictureParentControlSet));
    const int n = scs->sb_total_count;
    pcs->sb_total_count = (uint16_t)n; pcs->pic_avg_variance = (uint16_t)pic_avg_variance;
    pcs->variance = (uint16_t **)calloc(n, sizeof(void *)); pcs->y_mean = (uint8_t **)calloc(n, sizeof(void *));
    for (int i = 0; i < n; i++) { pcs->variance[i] = (uint16_t *)variance + (size_t)i * 85; pcs->y_mean[i] = (uint8_t *)y_mean + (size_t)i * 85; }
    pcs->edge_results_ptr = (EdgeSbResults *)calloc(n, sizeof(EdgeSbResults));
    pcs->sharp_edge_sb_flag = (uint8_t *)calloc(n, 1);
    edge_detection(scs, pcs);
    for (int i = 0; i < n; i++) { edge_block_num[i] = pcs->edge_results_ptr[i].edge_block_num; isolated[i] = pcs->edge_results_ptr[i].isolated_high_intensity_sb; sharp_edge[i] = pcs->sharp_edge_sb_flag[i]; }
    free(pcs->edge_results_ptr); free(pcs->sharp_edge_sb_flag); free(pcs->variance); free(pcs->y_mean); free(pcs); free(scs->sb_params_array); free(scs);
    return 0;
}

/* edge_detection_mean_luma_chroma16x16 (:4434); logo_flags[sb] = SbParams::potential_logo_sb (is_complete_sb comes from the picture size) */
int refdrv_edge_cu(const uint8_t *y_mean, const uint8_t *cb_mean, const uint8_t *cr_mean, const uint8_t *logo_flags, int w, int h, uint64_t picture_number,
                   uint8_t *edge_cu) {
    void edge_detection_mean_luma_chroma16x16(SequenceControlSet *, PictureParentControlSet *, PictureAnalysisContext *, uint32_t);
    SequenceControlSet *scs = pa_scs(w, h, logo_flags);
    PictureParentControlSet *pcs = (PictureParentControlSet *)calloc(1, sizeof(PictureParentControlSet));
    PictureAnalysisContext *pa = (PictureAnalysisContext *)calloc(1, sizeof(PictureAnalysisContext));
    const int n = scs->sb_total_count;
    pcs->sb_total_count = (uint16_t)n; pcs->picture_number = picture_number;
    pcs->y_mean = (uint8_t **)calloc(n, sizeof(void *)); pcs->cb_mean = (uint8_t **)calloc(n, sizeof(void *)); pcs->cr_mean = (uint8_t **)calloc(n, sizeof(void *));
    pa->grad = (uint16_t **)calloc(n, sizeof(void *));
    for (int i = 0; i < n; i++) {
        pcs->y_mean[i] = (uint8_t *)y_mean + (size_t)i * 85; pcs->cb_mean[i] = (uint8_t *)cb_mean + (size_t)i * 21; pcs->cr_mean[i] = (uint8_t *)cr_mean + (size_t)i * 21;
        pa->grad[i] = (uint16_t *)calloc(PA_BLOCK_MAX_COUNT, sizeof(uint16_t));
    }
    pcs->sb_stat_array = (SbStat *)calloc(n, sizeof(SbStat));
    edge_detection_mean_luma_chroma16x16(scs, pcs, pa, (uint32_t)n);
    for (int i = 0; i < n; i++) for (int b = 0; b < 16; b++) edge_cu[(size_t)i * 16 + b] = pcs->sb_stat_array[i].cu_stat_array[5 + b].edge_cu;
    for (int i = 0; i < n; i++) free(pa->grad[i]);
    free(pa->grad); free(pa); free(pcs->sb_stat_array); free(pcs->y_mean); free(pcs->cb_mean); free(pcs->cr_mean); free(pcs); free(scs->sb_params_array); free(scs);
    return 0;
}

/* ------------------------------------------------------------------------------------------------ intra reference samples
 * The reference's own generate_intra_reference_samples() (EbEncDecProcess.c:1128) on an EncDecContext filled the way the encode pass fills
 * it (block origin, ep block statistics, bmi_index, MACROBLOCKD edges) with neighbour arrays that hold the samples around the block
 * (top_array / left_array / top_left_array, everything else poisoned -- including the context's own scratch rows, which keep whatever
 * the previous block left there in the encoder).  out: the 128-byte records of svtb200_intra_edges_dev. */
typedef struct { int32_t mi_row, mi_col; uint8_t bsize, tx_size, bmi_index, is_4x4; } RefdrvIntraBlk;
void generate_intra_reference_samples(SequenceControlSet *sequence_control_set_ptr, EncDecContext *context_ptr, int plane);

int refdrv_intra_edges(void *h, const uint8_t *plane_buf, int stride, int plane, int luma_w, int luma_h, const RefdrvIntraBlk *blks, int n, uint8_t *out) {
    RefPredCtx *c = (RefPredCtx *)h;
    if (!g_init || !c) return -1;
    SequenceControlSet scs;
    memset(&scs, 0, sizeof(scs));
    scs.luma_width = (uint16_t)luma_w; scs.luma_height = (uint16_t)luma_h; scs.chroma_width = (uint16_t)(luma_w >> 1); scs.chroma_height = (uint16_t)(luma_h >> 1);
    const int mi_cols = (luma_w + 7) >> 3, mi_rows = (luma_h + 7) >> 3, ss = plane ? 1 : 0;
    NeighborArrayUnit na;
    memset(&na, 0, sizeof(na));
    const size_t top_n = (size_t)luma_w + 256, left_n = (size_t)luma_h + 256, tl_n = 2 * MAX_PICTURE_HEIGHT_SIZE + 256;
    na.top_array = (uint8_t *)malloc(top_n); na.left_array = (uint8_t *)malloc(left_n); na.top_left_array = (uint8_t *)malloc(tl_n);
    uint8_t *above_ref = (uint8_t *)malloc(256), *left_ref = (uint8_t *)malloc(256);
    c->ctx.intra_above_ref = above_ref; c->ctx.intra_left_ref = left_ref;
    c->ctx.luma_recon_neighbor_array = c->ctx.cb_recon_neighbor_array = c->ctx.cr_recon_neighbor_array = &na;
    for (int i = 0; i < n; i++) {
        const RefdrvIntraBlk *b = &blks[i];
        const BLOCK_SIZE mi_bsize = (BLOCK_SIZE)b->bsize; /* the mode-info block (>= 8x8) */
        const int sq = b->is_4x4 ? 4 : 4 * eb_vp9_num_4x4_blocks_wide_lookup[mi_bsize];
        memset(na.top_array, 0xA5, top_n); memset(na.left_array, 0x5A, left_n); memset(na.top_left_array, 0xC3, tl_n);
        memset(above_ref, 0x77, 256); memset(left_ref, 0x99, 256);
        memset(c->ctx.above_data, 0xEE, sizeof(c->ctx.above_data)); memset(c->ctx.left_col, 0xDD, sizeof(c->ctx.left_col));
        c->bs.bsize = b->is_4x4 ? BLOCK_4X4 : mi_bsize;
        c->bs.tx_size = c->bs.tx_size_uv = (TX_SIZE)b->tx_size;
        c->bs.sq_size = sq; c->bs.sq_size_uv = sq >> 1 < 4 ? 4 : sq >> 1;
        c->ctx.bmi_index = b->bmi_index;
        c->ctx.block_origin_x = (uint16_t)(b->mi_col * 8 + 4 * (b->bmi_index & 1)); c->ctx.block_origin_y = (uint16_t)(b->mi_row * 8 + 4 * (b->bmi_index >> 1));
        c->xd.mb_to_top_edge = -((b->mi_row * MI_SIZE) * 8);
        c->xd.mb_to_bottom_edge = ((mi_rows - eb_vp9_num_8x8_blocks_high_lookup[mi_bsize] - b->mi_row) * MI_SIZE) * 8;
        c->xd.mb_to_left_edge = -((b->mi_col * MI_SIZE) * 8);
        c->xd.mb_to_right_edge = ((mi_cols - eb_vp9_num_8x8_blocks_wide_lookup[mi_bsize] - b->mi_col) * MI_SIZE) * 8;
        /* the neighbour arrays around the block, in the plane's coordinates */
        const int ox = plane ? ROUND_UV(c->ctx.block_origin_x) >> 1 : c->ctx.block_origin_x, oy = plane ? ROUND_UV(c->ctx.block_origin_y) >> 1 : c->ctx.block_origin_y;
        const int cnt = (plane ? c->bs.sq_size_uv : c->bs.sq_size) << 1;
        if (oy > 0) for (int k = 0; k < cnt; k++) na.top_array[ox + k] = plane_buf[(ptrdiff_t)(oy - 1) * stride + ox + k];
        if (ox > 0) for (int k = 0; k < cnt; k++) na.left_array[oy + k] = plane_buf[(ptrdiff_t)(oy + k) * stride + ox - 1];
        if (ox > 0 && oy > 0) na.top_left_array[(plane ? MAX_PICTURE_HEIGHT_SIZE >> 1 : MAX_PICTURE_HEIGHT_SIZE) + ox - oy] = plane_buf[(ptrdiff_t)(oy - 1) * stride + ox - 1];
        (void)ss;
        generate_intra_reference_samples(&scs, &c->ctx, plane);
        uint8_t *rec = out + (size_t)i * 128;
        const int bs = 4 << b->tx_size;
        memset(rec, 0, 128);
        memcpy(rec + 15, c->ctx.const_above_row[plane] - 1, (size_t)(2 * bs + 1));
        memcpy(rec + 80, c->ctx.left_col[plane], (size_t)bs);
    }
    free(na.top_array); free(na.left_array); free(na.top_left_array); free(above_ref); free(left_ref);
    return 0;
}
