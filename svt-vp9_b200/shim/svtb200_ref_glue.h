/*
 * svtb200_ref_glue.h -- the reference-side binding: translates the reference encoder's own
 * structures (PictureParentControlSet, MeContext, EbPaReferenceObject, EbPictureBufferDesc) into
 * the plain-C arguments of include/svtb200.h.  Compiled ONLY against the reference's headers
 * (-I<reference>/Source/Lib/Codec ...); nothing from the reference is copied here.  It is what a
 * maintainer would add next to motion_estimate_sb()'s call site
 * (Source/Lib/Codec/EbMotionEstimationProcess.c:1037) -- see INTEGRATION.md.
 */
#ifndef SVTB200_REF_GLUE_H
#define SVTB200_REF_GLUE_H

#include "EbDefinitions.h"
#include "EbPictureControlSet.h"
#include "EbSequenceControlSet.h"
#include "EbMotionEstimationContext.h"
#include "EbReferenceObject.h"
#include "EbPictureBufferDesc.h"

#include "vp9_onyxc_int.h"
#include "vp9_blockd.h"
#include "vp9_loopfilter.h"

#include "svtb200.h"

static inline void svtb200_glue_plane(const EbPictureBufferDesc *d, SvtB200Plane *p) {
    p->buf           = d->buffer_y;
    p->stride        = d->stride_y;
    p->origin_x      = d->origin_x;
    p->origin_y      = d->origin_y;
    p->width         = d->width;
    p->height        = d->height;
    p->padded_height = d->stride_y ? (int32_t)(d->luma_size / d->stride_y) : 0;
}

static inline void svtb200_glue_pa_planes(const EbPaReferenceObject *o, SvtB200Plane out[3]) {
    svtb200_glue_plane(o->input_padded_picture_ptr, &out[0]);
    svtb200_glue_plane(o->quarter_decimated_picture_ptr, &out[1]);
    svtb200_glue_plane(o->sixteenth_decimated_picture_ptr, &out[2]);
}

/* Everything motion_estimate_sb() (EbMotionEstimation.c:4524) reads from its control structures. */
static inline void svtb200_glue_me_params(PictureParentControlSet *pcs, const MeContext *me, SvtB200MeParams *P) {
    SequenceControlSet *scs = (SequenceControlSet *)pcs->sequence_control_set_wrapper_ptr->object_ptr;
    memset(P, 0, sizeof(*P));
    P->luma_width           = scs->luma_width;
    P->luma_height          = scs->luma_height;
    P->num_lists            = (pcs->slice_type == P_SLICE) ? 1 : 2;
    P->temporal_layer_index = pcs->temporal_layer_index;
    P->hierarchical_levels  = pcs->hierarchical_levels;
    P->ref_poc[0]           = pcs->ref_pic_poc_array[0];
    P->ref_poc[1]           = pcs->ref_pic_poc_array[1];
    P->enable_hme_flag      = pcs->enable_hme_flag;
    P->enable_hme_level0    = pcs->enable_hme_level_0_flag;
    P->enable_hme_level1    = pcs->enable_hme_level_1_flag;
    P->enable_hme_level2    = pcs->enable_hme_level_2_flag;
    P->single_hme_quadrant  = me->single_hme_quadrant;
    P->number_hme_search_region_in_width   = me->number_hme_search_region_in_width;
    P->number_hme_search_region_in_height  = me->number_hme_search_region_in_height;
    P->hme_level0_total_search_area_width  = me->hme_level0_total_search_area_width;
    P->hme_level0_total_search_area_height = me->hme_level0_total_search_area_height;
    for (int i = 0; i < 2; i++) {
        P->hme_level0_search_area_in_width[i]  = me->hme_level0_search_area_in_width_array[i];
        P->hme_level0_search_area_in_height[i] = me->hme_level0_search_area_in_height_array[i];
        P->hme_level1_search_area_in_width[i]  = me->hme_level1_search_area_in_width_array[i];
        P->hme_level1_search_area_in_height[i] = me->hme_level1_search_area_in_height_array[i];
        P->hme_level2_search_area_in_width[i]  = me->hme_level2_search_area_in_width_array[i];
        P->hme_level2_search_area_in_height[i] = me->hme_level2_search_area_in_height_array[i];
    }
    P->hme_l0_multiplier_x = hme_level_0_search_area_multiplier_x[pcs->hierarchical_levels][pcs->temporal_layer_index];
    P->hme_l0_multiplier_y = hme_level_0_search_area_multiplier_y[pcs->hierarchical_levels][pcs->temporal_layer_index];
    P->search_area_width        = me->search_area_width;
    P->search_area_height       = me->search_area_height;
    P->fractional_search_method = me->fractional_search_method;
    P->fractional_search_model  = me->fractional_search_model;
    P->fractional_search64x64   = me->fractional_search64x64;
    P->cu8x8_mode               = pcs->cu8x8_mode;
    P->cu16x16_mode             = pcs->cu16x16_mode;
    P->rate_control_mode        = scs->static_config.rate_control_mode;
}


/* Arguments of eb_vp9_loop_filter_frame() (Source/Lib/VPX/vp9_loopfilter.c:1521) as plain data: picture geometry,
 * the per-SB LOOP_FILTER_MASK array (same layout as SvtB200LfMask) and the 64 threshold triples. */
static inline void svtb200_glue_lf(VP9_COMMON *cm, int filter_level, int y_only, SvtB200LfParams *P, SvtB200LfThresh thr[64]) {
    memset(P, 0, sizeof(*P));
    P->mi_rows = cm->mi_rows; P->mi_cols = cm->mi_cols;
    P->width = cm->mi_cols * 8; P->height = cm->mi_rows * 8;
    P->y_only = y_only; P->filter_level = filter_level;
    for (int l = 0; l < 64; l++) {
        thr[l].mblim = cm->lf_info.lfthr[l].mblim[0];
        thr[l].lim = cm->lf_info.lfthr[l].lim[0];
        thr[l].hev_thr = cm->lf_info.lfthr[l].hev_thr[0];
    }
}

#endif
