/*
 * TEST INFRASTRUCTURE ONLY.  Link-time interposer on the UNMODIFIED reference objects
 * (-Wl,--wrap=motion_estimate_sb): every call first runs the reference's own
 * motion_estimate_sb() and then, depending on SVTB200_SEAM_MODE,
 *   check : runs the CPU oracle (oracle/me_port.c) on the same inputs and counts mismatches
 *           (this is how the oracle restatement is pinned against the reference);
 *   dump  : writes the picture's planes, the translated parameters and the reference's results
 *           to $SVTB200_SEAM_DIR (golden fixtures for tests/golden, see oracle/make_golden.py).
 * The summary goes to $SVTB200_SEAM_DIR/summary.txt (the app closes stderr before exit).
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "svtb200_ref_glue.h"
#include "../oracle/oracle.h"

EbErrorType __real_motion_estimate_sb(PictureParentControlSet *pcs, uint32_t sb_index, uint32_t sb_origin_x,
                                      uint32_t sb_origin_y, MeContext *me, EbPictureBufferDesc *input);

static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;
static int   g_mode = -1; /* 0 off, 1 check, 2 dump */
static char  g_dir[512];
static long  g_sb_total, g_sb_bad, g_pu_bad, g_printed, g_stats_bad, g_hist_pics, g_hist_bad, g_chroma_bad, g_chroma_hist_bad;
static long  g_edge_bad, g_edge_cu_bad, g_edge_cu_set, g_noise_pics, g_noise_bad, g_noise_method = -1, g_noise_class_hist[12], g_noise_flat_set, g_dbg_row[4], g_dbg_col[4], g_dbg_ref1, g_dbg_mine1, g_homog_bad, g_homog_nonflat;
static int   g_max_dump_pics = 4, g_dump_from = 0;

#define MAX_TRACKED 64
typedef struct {
    uint64_t pic; int used; int n_sb, done; FILE *f; SvtB200MeCuResult *res;
} DumpPic;
static DumpPic g_dp[MAX_TRACKED];
static int g_dumped;

static void write_summary(void) {
    char path[600];
    snprintf(path, sizeof(path), "%s/summary.txt", g_dir);
    FILE *f = fopen(path, "w");
    if (!f) return;
    fprintf(f, "sb_total %ld\nsb_bad %ld\npu_bad %ld\ndumped %d\nstats_bad %ld\nhist_pics %ld\nhist_bad %ld\nchroma_bad %ld\nchroma_hist_bad %ld\n", g_sb_total, g_sb_bad, g_pu_bad, g_dumped, g_stats_bad, g_hist_pics, g_hist_bad, g_chroma_bad, g_chroma_hist_bad);
    fprintf(f, "edge_bad %ld\nedge_cu_bad %ld\nedge_cu_set %ld\nnoise_pics %ld\nnoise_bad %ld\n", g_edge_bad, g_edge_cu_bad, g_edge_cu_set, g_noise_pics, g_noise_bad);
    fprintf(f, "noise_flat_set %ld\nnoise_method %ld\nhomog_bad %ld\nhomog_not_homogeneous %ld\n", g_noise_flat_set, g_noise_method, g_homog_bad, g_homog_nonflat);
    if (g_dbg_mine1 + g_dbg_ref1) fprintf(f, "dbg_rows %ld,%ld,%ld,%ld\ndbg_cols %ld,%ld,%ld,%ld\ndbg_only_oracle %ld\ndbg_only_reference %ld\n", g_dbg_row[0], g_dbg_row[1], g_dbg_row[2], g_dbg_row[3], g_dbg_col[0], g_dbg_col[1], g_dbg_col[2], g_dbg_col[3], g_dbg_mine1, g_dbg_ref1); /* NOISE_DETECT_HALF / QUARTER / FULL_PRECISION = 0 / 1 / 2 */
    for (int i = 0; i < 12; i++) if (g_noise_class_hist[i]) fprintf(f, "noise_class_%d %ld\n", i, g_noise_class_hist[i]);
    fclose(f);
}

static void init_once(void) {
    if (g_mode >= 0) return;
    const char *m = getenv("SVTB200_SEAM_MODE"), *d = getenv("SVTB200_SEAM_DIR");
    g_mode = 0;
    if (m && !strcmp(m, "check")) g_mode = 1;
    if (m && !strcmp(m, "dump")) g_mode = 2;
    snprintf(g_dir, sizeof(g_dir), "%s", d ? d : "/tmp");
    if (getenv("SVTB200_SEAM_MAXPICS")) g_max_dump_pics = atoi(getenv("SVTB200_SEAM_MAXPICS"));
    if (getenv("SVTB200_SEAM_FROM")) g_dump_from = atoi(getenv("SVTB200_SEAM_FROM"));
    if (g_mode) atexit(write_summary);
}

static void ref_to_abi(const MeCuResults *r, SvtB200MeCuResult *o, int nl) {
    memset(o, 0, sizeof(*o));
    o->x_mv_l0 = r->x_mv_l0; o->y_mv_l0 = r->y_mv_l0;
    if (nl == 2) { o->x_mv_l1 = r->x_mv_l1; o->y_mv_l1 = r->y_mv_l1; } /* P pictures: stale scratch, ignored */
    o->total_me_candidate_index = r->total_me_candidate_index;
    for (int i = 0; i < r->total_me_candidate_index && i < 3; i++) {
        o->cand[i].distortion = r->distortion_direction[i].distortion;
        o->cand[i].direction  = r->distortion_direction[i].direction;
    }
}

static void write_plane(FILE *f, const SvtB200Plane *p) {
    int32_t hdr[6] = {p->stride, p->origin_x, p->origin_y, p->width, p->height, p->padded_height};
    fwrite(hdr, sizeof(hdr), 1, f);
    if (p->buf && p->padded_height > 0) fwrite(p->buf, 1, (size_t)p->stride * p->padded_height, f);
}

EbErrorType __wrap_motion_estimate_sb(PictureParentControlSet *pcs, uint32_t sb_index, uint32_t sb_origin_x,
                                      uint32_t sb_origin_y, MeContext *me, EbPictureBufferDesc *input) {
    EbErrorType rc = __real_motion_estimate_sb(pcs, sb_index, sb_origin_x, sb_origin_y, me, input);
    pthread_mutex_lock(&g_mu);
    init_once();
    int mode = g_mode;
    pthread_mutex_unlock(&g_mu);
    if (!mode) return rc;

    SvtB200MeParams P;
    svtb200_glue_me_params(pcs, me, &P);
    SvtB200Plane cur[3], r0[3], r1[3];
    svtb200_glue_pa_planes((EbPaReferenceObject *)pcs->pareference_picture_wrapper_ptr->object_ptr, cur);
    svtb200_glue_pa_planes((EbPaReferenceObject *)pcs->ref_pa_pic_ptr_array[0]->object_ptr, r0);
    if (P.num_lists == 2) svtb200_glue_pa_planes((EbPaReferenceObject *)pcs->ref_pa_pic_ptr_array[1]->object_ptr, r1);
    /* the source SB is read through the PA reference object's padded plane (aliases enhanced_picture) */
    SvtB200MeCuResult refres[85];
    for (int i = 0; i < 85; i++) ref_to_abi(&pcs->me_results[sb_index][i], &refres[i], P.num_lists);

    if (mode == 1) {
        SvtB200MeCuResult mine[85];
        uint32_t rcme = 0;
        SvtB200OracleSbTrace tr; memset(&tr, 0, sizeof(tr));
        svtb200_oracle_me_sb(&P, cur, r0, P.num_lists == 2 ? r1 : NULL, sb_origin_x / 64, sb_origin_y / 64, mine, &rcme, &tr);
        int bad = 0;
        for (int i = 0; i < 85; i++) if (memcmp(&mine[i], &refres[i], sizeof(mine[i]))) bad++;
        if (sb_index == 0) { /* once per picture: the luma histograms of the 1/16 plane against the oracle restatement */
            SequenceControlSet *scs = (SequenceControlSet *)pcs->sequence_control_set_wrapper_ptr->object_ptr;
            const int rw = (int)scs->picture_analysis_number_of_regions_per_width, rh = (int)scs->picture_analysis_number_of_regions_per_height;
            uint32_t *hg = (uint32_t *)malloc((size_t)rw * rh * 256 * sizeof(uint32_t));
            uint8_t av[256]; uint64_t s16 = 0;
            svtb200_oracle_luma_histogram(&cur[2], rw, rh, hg, av, &s16);
            long hb = 0;
            for (int iw = 0; iw < rw; iw++) for (int ih = 0; ih < rh; ih++) {
                for (int b = 0; b < 256; b++) if (hg[((size_t)iw * rh + ih) * 256 + b] != pcs->picture_histogram[iw][ih][0][b]) hb++;
                if (av[iw * rh + ih] != (uint8_t)pcs->average_intensity_per_region[iw][ih][0]) hb++;
            }
            free(hg);
            /* ... and the chroma statistics of the picture (block means of every super-block, region histograms) */
            EbPictureBufferDesc *ip = pcs->enhanced_picture_ptr;
            const uint8_t *cb = ip->buffer_cb + (size_t)(ip->origin_y >> 1) * ip->stride_cb + (ip->origin_x >> 1);
            const uint8_t *cr = ip->buffer_cr + (size_t)(ip->origin_y >> 1) * ip->stride_cr + (ip->origin_x >> 1);
            const int n_sb = (int)pcs->sb_total_count;
            uint8_t *mb = (uint8_t *)malloc((size_t)n_sb * 21), *mr = (uint8_t *)malloc((size_t)n_sb * 21);
            long cbad = 0, chb = 0;
            svtb200_oracle_chroma_block_mean(cb, cr, ip->stride_cb, ip->width, ip->height, mb, mr);
            for (int s = 0; s < n_sb; s++) for (int i = 0; i < 21; i++) if (mb[s * 21 + i] != pcs->cb_mean[s][i] || mr[s * 21 + i] != pcs->cr_mean[s][i]) cbad++;
            free(mb); free(mr);
            uint32_t *ch = (uint32_t *)malloc((size_t)rw * rh * 2 * 256 * sizeof(uint32_t));
            uint8_t cav[512]; uint64_t cs16[2];
            svtb200_oracle_chroma_histogram(cb, cr, ip->stride_cb, ip->width, ip->height, rw, rh, ch, cav, cs16);
            for (int iw = 0; iw < rw; iw++) for (int ih = 0; ih < rh; ih++) for (int p = 0; p < 2; p++) {
                for (int b = 0; b < 256; b++) if (ch[(((size_t)iw * rh + ih) * 2 + p) * 256 + b] != pcs->picture_histogram[iw][ih][1 + p][b]) chb++;
                if (cav[(iw * rh + ih) * 2 + p] != (uint8_t)pcs->average_intensity_per_region[iw][ih][1 + p]) chb++;
            }
            free(ch);
            /* ... the two edge maps of compute_picture_spatial_statistics (edge_detection, edge_detection_mean_luma_chroma16x16) ... */
            long ebad = 0, ecbad = 0, ecset = 0, nbad = 0, npics = 0, nflat = 0;
            {
                uint8_t *ym = (uint8_t *)malloc((size_t)n_sb * 85), *cbm = (uint8_t *)malloc((size_t)n_sb * 21), *crm = (uint8_t *)malloc((size_t)n_sb * 21);
                uint16_t *vr = (uint16_t *)malloc((size_t)n_sb * 85 * 2);
                uint8_t *o3 = (uint8_t *)malloc((size_t)n_sb * 3), *fl = (uint8_t *)malloc((size_t)n_sb), *ecu = (uint8_t *)malloc((size_t)n_sb * 16);
                for (int s = 0; s < n_sb; s++) {
                    memcpy(ym + (size_t)s * 85, pcs->y_mean[s], 85); memcpy(vr + (size_t)s * 85, pcs->variance[s], 85 * 2);
                    memcpy(cbm + (size_t)s * 21, pcs->cb_mean[s], 21); memcpy(crm + (size_t)s * 21, pcs->cr_mean[s], 21);
                    fl[s] = scs->sb_params_array[s].potential_logo_sb && scs->sb_params_array[s].is_complete_sb;
                }
                svtb200_oracle_edge_detection(ym, vr, scs->luma_width, scs->luma_height, pcs->pic_avg_variance, o3, o3 + n_sb, o3 + 2 * n_sb);
                for (int s = 0; s < n_sb; s++)
                    ebad += o3[s] != pcs->edge_results_ptr[s].edge_block_num || o3[n_sb + s] != pcs->edge_results_ptr[s].isolated_high_intensity_sb ||
                            o3[2 * n_sb + s] != pcs->sharp_edge_sb_flag[s];
                {   /* determine_homogeneous_region_in_picture: variance of variances, homogeneous-area flags, the two picture flags */
                    uint64_t *vv = (uint64_t *)malloc((size_t)n_sb * 32), cnt[3];
                    uint8_t *hm = (uint8_t *)malloc((size_t)n_sb);
                    long hbad = 0, hnf = 0;
                    svtb200_oracle_homogeneous_regions(vr, scs->luma_width, scs->luma_height, 5 /* SB_LOW_VAR_TH */, 64 * 64 /* VAR_BASED_DETAIL_PRESERVATION_SELECTOR_THRSLHD */, vv, hm, cnt);
                    for (int s = 0; s < n_sb; s++) {
                        for (int k = 0; k < 4; k++) hbad += vv[(size_t)s * 4 + k] != pcs->var_of_var_32x32_based_sb_array[s][k];
                        hbad += (hm[s] != 0) != (pcs->sb_homogeneous_area_array[s] != 0);
                        hnf += !hm[s];
                    }
                    hbad += (cnt[0] > 0 && (cnt[1] * 100) / cnt[0] > 60 /* PIC_LOW_VAR_PERCENTAGE_TH */) != (pcs->very_low_var_pic_flag != 0);
                    hbad += (cnt[0] > 0 && (cnt[1] * 100) / cnt[0] > 80) != (pcs->logo_pic_flag != 0);
                    free(vv); free(hm);
                    pthread_mutex_lock(&g_mu); g_homog_bad += hbad; g_homog_nonflat += hnf; pthread_mutex_unlock(&g_mu);
                }
                svtb200_oracle_edge_cu(ym, cbm, crm, fl, n_sb, pcs->picture_number, ecu);
                for (int s = 0; s < n_sb; s++)
                    for (int b = 0; b < 16; b++) { ecbad += ecu[s * 16 + b] != pcs->sb_stat_array[s].cu_stat_array[5 + b].edge_cu; ecset += ecu[s * 16 + b]; }
                free(ym); free(cbm); free(crm); free(vr); free(o3); free(fl); free(ecu);
            }
            /* ... and the noise detection (detect_input_picture_noise, full-sample method) -- on pictures the denoiser left alone: what motion
             * estimation sees is then still what picture analysis measured (classes from 3_1 up, and flat-noise super-blocks at 4K, are
             * filtered in place, EbPictureAnalysisProcess.c:3358) */
            if (pcs->noise_detection_method == NOISE_DETECT_FULL_PRECISION && pcs->pic_noise_class < PIC_NOISE_CLASS_3_1 && scs->input_resolution != INPUT_SIZE_4K_RANGE) {
                const uint8_t *luma = ip->buffer_y + (size_t)ip->origin_y * ip->stride_y + ip->origin_x;
                uint8_t *den = (uint8_t *)malloc((size_t)ip->width * ip->height), *fl = (uint8_t *)malloc((size_t)n_sb);
                uint64_t *nv = (uint64_t *)malloc((size_t)n_sb * 8), *dv = (uint64_t *)malloc((size_t)n_sb * 8), tot[2];
                svtb200_oracle_noise_detect(luma, ip->stride_y, ip->width, ip->height, den, ip->width, pcs->noise_detection_th == 1 ? 70000 : 120000 /* NOISE_MIN_LEVEL_0 / _1 */,
                                            50 /* FLAT_MAX_VAR */, nv, dv, fl, tot);
                for (int s = 0; s < n_sb; s++) { nbad += fl[s] != pcs->sb_flat_noise_array[s]; nflat += fl[s]; }
                const uint64_t v = tot[1] ? tot[0] / tot[1] : 0, th = scs->luma_height <= 720 ? 25 : 0; /* the class thresholds of :3587-3615 */
                int cls = v >= 80 + th ? 11 : v >= 70 + th ? 10 : v >= 60 + th ? 9 : v >= 50 + th ? 8 : v >= 40 + th ? 7 : v >= 30 + th ? 6 : v >= 20 + th ? 5 : v >= 17 + th ? 4 :
                          v >= 10 + th ? 3 : v >= 5 + th ? 2 : 1;
                if (cls >= PIC_NOISE_CLASS_4) cls = PIC_NOISE_CLASS_3_1;
                nbad += cls != (int)pcs->pic_noise_class;
                npics++;
                free(den); free(fl); free(nv); free(dv);
            }
            /* the decimated methods (quarter_sample_detect_noise on the 1/4 plane, sub_sample_detect_noise on the 1/16 plane): comparable on
             * pictures of class 1 and 3, which sub_sample_filter_noise leaves alone (:4089-4100) */
            if (pcs->noise_detection_method != NOISE_DETECT_FULL_PRECISION && (pcs->pic_noise_class == PIC_NOISE_CLASS_1 || pcs->pic_noise_class == PIC_NOISE_CLASS_3)) {
                const int quarter = pcs->noise_detection_method == NOISE_DETECT_QUARTER_PRECISION;
                const SvtB200Plane *dp = quarter ? &cur[1] : &cur[2];
                if (dp->buf) {
                    const int pw = (scs->luma_width + 63) / 64;
                    const uint8_t *d0 = dp->buf + (size_t)dp->origin_y * dp->stride + dp->origin_x;
                    uint8_t *den = (uint8_t *)malloc((size_t)dp->width * dp->height), *fl = (uint8_t *)malloc((size_t)n_sb);
                    uint64_t *nv = (uint64_t *)malloc((size_t)n_sb * 8), *dv = (uint64_t *)malloc((size_t)n_sb * 8), tot[2];
                    svtb200_oracle_noise_detect_decimated(d0, dp->stride, dp->width, dp->height, quarter ? 32 : 16, pw, n_sb, den, dp->width,
                                                          pcs->noise_detection_th == 1 ? 70000 : 120000 /* NOISE_MIN_LEVEL_DECIM_0 / _1 */, 50 /* FLAT_MAX_VAR_DECIM */, nv, dv, fl, tot);
                    for (int s = 0; s < n_sb; s++) { nbad += fl[s] != pcs->sb_flat_noise_array[s]; nflat += fl[s];
                        if (fl[s] != pcs->sb_flat_noise_array[s]) { g_dbg_row[(s / pw) & 3]++; g_dbg_col[(s % pw) & 3]++; if (fl[s]) g_dbg_mine1++; else g_dbg_ref1++; } }
                    const uint64_t v = tot[1] ? tot[0] / tot[1] : 0;
                    int cls;
                    if (quarter) cls = v > 60 ? PIC_NOISE_CLASS_3_1 : v >= 10 ? PIC_NOISE_CLASS_3 : v >= 5 ? PIC_NOISE_CLASS_2 : PIC_NOISE_CLASS_1;
                    else { const uint64_t th = scs->luma_height <= 720 ? 25 : scs->luma_height <= 1080 ? 10 : 0;
                           cls = v >= 55 + th ? PIC_NOISE_CLASS_3_1 : v >= 10 + th ? PIC_NOISE_CLASS_3 : v >= 5 + th ? PIC_NOISE_CLASS_2 : PIC_NOISE_CLASS_1; }
                    nbad += cls != (int)pcs->pic_noise_class;
                    npics++;
                    free(den); free(fl); free(nv); free(dv);
                }
            }
            pthread_mutex_lock(&g_mu); g_noise_method = pcs->noise_detection_method; if ((unsigned)pcs->pic_noise_class < 12) g_noise_class_hist[pcs->pic_noise_class]++; pthread_mutex_unlock(&g_mu);
            pthread_mutex_lock(&g_mu); g_hist_pics++; g_hist_bad += hb; g_chroma_bad += cbad; g_chroma_hist_bad += chb;
            g_edge_bad += ebad; g_edge_cu_bad += ecbad; g_edge_cu_set += ecset; g_noise_pics += npics; g_noise_bad += nbad; g_noise_flat_set += nflat; pthread_mutex_unlock(&g_mu);
        }
        {   /* the picture-analysis statistics of this SB (compute_block_mean_compute_variance) against the oracle restatement */
            SvtB200Plane one = cur[0];
            uint8_t ym[85]; uint16_t var[85];
            one.origin_x += (int32_t)sb_origin_x; one.origin_y += (int32_t)sb_origin_y;
            svtb200_oracle_block_stats_picture(&one, 64, 64, ym, var);
            int sbad = 0;
            for (int i = 0; i < 85; i++) if (ym[i] != pcs->y_mean[sb_index][i] || var[i] != pcs->variance[sb_index][i]) sbad++;
            if (sbad) { pthread_mutex_lock(&g_mu); g_stats_bad += sbad; pthread_mutex_unlock(&g_mu); }
        }
        if (P.rate_control_mode && rcme != pcs->rcme_distortion[sb_index]) bad++;
        pthread_mutex_lock(&g_mu);
        g_sb_total++;
        if ((g_sb_total & 1023) == 0) write_summary(); /* the 4K reference aborts in deinit: atexit never runs */
        if (bad) {
            g_sb_bad++; g_pu_bad += bad;
            if (g_printed < 20) {
                g_printed++;
                char path[600]; snprintf(path, sizeof(path), "%s/mismatch.txt", g_dir);
                FILE *f = fopen(path, "a");
                if (f) {
                    fprintf(f, "pic %llu sb %u (%u,%u) nl %d tl %d bad %d\n", (unsigned long long)pcs->picture_number, sb_index,
                            sb_origin_x, sb_origin_y, P.num_lists, P.temporal_layer_index, bad);
                    for (int l = 0; l < P.num_lists; l++)
                        fprintf(f, "  list %d: ref sa_org (%d,%d) | port hme (%d,%d) sa_org (%d,%d) %dx%d subpel %x\n", l,
                                me->x_search_area_origin[l][0], me->y_search_area_origin[l][0], tr.hme_center_x[l], tr.hme_center_y[l],
                                tr.sa_ox[l], tr.sa_oy[l], tr.sa_w[l], tr.sa_h[l], tr.subpel_enable[l]);
                    {   /* direct 64x64 sub-sampled SAD at the reference's list-0 MV (if full-pel) */
                        int mx = refres[0].x_mv_l0, my = refres[0].y_mv_l0;
                        const uint8_t *sp = cur[0].buf + (cur[0].origin_y + sb_origin_y) * cur[0].stride + cur[0].origin_x + sb_origin_x;
                        const uint8_t *rp = r0[0].buf + (r0[0].origin_y + (int)sb_origin_y + (my >> 2)) * r0[0].stride + r0[0].origin_x + (int)sb_origin_x + (mx >> 2);
                        unsigned sad = 0;
                        for (int y = 0; y < 64; y += 2) for (int x = 0; x < 64; x++) sad += abs(sp[y * cur[0].stride + x] - rp[y * r0[0].stride + x]);
                        for (int dx = -3; dx <= 6; dx++) {
                            unsigned s2 = 0;
                            for (int y = 0; y < 64; y += 2) for (int x = 0; x < 64; x++) s2 += abs(sp[y * cur[0].stride + x] - rp[y * r0[0].stride + x + dx]);
                            fprintf(f, "   dx %d -> %u\n", dx, 2 * s2);
                        }
                        fprintf(f, "  planes: cur stride %d ox %d w %d | ref stride %d ox %d w %d h %d ph %d\n", cur[0].stride, cur[0].origin_x, cur[0].width, r0[0].stride, r0[0].origin_x, r0[0].width, r0[0].height, r0[0].padded_height);
                        fprintf(f, "  direct sad@ref mv0 (%d,%d) = %u ; fullpel port sad[0]=%u mv=(%d,%d)\n", mx, my, 2 * sad, tr.fullpel_sad[0][0],
                                (int16_t)(tr.fullpel_mv[0][0] & 0xffff), (int16_t)(tr.fullpel_mv[0][0] >> 16));
                    }
                    for (int i = 0; i < 85; i++)
                        if (memcmp(&mine[i], &refres[i], sizeof(mine[i])))
                            fprintf(f, "  pu %d ref mv0(%d,%d) mv1(%d,%d) n%d d[%u:%u %u:%u %u:%u] | port mv0(%d,%d) mv1(%d,%d) n%d d[%u:%u %u:%u %u:%u]\n", i,
                                    refres[i].x_mv_l0, refres[i].y_mv_l0, refres[i].x_mv_l1, refres[i].y_mv_l1, refres[i].total_me_candidate_index,
                                    refres[i].cand[0].distortion, refres[i].cand[0].direction, refres[i].cand[1].distortion, refres[i].cand[1].direction,
                                    refres[i].cand[2].distortion, refres[i].cand[2].direction,
                                    mine[i].x_mv_l0, mine[i].y_mv_l0, mine[i].x_mv_l1, mine[i].y_mv_l1, mine[i].total_me_candidate_index,
                                    mine[i].cand[0].distortion, mine[i].cand[0].direction, mine[i].cand[1].distortion, mine[i].cand[1].direction,
                                    mine[i].cand[2].distortion, mine[i].cand[2].direction);
                    fclose(f);
                }
            }
        }
        pthread_mutex_unlock(&g_mu);
        return rc;
    }

    /* dump mode: one file per picture, written when its last SB arrives */
    if ((int)pcs->picture_number < g_dump_from) return rc;
    pthread_mutex_lock(&g_mu);
    DumpPic *dp = NULL;
    for (int i = 0; i < MAX_TRACKED; i++) if (g_dp[i].used && g_dp[i].pic == pcs->picture_number) dp = &g_dp[i];
    if (!dp && g_dumped < g_max_dump_pics) {
        for (int i = 0; i < MAX_TRACKED && !dp; i++) if (!g_dp[i].used) dp = &g_dp[i];
        if (dp) {
            memset(dp, 0, sizeof(*dp));
            dp->used = 1; dp->pic = pcs->picture_number; g_dumped++;
            dp->n_sb = ((P.luma_width + 63) / 64) * ((P.luma_height + 63) / 64);
            dp->res  = (SvtB200MeCuResult *)calloc((size_t)dp->n_sb * 85, sizeof(SvtB200MeCuResult));
            char path[600]; snprintf(path, sizeof(path), "%s/me_pic%04llu.bin", g_dir, (unsigned long long)pcs->picture_number);
            dp->f = fopen(path, "wb");
            uint32_t magic[4] = {0x42305653u /* 'SV0B' */, 1, (uint32_t)sizeof(P), (uint32_t)dp->n_sb};
            fwrite(magic, sizeof(magic), 1, dp->f);
            fwrite(&P, sizeof(P), 1, dp->f);
            for (int i = 0; i < 3; i++) write_plane(dp->f, &cur[i]);
            for (int i = 0; i < 3; i++) write_plane(dp->f, &r0[i]);
            if (P.num_lists == 2) for (int i = 0; i < 3; i++) write_plane(dp->f, &r1[i]);
        }
    }
    if (dp && dp->f) {
        memcpy(dp->res + (size_t)sb_index * 85, refres, sizeof(refres));
        if (++dp->done == dp->n_sb) {
            fwrite(dp->res, sizeof(SvtB200MeCuResult), (size_t)dp->n_sb * 85, dp->f);
            fclose(dp->f); dp->f = NULL; free(dp->res); dp->res = NULL;
        }
    }
    pthread_mutex_unlock(&g_mu);
    return rc;
}


/* ---------------------------------------------------------------------------------------------------------------
 * S2: eb_vp9_loop_filter_frame -- run the oracle port (oracle/lf_port.c) on a copy and compare the planes
 * ------------------------------------------------------------------------------------------------------------- */
void __real_eb_vp9_loop_filter_frame(VP9_COMMON *cm, MACROBLOCKD *xd, int frame_filter_level, int y_only, int partial_frame);
void svtb200_oracle_deblock_picture(const SvtB200LfParams *p, uint8_t *y, int y_stride, uint8_t *u, uint8_t *v, int uv_stride,
                                    SvtB200LfMask *lfm, int lfm_stride, const SvtB200LfThresh thr[64]);
static long g_lf_total, g_lf_bad, g_lf_dumped;

static void write_lf_summary(void) {
    char path[600];
    snprintf(path, sizeof(path), "%s/summary_lf.txt", g_dir);
    FILE *f = fopen(path, "w");
    if (!f) return;
    fprintf(f, "lf_pictures %ld\nlf_bad %ld\nlf_dumped %ld\n", g_lf_total, g_lf_bad, g_lf_dumped);
    fclose(f);
}

void __wrap_eb_vp9_loop_filter_frame(VP9_COMMON *cm, MACROBLOCKD *xd, int frame_filter_level, int y_only, int partial_frame) {
    pthread_mutex_lock(&g_mu);
    init_once();
    int mode = g_mode;
    pthread_mutex_unlock(&g_mu);
    if (!mode || !frame_filter_level || partial_frame || sizeof(LOOP_FILTER_MASK) != sizeof(SvtB200LfMask)) {
        __real_eb_vp9_loop_filter_frame(cm, xd, frame_filter_level, y_only, partial_frame);
        return;
    }
    SvtB200LfParams P; SvtB200LfThresh thr[64];
    svtb200_glue_lf(cm, frame_filter_level, y_only, &P, thr);
    const int W = P.width, H = P.height, ys = xd->plane[0].dst.stride, cs = xd->plane[1].dst.stride;
    /* the reference advances xd->plane[].dst.buf while filtering: remember the picture origins */
    uint8_t *const ry = xd->plane[0].dst.buf, *const ru = xd->plane[1].dst.buf, *const rv = xd->plane[2].dst.buf;
    const int sb_rows = (cm->mi_rows + 7) >> 3, n_lfm = cm->lf.lfm_stride * sb_rows;
    uint8_t *cy = malloc((size_t)W * H), *cu = malloc((size_t)W * H / 4), *cv = malloc((size_t)W * H / 4);
    SvtB200LfMask *lfm = malloc(sizeof(SvtB200LfMask) * n_lfm);
    memcpy(lfm, cm->lf.lfm, sizeof(SvtB200LfMask) * n_lfm);
    for (int r = 0; r < H; r++) memcpy(cy + (size_t)r * W, ry + (size_t)r * ys, W);
    for (int r = 0; r < H / 2; r++) { memcpy(cu + (size_t)r * (W / 2), ru + (size_t)r * cs, W / 2); memcpy(cv + (size_t)r * (W / 2), rv + (size_t)r * cs, W / 2); }
    FILE *df = NULL;
    if (mode == 2) {
        pthread_mutex_lock(&g_mu);
        if (g_lf_dumped < g_max_dump_pics) {
            char path[600]; snprintf(path, sizeof(path), "%s/lf_pic%04ld.bin", g_dir, g_lf_dumped++);
            df = fopen(path, "wb");
        }
        pthread_mutex_unlock(&g_mu);
        if (df) { /* inputs: params, thresholds, masks (before adjustment), planes */
            uint32_t hdr[4] = {0x4c305653u, 1, (uint32_t)cm->lf.lfm_stride, (uint32_t)n_lfm};
            fwrite(hdr, sizeof(hdr), 1, df); fwrite(&P, sizeof(P), 1, df); fwrite(thr, sizeof(thr), 1, df);
            fwrite(lfm, sizeof(SvtB200LfMask), n_lfm, df);
            fwrite(cy, 1, (size_t)W * H, df); fwrite(cu, 1, (size_t)W * H / 4, df); fwrite(cv, 1, (size_t)W * H / 4, df);
        }
    }
    __real_eb_vp9_loop_filter_frame(cm, xd, frame_filter_level, y_only, partial_frame);
    svtb200_oracle_deblock_picture(&P, cy, W, cu, cv, W / 2, lfm, cm->lf.lfm_stride, thr);
    long bad = 0;
    for (int r = 0; r < H; r++) bad += memcmp(cy + (size_t)r * W, ry + (size_t)r * ys, W) != 0;
    if (!y_only)
        for (int r = 0; r < H / 2; r++) {
            bad += memcmp(cu + (size_t)r * (W / 2), ru + (size_t)r * cs, W / 2) != 0;
            bad += memcmp(cv + (size_t)r * (W / 2), rv + (size_t)r * cs, W / 2) != 0;
        }
    if (bad && g_printed < 40) {
        char path[600]; snprintf(path, sizeof(path), "%s/mismatch_lf.txt", g_dir);
        FILE *f = fopen(path, "a");
        if (f) {
            int shown = 0;
            fprintf(f, "lf picture %ld level %d y_only %d mi %dx%d bad rows %ld\n", g_lf_total, frame_filter_level, y_only, cm->mi_cols, cm->mi_rows, bad);
            for (int r = 0; r < H && shown < 12; r++) for (int x = 0; x < W && shown < 12; x++)
                if (cy[(size_t)r * W + x] != ry[(size_t)r * ys + x]) { fprintf(f, "  Y (%d,%d) ref %d port %d\n", x, r, ry[(size_t)r * ys + x], cy[(size_t)r * W + x]); shown++; }
            shown = 0;
            for (int r = 0; r < H / 2 && shown < 8; r++) for (int x = 0; x < W / 2 && shown < 8; x++)
                if (cu[(size_t)r * (W / 2) + x] != ru[(size_t)r * cs + x]) { fprintf(f, "  U (%d,%d) ref %d port %d\n", x, r, ru[(size_t)r * cs + x], cu[(size_t)r * (W / 2) + x]); shown++; }
            fclose(f); g_printed++;
        }
    }
    if (df) { /* expected outputs = the reference's own filtered planes */
        for (int r = 0; r < H; r++) fwrite(ry + (size_t)r * ys, 1, W, df);
        for (int r = 0; r < H / 2; r++) fwrite(ru + (size_t)r * cs, 1, W / 2, df);
        for (int r = 0; r < H / 2; r++) fwrite(rv + (size_t)r * cs, 1, W / 2, df);
        fclose(df);
    }
    pthread_mutex_lock(&g_mu);
    g_lf_total++; if (bad) g_lf_bad++;
    write_lf_summary();
    pthread_mutex_unlock(&g_mu);
    free(cy); free(cu); free(cv); free(lfm);
}
