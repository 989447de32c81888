/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle ("port") for the motion-estimation hot path.
 *
 * A scalar restatement of what the reference's motion_estimate_sb() computes for one 64x64
 * super-block, written from the behaviour of the reference C path.  It is the checker that the
 * CUDA implementation is compared against (tests/, __graft_entry__.smoke(), bench.py's
 * cpu_baseline leg); nothing in the product path may call it.
 *
 * Parity pinning: this file is validated against the UNMODIFIED reference (oracle/_ref, built
 * from /root/reference by oracle/build_ref.sh): inside real encodes by the seam check (oracle/ref_seam_check.c,
 * oracle/run_seam_check.py), on whole pictures through the path driver (oracle/ref_path_driver.c,
 * tests/test_ref_path_driver.py), and through the fixtures under tests/golden/ that were dumped from the
 * reference encoder by the seam check's dump mode (oracle/make_golden.py; tests/test_me_oracle.py).
 *
 * Reference functions followed (paths relative to Source/Lib):
 *   motion_estimate_sb                Codec/EbMotionEstimation.c:4524-5305
 *   test_search_area_bounds           Codec/EbMotionEstimation.c:4260-4518
 *   hme_level0 / single_hme_quadrant_level0 / hme_level1 / hme_level2
 *                                     Codec/EbMotionEstimation.c:2872 / :2717 / :3058 / :3183
 *   eb_vp9_sad_loop_kernel            C_DEFAULT/EbComputeSAD_C.c:132-169
 *   check_zero_zero_center            Codec/EbMotionEstimation.c:3758-3837
 *   full_pel_search_sb + 8x8/16x16/32x32/64x64 SAD kernels
 *                                     Codec/EbMotionEstimation.c:951-990, C_DEFAULT/EbComputeSAD_C.c:172-375,
 *                                     C_DEFAULT/EbMeSadCalculation_C.c:16-113
 *   su_pel_enable                     Codec/EbMotionEstimation.c:3839-4258
 *   interpolate_search_region_avc     Codec/EbMotionEstimation.c:992-1064, C_DEFAULT/EbAvcStyleMcp_C.c:25-72
 *   half_pel_search_sb / pu_half_pel_refinement
 *                                     Codec/EbMotionEstimation.c:1565-1702 / :1076-1559
 *   quarter_pel_search_sb / pu_quarter_pel_refinement_on_the_fly / set_quarter_pel_refinement_inputs_on_the_fly
 *                                     Codec/EbMotionEstimation.c:2471-2715 / :1731-2288 / :2290-2465
 *   bi_prediction_search / bi_prediction_compensation / bi_pred_averging / select_buffer /
 *   quarter_pel_compensation / sort3_elements
 *                                     Codec/EbMotionEstimation.c:3310-3756
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/svtb200.h"
#include "oracle.h"

#define SB 64
#define MAX_SAD_VALUE (64 * 64 * 255)

enum { M_SUB_SAD = 0, M_FULL_SAD = 1, M_SSD = 2 }; /* EbDefinitions.h:664-666 */

/* z-order tables (EbMotionEstimation.c:51-54): raster PU index -> internal index */
static const uint8_t tab16[16] = {0, 1, 4, 5, 2, 3, 6, 7, 8, 9, 12, 13, 10, 11, 14, 15};
static const uint8_t tab8[64]  = {0,  1,  4,  5,  16, 17, 20, 21, 2,  3,  6,  7,  18, 19, 22, 23,
                                  8,  9,  12, 13, 24, 25, 28, 29, 10, 11, 14, 15, 26, 27, 30, 31,
                                  32, 33, 36, 37, 48, 49, 52, 53, 34, 35, 38, 39, 50, 51, 54, 55,
                                  40, 41, 44, 45, 56, 57, 60, 61, 42, 43, 46, 47, 58, 59, 62, 63};

static inline int16_t mvx(uint32_t mv) { return (int16_t)(mv & 0xffff); }
static inline int16_t mvy(uint32_t mv) { return (int16_t)(mv >> 16); }
static inline uint32_t mkmv(int x, int y) { return ((uint32_t)(uint16_t)y << 16) | (uint16_t)x; }

static inline const uint8_t *pix(const SvtB200Plane *p, int x, int y) {
    return p->buf + (ptrdiff_t)(p->origin_y + y) * p->stride + p->origin_x + x;
}

static uint32_t sad_wh(const uint8_t *a, int as, const uint8_t *b, int bs, int h, int w) {
    uint32_t s = 0;
    for (int y = 0; y < h; y++, a += as, b += bs)
        for (int x = 0; x < w; x++) s += (uint32_t)abs((int)a[x] - (int)b[x]);
    return s;
}
static uint32_t ssd_wh(const uint8_t *a, int as, const uint8_t *b, int bs, int h, int w) {
    uint32_t s = 0;
    for (int y = 0; y < h; y++, a += as, b += bs)
        for (int x = 0; x < w; x++) {
            int d = (int)a[x] - (int)b[x];
            s += (uint32_t)(d * d);
        }
    return s;
}
static uint32_t avg_sad_wh(const uint8_t *s, int ss, const uint8_t *a, int as, const uint8_t *b, int bs, int h, int w) {
    uint32_t r = 0;
    for (int y = 0; y < h; y++, s += ss, a += as, b += bs)
        for (int x = 0; x < w; x++) r += (uint32_t)abs((int)s[x] - ((a[x] + b[x] + 1) >> 1));
    return r;
}
static uint32_t avg_ssd_wh(const uint8_t *s, int ss, const uint8_t *a, int as, const uint8_t *b, int bs, int h, int w) {
    uint32_t r = 0;
    for (int y = 0; y < h; y++, s += ss, a += as, b += bs)
        for (int x = 0; x < w; x++) {
            int d = (int)s[x] - ((a[x] + b[x] + 1) >> 1);
            r += (uint32_t)(d * d);
        }
    return r;
}

/* clip a search window [org, org+size) so the block stays on the padded reference
 * (the four "Correct the ... edge" statements that recur in every search set-up) */
static void clip_window(int origin, int16_t *org, int16_t *size, int pad, int pic_dim) {
    int16_t o = *org, s = *size;
    /* left/top: the origin is moved first, so the reference's follow-up width test
     * ((origin + o) < -pad) is never true any more: the size is NOT reduced on this side */
    if (origin + o < -pad) o = (int16_t)(-pad - origin);
    /* right/bottom */
    if (origin + o > pic_dim - 1) o = (int16_t)(o - ((origin + o) - (pic_dim - 1)));
    if (origin + o + s > pic_dim) {
        int v = s - ((origin + o + s) - pic_dim);
        s = (int16_t)(v < 1 ? 1 : v);
    }
    *org = o;
    *size = s;
}
static int16_t clip_center(int origin, int16_t c, int pad, int pic_dim) {
    if (origin + c < -pad) c = (int16_t)(-pad - origin);
    if (origin + c > pic_dim - 1) c = (int16_t)(c - ((origin + c) - (pic_dim - 1)));
    return c;
}

/* eb_vp9_sad_loop_kernel: exhaustive search, first minimum in raster order wins. Outputs are
 * left untouched when no position beats 0xffffff. */
static void sad_loop(const uint8_t *src, int src_stride, const uint8_t *ref, int ref_stride, int h, int w,
                     uint64_t *best, int16_t *xc, int16_t *yc, int ref_row_step, int saw, int sah) {
    *best = 0xffffff;
    for (int ys = 0; ys < sah; ys++, ref += ref_row_step)
        for (int xs = 0; xs < saw; xs++) {
            uint32_t s = sad_wh(src, src_stride, ref + xs, ref_stride, h, w);
            if (s < *best) {
                *best = s;
                *xc = (int16_t)xs;
                *yc = (int16_t)ys;
            }
        }
}

typedef struct {
    const SvtB200MeParams *P;
    const SvtB200Plane    *cur[3];    /* full, quarter, sixteenth of the current picture */
    const SvtB200Plane    *ref[2][3]; /* per list */
    int                    ox, oy, sbw, sbh;
    const uint8_t         *src; /* source SB top-left in the padded full plane */
    int                    ss;  /* its stride */
    uint32_t               best_sad[2][85], best_mv[2][85], best_ssd[2][85], bipred_sad[85];
    uint8_t                dir[85];
    int16_t                sa_ox[2], sa_oy[2]; /* final search-area origin per list */
    int                    saw[2], sah[2];     /* clipped search area per list */
    /* interpolated planes per list; geometry in interp_region() */
    uint8_t *posb[2], *posh[2], *posj[2];
    int      istride[2];
    int      have_interp[2];
} SbCtx;

/* ---------------------------------------------------------------------------------------------
 * test_search_area_bounds
 * ------------------------------------------------------------------------------------------- */
static uint64_t sb_sad_at(const SbCtx *c, const SvtB200Plane *rp, int dx, int dy) {
    /* 64-wide (sb_width) x sb_height/2 rows (every other row), times two */
    return (uint64_t)sad_wh(c->src, c->ss * 2, pix(rp, c->ox + dx, c->oy + dy), rp->stride * 2, c->sbh >> 1, c->sbw) << 1;
}

static void search_area_bounds(const SbCtx *c, int list, int16_t *xsc, int16_t *ysc) {
    const SvtB200MeParams *P  = c->P;
    const SvtB200Plane    *rp = c->ref[list][0];
    const int              pad = SB - 1, W0 = P->hme_level0_total_search_area_width, H0 = P->hme_level0_total_search_area_height;
    uint64_t cost_o = sb_sad_at(c, rp, 0, 0);
    /* A is measured at the O position (the reference never recomputes its index): ties O */
    uint64_t cost_a = cost_o;
    int16_t  bx = clip_center(c->ox, (int16_t)W0, pad, rp->width), by = clip_center(c->oy, 0, pad, rp->height);
    uint64_t cost_b = sb_sad_at(c, rp, bx, by);
    int16_t  cx = clip_center(c->ox, 0, pad, rp->width), cy = clip_center(c->oy, (int16_t)(0 - H0), pad, rp->height);
    uint64_t cost_c = sb_sad_at(c, rp, cx, cy);
    int16_t  dx = clip_center(c->ox, 0, pad, rp->width), dy = clip_center(c->oy, (int16_t)H0, pad, rp->height);
    uint64_t cost_d = sb_sad_at(c, rp, dx, dy);
    uint64_t cost_direct = 0xFFFFFFFFFFFFFull;
    int16_t  direct_x = 0, direct_y = 0;
    if (list == 1) {
        direct_x = (int16_t)(0 - (mvx(c->best_mv[0][0]) >> 2));
        direct_y = (int16_t)(0 - (mvy(c->best_mv[0][0]) >> 2));
        int16_t ex = clip_center(c->ox, direct_x, pad, rp->width), ey = clip_center(c->oy, direct_y, pad, rp->height);
        cost_direct = sb_sad_at(c, rp, ex, ey);
    }
    uint64_t best = cost_o;
    if (cost_a < best) best = cost_a;
    if (cost_b < best) best = cost_b;
    if (cost_c < best) best = cost_c;
    if (cost_d < best) best = cost_d;
    if (cost_direct < best) best = cost_direct;
    /* tie order O, A, B, C, direct, D; the UNCLIPPED offsets are returned */
    if (best == cost_o) { *xsc = 0; *ysc = 0; }
    else if (best == cost_a) { *xsc = (int16_t)(0 - W0); *ysc = 0; }
    else if (best == cost_b) { *xsc = (int16_t)W0; *ysc = 0; }
    else if (best == cost_c) { *xsc = 0; *ysc = (int16_t)(0 - H0); }
    else if (best == cost_direct) { *xsc = direct_x; *ysc = direct_y; }
    else { *xsc = 0; *ysc = (int16_t)H0; }
}

/* ---------------------------------------------------------------------------------------------
 * HME levels
 * ------------------------------------------------------------------------------------------- */
static void hme_l0(const SbCtx *c, int list, int qx, int qy, int single, int16_t center_x, int16_t center_y,
                   uint64_t *best, int16_t *xc, int16_t *yc) {
    const SvtB200MeParams *P  = c->P;
    const SvtB200Plane    *rp = c->ref[list][2];
    const int ox = c->ox >> 2, oy = c->oy >> 2, sbw = c->sbw >> 2, sbh = c->sbh >> 2;
    int16_t saw, sah, x_org, y_org;
    if (single) {
        saw   = (int16_t)((P->hme_level0_total_search_area_width * P->hme_l0_multiplier_x) / 100);
        sah   = (int16_t)((P->hme_level0_total_search_area_height * P->hme_l0_multiplier_y) / 100);
        x_org = (int16_t)(-(int16_t)(saw >> 1) + center_x);
        y_org = (int16_t)(-(int16_t)(sah >> 1) + center_y);
    } else {
        saw = (int16_t)((P->hme_level0_search_area_in_width[qx] * P->hme_l0_multiplier_x) / 100);
        sah = (int16_t)((P->hme_level0_search_area_in_height[qy] * P->hme_l0_multiplier_y) / 100);
        int16_t xd = center_x, yd = center_y;
        for (int k = qx; k > 0; k--) xd = (int16_t)(xd + (int16_t)((P->hme_level0_search_area_in_width[k - 1] * P->hme_l0_multiplier_x) / 100));
        for (int k = qy; k > 0; k--) yd = (int16_t)(yd + (int16_t)((P->hme_level0_search_area_in_height[k - 1] * P->hme_l0_multiplier_y) / 100));
        x_org = (int16_t)(-(int16_t)(((P->hme_level0_total_search_area_width * P->hme_l0_multiplier_x) / 100) >> 1) + xd);
        y_org = (int16_t)(-(int16_t)(((P->hme_level0_total_search_area_height * P->hme_l0_multiplier_y) / 100) >> 1) + yd);
    }
    clip_window(ox, &x_org, &saw, rp->origin_x - 1, rp->width);
    clip_window(oy, &y_org, &sah, rp->origin_y - 1, rp->height);
    if (single && (saw & 15)) saw = (int16_t)((saw >> 4) << 4);
    /* source: the 1/16 SB, every other row (8 rows of sbw samples) */
    const uint8_t *s16 = pix(c->cur[2], ox, oy);
    sad_loop(s16, c->cur[2]->stride * 2, pix(rp, ox + x_org, oy + y_org), rp->stride * 2, sbh >> 1, sbw, best, xc, yc,
             rp->stride, saw, sah);
    *best *= 2;
    *xc = (int16_t)((*xc + x_org) * 4);
    *yc = (int16_t)((*yc + y_org) * 4);
}

static int16_t widen8(int16_t w) { /* EbMotionEstimation.c:3083-3086 */
    return (int16_t)((w < 8) ? 8 : (w & 7) ? w + (w - ((w >> 3) << 3)) : w);
}

static void hme_l1(const SbCtx *c, int list, int qx, int qy, int16_t l0x, int16_t l0y, uint64_t *best, int16_t *xc,
                   int16_t *yc) {
    const SvtB200MeParams *P  = c->P;
    const SvtB200Plane    *rp = c->ref[list][1];
    const int ox = c->ox >> 1, oy = c->oy >> 1, sbw = c->sbw >> 1, sbh = c->sbh >> 1;
    int16_t saw = widen8((int16_t)P->hme_level1_search_area_in_width[qx]);
    int16_t sah = (int16_t)P->hme_level1_search_area_in_height[qy];
    int16_t x_org = (int16_t)(-(saw >> 1) + l0x), y_org = (int16_t)(-(sah >> 1) + l0y);
    clip_window(ox, &x_org, &saw, rp->origin_x - 1, rp->width);
    clip_window(oy, &y_org, &sah, rp->origin_y - 1, rp->height);
    sad_loop(pix(c->cur[1], ox, oy), c->cur[1]->stride * 2, pix(rp, ox + x_org, oy + y_org), rp->stride * 2, sbh >> 1,
             sbw, best, xc, yc, rp->stride, saw, sah);
    *best *= 2;
    *xc = (int16_t)((*xc + x_org) * 2);
    *yc = (int16_t)((*yc + y_org) * 2);
}

static void hme_l2(const SbCtx *c, int list, int qx, int qy, int16_t l1x, int16_t l1y, uint64_t *best, int16_t *xc,
                   int16_t *yc) {
    const SvtB200MeParams *P  = c->P;
    const SvtB200Plane    *rp = c->ref[list][0];
    int16_t saw = widen8((int16_t)P->hme_level2_search_area_in_width[qx]);
    int16_t sah = (int16_t)P->hme_level2_search_area_in_height[qy];
    int16_t x_org = (int16_t)(-(saw >> 1) + l1x), y_org = (int16_t)(-(sah >> 1) + l1y);
    clip_window(c->ox, &x_org, &saw, SB - 1, rp->width);
    clip_window(c->oy, &y_org, &sah, SB - 1, rp->height);
    sad_loop(c->src, c->ss * 2, pix(rp, c->ox + x_org, c->oy + y_org), rp->stride * 2, c->sbh >> 1, c->sbw, best, xc, yc,
             rp->stride, saw, sah);
    *best *= 2;
    *xc = (int16_t)(*xc + x_org);
    *yc = (int16_t)(*yc + y_org);
}

/* ---------------------------------------------------------------------------------------------
 * full-pel search over the 85 PUs
 * ------------------------------------------------------------------------------------------- */
static void full_pel_search(SbCtx *c, int list) {
    const SvtB200Plane *rp = c->ref[list][0];
    uint32_t *bs = c->best_sad[list], *bm = c->best_mv[list];
    for (int i = 0; i < 85; i++) bs[i] = MAX_SAD_VALUE;
    for (int ys = 0; ys < c->sah[list]; ys++)
        for (int xs = 0; xs < c->saw[list]; xs++) {
            const uint8_t *r  = pix(rp, c->ox + c->sa_ox[list] + xs, c->oy + c->sa_oy[list] + ys);
            const uint32_t mv = mkmv((xs + c->sa_ox[list]) * 4, (ys + c->sa_oy[list]) * 4);
            uint32_t s16[16], s32[4], s64 = 0;
            /* columns beyond the last multiple of 8 go through get_search_point_results()
             * (EbMotionEstimation.c:662-945), which advances the reference offset twice before
             * 16x16 block 12 (:855-856): blocks 12 and 13 are then matched 16 samples further
             * right in the reference.  Reproduced, since both the C and AVX2 builds share it. */
            const int tail = xs >= (c->saw[list] & ~7);
            for (int b16 = 0; b16 < 16; b16++) { /* raster 16x16 blocks */
                const int bx = (b16 & 3) * 16, by = (b16 >> 2) * 16, z16 = tab16[b16];
                const int rskew = (tail && (z16 == 12 || z16 == 13)) ? 16 : 0;
                uint32_t  sum = 0;
                for (int k = 0; k < 4; k++) { /* 8x8 quadrants TL,TR,BL,BR: rows 0,2,4,6 only */
                    const int qx = bx + (k & 1) * 8, qy = by + (k >> 1) * 8;
                    uint32_t  s = sad_wh(c->src + qy * c->ss + qx, c->ss * 2, r + qy * rp->stride + qx + rskew, rp->stride * 2, 4, 8);
                    sum += s;
                    const int idx = 21 + z16 * 4 + k;
                    if (2 * s < bs[idx]) { bs[idx] = 2 * s; bm[idx] = mv; }
                }
                s16[z16] = sum;
                if (2 * sum < bs[5 + z16]) { bs[5 + z16] = 2 * sum; bm[5 + z16] = mv; }
            }
            for (int q = 0; q < 4; q++) {
                s32[q] = s16[4 * q] + s16[4 * q + 1] + s16[4 * q + 2] + s16[4 * q + 3];
                s64 += s32[q];
                if (2 * s32[q] < bs[1 + q]) { bs[1 + q] = 2 * s32[q]; bm[1 + q] = mv; }
            }
            if (2 * s64 < bs[0]) { bs[0] = 2 * s64; bm[0] = mv; }
        }
}

/* ---------------------------------------------------------------------------------------------
 * su_pel_enable (fractional_search_model == 1)
 * ------------------------------------------------------------------------------------------- */
static void su_pel_enable(const SbCtx *c, int list, int *e32, int *e16, int *e8) {
    const uint32_t *bs = c->best_sad[list], *bm = c->best_mv[list];
    int sx = 0, sy = 0;
    uint32_t ss = 0;
    for (int i = 1; i < 5; i++) { sx += mvx(bm[i]); sy += mvy(bm[i]); ss += bs[i]; }
    uint32_t ax = (uint32_t)(sx >> 2), ay = (uint32_t)(sy >> 2);
    const uint32_t mag32 = ax * ax + ay * ay, sad32 = ss >> 2;
    sx = sy = 0; ss = 0;
    for (int i = 5; i < 21; i++) { sx += mvx(bm[i]); sy += mvy(bm[i]); ss += bs[i]; }
    ax = (uint32_t)(sx >> 4); ay = (uint32_t)(sy >> 4);
    const uint32_t mag16 = ax * ax + ay * ay, sad16 = ss >> 4;
    sx = sy = 0; ss = 0;
    for (int i = 21; i < 85; i++) { sx += mvx(bm[i]); sy += mvy(bm[i]); ss += bs[i]; }
    ax = (uint32_t)(sx >> 6); ay = (uint32_t)(sy >> 6);
    const uint32_t mag8 = ax * ax + ay * ay, sad8 = ss >> 6;

    const int tl = c->P->temporal_layer_index;
    const uint32_t th = (tl == 1) ? 32 * 32 : (tl == 2) ? 80 * 80 : 48 * 48;
    const int near32 = mag32 < th, low32 = sad32 < 32 * 32 * 6;
    /* 32x32: on unless (class 1 or 3) per layer */
    if (tl == 0 || tl == 2) *e32 = low32;                       /* classes 0,2 on; 1,3 off */
    else if (tl == 1) *e32 = !(near32 && !low32);               /* only class 1 off */
    else *e32 = !(!near32 && !low32);                           /* only class 3 off */
    /* 16x16: on when SAD is not low (classes 1,3), every layer */
    *e16 = !(sad16 < 16 * 16 * 2);
    /* 8x8: same, except the highest layers switch class 3 off */
    const int near8 = mag8 < th, low8 = sad8 < 8 * 8 * 2;
    if (tl <= 2) *e8 = !low8;
    else *e8 = near8 && !low8;
    (void)mag16;
}

/* ---------------------------------------------------------------------------------------------
 * half-pel planes b, h, j of the search region (interpolate_search_region_avc)
 * R(x,y) = reference sample at the top-left of the clipped search area + (x,y).
 *   posb[x + y*is] = H-filter of R(x-2..x+1, y-2)            x<wasm, y<H+4   ("b" left of R(x, y-2))
 *   posh[x + y*is] = V-filter of R(x-1, y-2..y+1)            x<wasm, y<H+1   ("h" above R(x-1, y))
 *   posj[x + y*is] = V-filter of posb[x + (y..y+3)*is]       x<wasm, y<H+1   ("j" up-left of R(x, y))
 * taps {-2,18,18,-2}, +16 >> 5, clip to u8.
 * ------------------------------------------------------------------------------------------- */
static inline uint8_t clip8(int v) { return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }
static inline uint8_t tap4(int a, int b, int cc, int d) { return clip8((-2 * a + 18 * b + 18 * cc - 2 * d + 16) >> 5); }

static void interp_region(SbCtx *c, int list) {
    const SvtB200Plane *rp = c->ref[list][0];
    const int W = c->saw[list] + SB - 1, H = c->sah[list] + SB - 1;
    const int wasm = (W + 2 + 7) & ~7, is = wasm;
    const uint8_t *R = pix(rp, c->ox + c->sa_ox[list], c->oy + c->sa_oy[list]);
    const int rs = rp->stride;
    free(c->posb[list]); free(c->posh[list]); free(c->posj[list]);
    c->posb[list] = (uint8_t *)malloc((size_t)is * (H + 4));
    c->posh[list] = (uint8_t *)malloc((size_t)is * (H + 1));
    c->posj[list] = (uint8_t *)malloc((size_t)is * (H + 1));
    c->istride[list] = is;
    for (int y = 0; y < H + 4; y++)
        for (int x = 0; x < wasm; x++) {
            const uint8_t *p = R + (ptrdiff_t)(y - 2) * rs + (x - 2);
            c->posb[list][x + y * is] = tap4(p[0], p[1], p[2], p[3]);
        }
    for (int y = 0; y < H + 1; y++)
        for (int x = 0; x < wasm; x++) {
            const uint8_t *p = R + (ptrdiff_t)(y - 2) * rs + (x - 1);
            c->posh[list][x + y * is] = tap4(p[0], p[rs], p[2 * rs], p[3 * rs]);
            const uint8_t *q = c->posb[list] + x + y * is;
            c->posj[list][x + y * is] = tap4(q[0], q[is], q[2 * is], q[3 * is]);
        }
    c->have_interp[list] = 1;
}

/* pointers in the convention of half_pel_search_sb's arguments: pb[x+y*is] = b left of R(x,y),
 * ph = h above R(x,y), pj = j up-left of R(x,y) */
#define PB(c, l) ((c)->posb[l] + 2 * (c)->istride[l])
#define PH(c, l) ((c)->posh[l] + 1)
#define PJ(c, l) ((c)->posj[l])

static uint32_t frac_dist(const SbCtx *c, int method, const uint8_t *src, const uint8_t *r, int rs, int w, int h) {
    if (method == M_SSD) return ssd_wh(src, c->ss, r, rs, h, w);
    if (method == M_SUB_SAD) return sad_wh(src, c->ss * 2, r, rs * 2, h >> 1, w) << 1;
    return sad_wh(src, c->ss, r, rs, h, w);
}

static void half_pel_pu(SbCtx *c, int list, int idx, int px, int py, int w, int h) {
    const SvtB200Plane *rp = c->ref[list][0];
    const int method = c->P->fractional_search_method, is = c->istride[list];
    uint32_t *bsad = &c->best_sad[list][idx], *bmv = &c->best_mv[list][idx], *bssd = &c->best_ssd[list][idx];
    const int16_t x_mv = mvx(*bmv), y_mv = mvy(*bmv);
    const int xs = (x_mv >> 2) - c->sa_ox[list], ys = (y_mv >> 2) - c->sa_oy[list];
    const uint8_t *src = c->src + py * c->ss + px;
    const uint8_t *R   = pix(rp, c->ox + c->sa_ox[list] + px, c->oy + c->sa_oy[list] + py);
    const uint8_t *pb = PB(c, list) + px + py * is, *ph = PH(c, list) + px + py * is, *pj = PJ(c, list) + px + py * is;
    if (method == M_SSD) *bssd = ssd_wh(src, c->ss, R + ys * rp->stride + xs, rp->stride, h, w);
    const int base = xs + is * ys;
    /* order L, R, T, B, TL, TR, BR, BL */
    const uint8_t *cand[8] = {pb + base, pb + base + 1, ph + base, ph + base + is,
                              pj + base, pj + base + 1, pj + base + 1 + is, pj + base + is};
    static const int8_t dx[8] = {-2, 2, 0, 0, -2, 2, 2, -2}, dy[8] = {0, 0, -2, 2, -2, -2, 2, 2};
    uint64_t dist[8];
    for (int k = 0; k < 8; k++) {
        dist[k] = frac_dist(c, method, src, cand[k], is, w, h);
        if (method == M_SSD) {
            if (dist[k] < *bssd) {
                *bsad = sad_wh(src, c->ss, cand[k], is, h, w);
                *bmv  = mkmv(x_mv + dx[k], y_mv + dy[k]);
                *bssd = (uint32_t)dist[k];
            }
        } else if (dist[k] < *bsad) {
            *bsad = (uint32_t)dist[k];
            *bmv  = mkmv(x_mv + dx[k], y_mv + dy[k]);
        }
    }
    /* direction of the smallest half-pel distortion; tie order L,R,T,B,TL,TR,BL,BR.
     * codes: TL 0, T 1, TR 2, R 3, BR 4, B 5, BL 6, L 7 */
    static const uint8_t order[8] = {0, 1, 2, 3, 4, 5, 7, 6};
    static const uint8_t code[8]  = {7, 3, 1, 5, 0, 2, 4, 6}; /* indexed by candidate k */
    uint64_t m = dist[0];
    for (int k = 1; k < 8; k++) if (dist[k] < m) m = dist[k];
    for (int t = 0; t < 8; t++) if (dist[order[t]] == m) { c->dir[idx] = code[order[t]]; break; }
}

/* quarter-pel: two source buffers per candidate, averaged on the fly.
 * kind: 0 = integer plane A, 1 = b, 2 = h, 3 = j; (ox,oy) offsets in samples */
typedef struct { int8_t kind, ox, oy; } QSrc;
static const QSrc qtab[4][8][2] = {
    /* best MV on a full-pel position */
    {{{1, 0, 0}, {0, 0, 0}}, {{0, 0, 0}, {1, 1, 0}}, {{2, 0, 0}, {0, 0, 0}}, {{0, 0, 0}, {2, 0, 1}},
     {{1, 0, 0}, {2, 0, 0}}, {{2, 0, 0}, {1, 1, 0}}, {{2, 0, 1}, {1, 1, 0}}, {{1, 0, 0}, {2, 0, 1}}},
    /* half-pel horizontally */
    {{{0, -1, 0}, {1, 0, 0}}, {{1, 0, 0}, {0, 0, 0}}, {{3, 0, 0}, {1, 0, 0}}, {{1, 0, 0}, {3, 0, 1}},
     {{2, -1, 0}, {1, 0, 0}}, {{1, 0, 0}, {2, 0, 0}}, {{1, 0, 0}, {2, 0, 1}}, {{2, -1, 1}, {1, 0, 0}}},
    /* half-pel vertically */
    {{{3, 0, 0}, {2, 0, 0}}, {{2, 0, 0}, {3, 1, 0}}, {{0, 0, -1}, {2, 0, 0}}, {{2, 0, 0}, {0, 0, 0}},
     {{1, 0, -1}, {2, 0, 0}}, {{2, 0, 0}, {1, 1, -1}}, {{2, 0, 0}, {1, 1, 0}}, {{1, 0, 0}, {2, 0, 0}}},
    /* half-pel diagonally */
    {{{2, -1, 0}, {3, 0, 0}}, {{3, 0, 0}, {2, 0, 0}}, {{1, 0, -1}, {3, 0, 0}}, {{3, 0, 0}, {1, 0, 0}},
     {{2, -1, 0}, {1, 0, -1}}, {{1, 0, -1}, {2, 0, 0}}, {{1, 0, 0}, {2, 0, 0}}, {{2, -1, 0}, {1, 0, 0}}},
};

static const uint8_t *qsrc_ptr(const SbCtx *c, int list, QSrc q, int x, int y, int *stride) {
    const SvtB200Plane *rp = c->ref[list][0];
    const int is = c->istride[list];
    x += q.ox; y += q.oy;
    switch (q.kind) {
    case 0: *stride = rp->stride; return pix(rp, c->ox + c->sa_ox[list] + x, c->oy + c->sa_oy[list] + y);
    case 1: *stride = is; return PB(c, list) + x + y * is;
    case 2: *stride = is; return PH(c, list) + x + y * is;
    default: *stride = is; return PJ(c, list) + x + y * is;
    }
}

static void quarter_pel_pu(SbCtx *c, int list, int idx, int px, int py, int w, int h) {
    const int method = c->P->fractional_search_method;
    uint32_t *bsad = &c->best_sad[list][idx], *bmv = &c->best_mv[list][idx], *bssd = &c->best_ssd[list][idx];
    const int16_t x_mv = mvx(*bmv), y_mv = mvy(*bmv);
    const int xs = ((x_mv + 2) >> 2) - c->sa_ox[list], ys = ((y_mv + 2) >> 2) - c->sa_oy[list];
    const int phase = (y_mv & 2) + ((x_mv & 2) >> 1);
    /* the three quarter positions next to the winning half-pel direction; when the current best
     * is itself a half-pel position the direction is mirrored */
    const int d = phase ? (c->dir[idx] + 4) & 7 : c->dir[idx];
    /* candidate order L,R,T,B,TL,TR,BR,BL -> ring position (TL0 T1 TR2 R3 BR4 B5 BL6 L7) */
    static const uint8_t ring[8] = {7, 3, 1, 5, 0, 2, 4, 6};
    static const int8_t dx[8] = {-1, 1, 0, 0, -1, 1, 1, -1}, dy[8] = {0, 0, -1, 1, -1, -1, 1, 1};
    const uint8_t *src = c->src + py * c->ss + px; /* == sb_buffer copy of the same samples */
    for (int k = 0; k < 8; k++) {
        const int diff = (ring[k] - d) & 7;
        if (!(diff == 0 || diff == 1 || diff == 7)) continue;
        int s1, s2;
        const uint8_t *b1 = qsrc_ptr(c, list, qtab[phase][k][0], xs + px, ys + py, &s1);
        const uint8_t *b2 = qsrc_ptr(c, list, qtab[phase][k][1], xs + px, ys + py, &s2);
        uint64_t dist;
        if (method == M_SSD) dist = avg_ssd_wh(src, c->ss, b1, s1, b2, s2, h, w);
        else if (method == M_SUB_SAD) dist = (uint64_t)avg_sad_wh(src, c->ss * 2, b1, s1 * 2, b2, s2 * 2, h >> 1, w) << 1;
        else dist = avg_sad_wh(src, c->ss, b1, s1, b2, s2, h, w);
        if (method == M_SSD) {
            if (dist < *bssd) {
                *bsad = avg_sad_wh(src, c->ss, b1, s1, b2, s2, h, w);
                *bmv  = mkmv(x_mv + dx[k], y_mv + dy[k]);
                *bssd = (uint32_t)dist;
            }
        } else if (dist < *bsad) {
            *bsad = (uint32_t)dist;
            *bmv  = mkmv(x_mv + dx[k], y_mv + dy[k]);
        }
    }
}

static void subpel_search(SbCtx *c, int list, int e32, int e16, int e8, int eq) {
    const SvtB200MeParams *P = c->P;
    const int dis8 = P->cu8x8_mode == 1;
    e16 = e16 && P->cu16x16_mode == 0;
    /* half-pel */
    if (P->fractional_search64x64) half_pel_pu(c, list, 0, 0, 0, 64, 64);
    if (e32) for (int i = 0; i < 4; i++) half_pel_pu(c, list, 1 + i, (i & 1) * 32, (i >> 1) * 32, 32, 32);
    if (e16) for (int i = 0; i < 16; i++) half_pel_pu(c, list, 5 + tab16[i], (i & 3) * 16, (i >> 2) * 16, 16, 16);
    if (e8 && !dis8) for (int i = 0; i < 64; i++) half_pel_pu(c, list, 21 + tab8[i], (i & 7) * 8, (i >> 3) * 8, 8, 8);
    /* quarter-pel (the 64x64 PU is refined on its top-left 32x32 only, as the reference does) */
    if (P->fractional_search64x64) quarter_pel_pu(c, list, 0, 0, 0, 32, 32);
    if (eq && e32) for (int i = 0; i < 4; i++) quarter_pel_pu(c, list, 1 + i, (i & 1) * 32, (i >> 1) * 32, 32, 32);
    if (eq && e16) for (int i = 0; i < 16; i++) quarter_pel_pu(c, list, 5 + tab16[i], (i & 3) * 16, (i >> 2) * 16, 16, 16);
    if (eq && e8 && !dis8) for (int i = 0; i < 64; i++) quarter_pel_pu(c, list, 21 + tab8[i], (i & 7) * 8, (i >> 3) * 8, 8, 8);
}

/* ---------------------------------------------------------------------------------------------
 * bi-prediction: SAD(src, avg(pred L0, pred L1)) at the two uni-pred winners
 * ------------------------------------------------------------------------------------------- */
static const uint8_t *bipred_pred(const SbCtx *c, int list, uint32_t mv, int px, int py, int w, int h, uint8_t *tmp,
                                  int *stride) {
    const SvtB200Plane *rp = c->ref[list][0];
    const int is = c->istride[list];
    const int16_t x = mvx(mv), y = mvy(mv);
    const int xs = (x >> 2) - c->sa_ox[list], ys = (y >> 2) - c->sa_oy[list];
    const int frac = (x & 3) + ((y & 3) << 2);
    /* planes anchored at the integer sample I = R(xs,ys): b RIGHT of I, h BELOW I, j down-right */
    const uint8_t *A = pix(rp, c->ox + c->sa_ox[list] + xs + px, c->oy + c->sa_oy[list] + ys + py);
    const int as = rp->stride;
    const uint8_t *b = c->posb[list] + (xs + 1 + px) + (ys + 2 + py) * is;
    const uint8_t *hh = c->posh[list] + (xs + 1 + px) + (ys + 1 + py) * is;
    const uint8_t *j = c->posj[list] + (xs + 1 + px) + (ys + 1 + py) * is;
    const uint8_t *p1 = 0, *p2 = 0;
    int s1 = is, s2 = is;
    switch (frac) {
    case 0: *stride = as; return A;
    case 2: *stride = is; return b;
    case 8: *stride = is; return hh;
    case 10: *stride = is; return j;
    case 1: p1 = A; s1 = as; p2 = b; break;
    case 3: p1 = b; p2 = A + 1; s2 = as; break;
    case 4: p1 = A; s1 = as; p2 = hh; break;
    case 5: p1 = b; p2 = hh; break;
    case 6: p1 = b; p2 = j; break;
    case 7: p1 = b; p2 = hh + 1; break;
    case 9: p1 = hh; p2 = j; break;
    case 11: p1 = j; p2 = hh + 1; break;
    case 12: p1 = hh; p2 = A + as; s2 = as; break;
    case 13: p1 = hh; p2 = b + is; break;
    case 14: p1 = j; p2 = b + is; break;
    default: p1 = hh + 1; p2 = b + is; break; /* 15 */
    }
    for (int yy = 0; yy < h; yy++)
        for (int xx = 0; xx < w; xx++) tmp[yy * SB + xx] = (uint8_t)((p1[yy * s1 + xx] + p2[yy * s2 + xx] + 1) >> 1);
    *stride = SB;
    return tmp;
}

static uint32_t bipred_sad(const SbCtx *c, int n_idx, int px, int py, int w, int h) {
    uint8_t t0[SB * SB], t1[SB * SB];
    int s0, s1;
    /* fractional planes only exist for a list whose sub-pel search ran; a full-pel MV needs none */
    const uint8_t *p0 = bipred_pred(c, 0, c->best_mv[0][n_idx], px, py, w, h, t0, &s0);
    const uint8_t *p1 = bipred_pred(c, 1, c->best_mv[1][n_idx], px, py, w, h, t1, &s1);
    const uint8_t *src = c->src + py * c->ss + px;
    if (c->P->fractional_search_method == M_SUB_SAD) return avg_sad_wh(src, c->ss * 2, p0, s0 * 2, p1, s1 * 2, h >> 1, w) << 1;
    return avg_sad_wh(src, c->ss, p0, s0, p1, s1, h, w);
}

/* ---------------------------------------------------------------------------------------------
 * one super-block
 * ------------------------------------------------------------------------------------------- */
void svtb200_oracle_me_sb(const SvtB200MeParams *P, const SvtB200Plane cur[3], const SvtB200Plane ref0[3],
                          const SvtB200Plane ref1[3], int sb_x, int sb_y, SvtB200MeCuResult out[85],
                          uint32_t *rcme_distortion, SvtB200OracleSbTrace *trace) {
    SbCtx *c = (SbCtx *)calloc(1, sizeof(SbCtx));
    c->P = P;
    for (int i = 0; i < 3; i++) { c->cur[i] = &cur[i]; c->ref[0][i] = &ref0[i]; c->ref[1][i] = ref1 ? &ref1[i] : &ref0[i]; }
    c->ox = sb_x * SB; c->oy = sb_y * SB;
    c->sbw = (cur[0].width - c->ox) < SB ? cur[0].width - c->ox : SB;
    c->sbh = (cur[0].height - c->oy) < SB ? cur[0].height - c->oy : SB;
    c->src = pix(&cur[0], c->ox, c->oy);
    c->ss  = cur[0].stride;
    const int nl = P->num_lists;
    const int pic_w = P->luma_width, pic_h = P->luma_height;

    /* HME state survives from list 0 to list 1 (the quadrant counters are not reset) */
    int16_t  l0x[2][2] = {{0}}, l0y[2][2] = {{0}}, l1x[2][2] = {{0}}, l1y[2][2] = {{0}}, l2x[2][2] = {{0}}, l2y[2][2] = {{0}};
    uint64_t l0s[2][2] = {{0}}, l1s[2][2] = {{0}}, l2s[2][2] = {{0}};
    int      srh = 0; /* the reference's function-scope quadrant row counter (never reset per list) */
    const int nqx = P->number_hme_search_region_in_width, nqy = P->number_hme_search_region_in_height;
    int16_t  hme_x = 0, hme_y = 0;

    for (int list = 0; list < nl; list++) {
        int16_t xsc = 0, ysc = 0;
        if (P->temporal_layer_index > 0 || list == 0) {
            search_area_bounds(c, list, &xsc, &ysc);
            if (P->enable_hme_flag && c->sbh == SB) {
                if (srh < nqy) { /* centre initialisation loop, EbMotionEstimation.c:4668-4701 */
                    for (int qy = 0; qy < nqy; qy++)
                        for (int qx = 0; qx < nqx; qx++) {
                            l0x[qx][qy] = (int16_t)(xsc >> 2); l0y[qx][qy] = (int16_t)(ysc >> 2);
                            l1x[qx][qy] = (int16_t)(xsc >> 1); l1y[qx][qy] = (int16_t)(ysc >> 1);
                            l2x[qx][qy] = xsc; l2y[qx][qy] = ysc;
                        }
                    srh = nqy;
                }
                const int single = P->single_hme_quadrant && !P->enable_hme_level1 && !P->enable_hme_level2;
                /* the single-quadrant L0 call zeroes the counters and nothing advances them again,
                 * so the centres ARE re-initialised for list 1 in that mode (and only in that mode) */
                if (P->enable_hme_level0 && single) srh = 0;
                if (P->enable_hme_level0) {
                    if (single) hme_l0(c, list, 0, 0, 1, (int16_t)(xsc >> 2), (int16_t)(ysc >> 2), &l0s[0][0], &l0x[0][0], &l0y[0][0]);
                    else
                        for (int qy = 0; qy < nqy; qy++)
                            for (int qx = 0; qx < nqx; qx++)
                                hme_l0(c, list, qx, qy, 0, (int16_t)(xsc >> 2), (int16_t)(ysc >> 2), &l0s[qx][qy], &l0x[qx][qy], &l0y[qx][qy]);
                }
                if (P->enable_hme_level1)
                    for (int qy = 0; qy < nqy; qy++)
                        for (int qx = 0; qx < nqx; qx++)
                            hme_l1(c, list, qx, qy, (int16_t)(l0x[qx][qy] >> 1), (int16_t)(l0y[qx][qy] >> 1), &l1s[qx][qy], &l1x[qx][qy], &l1y[qx][qy]);
                if (P->enable_hme_level2)
                    for (int qy = 0; qy < nqy; qy++)
                        for (int qx = 0; qx < nqx; qx++)
                            hme_l2(c, list, qx, qy, l1x[qx][qy], l1y[qx][qy], &l2s[qx][qy], &l2x[qx][qy], &l2y[qx][qy]);

                /* best quadrant of the deepest enabled level; scan order [0][0],[1][0],[0][1],[1][1] */
                int16_t (*fx)[2] = 0, (*fy)[2] = 0;
                uint64_t (*fs)[2] = 0;
                if (P->enable_hme_level2) { fx = l2x; fy = l2y; fs = l2s; }
                else if (P->enable_hme_level1) { fx = l1x; fy = l1y; fs = l1s; }
                else if (P->enable_hme_level0) { fx = l0x; fy = l0y; fs = l0s; }
                if (fs) {
                    hme_x = fx[0][0]; hme_y = fy[0][0];
                    uint64_t best = fs[0][0];
                    if (!(single && fs == l0s))
                        for (int qy = 0; qy < nqy; qy++)
                            for (int qx = (qy == 0) ? 1 : 0; qx < nqx; qx++)
                                if (fs[qx][qy] < best) { best = fs[qx][qy]; hme_x = fx[qx][qy]; hme_y = fy[qx][qy]; }
                    if (P->enable_hme_level2 && P->ref_poc[0] == P->ref_poc[1] && list == 1 && nqx * nqy > 1) {
                        /* same picture in both lists: take the second best quadrant (the reference
                         * sorts with [q / nqx][q % nqx] indexing and then reads entry [0][1]) */
                        const int tq = nqx * nqy;
                        for (int a = 0; a < tq - 1; a++)
                            for (int b = a + 1; b < tq; b++)
                                if (l2s[a / nqx][a % nqx] > l2s[b / nqx][b % nqx]) {
                                    int16_t tx = l2x[a / nqx][a % nqx], ty = l2y[a / nqx][a % nqx];
                                    uint64_t ts = l2s[a / nqx][a % nqx];
                                    l2x[a / nqx][a % nqx] = l2x[b / nqx][b % nqx]; l2y[a / nqx][a % nqx] = l2y[b / nqx][b % nqx];
                                    l2s[a / nqx][a % nqx] = l2s[b / nqx][b % nqx];
                                    l2x[b / nqx][b % nqx] = tx; l2y[b / nqx][b % nqx] = ty; l2s[b / nqx][b % nqx] = ts;
                                }
                        hme_x = l2x[0][1]; hme_y = l2y[0][1];
                    }
                }
                xsc = hme_x; ysc = hme_y;
            }
        }
        if (trace) { trace->hme_center_x[list] = xsc; trace->hme_center_y[list] = ysc; }

        /* (0,0) versus the HME centre: ties go to (0,0) */
        if (xsc != 0 || ysc != 0) {
            const SvtB200Plane *rp = c->ref[list][0];
            uint64_t zero = sb_sad_at(c, rp, 0, 0);
            xsc = clip_center(c->ox, xsc, SB - 1, rp->width);
            ysc = clip_center(c->oy, ysc, SB - 1, rp->height);
            uint64_t hme = sb_sad_at(c, rp, xsc, ysc);
            if (zero <= hme) { xsc = 0; ysc = 0; }
        }
        int16_t saw = (int16_t)(P->search_area_width < 127 ? P->search_area_width : 127);
        int16_t sah = (int16_t)(P->search_area_height < 127 ? P->search_area_height : 127);
        int16_t xo = (int16_t)(xsc - (saw >> 1)), yo = (int16_t)(ysc - (sah >> 1));
        clip_window(c->ox, &xo, &saw, SB - 1, pic_w);
        clip_window(c->oy, &yo, &sah, SB - 1, pic_h);
        c->sa_ox[list] = xo; c->sa_oy[list] = yo; c->saw[list] = saw; c->sah[list] = sah;
        if (trace) { trace->sa_ox[list] = xo; trace->sa_oy[list] = yo; trace->sa_w[list] = saw; trace->sa_h[list] = sah; }

        full_pel_search(c, list);
        if (trace) { memcpy(trace->fullpel_sad[list], c->best_sad[list], sizeof(c->best_sad[list])); memcpy(trace->fullpel_mv[list], c->best_mv[list], sizeof(c->best_mv[list])); }

        int e32 = 0, e16 = 0, e8 = 0, eq = 0;
        if (P->fractional_search_model == 0) e32 = e16 = e8 = eq = 1;
        else if (P->fractional_search_model == 1) { su_pel_enable(c, list, &e32, &e16, &e8); eq = 1; }
        if (trace) trace->subpel_enable[list] = (uint8_t)(e32 | (e16 << 1) | (e8 << 2) | (eq << 3));
        if (e32 || e16 || e8 || eq) {
            interp_region(c, list);
            subpel_search(c, list, e32, e16, e8, eq);
        }
    }

    /* pack: output PU order is raster within each size class, internal order is z-order */
    for (int pu = 0; pu < 85; pu++) {
        const int n_idx = pu > 20 ? tab8[pu - 21] + 21 : pu > 4 ? tab16[pu - 5] + 5 : pu;
        int total = nl;
        if (nl == 2) {
            const int cond = (P->cu8x8_mode == 0 || pu < 21) && (P->cu16x16_mode == 0 || pu < 5);
            if (cond) {
                const int w = pu == 0 ? 64 : pu < 5 ? 32 : pu < 21 ? 16 : 8;
                int px, py;
                if (pu == 0) { px = py = 0; }
                else if (pu < 5) { px = ((pu - 1) & 1) * 32; py = ((pu - 1) >> 1) * 32; }
                else if (pu < 21) { px = ((pu - 5) & 3) * 16; py = ((pu - 5) >> 2) * 16; }
                else { px = ((pu - 21) & 7) * 8; py = ((pu - 21) >> 3) * 8; }
                c->bipred_sad[n_idx] = bipred_sad(c, n_idx, px, py, w, w);
                total = 3;
            }
        }
        SvtB200MeCuResult *r = &out[pu];
        memset(r, 0, sizeof(*r));
        r->total_me_candidate_index = (uint8_t)total;
        r->x_mv_l0 = mvx(c->best_mv[0][n_idx]); r->y_mv_l0 = mvy(c->best_mv[0][n_idx]);
        if (nl == 2) { r->x_mv_l1 = mvx(c->best_mv[1][n_idx]); r->y_mv_l1 = mvy(c->best_mv[1][n_idx]); }
        /* else: the reference leaves whatever an earlier SB put in its scratch; reported as 0 here */
        const uint32_t a = c->best_sad[0][n_idx], b = c->best_sad[1][n_idx], cc = c->bipred_sad[n_idx];
        if (total == 3) {
            uint32_t d[3]; uint32_t k[3];
            if (a <= b && a <= cc) { d[0] = a; k[0] = 0; if (b <= cc) { d[1] = b; k[1] = 1; d[2] = cc; k[2] = 2; } else { d[1] = cc; k[1] = 2; d[2] = b; k[2] = 1; } }
            else if (b <= a && b <= cc) { d[0] = b; k[0] = 1; if (a <= cc) { d[1] = a; k[1] = 0; d[2] = cc; k[2] = 2; } else { d[1] = cc; k[1] = 2; d[2] = a; k[2] = 0; } }
            else if (a <= b) { d[0] = cc; k[0] = 2; d[1] = a; k[1] = 0; d[2] = b; k[2] = 1; }
            else { d[0] = cc; k[0] = 2; d[1] = b; k[1] = 1; d[2] = a; k[2] = 0; }
            for (int i = 0; i < 3; i++) { r->cand[i].distortion = d[i]; r->cand[i].direction = k[i]; }
        } else if (total == 2) {
            if (a <= b) { r->cand[0].distortion = a; r->cand[0].direction = 0; r->cand[1].distortion = b; r->cand[1].direction = 1; }
            else { r->cand[0].distortion = b; r->cand[0].direction = 1; r->cand[1].distortion = a; r->cand[1].direction = 0; }
        } else {
            r->cand[0].distortion = a; r->cand[0].direction = 0;
        }
    }
    if (rcme_distortion) {
        uint32_t s = 0;
        for (int i = 0; i < 16; i++) s += out[5 + i].cand[0].distortion;
        *rcme_distortion = s;
    }
    for (int l = 0; l < 2; l++) { free(c->posb[l]); free(c->posh[l]); free(c->posj[l]); }
    free(c);
}

void svtb200_oracle_me_picture(const SvtB200MeParams *P, const SvtB200Plane cur[3], const SvtB200Plane ref0[3],
                               const SvtB200Plane ref1[3], SvtB200MeCuResult *out, uint32_t *rcme,
                               int sb_begin, int sb_end) {
    const int sbs_w = (P->luma_width + SB - 1) / SB;
    for (int sb = sb_begin; sb < sb_end; sb++)
        svtb200_oracle_me_sb(P, cur, ref0, ref1, sb % sbs_w, sb / sbs_w, out + (size_t)sb * 85, rcme ? rcme + sb : 0, 0);
}

/* compute_zz_sad (Source/Lib/Codec/EbMotionEstimationProcess.c:431-527, the ADAPTIVE_QP_INDEX_GEN branch): 16x16 SAD of the
 * current 1/16 plane against eb_vp9_decimation_2d(previous full plane, step 4) per complete SB, mapped to the non-moving
 * scores with thresholds (block_width * block_height * {2,4,8}) >> non_moving_th_shift[input_resolution] (:353, :503-515). */
void svtb200_oracle_zz_sad_picture(const SvtB200Plane *cur16, const SvtB200Plane *prev_full, int w, int h, int th_shift,
                                   uint8_t *scores, uint32_t *sads) {
    const int sbs_w = (w + SB - 1) / SB, sbs_h = (h + SB - 1) / SB;
    for (int sy = 0; sy < sbs_h; sy++)
        for (int sx = 0; sx < sbs_w; sx++) {
            const int ox = sx * SB, oy = sy * SB, sb = sx + sy * sbs_w;
            const int bw = (w - ox < SB ? w - ox : SB) >> 2, bh = (h - oy < SB ? h - oy : SB) >> 2;
            uint32_t sad = 0xffffffffu;
            if (w - ox >= SB && h - oy >= SB) {
                uint8_t dec[16 * 16];
                const uint8_t *p = prev_full->buf + (size_t)(prev_full->origin_y + oy) * prev_full->stride + prev_full->origin_x + ox;
                for (int y = 0; y < SB; y += 4)
                    for (int x = 0; x < SB; x += 4) dec[(y >> 2) * 16 + (x >> 2)] = p[(size_t)y * prev_full->stride + x];
                const uint8_t *c = cur16->buf + (size_t)(cur16->origin_y + (oy >> 2)) * cur16->stride + cur16->origin_x + (ox >> 2);
                sad = 0;
                for (int y = 0; y < 16; y++)
                    for (int x = 0; x < 16; x++) { const int d = (int)c[(size_t)y * cur16->stride + x] - (int)dec[y * 16 + x]; sad += (uint32_t)(d < 0 ? -d : d); }
            }
            const uint32_t area = (uint32_t)(bw * bh);
            scores[sb] = sad < ((area * 2) >> th_shift) ? 0 : sad < ((area * 4) >> th_shift) ? 10 : sad < ((area * 8) >> th_shift) ? 20 : 30;
            if (sads) sads[sb] = sad;
        }
}


/* compute_block_mean_compute_variance (Source/Lib/Codec/EbPictureAnalysisProcess.c:2115-3356): 8x8 means from rows 0,2,4,6
 * (eb_vp9_compute_sub_mean8x8_sse2_intrin: sum << 3; eb_vp9_compute_subd_mean_of_squared_values8x8_sse2_intrin: sum of squares
 * << 11, ASM_SSE2/EbComputeMean_Intrinsic_SSE2.c:10-53), quadrant averages (>> 2) up to 64x64 (:2732-2960), outputs
 * y_mean = mean >> MEAN_PRECISION(8) and variance = (mean_sq - mean^2) >> VARIANCE_PRECISION(16) (:2962-3350). */
void svtb200_oracle_block_stats_picture(const SvtB200Plane *full, int w, int h, uint8_t *y_mean, uint16_t *variance) {
    const int sbs_w = (w + SB - 1) / SB, sbs_h = (h + SB - 1) / SB;
    for (int sb = 0; sb < sbs_w * sbs_h; sb++) {
        const int ox = (sb % sbs_w) * SB, oy = (sb / sbs_w) * SB;
        uint64_t m[85], q[85];
        for (int b = 0; b < 64; b++) {
            const uint8_t *p = full->buf + (size_t)(full->origin_y + oy + (b >> 3) * 8) * full->stride + full->origin_x + ox + (b & 7) * 8;
            uint64_t s = 0, s2 = 0;
            for (int r = 0; r < 8; r += 2)
                for (int x = 0; x < 8; x++) { const uint64_t v = p[(size_t)r * full->stride + x]; s += v; s2 += v * v; }
            m[21 + b] = s << 3; q[21 + b] = s2 << 11;
        }
        for (int i = 0; i < 16; i++) {
            const int a = 21 + (2 * (i >> 2)) * 8 + 2 * (i & 3);
            m[5 + i] = (m[a] + m[a + 1] + m[a + 8] + m[a + 9]) >> 2; q[5 + i] = (q[a] + q[a + 1] + q[a + 8] + q[a + 9]) >> 2;
        }
        for (int i = 0; i < 4; i++) {
            const int a = 5 + (2 * (i >> 1)) * 4 + 2 * (i & 1);
            m[1 + i] = (m[a] + m[a + 1] + m[a + 4] + m[a + 5]) >> 2; q[1 + i] = (q[a] + q[a + 1] + q[a + 4] + q[a + 5]) >> 2;
        }
        m[0] = (m[1] + m[2] + m[3] + m[4]) >> 2; q[0] = (q[1] + q[2] + q[3] + q[4]) >> 2;
        for (int i = 0; i < 85; i++) { y_mean[(size_t)sb * 85 + i] = (uint8_t)(m[i] >> 8); variance[(size_t)sb * 85 + i] = (uint16_t)((q[i] - m[i] * m[i]) >> 16); }
    }
}


/* sub_sample_luma_generate_pixel_intensity_histogram_bins + calculate_histogram (EbPictureAnalysisProcess.c:4237-4312, :128-150)
 * on the 1/16 plane: bins start at 1, are counted with decimation step 1 and end up << 4; average = (sum + area/2) / area. */
void svtb200_oracle_luma_histogram(const SvtB200Plane *p16, int rw, int rh, uint32_t *hist, uint8_t *avg, uint64_t *sum16) {
    const int regw = p16->width / rw, regh = p16->height / rh;
    *sum16 = 0;
    for (int iw = 0; iw < rw; iw++)
        for (int ih = 0; ih < rh; ih++) {
            uint32_t *hg = hist + ((size_t)iw * rh + ih) * 256;
            const int cw = regw + (iw == rw - 1 ? p16->width - rw * regw : 0), ch = regh + (ih == rh - 1 ? p16->height - rh * regh : 0);
            const uint8_t *s = p16->buf + (size_t)(p16->origin_y + ih * regh) * p16->stride + p16->origin_x + iw * regw;
            uint64_t sum = 0;
            for (int b = 0; b < 256; b++) hg[b] = 1;
            for (int y = 0; y < ch; y++)
                for (int x = 0; x < cw; x++) { hg[s[(size_t)y * p16->stride + x]]++; sum += s[(size_t)y * p16->stride + x]; }
            avg[iw * rh + ih] = (uint8_t)((sum + (((uint64_t)cw * ch) >> 1)) / ((uint64_t)cw * ch));
            *sum16 += sum << 4;
            for (int b = 0; b < 256; b++) hg[b] <<= 4;
        }
}
