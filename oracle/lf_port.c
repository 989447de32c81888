/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle ("port") for the in-loop deblocking filter.
 *
 * Scalar restatement of eb_vp9_loop_filter_frame(): per 64x64 super-block in raster order, border adjustment of
 * the edge masks, then per plane all vertical edges followed by all horizontal edges, each 8-sample edge segment
 * filtered by the 16-, 8- or 4-wide filter its transform size selects (+ the internal 4x4 edge).
 *
 * Parity pinning: oracle/ref_seam_check.c (--wrap=eb_vp9_loop_filter_frame) runs this port next to the reference's
 * own filter on every loop-filtered picture of real encodes and compares the three planes byte for byte;
 * tests/golden/lf_*.npz hold pictures dumped the same way.
 *
 * Reference functions followed (Source/Lib/VPX):
 *   eb_vp9_loop_filter_frame / loop_filter_rows          vp9_loopfilter.c:1521 / :1456
 *   eb_vp9_adjust_mask                                   vp9_loopfilter.c:786-899
 *   eb_vp9_filter_block_plane_ss00 / _ss11               vp9_loopfilter.c:1238 / :1342
 *   filter_selectively_vert_row2 / filter_selectively_horiz   vp9_loopfilter.c:305 / :481
 *   filter_mask, flat_mask4/5, hev_mask, filter4/8/16, eb_vp9_lpf_*_c   loopfilter.c:31-329
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

static inline int iabs(int v) { return v < 0 ? -v : v; }
static inline int8_t sclamp(int t) { return (int8_t)(t < -128 ? -128 : t > 127 ? 127 : t); }

/* one sample position across an edge: s points at q0, 'a' is the distance between taps (1 or pitch).
 * width: 4, 8 or 16 = widest filter allowed at this edge. */
static void filter_across(uint8_t *s, ptrdiff_t a, int width, const SvtB200LfThresh *t) {
    const int p3 = s[-4 * a], p2 = s[-3 * a], p1 = s[-2 * a], p0 = s[-a], q0 = s[0], q1 = s[a], q2 = s[2 * a], q3 = s[3 * a];
    const int lim = t->lim, blim = t->mblim, thr = t->hev_thr;
    const int mask = !(iabs(p3 - p2) > lim || iabs(p2 - p1) > lim || iabs(p1 - p0) > lim || iabs(q1 - q0) > lim ||
                       iabs(q2 - q1) > lim || iabs(q3 - q2) > lim || iabs(p0 - q0) * 2 + iabs(p1 - q1) / 2 > blim);
    const int flat = width >= 8 && !(iabs(p1 - p0) > 1 || iabs(q1 - q0) > 1 || iabs(p2 - p0) > 1 || iabs(q2 - q0) > 1 ||
                                     iabs(p3 - p0) > 1 || iabs(q3 - q0) > 1);
    if (width == 16 && flat && mask) {
        const int p7 = s[-8 * a], p6 = s[-7 * a], p5 = s[-6 * a], p4 = s[-5 * a], q4 = s[4 * a], q5 = s[5 * a], q6 = s[6 * a], q7 = s[7 * a];
        const int flat2 = !(iabs(p4 - p0) > 1 || iabs(q4 - q0) > 1 || iabs(p5 - p0) > 1 || iabs(q5 - q0) > 1 ||
                            iabs(p6 - p0) > 1 || iabs(q6 - q0) > 1 || iabs(p7 - p0) > 1 || iabs(q7 - q0) > 1);
        if (flat2) { /* 15-tap smoothing: sliding window sum with replicated ends */
            const int v[16] = {p7, p6, p5, p4, p3, p2, p1, p0, q0, q1, q2, q3, q4, q5, q6, q7};
            int out[16];
            for (int i = 1; i < 15; i++) {
                int sum = v[i]; /* centre counted twice */
                for (int k = i - 7; k <= i + 7; k++) sum += v[k < 0 ? 0 : k > 15 ? 15 : k];
                out[i] = (sum + 8) >> 4;
            }
            for (int i = 1; i < 15; i++) s[(i - 8) * a] = (uint8_t)out[i];
            return;
        }
    }
    if (flat && mask) { /* 7-tap smoothing */
        const int v[8] = {p3, p2, p1, p0, q0, q1, q2, q3};
        int out[8];
        for (int i = 1; i < 7; i++) {
            int sum = v[i];
            for (int k = i - 3; k <= i + 3; k++) sum += v[k < 0 ? 0 : k > 7 ? 7 : k];
            out[i] = (sum + 4) >> 3;
        }
        for (int i = 1; i < 7; i++) s[(i - 4) * a] = (uint8_t)out[i];
        return;
    }
    { /* narrow filter on the signed representation */
        const int8_t ps1 = (int8_t)(p1 ^ 0x80), ps0 = (int8_t)(p0 ^ 0x80), qs0 = (int8_t)(q0 ^ 0x80), qs1 = (int8_t)(q1 ^ 0x80);
        const int8_t m = mask ? -1 : 0, hev = (iabs(p1 - p0) > thr || iabs(q1 - q0) > thr) ? -1 : 0;
        int8_t f = sclamp(ps1 - qs1) & hev;
        f = sclamp(f + 3 * (qs0 - ps0)) & m;
        const int8_t f1 = sclamp(f + 4) >> 3, f2 = sclamp(f + 3) >> 3;
        s[0]  = (uint8_t)(sclamp(qs0 - f1) ^ 0x80);
        s[-a] = (uint8_t)(sclamp(ps0 + f2) ^ 0x80);
        f = (int8_t)(((f1 + 1) >> 1) & ~hev);
        s[a]      = (uint8_t)(sclamp(qs1 - f) ^ 0x80);
        s[-2 * a] = (uint8_t)(sclamp(ps1 + f) ^ 0x80);
    }
}

/* 8 consecutive positions along the edge */
static void filter_segment(uint8_t *s, ptrdiff_t across, ptrdiff_t along, int width, const SvtB200LfThresh *t) {
    for (int i = 0; i < 8; i++) filter_across(s + i * along, across, width, t);
}

void svtb200_oracle_lf_segment(uint8_t *s, long across, long along, int width, const SvtB200LfThresh *t) {
    filter_segment(s, across, along, width, t);
}

void svtb200_oracle_lf_adjust_mask(int mi_rows, int mi_cols, int mi_row, int mi_col, SvtB200LfMask *m) {
    m->left_y[2] |= m->left_y[3]; m->above_y[2] |= m->above_y[3];
    m->left_uv[2] |= m->left_uv[3]; m->above_uv[2] |= m->above_uv[3];
    /* 4x4 edges on a 32x32 boundary are promoted to the 8-wide filter */
    const uint64_t lb = 0x1111111111111111ull, ab = 0x000000ff000000ffull;
    m->left_y[1] |= m->left_y[0] & lb; m->left_y[0] &= ~lb;
    m->above_y[1] |= m->above_y[0] & ab; m->above_y[0] &= ~ab;
    m->left_uv[1] |= m->left_uv[0] & 0x1111; m->left_uv[0] &= (uint16_t)~0x1111;
    m->above_uv[1] |= m->above_uv[0] & 0x000f; m->above_uv[0] &= (uint16_t)~0x000f;
    if (mi_row + 8 > mi_rows) {
        const uint64_t rows = (uint64_t)(mi_rows - mi_row);
        const uint64_t my = ((uint64_t)1 << (rows << 3)) - 1;
        const uint16_t muv = (uint16_t)(((uint16_t)1 << (((rows + 1) >> 1) << 2)) - 1);
        for (int i = 0; i < 3; i++) { m->left_y[i] &= my; m->above_y[i] &= my; m->left_uv[i] &= muv; m->above_uv[i] &= muv; }
        m->int_4x4_y &= my; m->int_4x4_uv &= muv;
        if (rows == 1) { m->above_uv[1] |= m->above_uv[2]; m->above_uv[2] = 0; }
        if (rows == 5) { m->above_uv[1] |= m->above_uv[2] & 0xff00; m->above_uv[2] &= (uint16_t)~(m->above_uv[2] & 0xff00); }
    }
    if (mi_col + 8 > mi_cols) {
        const uint64_t cols = (uint64_t)(mi_cols - mi_col);
        const uint64_t my = (((uint64_t)1 << cols) - 1) * 0x0101010101010101ull;
        const uint16_t muv = (uint16_t)(((1 << ((cols + 1) >> 1)) - 1) * 0x1111);
        const uint16_t muvi = (uint16_t)(((1 << (cols >> 1)) - 1) * 0x1111);
        for (int i = 0; i < 3; i++) { m->left_y[i] &= my; m->above_y[i] &= my; m->left_uv[i] &= muv; m->above_uv[i] &= muv; }
        m->int_4x4_y &= my; m->int_4x4_uv &= muvi;
        if (cols == 1) { m->left_uv[1] |= m->left_uv[2]; m->left_uv[2] = 0; }
        if (cols == 5) { m->left_uv[1] |= m->left_uv[2] & 0xcccc; m->left_uv[2] &= (uint16_t)~(m->left_uv[2] & 0xcccc); }
    }
    if (mi_col == 0) for (int i = 0; i < 3; i++) { m->left_y[i] &= 0xfefefefefefefefeull; m->left_uv[i] &= 0xeeee; }
}

static void hseg(uint8_t *s, ptrdiff_t across, int n, int width, const SvtB200LfThresh *t) {
    for (int i = 0; i < n; i++) filter_across(s + i, across, width, t);
}

/* One plane of one super-block.  bw = 8x8 blocks per row of the mask grid (8 luma, 4 chroma), masks given as
 * row-major bit grids; lfl[row*bw + col] = filter level of the block. */
static void filter_plane_sb(uint8_t *buf, int stride, int bw, int rows /* block rows to filter */, int px_rows /* sample rows that exist below buf */, int px_cols /* sample columns that exist */, int first_picture_row,
                            uint64_t l16, uint64_t l8, uint64_t l4, uint64_t a16, uint64_t a8, uint64_t a4, uint64_t i4_v,
                            uint64_t i4_h, const uint8_t *lfl, const SvtB200LfThresh *thr) {
    /* vertical edges: rows are handled in pairs; a 16-wide edge present in both rows of a pair takes the
     * thresholds of the pair's first row */
    for (int r = 0; r < rows; r++)
        for (int c = 0; c < bw; c++) {
            const int bit = r * bw + c;
            uint8_t *s = buf + (r * 8) * stride + c * 8;
            const int pr = r & ~1, other = r ^ 1;
            int lvl16 = lfl[bit];
            if ((l16 >> bit & 1) && (l16 >> (other * bw + c) & 1)) lvl16 = lfl[pr * bw + c];
            /* with an odd number of 8x8 luma rows the last chroma block row is only 4 samples high: the reference
             * also filters the 4 rows below the picture (inside its padded buffer, regenerated afterwards);
             * here nothing outside the picture is touched */
            const int nrow = px_rows - r * 8 < 8 ? px_rows - r * 8 : 8;
            for (int i = 0; i < nrow; i++) {
                uint8_t *q = s + i * stride;
                if (l16 >> bit & 1) filter_across(q, 1, 16, &thr[lvl16]);
                if (l8 >> bit & 1) filter_across(q, 1, 8, &thr[lfl[bit]]);
                if (l4 >> bit & 1) filter_across(q, 1, 4, &thr[lfl[bit]]);
                if (i4_v >> bit & 1) filter_across(q + 4, 1, 4, &thr[lfl[bit]]);
            }
        }
    /* horizontal edges, row by row; no edge filtering on the first picture row; adjacent 16-wide edges are
     * paired from the start of each run and share the first block's thresholds */
    for (int r = 0; r < rows; r++) {
        const int top = first_picture_row && r == 0;
        int run_start = -1;
        for (int c = 0; c < bw; c++) {
            const int bit = r * bw + c;
            uint8_t *s = buf + (r * 8) * stride + c * 8;
            const int m16 = !top && (a16 >> bit & 1), m8 = !top && (a8 >> bit & 1), m4 = !top && (a4 >> bit & 1);
            const int mi = (int)(i4_h >> bit & 1);
            const int ncol = px_cols - c * 8 < 8 ? px_cols - c * 8 : 8; /* the last chroma block column may be 4 samples wide */
            if (m16) {
                if (run_start < 0) run_start = c;
                const int lead = c - ((c - run_start) & 1);
                hseg(s, stride, ncol, 16, &thr[lfl[r * bw + lead]]);
            } else {
                run_start = -1;
                if (m8) hseg(s, stride, ncol, 8, &thr[lfl[bit]]);
                else if (m4) hseg(s, stride, ncol, 4, &thr[lfl[bit]]);
                if (mi && r * 8 + 4 < px_rows) hseg(s + 4 * stride, stride, ncol, 4, &thr[lfl[bit]]);
            }
        }
    }
}

/* In-place deblocking of a whole picture.  lfm (lfm_stride entries per SB row) is modified exactly as the
 * reference's eb_vp9_adjust_mask does. */
void svtb200_oracle_deblock_picture(const SvtB200LfParams *p, uint8_t *y, int y_stride, uint8_t *u, uint8_t *v,
                                    int uv_stride, SvtB200LfMask *lfm, int lfm_stride, const SvtB200LfThresh thr[64]) {
    if (!p->filter_level) return;
    for (int mi_row = 0; mi_row < p->mi_rows; mi_row += 8)
        for (int mi_col = 0; mi_col < p->mi_cols; mi_col += 8) {
            SvtB200LfMask *m = &lfm[(mi_col >> 3) + (mi_row >> 3) * lfm_stride];
            svtb200_oracle_lf_adjust_mask(p->mi_rows, p->mi_cols, mi_row, mi_col, m);
            const int rows = p->mi_rows - mi_row < 8 ? p->mi_rows - mi_row : 8;
            const int cols = p->mi_cols - mi_col < 8 ? p->mi_cols - mi_col : 8;
            filter_plane_sb(y + (size_t)mi_row * 8 * y_stride + mi_col * 8, y_stride, 8, rows, rows * 8, cols * 8, mi_row == 0, m->left_y[2], m->left_y[1],
                            m->left_y[0], m->above_y[2], m->above_y[1], m->above_y[0], m->int_4x4_y, m->int_4x4_y, m->lfl_y, thr);
            if (p->y_only) continue;
            /* chroma: 4x4 grid of 8x8 chroma blocks; levels sampled from the even luma rows / columns */
            uint8_t lfl_uv[16];
            for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) lfl_uv[r * 4 + c] = m->lfl_y[(r * 2) * 8 + c * 2];
            const int crows = (rows + 1) >> 1;
            /* the internal 4x4 edge of the last chroma row of the picture is skipped in the horizontal pass
             * (mi_row + r == mi_rows - 1 with r even, i.e. an odd number of luma block rows) */
            uint64_t i4h = m->int_4x4_uv;
            for (int r = 0; r < 8; r += 2)
                if (mi_row + r == p->mi_rows - 1) i4h &= ~((uint64_t)0xf << ((r >> 1) * 4));
            for (int pl = 0; pl < 2; pl++) {
                uint8_t *b = (pl ? v : u) + (size_t)(mi_row * 4) * uv_stride + mi_col * 4;
                filter_plane_sb(b, uv_stride, 4, crows, rows * 4, cols * 4, mi_row == 0, m->left_uv[2], m->left_uv[1], m->left_uv[0], m->above_uv[2],
                                m->above_uv[1], m->above_uv[0], m->int_4x4_uv, i4h, lfl_uv, thr);
            }
        }
}


/* ---------------------------------------------------------------------------------------------------------------------
 * Edge-mask construction: eb_vp9_build_mask_frame (Source/Lib/VPX/vp9_loopfilter.c:1548) -> eb_vp9_setup_mask (:901-1018)
 * -> build_masks (:677-748) / build_y_mask (:753-784), with the lookup tables the reference keeps at :36-200 restated as
 * explicit literals in the reference's own index order (BLOCK_SIZE 0..12, TX_SIZE 0..3).
 * ------------------------------------------------------------------------------------------------------------------- */
static const uint64_t o_left_pred[13] = {0x1ull, 0x1ull, 0x1ull, 0x1ull, 0x0101ull, 0x1ull, 0x0101ull, 0x01010101ull, 0x0101ull,
                                         0x01010101ull, 0x0101010101010101ull, 0x01010101ull, 0x0101010101010101ull};
static const uint64_t o_above_pred[13] = {0x1ull, 0x1ull, 0x1ull, 0x1ull, 0x1ull, 0x3ull, 0x3ull, 0x3ull, 0xfull, 0xfull, 0xfull, 0xffull, 0xffull};
static const uint64_t o_size_mask[13] = {0x1ull, 0x1ull, 0x1ull, 0x1ull, 0x0101ull, 0x3ull, 0x0303ull, 0x03030303ull, 0x0f0full, 0x0f0f0f0full,
                                         0x0f0f0f0f0f0f0f0full, 0xffffffffull, 0xffffffffffffffffull};
static const uint16_t o_left_pred_uv[13] = {0x1, 0x1, 0x1, 0x1, 0x1, 0x1, 0x1, 0x11, 0x1, 0x11, 0x1111, 0x11, 0x1111};
static const uint16_t o_above_pred_uv[13] = {0x1, 0x1, 0x1, 0x1, 0x1, 0x1, 0x1, 0x1, 0x3, 0x3, 0x3, 0xf, 0xf};
static const uint16_t o_size_mask_uv[13] = {0x1, 0x1, 0x1, 0x1, 0x1, 0x1, 0x1, 0x11, 0x3, 0x33, 0x3333, 0xff, 0xffff};
static const uint64_t o_left_tx[4] = {0xffffffffffffffffull, 0xffffffffffffffffull, 0x5555555555555555ull, 0x1111111111111111ull};
static const uint64_t o_above_tx[4] = {0xffffffffffffffffull, 0xffffffffffffffffull, 0x00ff00ff00ff00ffull, 0x000000ff000000ffull};
static const uint16_t o_left_tx_uv[4] = {0xffff, 0xffff, 0x5555, 0x1111};
static const uint16_t o_above_tx_uv[4] = {0xffff, 0xffff, 0x0f0f, 0x000f};
static const uint8_t o_w8[13] = {1, 1, 1, 1, 1, 2, 2, 2, 4, 4, 4, 8, 8}, o_h8[13] = {1, 1, 1, 1, 2, 1, 2, 4, 2, 4, 8, 4, 8};
/* eb_vp9_uv_txsize_lookup[bs][tx][1][1] (4:2:0) */
static const uint8_t o_uv_tx[13][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 1, 1, 1},
                                       {0, 1, 1, 1}, {0, 1, 1, 1}, {0, 1, 2, 2}, {0, 1, 2, 2}, {0, 1, 2, 2}, {0, 1, 2, 3}};

static void o_build(const SvtB200LfModeInfo *mi, int sy, int su, int uv, SvtB200LfMask *m) {
    const int bs = mi->sb_type, tx = mi->tx_size, txu = o_uv_tx[bs][tx];
    if (!mi->filter_level) return;
    for (int i = 0; i < o_h8[bs]; i++) memset(&m->lfl_y[sy + 8 * i], mi->filter_level, o_w8[bs]);
    m->above_y[tx] |= o_above_pred[bs] << sy; m->left_y[tx] |= o_left_pred[bs] << sy;
    if (uv) { m->above_uv[txu] |= (uint16_t)(o_above_pred_uv[bs] << su); m->left_uv[txu] |= (uint16_t)(o_left_pred_uv[bs] << su); }
    if (mi->skip_inter) return;
    m->above_y[tx] |= (o_size_mask[bs] & o_above_tx[tx]) << sy; m->left_y[tx] |= (o_size_mask[bs] & o_left_tx[tx]) << sy;
    if (uv) {
        m->above_uv[txu] |= (uint16_t)((o_size_mask_uv[bs] & o_above_tx_uv[txu]) << su);
        m->left_uv[txu] |= (uint16_t)((o_size_mask_uv[bs] & o_left_tx_uv[txu]) << su);
    }
    if (tx == 0) m->int_4x4_y |= o_size_mask[bs] << sy;
    if (uv && txu == 0) m->int_4x4_uv |= (uint16_t)((o_size_mask_uv[bs] & 0xffff) << su);
}

void svtb200_oracle_build_masks(int mi_rows, int mi_cols, const SvtB200LfModeInfo *mi, int stride, SvtB200LfMask *lfm, int lfm_stride) {
    static const int s32y[4] = {0, 4, 32, 36}, s16y[4] = {0, 2, 16, 18}, s8y[4] = {0, 1, 8, 9}, s32u[4] = {0, 2, 8, 10}, s16u[4] = {0, 1, 4, 5};
    for (int mr = 0; mr < mi_rows; mr += 8)
        for (int mc = 0; mc < mi_cols; mc += 8) {
            SvtB200LfMask *m = &lfm[(mr >> 3) * lfm_stride + (mc >> 3)];
            const SvtB200LfModeInfo *p = mi + (size_t)mr * stride + mc;
            const int max_rows = mr + 8 > mi_rows ? mi_rows - mr : 8, max_cols = mc + 8 > mi_cols ? mi_cols - mc : 8;
            memset(m, 0, sizeof(*m));
            switch (p->sb_type) {
            case 12: o_build(p, 0, 0, 1, m); break;
            case 11: o_build(p, 0, 0, 1, m); if (4 < max_rows) o_build(p + 4 * stride, 32, 8, 1, m); break;
            case 10: o_build(p, 0, 0, 1, m); if (4 < max_cols) o_build(p + 4, 4, 2, 1, m); break;
            default:
                for (int a = 0; a < 4; a++) {
                    const int c32 = (a & 1) << 2, r32 = (a >> 1) << 2;
                    const SvtB200LfModeInfo *q = p + r32 * stride + c32;
                    if (c32 >= max_cols || r32 >= max_rows) continue;
                    switch (q->sb_type) {
                    case 9: o_build(q, s32y[a], s32u[a], 1, m); break;
                    case 8: o_build(q, s32y[a], s32u[a], 1, m); if (r32 + 2 < max_rows) o_build(q + 2 * stride, s32y[a] + 16, s32u[a] + 4, 1, m); break;
                    case 7: o_build(q, s32y[a], s32u[a], 1, m); if (c32 + 2 < max_cols) o_build(q + 2, s32y[a] + 2, s32u[a] + 1, 1, m); break;
                    default:
                        for (int b = 0; b < 4; b++) {
                            const int c16 = c32 + ((b & 1) << 1), r16 = r32 + ((b >> 1) << 1);
                            const SvtB200LfModeInfo *t = p + r16 * stride + c16;
                            const int sy = s32y[a] + s16y[b], su = s32u[a] + s16u[b];
                            if (c16 >= max_cols || r16 >= max_rows) continue;
                            switch (t->sb_type) {
                            case 6: o_build(t, sy, su, 1, m); break;
                            case 5: o_build(t, sy, su, 1, m); if (r16 + 1 < max_rows) o_build(t + stride, sy + 8, 0, 0, m); break;
                            case 4: o_build(t, sy, su, 1, m); if (c16 + 1 < max_cols) o_build(t + 1, sy + 1, 0, 0, m); break;
                            default:
                                o_build(t, sy + s8y[0], su, 1, m);
                                for (int e = 1; e < 4; e++) {
                                    const int c8 = c16 + (e & 1), r8 = r16 + (e >> 1);
                                    if (c8 >= max_cols || r8 >= max_rows) continue;
                                    o_build(p + r8 * stride + c8, sy + s8y[e], 0, 0, m);
                                }
                            }
                        }
                    }
                }
            }
        }
}
