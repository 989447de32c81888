/* pred_port.c -- TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into or called by the product library).
 *
 * Restatement of the prediction step of the reference's encode pass and of its full-loop distortion:
 *   inter_predictor()                 Source/Lib/VPX/vp9_reconinter.h:23   (dispatch on subpel_x != 0 / subpel_y != 0)
 *   scale_factors::predict, unscaled  Source/Lib/VPX/vp9_scale.c:78-83,127-128
 *   convolve_horiz / convolve_vert    Source/Lib/VPX/vpx_convolve.c:20-38, 60-78
 *   eb_vp9_convolve8_c (two passes through an 8-bit intermediate of h + 7 rows)   vpx_convolve.c:133-168
 *   eb_vp9_convolve_copy_c / eb_vp9_convolve_avg_c / *_avg_* (second reference)   vpx_convolve.c:170-215
 *   sub_pel_filters_8 (the only kernel build_inter_predictors uses, vp9_reconinter.c:105)   Source/Lib/VPX/vp9_filter.c:32
 *   the intra predictors (d207 .. dc, generic :22-218 and the 4x4 forms :253-417)   Source/Lib/VPX/intrapred.c
 *     selected as build_intra_predictors does   Source/Lib/VPX/vp9_reconintra.c:397-407 (table :64-87)
 *   full_distortion_kernel32bit / full_distortion_kernel_eob_zero32bit   Source/Lib/C_DEFAULT/EbPictureOperators_C.c:289, 314
 * Pinned against those reference functions (C and AVX2 flavours) through oracle/_ref (tests/test_pred.py).
 */
#include "oracle.h"
#include <string.h>

/* the normative VP9 regular 8-tap filter, 16 phases */
static const int8_t k_regular[16][8] = {
    {0, 0, 0, 127, 0, 0, 0, 0} /* phase 0 is never applied (its centre tap is 128: a copy) */,
    {0, 1, -5, 126, 8, -3, 1, 0},    {-1, 3, -10, 122, 18, -6, 2, 0},  {-1, 4, -13, 118, 27, -9, 3, -1},
    {-1, 4, -16, 112, 37, -11, 4, -1}, {-1, 5, -18, 105, 48, -14, 4, -1}, {-1, 5, -19, 97, 58, -16, 5, -1},
    {-1, 6, -19, 88, 68, -18, 5, -1},  {-1, 6, -19, 78, 78, -19, 6, -1},  {-1, 5, -18, 68, 88, -19, 6, -1},
    {-1, 5, -16, 58, 97, -19, 5, -1},  {-1, 4, -14, 48, 105, -18, 5, -1}, {-1, 4, -11, 37, 112, -16, 4, -1},
    {-1, 3, -9, 27, 118, -13, 4, -1},  {0, 2, -6, 18, 122, -10, 3, -1},   {0, 1, -3, 8, 126, -5, 1, 0}};

static uint8_t clip8(int v) { return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }

/* one reference of one block into out (pitch 64) */
static void predict_one(const uint8_t *src, int stride, int sx, int sy, int w, int h, uint8_t out[64 * 64]) {
    uint8_t tmp[(64 + 7) * 64];
    if (!sx && !sy) { /* vpx_convolve.c:182 */
        for (int r = 0; r < h; r++) memcpy(out + r * 64, src + (ptrdiff_t)r * stride, (size_t)w);
        return;
    }
    if (sx && !sy) { /* horizontal only: h rows (vpx_convolve.c:101) */
        for (int r = 0; r < h; r++)
            for (int c = 0; c < w; c++) {
                int sum = 0;
                for (int k = 0; k < 8; k++) sum += src[(ptrdiff_t)r * stride + c + k - 3] * k_regular[sx][k];
                out[r * 64 + c] = clip8((sum + 64) >> 7);
            }
        return;
    }
    if (!sx) { /* vertical only (vpx_convolve.c:117) */
        for (int r = 0; r < h; r++)
            for (int c = 0; c < w; c++) {
                int sum = 0;
                for (int k = 0; k < 8; k++) sum += src[(ptrdiff_t)(r + k - 3) * stride + c] * k_regular[sy][k];
                out[r * 64 + c] = clip8((sum + 64) >> 7);
            }
        return;
    }
    /* both: rows -3 .. h+3 horizontally into the intermediate, rounded and clipped to 8 bits, then vertically (vpx_convolve.c:133) */
    for (int r = 0; r < h + 7; r++)
        for (int c = 0; c < w; c++) {
            int sum = 0;
            for (int k = 0; k < 8; k++) sum += src[(ptrdiff_t)(r - 3) * stride + c + k - 3] * k_regular[sx][k];
            tmp[r * 64 + c] = clip8((sum + 64) >> 7);
        }
    for (int r = 0; r < h; r++)
        for (int c = 0; c < w; c++) {
            int sum = 0;
            for (int k = 0; k < 8; k++) sum += tmp[(r + k) * 64 + c] * k_regular[sy][k];
            out[r * 64 + c] = clip8((sum + 64) >> 7);
        }
}

/* Same descriptor as the product's, with HOST pointers. */
void svtb200_oracle_inter_predict(const SvtB200PredBlock *b, int n) {
    for (int i = 0; i < n; i++, b++) {
        uint8_t p0[64 * 64], p1[64 * 64];
        predict_one(b->ref[0] + (ptrdiff_t)b->y[0] * b->ref_stride[0] + b->x[0], b->ref_stride[0], b->subpel_x[0], b->subpel_y[0], b->w, b->h, p0);
        if (b->compound)
            predict_one(b->ref[1] + (ptrdiff_t)b->y[1] * b->ref_stride[1] + b->x[1], b->ref_stride[1], b->subpel_x[1], b->subpel_y[1], b->w, b->h, p1);
        for (int r = 0; r < b->h; r++)
            for (int c = 0; c < b->w; c++)
                b->dst[(ptrdiff_t)r * b->dst_stride + c] = b->compound ? (uint8_t)((p0[r * 64 + c] + p1[r * 64 + c] + 1) >> 1) : p0[r * 64 + c];
    }
}

void svtb200_oracle_full_distortion(int tx_size, int n, const int16_t *coeff, const int16_t *recon_coeff, const uint16_t *eob, uint64_t *dist) {
    const int nn = (4 << tx_size) * (4 << tx_size);
    for (int k = 0; k < n; k++) {
        uint32_t res = 0, pred = 0; /* 32-bit sums, EbPictureOperators_C.c:293-294 */
        for (int i = 0; i < nn; i++) {
            const int32_t c = coeff[(size_t)k * nn + i], d = (int16_t)(c - recon_coeff[(size_t)k * nn + i]); /* sq_r16to32 takes an int16_t (:264) */
            res += (uint32_t)(d * d);
            pred += (uint32_t)(c * c);
        }
        dist[2 * k] = eob[k] ? res : pred;
        dist[2 * k + 1] = pred;
    }
}

/* ---- intra prediction.  Written as the fill order the predictors define (first rows / columns from the edge samples, the
 * rest by the diagonal recurrences), which is a different formulation from the closed forms the CUDA kernel evaluates. ---- */
#define A2(a, b) (((a) + (b) + 1) >> 1)
#define A3(a, b, c) (((a) + 2 * (b) + (c) + 2) >> 2)

static void intra_one(int mode, int bs, int have_left, int have_top, const uint8_t *a /* above, a[-1] valid */, const uint8_t *l, uint8_t *o /* pitch 32 */) {
    int r, c;
#define O(r, c) o[(r) * 32 + (c)]
    switch (mode) {
    case 0: { /* DC: dc_128 / dc_top / dc_left / dc (intrapred.c:167-218) */
        int sum = 0, v;
        if (have_left && have_top) { for (c = 0; c < bs; c++) sum += a[c] + l[c]; v = (sum + bs) / (2 * bs); }
        else if (have_left) { for (c = 0; c < bs; c++) sum += l[c]; v = (sum + (bs >> 1)) / bs; }
        else if (have_top) { for (c = 0; c < bs; c++) sum += a[c]; v = (sum + (bs >> 1)) / bs; }
        else v = 128;
        for (r = 0; r < bs; r++) for (c = 0; c < bs; c++) O(r, c) = (uint8_t)v;
        break;
    }
    case 1: for (r = 0; r < bs; r++) for (c = 0; c < bs; c++) O(r, c) = a[c]; break;                       /* V :137 */
    case 2: for (r = 0; r < bs; r++) for (c = 0; c < bs; c++) O(r, c) = l[r]; break;                       /* H :147 */
    case 9: for (r = 0; r < bs; r++) for (c = 0; c < bs; c++) { int v = l[r] + a[c] - a[-1]; O(r, c) = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); } break; /* TM :157 */
    case 3: /* D45 :59, 4x4 :313 (which runs on to above[7]) */
        if (bs == 4) {
            for (r = 0; r < 4; r++) for (c = 0; c < 4; c++) O(r, c) = (uint8_t)(r + c < 6 ? A3(a[r + c], a[r + c + 1], a[r + c + 2]) : a[7]);
        } else {
            for (c = 0; c < bs - 1; c++) O(0, c) = (uint8_t)A3(a[c], a[c + 1], a[c + 2]);
            O(0, bs - 1) = a[bs - 1];
            for (r = 1; r < bs; r++) for (c = 0; c < bs; c++) O(r, c) = c < bs - 1 - r ? O(0, r + c) : a[bs - 1];
        }
        break;
    case 8: /* D63 :43, 4x4 :268 (no replication of above[3]) */
        for (c = 0; c < bs; c++) { O(0, c) = (uint8_t)A2(a[c], a[c + 1]); O(1, c) = (uint8_t)A3(a[c], a[c + 1], a[c + 2]); }
        for (r = 2; r < bs; r++) {
            const int k = r >> 1, size = bs - 1 - k;
            for (c = 0; c < bs; c++) {
                if (bs == 4) O(r, c) = (uint8_t)((r & 1) ? A3(a[k + c], a[k + c + 1], a[k + c + 2]) : A2(a[k + c], a[k + c + 1]));
                else O(r, c) = c < size ? O(r & 1, k + c) : a[bs - 1];
            }
        }
        break;
    case 7: /* D207 :22 */
        for (r = 0; r < bs - 1; r++) O(r, 0) = (uint8_t)A2(l[r], l[r + 1]);
        O(bs - 1, 0) = l[bs - 1];
        for (r = 0; r < bs - 2; r++) O(r, 1) = (uint8_t)A3(l[r], l[r + 1], l[r + 2]);
        O(bs - 2, 1) = (uint8_t)A3(l[bs - 2], l[bs - 1], l[bs - 1]);
        O(bs - 1, 1) = l[bs - 1];
        for (c = 2; c < bs; c++) O(bs - 1, c) = l[bs - 1];
        for (r = bs - 2; r >= 0; r--) for (c = 2; c < bs; c++) O(r, c) = O(r + 1, c - 2);
        break;
    case 5: /* D117 :75 */
        for (c = 0; c < bs; c++) O(0, c) = (uint8_t)A2(a[c - 1], a[c]);
        O(1, 0) = (uint8_t)A3(l[0], a[-1], a[0]);
        for (c = 1; c < bs; c++) O(1, c) = (uint8_t)A3(a[c - 2], a[c - 1], a[c]);
        O(2, 0) = (uint8_t)A3(a[-1], l[0], l[1]);
        for (r = 3; r < bs; r++) O(r, 0) = (uint8_t)A3(l[r - 3], l[r - 2], l[r - 1]);
        for (r = 2; r < bs; r++) for (c = 1; c < bs; c++) O(r, c) = O(r - 2, c - 1);
        break;
    case 4: { /* D135 :98: the outer border from bottom-left to top-right, every row a window of it */
        uint8_t b[63];
        for (c = 0; c < bs - 2; c++) b[c] = (uint8_t)A3(l[bs - 3 - c], l[bs - 2 - c], l[bs - 1 - c]);
        b[bs - 2] = (uint8_t)A3(a[-1], l[0], l[1]);
        b[bs - 1] = (uint8_t)A3(l[0], a[-1], a[0]);
        b[bs] = (uint8_t)A3(a[-1], a[0], a[1]);
        for (c = 0; c < bs - 2; c++) b[bs + 1 + c] = (uint8_t)A3(a[c], a[c + 1], a[c + 2]);
        for (r = 0; r < bs; r++) for (c = 0; c < bs; c++) O(r, c) = b[bs - 1 - r + c];
        break;
    }
    case 6: /* D153 :117 */
        O(0, 0) = (uint8_t)A2(a[-1], l[0]);
        for (r = 1; r < bs; r++) O(r, 0) = (uint8_t)A2(l[r - 1], l[r]);
        O(0, 1) = (uint8_t)A3(l[0], a[-1], a[0]);
        O(1, 1) = (uint8_t)A3(a[-1], l[0], l[1]);
        for (r = 2; r < bs; r++) O(r, 1) = (uint8_t)A3(l[r - 2], l[r - 1], l[r]);
        for (c = 2; c < bs; c++) O(0, c) = (uint8_t)A3(a[c - 3], a[c - 2], a[c - 1]);
        for (r = 1; r < bs; r++) for (c = 2; c < bs; c++) O(r, c) = O(r - 1, c - 2);
        break;
    }
#undef O
}

void svtb200_oracle_intra_predict(const SvtB200IntraBlock *b, int n) {
    for (int i = 0; i < n; i++, b++) {
        uint8_t o[32 * 32];
        const int bs = 4 << b->tx_size;
        intra_one(b->mode, bs, b->have_left, b->have_top, b->above, b->left, o);
        for (int r = 0; r < bs; r++) memcpy(b->dst + (ptrdiff_t)r * b->dst_stride, o + r * 32, (size_t)bs);
    }
}

/* generate_intra_reference_samples() (Source/Lib/Codec/EbEncDecProcess.c:1128-1315) for one transform block, with the neighbour samples
 * read from the plane itself: the reference keeps them in its neighbour arrays (top_array[x] = the sample above the block at column x,
 * left_array[y] = the sample left of it in row y, top_left_array = the corner; EbEncDecProcess.c:1179-1218), which hold the plane's
 * samples at those positions whenever the block's neighbours have been written.  Sections NEED_LEFT (:1221), NEED_ABOVE (:1243) and
 * NEED_ABOVERIGHT (:1269: it rewrites the whole above row) restated with their frame-edge extension rules.
 * out: 128 bytes per job: [15] = above[-1], [16 .. 16 + 2 bs) = above, [80 .. 80 + bs) = left, everything else 0. */
void svtb200_oracle_intra_edges(const uint8_t *plane, int stride, int frame_width, int frame_height, const SvtB200IntraEdgeJob *jobs, int n, uint8_t *out) {
    for (int j = 0; j < n; j++) {
        const SvtB200IntraEdgeJob *b = &jobs[j];
        const int bs = 4 << b->tx_size, x0 = b->x0, y0 = b->y0;
        uint8_t *rec = out + (size_t)j * 128, *row = rec + 16, *left = rec + 80;
        const uint8_t *A = plane + (ptrdiff_t)(y0 - 1) * stride + x0; /* above_ref */
        memset(rec, 0, 128);
        if (b->have_left) {
            if (b->extend_bottom && y0 + bs > frame_height) {
                const int ext = frame_height - y0;
                int i = 0;
                for (; i < ext; i++) left[i] = plane[(ptrdiff_t)(y0 + i) * stride + x0 - 1];
                for (; i < bs; i++) left[i] = plane[(ptrdiff_t)(y0 + ext - 1) * stride + x0 - 1];
            } else
                for (int i = 0; i < bs; i++) left[i] = plane[(ptrdiff_t)(y0 + i) * stride + x0 - 1];
        } else
            memset(left, 129, (size_t)bs);
        if (!b->have_top) { memset(row - 1, 127, (size_t)(2 * bs + 1)); continue; }
        const int four_right = bs == 4 && b->have_right;
        if (b->extend_right) {
            if (x0 + 2 * bs <= frame_width) {
                if (four_right) memcpy(row, A, (size_t)(2 * bs));
                else { memcpy(row, A, (size_t)bs); memset(row + bs, row[bs - 1], (size_t)bs); }
            } else if (x0 + bs <= frame_width) {
                const int r = frame_width - x0;
                if (four_right) { memcpy(row, A, (size_t)r); memset(row + r, row[r - 1], (size_t)(x0 + 2 * bs - frame_width)); }
                else { memcpy(row, A, (size_t)bs); memset(row + bs, row[bs - 1], (size_t)bs); }
            } else {
                const int r = frame_width - x0;
                memcpy(row, A, (size_t)r);
                memset(row + r, row[r - 1], (size_t)(x0 + 2 * bs - frame_width));
            }
        } else {
            memcpy(row, A, (size_t)bs);
            if (four_right) memcpy(row + bs, A + bs, (size_t)bs); /* (with have_left the reference aliases above_ref itself: same samples) */
            else memset(row + bs, row[bs - 1], (size_t)bs);
        }
        row[-1] = b->have_left ? A[-1] : 129;
    }
}
