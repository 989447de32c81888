/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement of the block walk in front of the reference's inter predictors:
 *   svtb200_oracle_inter_expand   inter_prediction() -> build_inter_predictors() up to the inter_predictor() call
 *                                 (Source/Lib/Codec/EbIntraPrediction.c:49-73, Source/Lib/VPX/vp9_reconinter.c:102-250), with
 *                                 eb_vp9_clamp_mv_to_umv_border_sb (:68-87) and clamp_mv (vp9_mv.h:43) written out, and the
 *                                 mb_to_*_edge values of the encode pass (Source/Lib/Codec/EbEncDecProcess.c:2052-2066)
 *   svtb200_oracle_me_best_blocks the bench's open-loop partition (32x32 blocks, best motion-estimation candidate each;
 *                                 vector conversion of Source/Lib/Codec/EbModeDecision.c:222-227)
 * Pinned against the reference's own inter_prediction() through oracle/ref_path_driver.c (tests/test_blocks.py).
 */
#include <stdint.h>
#include <stddef.h>
#include "oracle.h"

typedef struct { int16_t row, col; } Mv;

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* eb_vp9_clamp_mv_to_umv_border_sb: the vector in 1/16 sample of the plane, limited so that the (bw x bh) block keeps within
 * VP9_INTERP_EXTEND + block size of the picture */
static Mv clamp_to_umv_border(int to_left, int to_right, int to_top, int to_bottom, Mv src, int bw, int bh, int ss_x, int ss_y) {
    const int spel_left = (4 + bw) << 4, spel_right = spel_left - 16;
    const int spel_top = (4 + bh) << 4, spel_bottom = spel_top - 16;
    Mv m;
    m.row = (int16_t)(src.row * (1 << (1 - ss_y)));
    m.col = (int16_t)(src.col * (1 << (1 - ss_x)));
    m.col = (int16_t)clampi(m.col, to_left * (1 << (1 - ss_x)) - spel_left, to_right * (1 << (1 - ss_x)) + spel_right);
    m.row = (int16_t)clampi(m.row, to_top * (1 << (1 - ss_y)) - spel_top, to_bottom * (1 << (1 - ss_y)) + spel_bottom);
    return m;
}

void svtb200_oracle_inter_expand(const SvtB200InterBlock *blocks, int n, const SvtB200PredPicture *pic, SvtB200PredBlock *out) {
    const int mi_cols = (pic->width + 7) >> 3, mi_rows = (pic->height + 7) >> 3;
    for (int i = 0; i < n; i++) {
        const SvtB200InterBlock *b = &blocks[i];
        const int mi_col = b->x >> 3, mi_row = b->y >> 3;
        const int to_top = -((mi_row * 8) * 8), to_bottom = ((mi_rows - (b->bh >> 3) - mi_row) * 8) * 8;
        const int to_left = -((mi_col * 8) * 8), to_right = ((mi_cols - (b->bw >> 3) - mi_col) * 8) * 8;
        const int compound = b->list[1] != 0xff;
        for (int plane = 0; plane < 3; plane++) {
            SvtB200PredBlock *o = &out[3 * i + plane];
            const int ss = plane > 0;
            const int bw = b->bw >> ss, bh = b->bh >> ss;
            const int ox = plane ? ((b->x >> 3) << 3) >> 1 : b->x, oy = plane ? ((b->y >> 3) << 3) >> 1 : b->y;
            for (int ref = 0; ref < 2; ref++) {
                const int used = ref < 1 + compound;
                const int list_index = used && b->list[ref] ? 1 : 0;
                Mv mv = {used ? b->mv[ref][0] : 0, used ? b->mv[ref][1] : 0};
                const Mv q4 = clamp_to_umv_border(to_left, to_right, to_top, to_bottom, mv, bw, bh, ss, ss);
                int row = q4.row, col = q4.col, subpel_x, subpel_y;
                if (pic->use_subpel_flag) { subpel_x = col & 15; subpel_y = row & 15; }
                else if (plane) { row = (row + 4) & ~0x07; col = (col + 4) & ~0x07; subpel_x = col & 0x07; subpel_y = row & 0x07; }
                else { row = (row + 8) & ~0x0F; col = (col + 8) & ~0x0F; subpel_x = col & 15; subpel_y = row & 15; }
                o->ref[ref] = pic->ref[list_index][plane];
                o->ref_stride[ref] = pic->ref_stride[list_index][plane];
                o->x[ref] = ox + (col >> 4);
                o->y[ref] = oy + (row >> 4);
                o->subpel_x[ref] = (uint8_t)subpel_x;
                o->subpel_y[ref] = (uint8_t)subpel_y;
            }
            o->dst = pic->dst[plane] + (ptrdiff_t)oy * pic->dst_stride[plane] + ox;
            o->dst_stride = pic->dst_stride[plane];
            o->w = (uint8_t)bw; o->h = (uint8_t)bh;
            o->compound = (uint8_t)compound;
            o->reserved = 0;
        }
    }
}

int svtb200_oracle_me_best_blocks(const SvtB200MeCuResult *me, int width, int height, SvtB200InterBlock *out) {
    const int sbs_w = (width + 63) / 64;
    int n = 0;
    /* pass 0: 32x32 blocks; pass 1: 16x16 blocks of the bottom strip; pass 2: of the right strip */
    for (int pass = 0; pass < 3; pass++) {
        const int size = pass ? 16 : 32;
        const int x_end = (width / 32) * 32, y_end = (height / 32) * 32;
        const int x0 = pass == 2 ? x_end : 0, x1 = pass == 0 ? x_end : width;
        const int y0 = pass == 1 ? y_end : 0, y1 = pass == 1 ? height : y_end;
        for (int y = y0; y + size <= y1; y += size)
            for (int x = x0; x + size <= x1; x += size) {
                const SvtB200MeCuResult *sb = me + (size_t)((y / 64) * sbs_w + x / 64) * 85;
                const SvtB200MeCuResult *r = size == 32 ? &sb[1 + ((y % 64) / 32) * 2 + (x % 64) / 32] : &sb[5 + ((y % 64) / 16) * 4 + (x % 64) / 16];
                SvtB200InterBlock *b = &out[n++];
                const unsigned dir = r->cand[0].direction;
                b->x = (uint16_t)x; b->y = (uint16_t)y; b->bw = b->bh = (uint8_t)size;
                if (dir == 2) {
                    b->list[0] = 0; b->list[1] = 1;
                    b->mv[0][0] = (int16_t)(r->y_mv_l0 << 1); b->mv[0][1] = (int16_t)(r->x_mv_l0 << 1);
                    b->mv[1][0] = (int16_t)(r->y_mv_l1 << 1); b->mv[1][1] = (int16_t)(r->x_mv_l1 << 1);
                } else {
                    b->list[0] = dir == 1; b->list[1] = 0xff;
                    b->mv[0][0] = (int16_t)((dir == 1 ? r->y_mv_l1 : r->y_mv_l0) << 1);
                    b->mv[0][1] = (int16_t)((dir == 1 ? r->x_mv_l1 : r->x_mv_l0) << 1);
                    b->mv[1][0] = b->mv[1][1] = 0;
                }
            }
    }
    return n;
}
