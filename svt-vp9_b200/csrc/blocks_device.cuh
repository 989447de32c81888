/*
 * blocks_device.cuh -- per-block arithmetic of blocks.cu (host + device functions; also compiled by g++ under csrc/emu/cuda_emu.h for
 * the CPU-side logic test, tests/emu/blocks_emu.cpp).
 */
#ifndef SVTB200_BLOCKS_DEVICE_CUH
#define SVTB200_BLOCKS_DEVICE_CUH

#include <stdint.h>
#include "../../include/svtb200.h"

#ifndef SVTB200_HD
#if defined(__CUDACC__)
#define SVTB200_HD __host__ __device__ __forceinline__
#else
#define SVTB200_HD static inline
#endif
#endif

#define BLK_INTERP_EXTEND 4 /* VP9_INTERP_EXTEND, Source/Lib/VPX/yv12config.h:25 */
#define BLK_SUBPEL_BITS 4   /* SUBPEL_BITS, Source/Lib/VPX/vpx_filter.h:22 */

SVTB200_HD int blk_clamp(int v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : v; }

/* one plane of one block: build_inter_predictors() up to the inter_predictor() call */
SVTB200_HD SvtB200PredBlock blocks_expand_one(const SvtB200InterBlock &B, int plane, const SvtB200PredPicture &pic) {
    SvtB200PredBlock o;
    const int ss = plane ? 1 : 0;
    const int bw = B.bw >> ss, bh = B.bh >> ss; /* 4 * num_4x4 of bsize / bsize_uv */
    /* mb_to_*_edge in 1/8 luma sample (EbEncDecProcess.c:2052-2066); mi_cols / mi_rows: the picture rounded up to 8 samples */
    const int mi_w = (pic.width + 7) & ~7, mi_h = (pic.height + 7) & ~7;
    const int to_left = -((int)B.x * 8), to_right = (mi_w - B.bw - (int)B.x) * 8;
    const int to_top = -((int)B.y * 8), to_bottom = (mi_h - B.bh - (int)B.y) * 8;
    const int spel_left = (BLK_INTERP_EXTEND + bw) << BLK_SUBPEL_BITS, spel_right = spel_left - (1 << BLK_SUBPEL_BITS);
    const int spel_top = (BLK_INTERP_EXTEND + bh) << BLK_SUBPEL_BITS, spel_bottom = spel_top - (1 << BLK_SUBPEL_BITS);
    const int sc = 1 << (1 - ss);
    /* position of the block in the plane: luma block origin, chroma (origin & ~7) >> 1 (ROUND_UV) */
    const int px = plane ? ((int)B.x & ~7) >> 1 : (int)B.x, py = plane ? ((int)B.y & ~7) >> 1 : (int)B.y;
    o.compound = B.list[1] != 0xff;
    for (int r = 0; r < 2; r++) {
        const int live = r == 0 || o.compound;
        const int list = live ? (B.list[r] ? 1 : 0) : 0;
        /* MV is two int16: the scaled vector wraps like the reference's struct assignment */
        int row = (int16_t)((live ? B.mv[r][0] : 0) * sc), col = (int16_t)((live ? B.mv[r][1] : 0) * sc);
        col = (int16_t)blk_clamp(col, to_left * sc - spel_left, to_right * sc + spel_right);
        row = (int16_t)blk_clamp(row, to_top * sc - spel_top, to_bottom * sc + spel_bottom);
        int sx, sy;
        if (pic.use_subpel_flag) { sx = col & 15; sy = row & 15; }
        else if (plane) { row = (row + 4) & ~7; col = (col + 4) & ~7; sx = col & 7; sy = row & 7; }
        else { row = (row + 8) & ~15; col = (col + 8) & ~15; sx = col & 15; sy = row & 15; }
        o.ref[r] = pic.ref[list][plane];
        o.ref_stride[r] = pic.ref_stride[list][plane];
        o.x[r] = px + (col >> BLK_SUBPEL_BITS); o.y[r] = py + (row >> BLK_SUBPEL_BITS);
        o.subpel_x[r] = (uint8_t)sx; o.subpel_y[r] = (uint8_t)sy;
    }
    o.dst = pic.dst[plane] + (ptrdiff_t)py * pic.dst_stride[plane] + px;
    o.dst_stride = pic.dst_stride[plane];
    o.w = (uint8_t)bw; o.h = (uint8_t)bh;
    o.reserved = 0;
    return o;
}

/* the open-loop tiling: 32x32 blocks in raster order, then the 16x16 blocks of the bottom strip, then those of the right strip */
SVTB200_HD int blocks_tile_count(int width, int height) {
    const int w32 = width >> 5, h32 = height >> 5, w16 = width >> 4, h16 = height >> 4;
    return w32 * h32 + (h16 - 2 * h32) * w16 + (w16 - 2 * w32) * 2 * h32;
}
SVTB200_HD void blocks_tile_geometry(int width, int height, int i, int *x, int *y, int *size) {
    const int w32 = width >> 5, h32 = height >> 5, w16 = width >> 4, h16 = height >> 4;
    const int n32 = w32 * h32, nbot = (h16 - 2 * h32) * w16;
    if (i < n32) { *y = (i / w32) * 32; *x = (i - (i / w32) * w32) * 32; *size = 32; }
    else if (i < n32 + nbot) { const int j = i - n32; *y = (2 * h32 + j / w16) * 16; *x = (j - (j / w16) * w16) * 16; *size = 16; }
    else { const int j = i - n32 - nbot, wr = w16 - 2 * w32; *y = (j / wr) * 16; *x = (2 * w32 + j - (j / wr) * wr) * 16; *size = 16; }
}
/* the best motion-estimation candidate of a partition as an inter block (vectors: quarter-pel -> 1/8, EbModeDecision.c:222-227) */
SVTB200_HD SvtB200InterBlock blocks_from_me(const SvtB200MeCuResult &r, int x, int y, int size) {
    SvtB200InterBlock b;
    const uint32_t d = r.cand[0].direction;
    b.x = (uint16_t)x; b.y = (uint16_t)y; b.bw = b.bh = (uint8_t)size;
    b.list[0] = d == 1 ? 1 : 0;
    b.list[1] = d == 2 ? 1 : 0xff;
    b.mv[0][0] = (int16_t)((d == 1 ? r.y_mv_l1 : r.y_mv_l0) << 1); b.mv[0][1] = (int16_t)((d == 1 ? r.x_mv_l1 : r.x_mv_l0) << 1);
    b.mv[1][0] = (int16_t)(d == 2 ? r.y_mv_l1 << 1 : 0); b.mv[1][1] = (int16_t)(d == 2 ? r.x_mv_l1 << 1 : 0);
    return b;
}

#endif
