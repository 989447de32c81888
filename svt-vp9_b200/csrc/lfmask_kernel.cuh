/*
 * lfmask_kernel.cuh -- the edge masks of one picture from its mode-info grid: eb_vp9_build_mask_frame()
 * (Source/Lib/VPX/vp9_loopfilter.c:1548) -> eb_vp9_setup_mask (:901) -> build_masks (:677) / build_y_mask (:750).
 *
 * One thread per 64x64 super-block walks the partition quadtree (64 -> 32 -> 16 -> 8) exactly as the reference does and
 * ORs every block's contribution into a LOOP_FILTER_MASK held in registers / local memory.  The bit patterns of the
 * reference's lookup tables are VP9 format constants; here they are generated from the block geometry:
 *   prediction edge of a w8 x h8 block (8-sample units): left = bit 0 of each of its rows, above = its first row;
 *   transform edges inside a 64x64: columns / rows that are multiples of the transform size;
 *   chroma (4:2:0): the same on a 4x4 grid with halved block dimensions (at least one unit).
 * The file also compiles as plain C++ under csrc/emu/cuda_emu.h for the CPU-side logic tests.
 */
#ifndef SVTB200_LFMASK_KERNEL_CUH
#define SVTB200_LFMASK_KERNEL_CUH
#include <stdint.h>
#include <string.h>
#include "../../include/svtb200.h"

/* BLOCK_SIZE order of the format: 4x4 4x8 8x4 8x8 8x16 16x8 16x16 16x32 32x16 32x32 32x64 64x32 64x64 */
__device__ __forceinline__ int lfm_w8(int bs) { return bs < 5 ? 1 : bs < 8 ? 2 : bs < 11 ? 4 : 8; }   /* width in 8-sample units */
__device__ __forceinline__ int lfm_h8(int bs) {
    return bs < 4 ? 1 : bs == 4 ? 2 : bs == 5 ? 1 : bs == 6 ? 2 : bs == 7 ? 4 : bs == 8 ? 2 : bs == 9 ? 4 : bs == 10 ? 8 : bs == 11 ? 4 : 8;
}
__device__ __forceinline__ uint64_t lfm_rect(int w, int h, int pitch) { /* w x h ones on a grid with 'pitch' bits per row */
    uint64_t row = ((uint64_t)1 << w) - 1, m = 0;
    for (int r = 0; r < h; r++) m |= row << (r * pitch);
    return m;
}
/* columns (vertical == 1) or rows that start a transform block of 'tx8' units on an n x n grid */
__device__ __forceinline__ uint64_t lfm_tx_lines(int tx8, int n, int vertical) {
    uint64_t m = 0;
    for (int r = 0; r < n; r++)
        for (int c = 0; c < n; c++)
            if (((vertical ? c : r) % tx8) == 0) m |= (uint64_t)1 << (r * n + c);
    return m;
}
/* transform size of the chroma planes of a block (4:2:0): the luma transform, capped by the chroma block's smaller side */
__device__ __forceinline__ int lfm_uv_tx(int bs, int tx) {
    const int wpx = bs < 2 ? 4 : 8 * lfm_w8(bs), hpx = (bs == 0 || bs == 2) ? 4 : 8 * lfm_h8(bs);
    const int m = (wpx < hpx ? wpx : hpx) >> 1; /* smaller chroma side in samples (>= 2) */
    const int cap = m >= 32 ? 3 : m >= 16 ? 2 : m >= 8 ? 1 : 0;
    return tx < cap ? tx : cap;
}

struct LfmAcc { SvtB200LfMask m; };

__device__ static void lfm_add(LfmAcc &A, const SvtB200LfModeInfo &mi, int shift_y, int shift_uv, int with_uv) {
    const int bs = mi.sb_type, tx = mi.tx_size, lvl = mi.filter_level;
    if (!lvl) return;
    const int w = lfm_w8(bs), h = lfm_h8(bs);
    for (int r = 0; r < h; r++)
        for (int c = 0; c < w; c++) A.m.lfl_y[shift_y + r * 8 + c] = (uint8_t)lvl;
    const int wu = w > 1 ? w >> 1 : 1, hu = h > 1 ? h >> 1 : 1;
    const int txu = lfm_uv_tx(bs, tx);
    A.m.above_y[tx] |= lfm_rect(w, 1, 8) << shift_y;
    A.m.left_y[tx] |= lfm_rect(1, h, 8) << shift_y;
    if (with_uv) {
        A.m.above_uv[txu] |= (uint16_t)(lfm_rect(wu, 1, 4) << shift_uv);
        A.m.left_uv[txu] |= (uint16_t)(lfm_rect(1, hu, 4) << shift_uv);
    }
    if (mi.skip_inter) return; /* no coefficients and not intra: only the prediction edges */
    const int tx8 = tx < 2 ? 1 : tx == 2 ? 2 : 4, txu8 = txu < 2 ? 1 : txu == 2 ? 2 : 4;
    const uint64_t size = lfm_rect(w, h, 8), size_uv = lfm_rect(wu, hu, 4);
    A.m.above_y[tx] |= (size & lfm_tx_lines(tx8, 8, 0)) << shift_y;
    A.m.left_y[tx] |= (size & lfm_tx_lines(tx8, 8, 1)) << shift_y;
    if (tx == 0) A.m.int_4x4_y |= size << shift_y;
    if (with_uv) {
        A.m.above_uv[txu] |= (uint16_t)((size_uv & lfm_tx_lines(txu8, 4, 0)) << shift_uv);
        A.m.left_uv[txu] |= (uint16_t)((size_uv & lfm_tx_lines(txu8, 4, 1)) << shift_uv);
        if (txu == 0) A.m.int_4x4_uv |= (uint16_t)((size_uv & 0xffff) << shift_uv);
    }
}

/* eb_vp9_setup_mask for the super-block whose top-left mode-info unit is mi0 (row pitch 'stride' units) */
__device__ static void lfm_setup(const SvtB200LfModeInfo *mi0, int stride, int max_rows, int max_cols, SvtB200LfMask *out) {
    LfmAcc A;
    memset(&A.m, 0, sizeof(A.m)); /* vp9_zero(*lfm), padding included */
    auto at = [&](int r, int c) -> const SvtB200LfModeInfo & { return mi0[r * stride + c]; };
    const int bs0 = mi0->sb_type;
    if (bs0 == 12) lfm_add(A, at(0, 0), 0, 0, 1);
    else if (bs0 == 11) { lfm_add(A, at(0, 0), 0, 0, 1); if (4 < max_rows) lfm_add(A, at(4, 0), 32, 8, 1); }
    else if (bs0 == 10) { lfm_add(A, at(0, 0), 0, 0, 1); if (4 < max_cols) lfm_add(A, at(0, 4), 4, 2, 1); }
    else
        for (int q = 0; q < 4; q++) {
            const int r32 = (q >> 1) * 4, c32 = (q & 1) * 4;
            if (c32 >= max_cols || r32 >= max_rows) continue;
            const int sy = r32 * 8 + c32, su = (r32 >> 1) * 4 + (c32 >> 1);
            const int bs1 = at(r32, c32).sb_type;
            if (bs1 == 9) lfm_add(A, at(r32, c32), sy, su, 1);
            else if (bs1 == 8) { lfm_add(A, at(r32, c32), sy, su, 1); if (r32 + 2 < max_rows) lfm_add(A, at(r32 + 2, c32), sy + 16, su + 4, 1); }
            else if (bs1 == 7) { lfm_add(A, at(r32, c32), sy, su, 1); if (c32 + 2 < max_cols) lfm_add(A, at(r32, c32 + 2), sy + 2, su + 1, 1); }
            else
                for (int s = 0; s < 4; s++) {
                    const int r16 = r32 + (s >> 1) * 2, c16 = c32 + (s & 1) * 2;
                    if (c16 >= max_cols || r16 >= max_rows) continue;
                    const int sy16 = r16 * 8 + c16, su16 = (r16 >> 1) * 4 + (c16 >> 1);
                    const int bs2 = at(r16, c16).sb_type;
                    if (bs2 == 6) lfm_add(A, at(r16, c16), sy16, su16, 1);
                    else if (bs2 == 5) { lfm_add(A, at(r16, c16), sy16, su16, 1); if (r16 + 1 < max_rows) lfm_add(A, at(r16 + 1, c16), sy16 + 8, 0, 0); }
                    else if (bs2 == 4) { lfm_add(A, at(r16, c16), sy16, su16, 1); if (c16 + 1 < max_cols) lfm_add(A, at(r16, c16 + 1), sy16 + 1, 0, 0); }
                    else { /* four 8x8 (or smaller) units: chroma from the first only */
                        lfm_add(A, at(r16, c16), sy16, su16, 1);
                        for (int e = 1; e < 4; e++) {
                            const int r8 = r16 + (e >> 1), c8 = c16 + (e & 1);
                            if (c8 >= max_cols || r8 >= max_rows) continue;
                            lfm_add(A, at(r8, c8), r8 * 8 + c8, 0, 0);
                        }
                    }
                }
        }
    *out = A.m;
}

struct LfmArgs {
    const SvtB200LfModeInfo *mi;
    int32_t mi_rows, mi_cols, mi_stride;
    SvtB200LfMask *lfm;
    int32_t lfm_stride, sbs_w, n_sb;
};

__global__ void __launch_bounds__(64) lf_build_masks_kernel(const LfmArgs A) {
    const int sb = blockIdx.x * blockDim.x + threadIdx.x;
    if (sb >= A.n_sb) return;
    const int r = sb / A.sbs_w, c = sb - r * A.sbs_w, mi_row = r * 8, mi_col = c * 8;
    const int max_rows = A.mi_rows - mi_row < 8 ? A.mi_rows - mi_row : 8, max_cols = A.mi_cols - mi_col < 8 ? A.mi_cols - mi_col : 8;
    lfm_setup(A.mi + (size_t)mi_row * A.mi_stride + mi_col, A.mi_stride, max_rows, max_cols, A.lfm + (size_t)r * A.lfm_stride + c);
}
#endif
