/* TEST INFRASTRUCTURE ONLY (oracle/_ref link shim) -- not product code.
 *
 * The reference ships six NASM files (Source/Lib/ASM_SSE2/*.asm, ASM_SSSE3/*.asm) that
 * cannot be assembled in this image (no yasm/nasm).  They export 59 symbols, none of which
 * is on the ME / transform / quant / deblock hot path.  This shim forwards each one to the
 * plain-C twin that already exists in the reference objects, so that libSvtVp9Enc_ref.so
 * links.  Generated once from the linker's undefined-symbol list (see oracle/build_ref.sh).
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>

uint64_t eb_vp9_Log2f_SSE2(uint64_t x) { return 63 - __builtin_clzll(x); } /* bsr */

void eb_vp9_subtract_block_c(int rows, int cols, int16_t *diff, ptrdiff_t diff_stride, const uint8_t *src,
                             ptrdiff_t src_stride, const uint8_t *pred, ptrdiff_t pred_stride);
void eb_vp9_subtract_block_sse2(int rows, int cols, int16_t *diff, ptrdiff_t diff_stride, const uint8_t *src,
                                ptrdiff_t src_stride, const uint8_t *pred, ptrdiff_t pred_stride) {
    eb_vp9_subtract_block_c(rows, cols, diff, diff_stride, src, src_stride, pred, pred_stride);
}

void eb_vp9_picture_copy_kernel_sse2(uint8_t *src, uint32_t src_stride, uint8_t *dst, uint32_t dst_stride,
                                     uint32_t w, uint32_t h) {
    for (uint32_t y = 0; y < h; y++) memcpy(dst + (size_t)y * dst_stride, src + (size_t)y * src_stride, w);
}

static uint8_t clip8(int v) { return v < 0 ? 0 : v > 255 ? 255 : (uint8_t)v; }
/* 2-tap (bilinear) 1-D filters: taps 3 and 4 of the 8-tap kernel, round 64, shift 7. */
#define BIL(name, W, STEP, AVG)                                                                      \
    void name(const uint8_t *src, ptrdiff_t src_pitch, uint8_t *out, ptrdiff_t out_pitch,           \
              uint32_t h, const int16_t *f) {                                                        \
        for (uint32_t y = 0; y < h; y++, src += src_pitch, out += out_pitch)                         \
            for (int x = 0; x < W; x++) {                                                            \
                int v = clip8((src[x] * f[3] + src[x + (STEP)] * f[4] + 64) >> 7);                   \
                out[x] = (AVG) ? (uint8_t)((out[x] + v + 1) >> 1) : (uint8_t)v;                      \
            }                                                                                        \
    }
void eb_vp9_d153_predictor_16x16_c(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
void eb_vp9_d153_predictor_16x16_ssse3(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left) { eb_vp9_d153_predictor_16x16_c(dst, stride, above, left); }
void eb_vp9_d153_predictor_32x32_c(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
void eb_vp9_d153_predictor_32x32_ssse3(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left) { eb_vp9_d153_predictor_32x32_c(dst, stride, above, left); }
void eb_vp9_d153_predictor_4x4_c(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
void eb_vp9_d153_predictor_4x4_ssse3(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left) { eb_vp9_d153_predictor_4x4_c(dst, stride, above, left); }
void eb_vp9_d153_predictor_8x8_c(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
void eb_vp9_d153_predictor_8x8_ssse3(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left) { eb_vp9_d153_predictor_8x8_c(dst, stride, above, left); }
void eb_vp9_d207_predictor_16x16_c(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
void eb_vp9_d207_predictor_16x16_ssse3(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left) { eb_vp9_d207_predictor_16x16_c(dst, stride, above, left); }
void eb_vp9_d207_predictor_32x32_c(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
void eb_vp9_d207_predictor_32x32_ssse3(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left) { eb_vp9_d207_predictor_32x32_c(dst, stride, above, left); }
void eb_vp9_d207_predictor_4x4_c(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
void eb_vp9_d207_predictor_4x4_sse2(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left) { eb_vp9_d207_predictor_4x4_c(dst, stride, above, left); }
void eb_vp9_d207_predictor_8x8_c(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
void eb_vp9_d207_predictor_8x8_ssse3(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left) { eb_vp9_d207_predictor_8x8_c(dst, stride, above, left); }
void eb_vp9_d45_predictor_16x16_c(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
void eb_vp9_d45_predictor_16x16_ssse3(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left) { eb_vp9_d45_predictor_16x16_c(dst, stride, above, left); }
void eb_vp9_d45_predictor_32x32_c(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
void eb_vp9_d45_predictor_32x32_ssse3(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left) { eb_vp9_d45_predictor_32x32_c(dst, stride, above, left); }
void eb_vp9_d45_predictor_4x4_c(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
void eb_vp9_d45_predictor_4x4_sse2(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left) { eb_vp9_d45_predictor_4x4_c(dst, stride, above, left); }
void eb_vp9_d45_predictor_8x8_c(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
void eb_vp9_d45_predictor_8x8_sse2(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left) { eb_vp9_d45_predictor_8x8_c(dst, stride, above, left); }
void eb_vp9_d63_predictor_16x16_c(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
void eb_vp9_d63_predictor_16x16_ssse3(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left) { eb_vp9_d63_predictor_16x16_c(dst, stride, above, left); }
void eb_vp9_d63_predictor_32x32_c(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
void eb_vp9_d63_predictor_32x32_ssse3(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left) { eb_vp9_d63_predictor_32x32_c(dst, stride, above, left); }
void eb_vp9_d63_predictor_4x4_c(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
void eb_vp9_d63_predictor_4x4_ssse3(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left) { eb_vp9_d63_predictor_4x4_c(dst, stride, above, left); }
void eb_vp9_d63_predictor_8x8_c(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
void eb_vp9_d63_predictor_8x8_ssse3(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left) { eb_vp9_d63_predictor_8x8_c(dst, stride, above, left); }
void eb_vp9_dc_128_predictor_16x16_c(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
void eb_vp9_dc_128_predictor_16x16_sse2(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left) { eb_vp9_dc_128_predictor_16x16_c(dst, stride, above, left); }
void eb_vp9_dc_128_predictor_32x32_c(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
void eb_vp9_dc_128_predictor_32x32_sse2(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left) { eb_vp9_dc_128_predictor_32x32_c(dst, stride, above, left); }
void eb_vp9_dc_128_predictor_4x4_c(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
void eb_vp9_dc_128_predictor_4x4_sse2(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left) { eb_vp9_dc_128_predictor_4x4_c(dst, stride, above, left); }
void eb_vp9_dc_128_predictor_8x8_c(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
void eb_vp9_dc_128_predictor_8x8_sse2(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left) { eb_vp9_dc_128_predictor_8x8_c(dst, stride, above, left); }
void eb_vp9_dc_left_predictor_16x16_c(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
void eb_vp9_dc_left_predictor_16x16_sse2(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left) { eb_vp9_dc_left_predictor_16x16_c(dst, stride, above, left); }
void eb_vp9_dc_left_predictor_32x32_c(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
void eb_vp9_dc_left_predictor_32x32_sse2(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left) { eb_vp9_dc_left_predictor_32x32_c(dst, stride, above, left); }
void eb_vp9_dc_left_predictor_4x4_c(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
void eb_vp9_dc_left_predictor_4x4_sse2(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left) { eb_vp9_dc_left_predictor_4x4_c(dst, stride, above, left); }
void eb_vp9_dc_left_predictor_8x8_c(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
void eb_vp9_dc_left_predictor_8x8_sse2(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left) { eb_vp9_dc_left_predictor_8x8_c(dst, stride, above, left); }
void eb_vp9_dc_predictor_16x16_c(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
void eb_vp9_dc_predictor_16x16_sse2(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left) { eb_vp9_dc_predictor_16x16_c(dst, stride, above, left); }
void eb_vp9_dc_predictor_32x32_c(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
void eb_vp9_dc_predictor_32x32_sse2(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left) { eb_vp9_dc_predictor_32x32_c(dst, stride, above, left); }
void eb_vp9_dc_predictor_4x4_c(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
void eb_vp9_dc_predictor_4x4_sse2(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left) { eb_vp9_dc_predictor_4x4_c(dst, stride, above, left); }
void eb_vp9_dc_predictor_8x8_c(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
void eb_vp9_dc_predictor_8x8_sse2(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left) { eb_vp9_dc_predictor_8x8_c(dst, stride, above, left); }
void eb_vp9_dc_top_predictor_16x16_c(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
void eb_vp9_dc_top_predictor_16x16_sse2(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left) { eb_vp9_dc_top_predictor_16x16_c(dst, stride, above, left); }
void eb_vp9_dc_top_predictor_32x32_c(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
void eb_vp9_dc_top_predictor_32x32_sse2(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left) { eb_vp9_dc_top_predictor_32x32_c(dst, stride, above, left); }
void eb_vp9_dc_top_predictor_4x4_c(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
void eb_vp9_dc_top_predictor_4x4_sse2(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left) { eb_vp9_dc_top_predictor_4x4_c(dst, stride, above, left); }
void eb_vp9_dc_top_predictor_8x8_c(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
void eb_vp9_dc_top_predictor_8x8_sse2(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left) { eb_vp9_dc_top_predictor_8x8_c(dst, stride, above, left); }
BIL(eb_vp9_filter_block1d16_h2_avg_ssse3, 16, 1, 1)
BIL(eb_vp9_filter_block1d16_h2_ssse3, 16, 1, 0)
BIL(eb_vp9_filter_block1d16_v2_avg_ssse3, 16, src_pitch, 1)
BIL(eb_vp9_filter_block1d16_v2_ssse3, 16, src_pitch, 0)
BIL(eb_vp9_filter_block1d4_h2_avg_ssse3, 4, 1, 1)
BIL(eb_vp9_filter_block1d4_h2_ssse3, 4, 1, 0)
BIL(eb_vp9_filter_block1d4_v2_avg_ssse3, 4, src_pitch, 1)
BIL(eb_vp9_filter_block1d4_v2_ssse3, 4, src_pitch, 0)
BIL(eb_vp9_filter_block1d8_h2_avg_ssse3, 8, 1, 1)
BIL(eb_vp9_filter_block1d8_h2_ssse3, 8, 1, 0)
BIL(eb_vp9_filter_block1d8_v2_avg_ssse3, 8, src_pitch, 1)
BIL(eb_vp9_filter_block1d8_v2_ssse3, 8, src_pitch, 0)
void eb_vp9_h_predictor_16x16_c(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
void eb_vp9_h_predictor_16x16_sse2(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left) { eb_vp9_h_predictor_16x16_c(dst, stride, above, left); }
void eb_vp9_h_predictor_32x32_c(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
void eb_vp9_h_predictor_32x32_sse2(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left) { eb_vp9_h_predictor_32x32_c(dst, stride, above, left); }
void eb_vp9_h_predictor_4x4_c(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
void eb_vp9_h_predictor_4x4_sse2(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left) { eb_vp9_h_predictor_4x4_c(dst, stride, above, left); }
void eb_vp9_h_predictor_8x8_c(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
void eb_vp9_h_predictor_8x8_sse2(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left) { eb_vp9_h_predictor_8x8_c(dst, stride, above, left); }
void eb_vp9_tm_predictor_16x16_c(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
void eb_vp9_tm_predictor_16x16_sse2(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left) { eb_vp9_tm_predictor_16x16_c(dst, stride, above, left); }
void eb_vp9_tm_predictor_32x32_c(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
void eb_vp9_tm_predictor_32x32_sse2(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left) { eb_vp9_tm_predictor_32x32_c(dst, stride, above, left); }
void eb_vp9_tm_predictor_4x4_c(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
void eb_vp9_tm_predictor_4x4_sse2(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left) { eb_vp9_tm_predictor_4x4_c(dst, stride, above, left); }
void eb_vp9_tm_predictor_8x8_c(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
void eb_vp9_tm_predictor_8x8_sse2(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left) { eb_vp9_tm_predictor_8x8_c(dst, stride, above, left); }
void eb_vp9_v_predictor_16x16_c(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
void eb_vp9_v_predictor_16x16_sse2(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left) { eb_vp9_v_predictor_16x16_c(dst, stride, above, left); }
void eb_vp9_v_predictor_32x32_c(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
void eb_vp9_v_predictor_32x32_sse2(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left) { eb_vp9_v_predictor_32x32_c(dst, stride, above, left); }
void eb_vp9_v_predictor_4x4_c(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
void eb_vp9_v_predictor_4x4_sse2(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left) { eb_vp9_v_predictor_4x4_c(dst, stride, above, left); }
void eb_vp9_v_predictor_8x8_c(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left);
void eb_vp9_v_predictor_8x8_sse2(uint8_t *dst, ptrdiff_t stride, const uint8_t *above, const uint8_t *left) { eb_vp9_v_predictor_8x8_c(dst, stride, above, left); }
