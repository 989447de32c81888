/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement of the chroma part of the reference's picture-analysis statistics:
 *   svtb200_oracle_chroma_block_mean   compute_chroma_block_mean / zero_out_chroma_block_mean
 *                                      (Source/Lib/Codec/EbPictureAnalysisProcess.c:1828-2090, :1760-1823; driver :4810-4822):
 *                                      per super-block the means of its 16 chroma 8x8 blocks (rows 0, 2, 4, 6 only:
 *                                      eb_vp9_compute_sub_mean8x8_sse2_intrin, ASM_SSE2/EbComputeMean_Intrinsic_SSE2.c:36), of the
 *                                      four 16x16 and of the 32x32 chroma block, in ME_TIER_ZERO_PU order (21 entries)
 *   svtb200_oracle_chroma_histogram    sub_sample_chroma_generate_pixel_intensity_histogram_bins (:4313-4440) + calculate_histogram
 *                                      (:127): 256-bin histograms of every 4th sample of every 4th row of the chroma planes
 *                                      per picture region
 * Pinned inside real encodes by the seam check (oracle/ref_seam_check.c: chroma_bad 0, chroma_hist_bad 0).
 */
#include <stdint.h>
#include <string.h>
#include "oracle.h"

/* sum of rows 0, 2, 4, 6 of an 8x8 block, << 3: the 8.8 fixed-point mean of the 32 samples */
static uint64_t sub_mean8x8(const uint8_t *s, int stride) {
    uint32_t sum = 0;
    for (int y = 0; y < 8; y += 2) for (int x = 0; x < 8; x++) sum += s[y * stride + x];
    return (uint64_t)sum << 3;
}

/* plane: the chroma plane at sample (0,0); (cx, cy): chroma position of the super-block; mean[21]: 64x64, 32x32 x 4, 16x16 x 16 (luma names) */
static void plane_means(const uint8_t *plane, int stride, int cx, int cy, uint8_t *mean) {
    uint64_t m16[16], m32[4], m64;
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) m16[r * 4 + c] = sub_mean8x8(plane + (size_t)(cy + 8 * r) * stride + cx + 8 * c, stride);
    m32[0] = (m16[0] + m16[1] + m16[4] + m16[5]) >> 2;
    m32[1] = (m16[2] + m16[3] + m16[6] + m16[7]) >> 2;
    m32[2] = (m16[8] + m16[9] + m16[12] + m16[13]) >> 2;
    m32[3] = (m16[10] + m16[11] + m16[14] + m16[15]) >> 2;
    m64 = (m32[0] + m32[1] + m32[3] + m32[3]) >> 2; /* the reference adds quadrant 3 twice and never quadrant 2 (:2011-2016) */
    mean[0] = (uint8_t)(m64 >> 8);
    for (int i = 0; i < 4; i++) mean[1 + i] = (uint8_t)(m32[i] >> 8);
    for (int i = 0; i < 16; i++) mean[5 + i] = (uint8_t)(m16[i] >> 8);
}

void svtb200_oracle_chroma_block_mean(const uint8_t *cb, const uint8_t *cr, int stride, int luma_width, int luma_height, uint8_t *cb_mean, uint8_t *cr_mean) {
    const int sbs_w = (luma_width + 63) / 64, sbs_h = (luma_height + 63) / 64;
    for (int sy = 0; sy < sbs_h; sy++)
        for (int sx = 0; sx < sbs_w; sx++) {
            uint8_t *mb = cb_mean + (size_t)(sy * sbs_w + sx) * 21, *mr = cr_mean + (size_t)(sy * sbs_w + sx) * 21;
            const int complete = sx * 64 + 64 <= luma_width && sy * 64 + 64 <= luma_height; /* SbParams::is_complete_sb */
            if (!complete) { memset(mb, 0, 21); memset(mr, 0, 21); continue; }
            plane_means(cb, stride, sx * 32, sy * 32, mb);
            plane_means(cr, stride, sx * 32, sy * 32, mr);
        }
}

/* hist: [rw][rh][2][256] (Cb, Cr), avg: [rw][rh][2], sum16[2]: sums over the regions of (region sum << 4) */
void svtb200_oracle_chroma_histogram(const uint8_t *cb, const uint8_t *cr, int stride, int luma_width, int luma_height, int rw, int rh,
                                     uint32_t *hist, uint8_t *avg, uint64_t sum16[2]) {
    const int regw = luma_width / rw, regh = luma_height / rh;
    sum16[0] = sum16[1] = 0;
    for (int iw = 0; iw < rw; iw++)
        for (int ih = 0; ih < rh; ih++) {
            const int w = regw + (iw == rw - 1 ? luma_width - rw * regw : 0), h = regh + (ih == rh - 1 ? luma_height - rh * regh : 0);
            for (int p = 0; p < 2; p++) {
                const uint8_t *s = (p ? cr : cb) + (size_t)((ih * regh) >> 1) * stride + ((iw * regw) >> 1);
                uint32_t *hg = hist + (((size_t)iw * rh + ih) * 2 + p) * 256;
                uint64_t sum = 0;
                for (int b = 0; b < 256; b++) hg[b] = 1;
                for (int y = 0; y < (h >> 1); y += 4)
                    for (int x = 0; x < (w >> 1); x += 4) { hg[s[(size_t)y * stride + x]]++; sum += s[(size_t)y * stride + x]; }
                sum <<= 4;
                sum16[p] += sum;
                avg[(iw * rh + ih) * 2 + p] = (uint8_t)((sum + (((uint64_t)w * h) >> 3)) / (((uint64_t)w * h) >> 2));
                for (int b = 0; b < 256; b++) hg[b] <<= 4;
            }
        }
}
