/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement of the rest of the reference's picture-analysis front end:
 *   svtb200_oracle_noise_filter     eb_vp9_noise_extract_luma_weak / _luma_strong / _chroma_weak / _chroma_strong
 *                                   (Source/Lib/Codec/EbPictureAnalysisProcess.c:1653, :1450, :1575, :1498) applied to every super-block row of a
 *                                   plane, i.e. get_filtered_types (:1403) with types 0 / 4 / 4 / 6 on every sample that is not on the
 *                                   picture's border, border samples copied; the weak luma filter also leaves the noise plane
 *                                   clip(in - denoised, 0, 255)
 *   svtb200_oracle_noise_detect     detect_input_picture_noise (:3512-3560): per COMPLETE super-block compute_variance64x64 (:408) of the
 *                                   noise and of the denoised samples, the flat-noise flag, the picture sums
 *   svtb200_oracle_noise_detect_decimated   quarter_sample_detect_noise (:3809) / sub_sample_detect_noise (:3930): the same on the 1/4 / 1/16 plane with
 *                                   32x32 / 16x16 blocks (the methods the reference picks at 1080p / 4K for tune 1 above enc-mode 3)
 *   svtb200_oracle_edge_detection   edge_detection (:4545)
 *   svtb200_oracle_edge_cu          edge_detection_mean_luma_chroma16x16 (:4434)
 * Pinned against the reference's own functions (plain C and the AVX2 entries of its dispatch tables) through the path driver
 * (oracle/ref_path_driver.c: refdrv_noise_filter, refdrv_edge_detection, refdrv_edge_cu; tests/test_pa_noise_edge.py).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

/* get_filtered_types (:1403): p = centre sample */
static uint8_t filtered(const uint8_t *p, int s, int type) {
    uint32_t a;
    if (type == 0) a = (p[-s] + p[-1] + 4 * p[0] + p[1] + p[s]) / 8;
    else if (type == 4) a = (p[-s - 1] + 2 * p[-s] + p[-s + 1] + 2 * p[-1] + 4 * p[0] + 2 * p[1] + p[s - 1] + 2 * p[s] + p[s + 1]) / 16;
    else a = (4 * (p[-s - 1] + p[-s] + p[-s + 1] + p[-1] + p[0] + p[1] + p[s - 1] + p[s] + p[s + 1])) / 36;
    return (uint8_t)(a > 255 ? 255 : a);
}

/* kind: 0 luma weak, 1 luma strong, 2 chroma weak, 3 chroma strong; noise (luma weak only, may be NULL) has the geometry of den */
void svtb200_oracle_noise_filter(int kind, const uint8_t *in, int in_stride, int w, int h, uint8_t *den, int den_stride, uint8_t *noise) {
    const int type = kind == 0 ? 0 : kind == 3 ? 6 : 4;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const uint8_t *p = in + (size_t)y * in_stride + x;
            const int inner = y > 0 && y < h - 1 && x > 0 && x < w - 1;
            const uint8_t d = inner ? filtered(p, in_stride, type) : *p;
            den[(size_t)y * den_stride + x] = d;
            if (noise) noise[(size_t)y * den_stride + x] = (uint8_t)(inner && *p > d ? *p - d : 0);
        }
}

/* compute_variance64x64 (:408): the 64x64 point of the block statistics' hierarchy (8x8 means from rows 0, 2, 4, 6) BEFORE the precision shift */
static uint64_t variance64(const uint8_t *p, int stride) {
    uint64_t m[64], q[64], m16[16], q16[16], m32[4], q32[4];
    for (int b = 0; b < 64; b++) {
        const uint8_t *s = p + (size_t)(b >> 3) * 8 * stride + (b & 7) * 8;
        uint64_t a = 0, a2 = 0;
        for (int r = 0; r < 8; r += 2) for (int x = 0; x < 8; x++) { const uint64_t v = s[(size_t)r * stride + x]; a += v; a2 += v * v; }
        m[b] = a << 3; q[b] = a2 << 11;
    }
    for (int i = 0; i < 16; i++) { const int a = (2 * (i >> 2)) * 8 + 2 * (i & 3); m16[i] = (m[a] + m[a + 1] + m[a + 8] + m[a + 9]) >> 2; q16[i] = (q[a] + q[a + 1] + q[a + 8] + q[a + 9]) >> 2; }
    for (int i = 0; i < 4; i++) { const int a = (2 * (i >> 1)) * 4 + 2 * (i & 1); m32[i] = (m16[a] + m16[a + 1] + m16[a + 4] + m16[a + 5]) >> 2; q32[i] = (q16[a] + q16[a + 1] + q16[a + 4] + q16[a + 5]) >> 2; }
    const uint64_t m64 = (m32[0] + m32[1] + m32[2] + m32[3]) >> 2, q64 = (q32[0] + q32[1] + q32[2] + q32[3]) >> 2;
    return q64 - m64 * m64;
}

/* in: luma picture area; den: receives the weak-filtered picture.  Per super-block (0 for incomplete ones): noise_var = variance of
 * the noise samples (raw), den_var = variance of the denoised samples >> 16, flat = den_var < den_th && noise_var > noise_th;
 * totals[0] = sum of noise_var >> 16 over the complete super-blocks, totals[1] = their number. */
void svtb200_oracle_noise_detect(const uint8_t *in, int in_stride, int w, int h, uint8_t *den, int den_stride, uint64_t noise_th, uint64_t den_th,
                                 uint64_t *noise_var, uint64_t *den_var, uint8_t *flat, uint64_t totals[2]) {
    const int sbs_w = (w + 63) / 64, sbs_h = (h + 63) / 64;
    uint8_t noise[64 * 64];
    svtb200_oracle_noise_filter(0, in, in_stride, w, h, den, den_stride, NULL);
    totals[0] = totals[1] = 0;
    for (int sy = 0; sy < sbs_h; sy++)
        for (int sx = 0; sx < sbs_w; sx++) {
            const int sb = sy * sbs_w + sx;
            noise_var[sb] = den_var[sb] = 0; flat[sb] = 0;
            if (sx * 64 + 64 > w || sy * 64 + 64 > h) continue;
            for (int y = 0; y < 64; y++)
                for (int x = 0; x < 64; x++) {
                    const int py = sy * 64 + y, px = sx * 64 + x;
                    const int inner = py > 0 && py < h - 1 && px > 0 && px < w - 1;
                    const int d = (int)in[(size_t)py * in_stride + px] - (int)den[(size_t)py * den_stride + px];
                    noise[y * 64 + x] = (uint8_t)(inner && d > 0 ? d : 0);
                }
            noise_var[sb] = variance64(noise, 64);
            den_var[sb] = variance64(den + (size_t)sy * 64 * den_stride + sx * 64, den_stride) >> 16;
            flat[sb] = den_var[sb] < den_th && noise_var[sb] > noise_th;
            totals[0] += noise_var[sb] >> 16;
            totals[1]++;
        }
}

/* compute_variance32x32 (:152) / compute_variance16x16 (:345): the same hierarchy cut off at 32x32 / 16x16 */
static uint64_t variance_n(const uint8_t *p, int stride, int n /* 16, 32 or 64 */) {
    if (n == 64) return variance64(p, stride);
    uint64_t m[16], q[16];
    const int nb = n / 8;
    for (int b = 0; b < nb * nb; b++) {
        const uint8_t *s = p + (size_t)(b / nb) * 8 * stride + (b % nb) * 8;
        uint64_t a = 0, a2 = 0;
        /* compute_mean_func[0] / [1] (Codec/EbComputeMean.h:21; compute_mean, compute_mean_of_squared_values, C_DEFAULT/EbComputeMean_C.c:15-60): ALL
         * eight rows of the 8x8 block here -- (sum << 8) / 64 and (sum of squares << 16) / 64 --, not the rows 0, 2, 4, 6 of the 64x64 routine */
        for (int r = 0; r < 8; r++) for (int x = 0; x < 8; x++) { const uint64_t v = s[(size_t)r * stride + x]; a += v; a2 += v * v; }
        m[b] = a << 2; q[b] = a2 << 10;
    }
    if (n == 32) {
        uint64_t m16[4], q16[4];
        /* the reference groups the sixteen 8x8 blocks (numbered row by row, four per row) as {2i, 2i+1, 2i+8, 2i+9} -- the pairing of its 64x64
         * routine, whose rows hold eight blocks (:300-312): two blocks of block row 0 / 1 with the two below them two rows down, not one */
        for (int i = 0; i < 4; i++) { const int a = 2 * i; m16[i] = (m[a] + m[a + 1] + m[a + 8] + m[a + 9]) >> 2; q16[i] = (q[a] + q[a + 1] + q[a + 8] + q[a + 9]) >> 2; }
        const uint64_t mm = (m16[0] + m16[1] + m16[2] + m16[3]) >> 2, qq = (q16[0] + q16[1] + q16[2] + q16[3]) >> 2;
        return qq - mm * mm;
    }
    const uint64_t mm = (m[0] + m[1] + m[2] + m[3]) >> 2, qq = (q[0] + q[1] + q[2] + q[3]) >> 2;
    return qq - mm * mm;
}

/* quarter_sample_detect_noise (:3809, block 32 on the 1/4 plane) / sub_sample_detect_noise (:3930, block 16 on the 1/16 plane) up to their
 * picture-level decision.  in: the decimated plane (w x h); one block x block area = one full-resolution super-block, sb index
 * (by / block) * picture_width_in_sb + bx / block.  As the reference does it: the plane is walked in 64x64 areas, floor(w / 64) x
 * floor(h / 64) of them -- blocks further right / down are never visited (at 3840x2160 the complete super-block row 32 is not) --, each
 * 64-row band is weak-filtered into den (+ the band's noise rows), and the NOISE variance of every block of the band is taken from the
 * band's FIRST `block` noise rows at the block's columns (noise_origin_index carries no row offset for the block inside the band, :3866 /
 * :3988), the variance of the denoised samples from the block's own position.  Unvisited super-blocks: zeros. */
void svtb200_oracle_noise_detect_decimated(const uint8_t *in, int in_stride, int w, int h, int block, int picture_width_in_sb, int n_sb, uint8_t *den, int den_stride,
                                           uint64_t noise_th, uint64_t den_th, uint64_t *noise_var, uint64_t *den_var, uint8_t *flat, uint64_t totals[2]) {
    uint8_t *noise = (uint8_t *)malloc((size_t)w * 64);
    memset(noise_var, 0, (size_t)n_sb * 8); memset(den_var, 0, (size_t)n_sb * 8); memset(flat, 0, (size_t)n_sb);
    totals[0] = totals[1] = 0;
    svtb200_oracle_noise_filter(0, in, in_stride, w, h, den, den_stride, NULL);
    for (int by64 = 0; by64 < h / 64; by64++) {
        for (int y = 0; y < 64; y++) /* the band's noise rows */
            for (int x = 0; x < w; x++) {
                const int py = by64 * 64 + y;
                const int inner = py > 0 && py < h - 1 && x > 0 && x < w - 1;
                const int d = (int)in[(size_t)py * in_stride + x] - (int)den[(size_t)py * den_stride + x];
                noise[(size_t)y * w + x] = (uint8_t)(inner && d > 0 ? d : 0);
            }
        for (int bx64 = 0; bx64 < w / 64; bx64++)
            for (int vi = 0; vi < 64 / block; vi++)
                for (int hi = 0; hi < 64 / block; hi++) {
                    const int bx = bx64 * 64 + hi * block, by = by64 * 64 + vi * block;
                    if (bx + block > w || by + block > h) continue;
                    const int sb = (by / block) * picture_width_in_sb + bx / block;
                    if (sb >= n_sb) continue;
                    noise_var[sb] = variance_n(noise + bx, w, block); /* rows 0 .. block-1 of the band, whatever vi is */
                    den_var[sb] = variance_n(den + (size_t)by * den_stride + bx, den_stride, block) >> 16;
                    flat[sb] = den_var[sb] < den_th && noise_var[sb] > noise_th;
                    totals[0] += noise_var[sb] >> 16;
                    totals[1]++;
                }
    }
    free(noise);
}

/* edge_detection (:4545).  y_mean / variance: [n_sb][85] as the block statistics leave them.  The reference walks the super-blocks in
 * raster order, clears isolated_high_intensity_sb of the super-block it is at and THEN sets the 9x9 neighbourhood of a bright
 * super-block next to a dark one: marks a later super-block made on earlier ones survive, marks made on later ones are cleared
 * when the walk reaches them (unless made again).  Restated as that walk. */
void svtb200_oracle_edge_detection(const uint8_t *y_mean, const uint16_t *variance, int w, int h, uint64_t pic_avg_variance, uint8_t *edge_block_num,
                                   uint8_t *isolated, uint8_t *sharp_edge) {
    const int sbs_w = (w + 63) / 64, sbs_h = (h + 63) / 64;
    const uint64_t th0 = (pic_avg_variance * 70) / 100;
    for (int sb = 0; sb < sbs_w * sbs_h; sb++) {
        const int x = sb % sbs_w, y = sb / sbs_w;
        edge_block_num[sb] = 0; isolated[sb] = 0; sharp_edge[sb] = 0;
        if (!(x > 0 && x < sbs_w - 1 && y > 0 && y < sbs_h - 1)) continue;
        const uint16_t *v = variance + (size_t)sb * 85;
        edge_block_num[sb] = v[0] > th0;
        if (v[0] > 200) {
            int n = 0;
            for (int i = 5; i <= 20; i++) n += v[i] < 20;
            if (n > 4) sharp_edge[sb] = 1;
        }
        if (x > 3 && x < sbs_w - 4 && y > 3 && y < sbs_h - 4 && y_mean[(size_t)sb * 85] > 180) {
            const int dark = (y_mean[(size_t)(sb - 1) * 85] < 120) + (y_mean[(size_t)(sb + 1) * 85] < 120) + (y_mean[(size_t)(sb - sbs_w) * 85] < 120) +
                             (y_mean[(size_t)(sb + sbs_w) * 85] < 120);
            if (dark > 0)
                for (int i = -4; i < 5; i++)
                    for (int j = -4; j < 5; j++) isolated[sb + i * sbs_w + j] = 1;
        }
    }
}

/* edge_detection_mean_luma_chroma16x16 (:4434).  sb_flags[sb] = potential_logo_sb && is_complete_sb (SbParams, EbSequenceControlSet.c:332-380);
 * y_mean [n_sb][85], cb_mean / cr_mean [n_sb][21]; edge_cu [n_sb][16] (the sixteen 16x16 blocks in raster order); all zero when
 * picture_number is not a multiple of four. */
void svtb200_oracle_edge_cu(const uint8_t *y_mean, const uint8_t *cb_mean, const uint8_t *cr_mean, const uint8_t *sb_flags, int n_sb, uint64_t picture_number,
                            uint8_t *edge_cu) {
    memset(edge_cu, 0, (size_t)n_sb * 16);
    if (picture_number & 3) return;
    uint32_t max_grad = 1;
    static uint16_t grad_store[1 << 16][16];
    uint16_t(*grad)[16] = grad_store;
    if (n_sb > (1 << 16)) return;
    for (int sb = 0; sb < n_sb; sb++) {
        if (!sb_flags[sb]) continue;
        const uint8_t *ym = y_mean + (size_t)sb * 85, *cr = cr_mean + (size_t)sb * 21, *cb = cb_mean + (size_t)sb * 21;
        for (int k = 5; k <= 20; k++) {
            const int b = k - 5, x = b & 3, y = b >> 2;
            int gx = 0, gy = 0, nx = 0, ny = 0;
#define AD(a, i, j) ((a)[i] > (a)[j] ? (a)[i] - (a)[j] : (a)[j] - (a)[i])
            if (x != 0) { gx += AD(ym, k, k - 1) + AD(cr, k, k - 1) + AD(cb, k, k - 1); nx++; }
            if (x != 3) { gx += AD(ym, k + 1, k) + AD(cr, k + 1, k) + AD(cb, k + 1, k); nx++; }
            if (y != 0) { gy += AD(ym, k, k - 4) + AD(cr, k, k - 4) + AD(cb, k, k - 4); ny++; }
            if (y != 3) { gy += AD(ym, k + 4, k) + AD(cr, k + 4, k) + AD(cb, k + 4, k); ny++; }
#undef AD
            grad[sb][b] = (uint16_t)(gx / nx + gy / ny);
            if (grad[sb][b] > max_grad) max_grad = grad[sb][b];
        }
    }
    for (int sb = 0; sb < n_sb; sb++)
        if (sb_flags[sb])
            for (int b = 0; b < 16; b++) {
                const uint32_t v = (grad[sb][b] * (255u * 3)) / max_grad;
                edge_cu[(size_t)sb * 16 + b] = (uint16_t)(v < 255 ? v : 255) < 30 ? 0 : 1;
            }
}

/* determine_homogeneous_region_in_picture (:4660): the variance of the 8x8 variances inside each 32x32 quadrant and inside the 64x64 of every
 * COMPLETE super-block (incomplete: all ones / homogeneous), the homogeneous-area flag (64x64 variance of variance <= detail_th,
 * VAR_BASED_DETAIL_PRESERVATION_SELECTOR_THRSLHD), and the picture counts very_low_var_pic_flag / logo_pic_flag are thresholds of:
 * counts[0] = complete super-blocks, counts[1] = those whose 64x64 variance is below low_var_th (SB_LOW_VAR_TH), counts[2] = those with variance 0.
 * variance: [n_sb][85] as the block statistics leave them (8x8 entries 21..84, row-major 8 per row). */
void svtb200_oracle_homogeneous_regions(const uint16_t *variance, int w, int h, uint32_t low_var_th, uint64_t detail_th, uint64_t *var_of_var32, uint8_t *homogeneous,
                                        uint64_t counts[3]) {
    const int sbs_w = (w + 63) / 64, sbs_h = (h + 63) / 64;
    counts[0] = counts[1] = counts[2] = 0;
    for (int sb = 0; sb < sbs_w * sbs_h; sb++) {
        const uint16_t *v = variance + (size_t)sb * 85;
        homogeneous[sb] = 1;
        if ((sb % sbs_w) * 64 + 64 > w || (sb / sbs_w) * 64 + 64 > h) { for (int k = 0; k < 4; k++) var_of_var32[(size_t)sb * 4 + k] = ~(uint64_t)0; continue; }
        counts[0]++; counts[1] += v[0] < low_var_th; counts[2] += v[0] == 0;
        uint64_t s64 = 0, q64 = 0;
        for (int k = 0; k < 4; k++) {
            uint64_t s = 0, q = 0;
            for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) { const uint64_t x = v[21 + ((k >> 1) * 4 + r) * 8 + (k & 1) * 4 + c]; s += x; q += x * x; }
            s >>= 4; q >>= 4;
            var_of_var32[(size_t)sb * 4 + k] = q - s * s;
        }
        for (int i = 21; i < 85; i++) { s64 += v[i]; q64 += (uint64_t)v[i] * v[i]; }
        s64 >>= 6; q64 >>= 6;
        if (q64 - s64 * s64 > detail_th) homogeneous[sb] = 0;
    }
}
