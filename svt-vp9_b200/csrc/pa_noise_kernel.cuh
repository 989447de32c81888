/* pa_noise_kernel.cuh -- device code of pa_noise.cu (noise filters / noise detection statistics / edge maps of the picture-analysis
 * stage); a header of its own so that tests/emu/pa_emu.cpp can run the same source on the CPU. */
#ifndef SVTB200_PA_NOISE_KERNEL_CUH
#define SVTB200_PA_NOISE_KERNEL_CUH
#include <stdint.h>
#include <stddef.h>

/* get_filtered_types (:1403) on the sample at p: type 0 = cross (1 1 4 1 1) / 8, 4 = 3x3 gaussian / 16, 6 = 3x3 box (4 * sum / 36) */
template <int TYPE>
__device__ __forceinline__ uint32_t nz_filtered(const uint8_t *p, int s) {
    const uint32_t c = __ldg(p), n = __ldg(p - s), so = __ldg(p + s), w = __ldg(p - 1), e = __ldg(p + 1);
    if (TYPE == 0) return (n + w + 4 * c + e + so) >> 3;
    const uint32_t nw = __ldg(p - s - 1), ne = __ldg(p - s + 1), sw = __ldg(p + s - 1), se = __ldg(p + s + 1);
    if (TYPE == 4) return (nw + ne + sw + se + 2 * (n + w + e + so) + 4 * c) >> 4;
    return (4 * (nw + n + ne + w + c + e + sw + so + se)) / 36; /* <= 255 */
}

/* one thread per four horizontally adjacent samples; border samples of the plane are copied, their noise is 0 */
template <int TYPE>
__global__ void __launch_bounds__(128) noise_filter_kernel(const uint8_t *in, int is, int w, int h, uint8_t *den, int ds, uint8_t *noise) {
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y;
    if (x0 >= w) return;
    const bool row_inner = y > 0 && y < h - 1;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int x = x0 + k;
        if (x >= w) break;
        const uint8_t *p = in + (size_t)y * is + x;
        const uint32_t v = __ldg(p);
        const bool inner = row_inner && x > 0 && x < w - 1;
        const uint32_t d = inner ? nz_filtered<TYPE>(p, is) : v;
        den[(size_t)y * ds + x] = (uint8_t)d;
        if (noise) noise[(size_t)y * ds + x] = (uint8_t)(inner && v > d ? v - d : 0);
    }
}

/* ---- detect_input_picture_noise (:3512) and its decimated twins quarter_sample_detect_noise (:3809, block 32 on the 1/4 plane) /
 * sub_sample_detect_noise (:3930, block 16 on the 1/16 plane): weak filter + the variances of the noise and of the denoised samples of every
 * block x block area (= one full-resolution super-block), one CTA per 64x64 area of the plane.  What the reference does differently for
 * block < 64, reproduced because sb_flat_noise_array depends on it: (1) only floor(w / 64) x floor(h / 64) areas are walked (grid);
 * (2) the 8x8 means take ALL eight rows (compute_mean_func, EbComputeMean.h:21) where the 64x64 routine takes rows 0, 2, 4, 6;
 * (3) compute_variance32x32 pairs the 8x8 blocks {2i, 2i+1, 2i+8, 2i+9} of its 4-wide numbering (:300-312); (4) the NOISE variance of
 * every block of a 64-row band is taken from the band's first `block` noise rows (noise_origin_index has no row offset inside the band). ---- */
struct NoiseDetArgs {
    const uint8_t *in; int32_t is, w, h;
    uint8_t *den; int32_t ds;
    int32_t sbs_w;            /* 64x64 areas per row of the grid */
    int32_t block, pw, n_sb;  /* block 64: sb = area index; block 32 / 16: sb = (by / block) * pw + bx / block */
    unsigned long long noise_th, den_th;
    unsigned long long *noise_var, *den_var, *totals;
    uint8_t *flat;
};
/* variance of sub-block (vi, hi) of the tile from the 8x8 tables (64 entries, 8 per row); noise_rows: take block row 0 of the band (quirk 4) */
__device__ __forceinline__ unsigned long long nz_block_variance(const unsigned long long *m, const unsigned long long *q, int block, int vi, int hi) {
    if (block == 64) {
        unsigned long long m16[16], q16[16], m32[4], q32[4];
        for (int i = 0; i < 16; i++) { const int a = (2 * (i >> 2)) * 8 + 2 * (i & 3); m16[i] = (m[a] + m[a + 1] + m[a + 8] + m[a + 9]) >> 2; q16[i] = (q[a] + q[a + 1] + q[a + 8] + q[a + 9]) >> 2; }
        for (int i = 0; i < 4; i++) { const int a = (2 * (i >> 1)) * 4 + 2 * (i & 1); m32[i] = (m16[a] + m16[a + 1] + m16[a + 4] + m16[a + 5]) >> 2; q32[i] = (q16[a] + q16[a + 1] + q16[a + 4] + q16[a + 5]) >> 2; }
        const unsigned long long mm = (m32[0] + m32[1] + m32[2] + m32[3]) >> 2, qq = (q32[0] + q32[1] + q32[2] + q32[3]) >> 2;
        return qq - mm * mm;
    }
    if (block == 32) {
        unsigned long long m16[4], q16[4];
#define NZ_AT(k) (((vi) * 4 + ((k) >> 2)) * 8 + (hi) * 4 + ((k) & 3)) /* 8x8 block k (four per row) of the 32x32 block, in the tile's table */
        for (int i = 0; i < 4; i++) {
            const int a = 2 * i;
            m16[i] = (m[NZ_AT(a)] + m[NZ_AT(a + 1)] + m[NZ_AT(a + 8)] + m[NZ_AT(a + 9)]) >> 2;
            q16[i] = (q[NZ_AT(a)] + q[NZ_AT(a + 1)] + q[NZ_AT(a + 8)] + q[NZ_AT(a + 9)]) >> 2;
        }
#undef NZ_AT
        const unsigned long long mm = (m16[0] + m16[1] + m16[2] + m16[3]) >> 2, qq = (q16[0] + q16[1] + q16[2] + q16[3]) >> 2;
        return qq - mm * mm;
    }
    const int a = (vi * 2) * 8 + hi * 2;
    const unsigned long long mm = (m[a] + m[a + 1] + m[a + 8] + m[a + 9]) >> 2, qq = (q[a] + q[a + 1] + q[a + 8] + q[a + 9]) >> 2;
    return qq - mm * mm;
}
__global__ void __launch_bounds__(256) noise_detect_kernel(const NoiseDetArgs A) {
    __shared__ uint8_t den_s[64][64 + 4], nz_s[64][64 + 4];
    __shared__ unsigned long long m_s[2][64], q_s[2][64]; /* [noise, denoised][8x8 block]: 8.8 means, 16.16 means of squares */
    __shared__ unsigned long long var_s[2][16];
    const int tile = blockIdx.x, sx = (tile % A.sbs_w) * 64, sy = (tile / A.sbs_w) * 64, t = threadIdx.x;
    const bool complete = sx + 64 <= A.w && sy + 64 <= A.h;
    for (int i = t; i < 4096; i += 256) {
        const int y = i >> 6, x = i & 63, px = sx + x, py = sy + y;
        if (px >= A.w || py >= A.h) continue;
        const uint8_t *p = A.in + (size_t)py * A.is + px;
        const uint32_t v = __ldg(p);
        const bool inner = py > 0 && py < A.h - 1 && px > 0 && px < A.w - 1;
        const uint32_t d = inner ? nz_filtered<0>(p, A.is) : v;
        A.den[(size_t)py * A.ds + px] = (uint8_t)d;
        den_s[y][x] = (uint8_t)d;
        nz_s[y][x] = (uint8_t)(inner && v > d ? v - d : 0);
    }
    if (!complete) { /* "do it only for complete 64x64 blocks" (:3549); the decimated walks never get here (their grid holds whole areas only) */
        if (t == 0 && A.block == 64) { A.noise_var[tile] = 0; A.den_var[tile] = 0; A.flat[tile] = 0; }
        return;
    }
    __syncthreads();
    const bool all_rows = A.block != 64;
    if (t < 128) { /* 8x8 sums: rows 0, 2, 4, 6 (eb_vp9_compute_sub_mean8x8 ...) for the 64x64 routine, every row (compute_mean ...) for the others */
        const int which = t >> 6, b = t & 63, bx = (b & 7) * 8, by = (b >> 3) * 8;
        uint32_t sum = 0, sq = 0;
#pragma unroll
        for (int r = 0; r < 8; r++) {
            if ((r & 1) && !all_rows) continue;
#pragma unroll
            for (int k = 0; k < 8; k++) { const uint32_t v = which ? den_s[by + r][bx + k] : nz_s[by + r][bx + k]; sum += v; sq += v * v; }
        }
        m_s[which][b] = (unsigned long long)sum << (all_rows ? 2 : 3);
        q_s[which][b] = (unsigned long long)sq << (all_rows ? 10 : 11);
    }
    __syncthreads();
    const int nb = 64 / A.block, nsub = nb * nb; /* sub-blocks of the area */
    if (t < 2 * nsub) {
        const int which = t / nsub, k = t % nsub, vi = k / nb, hi = k % nb;
        var_s[which][k] = nz_block_variance(m_s[which], q_s[which], A.block, which ? vi : 0, hi); /* noise: the band's first block row */
    }
    __syncthreads();
    if (t < nsub) {
        const int vi = t / nb, hi = t % nb;
        const int sb = A.block == 64 ? tile : ((sy / A.block) + vi) * A.pw + (sx / A.block) + hi;
        if (sb < A.n_sb) {
            const unsigned long long nv = var_s[0][t], dv = var_s[1][t] >> 16;
            A.noise_var[sb] = nv; A.den_var[sb] = dv;
            A.flat[sb] = dv < A.den_th && nv > A.noise_th;
            atomicAdd(&A.totals[0], nv >> 16);
            atomicAdd(&A.totals[1], 1ull);
        }
    }
}

/* ---- edge_detection (:4545) ---- */
__global__ void __launch_bounds__(1024) pa_avg_variance_kernel(const uint16_t *variance, int n_sb, uint16_t *avg) {
    __shared__ unsigned long long part[32];
    unsigned long long s = 0;
    for (int i = threadIdx.x; i < n_sb; i += blockDim.x) s += variance[(size_t)i * 85];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { for (int i = 1; i < (int)(blockDim.x >> 5); i++) s += part[i]; *avg = (uint16_t)(s / (unsigned long long)n_sb); } /* compute_picture_spatial_statistics :4827 */
}
__device__ __forceinline__ bool edge_bright_beside_dark(const uint8_t *y_mean, int sb, int x, int y, int sbs_w, int sbs_h) {
    if (!(x > 3 && x < sbs_w - 4 && y > 3 && y < sbs_h - 4)) return false;
    if (!(y_mean[(size_t)sb * 85] > 180)) return false;
    return y_mean[(size_t)(sb - 1) * 85] < 120 || y_mean[(size_t)(sb + 1) * 85] < 120 || y_mean[(size_t)(sb - sbs_w) * 85] < 120 || y_mean[(size_t)(sb + sbs_w) * 85] < 120;
}
__global__ void edge_detection_kernel(const uint8_t *y_mean, const uint16_t *variance, int sbs_w, int sbs_h, const uint16_t *avg, uint8_t *edge_block_num,
                                      uint8_t *isolated, uint8_t *sharp) {
    const int sb = blockIdx.x * blockDim.x + threadIdx.x;
    if (sb >= sbs_w * sbs_h) return;
    const int x = sb % sbs_w, y = sb / sbs_w;
    uint8_t ebn = 0, sh = 0, iso = 0;
    if (x > 0 && x < sbs_w - 1 && y > 0 && y < sbs_h - 1) {
        const uint16_t *v = variance + (size_t)sb * 85;
        const unsigned long long th0 = ((unsigned long long)*avg * 70) / 100;
        ebn = v[0] > th0;
        if (v[0] > 200) {
            int n = 0;
            for (int i = 5; i <= 20; i++) n += v[i] < 20;
            sh = n > 4;
        }
    }
    /* The reference walks the super-blocks in raster order, clears the flag of the one it is at and then marks the 9x9 neighbourhood of a
     * bright super-block beside a dark one: what stays set on super-block n are the marks made from n itself or from a LATER super-block. */
    for (int j = -4; j <= 4 && !iso; j++)
        for (int i = -4; i <= 4; i++) {
            const int mx = x + i, my = y + j, m = my * sbs_w + mx;
            if (mx < 0 || mx >= sbs_w || my < 0 || my >= sbs_h || m < sb) continue;
            if (edge_bright_beside_dark(y_mean, m, mx, my, sbs_w, sbs_h)) { iso = 1; break; }
        }
    edge_block_num[sb] = ebn; sharp[sb] = sh; isolated[sb] = iso;
}

/* ---- edge_detection_mean_luma_chroma16x16 (:4434): gradient of the 16x16 means, relative to the picture's largest ---- */
__device__ __forceinline__ uint32_t edge_cu_grad(const uint8_t *ym, const uint8_t *cb, const uint8_t *cr, int b) {
    const int k = 5 + b, x = b & 3, y = b >> 2;
    int gx = 0, gy = 0, nx = 0, ny = 0;
#define AD(a, i, j) abs((int)(a)[i] - (int)(a)[j])
    if (x != 0) { gx += AD(ym, k, k - 1) + AD(cr, k, k - 1) + AD(cb, k, k - 1); nx++; }
    if (x != 3) { gx += AD(ym, k + 1, k) + AD(cr, k + 1, k) + AD(cb, k + 1, k); nx++; }
    if (y != 0) { gy += AD(ym, k, k - 4) + AD(cr, k, k - 4) + AD(cb, k, k - 4); ny++; }
    if (y != 3) { gy += AD(ym, k + 4, k) + AD(cr, k + 4, k) + AD(cb, k + 4, k); ny++; }
#undef AD
    return (uint32_t)(gx / nx + gy / ny) & 0xffffu;
}
/* pass 0: the picture's largest gradient (at least 1); pass 1: the flags.  One thread per (super-block, 16x16 block). */
template <int PASS>
__global__ void edge_cu_kernel(const uint8_t *y_mean, const uint8_t *cb_mean, const uint8_t *cr_mean, const uint8_t *logo, int w, int h, int sbs_w, int n_sb,
                               uint32_t *max_grad, uint8_t *edge_cu) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, sb = i >> 4, b = i & 15;
    const bool valid = sb < n_sb; /* (no early return: the warp reduction below needs every lane) */
    const int sx = valid ? (sb % sbs_w) * 64 : 0, sy = valid ? (sb / sbs_w) * 64 : 0;
    const bool on = valid && logo[sb] && sx + 64 <= w && sy + 64 <= h; /* potential_logo_sb && is_complete_sb */
    const uint32_t g = on ? edge_cu_grad(y_mean + (size_t)sb * 85, cb_mean + (size_t)sb * 21, cr_mean + (size_t)sb * 21, b) : 0;
    if (PASS == 0) {
        uint32_t m = g;
        for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
        if ((threadIdx.x & 31) == 0 && m > 1) atomicMax(max_grad, m);
    } else if (valid) {
        const uint32_t v = min((g * 765u) / max(*max_grad, 1u), 255u); /* max_grad starts at 1 in the reference */
        edge_cu[i] = on && v >= 30;
    }
}
/* ---- determine_homogeneous_region_in_picture (:4660): one thread per super-block on the 64 8x8 variances the block statistics left ---- */
__global__ void homogeneous_regions_kernel(const uint16_t *variance, int w, int h, int sbs_w, int n_sb, uint32_t low_var_th, unsigned long long detail_th,
                                           unsigned long long *var_of_var32, uint8_t *homogeneous, unsigned long long *counts) {
    const int sb = blockIdx.x * blockDim.x + threadIdx.x;
    if (sb >= n_sb) return;
    const uint16_t *v = variance + (size_t)sb * 85;
    uint8_t hom = 1;
    if ((sb % sbs_w) * 64 + 64 > w || (sb / sbs_w) * 64 + 64 > h) { /* "should be re-calculated and scaled properly" (:4762): all ones */
#pragma unroll
        for (int k = 0; k < 4; k++) var_of_var32[(size_t)sb * 4 + k] = ~0ull;
    } else {
        unsigned long long s64 = 0, q64 = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            unsigned long long s = 0, q = 0;
            for (int r = 0; r < 4; r++)
                for (int c = 0; c < 4; c++) { const unsigned long long x = v[21 + ((k >> 1) * 4 + r) * 8 + (k & 1) * 4 + c]; s += x; q += x * x; }
            s64 += s; q64 += q;
            s >>= 4; q >>= 4;
            var_of_var32[(size_t)sb * 4 + k] = q - s * s;
        }
        s64 >>= 6; q64 >>= 6;
        hom = !(q64 - s64 * s64 > detail_th);
        atomicAdd(&counts[0], 1ull);
        if (v[0] < low_var_th) atomicAdd(&counts[1], 1ull);
        if (v[0] == 0) atomicAdd(&counts[2], 1ull);
    }
    homogeneous[sb] = hom;
}
#endif
