/* zz_kernel.cuh -- device code of zz.cu (compute_zz_sad, block statistics, luma histograms of the picture-analysis side); a header of its own so
 * that tests/emu/pa_emu.cpp can run the same source on the CPU. */
#ifndef SVTB200_ZZ_KERNEL_CUH
#define SVTB200_ZZ_KERNEL_CUH
#include <stdint.h>
#include <stddef.h>

struct ZzArgs {
    const uint8_t *cur16; int32_t c_stride, c_ox, c_oy;   /* current picture, 1/16 plane */
    const uint8_t *prev;  int32_t p_stride, p_ox, p_oy;   /* previous picture, full plane */
    int32_t w, h, sbs_w, n_sb, th_shift;
    uint8_t *scores; uint32_t *sads;
};

__global__ void __launch_bounds__(256) zz_sad_kernel(const ZzArgs A) {
    const int sb = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (sb >= A.n_sb) return;
    const int sx = (sb % A.sbs_w) * 64, sy = (sb / A.sbs_w) * 64;
    const bool complete = sx + 64 <= A.w && sy + 64 <= A.h;
    uint32_t sad = 0xffffffffu;
    if (complete) {
        /* lane: row y of the 16x16 block, 8 samples from x0 */
        const int y = lane >> 1, x0 = (lane & 1) * 8;
        const uint8_t *c = A.cur16 + (size_t)(A.c_oy + (sy >> 2) + y) * A.c_stride + A.c_ox + (sx >> 2) + x0;
        const uint8_t *p = A.prev + (size_t)(A.p_oy + sy + 4 * y) * A.p_stride + A.p_ox + sx + 4 * x0;
        uint32_t acc = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int a = __ldg(c + k), b = __ldg(p + 4 * k);
            acc += (uint32_t)(a > b ? a - b : b - a);
        }
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        sad = acc;
    }
    if (lane == 0) {
        const uint32_t t0 = 512u >> A.th_shift, t1 = 1024u >> A.th_shift, t2 = 2048u >> A.th_shift;
        A.scores[sb] = sad < t0 ? 0 : sad < t1 ? 10 : sad < t2 ? 20 : 30;
        if (A.sads) A.sads[sb] = sad;
    }
}

/* ---- block statistics: one warp per super-block, two 8x8 blocks per lane ---- */
struct StatArgs {
    const uint8_t *full; int32_t stride, ox, oy;
    int32_t sbs_w, n_sb;
    uint8_t *y_mean; uint16_t *variance;
};
__global__ void __launch_bounds__(256) block_stats_kernel(const StatArgs A) {
    __shared__ unsigned long long m_s[8][85], q_s[8][85]; /* per warp: means (8.8) and means of squares (16.16), ME_TIER_ZERO_PU order */
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, sb = blockIdx.x * 8 + w;
    if (sb >= A.n_sb) return;
    const int sx = (sb % A.sbs_w) * 64, sy = (sb / A.sbs_w) * 64;
    unsigned long long *m = m_s[w], *q = q_s[w];
    for (int h = 0; h < 2; h++) {
        const int b = lane + 32 * h, bx = (b & 7) * 8, by = (b >> 3) * 8;
        const uint8_t *p = A.full + (size_t)(A.oy + sy + by) * A.stride + A.ox + sx + bx;
        uint32_t sum = 0, sq = 0;
#pragma unroll
        for (int r = 0; r < 8; r += 2) { /* the reference sub-samples: rows 0, 2, 4, 6 */
#pragma unroll
            for (int k = 0; k < 8; k++) { const uint32_t v = __ldg(p + (size_t)r * A.stride + k); sum += v; sq += v * v; }
        }
        m[21 + b] = (unsigned long long)sum << 3;
        q[21 + b] = (unsigned long long)sq << 11;
    }
    __syncwarp();
    if (lane < 16) { /* 16x16: raster index lane, its four 8x8 */
        const int r = lane >> 2, c = lane & 3, a = 21 + (2 * r) * 8 + 2 * c;
        m[5 + lane] = (m[a] + m[a + 1] + m[a + 8] + m[a + 9]) >> 2;
        q[5 + lane] = (q[a] + q[a + 1] + q[a + 8] + q[a + 9]) >> 2;
    }
    __syncwarp();
    if (lane < 4) {
        const int r = lane >> 1, c = lane & 1, a = 5 + (2 * r) * 4 + 2 * c;
        m[1 + lane] = (m[a] + m[a + 1] + m[a + 4] + m[a + 5]) >> 2;
        q[1 + lane] = (q[a] + q[a + 1] + q[a + 4] + q[a + 5]) >> 2;
    }
    __syncwarp();
    if (lane == 0) { m[0] = (m[1] + m[2] + m[3] + m[4]) >> 2; q[0] = (q[1] + q[2] + q[3] + q[4]) >> 2; }
    __syncwarp();
    for (int i = lane; i < 85; i += 32) {
        A.y_mean[(size_t)sb * 85 + i] = (uint8_t)(m[i] >> 8);
        A.variance[(size_t)sb * 85 + i] = (uint16_t)((q[i] - m[i] * m[i]) >> 16);
    }
}

/* ---- luma histograms of the 1/16 plane: one CTA per (region, slice of its rows), shared-memory bins ---- */
struct HistArgs {
    const uint8_t *p16; int32_t stride, ox, oy, w, h;
    int32_t rw, rh, slices;
    uint32_t *count;            /* [rw * rh][256], zeroed */
    unsigned long long *sum;    /* [rw * rh], zeroed */
};
__device__ __forceinline__ void hist_region(const HistArgs &A, int ri, int *x0, int *y0, int *cw, int *ch) {
    const int iw = ri / A.rh, ih = ri - iw * A.rh, regw = A.w / A.rw, regh = A.h / A.rh;
    *x0 = iw * regw; *y0 = ih * regh;
    *cw = regw + (iw == A.rw - 1 ? A.w - A.rw * regw : 0);
    *ch = regh + (ih == A.rh - 1 ? A.h - A.rh * regh : 0);
}
__global__ void __launch_bounds__(256) hist_count_kernel(const HistArgs A) {
    __shared__ uint32_t bins[256];
    __shared__ unsigned long long ssum;
    const int ri = blockIdx.x / A.slices, sl = blockIdx.x - ri * A.slices;
    int x0, y0, cw, ch;
    hist_region(A, ri, &x0, &y0, &cw, &ch);
    bins[threadIdx.x] = 0;
    if (threadIdx.x == 0) ssum = 0;
    __syncthreads();
    const int r0 = (int)((long long)ch * sl / A.slices), r1 = (int)((long long)ch * (sl + 1) / A.slices);
    unsigned long long s = 0;
    for (int i = threadIdx.x; i < (r1 - r0) * cw; i += 256) {
        const int y = r0 + i / cw, x = i - (i / cw) * cw;
        const uint32_t v = __ldg(A.p16 + (size_t)(A.oy + y0 + y) * A.stride + A.ox + x0 + x);
        atomicAdd(&bins[v], 1u);
        s += v;
    }
    atomicAdd(&ssum, s);
    __syncthreads();
    if (bins[threadIdx.x]) atomicAdd(&A.count[(size_t)ri * 256 + threadIdx.x], bins[threadIdx.x]);
    if (threadIdx.x == 0 && ssum) atomicAdd(&A.sum[ri], ssum);
}
__global__ void hist_finish_kernel(const HistArgs A, uint32_t *hist, uint8_t *avg, unsigned long long *sum16) {
    const int ri = blockIdx.x;
    int x0, y0, cw, ch;
    hist_region(A, ri, &x0, &y0, &cw, &ch);
    hist[(size_t)ri * 256 + threadIdx.x] = (1u + A.count[(size_t)ri * 256 + threadIdx.x]) << 4;
    if (threadIdx.x == 0) {
        const unsigned long long area = (unsigned long long)cw * ch;
        avg[ri] = (uint8_t)((A.sum[ri] + (area >> 1)) / area);
        atomicAdd(sum16, A.sum[ri] << 4);
    }
}
#endif
