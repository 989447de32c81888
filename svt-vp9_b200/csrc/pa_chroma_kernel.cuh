/* pa_chroma_kernel.cuh -- device code of pa_chroma.cu (chroma block means / region histograms of picture analysis); a header of its own so that
 * tests/emu/pa_emu.cpp can run the same source on the CPU. */
#ifndef SVTB200_PA_CHROMA_KERNEL_CUH
#define SVTB200_PA_CHROMA_KERNEL_CUH
#include <stdint.h>
#include <stddef.h>

struct ChromaMeanArgs {
    const uint8_t *pl[2];
    int32_t stride, w, h, sbs_w, n_sb;
    uint8_t *mean[2]; /* n_sb x 21 each */
};
__global__ void __launch_bounds__(256) chroma_mean_kernel(const ChromaMeanArgs A) {
    const int sb = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (sb >= A.n_sb) return;
    const int sx = sb % A.sbs_w, sy = sb / A.sbs_w, p = lane >> 4, k = lane & 15;
    uint8_t *out = A.mean[p] + (size_t)sb * 21;
    if (sx * 64 + 64 > A.w || sy * 64 + 64 > A.h) { /* incomplete super-block: all zero */
        if (k == 0) for (int i = 0; i < 21; i++) out[i] = 0;
        return;
    }
    const uint8_t *s = A.pl[p] + (size_t)(sy * 32 + 8 * (k >> 2)) * A.stride + sx * 32 + 8 * (k & 3);
    uint32_t sum = 0;
#pragma unroll
    for (int y = 0; y < 8; y += 2) {
        const uint2 v = *(const uint2 *)(s + (size_t)y * A.stride); /* 8-byte aligned: stride and plane base are multiples of 8 */
        sum += __vsadu4(v.x, 0u) + __vsadu4(v.y, 0u);
    }
    const uint32_t m16 = sum << 3;
    /* 32x32 quadrant q = (k >> 3) * 2 + ((k >> 1) & 1): lanes k, k ^ 1, k ^ 4, k ^ 5 */
    uint32_t m32 = m16 + __shfl_xor_sync(0xffffffffu, m16, 1);
    m32 = (m32 + __shfl_xor_sync(0xffffffffu, m32, 4)) >> 2;
    const int base = lane & 16;
    const uint32_t q0 = __shfl_sync(0xffffffffu, m32, base + 0), q1 = __shfl_sync(0xffffffffu, m32, base + 2), q3 = __shfl_sync(0xffffffffu, m32, base + 10);
    out[5 + k] = (uint8_t)(m16 >> 8);
    if ((k & 5) == 0) out[1 + (k >> 3) * 2 + ((k >> 1) & 1)] = (uint8_t)(m32 >> 8);
    if (k == 0) out[0] = (uint8_t)(((q0 + q1 + q3 + q3) >> 2) >> 8);
}

struct ChromaHistArgs {
    const uint8_t *pl[2];
    int32_t stride, w, h, rw, rh, slices; /* w, h: luma size */
    uint32_t *count;           /* [rw * rh * 2][256], zeroed */
    unsigned long long *sum;   /* [rw * rh * 2], zeroed */
};
__device__ __forceinline__ void chroma_region(const ChromaHistArgs &A, int ri, int *x0, int *y0, int *w, int *h) {
    const int iw = ri / A.rh, ih = ri - iw * A.rh, regw = A.w / A.rw, regh = A.h / A.rh;
    *x0 = (iw * regw) >> 1; *y0 = (ih * regh) >> 1;
    *w = regw + (iw == A.rw - 1 ? A.w - A.rw * regw : 0);
    *h = regh + (ih == A.rh - 1 ? A.h - A.rh * regh : 0);
}
__global__ void __launch_bounds__(256) chroma_hist_count_kernel(const ChromaHistArgs A) {
    __shared__ uint32_t bins[256];
    __shared__ unsigned long long ssum;
    const int rp = blockIdx.x / A.slices, sl = blockIdx.x - rp * A.slices, ri = rp >> 1, p = rp & 1;
    int x0, y0, w, h;
    chroma_region(A, ri, &x0, &y0, &w, &h);
    bins[threadIdx.x] = 0;
    if (threadIdx.x == 0) ssum = 0;
    __syncthreads();
    const int nx = ((w >> 1) + 3) >> 2, ny = ((h >> 1) + 3) >> 2; /* samples x = 0, 4, .. < w / 2; rows likewise */
    const int r0 = (int)((long long)ny * sl / A.slices), r1 = (int)((long long)ny * (sl + 1) / A.slices);
    unsigned long long s = 0;
    for (int i = threadIdx.x; i < (r1 - r0) * nx; i += 256) {
        const int y = r0 + i / nx, x = i - (i / nx) * nx;
        const uint32_t v = __ldg(A.pl[p] + (size_t)(y0 + 4 * y) * A.stride + x0 + 4 * x);
        atomicAdd(&bins[v], 1u);
        s += v;
    }
    atomicAdd(&ssum, s);
    __syncthreads();
    if (bins[threadIdx.x]) atomicAdd(&A.count[(size_t)rp * 256 + threadIdx.x], bins[threadIdx.x]);
    if (threadIdx.x == 0 && ssum) atomicAdd(&A.sum[rp], ssum);
}
__global__ void chroma_hist_finish_kernel(const ChromaHistArgs A, uint32_t *hist, uint8_t *avg, unsigned long long *sum16) {
    const int rp = blockIdx.x, ri = rp >> 1, p = rp & 1;
    int x0, y0, w, h;
    chroma_region(A, ri, &x0, &y0, &w, &h);
    hist[(size_t)rp * 256 + threadIdx.x] = (1u + A.count[(size_t)rp * 256 + threadIdx.x]) << 4;
    if (threadIdx.x == 0) {
        const unsigned long long area = (unsigned long long)w * h, s = A.sum[rp] << 4;
        avg[rp] = (uint8_t)((s + (area >> 3)) / (area >> 2));
        atomicAdd(&sum16[p], s);
    }
}
#endif
