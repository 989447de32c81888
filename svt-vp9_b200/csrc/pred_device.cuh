/* pred_device.cuh -- the inter-prediction kernel of pred.cu (device code only; tests/emu/pred_emu.cpp runs the same source on
 * the CPU under csrc/emu/cuda_emu.h). */
#ifndef SVTB200_PRED_DEVICE_CUH
#define SVTB200_PRED_DEVICE_CUH
#include <stdint.h>
#include "../../include/svtb200.h"

#ifdef SVTB200_CPU_EMU
#define PRED_SHARED16 alignas(16) static
#define PRED_CONST8 alignas(8) static const
#else
#define PRED_SHARED16 __shared__ __align__(16)
#define PRED_CONST8 __constant__ __align__(8)
#endif

#define PRED_THREADS 128
#define PRED_WP 80 /* window pitch: sample x of the block sits at byte x + 4, so that 4-sample groups are word aligned */
#define PRED_ROWS 73 /* 64 + 7 staged rows, + slack the 2-row groups of the vertical pass may touch */

/* the normative VP9 regular 8-tap filter (sub_pel_filters_8, Source/Lib/VPX/vp9_filter.c:32); phase 0 (a copy) never filters */
PRED_CONST8 int8_t c_pred_taps[16][8] = {
    {0, 0, 0, 127, 0, 0, 0, 0},
    {0, 1, -5, 126, 8, -3, 1, 0},      {-1, 3, -10, 122, 18, -6, 2, 0},   {-1, 4, -13, 118, 27, -9, 3, -1},
    {-1, 4, -16, 112, 37, -11, 4, -1}, {-1, 5, -18, 105, 48, -14, 4, -1}, {-1, 5, -19, 97, 58, -16, 5, -1},
    {-1, 6, -19, 88, 68, -18, 5, -1},  {-1, 6, -19, 78, 78, -19, 6, -1},  {-1, 5, -18, 68, 88, -19, 6, -1},
    {-1, 5, -16, 58, 97, -19, 5, -1},  {-1, 4, -14, 48, 105, -18, 5, -1}, {-1, 4, -11, 37, 112, -16, 4, -1},
    {-1, 3, -9, 27, 118, -13, 4, -1},  {0, 2, -6, 18, 122, -10, 3, -1},   {0, 1, -3, 8, 126, -5, 1, 0}};

__device__ __forceinline__ int pred_dp4a_us(uint32_t a /* 4 unsigned bytes */, uint32_t b /* 4 signed bytes */, int c) {
#ifdef SVTB200_CPU_EMU
    for (int i = 0; i < 4; i++) c += (int)((a >> (8 * i)) & 255) * (int)(int8_t)((b >> (8 * i)) & 255);
    return c;
#else
    int r;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
#endif
}
/* i / d for i < 5760, d <= 71 with rcp = ceil(2^32 / d) (exact over that range) */
__device__ __forceinline__ int pred_div(int i, uint32_t rcp) { return rcp ? (int)__umulhi((uint32_t)i, rcp) : i; }
__device__ __forceinline__ uint32_t pred_rcp(int d) { return d == 1 ? 0u : 0xffffffffu / (uint32_t)d + 1u; } /* 0: d == 1 */
__device__ __forceinline__ uint32_t pred_clip8(int v) { return (uint32_t)(v < 0 ? 0 : v > 255 ? 255 : v); }

/* one reference of the block -> res (pitch 64, shared).  Horizontal pass: four adjacent outputs per thread from three words
 * of the row, two 4-tap dot products each.  Vertical pass: a 4 x 2 tile per thread from nine words (4 columns each); even
 * and odd columns ride in the two 16-bit halves of a register (the sums are biased by 16384 so that a half never borrows from
 * its neighbour: 255 x the taps' negative / positive sums stays inside -10200 .. 42840). */
__device__ static void pred_one(const uint8_t *src /* at the block's integer position */, int stride, int sx, int sy, int w, int h,
                                uint8_t *win, uint8_t *tmp, uint8_t *res) {
    const int t = threadIdx.x;
    /* rows are fetched as aligned 32-bit words and re-aligned with a funnel shift when the geometry allows it (block width a
     * multiple of 4, word-aligned plane): a quarter of the load instructions of a byte loop.  The word path starts a filtered row at
     * sample -4 instead of -3 and may read up to 3 bytes past the last sample the reference functions read -- inside the plane's
     * padding (the motion vector clamp keeps the block within VP9_INTERP_EXTEND + block size of the picture, the padding is larger) */
    const bool words = (w & 3) == 0;
    if (!sx && !sy) { /* eb_vp9_convolve_copy_c */
        if (words) {
            const int wq = w >> 2;
            const uint32_t rwq = pred_rcp(wq);
            for (int i = t; i < wq * h; i += PRED_THREADS) {
                const int r = pred_div(i, rwq), c = i - r * wq;
                const uintptr_t a = (uintptr_t)(src + (ptrdiff_t)r * stride + 4 * c);
                const uint32_t *g = (const uint32_t *)(a & ~(uintptr_t)3);
                const uint32_t sh = (uint32_t)(a & 3) * 8, lo = __ldg(g), hi = sh ? __ldg(g + 1) : 0u;
                *(uint32_t *)(res + r * 64 + 4 * c) = __funnelshift_r(lo, hi, sh);
            }
        } else {
            const uint32_t rw = pred_rcp(w);
            for (int i = t; i < w * h; i += PRED_THREADS) {
                const int r = pred_div(i, rw), c = i - r * w;
                res[r * 64 + c] = __ldg(src + (ptrdiff_t)r * stride + c);
            }
        }
        __syncthreads();
        return;
    }
    const int x_lo = sx ? -3 : 0, nx = sx ? w + 7 : w;
    const int y_lo = sy ? -3 : 0, ny = sy ? h + 7 : h;
    if (words) {
        const int xa = sx ? -4 : 0, nwr = sx ? (w >> 2) + 2 : (w >> 2); /* samples xa .. xa + 4 * nwr - 1 of every staged row */
        const uint32_t rnw = pred_rcp(nwr);
        for (int i = t; i < nwr * ny; i += PRED_THREADS) {
            const int r = pred_div(i, rnw), c = i - r * nwr;
            const uintptr_t a = (uintptr_t)(src + (ptrdiff_t)(r + y_lo) * stride + xa + 4 * c);
            const uint32_t *g = (const uint32_t *)(a & ~(uintptr_t)3);
            const uint32_t sh = (uint32_t)(a & 3) * 8, lo = __ldg(g), hi = sh ? __ldg(g + 1) : 0u;
            *(uint32_t *)(win + r * PRED_WP + xa + 4 + 4 * c) = __funnelshift_r(lo, hi, sh);
        }
    } else {
        const uint32_t rnx = pred_rcp(nx);
        for (int i = t; i < nx * ny; i += PRED_THREADS) { /* only the samples the reference functions read */
            const int r = pred_div(i, rnx), c = i - r * nx;
            win[r * PRED_WP + c + x_lo + 4] = __ldg(src + (ptrdiff_t)(r + y_lo) * stride + c + x_lo);
        }
    }
    __syncthreads();
    const int wq = (w + 3) >> 2;
    const uint32_t rwq = pred_rcp(wq);
    const uint8_t *v_in = win + 4;
    int v_pitch = PRED_WP;
    if (sx) { /* over every staged row: h, or h + 7 when a vertical pass follows (vpx_convolve.c:153-162) */
        const uint32_t K0 = ((const uint32_t *)c_pred_taps[sx])[0], K1 = ((const uint32_t *)c_pred_taps[sx])[1];
        uint8_t *out = sy ? tmp : res;
        for (int u = t; u < wq * ny; u += PRED_THREADS) {
            const int r = pred_div(u, rwq), c4 = (u - r * wq) * 4;
            const uint32_t *p = (const uint32_t *)(win + r * PRED_WP + c4); /* samples c4 - 4 .. c4 + 7 */
            const uint32_t W0 = p[0], W1 = p[1], W2 = p[2];
            const uint32_t v0 = pred_clip8(pred_dp4a_us(__funnelshift_r(W0, W1, 8), K0, pred_dp4a_us(__funnelshift_r(W1, W2, 8), K1, 64)) >> 7);
            const uint32_t v1 = pred_clip8(pred_dp4a_us(__funnelshift_r(W0, W1, 16), K0, pred_dp4a_us(__funnelshift_r(W1, W2, 16), K1, 64)) >> 7);
            const uint32_t v2 = pred_clip8(pred_dp4a_us(__funnelshift_r(W0, W1, 24), K0, pred_dp4a_us(__funnelshift_r(W1, W2, 24), K1, 64)) >> 7);
            const uint32_t v3 = pred_clip8(pred_dp4a_us(W1, K0, pred_dp4a_us(W2, K1, 64)) >> 7);
            *(uint32_t *)(out + r * 64 + c4) = v0 | v1 << 8 | v2 << 16 | v3 << 24;
        }
        __syncthreads();
        if (!sy) return;
        v_in = tmp; v_pitch = 64;
    }
    {
        int tap[8];
#pragma unroll
        for (int k = 0; k < 8; k++) tap[k] = c_pred_taps[sy][k];
        const int hq = (h + 1) >> 1;
        for (int u = t; u < wq * hq; u += PRED_THREADS) {
            const int rq = pred_div(u, rwq), c4 = (u - rq * wq) * 4, r0 = rq * 2;
            uint32_t E[9], O[9];
#pragma unroll
            for (int k = 0; k < 9; k++) {
                const uint32_t W = *(const uint32_t *)(v_in + (r0 + k) * v_pitch + c4);
                E[k] = W & 0x00ff00ffu; O[k] = (W >> 8) & 0x00ff00ffu;
            }
#pragma unroll
            for (int i = 0; i < 2; i++) {
                uint32_t aE = 0x40404040u, aO = 0x40404040u; /* per half: 16384 + the rounding 64 */
#pragma unroll
                for (int k = 0; k < 8; k++) { aE += E[i + k] * (uint32_t)tap[k]; aO += O[i + k] * (uint32_t)tap[k]; }
                /* per half: ((sum + 64) >> 7) + 128, clamped to 128 .. 383, minus 128 */
                uint32_t qE = (aE >> 7) & 0x01ff01ffu, qO = (aO >> 7) & 0x01ff01ffu;
                qE = __vminu2(__vmaxu2(qE, 0x00800080u), 0x017f017fu) - 0x00800080u;
                qO = __vminu2(__vmaxu2(qO, 0x00800080u), 0x017f017fu) - 0x00800080u;
                *(uint32_t *)(res + (r0 + i) * 64 + c4) = qE | qO << 8;
            }
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(PRED_THREADS) pred_kernel(const SvtB200PredBlock *blocks, int n) {
    PRED_SHARED16 uint8_t win[PRED_ROWS * PRED_WP];
    PRED_SHARED16 uint8_t tmp[PRED_ROWS * 64];
    PRED_SHARED16 uint8_t res[2][68 * 64];
    for (int b = blockIdx.x; b < n; b += gridDim.x) {
        const SvtB200PredBlock B = blocks[b];
        pred_one(B.ref[0] + (ptrdiff_t)B.y[0] * B.ref_stride[0] + B.x[0], B.ref_stride[0], B.subpel_x[0], B.subpel_y[0], B.w, B.h, win, tmp, res[0]);
        if (B.compound)
            pred_one(B.ref[1] + (ptrdiff_t)B.y[1] * B.ref_stride[1] + B.x[1], B.ref_stride[1], B.subpel_x[1], B.subpel_y[1], B.w, B.h, win, tmp, res[1]);
        if (((uintptr_t)B.dst & 3) == 0 && (B.dst_stride & 3) == 0 && (B.w & 3) == 0) {
            const int wq = B.w >> 2;
            const uint32_t rwq = pred_rcp(wq);
            for (int i = threadIdx.x; i < wq * B.h; i += PRED_THREADS) {
                const int r = pred_div(i, rwq), c4 = (i - r * wq) * 4;
                uint32_t v = *(const uint32_t *)(res[0] + r * 64 + c4);
                if (B.compound) v = __vavgu4(v, *(const uint32_t *)(res[1] + r * 64 + c4)); /* (a + b + 1) >> 1: eb_vp9_convolve_avg_c, vpx_convolve.c:211 */
                *(uint32_t *)(B.dst + (ptrdiff_t)r * B.dst_stride + c4) = v;
            }
        } else {
            for (int i = threadIdx.x; i < B.w * B.h; i += PRED_THREADS) {
                const int r = i / B.w, c = i - r * B.w;
                int v = res[0][r * 64 + c];
                if (B.compound) v = (v + res[1][r * 64 + c] + 1) >> 1;
                B.dst[(ptrdiff_t)r * B.dst_stride + c] = (uint8_t)v;
            }
        }
        __syncthreads();
    }
}
/* ---- full-loop distortion: one warp per transform block ---- */
__global__ void __launch_bounds__(256) full_dist_kernel(int nn, int n, const int16_t *coeff, const int16_t *rc, const uint16_t *eob, uint64_t *dist) {
    const int k = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (k >= n) return;
    const int16_t *a = coeff + (size_t)k * nn, *b = rc + (size_t)k * nn;
    uint32_t res = 0, pred = 0;
    for (int i = lane; i < nn; i += 32) {
        const int c = a[i], d = (int16_t)(c - b[i]); /* sq_r16to32 takes an int16_t (EbPictureOperators_C.c:264) */
        res += (uint32_t)(d * d);
        pred += (uint32_t)(c * c);
    }
    for (int o = 16; o > 0; o >>= 1) { res += __shfl_xor_sync(0xffffffffu, res, o); pred += __shfl_xor_sync(0xffffffffu, pred, o); }
    if (lane == 0) { dist[2 * k] = eob[k] ? res : pred; dist[2 * k + 1] = pred; }
}
#endif
