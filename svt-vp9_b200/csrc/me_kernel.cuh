/*
 * me_kernel.cuh -- motion estimation for one picture: one CTA (256 threads) per 64x64 super-block
 * runs the whole search chain of the reference's motion_estimate_sb()
 * (Source/Lib/Codec/EbMotionEstimation.c:4524): search-centre test, hierarchical ME on the 1/16,
 * 1/4 and full-resolution planes, the (0,0) check, the 85-PU full-pel search, half/quarter-pel
 * refinement on AVC-style interpolated planes, bi-prediction and result packing.
 *
 * B200 mapping (not a translation of the AVX2 code):
 *   - the source SB lives in shared memory for the whole chain; every search window is staged
 *     once into shared memory with 128-bit-friendly pitch and consumed from there;
 *   - block matching works on packed 4-byte words: funnel-shift re-alignment + VABSDIFF4.U8.ACC
 *     (4 absolute differences and the accumulate in one instruction);
 *   - "first minimum in raster order" is a single 64-bit atomicMin on (sad << 32 | raster index);
 *   - the 8x8 -> 16x16 -> 32x32 -> 64x64 SAD pyramid is built once per search position in shared
 *     memory, all 85 partitions reduce in parallel;
 *   - the three half-pel planes of both reference lists stay in shared memory (global scratch only
 *     for search areas too large to fit) so that quarter-pel and bi-prediction never re-filter.
 *
 * The file also compiles as plain C++ under csrc/emu/cuda_emu.h for the CPU-side logic tests.
 */
#ifndef SVTB200_ME_KERNEL_CUH
#define SVTB200_ME_KERNEL_CUH

#include <stdint.h>
#include "../../include/svtb200.h"

#define ME_THREADS 256
#define ME_SB 64
#define ME_MAX_SAD (64 * 64 * 255)
#define ME_CHUNK 32 /* search positions whose SAD pyramid is resident at once */

struct MePlaneDev {
    const uint8_t *buf;
    int32_t stride, ox, oy, w, h;
};

struct MeTraceDev { /* optional per-SB stage outputs, same fields as the oracle's trace */
    int16_t  hme_center_x[2], hme_center_y[2];
    int16_t  sa_ox[2], sa_oy[2], sa_w[2], sa_h[2];
    uint32_t fullpel_sad[2][85], fullpel_mv[2][85];
    uint8_t  subpel_enable[2];
};

struct MeKernelArgs {
    SvtB200MeParams    P;
    MePlaneDev         cur[3], ref[2][3];
    SvtB200MeCuResult *out;
    uint32_t          *rcme;
    MeTraceDev        *trace;
    uint8_t           *interp_scratch;   /* used when !interp_in_smem: n_sb * 2 * 3 * interp_plane_bytes */
    int32_t            interp_in_smem;
    int32_t            interp_plane_bytes; /* capacity of one half-pel plane */
    int32_t            win_bytes;          /* capacity of the search-window staging buffer */
    int32_t            sbs_w, n_sb;
    int32_t            sb_first;           /* SB index of block 0 (partial launches in the logic tests) */
};

/* z-order tables (raster PU index -> internal index), EbMotionEstimation.c:51-54 */
__device__ static const uint8_t me_tab16[16] = {0, 1, 4, 5, 2, 3, 6, 7, 8, 9, 12, 13, 10, 11, 14, 15};
__device__ static const uint8_t me_tab8[64]  = {0,  1,  4,  5,  16, 17, 20, 21, 2,  3,  6,  7,  18, 19, 22, 23,
                                                8,  9,  12, 13, 24, 25, 28, 29, 10, 11, 14, 15, 26, 27, 30, 31,
                                                32, 33, 36, 37, 48, 49, 52, 53, 34, 35, 38, 39, 50, 51, 54, 55,
                                                40, 41, 44, 45, 56, 57, 60, 61, 42, 43, 46, 47, 58, 59, 62, 63};
/* inverse of me_tab8: internal z index -> raster 8x8 index */
__device__ static const uint8_t me_itab8[64] = {0,  1,  8,  9,  2,  3,  10, 11, 16, 17, 24, 25, 18, 19, 26, 27,
                                                4,  5,  12, 13, 6,  7,  14, 15, 20, 21, 28, 29, 22, 23, 30, 31,
                                                32, 33, 40, 41, 34, 35, 42, 43, 48, 49, 56, 57, 50, 51, 58, 59,
                                                36, 37, 44, 45, 38, 39, 46, 47, 52, 53, 60, 61, 54, 55, 62, 63};

struct MeSbState { /* scalar control state of one SB, written by thread 0 */
    int32_t  ox, oy, sbw, sbh;
    int16_t  xsc, ysc;              /* running search centre of the current list */
    int16_t  sa_ox[2], sa_oy[2];    /* clipped search-area origin */
    int32_t  saw[2], sah[2];        /* clipped search-area size */
    int16_t  l0x[2][2], l0y[2][2], l1x[2][2], l1y[2][2], l2x[2][2], l2y[2][2];
    uint64_t l0s[2][2], l1s[2][2], l2s[2][2];
    int16_t  hme_x, hme_y;
    int32_t  srh;                   /* the reference's never-reset quadrant row counter */
    int32_t  e32, e16, e8, eq;      /* sub-pel enables of the current list */
    int32_t  istride[2], interp_ok[2];
    /* current windowed search */
    int32_t  s_w, s_h, s_pitch;     /* source block (bytes per row, rows, pitch) */
    int32_t  w_x, w_y;              /* window origin relative to the picture (0,0) of its plane */
    int32_t  w_saw, w_sah, w_pitch, w_rows;
    unsigned long long best_key;
    uint32_t red[9];
};

__device__ __forceinline__ const uint8_t *me_pix(const MePlaneDev &p, int x, int y) {
    return p.buf + (ptrdiff_t)(p.oy + y) * p.stride + p.ox + x;
}
__device__ __forceinline__ int16_t me_mvx(uint32_t mv) { return (int16_t)(mv & 0xffff); }
__device__ __forceinline__ int16_t me_mvy(uint32_t mv) { return (int16_t)(mv >> 16); }
__device__ __forceinline__ uint32_t me_mkmv(int x, int y) { return ((uint32_t)(uint16_t)y << 16) | (uint16_t)x; }
__device__ __forceinline__ uint32_t me_absdiff_acc(uint32_t a, uint32_t b, uint32_t acc) {
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(acc));
    return r;
#else
    return __vsadu4(a, b) + acc;
#endif
}

/* sum over the CTA; every thread gets the total */
__device__ __forceinline__ uint32_t me_block_sum(uint32_t v, uint32_t *red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5;
    __syncthreads(); /* protect red[] from the previous use */
    if ((threadIdx.x & 31) == 0) red[warp] = v;
    __syncthreads();
    uint32_t t = 0;
    for (int i = 0; i < ME_THREADS / 32; i++) t += red[i];
    return t;
}

/* window clipping exactly as the reference's "Correct the ... edge" statements (the left/top
 * branch never shrinks the size because the origin is moved first) */
__device__ __forceinline__ void me_clip_window(int origin, int16_t *org, int16_t *size, int pad, int pic_dim) {
    int16_t o = *org, s = *size;
    if (origin + o < -pad) o = (int16_t)(-pad - origin);
    if (origin + o > pic_dim - 1) o = (int16_t)(o - ((origin + o) - (pic_dim - 1)));
    if (origin + o + s > pic_dim) {
        int v = s - ((origin + o + s) - pic_dim);
        s = (int16_t)(v < 1 ? 1 : v);
    }
    *org = o;
    *size = s;
}
__device__ __forceinline__ int16_t me_clip_center(int origin, int16_t c, int pad, int pic_dim) {
    if (origin + c < -pad) c = (int16_t)(-pad - origin);
    if (origin + c > pic_dim - 1) c = (int16_t)(c - ((origin + c) - (pic_dim - 1)));
    return c;
}
__device__ __forceinline__ int16_t me_widen8(int16_t w) {
    return (int16_t)((w < 8) ? 8 : (w & 7) ? w + (w - ((w >> 3) << 3)) : w);
}

/* SAD of the source SB (every other row) against the reference at displacement (dx,dy):
 * sbw x sbh/2 samples, times two.  All threads must call; all get the result. */
__device__ static uint32_t me_sb_sad_at(const uint8_t *src, const MeSbState *st, const MePlaneDev &rp, int dx, int dy,
                                        uint32_t *red) {
    const uint8_t *r = me_pix(rp, st->ox + dx, st->oy + dy);
    const int n = (st->sbh >> 1) * st->sbw;
    uint32_t acc = 0;
    for (int i = threadIdx.x; i < n; i += ME_THREADS) {
        const int y = i / st->sbw, x = i - y * st->sbw;
        const int a = src[(2 * y) * ME_SB + x], b = __ldg(r + (ptrdiff_t)(2 * y) * rp.stride + x);
        acc += (uint32_t)(a > b ? a - b : b - a);
    }
    return me_block_sum(acc, red) << 1;
}

/* Stage a rows x cols byte window of plane pl starting at picture coordinates (x0,y0) into
 * shared memory with the given pitch (pitch >= cols + 4, multiple of 4). */
__device__ static void me_stage_window(uint8_t *win, int pitch, const MePlaneDev &pl, int x0, int y0, int cols, int rows) {
    const uint8_t *g = me_pix(pl, x0, y0);
    const int cw = cols + 4; /* one spare word for the funnel shift */
    for (int i = threadIdx.x; i < rows * cw; i += ME_THREADS) {
        const int y = i / cw, x = i - y * cw;
        win[y * pitch + x] = __ldg(g + (ptrdiff_t)y * pl.stride + x);
    }
}

/* SAD of one candidate row: wwords packed source words vs the window row at byte offset xoff */
__device__ __forceinline__ uint32_t me_row_sad(const uint32_t *srow, const uint8_t *wrow, int xoff, int wwords, uint32_t acc) {
    const uint32_t *wp = (const uint32_t *)(wrow + (xoff & ~3));
    const uint32_t  sh = (uint32_t)(xoff & 3) * 8;
    uint32_t lo = wp[0];
    for (int k = 0; k < wwords; k++) {
        const uint32_t hi = wp[k + 1];
        acc = me_absdiff_acc(__funnelshift_r(lo, hi, sh), srow[k], acc);
        lo = hi;
    }
    return acc;
}

/* Exhaustive window search (eb_vp9_sad_loop_kernel semantics): source block s_w x s_h in sblk
 * (pitch s_pitch, rows already sub-sampled), candidate (xs,ys) uses window rows ys + 2r.
 * Result: st->best_key = (sad << 32 | ys * saw + xs) of the first minimum, or ~0 if none. */
__device__ static void me_window_search(const uint8_t *sblk, const uint8_t *win, MeSbState *st, uint32_t *posacc /* >= 256 */) {
    const int saw = st->w_saw, sah = st->w_sah, npos = saw * sah, w = st->s_w, h = st->s_h;
    const int wwords = w >> 2;
    if (threadIdx.x == 0) st->best_key = ~0ull;
    __syncthreads();
    if (npos <= 0) return;
    if (npos >= ME_THREADS || (w & 3)) {
        unsigned long long best = ~0ull;
        for (int pos = threadIdx.x; pos < npos; pos += ME_THREADS) {
            const int ys = pos / saw, xs = pos - ys * saw;
            uint32_t acc = 0;
            if (w & 3) {
                for (int r = 0; r < h; r++)
                    for (int x = 0; x < w; x++) {
                        const int a = sblk[r * st->s_pitch + x], b = win[(ys + 2 * r) * st->w_pitch + xs + x];
                        acc += (uint32_t)(a > b ? a - b : b - a);
                    }
            } else {
                for (int r = 0; r < h; r++)
                    acc = me_row_sad((const uint32_t *)(sblk + r * st->s_pitch), win + (ys + 2 * r) * st->w_pitch, xs, wwords, acc);
            }
            if (acc < 0xffffffu) {
                const unsigned long long key = ((unsigned long long)acc << 32) | (uint32_t)pos;
                if (key < best) best = key;
            }
        }
        if (best != ~0ull) atomicMin(&st->best_key, best);
    } else {
        /* few positions: split every candidate's rows over P threads */
        int P = 1;
        while (P * 2 * npos <= ME_THREADS && P * 2 <= h) P *= 2;
        for (int i = threadIdx.x; i < npos; i += ME_THREADS) posacc[i] = 0;
        __syncthreads();
        const int task = threadIdx.x;
        if (task < npos * P) {
            const int pos = task / P, part = task - pos * P;
            const int ys = pos / saw, xs = pos - ys * saw;
            uint32_t acc = 0;
            for (int r = part; r < h; r += P)
                acc = me_row_sad((const uint32_t *)(sblk + r * st->s_pitch), win + (ys + 2 * r) * st->w_pitch, xs, wwords, acc);
            atomicAdd(&posacc[pos], acc);
        }
        __syncthreads();
        if (threadIdx.x < (unsigned)npos && posacc[threadIdx.x] < 0xffffffu)
            atomicMin(&st->best_key, ((unsigned long long)posacc[threadIdx.x] << 32) | threadIdx.x);
    }
    __syncthreads();
}

/* one HME level: thread 0 prepared st->s_*, st->w_*; all threads stage and search.
 * level 0/1 read the source block from the decimated planes of the current picture, level 2
 * from the shared-memory SB. */
__device__ static void me_hme_search(const MeKernelArgs &A, int list, int level, MeSbState *st, const uint8_t *src,
                                     uint8_t *hsrc, uint8_t *win, uint32_t *posacc) {
    const int w = st->s_w, h = st->s_h;
    /* source block: every other row of the (decimated) SB */
    if (level == 2) {
        for (int i = threadIdx.x; i < w * h; i += ME_THREADS) {
            const int y = i / w, x = i - y * w;
            hsrc[y * st->s_pitch + x] = src[(2 * y) * ME_SB + x];
        }
    } else {
        const MePlaneDev &cp = A.cur[level == 0 ? 2 : 1];
        const int sh = level == 0 ? 2 : 1;
        const uint8_t *g = me_pix(cp, st->ox >> sh, st->oy >> sh);
        for (int i = threadIdx.x; i < w * h; i += ME_THREADS) {
            const int y = i / w, x = i - y * w;
            hsrc[y * st->s_pitch + x] = __ldg(g + (ptrdiff_t)(2 * y) * cp.stride + x);
        }
    }
    const MePlaneDev &rp = A.ref[list][level == 0 ? 2 : level == 1 ? 1 : 0];
    me_stage_window(win, st->w_pitch, rp, st->w_x, st->w_y, st->w_saw + w - 1, st->w_rows);
    __syncthreads();
    me_window_search(hsrc, win, st, posacc);
}

/* prepare the window geometry for a search of saw x sah candidates (thread 0) */
__device__ __forceinline__ void me_set_window(MeSbState *st, int w, int h, int wx, int wy, int saw, int sah) {
    st->s_w = w; st->s_h = h; st->s_pitch = (w + 3) & ~3;
    st->w_x = wx; st->w_y = wy; st->w_saw = saw; st->w_sah = sah;
    st->w_pitch = ((saw + w - 1 + 4) + 3) & ~3;
    st->w_rows  = sah + 2 * (h - 1);
}


/* ---------------------------------------------------------------------------------------------
 * Full-pel search over the 85 partitions.
 * win holds the reference region starting at R(0,0) = top-left of the clipped search area
 * (cols saw+63(+16 for the tail quirk), rows sah+63).  Search positions are processed in chunks
 * of ME_CHUNK; per chunk the 64 sub-sampled 8x8 SADs of every position are written to shared
 * memory and summed up the pyramid; every partition keeps (2*sad << 32 | raster index) minima.
 * ------------------------------------------------------------------------------------------- */
__device__ static void me_fullpel(const uint8_t *src, const uint8_t *win, int wpitch, MeSbState *st, int list,
                                  unsigned long long *pu_key /* [85] smem */, uint16_t *s8 /* [CHUNK][64] */,
                                  uint16_t *s16 /* [CHUNK][16] */, uint32_t *s32 /* [CHUNK][4] */) {
    const int saw = st->saw[list], sah = st->sah[list], npos = saw * sah;
    const int tail_from = saw & ~7; /* columns handled by the reference's one-position routine */
    for (int i = threadIdx.x; i < 85; i += ME_THREADS) pu_key[i] = ((unsigned long long)ME_MAX_SAD << 32) | 0xffffffffu;
    __syncthreads();
    const int t = threadIdx.x;
    for (int c0 = 0; c0 < npos; c0 += ME_CHUNK) {
        const int cn = min(ME_CHUNK, npos - c0);
        /* phase 1: 8x8 SADs (rows 0,2,4,6), z-ordered so that 4 consecutive entries form a 16x16 */
        {
            const int z8 = t & 63, lane = t >> 6;
            const int z16 = z8 >> 2, k = z8 & 3;
            /* raster position of the 16x16 block with internal index z16 */
            const int b16 = (z16 & 1) | ((z16 & 4) >> 1) | ((z16 & 2) << 1) | (z16 & 8);
            const int bx = (b16 & 3) * 16 + (k & 1) * 8, by = (b16 >> 2) * 16 + (k >> 1) * 8;
            const uint32_t *sp = (const uint32_t *)(src + by * ME_SB + bx);
            for (int pp = lane; pp < cn; pp += 4) {
                const int pos = c0 + pp, ys = pos / saw, xs = pos - ys * saw;
                /* the reference's single-position routine matches 16x16 blocks 12 and 13 sixteen
                 * samples further right (EbMotionEstimation.c:855-856); reproduced */
                const int skew = (xs >= tail_from && (z16 == 12 || z16 == 13)) ? 16 : 0;
                const uint8_t *wr = win + (ys + by) * wpitch;
                const int xo = xs + bx + skew;
                uint32_t acc = 0;
#pragma unroll
                for (int r = 0; r < 4; r++) acc = me_row_sad(sp + r * (2 * ME_SB / 4), wr + (2 * r) * wpitch, xo, 2, acc);
                s8[pp * 64 + z8] = (uint16_t)acc;
            }
        }
        __syncthreads();
        /* phase 2: 8x8 minima and 16x16 sums */
        {
            const int z8 = t & 63, lane = t >> 6;
            unsigned long long best = ~0ull;
            for (int pp = lane; pp < cn; pp += 4) {
                const uint32_t v = 2u * s8[pp * 64 + z8];
                if (v < ME_MAX_SAD) { const unsigned long long key = ((unsigned long long)v << 32) | (uint32_t)(c0 + pp); if (key < best) best = key; }
            }
            if (best != ~0ull) atomicMin(&pu_key[21 + z8], best);
            const int z16 = t & 15, l16 = t >> 4;
            best = ~0ull;
            for (int pp = l16; pp < cn; pp += 16) {
                const uint16_t *q = s8 + pp * 64 + z16 * 4;
                const uint32_t sum = (uint32_t)q[0] + q[1] + q[2] + q[3];
                s16[pp * 16 + z16] = (uint16_t)sum;
                const uint32_t v = 2u * sum;
                if (v < ME_MAX_SAD) { const unsigned long long key = ((unsigned long long)v << 32) | (uint32_t)(c0 + pp); if (key < best) best = key; }
            }
            if (best != ~0ull) atomicMin(&pu_key[5 + z16], best);
        }
        __syncthreads();
        /* phase 3: 32x32 */
        {
            const int q32 = t & 3, l32 = t >> 2;
            unsigned long long best = ~0ull;
            for (int pp = l32; pp < cn; pp += 64) {
                const uint16_t *q = s16 + pp * 16 + q32 * 4;
                const uint32_t sum = (uint32_t)q[0] + q[1] + q[2] + q[3];
                s32[pp * 4 + q32] = sum;
                const uint32_t v = 2u * sum;
                if (v < ME_MAX_SAD) { const unsigned long long key = ((unsigned long long)v << 32) | (uint32_t)(c0 + pp); if (key < best) best = key; }
            }
            if (best != ~0ull) atomicMin(&pu_key[1 + q32], best);
        }
        __syncthreads();
        /* phase 4: 64x64 */
        if (t < cn) {
            const uint32_t *q = s32 + t * 4;
            const uint32_t v = 2u * (q[0] + q[1] + q[2] + q[3]);
            if (v < ME_MAX_SAD) atomicMin(&pu_key[0], ((unsigned long long)v << 32) | (uint32_t)(c0 + t));
        }
        __syncthreads();
    }
}

/* ---------------------------------------------------------------------------------------------
 * su_pel_enable (EbMotionEstimation.c:3839-4258), thread 0
 * ------------------------------------------------------------------------------------------- */
__device__ static void me_su_pel_enable(const uint32_t *bs, const uint32_t *bm, int tl, int *e32, int *e16, int *e8) {
    int sx = 0, sy = 0;
    uint32_t ss = 0;
    for (int i = 1; i < 5; i++) { sx += me_mvx(bm[i]); sy += me_mvy(bm[i]); ss += bs[i]; }
    uint32_t ax = (uint32_t)(sx >> 2), ay = (uint32_t)(sy >> 2);
    const uint32_t mag32 = ax * ax + ay * ay, sad32 = ss >> 2;
    ss = 0;
    for (int i = 5; i < 21; i++) ss += bs[i];
    const uint32_t sad16 = ss >> 4;
    sx = sy = 0; ss = 0;
    for (int i = 21; i < 85; i++) { sx += me_mvx(bm[i]); sy += me_mvy(bm[i]); ss += bs[i]; }
    ax = (uint32_t)(sx >> 6); ay = (uint32_t)(sy >> 6);
    const uint32_t mag8 = ax * ax + ay * ay, sad8 = ss >> 6;
    const uint32_t th = (tl == 1) ? 32u * 32u : (tl == 2) ? 80u * 80u : 48u * 48u;
    const int near32 = mag32 < th, low32 = sad32 < 32 * 32 * 6;
    if (tl == 0 || tl == 2) *e32 = low32;
    else if (tl == 1) *e32 = !(near32 && !low32);
    else *e32 = !(!near32 && !low32);
    *e16 = !(sad16 < 16 * 16 * 2);
    const int near8 = mag8 < th, low8 = sad8 < 8 * 8 * 2;
    *e8 = (tl <= 2) ? !low8 : (near8 && !low8);
}

/* ---------------------------------------------------------------------------------------------
 * Half-pel planes of the search region (interpolate_search_region_avc).  With R(x,y) the
 * reference sample at the top-left of the clipped search area + (x,y):
 *   pb[x + y*is] = H(R(x-2..x+1, y-2))       x < wasm, y < H+4
 *   ph[x + y*is] = V(R(x-1, y-2..y+1))       x < wasm, y < H+1
 *   pj[x + y*is] = V(pb[x + (y..y+3)*is])    x < wasm, y < H+1
 * taps {-2,18,18,-2}, +16 >> 5, clipped (two rounded 8-bit passes for j).
 * ------------------------------------------------------------------------------------------- */
__device__ __forceinline__ uint8_t me_tap4(int a, int b, int c, int d) {
    int v = (-2 * a + 18 * b + 18 * c - 2 * d + 16) >> 5;
    return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
}
__device__ static void me_interp(const MePlaneDev &rp, const MeSbState *st, int list, uint8_t *pb, uint8_t *ph, uint8_t *pj) {
    const int W = st->saw[list] + ME_SB - 1, H = st->sah[list] + ME_SB - 1;
    const int wasm = (W + 2 + 7) & ~7, is = st->istride[list];
    const uint8_t *R = me_pix(rp, st->ox + st->sa_ox[list], st->oy + st->sa_oy[list]);
    const int rs = rp.stride;
    for (int i = threadIdx.x; i < wasm * (H + 4); i += ME_THREADS) {
        const int y = i / wasm, x = i - y * wasm;
        const uint8_t *p = R + (ptrdiff_t)(y - 2) * rs + (x - 2);
        pb[x + y * is] = me_tap4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
    }
    for (int i = threadIdx.x; i < wasm * (H + 1); i += ME_THREADS) {
        const int y = i / wasm, x = i - y * wasm;
        const uint8_t *p = R + (ptrdiff_t)(y - 2) * rs + (x - 1);
        ph[x + y * is] = me_tap4(__ldg(p), __ldg(p + rs), __ldg(p + 2 * rs), __ldg(p + 3 * rs));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < wasm * (H + 1); i += ME_THREADS) {
        const int y = i / wasm, x = i - y * wasm;
        const uint8_t *q = pb + x + y * is;
        pj[x + y * is] = me_tap4(q[0], q[is], q[2 * is], q[3 * is]);
    }
    __syncthreads();
}

/* generic distortion between the source PU (shared memory, pitch 64) and up to two prediction
 * sources averaged on the fly; rows [r0, r0+nr) with the given row step; method as the reference */
struct MePred { const uint8_t *p1, *p2; int s1, s2; }; /* p2 == nullptr: single source */

__device__ __forceinline__ uint32_t me_pred_dist(const uint8_t *s, const MePred &pr, int w, int y, int sq) {
    const uint8_t *a = pr.p1 + (ptrdiff_t)y * pr.s1;
    const uint8_t *sr = s + y * ME_SB;
    uint32_t acc = 0;
    if (pr.p2) {
        const uint8_t *b = pr.p2 + (ptrdiff_t)y * pr.s2;
        for (int x = 0; x < w; x++) {
            const int d = (int)sr[x] - ((a[x] + b[x] + 1) >> 1);
            acc += sq ? (uint32_t)(d * d) : (uint32_t)(d < 0 ? -d : d);
        }
    } else {
        for (int x = 0; x < w; x++) {
            const int d = (int)sr[x] - (int)a[x];
            acc += sq ? (uint32_t)(d * d) : (uint32_t)(d < 0 ? -d : d);
        }
    }
    return acc;
}

/* Evaluate ntask (PU, candidate) tasks in parallel: task i covers PU geometry geo[i] and
 * prediction preds[i]; result dist[i] (metric per 'method'; for SSD also sad[i] = full SAD). */
struct MeTask { MePred pr; int16_t px, py, w, h; int32_t valid; };

__device__ static void me_run_tasks(const uint8_t *src, const MeTask *tasks, int ntask, int method, uint32_t *dist,
                                    uint32_t *sad /* may be null unless method==SSD */) {
    for (int i = threadIdx.x; i < ntask; i += ME_THREADS) { dist[i] = 0; if (sad) sad[i] = 0; }
    __syncthreads();
    /* rows of all tasks are dealt round-robin to the threads */
    int T = 1;
    while (T * 2 * ntask <= ME_THREADS && T < 32) T *= 2;
    for (int base = 0; base < ntask * T; base += ME_THREADS) {
        const int id = base + threadIdx.x;
        if (id < ntask * T) {
            const int ti = id / T, part = id - ti * T;
            const MeTask &tk = tasks[ti];
            if (tk.valid) {
                const uint8_t *s = src + tk.py * ME_SB + tk.px;
                uint32_t d = 0, a = 0;
                if (method == 0) { /* SUB_SAD: even rows, doubled */
                    for (int y = 2 * part; y < tk.h; y += 2 * T) d += me_pred_dist(s, tk.pr, tk.w, y, 0);
                    d <<= 1;
                } else if (method == 1) {
                    for (int y = part; y < tk.h; y += T) d += me_pred_dist(s, tk.pr, tk.w, y, 0);
                } else {
                    for (int y = part; y < tk.h; y += T) { d += me_pred_dist(s, tk.pr, tk.w, y, 1); a += me_pred_dist(s, tk.pr, tk.w, y, 0); }
                }
                atomicAdd(&dist[ti], d);
                if (sad && method == 2) atomicAdd(&sad[ti], a);
            }
        }
    }
    __syncthreads();
}

/* quarter-pel source table: kind 0 = integer plane, 1 = b, 2 = h, 3 = j; offsets in samples.
 * [phase][candidate L,R,T,B,TL,TR,BR,BL][2]  (set_quarter_pel_refinement_inputs_on_the_fly) */
struct MeQSrc { int8_t kind, ox, oy; };
__device__ static const MeQSrc me_qtab[4][8][2] = {
    {{{1, 0, 0}, {0, 0, 0}}, {{0, 0, 0}, {1, 1, 0}}, {{2, 0, 0}, {0, 0, 0}}, {{0, 0, 0}, {2, 0, 1}},
     {{1, 0, 0}, {2, 0, 0}}, {{2, 0, 0}, {1, 1, 0}}, {{2, 0, 1}, {1, 1, 0}}, {{1, 0, 0}, {2, 0, 1}}},
    {{{0, -1, 0}, {1, 0, 0}}, {{1, 0, 0}, {0, 0, 0}}, {{3, 0, 0}, {1, 0, 0}}, {{1, 0, 0}, {3, 0, 1}},
     {{2, -1, 0}, {1, 0, 0}}, {{1, 0, 0}, {2, 0, 0}}, {{1, 0, 0}, {2, 0, 1}}, {{2, -1, 1}, {1, 0, 0}}},
    {{{3, 0, 0}, {2, 0, 0}}, {{2, 0, 0}, {3, 1, 0}}, {{0, 0, -1}, {2, 0, 0}}, {{2, 0, 0}, {0, 0, 0}},
     {{1, 0, -1}, {2, 0, 0}}, {{2, 0, 0}, {1, 1, -1}}, {{2, 0, 0}, {1, 1, 0}}, {{1, 0, 0}, {2, 0, 0}}},
    {{{2, -1, 0}, {3, 0, 0}}, {{3, 0, 0}, {2, 0, 0}}, {{1, 0, -1}, {3, 0, 0}}, {{3, 0, 0}, {1, 0, 0}},
     {{2, -1, 0}, {1, 0, -1}}, {{1, 0, -1}, {2, 0, 0}}, {{1, 0, 0}, {2, 0, 0}}, {{2, -1, 0}, {1, 0, 0}}},
};

struct MeInterp { uint8_t *pb, *ph, *pj; }; /* plane bases as written by me_interp */

/* pointer conventions of the reference's half_pel_search_sb arguments */
__device__ __forceinline__ const uint8_t *me_PB(const MeInterp &ip, int is) { return ip.pb + 2 * is; }
__device__ __forceinline__ const uint8_t *me_PH(const MeInterp &ip) { return ip.ph + 1; }
__device__ __forceinline__ const uint8_t *me_PJ(const MeInterp &ip) { return ip.pj; }

__device__ __forceinline__ const uint8_t *me_qsrc_ptr(const MePlaneDev &rp, const MeSbState *st, int list, const MeInterp &ip,
                                                      MeQSrc q, int x, int y, int *stride) {
    const int is = st->istride[list];
    x += q.ox; y += q.oy;
    switch (q.kind) {
    case 0: *stride = rp.stride; return me_pix(rp, st->ox + st->sa_ox[list] + x, st->oy + st->sa_oy[list] + y);
    case 1: *stride = is; return me_PB(ip, is) + x + y * is;
    case 2: *stride = is; return me_PH(ip) + x + y * is;
    default: *stride = is; return me_PJ(ip) + x + y * is;
    }
}

/* PU geometry by internal index: 0 = 64x64, 1..4 = 32x32 (raster), 5..20 = 16x16 (z), 21..84 = 8x8 (z) */
__device__ __forceinline__ void me_pu_geom(int idx, int *px, int *py, int *w) {
    if (idx == 0) { *px = 0; *py = 0; *w = 64; }
    else if (idx < 5) { *px = ((idx - 1) & 1) * 32; *py = ((idx - 1) >> 1) * 32; *w = 32; }
    else if (idx < 21) {
        const int z = idx - 5, b = (z & 1) | ((z & 4) >> 1) | ((z & 2) << 1) | (z & 8);
        *px = (b & 3) * 16; *py = (b >> 2) * 16; *w = 16;
    } else { const int b = me_itab8[idx - 21]; *px = (b & 7) * 8; *py = (b >> 3) * 8; *w = 8; }
}

/* Half-pel then quarter-pel refinement of the PUs idx0..idx1-1 (one size class), all at once. */
__device__ static void me_subpel_class(const MeKernelArgs &A, const uint8_t *src, MeSbState *st, int list, const MeInterp &ip,
                                       int idx0, int idx1, int do_quarter, int quarter_w_override, uint32_t *bsad, uint32_t *bmv,
                                       uint32_t *bssd, uint8_t *dir, MeTask *tasks, uint32_t *dist, uint32_t *sad) {
    const MePlaneDev &rp = A.ref[list][0];
    const int method = A.P.fractional_search_method, is = st->istride[list];
    const int npu = idx1 - idx0;
    /* ---- half-pel: 8 candidates per PU (+ the full-pel SSD when the metric is SSD) ---- */
    for (int i = threadIdx.x; i < npu * 9; i += ME_THREADS) {
        const int pu = i / 9, k = i - pu * 9, idx = idx0 + pu;
        int px, py, w;
        me_pu_geom(idx, &px, &py, &w);
        const int16_t x_mv = me_mvx(bmv[idx]), y_mv = me_mvy(bmv[idx]);
        const int xs = (x_mv >> 2) - st->sa_ox[list], ys = (y_mv >> 2) - st->sa_oy[list];
        MeTask tk;
        tk.px = (int16_t)px; tk.py = (int16_t)py; tk.w = (int16_t)w; tk.h = (int16_t)w; tk.valid = 1;
        tk.pr.p2 = 0; tk.pr.s2 = 0;
        const int base = xs + px + (ys + py) * is;
        if (k == 8) { /* integer position (only needed for SSD) */
            tk.pr.p1 = me_pix(rp, st->ox + st->sa_ox[list] + xs + px, st->oy + st->sa_oy[list] + ys + py);
            tk.pr.s1 = rp.stride; tk.valid = (method == 2);
        } else {
            const uint8_t *pb = me_PB(ip, is), *ph = me_PH(ip), *pj = me_PJ(ip);
            const uint8_t *c;
            switch (k) { /* L, R, T, B, TL, TR, BR, BL */
            case 0: c = pb + base; break;
            case 1: c = pb + base + 1; break;
            case 2: c = ph + base; break;
            case 3: c = ph + base + is; break;
            case 4: c = pj + base; break;
            case 5: c = pj + base + 1; break;
            case 6: c = pj + base + 1 + is; break;
            default: c = pj + base + is; break;
            }
            tk.pr.p1 = c; tk.pr.s1 = is;
        }
        tasks[i] = tk;
    }
    __syncthreads();
    me_run_tasks(src, tasks, npu * 9, method, dist, sad);
    for (int pu = threadIdx.x; pu < npu; pu += ME_THREADS) {
        const int idx = idx0 + pu;
        const uint32_t *d = dist + pu * 9;
        const int16_t x_mv = me_mvx(bmv[idx]), y_mv = me_mvy(bmv[idx]);
        const int dx[8] = {-2, 2, 0, 0, -2, 2, 2, -2}, dy[8] = {0, 0, -2, 2, -2, -2, 2, 2};
        uint32_t best = (method == 2) ? d[8] : bsad[idx];
        int win = -1;
        for (int k = 0; k < 8; k++) if (d[k] < best) { best = d[k]; win = k; }
        if (method == 2) bssd[idx] = best;
        if (win >= 0) {
            bsad[idx] = (method == 2) ? sad[pu * 9 + win] : best;
            bmv[idx]  = me_mkmv(x_mv + dx[win], y_mv + dy[win]);
        }
        /* direction of the smallest half-pel distortion, tie order L,R,T,B,TL,TR,BL,BR;
         * codes TL0 T1 TR2 R3 BR4 B5 BL6 L7 */
        const int order[8] = {0, 1, 2, 3, 4, 5, 7, 6};
        const int code[8]  = {7, 3, 1, 5, 0, 2, 4, 6};
        uint32_t m = d[0];
        for (int k = 1; k < 8; k++) m = d[k] < m ? d[k] : m;
        for (int t = 0; t < 8; t++) if (d[order[t]] == m) { dir[idx] = (uint8_t)code[order[t]]; break; }
    }
    __syncthreads();
    if (!do_quarter) return;
    /* ---- quarter-pel: the three positions next to the winning half-pel direction ---- */
    for (int i = threadIdx.x; i < npu * 8; i += ME_THREADS) {
        const int pu = i >> 3, k = i & 7, idx = idx0 + pu;
        int px, py, w;
        me_pu_geom(idx, &px, &py, &w);
        if (quarter_w_override) w = quarter_w_override;
        const int16_t x_mv = me_mvx(bmv[idx]), y_mv = me_mvy(bmv[idx]);
        const int xs = ((x_mv + 2) >> 2) - st->sa_ox[list], ys = ((y_mv + 2) >> 2) - st->sa_oy[list];
        const int phase = (y_mv & 2) + ((x_mv & 2) >> 1);
        const int d = phase ? (dir[idx] + 4) & 7 : dir[idx];
        const int ring[8] = {7, 3, 1, 5, 0, 2, 4, 6};
        const int diff = (ring[k] - d) & 7;
        MeTask tk;
        tk.px = (int16_t)px; tk.py = (int16_t)py; tk.w = (int16_t)w; tk.h = (int16_t)w;
        tk.valid = (diff == 0 || diff == 1 || diff == 7);
        int s1 = 0, s2 = 0;
        tk.pr.p1 = me_qsrc_ptr(rp, st, list, ip, me_qtab[phase][k][0], xs + px, ys + py, &s1);
        tk.pr.p2 = me_qsrc_ptr(rp, st, list, ip, me_qtab[phase][k][1], xs + px, ys + py, &s2);
        tk.pr.s1 = s1; tk.pr.s2 = s2;
        tasks[i] = tk;
    }
    __syncthreads();
    me_run_tasks(src, tasks, npu * 8, method, dist, sad);
    for (int pu = threadIdx.x; pu < npu; pu += ME_THREADS) {
        const int idx = idx0 + pu;
        const int16_t x_mv = me_mvx(bmv[idx]), y_mv = me_mvy(bmv[idx]);
        const int dx[8] = {-1, 1, 0, 0, -1, 1, 1, -1}, dy[8] = {0, 0, -1, 1, -1, -1, 1, 1};
        uint32_t best = (method == 2) ? bssd[idx] : bsad[idx];
        int win = -1;
        for (int k = 0; k < 8; k++)
            if (tasks[pu * 8 + k].valid && dist[pu * 8 + k] < best) { best = dist[pu * 8 + k]; win = k; }
        if (win >= 0) {
            if (method == 2) { bssd[idx] = best; bsad[idx] = sad[pu * 8 + win]; }
            else bsad[idx] = best;
            bmv[idx] = me_mkmv(x_mv + dx[win], y_mv + dy[win]);
        }
    }
    __syncthreads();
}

/* prediction of list 'list' at quarter-pel mv for the PU at (px,py): at most two planes to average
 * (select_buffer / quarter_pel_compensation) */
__device__ static void me_bipred_pred(const MePlaneDev &rp, const MeSbState *st, int list, const MeInterp &ip, uint32_t mv,
                                      int px, int py, MePred *out) {
    const int is = st->istride[list];
    const int16_t x = me_mvx(mv), y = me_mvy(mv);
    const int xs = (x >> 2) - st->sa_ox[list], ys = (y >> 2) - st->sa_oy[list];
    const int frac = (x & 3) + ((y & 3) << 2);
    const uint8_t *Ap = me_pix(rp, st->ox + st->sa_ox[list] + xs + px, st->oy + st->sa_oy[list] + ys + py);
    const int as = rp.stride;
    /* b RIGHT of the integer sample, h BELOW it, j down-right */
    const uint8_t *b = ip.pb + (xs + 1 + px) + (ys + 2 + py) * is;
    const uint8_t *h = ip.ph + (xs + 1 + px) + (ys + 1 + py) * is;
    const uint8_t *j = ip.pj + (xs + 1 + px) + (ys + 1 + py) * is;
    MePred p; p.p1 = Ap; p.s1 = as; p.p2 = 0; p.s2 = is;
    switch (frac) {
    case 0: break;
    case 2: p.p1 = b; p.s1 = is; break;
    case 8: p.p1 = h; p.s1 = is; break;
    case 10: p.p1 = j; p.s1 = is; break;
    case 1: p.p2 = b; break;
    case 3: p.p1 = b; p.s1 = is; p.p2 = Ap + 1; p.s2 = as; break;
    case 4: p.p2 = h; break;
    case 5: p.p1 = b; p.s1 = is; p.p2 = h; break;
    case 6: p.p1 = b; p.s1 = is; p.p2 = j; break;
    case 7: p.p1 = b; p.s1 = is; p.p2 = h + 1; break;
    case 9: p.p1 = h; p.s1 = is; p.p2 = j; break;
    case 11: p.p1 = j; p.s1 = is; p.p2 = h + 1; break;
    case 12: p.p1 = h; p.s1 = is; p.p2 = Ap + as; p.s2 = as; break;
    case 13: p.p1 = h; p.s1 = is; p.p2 = b + is; break;
    case 14: p.p1 = j; p.s1 = is; p.p2 = b + is; break;
    default: p.p1 = h + 1; p.s1 = is; p.p2 = b + is; break;
    }
    *out = p;
}

__device__ __forceinline__ uint32_t me_pred_px(const MePred &p, int x, int y) {
    const uint32_t a = p.p1[(ptrdiff_t)y * p.s1 + x];
    if (!p.p2) return a;
    return (a + p.p2[(ptrdiff_t)y * p.s2 + x] + 1) >> 1;
}


/* ---------------------------------------------------------------------------------------------
 * shared-memory layout of one CTA
 * ------------------------------------------------------------------------------------------- */
struct MeSmemLayout {
    int src, hsrc, win, interp, s8, s16, s32, pu_key, bsad, bmv, bssd, bip, dir, posacc, tasks, dist, sad, preds, state, total;
};
#define ME_MAX_TASKS 576
__host__ __device__ static inline int me_align16(int v) { return (v + 15) & ~15; }
__host__ __device__ static inline MeSmemLayout me_smem_layout(int win_bytes, int interp_plane_bytes, int interp_in_smem) {
    MeSmemLayout L;
    int o = 0;
    L.src = o; o += ME_SB * ME_SB;
    L.hsrc = o; o += 64 * 32;
    L.win = o; o += me_align16(win_bytes);
    L.interp = o; o += interp_in_smem ? 6 * me_align16(interp_plane_bytes) : 0;
    L.s8 = o; o += ME_CHUNK * 64 * 2;
    L.s16 = o; o += ME_CHUNK * 16 * 2;
    L.s32 = o; o += ME_CHUNK * 4 * 4;
    L.pu_key = o; o += me_align16(85 * 8);
    L.bsad = o; o += me_align16(2 * 85 * 4);
    L.bmv = o; o += me_align16(2 * 85 * 4);
    L.bssd = o; o += me_align16(2 * 85 * 4);
    L.bip = o; o += me_align16(85 * 4);
    L.dir = o; o += me_align16(85);
    L.posacc = o; o += 256 * 4;
    L.tasks = o; o += me_align16(ME_MAX_TASKS * (int)sizeof(MeTask));
    L.dist = o; o += ME_MAX_TASKS * 4;
    L.sad = o; o += ME_MAX_TASKS * 4;
    L.preds = o; o += me_align16(85 * 2 * (int)sizeof(MePred));
    L.state = o; o += me_align16((int)sizeof(MeSbState));
    L.total = o;
    return L;
}

#ifdef SVTB200_CPU_EMU
#define ME_DYN_SMEM(name) uint8_t *name = emu_dyn_smem
#else
#define ME_DYN_SMEM(name) extern __shared__ __align__(16) uint8_t name[]
#endif

/* one HME level over all quadrants of one list (uniform control flow: every thread runs this) */
__device__ static void me_hme_level(const MeKernelArgs &A, int list, int level, int single, MeSbState *st, const uint8_t *src,
                                    uint8_t *hsrc, uint8_t *win, uint32_t *posacc) {
    const SvtB200MeParams &P = A.P;
    const int nqx = single ? 1 : P.number_hme_search_region_in_width, nqy = single ? 1 : P.number_hme_search_region_in_height;
    for (int qy = 0; qy < nqy; qy++)
        for (int qx = 0; qx < nqx; qx++) {
            __shared__ int16_t x_org_s, y_org_s;
            if (threadIdx.x == 0) {
                int16_t saw, sah, x_org, y_org;
                if (level == 0) {
                    const MePlaneDev &rp = A.ref[list][2];
                    const int ox = st->ox >> 2, oy = st->oy >> 2;
                    const int16_t cx = (int16_t)(st->xsc >> 2), cy = (int16_t)(st->ysc >> 2);
                    if (single) {
                        saw   = (int16_t)((P.hme_level0_total_search_area_width * P.hme_l0_multiplier_x) / 100);
                        sah   = (int16_t)((P.hme_level0_total_search_area_height * P.hme_l0_multiplier_y) / 100);
                        x_org = (int16_t)(-(int16_t)(saw >> 1) + cx);
                        y_org = (int16_t)(-(int16_t)(sah >> 1) + cy);
                    } else {
                        saw = (int16_t)((P.hme_level0_search_area_in_width[qx] * P.hme_l0_multiplier_x) / 100);
                        sah = (int16_t)((P.hme_level0_search_area_in_height[qy] * P.hme_l0_multiplier_y) / 100);
                        int16_t xd = cx, yd = cy;
                        for (int k = qx; k > 0; k--) xd = (int16_t)(xd + (int16_t)((P.hme_level0_search_area_in_width[k - 1] * P.hme_l0_multiplier_x) / 100));
                        for (int k = qy; k > 0; k--) yd = (int16_t)(yd + (int16_t)((P.hme_level0_search_area_in_height[k - 1] * P.hme_l0_multiplier_y) / 100));
                        x_org = (int16_t)(-(int16_t)(((P.hme_level0_total_search_area_width * P.hme_l0_multiplier_x) / 100) >> 1) + xd);
                        y_org = (int16_t)(-(int16_t)(((P.hme_level0_total_search_area_height * P.hme_l0_multiplier_y) / 100) >> 1) + yd);
                    }
                    me_clip_window(ox, &x_org, &saw, rp.ox - 1, rp.w);
                    me_clip_window(oy, &y_org, &sah, rp.oy - 1, rp.h);
                    if (single && (saw & 15)) saw = (int16_t)((saw >> 4) << 4);
                    me_set_window(st, st->sbw >> 2, (st->sbh >> 2) >> 1, ox + x_org, oy + y_org, saw, sah);
                } else if (level == 1) {
                    const MePlaneDev &rp = A.ref[list][1];
                    const int ox = st->ox >> 1, oy = st->oy >> 1;
                    saw = me_widen8((int16_t)P.hme_level1_search_area_in_width[qx]);
                    sah = (int16_t)P.hme_level1_search_area_in_height[qy];
                    x_org = (int16_t)(-(saw >> 1) + (int16_t)(st->l0x[qx][qy] >> 1));
                    y_org = (int16_t)(-(sah >> 1) + (int16_t)(st->l0y[qx][qy] >> 1));
                    me_clip_window(ox, &x_org, &saw, rp.ox - 1, rp.w);
                    me_clip_window(oy, &y_org, &sah, rp.oy - 1, rp.h);
                    me_set_window(st, st->sbw >> 1, (st->sbh >> 1) >> 1, ox + x_org, oy + y_org, saw, sah);
                } else {
                    const MePlaneDev &rp = A.ref[list][0];
                    saw = me_widen8((int16_t)P.hme_level2_search_area_in_width[qx]);
                    sah = (int16_t)P.hme_level2_search_area_in_height[qy];
                    x_org = (int16_t)(-(saw >> 1) + st->l1x[qx][qy]);
                    y_org = (int16_t)(-(sah >> 1) + st->l1y[qx][qy]);
                    me_clip_window(st->ox, &x_org, &saw, ME_SB - 1, rp.w);
                    me_clip_window(st->oy, &y_org, &sah, ME_SB - 1, rp.h);
                    me_set_window(st, st->sbw, st->sbh >> 1, st->ox + x_org, st->oy + y_org, saw, sah);
                }
                x_org_s = x_org; y_org_s = y_org;
            }
            __syncthreads();
            me_hme_search(A, list, level, st, src, hsrc, win, posacc);
            if (threadIdx.x == 0) {
                /* outputs keep their previous contents when no candidate beat 0xffffff */
                int16_t *px = level == 0 ? &st->l0x[qx][qy] : level == 1 ? &st->l1x[qx][qy] : &st->l2x[qx][qy];
                int16_t *py = level == 0 ? &st->l0y[qx][qy] : level == 1 ? &st->l1y[qx][qy] : &st->l2y[qx][qy];
                uint64_t *ps = level == 0 ? &st->l0s[qx][qy] : level == 1 ? &st->l1s[qx][qy] : &st->l2s[qx][qy];
                uint64_t best = 0xffffff;
                if (st->best_key != ~0ull) {
                    best = st->best_key >> 32;
                    const int pos = (int)(st->best_key & 0xffffffffu);
                    *px = (int16_t)(pos % st->w_saw);
                    *py = (int16_t)(pos / st->w_saw);
                }
                *ps = best * 2;
                const int mul = level == 0 ? 4 : level == 1 ? 2 : 1;
                *px = (int16_t)((*px + x_org_s) * mul);
                *py = (int16_t)((*py + y_org_s) * mul);
            }
            __syncthreads();
        }
}

/* ---------------------------------------------------------------------------------------------
 * the kernel: one CTA per super-block
 * ------------------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(ME_THREADS) me_sb_kernel(const __grid_constant__ MeKernelArgs A) {
    ME_DYN_SMEM(smem);
    const SvtB200MeParams &P = A.P;
    const MeSmemLayout L = me_smem_layout(A.win_bytes, A.interp_plane_bytes, A.interp_in_smem);
    uint8_t  *src = smem + L.src, *hsrc = smem + L.hsrc, *win = smem + L.win;
    uint16_t *s8 = (uint16_t *)(smem + L.s8), *s16 = (uint16_t *)(smem + L.s16);
    uint32_t *s32 = (uint32_t *)(smem + L.s32);
    unsigned long long *pu_key = (unsigned long long *)(smem + L.pu_key);
    uint32_t *bsad = (uint32_t *)(smem + L.bsad), *bmv = (uint32_t *)(smem + L.bmv), *bssd = (uint32_t *)(smem + L.bssd);
    uint32_t *bip = (uint32_t *)(smem + L.bip), *posacc = (uint32_t *)(smem + L.posacc);
    uint8_t  *dir = smem + L.dir;
    MeTask   *tasks = (MeTask *)(smem + L.tasks);
    uint32_t *dist = (uint32_t *)(smem + L.dist), *sadv = (uint32_t *)(smem + L.sad);
    MePred   *preds = (MePred *)(smem + L.preds);
    MeSbState *st = (MeSbState *)(smem + L.state);
    const int sb = blockIdx.x + A.sb_first, t = threadIdx.x, nl = P.num_lists;
    const int ipb = me_align16(A.interp_plane_bytes);
    MeInterp ip[2];
    for (int l = 0; l < 2; l++) {
        uint8_t *base = A.interp_in_smem ? smem + L.interp + (size_t)l * 3 * ipb
                                         : A.interp_scratch + ((size_t)sb * 2 + l) * 3 * ipb;
        ip[l].pb = base; ip[l].ph = base + ipb; ip[l].pj = base + 2 * ipb;
    }

    if (t == 0) {
        st->ox = (sb % A.sbs_w) * ME_SB; st->oy = (sb / A.sbs_w) * ME_SB;
        st->sbw = min(ME_SB, A.cur[0].w - st->ox); st->sbh = min(ME_SB, A.cur[0].h - st->oy);
        st->srh = 0; st->hme_x = 0; st->hme_y = 0;
        for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) {
            st->l0x[a][b] = st->l0y[a][b] = st->l1x[a][b] = st->l1y[a][b] = st->l2x[a][b] = st->l2y[a][b] = 0;
            st->l0s[a][b] = st->l1s[a][b] = st->l2s[a][b] = 0;
        }
        st->interp_ok[0] = st->interp_ok[1] = 0;
    }
    for (int i = t; i < 2 * 85; i += ME_THREADS) { bsad[i] = ME_MAX_SAD; bmv[i] = 0; bssd[i] = 0; }
    for (int i = t; i < 85; i += ME_THREADS) { bip[i] = 0; dir[i] = 0; }
    __syncthreads();
    { /* source SB: all 64 rows from the padded plane */
        const uint8_t *g = me_pix(A.cur[0], st->ox, st->oy);
        for (int i = t; i < ME_SB * ME_SB / 4; i += ME_THREADS) {
            const int y = i >> 4, x = (i & 15) * 4;
            const uint8_t *p = g + (ptrdiff_t)y * A.cur[0].stride + x;
            src[y * ME_SB + x] = __ldg(p); src[y * ME_SB + x + 1] = __ldg(p + 1);
            src[y * ME_SB + x + 2] = __ldg(p + 2); src[y * ME_SB + x + 3] = __ldg(p + 3);
        }
    }
    __syncthreads();

    for (int list = 0; list < nl; list++) {
        const MePlaneDev &rp = A.ref[list][0];
        uint32_t *ls = bsad + list * 85, *lm = bmv + list * 85, *lq = bssd + list * 85;
        if (t == 0) { st->xsc = 0; st->ysc = 0; }
        __syncthreads();
        if (P.temporal_layer_index > 0 || list == 0) {
            /* ---- test_search_area_bounds: O, (A == O), B, C, D and, for list 1, the mirrored list-0 MV */
            const int pad = ME_SB - 1, W0 = P.hme_level0_total_search_area_width, H0 = P.hme_level0_total_search_area_height;
            const uint32_t c_o = me_sb_sad_at(src, st, rp, 0, 0, st->red);
            const uint32_t c_b = me_sb_sad_at(src, st, rp, me_clip_center(st->ox, (int16_t)W0, pad, rp.w), me_clip_center(st->oy, 0, pad, rp.h), st->red);
            const uint32_t c_c = me_sb_sad_at(src, st, rp, me_clip_center(st->ox, 0, pad, rp.w), me_clip_center(st->oy, (int16_t)(0 - H0), pad, rp.h), st->red);
            const uint32_t c_d = me_sb_sad_at(src, st, rp, me_clip_center(st->ox, 0, pad, rp.w), me_clip_center(st->oy, (int16_t)H0, pad, rp.h), st->red);
            uint64_t c_dir = 0xFFFFFFFFFFFFFull;
            int16_t dirx = 0, diry = 0;
            if (list == 1) {
                dirx = (int16_t)(0 - (me_mvx(bmv[0]) >> 2)); diry = (int16_t)(0 - (me_mvy(bmv[0]) >> 2));
                c_dir = me_sb_sad_at(src, st, rp, me_clip_center(st->ox, dirx, pad, rp.w), me_clip_center(st->oy, diry, pad, rp.h), st->red);
            }
            if (t == 0) {
                uint64_t best = c_o;
                if (c_b < best) best = c_b;
                if (c_c < best) best = c_c;
                if (c_d < best) best = c_d;
                if (c_dir < best) best = c_dir;
                int16_t xsc, ysc; /* tie order O, A(==O), B, C, direct, D; unclipped offsets */
                if (best == c_o) { xsc = 0; ysc = 0; }
                else if (best == c_b) { xsc = (int16_t)W0; ysc = 0; }
                else if (best == c_c) { xsc = 0; ysc = (int16_t)(0 - H0); }
                else if (best == c_dir) { xsc = dirx; ysc = diry; }
                else { xsc = 0; ysc = (int16_t)H0; }
                st->xsc = xsc; st->ysc = ysc;
            }
            __syncthreads();
            if (P.enable_hme_flag && st->sbh == ME_SB) {
                const int nqx = P.number_hme_search_region_in_width, nqy = P.number_hme_search_region_in_height;
                const int single = P.single_hme_quadrant && !P.enable_hme_level1 && !P.enable_hme_level2;
                if (t == 0) {
                    if (st->srh < nqy) {
                        for (int qy = 0; qy < nqy; qy++) for (int qx = 0; qx < nqx; qx++) {
                            st->l0x[qx][qy] = (int16_t)(st->xsc >> 2); st->l0y[qx][qy] = (int16_t)(st->ysc >> 2);
                            st->l1x[qx][qy] = (int16_t)(st->xsc >> 1); st->l1y[qx][qy] = (int16_t)(st->ysc >> 1);
                            st->l2x[qx][qy] = st->xsc; st->l2y[qx][qy] = st->ysc;
                        }
                        st->srh = nqy;
                    }
                    if (P.enable_hme_level0 && single) st->srh = 0;
                }
                __syncthreads();
                if (P.enable_hme_level0) me_hme_level(A, list, 0, single, st, src, hsrc, win, posacc);
                if (P.enable_hme_level1) me_hme_level(A, list, 1, 0, st, src, hsrc, win, posacc);
                if (P.enable_hme_level2) me_hme_level(A, list, 2, 0, st, src, hsrc, win, posacc);
                if (t == 0) {
                    int16_t (*fx)[2] = 0, (*fy)[2] = 0;
                    uint64_t (*fs)[2] = 0;
                    if (P.enable_hme_level2) { fx = st->l2x; fy = st->l2y; fs = st->l2s; }
                    else if (P.enable_hme_level1) { fx = st->l1x; fy = st->l1y; fs = st->l1s; }
                    else if (P.enable_hme_level0) { fx = st->l0x; fy = st->l0y; fs = st->l0s; }
                    if (fs) {
                        st->hme_x = fx[0][0]; st->hme_y = fy[0][0];
                        uint64_t best = fs[0][0];
                        if (!(single && fs == st->l0s))
                            for (int qy = 0; qy < nqy; qy++)
                                for (int qx = (qy == 0) ? 1 : 0; qx < nqx; qx++)
                                    if (fs[qx][qy] < best) { best = fs[qx][qy]; st->hme_x = fx[qx][qy]; st->hme_y = fy[qx][qy]; }
                        if (P.enable_hme_level2 && P.ref_poc[0] == P.ref_poc[1] && list == 1 && nqx * nqy > 1) {
                            const int tq = nqx * nqy;
                            for (int a = 0; a < tq - 1; a++)
                                for (int b = a + 1; b < tq; b++)
                                    if (st->l2s[a / nqx][a % nqx] > st->l2s[b / nqx][b % nqx]) {
                                        const int16_t tx = st->l2x[a / nqx][a % nqx], ty = st->l2y[a / nqx][a % nqx];
                                        const uint64_t ts = st->l2s[a / nqx][a % nqx];
                                        st->l2x[a / nqx][a % nqx] = st->l2x[b / nqx][b % nqx]; st->l2y[a / nqx][a % nqx] = st->l2y[b / nqx][b % nqx];
                                        st->l2s[a / nqx][a % nqx] = st->l2s[b / nqx][b % nqx];
                                        st->l2x[b / nqx][b % nqx] = tx; st->l2y[b / nqx][b % nqx] = ty; st->l2s[b / nqx][b % nqx] = ts;
                                    }
                            st->hme_x = st->l2x[0][1]; st->hme_y = st->l2y[0][1];
                        }
                    }
                    st->xsc = st->hme_x; st->ysc = st->hme_y;
                }
                __syncthreads();
            }
        }
        if (A.trace && t == 0) { A.trace[sb].hme_center_x[list] = st->xsc; A.trace[sb].hme_center_y[list] = st->ysc; }
        /* ---- (0,0) versus the HME centre */
        if (st->xsc != 0 || st->ysc != 0) {
            const int16_t cx = me_clip_center(st->ox, st->xsc, ME_SB - 1, rp.w), cy = me_clip_center(st->oy, st->ysc, ME_SB - 1, rp.h);
            const uint32_t zero = me_sb_sad_at(src, st, rp, 0, 0, st->red);
            const uint32_t hme  = me_sb_sad_at(src, st, rp, cx, cy, st->red);
            __syncthreads();
            if (t == 0) { if (zero <= hme) { st->xsc = 0; st->ysc = 0; } else { st->xsc = cx; st->ysc = cy; } }
            __syncthreads();
        }
        if (t == 0) {
            int16_t saw = (int16_t)min(P.search_area_width, 127), sah = (int16_t)min(P.search_area_height, 127);
            int16_t xo = (int16_t)(st->xsc - (saw >> 1)), yo = (int16_t)(st->ysc - (sah >> 1));
            me_clip_window(st->ox, &xo, &saw, ME_SB - 1, P.luma_width);
            me_clip_window(st->oy, &yo, &sah, ME_SB - 1, P.luma_height);
            st->sa_ox[list] = xo; st->sa_oy[list] = yo; st->saw[list] = saw; st->sah[list] = sah;
            st->w_pitch = ((saw + 63 + 16 + 4) + 3) & ~3;
            if (A.trace) { A.trace[sb].sa_ox[list] = xo; A.trace[sb].sa_oy[list] = yo; A.trace[sb].sa_w[list] = saw; A.trace[sb].sa_h[list] = sah; }
        }
        __syncthreads();
        /* ---- full-pel search */
        me_stage_window(win, st->w_pitch, rp, st->ox + st->sa_ox[list], st->oy + st->sa_oy[list], st->saw[list] + 63 + 16, st->sah[list] + 63);
        __syncthreads();
        me_fullpel(src, win, st->w_pitch, st, list, pu_key, s8, s16, s32);
        if (t < 85) {
            const unsigned long long k = pu_key[t];
            const uint32_t pos = (uint32_t)(k & 0xffffffffu);
            if (pos != 0xffffffffu) {
                const int ys = pos / st->saw[list], xs = pos - ys * st->saw[list];
                ls[t] = (uint32_t)(k >> 32);
                lm[t] = me_mkmv((xs + st->sa_ox[list]) * 4, (ys + st->sa_oy[list]) * 4);
            } else { ls[t] = ME_MAX_SAD; }
            if (A.trace) { A.trace[sb].fullpel_sad[list][t] = ls[t]; A.trace[sb].fullpel_mv[list][t] = lm[t]; }
        }
        __syncthreads();
        /* ---- sub-pel */
        if (t == 0) {
            int e32 = 0, e16 = 0, e8 = 0, eq = 0;
            if (P.fractional_search_model == 0) e32 = e16 = e8 = eq = 1;
            else if (P.fractional_search_model == 1) { me_su_pel_enable(ls, lm, P.temporal_layer_index, &e32, &e16, &e8); eq = 1; }
            st->e32 = e32; st->e16 = e16; st->e8 = e8; st->eq = eq;
            const int W = st->saw[list] + ME_SB - 1;
            st->istride[list] = (W + 2 + 7) & ~7;
            if (A.trace) A.trace[sb].subpel_enable[list] = (uint8_t)(e32 | (e16 << 1) | (e8 << 2) | (eq << 3));
        }
        __syncthreads();
        if (st->e32 || st->e16 || st->e8 || st->eq) {
            me_interp(rp, st, list, ip[list].pb, ip[list].ph, ip[list].pj);
            if (t == 0) st->interp_ok[list] = 1;
            const int e16 = st->e16 && P.cu16x16_mode == 0, e8 = st->e8 && P.cu8x8_mode != 1, eq = st->eq;
            if (P.fractional_search64x64) me_subpel_class(A, src, st, list, ip[list], 0, 1, 1, 32, ls, lm, lq, dir, tasks, dist, sadv);
            if (st->e32) me_subpel_class(A, src, st, list, ip[list], 1, 5, eq, 0, ls, lm, lq, dir, tasks, dist, sadv);
            if (e16) me_subpel_class(A, src, st, list, ip[list], 5, 21, eq, 0, ls, lm, lq, dir, tasks, dist, sadv);
            if (e8) me_subpel_class(A, src, st, list, ip[list], 21, 85, eq, 0, ls, lm, lq, dir, tasks, dist, sadv);
        }
        __syncthreads();
    }

    /* ---- bi-prediction: SAD(src, avg(pred L0, pred L1)) at the two uni-pred winners */
    if (nl == 2) {
        const int sub = (P.fractional_search_method == 0);
        for (int i = t; i < 85 * 2; i += ME_THREADS) {
            const int idx = i >> 1, l = i & 1;
            int px, py, w;
            me_pu_geom(idx, &px, &py, &w);
            me_bipred_pred(A.ref[l][0], st, l, ip[l], bmv[l * 85 + idx], px, py, &preds[i]);
        }
        __syncthreads();
        for (int cls = 0; cls < 4; cls++) {
            if (cls == 3 && P.cu8x8_mode != 0) continue;
            if (cls >= 2 && P.cu16x16_mode != 0) continue;
            for (int seg = t; seg < 512; seg += ME_THREADS) {
                const int y = seg >> 3, x8 = (seg & 7) * 8;
                if (sub && (y & 1)) continue;
                int idx;
                if (cls == 0) idx = 0;
                else if (cls == 1) idx = 1 + (x8 >> 5) + 2 * (y >> 5);
                else if (cls == 2) idx = 5 + me_tab16[(x8 >> 4) + 4 * (y >> 4)];
                else idx = 21 + me_tab8[(x8 >> 3) + 8 * (y >> 3)];
                int px, py, w;
                me_pu_geom(idx, &px, &py, &w);
                const MePred &p0 = preds[idx * 2], &p1 = preds[idx * 2 + 1];
                uint32_t acc = 0;
                for (int x = 0; x < 8; x++) {
                    const int d = (int)src[y * ME_SB + x8 + x] - (int)((me_pred_px(p0, x8 + x - px, y - py) + me_pred_px(p1, x8 + x - px, y - py) + 1) >> 1);
                    acc += (uint32_t)(d < 0 ? -d : d);
                }
                atomicAdd(&bip[idx], sub ? acc << 1 : acc);
            }
        }
        __syncthreads();
    }

    /* ---- pack (output PU order is raster inside each size class) */
    if (t < 85) {
        const int pu = t;
        const int n_idx = pu > 20 ? me_tab8[pu - 21] + 21 : pu > 4 ? me_tab16[pu - 5] + 5 : pu;
        int total = nl;
        if (nl == 2 && (P.cu8x8_mode == 0 || pu < 21) && (P.cu16x16_mode == 0 || pu < 5)) total = 3;
        SvtB200MeCuResult r;
        r.x_mv_l0 = me_mvx(bmv[n_idx]); r.y_mv_l0 = me_mvy(bmv[n_idx]);
        r.x_mv_l1 = nl == 2 ? me_mvx(bmv[85 + n_idx]) : 0; r.y_mv_l1 = nl == 2 ? me_mvy(bmv[85 + n_idx]) : 0;
        for (int i = 0; i < 3; i++) { r.cand[i].distortion = 0; r.cand[i].direction = 0; }
        for (int i = 0; i < 7; i++) r.pad_[i] = 0;
        r.total_me_candidate_index = (uint8_t)total;
        const uint32_t a = bsad[n_idx], b = bsad[85 + n_idx], c = bip[n_idx];
        if (total == 3) { /* sort3_elements: exact <= pattern */
            uint32_t d[3], k[3];
            if (a <= b && a <= c) { d[0] = a; k[0] = 0; if (b <= c) { d[1] = b; k[1] = 1; d[2] = c; k[2] = 2; } else { d[1] = c; k[1] = 2; d[2] = b; k[2] = 1; } }
            else if (b <= a && b <= c) { d[0] = b; k[0] = 1; if (a <= c) { d[1] = a; k[1] = 0; d[2] = c; k[2] = 2; } else { d[1] = c; k[1] = 2; d[2] = a; k[2] = 0; } }
            else if (a <= b) { d[0] = c; k[0] = 2; d[1] = a; k[1] = 0; d[2] = b; k[2] = 1; }
            else { d[0] = c; k[0] = 2; d[1] = b; k[1] = 1; d[2] = a; k[2] = 0; }
            for (int i = 0; i < 3; i++) { r.cand[i].distortion = d[i]; r.cand[i].direction = k[i]; }
        } else if (total == 2) {
            if (a <= b) { r.cand[0].distortion = a; r.cand[0].direction = 0; r.cand[1].distortion = b; r.cand[1].direction = 1; }
            else { r.cand[0].distortion = b; r.cand[0].direction = 1; r.cand[1].distortion = a; r.cand[1].direction = 0; }
        } else { r.cand[0].distortion = a; r.cand[0].direction = 0; }
        A.out[(size_t)sb * 85 + pu] = r;
        if (pu >= 5 && pu < 21) dist[pu] = r.cand[0].distortion;
    }
    __syncthreads();
    if (A.rcme && t == 0) {
        uint32_t s = 0;
        for (int i = 5; i < 21; i++) s += dist[i];
        A.rcme[sb] = s;
    }
}

#endif /* SVTB200_ME_KERNEL_CUH */
