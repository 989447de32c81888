/*
 * me_kernel.cuh -- motion estimation for one picture: one CTA (256 threads) per 64x64 super-block
 * runs the whole search chain of the reference's motion_estimate_sb()
 * (Source/Lib/Codec/EbMotionEstimation.c:4524): search-centre test, hierarchical ME on the 1/16,
 * 1/4 and full-resolution planes, the (0,0) check, the 85-PU full-pel search, half/quarter-pel
 * refinement on AVC-style interpolated planes, bi-prediction and result packing.
 *
 * B200 mapping (not a translation of the AVX2 code):
 *   - the source SB lives in shared memory for the whole chain; every search window is staged
 *     once into shared memory, re-aligned to 32-bit words on the way in, and consumed from there
 *     (block matching, the three half-pel filters and the integer samples of the quarter-pel
 *     averages all read the same staged copy);
 *   - everything that touches samples works on packed 4-byte words: funnel-shift re-alignment,
 *     VABSDIFF4.U8.ACC (4 absolute differences + accumulate), byte averages, DP4A for the
 *     horizontal 4-tap filter and 2x16-bit lane arithmetic for the vertical one;
 *   - exhaustive searches evaluate 4 horizontally adjacent candidates per thread from one set of
 *     window words; "first minimum in raster order" is a single 64-bit atomicMin on
 *     (sad << 32 | raster index);
 *   - the 8x8 -> 16x16 -> 32x32 -> 64x64 SAD pyramid is built once per search position in shared
 *     memory, all 85 partitions reduce in parallel;
 *   - sub-pel refinement of all enabled partition sizes runs as ONE pass over equally sized work
 *     units (16 words each) for the half-pel ring and one for the quarter-pel ring;
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
#define ME_SRC_PITCH 72   /* shared-memory pitch of the source SB and of the stashed list-0 predictions: rows two apart (the sub-sampled
                            metrics, one row per lane) are 36 words apart, i.e. eight lanes hit eight different banks (at 64 they all hit one) */
#define ME_MAX_SAD (64 * 64 * 255)
#define ME_HALO_X 4       /* staged full-pel window: columns left of the search-area origin (word aligned) */
#define ME_HALO_Y 2       /* rows above / below (the 4-tap filters reach 2 rows up, 1+1 down) */
#define ME_HME_STRIPE 32  /* HME candidate rows staged at once (bounds the window for wide HME areas) */

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

/* a CUtensorMap (cuda.h) without the header: 128 opaque bytes, 64-byte aligned, passed inside the __grid_constant__ kernel argument */
struct alignas(64) MeTensorMap { uint64_t opaque[16]; };
#define ME_TMA_SRC 1  /* source super-block: 64x64 box of the current picture's full plane */
/* The search WINDOWS cannot come through TMA (nor through cp.async): a tiled copy must start at a 16-byte aligned byte offset of
 * the row -- measured on B200 (tools/tma_probe2.cu: box start columns 16 and 32 land, 1, 4, 8 and 13 raise "illegal instruction",
 * the same when the row is described as 32-bit elements) -- while a window starts at any byte (search centre + clipping), and
 * everything downstream reads the staged rows as aligned 32-bit words.  Their re-alignment therefore stays in registers
 * (aligned word loads + funnel shift, me_stage_window).  The source super-block always starts at a multiple of 64 samples of a
 * plane whose origin is a multiple of 16 (svtb200_picture_upload_luma's geometry): it is the TMA case. */

struct MeKernelArgs {
    MeTensorMap        tm_cur;             /* valid when tma & ME_TMA_SRC */
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
    int32_t            n_work;             /* super-blocks of this launch (n_sb for a whole picture) */
    int32_t            sb_first;           /* SB index of block 0 (partial launches in the logic tests) */
    int32_t            tma;                /* ME_TMA_* bits: staged by cp.async.bulk.tensor + mbarrier instead of by the threads */
    uint32_t          *sb_counter;         /* persistent grid: next super-block to take (null: block b takes super-block b) */
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
    int32_t  next_sb;               /* persistent grid: the super-block this CTA took from the counter */
    int32_t  srh;                   /* the reference's never-reset quadrant row counter */
    int32_t  e32, e16, e8, eq;      /* sub-pel enables of the current list */
    int32_t  istride[2], interp_ok[2];
    /* current windowed search */
    int32_t  s_w, s_h, s_pitch;     /* source block (bytes per row, rows, pitch) */
    int32_t  w_x, w_y;              /* window origin relative to the picture (0,0) of its plane */
    int32_t  w_saw, w_sah, w_pitch;
    int32_t  fp_pitch;              /* pitch of the staged full-pel window */
    unsigned long long best_key;
    uint32_t red[8 * 5];
};

__device__ __forceinline__ const uint8_t *me_pix(const MePlaneDev &p, int x, int y) {
    return p.buf + (ptrdiff_t)(p.oy + y) * p.stride + p.ox + x;
}
__device__ __forceinline__ int16_t me_mvx(uint32_t mv) { return (int16_t)(mv & 0xffff); }
__device__ __forceinline__ int16_t me_mvy(uint32_t mv) { return (int16_t)(mv >> 16); }
__device__ __forceinline__ uint32_t me_mkmv(int x, int y) { return ((uint32_t)(uint16_t)y << 16) | (uint16_t)x; }

/* ---------------------------------------------------------------------------------------------
 * TMA staging: one elected thread arms the CTA's mbarrier with the box size and issues cp.async.bulk.tensor; every thread then
 * waits on the barrier's phase.  (Box start = a multiple of 16 bytes, see ME_TMA_SRC.)
 * ------------------------------------------------------------------------------------------- */
struct MeTmaState { uint32_t mbar, phase; }; /* shared-memory address of the mbarrier, parity of the next completion */
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ uint32_t me_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void me_mbar_init(uint32_t mbar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
/* thread 0 only.  Every thread of the CTA has passed a barrier since its last access to dst. */
__device__ __forceinline__ void me_tma_issue(const MeTensorMap *tm, uint8_t *dst, int x, int y, uint32_t bytes, const MeTmaState &ts) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); /* generic-proxy accesses to dst before the async-proxy write */
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ts.mbar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(me_smem_u32(dst)), "l"(tm), "r"(ts.mbar), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void me_tma_wait(MeTmaState &ts) {
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; spins++) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(ts.mbar), "r"(ts.phase) : "memory");
        if (spins > (1u << 24)) __trap(); /* a copy that never lands (bad descriptor) must end the launch with an error, not hang the device */
    }
    ts.phase ^= 1;
}
#else
__device__ __forceinline__ void me_mbar_init(uint32_t) {}
__device__ __forceinline__ void me_tma_issue(const MeTensorMap *, uint8_t *, int, int, uint32_t, const MeTmaState &) {}
__device__ __forceinline__ void me_tma_wait(MeTmaState &) {}
#endif

/* ---------------------------------------------------------------------------------------------
 * packed-byte primitives
 * ------------------------------------------------------------------------------------------- */
__device__ __forceinline__ uint32_t me_absdiff_acc(uint32_t a, uint32_t b, uint32_t acc) {
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(acc));
    return r;
#else
    return __vsadu4(a, b) + acc;
#endif
}
/* sum_i a.byte[i] (unsigned) * b.byte[i] (signed) + c */
__device__ __forceinline__ int me_dp4a_us(uint32_t a, uint32_t b, int c) {
#if defined(__CUDA_ARCH__)
    int r;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
#else
    for (int i = 0; i < 4; i++) c += (int)((a >> (8 * i)) & 0xff) * (int)(int8_t)((b >> (8 * i)) & 0xff);
    return c;
#endif
}
/* sum_i a.byte[i] * b.byte[i] (both unsigned) + c */
__device__ __forceinline__ uint32_t me_dp4a_uu(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
#else
    for (int i = 0; i < 4; i++) c += ((a >> (8 * i)) & 0xff) * ((b >> (8 * i)) & 0xff);
    return c;
#endif
}
/* per-lane max / min of two unsigned 16-bit lanes */
__device__ __forceinline__ uint32_t me_max_u16x2(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __vimax3_u16x2(a, b, b);
#else
    const uint32_t al = a & 0xffff, bl = b & 0xffff, ah = a >> 16, bh = b >> 16;
    return (al > bl ? al : bl) | ((ah > bh ? ah : bh) << 16);
#endif
}
__device__ __forceinline__ uint32_t me_min_u16x2(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __vimin3_u16x2(a, b, b);
#else
    const uint32_t al = a & 0xffff, bl = b & 0xffff, ah = a >> 16, bh = b >> 16;
    return (al < bl ? al : bl) | ((ah < bh ? ah : bh) << 16);
#endif
}
/* 4 bytes at any byte address (shared or global): the two enclosing aligned words, re-aligned.
 * Reads up to 4 bytes past p + 3; every buffer this is used on carries that much slack. */
__device__ __forceinline__ uint32_t me_ld4(const uint8_t *p) {
    const uintptr_t a = (uintptr_t)p;
    const uint32_t *w = (const uint32_t *)(a & ~(uintptr_t)3);
    return __funnelshift_r(w[0], w[1], (uint32_t)(a & 3) * 8);
}
/* a row of consecutive words starting at any byte address, one new aligned load per word */
struct MeStream {
    const uint32_t *w;
    uint32_t        lo, sh;
    __device__ __forceinline__ void open(const uint8_t *p) {
        const uintptr_t a = (uintptr_t)p;
        w = (const uint32_t *)(a & ~(uintptr_t)3); sh = (uint32_t)(a & 3) * 8; lo = *w++;
    }
    __device__ __forceinline__ uint32_t next() {
        const uint32_t hi = *w++, v = __funnelshift_r(lo, hi, sh);
        lo = hi;
        return v;
    }
};

/* sum over the CTA of n (<= 5) values per thread; every thread gets the totals */
template <int N>
__device__ __forceinline__ void me_block_sum_n(uint32_t (&v)[N], int n, uint32_t *red) {
    const int warp = threadIdx.x >> 5;
    __syncthreads(); /* protect red[] from the previous use */
#pragma unroll
    for (int c = 0; c < N; c++) {
        if (c < n) {
            uint32_t x = v[c];
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) == 0) red[c * 8 + warp] = x;
        }
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < N; c++) {
        if (c < n) {
            uint32_t s = 0;
            for (int i = 0; i < ME_THREADS / 32; i++) s += red[c * 8 + i];
            v[c] = s;
        }
    }
}

/* minimum of 64-bit keys over lanes whose index differs in the bits of 'mask' (xor butterfly) */
__device__ __forceinline__ unsigned long long me_shfl_min64(unsigned long long v, int mask) {
    for (int o = 1; o < 32; o <<= 1)
        if (mask & o) {
            const unsigned long long x = __shfl_xor_sync(0xffffffffu, v, o);
            v = x < v ? x : v;
        }
    return v;
}
/* minimum over the CTA; red64: 8 words of shared memory; every thread gets the result */
__device__ __forceinline__ unsigned long long me_block_min64(unsigned long long v, unsigned long long *red64) {
    v = me_shfl_min64(v, 31);
    __syncthreads(); /* protect red64[] from the previous use */
    if ((threadIdx.x & 31) == 0) red64[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned long long m = red64[0];
    for (int i = 1; i < ME_THREADS / 32; i++) m = red64[i] < m ? red64[i] : m;
    return m;
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

/* SADs of the source SB (every other row) against the reference at n (<= 5) displacements:
 * sbw x sbh/2 samples, times two.  All threads must call; all get the results in out[]. */
__device__ static void me_sb_sad_multi(const uint8_t *src, const MeSbState *st, const MePlaneDev &rp, int n, const int16_t *dx,
                                       const int16_t *dy, uint32_t (&out)[5], uint32_t *red) {
    const int sbw = st->sbw, rows = st->sbh >> 1;
    uint32_t acc[5] = {0, 0, 0, 0, 0};
    if (((sbw | rp.stride) & 3) == 0 && ((uintptr_t)rp.buf & 3) == 0 && (sbw & 7) == 0) {
        /* one 8-sample segment per thread */
        const int nseg = sbw >> 3, seg = threadIdx.x;
        const int y = (nseg == 8) ? seg >> 3 : seg / nseg, x = (seg - y * nseg) * 8;
        if (y < rows) {
            const uint32_t *sp = (const uint32_t *)(src + (2 * y) * ME_SRC_PITCH + x);
            const uint32_t s0 = sp[0], s1 = sp[1];
#pragma unroll
            for (int c = 0; c < 5; c++) {
                if (c < n) {
                    const uint8_t *r = me_pix(rp, st->ox + dx[c] + x, st->oy + dy[c] + 2 * y);
                    const uintptr_t a = (uintptr_t)r;
                    const uint32_t *w = (const uint32_t *)(a & ~(uintptr_t)3);
                    const uint32_t sh = (uint32_t)(a & 3) * 8, w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
                    acc[c] = me_absdiff_acc(__funnelshift_r(w0, w1, sh), s0, acc[c]);
                    acc[c] = me_absdiff_acc(__funnelshift_r(w1, w2, sh), s1, acc[c]);
                }
            }
        }
    } else {
        const int cnt = rows * sbw;
        for (int c = 0; c < n; c++) {
            const uint8_t *r = me_pix(rp, st->ox + dx[c], st->oy + dy[c]);
            for (int i = threadIdx.x; i < cnt; i += ME_THREADS) {
                const int y = i / sbw, x = i - y * sbw;
                const int a = src[(2 * y) * ME_SRC_PITCH + x], b = __ldg(r + (ptrdiff_t)(2 * y) * rp.stride + x);
                acc[c] += (uint32_t)(a > b ? a - b : b - a);
            }
        }
    }
    me_block_sum_n(acc, n, red);
    for (int c = 0; c < 5; c++) out[c] = c < n ? acc[c] << 1 : 0;
}

/* Stage a rows x cols byte window of plane pl starting at picture coordinates (x0,y0) into shared
 * memory: byte x of row y of the staged copy is R(x0 + x, y0 + y), rows are word aligned (the
 * re-alignment happens here, once).  pitch is a multiple of 4 and >= cols + 8. */
__device__ static void me_stage_window(uint8_t *win, int pitch, const MePlaneDev &pl, int x0, int y0, int cols, int rows) {
    const uint8_t *g = me_pix(pl, x0, y0);
    if ((pl.stride & 3) == 0 && ((uintptr_t)pl.buf & 3) == 0) {
        const uint32_t sh = (uint32_t)((uintptr_t)g & 3) * 8;
        const uint8_t *ga = g - (sh >> 3);
        const int nw = ((cols + 3) >> 2) + 1, total = nw * rows;
        const uint32_t rcp = 0xffffffffu / (uint32_t)nw + 1u; /* ceil(2^32 / nw): __umulhi(i, rcp) == i / nw while i * nw < 2^32 */
        for (int i = threadIdx.x; i < total; i += ME_THREADS) {
            const int y = (int)__umulhi((uint32_t)i, rcp), x = i - y * nw;
            const uint32_t *gr = (const uint32_t *)(ga + (ptrdiff_t)y * pl.stride) + x;
            const uint32_t lo = __ldg(gr), hi = sh ? __ldg(gr + 1) : 0u;
            ((uint32_t *)(win + y * pitch))[x] = __funnelshift_r(lo, hi, sh);
        }
    } else {
        const int cw = cols + 4;
        for (int i = threadIdx.x; i < rows * cw; i += ME_THREADS) {
            const int y = i / cw, x = i - y * cw;
            win[y * pitch + x] = __ldg(g + (ptrdiff_t)y * pl.stride + x);
        }
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

/* Exhaustive search of one stripe of a window (eb_vp9_sad_loop_kernel semantics): source block
 * s_w x s_h in sblk (pitch s_pitch, rows already sub-sampled); the stripe holds candidate rows
 * [y0, y0 + nrows) of the w_saw-wide area, candidate (xs,ys) uses staged rows (ys - y0) + 2r.
 * Folds (sad << 32 | ys * saw + xs) of the first minimum into st->best_key. */
template <int WWORDS>
__device__ __forceinline__ void me_search4(const uint32_t *sblk, int sp32, const uint32_t *wrow0, int wp32, int h, uint32_t (&acc)[4]) {
    for (int r = 0; r < h; r++) {
        const uint32_t *wrow = wrow0 + (2 * r) * wp32, *srow = sblk + r * sp32;
        uint32_t lo = wrow[0];
#pragma unroll
        for (int k = 0; k < WWORDS; k++) {
            const uint32_t hi = wrow[k + 1], s = srow[k];
            acc[0] = me_absdiff_acc(lo, s, acc[0]);
            acc[1] = me_absdiff_acc(__funnelshift_r(lo, hi, 8), s, acc[1]);
            acc[2] = me_absdiff_acc(__funnelshift_r(lo, hi, 16), s, acc[2]);
            acc[3] = me_absdiff_acc(__funnelshift_r(lo, hi, 24), s, acc[3]);
            lo = hi;
        }
    }
}
__device__ static void me_window_search(const uint8_t *sblk, const uint8_t *win, MeSbState *st, uint32_t *posacc /* >= 256 */, int y0,
                                        int nrows) {
    const int saw = st->w_saw, npos = saw * nrows, w = st->s_w, h = st->s_h;
    const int wwords = w >> 2;
    if (npos <= 0) return;
    unsigned long long best = ~0ull;
    if (w & 3) {
        for (int pos = threadIdx.x; pos < npos; pos += ME_THREADS) {
            const int ys = pos / saw, xs = pos - ys * saw;
            uint32_t acc = 0;
            for (int r = 0; r < h; r++)
                for (int x = 0; x < w; x++) {
                    const int a = sblk[r * st->s_pitch + x], b = win[(ys + 2 * r) * st->w_pitch + xs + x];
                    acc += (uint32_t)(a > b ? a - b : b - a);
                }
            if (acc < 0xffffffu) {
                const unsigned long long key = ((unsigned long long)acc << 32) | (uint32_t)(pos + y0 * saw);
                if (key < best) best = key;
            }
        }
    } else if (npos >= ME_THREADS) {
        /* 4 horizontally adjacent candidates per thread share every window word */
        const int ngx = (saw + 3) >> 2, ngroups = ngx * nrows, sp32 = st->s_pitch >> 2, wp32 = st->w_pitch >> 2;
        const uint32_t rcp = 0xffffffffu / (uint32_t)ngx + 1u; /* __umulhi(g, rcp) == g / ngx */
        uint32_t bv = 0xffffffu, bp = 0;
        for (int g = threadIdx.x; g < ngroups; g += ME_THREADS) {
            const int ys = (int)__umulhi((uint32_t)g, rcp), gx = g - ys * ngx;
            const uint32_t *wrow0 = (const uint32_t *)win + ys * wp32 + gx;
            uint32_t acc[4] = {0, 0, 0, 0};
            if (wwords == 4) me_search4<4>((const uint32_t *)sblk, sp32, wrow0, wp32, h, acc);
            else if (wwords == 8) me_search4<8>((const uint32_t *)sblk, sp32, wrow0, wp32, h, acc);
            else if (wwords == 16) me_search4<16>((const uint32_t *)sblk, sp32, wrow0, wp32, h, acc);
            else {
                for (int r = 0; r < h; r++) {
                    const uint32_t *wrow = wrow0 + (2 * r) * wp32, *srow = (const uint32_t *)sblk + r * sp32;
                    uint32_t lo = wrow[0];
                    for (int k = 0; k < wwords; k++) {
                        const uint32_t hi = wrow[k + 1], s = srow[k];
                        acc[0] = me_absdiff_acc(lo, s, acc[0]);
                        acc[1] = me_absdiff_acc(__funnelshift_r(lo, hi, 8), s, acc[1]);
                        acc[2] = me_absdiff_acc(__funnelshift_r(lo, hi, 16), s, acc[2]);
                        acc[3] = me_absdiff_acc(__funnelshift_r(lo, hi, 24), s, acc[3]);
                        lo = hi;
                    }
                }
            }
            const uint32_t pos0 = (uint32_t)((y0 + ys) * saw + 4 * gx);
#pragma unroll
            for (int j = 0; j < 4; j++) /* raster order is increasing along a thread's groups: strict < keeps the first minimum */
                if (4 * gx + j < saw && acc[j] < bv) { bv = acc[j]; bp = pos0 + j; }
        }
        if (bv < 0xffffffu) best = ((unsigned long long)bv << 32) | bp;
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
            best = ((unsigned long long)posacc[threadIdx.x] << 32) | (uint32_t)(threadIdx.x + y0 * saw);
    }
    /* first minimum in raster order over the CTA (the key carries the raster index); posacc doubles as scratch */
    const unsigned long long m = me_block_min64(best, (unsigned long long *)posacc + 64);
    if (threadIdx.x == 0 && m < st->best_key) st->best_key = m;
}

/* one HME level: thread 0 prepared st->s_*, st->w_*; all threads stage and search, stripe by stripe.
 * level 0/1 read the source block from the decimated planes of the current picture, level 2
 * from the shared-memory SB. */
__device__ static void me_hme_search(const MeKernelArgs &A, int list, int level, MeSbState *st, const uint8_t *src,
                                     uint8_t *hsrc, uint8_t *win, uint32_t *posacc) {
    const int w = st->s_w, h = st->s_h;
    /* source block: every other row of the (decimated) SB */
    if (level == 2) {
        for (int i = threadIdx.x; i < w * h; i += ME_THREADS) {
            const int y = i / w, x = i - y * w;
            hsrc[y * st->s_pitch + x] = src[(2 * y) * ME_SRC_PITCH + x];
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
    if (threadIdx.x == 0) st->best_key = ~0ull;
    const MePlaneDev &rp = A.ref[list][level == 0 ? 2 : level == 1 ? 1 : 0];
    const int extra = 2 * (h - 1), sah = st->w_sah;
    int stripe = A.win_bytes / st->w_pitch - extra - 1; /* candidate rows the buffer can hold */
    if (stripe < 1) stripe = 1;
    for (int y0 = 0; y0 < sah; y0 += stripe) {
        const int n = min(stripe, sah - y0);
        me_stage_window(win, st->w_pitch, rp, st->w_x, st->w_y + y0, st->w_saw + w - 1, n + extra);
        __syncthreads();
        me_window_search(hsrc, win, st, posacc, y0, n);
        __syncthreads();
    }
}

/* prepare the window geometry for a search of saw x sah candidates (thread 0) */
__device__ __forceinline__ void me_set_window(MeSbState *st, int w, int h, int wx, int wy, int saw, int sah) {
    st->s_w = w; st->s_h = h; st->s_pitch = (w + 3) & ~3;
    st->w_x = wx; st->w_y = wy; st->w_saw = saw; st->w_sah = sah;
    st->w_pitch = (saw + w + 8 + 3) & ~3;
}


/* ---------------------------------------------------------------------------------------------
 * Full-pel search over the 85 partitions.
 * rb points at R(0,0) = top-left of the clipped search area inside the staged window (pitch
 * wpitch; the window also holds ME_HALO_X / ME_HALO_Y samples around it and the +16 columns of
 * the tail quirk).  Work item = (group of 4 horizontally adjacent candidates, one 16x16 block):
 * the thread accumulates the four 8x8 quadrants of its block for the four candidates from one
 * set of window words (every other row), sums them to the 16x16 SAD in registers, and the
 * 32x32 / 64x64 SADs come from shuffles over the 4 / 16 lanes that hold the other blocks of the
 * same candidates -- no shared-memory pyramid, no barrier inside the search.  Every thread keeps
 * running minima (2*sad << 32 | raster index) of "its" partitions; warps merge with shuffles,
 * the CTA with one atomicMin per warp and partition.
 * ------------------------------------------------------------------------------------------- */
/* a thread visits its candidates in increasing raster order, so "first minimum" is a strict < on the SAD alone; the
 * 64-bit (sad, index) key is only built when threads merge */
struct MeBest { uint32_t v, pos; };
__device__ __forceinline__ void me_best_upd(MeBest &b, uint32_t sad, uint32_t pos, bool valid) {
    const uint32_t v = 2u * sad;
    if (valid && v < b.v) { b.v = v; b.pos = pos; }
}
__device__ __forceinline__ unsigned long long me_best_key(const MeBest &b) {
    return b.v < ME_MAX_SAD ? ((unsigned long long)b.v << 32) | b.pos : ~0ull;
}
__device__ static void me_fullpel(const uint8_t *src, const uint8_t *rb, int wpitch, MeSbState *st, int list,
                                  unsigned long long *pu_key /* [85] smem */) {
    const int saw = st->saw[list], sah = st->sah[list];
    const int tail_from = saw & ~7; /* columns handled by the reference's one-position routine */
    const int ngx = (saw + 3) >> 2, ngroups = ngx * sah;
    const int t = threadIdx.x;
    for (int i = t; i < 85; i += ME_THREADS) pu_key[i] = ((unsigned long long)ME_MAX_SAD << 32) | 0xffffffffu;
    const int z16 = t & 15, gl = t >> 4;
    /* raster position of the 16x16 block with internal index z16 */
    const int b16 = (z16 & 1) | ((z16 & 4) >> 1) | ((z16 & 2) << 1) | (z16 & 8);
    const int bx = (b16 & 3) * 16, by = (b16 >> 2) * 16;
    const uint32_t *sp = (const uint32_t *)(src + by * ME_SRC_PITCH + bx);
    const int p2 = wpitch >> 1; /* two rows, in words */
    const bool skewed = (z16 == 12 || z16 == 13);
    const MeBest none = {ME_MAX_SAD, 0xffffffffu};
    MeBest r8[4] = {none, none, none, none}, r16 = none, r32 = none, r64 = none;
    for (int g0 = 0; g0 < ngroups; g0 += ME_THREADS / 16) {
        const int g = g0 + gl;
        const bool active = g < ngroups;
        const int ga = active ? g : 0, ys = ga / ngx, gx = ga - ys * ngx;
        /* the reference's single-position routine matches 16x16 blocks 12 and 13 sixteen samples further right
         * (EbMotionEstimation.c:855-856); reproduced */
        const int skew = (skewed && 4 * gx >= tail_from) ? 16 : 0;
        const uint32_t *wp = (const uint32_t *)(rb + (ys + by) * wpitch + 4 * gx + bx + skew);
        uint32_t acc[4][4]; /* [candidate][quadrant: 0 TL, 1 TR, 2 BL, 3 BR] */
#pragma unroll
        for (int j = 0; j < 4; j++) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0;
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const uint32_t *w = wp + r * p2;
            const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = w[4];
            const uint32_t s0 = sp[r * (ME_SRC_PITCH / 2)], s1 = sp[r * (ME_SRC_PITCH / 2) + 1], s2 = sp[r * (ME_SRC_PITCH / 2) + 2], s3 = sp[r * (ME_SRC_PITCH / 2) + 3];
            const int q = r < 4 ? 0 : 2;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t sh = 8 * j;
                acc[j][q] = me_absdiff_acc(j ? __funnelshift_r(w0, w1, sh) : w0, s0, acc[j][q]);
                acc[j][q] = me_absdiff_acc(j ? __funnelshift_r(w1, w2, sh) : w1, s1, acc[j][q]);
                acc[j][q + 1] = me_absdiff_acc(j ? __funnelshift_r(w2, w3, sh) : w2, s2, acc[j][q + 1]);
                acc[j][q + 1] = me_absdiff_acc(j ? __funnelshift_r(w3, w4, sh) : w3, s3, acc[j][q + 1]);
            }
        }
        const uint32_t pos0 = (uint32_t)(ys * saw + 4 * gx);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const bool valid = active && 4 * gx + j < saw;
            const uint32_t s16 = acc[j][0] + acc[j][1] + acc[j][2] + acc[j][3];
            uint32_t s32 = s16 + __shfl_xor_sync(0xffffffffu, s16, 1);
            s32 += __shfl_xor_sync(0xffffffffu, s32, 2);
            uint32_t s64 = s32 + __shfl_xor_sync(0xffffffffu, s32, 4);
            s64 += __shfl_xor_sync(0xffffffffu, s64, 8);
#pragma unroll
            for (int q = 0; q < 4; q++) me_best_upd(r8[q], acc[j][q], pos0 + j, valid);
            me_best_upd(r16, s16, pos0 + j, valid);
            me_best_upd(r32, s32, pos0 + j, valid);
            me_best_upd(r64, s64, pos0 + j, valid);
        }
    }
    /* the two half-warps hold different candidate groups of the same blocks */
    unsigned long long b8[4], b16k, b32k, b64k;
#pragma unroll
    for (int q = 0; q < 4; q++) b8[q] = me_shfl_min64(me_best_key(r8[q]), 16);
    b16k = me_shfl_min64(me_best_key(r16), 16); b32k = me_shfl_min64(me_best_key(r32), 16); b64k = me_shfl_min64(me_best_key(r64), 16);
    __syncthreads(); /* pu_key initialised */
    if ((t & 31) < 16) {
#pragma unroll
        for (int q = 0; q < 4; q++) if (b8[q] != ~0ull) atomicMin(&pu_key[21 + z16 * 4 + q], b8[q]);
        if (b16k != ~0ull) atomicMin(&pu_key[5 + z16], b16k);
        if ((z16 & 3) == 0 && b32k != ~0ull) atomicMin(&pu_key[1 + (z16 >> 2)], b32k);
        if (z16 == 0 && b64k != ~0ull) atomicMin(&pu_key[0], b64k);
    }
    __syncthreads();
}

/* ---------------------------------------------------------------------------------------------
 * su_pel_enable (EbMotionEstimation.c:3839-4258), one warp
 * ------------------------------------------------------------------------------------------- */
__device__ static void me_su_pel_enable(const uint32_t *bs, const uint32_t *bm, int tl, int *e32, int *e16, int *e8) {
    /* called by the 32 lanes of one warp: the sums over the 4 + 16 + 64 partitions by shuffles (integer sums: any order) */
    const int lane = threadIdx.x & 31;
    int sx32 = 0, sy32 = 0, sx8 = 0, sy8 = 0;
    uint32_t ss32 = 0, ss16 = 0, ss8 = 0;
    if (lane < 4) { sx32 = me_mvx(bm[1 + lane]); sy32 = me_mvy(bm[1 + lane]); ss32 = bs[1 + lane]; }
    if (lane < 16) ss16 = bs[5 + lane];
    for (int i = 21 + lane; i < 85; i += 32) { sx8 += me_mvx(bm[i]); sy8 += me_mvy(bm[i]); ss8 += bs[i]; }
    for (int o = 16; o > 0; o >>= 1) {
        sx32 += __shfl_xor_sync(0xffffffffu, sx32, o); sy32 += __shfl_xor_sync(0xffffffffu, sy32, o); ss32 += __shfl_xor_sync(0xffffffffu, ss32, o);
        ss16 += __shfl_xor_sync(0xffffffffu, ss16, o);
        sx8 += __shfl_xor_sync(0xffffffffu, sx8, o); sy8 += __shfl_xor_sync(0xffffffffu, sy8, o); ss8 += __shfl_xor_sync(0xffffffffu, ss8, o);
    }
    uint32_t ax = (uint32_t)(sx32 >> 2), ay = (uint32_t)(sy32 >> 2);
    const uint32_t mag32 = ax * ax + ay * ay, sad32 = ss32 >> 2;
    const uint32_t sad16 = ss16 >> 4;
    ax = (uint32_t)(sx8 >> 6); ay = (uint32_t)(sy8 >> 6);
    const uint32_t mag8 = ax * ax + ay * ay, sad8 = ss8 >> 6;
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
 * taps {-2,18,18,-2}, +16 >> 5, clipped (two rounded 8-bit passes for j).  R is read from the
 * staged full-pel window (rb = R(0,0), pitch rp); four output samples per step.
 * ------------------------------------------------------------------------------------------- */
__device__ __forceinline__ uint32_t me_clip8(int v) { return (uint32_t)min(max(v, 0), 255); }
/* horizontal: outputs x..x+3 from the words holding R(x-4..x-1), R(x..x+3), R(x+4..x+7) */
__device__ __forceinline__ uint32_t me_hfilt4(uint32_t wm, uint32_t w0, uint32_t wp) {
    const uint32_t K = 0xFE1212FEu; /* -2, 18, 18, -2 */
    const uint32_t v0 = me_clip8(me_dp4a_us(__funnelshift_r(wm, w0, 16), K, 16) >> 5);
    const uint32_t v1 = me_clip8(me_dp4a_us(__funnelshift_r(wm, w0, 24), K, 16) >> 5);
    const uint32_t v2 = me_clip8(me_dp4a_us(w0, K, 16) >> 5);
    const uint32_t v3 = me_clip8(me_dp4a_us(__funnelshift_r(w0, wp, 8), K, 16) >> 5);
    return v0 | (v1 << 8) | (v2 << 16) | (v3 << 24);
}
/* vertical: the same four columns on rows a,b,c,d; even and odd bytes as two 16-bit lanes each.
 * lane value = 18*(b+c) - 2*(a+d) + 16 + 1024 stays in [20, 10220]: no borrow between lanes;
 * (v + 1024) >> 5 == (v >> 5) + 32 also for negative v. */
__device__ __forceinline__ uint32_t me_vfilt4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    const uint32_t M = 0x00ff00ffu;
    const uint32_t se = (b & M) + (c & M), so = ((b >> 8) & M) + ((c >> 8) & M);
    const uint32_t te = (a & M) + (d & M), to = ((a >> 8) & M) + ((d >> 8) & M);
    uint32_t ve = ((se * 18u + 0x04100410u - 2u * te) >> 5) & 0x07ff07ffu;
    uint32_t vo = ((so * 18u + 0x04100410u - 2u * to) >> 5) & 0x07ff07ffu;
    ve = me_min_u16x2(me_max_u16x2(ve, 0x00200020u), 0x011f011fu) - 0x00200020u;
    vo = me_min_u16x2(me_max_u16x2(vo, 0x00200020u), 0x011f011fu) - 0x00200020u;
    return ve | (vo << 8);
}
__device__ static void me_interp(const uint8_t *rb, int rp, const MeSbState *st, int list, uint8_t *pb, uint8_t *ph, uint8_t *pj) {
    const int W = st->saw[list] + ME_SB - 1, H = st->sah[list] + ME_SB - 1;
    const int wasm = (W + 2 + 7) & ~7, is = st->istride[list];
    const int nw = wasm >> 2, is32 = is >> 2;
    const uint32_t rcp = 0xffffffffu / (uint32_t)nw + 1u; /* __umulhi(i, rcp) == i / nw */
    /* b: one word of four outputs per step */
    for (int it = threadIdx.x; it < nw * (H + 4); it += ME_THREADS) {
        const int y = (int)__umulhi((uint32_t)it, rcp), i = it - y * nw;
        const uint32_t *r = (const uint32_t *)(rb + (y - 2) * rp) + i;
        ((uint32_t *)pb)[y * is32 + i] = me_hfilt4(r[-1], r[0], r[1]);
    }
    /* h: a thread walks down a word column with a rolling window of four rows */
    const int nseg = nw >= ME_THREADS ? 1 : ME_THREADS / nw, per = (H + 1 + nseg - 1) / nseg;
    for (int it = threadIdx.x; it < nw * nseg; it += ME_THREADS) {
        const int seg = (int)__umulhi((uint32_t)it, rcp), i = it - seg * nw;
        const int ya = seg * per, yb = min(H + 1, ya + per);
        if (ya < yb) {
            const uint8_t *r = rb + (ya - 2) * rp + 4 * i - 1;
            uint32_t a = me_ld4(r), b = me_ld4(r + rp), c = me_ld4(r + 2 * rp);
            r += 3 * rp;
            for (int y = ya; y < yb; y++, r += rp) {
                const uint32_t d = me_ld4(r);
                ((uint32_t *)ph)[y * is32 + i] = me_vfilt4(a, b, c, d);
                a = b; b = c; c = d;
            }
        }
    }
    __syncthreads();
    for (int it = threadIdx.x; it < nw * nseg; it += ME_THREADS) {
        const int seg = (int)__umulhi((uint32_t)it, rcp), i = it - seg * nw;
        const int ya = seg * per, yb = min(H + 1, ya + per);
        if (ya < yb) {
            const uint32_t *q = (const uint32_t *)pb + ya * is32 + i;
            uint32_t a = q[0], b = q[is32], c = q[2 * is32];
            q += 3 * is32;
            for (int y = ya; y < yb; y++, q += is32) {
                const uint32_t d = *q;
                ((uint32_t *)pj)[y * is32 + i] = me_vfilt4(a, b, c, d);
                a = b; b = c; c = d;
            }
        }
    }
    __syncthreads();
}

/* ---------------------------------------------------------------------------------------------
 * Sub-pel refinement.  A prediction is one plane or the rounded average of two (MePred); the
 * distortion of a PU against it is computed by T threads, each taking every T-th (sub-sampled)
 * row, then reduced with shuffles inside the aligned lane group.
 * ------------------------------------------------------------------------------------------- */
struct MePred { const uint8_t *p1, *p2; int s1, s2; }; /* p2 == nullptr: single source */

/* SAD (sq == 0) or SSD (sq == 1) of one row of wwords words; srow is word aligned */
__device__ __forceinline__ uint32_t me_pred_row(const uint32_t *srow, const MePred &pr, int y, int wwords, int sq) {
    MeStream a, b;
    a.open(pr.p1 + (ptrdiff_t)y * pr.s1);
    uint32_t acc = 0;
    if (pr.p2) {
        b.open(pr.p2 + (ptrdiff_t)y * pr.s2);
        for (int k = 0; k < wwords; k++) {
            const uint32_t v = __vavgu4(a.next(), b.next());
            if (sq) { const uint32_t d = __vabsdiffu4(v, srow[k]); acc = me_dp4a_uu(d, d, acc); }
            else acc = me_absdiff_acc(v, srow[k], acc);
        }
    } else {
        for (int k = 0; k < wwords; k++) {
            const uint32_t v = a.next();
            if (sq) { const uint32_t d = __vabsdiffu4(v, srow[k]); acc = me_dp4a_uu(d, d, acc); }
            else acc = me_absdiff_acc(v, srow[k], acc);
        }
    }
    return acc;
}
/* distortion share of thread 'part' of T for the w x w PU at s (pitch ME_SRC_PITCH); method as the reference:
 * 0 SUB_SAD (even rows, doubled), 1 full SAD, 2 SSD (+ the full SAD in *sad) */
__device__ __forceinline__ uint32_t me_pred_part(const uint8_t *s, const MePred &pr, int w, int part, int T, int method, uint32_t *sad) {
    const int wwords = w >> 2;
    uint32_t d = 0, a = 0;
    if (method == 0) {
        for (int y = 2 * part; y < w; y += 2 * T) d += me_pred_row((const uint32_t *)(s + y * ME_SRC_PITCH), pr, y, wwords, 0);
        d <<= 1;
    } else if (method == 1) {
        for (int y = part; y < w; y += T) d += me_pred_row((const uint32_t *)(s + y * ME_SRC_PITCH), pr, y, wwords, 0);
    } else {
        for (int y = part; y < w; y += T) {
            d += me_pred_row((const uint32_t *)(s + y * ME_SRC_PITCH), pr, y, wwords, 1);
            a += me_pred_row((const uint32_t *)(s + y * ME_SRC_PITCH), pr, y, wwords, 0);
        }
    }
    *sad = a;
    return d;
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

/* sample sources of one reference list: the staged integer window and the three half-pel planes */
struct MeInterp {
    uint8_t       *pb, *ph, *pj; /* plane bases as written by me_interp */
    const uint8_t *rb;           /* R(0,0) of the list's search area: staged window, or the plane itself when the
                                    other list's window has replaced it */
    int            rp;           /* pitch of rb */
};

/* pointer conventions of the reference's half_pel_search_sb arguments */
__device__ __forceinline__ const uint8_t *me_PB(const MeInterp &ip, int is) { return ip.pb + 2 * is; }
__device__ __forceinline__ const uint8_t *me_PH(const MeInterp &ip) { return ip.ph + 1; }
__device__ __forceinline__ const uint8_t *me_PJ(const MeInterp &ip) { return ip.pj; }

__device__ __forceinline__ const uint8_t *me_qsrc_ptr(int is, const MeInterp &ip, MeQSrc q, int x, int y, int *stride) {
    x += q.ox; y += q.oy;
    switch (q.kind) {
    case 0: *stride = ip.rp; return ip.rb + x + y * ip.rp;
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

/* work units of the four size classes: class c has npu = 4^c PUs from index idx0, T = 32 >> ... threads per
 * (PU, candidate) task so that every unit is 16 words (8 for the 8x8 class) */
__device__ __forceinline__ int me_cls_idx0(int c) { return c == 0 ? 0 : c == 1 ? 1 : c == 2 ? 5 : 21; }
__device__ __forceinline__ int me_cls_npu(int c) { return 1 << (2 * c); }
__device__ __forceinline__ int me_cls_tlog(int c) { return c == 0 ? 5 : c == 1 ? 3 : c == 2 ? 1 : 0; }

/* sum over the aligned group of 2^tlog lanes (all 32 lanes call) */
__device__ __forceinline__ uint32_t me_group_sum(uint32_t v, int tlog) {
    for (int o = 0; o < tlog; o++) v += __shfl_xor_sync(0xffffffffu, v, 1 << o); /* tlog is warp-uniform */
    return v;
}

/* Half-pel then quarter-pel refinement of every enabled size class (cmask bit c), all at once.
 * dist/sadv: [85 * 9] indexed idx * 9 + candidate. */
__device__ static void me_subpel(const MeKernelArgs &A, const uint8_t *src, MeSbState *st, int list, const MeInterp &ip, int cmask,
                                 int qmask /* classes that also run quarter-pel */, uint32_t *bsad, uint32_t *bmv, uint32_t *bssd,
                                 uint8_t *dir, uint32_t *dist, uint32_t *sadv) {
    const int method = A.P.fractional_search_method, is = st->istride[list];
    const int sa_ox = st->sa_ox[list], sa_oy = st->sa_oy[list];
    const int ncand = (method == 2) ? 9 : 8;
    /* ---- half-pel: 8 candidates per PU (+ the full-pel position when the metric is SSD) ---- */
    for (int c = 0; c < 4; c++) {
        if (!(cmask & (1 << c))) continue;
        const int tlog = me_cls_tlog(c), T = 1 << tlog, idx0 = me_cls_idx0(c), nunits = me_cls_npu(c) * ncand * T;
        for (int u = threadIdx.x; u < nunits; u += ME_THREADS) {
            const int task = u >> tlog, part = u & (T - 1);
            const int pu = (ncand == 8) ? task >> 3 : task / 9, k = task - pu * ncand, idx = idx0 + pu;
            int px, py, w;
            me_pu_geom(idx, &px, &py, &w);
            const uint32_t mv = bmv[idx];
            const int xs = (me_mvx(mv) >> 2) - sa_ox, ys = (me_mvy(mv) >> 2) - sa_oy;
            const int base = xs + px + (ys + py) * is;
            MePred pr;
            pr.p2 = 0; pr.s2 = 0; pr.s1 = is;
            switch (k) { /* L, R, T, B, TL, TR, BR, BL, integer */
            case 0: pr.p1 = me_PB(ip, is) + base; break;
            case 1: pr.p1 = me_PB(ip, is) + base + 1; break;
            case 2: pr.p1 = me_PH(ip) + base; break;
            case 3: pr.p1 = me_PH(ip) + base + is; break;
            case 4: pr.p1 = me_PJ(ip) + base; break;
            case 5: pr.p1 = me_PJ(ip) + base + 1; break;
            case 6: pr.p1 = me_PJ(ip) + base + 1 + is; break;
            case 7: pr.p1 = me_PJ(ip) + base + is; break;
            default: pr.p1 = ip.rb + xs + px + (ys + py) * ip.rp; pr.s1 = ip.rp; break;
            }
            uint32_t a = 0;
            uint32_t d = me_pred_part(src + py * ME_SRC_PITCH + px, pr, w, part, T, method, &a);
            d = me_group_sum(d, tlog);
            if (method == 2) a = me_group_sum(a, tlog);
            if (part == 0) { dist[idx * 9 + k] = d; if (method == 2) sadv[idx * 9 + k] = a; }
        }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 85; idx += ME_THREADS) {
        const int c = idx == 0 ? 0 : idx < 5 ? 1 : idx < 21 ? 2 : 3;
        if (!(cmask & (1 << c))) continue;
        const uint32_t *d = dist + idx * 9;
        const int16_t x_mv = me_mvx(bmv[idx]), y_mv = me_mvy(bmv[idx]);
        const int dx[8] = {-2, 2, 0, 0, -2, 2, 2, -2}, dy[8] = {0, 0, -2, 2, -2, -2, 2, 2};
        uint32_t best = (method == 2) ? d[8] : bsad[idx];
        int win = -1;
        for (int k = 0; k < 8; k++) if (d[k] < best) { best = d[k]; win = k; }
        if (method == 2) bssd[idx] = best;
        if (win >= 0) {
            bsad[idx] = (method == 2) ? sadv[idx * 9 + win] : best;
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
    cmask &= qmask;
    if (!cmask) return;
    /* ---- quarter-pel: the three positions next to the winning half-pel direction.  The units of all enabled classes form
     * one index space (each class padded to whole warps) so that they run side by side instead of class after class ---- */
    {
        int off[5];
        off[0] = 0;
        for (int c = 0; c < 4; c++) {
            /* the 64x64 PU is refined on its top-left 32x32 only (the reference passes 32 as the size) */
            const int tl = (c == 0) ? 3 : me_cls_tlog(c);
            off[c + 1] = off[c] + ((cmask & (1 << c)) ? ((me_cls_npu(c) * 3 << tl) + 31) & ~31 : 0);
        }
        for (int g = threadIdx.x; g < off[4]; g += ME_THREADS) {
            const int c = g < off[1] ? 0 : g < off[2] ? 1 : g < off[3] ? 2 : 3;
            const int u = g - off[c];
            const int tlog = (c == 0) ? 3 : me_cls_tlog(c), T = 1 << tlog, idx0 = me_cls_idx0(c);
            const int nunits = me_cls_npu(c) * 3 * T;
            const int active = u < nunits;
            const int task = active ? u >> tlog : 0, part = u & (T - 1);
            const int pu = task / 3, j = task - pu * 3, idx = idx0 + pu;
            int px, py, w;
            me_pu_geom(idx, &px, &py, &w);
            if (c == 0) w = 32;
            const uint32_t mv = bmv[idx];
            const int16_t x_mv = me_mvx(mv), y_mv = me_mvy(mv);
            const int xs = ((x_mv + 2) >> 2) - sa_ox, ys = ((y_mv + 2) >> 2) - sa_oy;
            const int phase = (y_mv & 2) + ((x_mv & 2) >> 1);
            const int d0 = phase ? (dir[idx] + 4) & 7 : dir[idx];
            const int inv[8] = {4, 2, 5, 1, 6, 3, 7, 0}; /* ring code -> candidate L,R,T,B,TL,TR,BR,BL */
            const int k = inv[(d0 + j - 1) & 7];
            MePred pr;
            pr.p1 = me_qsrc_ptr(is, ip, me_qtab[phase][k][0], xs + px, ys + py, &pr.s1);
            pr.p2 = me_qsrc_ptr(is, ip, me_qtab[phase][k][1], xs + px, ys + py, &pr.s2);
            uint32_t a = 0, d = 0;
            if (active) d = me_pred_part(src + py * ME_SRC_PITCH + px, pr, w, part, T, method, &a);
            d = me_group_sum(d, tlog);
            if (method == 2) a = me_group_sum(a, tlog);
            if (active && part == 0) { dist[idx * 9 + k] = d; if (method == 2) sadv[idx * 9 + k] = a; }
        }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 85; idx += ME_THREADS) {
        const int c = idx == 0 ? 0 : idx < 5 ? 1 : idx < 21 ? 2 : 3;
        if (!(cmask & (1 << c))) continue;
        const int16_t x_mv = me_mvx(bmv[idx]), y_mv = me_mvy(bmv[idx]);
        const int dx[8] = {-1, 1, 0, 0, -1, 1, 1, -1}, dy[8] = {0, 0, -1, 1, -1, -1, 1, 1};
        const int ring[8] = {7, 3, 1, 5, 0, 2, 4, 6};
        const int phase = (y_mv & 2) + ((x_mv & 2) >> 1);
        const int d0 = phase ? (dir[idx] + 4) & 7 : dir[idx];
        uint32_t best = (method == 2) ? bssd[idx] : bsad[idx];
        int win = -1;
        for (int k = 0; k < 8; k++) {
            const int diff = (ring[k] - d0) & 7;
            if ((diff == 0 || diff == 1 || diff == 7) && dist[idx * 9 + k] < best) { best = dist[idx * 9 + k]; win = k; }
        }
        if (win >= 0) {
            if (method == 2) { bssd[idx] = best; bsad[idx] = sadv[idx * 9 + win]; }
            else bsad[idx] = best;
            bmv[idx] = me_mkmv(x_mv + dx[win], y_mv + dy[win]);
        }
    }
    __syncthreads();
}

/* prediction of one list at quarter-pel mv for the PU at (px,py): at most two planes to average
 * (select_buffer / quarter_pel_compensation) */
__device__ static void me_bipred_pred(int sa_ox, int sa_oy, int is, const MeInterp &ip, uint32_t mv, int px, int py, MePred *out) {
    const int16_t x = me_mvx(mv), y = me_mvy(mv);
    const int xs = (x >> 2) - sa_ox, ys = (y >> 2) - sa_oy;
    const int frac = (x & 3) + ((y & 3) << 2);
    const int as = ip.rp;
    const uint8_t *Ap = ip.rb + (xs + px) + (ys + py) * as;
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
/* one word of a prediction row */
__device__ __forceinline__ uint32_t me_pred_word(const MePred &p, int y, int k) {
    const uint32_t a = me_ld4(p.p1 + (ptrdiff_t)y * p.s1 + 4 * k);
    if (!p.p2) return a;
    return __vavgu4(a, me_ld4(p.p2 + (ptrdiff_t)y * p.s2 + 4 * k));
}


/* ---------------------------------------------------------------------------------------------
 * shared-memory layout of one CTA and the launch plan (shared by the host code and the emulation)
 * ------------------------------------------------------------------------------------------- */
struct MeSmemLayout {
    int src, hsrc, win, interp, bip0, pu_key, bsad, bmv, bssd, bip, dir, posacc, dist, sad, state, mbar, total;
};
__host__ __device__ static inline int me_align16(int v) { return (v + 15) & ~15; }
__host__ __device__ static inline MeSmemLayout me_smem_layout(int win_bytes, int interp_plane_bytes, int interp_in_smem, int bip0_bytes) {
    MeSmemLayout L;
    int o = 0;
    L.src = o; o += ME_SB * ME_SRC_PITCH;
    L.hsrc = o; o += 64 * 32;
    L.win = o; o += (win_bytes + 127) & ~127; /* TMA destinations are 128-byte aligned (src and hsrc keep it so) */
    L.interp = o; o += interp_in_smem ? 3 * me_align16(interp_plane_bytes) : 0; /* the planes of the list being searched */
    L.bip0 = o; o += me_align16(bip0_bytes);                                  /* list-0 predictions kept for bi-prediction */
    L.pu_key = o; o += me_align16(85 * 8);
    L.bsad = o; o += me_align16(2 * 85 * 4);
    L.bmv = o; o += me_align16(2 * 85 * 4);
    L.bssd = o; o += me_align16(2 * 85 * 4);
    L.bip = o; o += me_align16(85 * 4);
    L.dir = o; o += me_align16(85);
    L.posacc = o; o += 256 * 4;
    L.dist = o; o += me_align16(85 * 9 * 4);
    L.sad = o; o += me_align16(85 * 9 * 4);
    L.state = o; o += me_align16((int)sizeof(MeSbState));
    L.mbar = o; o += 16;
    L.total = o;
    return L;
}
/* pitch of the half-pel planes of a W-column search region: the columns the reference computes (W + 2 rounded up to 8) rounded up
 * to an ODD multiple of 8 bytes, so that rows two apart (one per lane in the sub-sampled metrics) fall into different banks */
__host__ __device__ static inline int me_istride(int W) { const int s = (W + 2 + 7) & ~7; return (s & 8) ? s : s + 8; }
/* pitch / rows of the staged full-pel window of a saw x sah search area */
__host__ __device__ static inline int me_fp_pitch(int saw) { return (saw + ME_SB - 1 + 16 + ME_HALO_X + 8 + 3) & ~3; }
__host__ __device__ static inline int me_fp_rows(int sah) { return sah + ME_SB - 1 + 2 * ME_HALO_Y; }
/* capacity of the window buffer: the full-pel window, and the largest HME stripe */
static inline int me_window_bytes(const SvtB200MeParams &P) {
    const int saw = P.search_area_width < 127 ? P.search_area_width : 127, sah = P.search_area_height < 127 ? P.search_area_height : 127;
    const int fp = me_fp_pitch(saw) * (me_fp_rows(sah) + 1);
    /* an HME area is searched in stripes of candidate rows; shorter stripes while that keeps the buffer at the size the
     * full-pel window needs anyway (more CTAs per SM), never below 8 rows */
    auto hme = [fp](int hsaw, int hsah, int w, int h) {
        const int pitch = (hsaw + w + 8 + 3) & ~3;
        int stripe = hsah < ME_HME_STRIPE ? hsah : ME_HME_STRIPE;
        while (stripe > 8 && pitch * (stripe + 2 * (h - 1) + 2) > fp) stripe -= 8;
        return pitch * (stripe + 2 * (h - 1) + 2);
    };
    auto mx = [](int a, int b) { return a > b ? a : b; };
    int best = fp;
    if (P.enable_hme_flag) {
        if (P.enable_hme_level0) {
            int tw = (P.hme_level0_total_search_area_width * P.hme_l0_multiplier_x) / 100, th = (P.hme_level0_total_search_area_height * P.hme_l0_multiplier_y) / 100;
            for (int i = 0; i < 2; i++) {
                tw = mx(tw, (P.hme_level0_search_area_in_width[i] * P.hme_l0_multiplier_x) / 100);
                th = mx(th, (P.hme_level0_search_area_in_height[i] * P.hme_l0_multiplier_y) / 100);
            }
            best = mx(best, hme(tw, th, 16, 8));
        }
        if (P.enable_hme_level1)
            for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++)
                best = mx(best, hme(mx(8, 2 * P.hme_level1_search_area_in_width[i]), P.hme_level1_search_area_in_height[j], 32, 16));
        if (P.enable_hme_level2)
            for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++)
                best = mx(best, hme(mx(8, 2 * P.hme_level2_search_area_in_width[i]), P.hme_level2_search_area_in_height[j], 64, 32));
    }
    return best + 16;
}
/* list-0 predictions of the partitions bi-prediction evaluates: one 64-wide image per size class, sub-sampled rows when
 * the metric is the sub-sampled SAD */
__host__ __device__ static inline int me_bip0_bytes(const SvtB200MeParams &P) {
    if (P.num_lists < 2) return 0;
    const int ncls = 2 + (P.cu16x16_mode == 0 ? 1 : 0) + ((P.cu16x16_mode == 0 && P.cu8x8_mode == 0) ? 1 : 0);
    return ncls * ME_SRC_PITCH * (P.fractional_search_method == 0 ? ME_SB / 2 : ME_SB);
}
static inline int me_interp_plane_bytes(const SvtB200MeParams &P) {
    const int saw = P.search_area_width < 127 ? P.search_area_width : 127, sah = P.search_area_height < 127 ? P.search_area_height : 127;
    return me_istride(saw + ME_SB - 1) * (sah + ME_SB - 1 + 4);
}

#ifdef SVTB200_CPU_EMU
#define ME_DYN_SMEM(name) uint8_t *name = emu_dyn_smem
#else
#define ME_DYN_SMEM(name) extern __shared__ __align__(128) uint8_t name[]
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
            /* no barrier here: me_hme_search ended with one, the fields above are thread 0's alone, and the window geometry thread 0
             * writes next is not read by the others before the barrier that follows it */
        }
}
/* ---------------------------------------------------------------------------------------------
 * the kernel: one CTA per super-block
 * ------------------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(ME_THREADS, 4) me_sb_kernel(const __grid_constant__ MeKernelArgs A) {
    ME_DYN_SMEM(smem);
    const SvtB200MeParams &P = A.P;
    const MeSmemLayout L = me_smem_layout(A.win_bytes, A.interp_plane_bytes, A.interp_in_smem, me_bip0_bytes(A.P));
    uint8_t  *src = smem + L.src, *hsrc = smem + L.hsrc, *win = smem + L.win;
    unsigned long long *pu_key = (unsigned long long *)(smem + L.pu_key);
    uint32_t *bsad = (uint32_t *)(smem + L.bsad), *bmv = (uint32_t *)(smem + L.bmv), *bssd = (uint32_t *)(smem + L.bssd);
    uint32_t *bip = (uint32_t *)(smem + L.bip), *posacc = (uint32_t *)(smem + L.posacc);
    uint8_t  *dir = smem + L.dir;
    uint32_t *dist = (uint32_t *)(smem + L.dist), *sadv = (uint32_t *)(smem + L.sad);
    uint8_t  *bip0 = smem + L.bip0;
    MeSbState *st = (MeSbState *)(smem + L.state);
    const int t = threadIdx.x, nl = P.num_lists;
    const int ipb = me_align16(A.interp_plane_bytes);
    MeTmaState tma;
    tma.phase = 0;
#if defined(__CUDA_ARCH__)
    tma.mbar = me_smem_u32(smem + L.mbar);
    if (A.tma && t == 0) me_mbar_init(tma.mbar);
#else
    tma.mbar = 0;
#endif
    /* persistent grid: a CTA keeps taking super-blocks from the picture's counter until none is left (without a counter block b
     * does super-block b) -- no tail of partly filled waves, and the work of a slow super-block is balanced by the others */
    for (int round = 0;; round++) {
    int sb;
    if (A.sb_counter) {
        __syncthreads(); /* everyone has finished the previous super-block (st, the result arrays and the windows are reused) */
        if (t == 0) st->next_sb = (int)atomicAdd(A.sb_counter, 1u);
        __syncthreads();
        sb = st->next_sb;
    } else {
        sb = round ? A.n_sb : (int)blockIdx.x;
    }
    if (sb >= A.n_work) break;
    sb += A.sb_first;
    MeInterp ip[2];
    for (int l = 0; l < 2; l++) {
        uint8_t *base = A.interp_in_smem ? smem + L.interp /* one set, reused by list 1 */
                                         : A.interp_scratch + ((size_t)sb * 2 + l) * 3 * ipb;
        ip[l].pb = base; ip[l].ph = base + ipb; ip[l].pj = base + 2 * ipb;
        ip[l].rb = 0; ip[l].rp = 0;
    }

    if (t == 0) {
        st->ox = (sb % A.sbs_w) * ME_SB; st->oy = (sb / A.sbs_w) * ME_SB;
        st->sbw = min(ME_SB, A.cur[0].w - st->ox); st->sbh = min(ME_SB, A.cur[0].h - st->oy);
        st->srh = 0; st->hme_x = 0; st->hme_y = 0;
        st->xsc = 0; st->ysc = 0; /* search centre of list 0 (list 1: reset where list 0 has finished with it) */
        for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) {
            st->l0x[a][b] = st->l0y[a][b] = st->l1x[a][b] = st->l1y[a][b] = st->l2x[a][b] = st->l2y[a][b] = 0;
            st->l0s[a][b] = st->l1s[a][b] = st->l2s[a][b] = 0;
        }
        st->interp_ok[0] = st->interp_ok[1] = 0;
        /* source SB: all 64 rows from the padded plane */
        if (A.tma & ME_TMA_SRC) me_tma_issue(&A.tm_cur, win, A.cur[0].ox + st->ox, A.cur[0].oy + st->oy, ME_SB * ME_SB, tma); /* dense box into the idle window buffer */
    }
    for (int i = t; i < 2 * 85; i += ME_THREADS) { bsad[i] = ME_MAX_SAD; bmv[i] = 0; bssd[i] = 0; }
    for (int i = t; i < 85; i += ME_THREADS) { bip[i] = 0; dir[i] = 0; }
    __syncthreads();
    if (A.tma & ME_TMA_SRC) {
        me_tma_wait(tma);
        for (int i = t; i < ME_SB * ME_SB / 4; i += ME_THREADS) ((uint32_t *)(src + (i >> 4) * ME_SRC_PITCH))[i & 15] = ((const uint32_t *)win)[i]; /* to the padded pitch */
    } else {
        me_stage_window(src, ME_SRC_PITCH, A.cur[0], st->ox, st->oy, ME_SB - 4, ME_SB);
    }
    __syncthreads();

    for (int list = 0; list < nl; list++) {
        const MePlaneDev &rp = A.ref[list][0];
        uint32_t *ls = bsad + list * 85, *lm = bmv + list * 85, *lq = bssd + list * 85;
        /* the running search centre (st->xsc, st->ysc) is (0,0) here; every thread carries a copy in registers: the decisions below
         * depend on CTA-wide sums every thread holds, so all threads take them redundantly instead of waiting for thread 0 */
        int16_t xsc_r = 0, ysc_r = 0;
        uint32_t sad_zero = 0;
        int have_zero = 0;
        if (P.temporal_layer_index > 0 || list == 0) {
            /* ---- test_search_area_bounds: O, (A == O), B, C, D and, for list 1, the mirrored list-0 MV */
            const int pad = ME_SB - 1, W0 = P.hme_level0_total_search_area_width, H0 = P.hme_level0_total_search_area_height;
            int16_t cdx[5], cdy[5];
            cdx[0] = 0; cdy[0] = 0;
            cdx[1] = me_clip_center(st->ox, (int16_t)W0, pad, rp.w); cdy[1] = me_clip_center(st->oy, 0, pad, rp.h);
            cdx[2] = me_clip_center(st->ox, 0, pad, rp.w); cdy[2] = me_clip_center(st->oy, (int16_t)(0 - H0), pad, rp.h);
            cdx[3] = me_clip_center(st->ox, 0, pad, rp.w); cdy[3] = me_clip_center(st->oy, (int16_t)H0, pad, rp.h);
            int16_t dirx = 0, diry = 0;
            int nc = 4;
            if (list == 1) {
                dirx = (int16_t)(0 - (me_mvx(bmv[0]) >> 2)); diry = (int16_t)(0 - (me_mvy(bmv[0]) >> 2));
                cdx[4] = me_clip_center(st->ox, dirx, pad, rp.w); cdy[4] = me_clip_center(st->oy, diry, pad, rp.h);
                nc = 5;
            } else { cdx[4] = 0; cdy[4] = 0; }
            uint32_t cs[5];
            me_sb_sad_multi(src, st, rp, nc, cdx, cdy, cs, st->red);
            const uint32_t c_o = cs[0], c_b = cs[1], c_c = cs[2], c_d = cs[3];
            const uint64_t c_dir = (list == 1) ? (uint64_t)cs[4] : 0xFFFFFFFFFFFFFull;
            sad_zero = c_o; have_zero = 1;
            {
                uint64_t best = c_o;
                if (c_b < best) best = c_b;
                if (c_c < best) best = c_c;
                if (c_d < best) best = c_d;
                if (c_dir < best) best = c_dir;
                /* tie order O, A(==O), B, C, direct, D; unclipped offsets */
                if (best == c_o) { xsc_r = 0; ysc_r = 0; }
                else if (best == c_b) { xsc_r = (int16_t)W0; ysc_r = 0; }
                else if (best == c_c) { xsc_r = 0; ysc_r = (int16_t)(0 - H0); }
                else if (best == c_dir) { xsc_r = dirx; ysc_r = diry; }
                else { xsc_r = 0; ysc_r = (int16_t)H0; }
                if (t == 0) { st->xsc = xsc_r; st->ysc = ysc_r; } /* thread 0's hierarchical-search set-up reads it from there */
            }
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
                /* (no barrier: only thread 0 touches these fields before me_hme_level's own) */
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
                xsc_r = st->xsc; ysc_r = st->ysc;
            }
        }
        if (A.trace && t == 0) { A.trace[sb].hme_center_x[list] = xsc_r; A.trace[sb].hme_center_y[list] = ysc_r; }
        /* ---- (0,0) versus the HME centre */
        if (xsc_r != 0 || ysc_r != 0) {
            int16_t cdx[5] = {0, 0, 0, 0, 0}, cdy[5] = {0, 0, 0, 0, 0};
            const int16_t cx = me_clip_center(st->ox, xsc_r, ME_SB - 1, rp.w), cy = me_clip_center(st->oy, ysc_r, ME_SB - 1, rp.h);
            cdx[0] = cx; cdy[0] = cy;
            uint32_t cs[5];
            me_sb_sad_multi(src, st, rp, have_zero ? 1 : 2, cdx, cdy, cs, st->red); /* candidate 1 = (0,0) when not yet known */
            const uint32_t hme = cs[0], zero = have_zero ? sad_zero : cs[1];
            if (zero <= hme) { xsc_r = 0; ysc_r = 0; } else { xsc_r = cx; ysc_r = cy; } /* every thread holds the sums: no hand-off */
        }
        if (t == 0) {
            int16_t saw = (int16_t)min(P.search_area_width, 127), sah = (int16_t)min(P.search_area_height, 127);
            int16_t xo = (int16_t)(xsc_r - (saw >> 1)), yo = (int16_t)(ysc_r - (sah >> 1));
            me_clip_window(st->ox, &xo, &saw, ME_SB - 1, P.luma_width);
            me_clip_window(st->oy, &yo, &sah, ME_SB - 1, P.luma_height);
            st->sa_ox[list] = xo; st->sa_oy[list] = yo; st->saw[list] = saw; st->sah[list] = sah;
            st->fp_pitch = me_fp_pitch(saw);
            if (A.trace) { A.trace[sb].sa_ox[list] = xo; A.trace[sb].sa_oy[list] = yo; A.trace[sb].sa_w[list] = saw; A.trace[sb].sa_h[list] = sah; }
        }
        __syncthreads();
        /* ---- full-pel search: the window (with the halo the half-pel filters need) is staged once */
        const int fpp = st->fp_pitch;
        const uint8_t *rb = win + ME_HALO_Y * fpp + ME_HALO_X;
        me_stage_window(win, fpp, rp, st->ox + st->sa_ox[list] - ME_HALO_X, st->oy + st->sa_oy[list] - ME_HALO_Y,
                        st->saw[list] + ME_SB - 1 + 16 + ME_HALO_X, me_fp_rows(st->sah[list]));
        __syncthreads();
        me_fullpel(src, rb, fpp, st, list, pu_key);
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
        if (t < 32) { /* warp 0 */
            int e32 = 0, e16 = 0, e8 = 0, eq = 0;
            if (P.fractional_search_model == 0) e32 = e16 = e8 = eq = 1;
            else if (P.fractional_search_model == 1) { me_su_pel_enable(ls, lm, P.temporal_layer_index, &e32, &e16, &e8); eq = 1; }
            if (t == 0) {
                st->xsc = 0; st->ysc = 0; /* for the next list: nothing reads the centre of this list any more */
                st->e32 = e32; st->e16 = e16; st->e8 = e8; st->eq = eq;
                const int W = st->saw[list] + ME_SB - 1;
                st->istride[list] = me_istride(W);
                if (A.trace) A.trace[sb].subpel_enable[list] = (uint8_t)(e32 | (e16 << 1) | (e8 << 2) | (eq << 3));
            }
        }
        __syncthreads();
        ip[list].rb = rb; ip[list].rp = fpp;
        if (st->e32 || st->e16 || st->e8 || st->eq) {
            me_interp(rb, fpp, st, list, ip[list].pb, ip[list].ph, ip[list].pj);
            if (t == 0) st->interp_ok[list] = 1;
            int cmask = 0;
            if (P.fractional_search64x64) cmask |= 1;
            if (st->e32) cmask |= 2;
            if (st->e16 && P.cu16x16_mode == 0) cmask |= 4;
            if (st->e8 && P.cu8x8_mode != 1) cmask |= 8;
            if (cmask) me_subpel(A, src, st, list, ip[list], cmask, st->eq ? 15 : 1, ls, lm, lq, dir, dist, sadv);
        }
        __syncthreads();
        if (list == 0 && nl == 2) {
            /* the list-0 prediction of every partition bi-prediction will evaluate, at the final list-0 vectors: one 64-wide
             * image per size class (even rows only for the sub-sampled metric) */
            const int sub = (P.fractional_search_method == 0), rsh = sub ? 1 : 0;
            for (int cls = 0; cls < 4; cls++) {
                if (cls == 3 && P.cu8x8_mode != 0) continue;
                if (cls >= 2 && P.cu16x16_mode != 0) continue;
                const int tlog = 5 - cls, T = 1 << tlog, idx0 = me_cls_idx0(cls), nunits = me_cls_npu(cls) * T;
                uint8_t *img = bip0 + cls * (ME_SRC_PITCH * (ME_SB >> rsh));
                for (int u = t; u < nunits; u += ME_THREADS) {
                    const int idx = idx0 + (u >> tlog), part = u & (T - 1);
                    int px, py, w;
                    me_pu_geom(idx, &px, &py, &w);
                    MePred p0;
                    me_bipred_pred(st->sa_ox[0], st->sa_oy[0], st->istride[0], ip[0], bmv[idx], px, py, &p0);
                    const int wwords = w >> 2;
                    for (int y = sub ? 2 * part : part; y < w; y += sub ? 2 * T : T) {
                        uint32_t *d = (uint32_t *)(img + ((py + y) >> rsh) * ME_SRC_PITCH + px);
                        for (int k = 0; k < wwords; k++) d[k] = me_pred_word(p0, y, k);
                    }
                }
            }
            __syncthreads();
        }
    }

    /* ---- bi-prediction: SAD(src, avg(pred L0, pred L1)) at the two uni-pred winners; the list-0 predictions were put
     * aside (bip0) before list 1 took over the half-pel planes */
    if (nl == 2) {
        const int sub = (P.fractional_search_method == 0), rsh = sub ? 1 : 0;
        /* one (sub-sampled) row of one PU per unit: 64x64 -> 32 lanes, 32x32 -> 16, 16x16 -> 8, 8x8 -> 4 */
        for (int cls = 0; cls < 4; cls++) {
            if (cls == 3 && P.cu8x8_mode != 0) continue;
            if (cls >= 2 && P.cu16x16_mode != 0) continue;
            const int tlog = 5 - cls, T = 1 << tlog, idx0 = me_cls_idx0(cls), nunits = me_cls_npu(cls) * T;
            const uint8_t *img = bip0 + cls * (ME_SRC_PITCH * (ME_SB >> rsh));
            for (int u = t; u < nunits; u += ME_THREADS) {
                const int idx = idx0 + (u >> tlog), part = u & (T - 1);
                int px, py, w;
                me_pu_geom(idx, &px, &py, &w);
                MePred p1;
                me_bipred_pred(st->sa_ox[1], st->sa_oy[1], st->istride[1], ip[1], bmv[85 + idx], px, py, &p1);
                const int wwords = w >> 2;
                uint32_t acc = 0;
                for (int y = sub ? 2 * part : part; y < w; y += sub ? 2 * T : T) {
                    const uint32_t *srow = (const uint32_t *)(src + (py + y) * ME_SRC_PITCH + px);
                    const uint32_t *p0 = (const uint32_t *)(img + ((py + y) >> rsh) * ME_SRC_PITCH + px);
                    for (int k = 0; k < wwords; k++) acc = me_absdiff_acc(__vavgu4(p0[k], me_pred_word(p1, y, k)), srow[k], acc);
                }
                acc = me_group_sum(acc, tlog);
                if (part == 0) bip[idx] = sub ? acc << 1 : acc;
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
    } /* next super-block */
}

#endif /* SVTB200_ME_KERNEL_CUH */
