/*
 * deblock_kernel.cuh -- S2: in-loop deblocking of one picture, replacing eb_vp9_loop_filter_frame()
 * (Source/Lib/VPX/vp9_loopfilter.c:1521 -> loop_filter_rows :1456 -> eb_vp9_adjust_mask :786,
 * eb_vp9_filter_block_plane_ss00/_ss11 :1238/:1342, eb_vp9_lpf_* Source/Lib/VPX/loopfilter.c:103-329).
 *
 * VP9 fixes the filtering order: super-blocks in raster order, and inside each the vertical edges of a plane
 * before its horizontal edges.  Filters reach 8 samples across super-block borders, so SB (r,c) depends on
 * (r,c-1) and (r-1,c+1): a wavefront of sbs_w + 2*(sbs_h-1) dependent steps, i.e. a latency chain.
 *
 * One persistent CTA (256 threads) per super-block ROW walks its row left to right, so the left dependency never
 * leaves shared memory: the right 8 columns of tile c become the left context of tile c+1 by a shared-memory copy.
 * Warps 0-3 filter (one thread per sample row for the vertical edges, one per sample column for the horizontal
 * edges of luma / U / V, edges sequential along the thread's line); warps 4-7 move data: while tile c is filtered
 * they write tile c-1 back (8-byte stores), publish the row's progress after a device-scope fence, and fetch tile
 * c+1 (its own samples, which nobody has modified yet, and the 8-sample strip above it once the row above has
 * progressed far enough) with 8-byte L1-bypassing loads into the other tile buffer.  Rows take their index from an
 * atomic ticket, so a CTA only ever waits for CTAs that already run -- no deadlock.  Two CTA barriers per SB.
 *
 * The file also compiles as plain C++ under csrc/emu/cuda_emu.h for the CPU-side logic tests.
 */
#ifndef SVTB200_DEBLOCK_KERNEL_CUH
#define SVTB200_DEBLOCK_KERNEL_CUH
#include <stdint.h>
#include "../../include/svtb200.h"

#define LF_THREADS 256       /* 128 filtering + 128 data-moving threads */
#define LF_HALF 128
#define LF_YP 76 /* luma tile pitch (odd number of words: conflict-free column access) */
#define LF_CP 44 /* chroma tile pitch */
#define LF_FLAG_PAD 32 /* progress words are 128 bytes apart: rows poll and publish without sharing a line */

struct LfKernelArgs {
    SvtB200LfParams P;
    uint8_t *y, *u, *v;
    int32_t  y_stride, uv_stride;
    const SvtB200LfMask *lfm;
    int32_t  lfm_stride;
    const SvtB200LfThresh *thr; /* [64] */
    int32_t *ticket;            /* zeroed before the launch */
    int32_t *done;              /* one word per SB row: (gen << 10) | number of SBs of the row that are final in memory */
    int32_t  gen;
    int32_t  sbs_w, sbs_h;
    unsigned long long *trace;  /* optional (tools/lf_trace.py): [row][sb][8] SM clock stamps */
};

#ifdef SVTB200_CPU_EMU
static inline unsigned long long lf_clock() { return 0; }
#define LF_SHARED16 alignas(16) static
#define LF_DYN_SMEM(name) uint8_t *name = emu_dyn_smem
static inline uint8_t lf_ldcg(const uint8_t *p) { return *p; }
static inline uint2 lf_ldcg8(const uint8_t *p) { uint2 v; memcpy(&v, p, 8); return v; }
static inline void lf_fence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void lf_sleep() {}
static inline void lf_sync_movers() { emu_sync_upper_half(); }
static inline int lf_flag_load(const int32_t *p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
static inline void lf_flag_store(int32_t *p, int v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
#else
__device__ __forceinline__ unsigned long long lf_clock() { return (unsigned long long)clock64(); }
#define LF_SHARED16 __shared__ __align__(16)
#define LF_DYN_SMEM(name) extern __shared__ __align__(16) uint8_t name[]
__device__ __forceinline__ uint8_t lf_ldcg(const uint8_t *p) { return __ldcg(p); }
__device__ __forceinline__ uint2 lf_ldcg8(const uint8_t *p) { return __ldcg((const uint2 *)p); }
__device__ __forceinline__ void lf_fence() { __threadfence(); }
__device__ __forceinline__ void lf_sleep() { __nanosleep(20); }
__device__ __forceinline__ void lf_sync_movers() { asm volatile("bar.sync 1, 128;" ::: "memory"); } /* threads 128..255 */
__device__ __forceinline__ int lf_flag_load(const int32_t *p) { return *(const volatile int32_t *)p; }
__device__ __forceinline__ void lf_flag_store(int32_t *p, int v) { *(volatile int32_t *)p = v; }
#endif

__device__ __forceinline__ int lf_abs(int v) { return v < 0 ? -v : v; }
__device__ __forceinline__ int lf_sc(int t) { return t < -128 ? -128 : t > 127 ? 127 : t; } /* signed char clamp */

/* one sample position across an edge; s points at q0, a = distance between taps; width = 4 / 8 / 16 */
__device__ __forceinline__ void lf_filter_across(uint8_t *s, int a, int width, int blim, int lim, int thr) {
    const int p3 = s[-4 * a], p2 = s[-3 * a], p1 = s[-2 * a], p0 = s[-a], q0 = s[0], q1 = s[a], q2 = s[2 * a], q3 = s[3 * a];
    const bool mask = !(lf_abs(p3 - p2) > lim || lf_abs(p2 - p1) > lim || lf_abs(p1 - p0) > lim || lf_abs(q1 - q0) > lim ||
                        lf_abs(q2 - q1) > lim || lf_abs(q3 - q2) > lim || lf_abs(p0 - q0) * 2 + lf_abs(p1 - q1) / 2 > blim);
    const bool flat = width >= 8 && !(lf_abs(p1 - p0) > 1 || lf_abs(q1 - q0) > 1 || lf_abs(p2 - p0) > 1 || lf_abs(q2 - q0) > 1 ||
                                      lf_abs(p3 - p0) > 1 || lf_abs(q3 - q0) > 1);
    if (flat && mask) {
        if (width == 16) {
            const int p7 = s[-8 * a], p6 = s[-7 * a], p5 = s[-6 * a], p4 = s[-5 * a], q4 = s[4 * a], q5 = s[5 * a], q6 = s[6 * a], q7 = s[7 * a];
            const bool flat2 = !(lf_abs(p4 - p0) > 1 || lf_abs(q4 - q0) > 1 || lf_abs(p5 - p0) > 1 || lf_abs(q5 - q0) > 1 ||
                                 lf_abs(p6 - p0) > 1 || lf_abs(q6 - q0) > 1 || lf_abs(p7 - p0) > 1 || lf_abs(q7 - q0) > 1);
            if (flat2) { /* 15-tap smoothing as a running window sum with replicated ends */
                int sum = p7 * 7 + p6 * 2 + p5 + p4 + p3 + p2 + p1 + p0 + q0; /* window of the p6 output */
                s[-7 * a] = (uint8_t)((sum + 8) >> 4);
                sum += q1 - p7 + p5 - p6; s[-6 * a] = (uint8_t)((sum + 8) >> 4);
                sum += q2 - p7 + p4 - p5; s[-5 * a] = (uint8_t)((sum + 8) >> 4);
                sum += q3 - p7 + p3 - p4; s[-4 * a] = (uint8_t)((sum + 8) >> 4);
                sum += q4 - p7 + p2 - p3; s[-3 * a] = (uint8_t)((sum + 8) >> 4);
                sum += q5 - p7 + p1 - p2; s[-2 * a] = (uint8_t)((sum + 8) >> 4);
                sum += q6 - p7 + p0 - p1; s[-1 * a] = (uint8_t)((sum + 8) >> 4);
                sum += q7 - p7 + q0 - p0; s[0] = (uint8_t)((sum + 8) >> 4);
                sum += q7 - p6 + q1 - q0; s[1 * a] = (uint8_t)((sum + 8) >> 4);
                sum += q7 - p5 + q2 - q1; s[2 * a] = (uint8_t)((sum + 8) >> 4);
                sum += q7 - p4 + q3 - q2; s[3 * a] = (uint8_t)((sum + 8) >> 4);
                sum += q7 - p3 + q4 - q3; s[4 * a] = (uint8_t)((sum + 8) >> 4);
                sum += q7 - p2 + q5 - q4; s[5 * a] = (uint8_t)((sum + 8) >> 4);
                sum += q7 - p1 + q6 - q5; s[6 * a] = (uint8_t)((sum + 8) >> 4);
                return;
            }
        }
        /* 7-tap smoothing */
        s[-3 * a] = (uint8_t)((3 * p3 + 2 * p2 + p1 + p0 + q0 + 4) >> 3);
        s[-2 * a] = (uint8_t)((2 * p3 + p2 + 2 * p1 + p0 + q0 + q1 + 4) >> 3);
        s[-1 * a] = (uint8_t)((p3 + p2 + p1 + 2 * p0 + q0 + q1 + q2 + 4) >> 3);
        s[0]      = (uint8_t)((p2 + p1 + p0 + 2 * q0 + q1 + q2 + q3 + 4) >> 3);
        s[1 * a]  = (uint8_t)((p1 + p0 + q0 + 2 * q1 + q2 + 2 * q3 + 4) >> 3);
        s[2 * a]  = (uint8_t)((p0 + q0 + q1 + 2 * q2 + 3 * q3 + 4) >> 3);
        return;
    }
    { /* narrow filter on the sign-flipped representation, every step saturated to int8 */
        const int ps1 = (int8_t)(p1 ^ 0x80), ps0 = (int8_t)(p0 ^ 0x80), qs0 = (int8_t)(q0 ^ 0x80), qs1 = (int8_t)(q1 ^ 0x80);
        const bool hev = lf_abs(p1 - p0) > thr || lf_abs(q1 - q0) > thr;
        int f = hev ? lf_sc(ps1 - qs1) : 0;
        f = mask ? lf_sc(f + 3 * (qs0 - ps0)) : 0;
        const int f1 = lf_sc(f + 4) >> 3, f2 = lf_sc(f + 3) >> 3;
        s[0]  = (uint8_t)(lf_sc(qs0 - f1) ^ 0x80);
        s[-a] = (uint8_t)(lf_sc(ps0 + f2) ^ 0x80);
        f = hev ? 0 : (f1 + 1) >> 1;
        s[a]      = (uint8_t)(lf_sc(qs1 - f) ^ 0x80);
        s[-2 * a] = (uint8_t)(lf_sc(ps1 + f) ^ 0x80);
    }
}

/* eb_vp9_adjust_mask (vp9_loopfilter.c:786-899) on a shared-memory copy of the SB's mask */
__device__ static void lf_adjust_mask(int mi_rows, int mi_cols, int mi_row, int mi_col, SvtB200LfMask *m) {
    m->left_y[2] |= m->left_y[3]; m->above_y[2] |= m->above_y[3];
    m->left_uv[2] |= m->left_uv[3]; m->above_uv[2] |= m->above_uv[3];
    const uint64_t lb = 0x1111111111111111ull, ab = 0x000000ff000000ffull;
    m->left_y[1] |= m->left_y[0] & lb; m->left_y[0] &= ~lb;
    m->above_y[1] |= m->above_y[0] & ab; m->above_y[0] &= ~ab;
    m->left_uv[1] |= m->left_uv[0] & 0x1111; m->left_uv[0] &= (uint16_t)~0x1111;
    m->above_uv[1] |= m->above_uv[0] & 0x000f; m->above_uv[0] &= (uint16_t)~0x000f;
    if (mi_row + 8 > mi_rows) {
        const uint64_t rows = (uint64_t)(mi_rows - mi_row);
        const uint64_t my = ((uint64_t)1 << (rows << 3)) - 1;
        const uint16_t muv = (uint16_t)(((uint16_t)1 << (((rows + 1) >> 1) << 2)) - 1);
        for (int i = 0; i < 3; i++) { m->left_y[i] &= my; m->above_y[i] &= my; m->left_uv[i] &= muv; m->above_uv[i] &= muv; }
        m->int_4x4_y &= my; m->int_4x4_uv &= muv;
        if (rows == 1) { m->above_uv[1] |= m->above_uv[2]; m->above_uv[2] = 0; }
        if (rows == 5) { m->above_uv[1] |= m->above_uv[2] & 0xff00; m->above_uv[2] &= (uint16_t)~(m->above_uv[2] & 0xff00); }
    }
    if (mi_col + 8 > mi_cols) {
        const uint64_t cols = (uint64_t)(mi_cols - mi_col);
        const uint64_t my = (((uint64_t)1 << cols) - 1) * 0x0101010101010101ull;
        const uint16_t muv = (uint16_t)(((1 << ((cols + 1) >> 1)) - 1) * 0x1111);
        const uint16_t muvi = (uint16_t)(((1 << (cols >> 1)) - 1) * 0x1111);
        for (int i = 0; i < 3; i++) { m->left_y[i] &= my; m->above_y[i] &= my; m->left_uv[i] &= muv; m->above_uv[i] &= muv; }
        m->int_4x4_y &= my; m->int_4x4_uv &= muvi;
        if (cols == 1) { m->left_uv[1] |= m->left_uv[2]; m->left_uv[2] = 0; }
        if (cols == 5) { m->left_uv[1] |= m->left_uv[2] & 0xcccc; m->left_uv[2] &= (uint16_t)~(m->left_uv[2] & 0xcccc); }
    }
    if (mi_col == 0) for (int i = 0; i < 3; i++) { m->left_y[i] &= 0xfefefefefefefefeull; m->left_uv[i] &= 0xeeee; }
}

/* per-plane view of the adjusted masks: row-major bit grids of bw x bw blocks */
struct LfPlaneMasks {
    uint64_t l16, l8, l4, a16, a8, a4, i4v, i4h;
    const uint8_t *lfl;
    int bw, rows /* block rows present */, px_rows /* sample rows present */, top;
};

/* lowest set bit of a non-zero word */
__device__ __forceinline__ int lf_ctz(uint32_t v) {
#ifdef SVTB200_CPU_EMU
    return __builtin_ctz(v);
#else
    return __ffs((int)v) - 1;
#endif
}

/* all vertical edges crossing sample row 'row' (0..8*bw-1) of the tile; t points at tile sample (0,row) of the SB.
 * The row's bits of the four masks are extracted once; only block columns that carry an edge are visited. */
__device__ static void lf_vertical_row(uint8_t *t, int row, const LfPlaneMasks &M, const SvtB200LfThresh *thr) {
    if (row >= M.px_rows) return;
    const int r = row >> 3, pr = r & ~1, sh = r * M.bw, osh = (r ^ 1) * M.bw;
    const uint32_t bwm = (1u << M.bw) - 1u;
    const uint32_t l16 = (uint32_t)(M.l16 >> sh) & bwm, l8 = (uint32_t)(M.l8 >> sh) & bwm, l4 = (uint32_t)(M.l4 >> sh) & bwm;
    const uint32_t i4 = (uint32_t)(M.i4v >> sh) & bwm, o16 = (uint32_t)(M.l16 >> osh) & bwm;
    for (uint32_t any = l16 | l8 | l4 | i4; any; any &= any - 1) {
        const int c = lf_ctz(any);
        uint8_t *s = t + c * 8;
        const SvtB200LfThresh own = thr[M.lfl[sh + c]];
        if (l16 >> c & 1) {
            /* a 16-wide edge present in both rows of a row pair takes the thresholds of the pair's first row */
            const SvtB200LfThresh th = (o16 >> c & 1) ? thr[M.lfl[pr * M.bw + c]] : own;
            lf_filter_across(s, 1, 16, th.mblim, th.lim, th.hev_thr);
        }
        if (l8 >> c & 1) lf_filter_across(s, 1, 8, own.mblim, own.lim, own.hev_thr);
        if (l4 >> c & 1) lf_filter_across(s, 1, 4, own.mblim, own.lim, own.hev_thr);
        if (i4 >> c & 1) lf_filter_across(s + 4, 1, 4, own.mblim, own.lim, own.hev_thr);
    }
}

/* all horizontal edges crossing sample column 'col'; t points at tile sample (col,0) of the SB.  The column's bits are
 * gathered once (one bit per block row); only block rows that carry an edge are visited. */
__device__ static void lf_horizontal_col(uint8_t *t, int col, int pitch, const LfPlaneMasks &M, const SvtB200LfThresh *thr) {
    const int c = col >> 3;
    uint32_t a16 = 0, a8 = 0, a4 = 0, i4 = 0;
    for (int r = 0; r < M.rows; r++) {
        const int bit = r * M.bw + c;
        a16 |= (uint32_t)(M.a16 >> bit & 1) << r; a8 |= (uint32_t)(M.a8 >> bit & 1) << r;
        a4 |= (uint32_t)(M.a4 >> bit & 1) << r; i4 |= (uint32_t)(M.i4h >> bit & 1) << r;
    }
    for (uint32_t any = a16 | a8 | a4 | i4; any; any &= any - 1) {
        const int r = lf_ctz(any), bit = r * M.bw + c;
        uint8_t *s = t + (r * 8) * pitch;
        const bool top = M.top && r == 0;
        const SvtB200LfThresh own = thr[M.lfl[bit]];
        if (!top && (a16 >> r & 1)) {
            /* adjacent 16-wide edges are paired from the start of each run; a pair shares the first block's thresholds */
            int start = c;
            while (start > 0 && (M.a16 >> (r * M.bw + start - 1) & 1)) start--;
            const SvtB200LfThresh th = thr[M.lfl[r * M.bw + c - ((c - start) & 1)]];
            lf_filter_across(s, pitch, 16, th.mblim, th.lim, th.hev_thr);
        } else {
            if (!top && (a8 >> r & 1)) lf_filter_across(s, pitch, 8, own.mblim, own.lim, own.hev_thr);
            else if (!top && (a4 >> r & 1)) lf_filter_across(s, pitch, 4, own.mblim, own.lim, own.hev_thr);
            if ((i4 >> r & 1) && r * 8 + 4 < M.px_rows) lf_filter_across(s + 4 * pitch, pitch, 4, own.mblim, own.lim, own.hev_thr);
        }
    }
}

/* Copy the samples [x_lo, x_hi) x [y_lo, y_hi) (SB-relative; the tile keeps SB sample (0,0) at (8,8)) between the
 * picture plane and the shared-memory tile.  vec: x bounds, plane pointer and stride are multiples of 8 -> 8-byte
 * global accesses (the tile pitch is only 4-byte aligned: two words on the shared side). */
template <int PITCH>
__device__ __forceinline__ void lf_tile_load(uint8_t *tile, const uint8_t *g, int gstride, int x_lo, int x_hi, int y_lo, int y_hi, bool vec, int tid,
                                             int nthr) {
    if (x_hi <= x_lo || y_hi <= y_lo) return;
    if (vec) {
        const int nx = (x_hi - x_lo) >> 3, n = nx * (y_hi - y_lo);
        for (int i = tid; i < n; i += nthr) {
            const int ry = i / nx, yy = y_lo + ry, xx = x_lo + (i - ry * nx) * 8;
            const uint2 v = lf_ldcg8(g + (ptrdiff_t)yy * gstride + xx);
            uint32_t *d = (uint32_t *)(tile + (yy + 8) * PITCH + xx + 8);
            d[0] = v.x; d[1] = v.y;
        }
    } else {
        const int nx = x_hi - x_lo, n = nx * (y_hi - y_lo);
        for (int i = tid; i < n; i += nthr) {
            const int ry = i / nx, yy = y_lo + ry, xx = x_lo + (i - ry * nx);
            tile[(yy + 8) * PITCH + xx + 8] = lf_ldcg(g + (ptrdiff_t)yy * gstride + xx);
        }
    }
}
template <int PITCH, int NX = 0 /* (x_hi - x_lo) / 8 when known at compile time */>
__device__ __forceinline__ void lf_tile_store(const uint8_t *tile, uint8_t *g, int gstride, int x_lo, int x_hi, int y_lo, int y_hi, bool vec, int tid,
                                              int nthr) {
    if (x_hi <= x_lo || y_hi <= y_lo) return;
    if (vec) {
        const int nx = NX ? NX : (x_hi - x_lo) >> 3, n = nx * (y_hi - y_lo);
        for (int i = tid; i < n; i += nthr) {
            const int ry = i / nx, yy = y_lo + ry, xx = x_lo + (i - ry * nx) * 8;
            const uint32_t *sp = (const uint32_t *)(tile + (yy + 8) * PITCH + xx + 8);
            uint2 v; v.x = sp[0]; v.y = sp[1];
            *(uint2 *)(g + (ptrdiff_t)yy * gstride + xx) = v;
        }
    } else {
        const int nx = x_hi - x_lo, n = nx * (y_hi - y_lo);
        for (int i = tid; i < n; i += nthr) {
            const int ry = i / nx, yy = y_lo + ry, xx = x_lo + (i - ry * nx);
            g[(ptrdiff_t)yy * gstride + xx] = tile[(yy + 8) * PITCH + xx + 8];
        }
    }
}

struct LfTile { uint8_t *y, *u, *v; };

/* geometry of super-block c of row r */
struct LfSbGeom { int mi_col, cols, lw, cw; };
__device__ __forceinline__ LfSbGeom lf_geom(const LfKernelArgs &A, int c) {
    LfSbGeom g;
    g.mi_col = c * 8;
    g.cols = A.P.mi_cols - g.mi_col < 8 ? A.P.mi_cols - g.mi_col : 8;
    g.lw = g.cols * 8; g.cw = g.cols * 4;
    return g;
}

__global__ void __launch_bounds__(LF_THREADS) lf_row_kernel(const __grid_constant__ LfKernelArgs A) {
    LF_SHARED16 uint8_t ty[2][72 * LF_YP];
    LF_SHARED16 uint8_t tu[2][40 * LF_CP];
    LF_SHARED16 uint8_t tv[2][40 * LF_CP];
    LF_DYN_SMEM(dyn); /* sbs_w adjusted masks (160 B each), then sbs_w x 16 chroma filter levels */
    SvtB200LfMask *msk = (SvtB200LfMask *)dyn;
    uint8_t (*lfl_uv)[16] = (uint8_t (*)[16])(dyn + (size_t)A.sbs_w * sizeof(SvtB200LfMask));
    __shared__ SvtB200LfThresh thr[64];
    __shared__ int row_s, seen_s;
    const int t = threadIdx.x;
    const bool mover = t >= LF_HALF;
    const int mt = t - LF_HALF;
    if (t == 0) row_s = atomicAdd(A.ticket, 1);
    if (t < 64) thr[t] = A.thr[t];
    __syncthreads();
    const int r = row_s;
    if (r >= A.sbs_h) return;
    const int mi_row = r * 8;
    const int rows = A.P.mi_rows - mi_row < 8 ? A.P.mi_rows - mi_row : 8;
    const int y0 = mi_row * 8, cy0 = mi_row * 4, lh = rows * 8, ch = rows * 4, lpy = r > 0 ? 8 : 0;
    const bool vy = (((uintptr_t)A.y | (uintptr_t)A.y_stride) & 7) == 0;
    const bool vc0 = !A.P.y_only && ((((uintptr_t)A.u | (uintptr_t)A.v | (uintptr_t)A.uv_stride) & 7) == 0);
    const bool chroma = !A.P.y_only;
    const int32_t *above = r > 0 ? A.done + (size_t)(r - 1) * LF_FLAG_PAD : nullptr;
    int32_t *mine = A.done + (size_t)r * LF_FLAG_PAD;

    /* the row above must have finished 'need' super-blocks (their right 8 columns included).  One thread polls: it reads the
     * progress word before its own stores of the step (a fence behind outstanding stores is expensive), keeps the count it saw
     * and does not look again while that covers the request; the other movers are ordered behind it by their barrier. */
    int seen = above ? 0 : 1 << 20;
    auto wait_above = [&](int need) {
        if (need > A.sbs_w) need = A.sbs_w;
        if (seen >= need) return;
        for (;;) {
            const int f = lf_flag_load(above);
            if ((f >> 10) == A.gen && (f & 1023) >= need) { seen = f & 1023; break; }
            lf_sleep();
        }
        lf_fence();
    };
    /* the own samples of a whole-width tile (any height) as two steps: the loads are issued half a step before the buffer they go to is free */
    auto own_fast = [&](int c) {
        const LfSbGeom g = lf_geom(A, c);
        return g.cols == 8 && c > 0 && vy && (!chroma || vc0);
    };
    auto own_issue = [&](int c, uint2 (&v)[6], int tid) {
        const uint8_t *gy = A.y + (size_t)y0 * A.y_stride + c * 64;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int i = tid + 128 * k;
            if ((i >> 3) < lh) v[k] = lf_ldcg8(gy + (ptrdiff_t)(i >> 3) * A.y_stride + (i & 7) * 8);
        }
        if (chroma && (tid >> 2) < ch) {
            const uint8_t *gu = A.u + (size_t)cy0 * A.uv_stride + c * 32, *gv = A.v + (size_t)cy0 * A.uv_stride + c * 32;
            v[4] = lf_ldcg8(gu + (ptrdiff_t)(tid >> 2) * A.uv_stride + (tid & 3) * 8);
            v[5] = lf_ldcg8(gv + (ptrdiff_t)(tid >> 2) * A.uv_stride + (tid & 3) * 8);
        }
    };
    auto own_commit = [&](int b, const uint2 (&v)[6], int tid) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int i = tid + 128 * k;
            if ((i >> 3) >= lh) continue;
            uint32_t *d = (uint32_t *)(ty[b] + ((i >> 3) + 8) * LF_YP + (i & 7) * 8 + 8);
            d[0] = v[k].x; d[1] = v[k].y;
        }
        if (chroma && (tid >> 2) < ch) {
            uint32_t *d = (uint32_t *)(tu[b] + ((tid >> 2) + 8) * LF_CP + (tid & 3) * 8 + 8);
            d[0] = v[4].x; d[1] = v[4].y;
            d = (uint32_t *)(tv[b] + ((tid >> 2) + 8) * LF_CP + (tid & 3) * 8 + 8);
            d[0] = v[5].x; d[1] = v[5].y;
        }
    };
    /* the 8 rows above a whole-width tile with the top-left corner (none for the first tile), as two steps like the own
     * samples: luma 8 rows x 9 words of 8 bytes, chroma 2 planes x 8 rows x 5 */
    auto strip_fast = [&](int c) {
        const LfSbGeom g = lf_geom(A, c);
        return lpy && g.cols == 8 && vy && (!chroma || vc0);
    };
    auto strip_issue = [&](int c, uint2 &a, uint2 &q, int tid) {
        const int ty_y = tid / 9, ty_x = tid - ty_y * 9;
        const int tc_p = tid >> 6, tc_i = tid & 63, tc_y = tc_i / 5, tc_x = tc_i - tc_y * 5;
        a.x = a.y = q.x = q.y = 0;
        if (tid < 72 && (ty_x || c)) a = lf_ldcg8(A.y + (size_t)y0 * A.y_stride + c * 64 + (ptrdiff_t)(ty_y - 8) * A.y_stride + (ty_x - 1) * 8);
        if (chroma && tc_i < 40 && (tc_x || c))
            q = lf_ldcg8((tc_p ? A.v : A.u) + (size_t)cy0 * A.uv_stride + c * 32 + (ptrdiff_t)(tc_y - 8) * A.uv_stride + (tc_x - 1) * 8);
    };
    auto strip_commit = [&](int c, int b, const uint2 &a, const uint2 &q, int tid) {
        const int ty_y = tid / 9, ty_x = tid - ty_y * 9;
        const int tc_p = tid >> 6, tc_i = tid & 63, tc_y = tc_i / 5, tc_x = tc_i - tc_y * 5;
        if (tid < 72 && (ty_x || c)) { uint32_t *d = (uint32_t *)(ty[b] + ty_y * LF_YP + ty_x * 8); d[0] = a.x; d[1] = a.y; }
        if (chroma && tc_i < 40 && (tc_x || c)) { uint32_t *d = (uint32_t *)((tc_p ? tv[b] : tu[b]) + tc_y * LF_CP + tc_x * 8); d[0] = q.x; d[1] = q.y; }
    };
    /* fetch into buffer b the samples of tile c nobody has modified yet (its own), or (strip) the 8 rows above it with the
     * top-left corner, final once the row above has published c + 1; the strip left of the tile comes from the previous
     * tile through shared memory */
    auto fetch = [&](int c, int b, bool strip, int tid, int nthr) {
        const LfSbGeom g = lf_geom(A, c);
        const int lpx = c > 0 ? 8 : 0;
        const bool vc = vc0 && (g.cw & 7) == 0;
        uint8_t *gy = A.y + (size_t)y0 * A.y_stride + c * 64;
        uint8_t *gu = chroma ? A.u + (size_t)cy0 * A.uv_stride + c * 32 : nullptr, *gv = chroma ? A.v + (size_t)cy0 * A.uv_stride + c * 32 : nullptr;
        if (nthr == LF_HALF && g.cols == 8 && (lpx || strip) && vy && (vc || !chroma)) {
            /* whole interior tile, 128 threads: independent 8-byte loads issued together, then stored */
            if (!strip) {
                uint2 v[6];
                own_issue(c, v, tid);
                own_commit(b, v, tid);
            } else if (lpy) {
                uint2 a, q;
                strip_issue(c, a, q, tid);
                strip_commit(c, b, a, q, tid);
            }
            return;
        }
        if (!strip) {
            lf_tile_load<LF_YP>(ty[b], gy, A.y_stride, 0, g.lw, 0, lh, vy, tid, nthr);
            if (chroma) {
                lf_tile_load<LF_CP>(tu[b], gu, A.uv_stride, 0, g.cw, 0, ch, vc, tid, nthr);
                lf_tile_load<LF_CP>(tv[b], gv, A.uv_stride, 0, g.cw, 0, ch, vc, tid, nthr);
            }
        } else {
            lf_tile_load<LF_YP>(ty[b], gy, A.y_stride, -lpx, g.lw, -lpy, 0, vy, tid, nthr);
            if (chroma) {
                lf_tile_load<LF_CP>(tu[b], gu, A.uv_stride, -lpx, g.cw, -lpy, 0, vc, tid, nthr);
                lf_tile_load<LF_CP>(tv[b], gv, A.uv_stride, -lpx, g.cw, -lpy, 0, vc, tid, nthr);
            }
        }
    };
    /* write back a whole-width tile c-1 but for its right 8 columns, and the strip above it (the common case with
     * compile-time extents); or arbitrary extents */
    auto put = [&](int c, int b, int x_lo, int x_hi, int y_lo, int y_hi, int cx_lo, int cx_hi, int cy_lo, int cy_hi, int tid, int nthr) {
        const LfSbGeom g = lf_geom(A, c);
        const bool vc = vc0 && (g.cw & 7) == 0;
        uint8_t *gy = A.y + (size_t)y0 * A.y_stride + c * 64;
        lf_tile_store<LF_YP>(ty[b], gy, A.y_stride, x_lo, x_hi, y_lo, y_hi, vy, tid, nthr);
        if (chroma) {
            uint8_t *gu = A.u + (size_t)cy0 * A.uv_stride + c * 32, *gv = A.v + (size_t)cy0 * A.uv_stride + c * 32;
            lf_tile_store<LF_CP>(tu[b], gu, A.uv_stride, cx_lo, cx_hi, cy_lo, cy_hi, vc, tid, nthr);
            lf_tile_store<LF_CP>(tv[b], gv, A.uv_stride, cx_lo, cx_hi, cy_lo, cy_hi, vc, tid, nthr);
        }
    };
    auto put_body = [&](int c, int b, int tid, int nthr) { /* tile c is 64 luma samples wide */
        uint8_t *gy = A.y + (size_t)y0 * A.y_stride + c * 64;
        lf_tile_store<LF_YP, 7>(ty[b], gy, A.y_stride, 0, 56, 0, lh, vy, tid, nthr);
        lf_tile_store<LF_YP, 8>(ty[b], gy, A.y_stride, 0, 64, -lpy, 0, vy, tid, nthr);
        if (chroma) {
            uint8_t *gu = A.u + (size_t)cy0 * A.uv_stride + c * 32, *gv = A.v + (size_t)cy0 * A.uv_stride + c * 32;
            lf_tile_store<LF_CP, 3>(tu[b], gu, A.uv_stride, 0, 24, 0, ch, vc0, tid, nthr);
            lf_tile_store<LF_CP, 4>(tu[b], gu, A.uv_stride, 0, 32, -lpy, 0, vc0, tid, nthr);
            lf_tile_store<LF_CP, 3>(tv[b], gv, A.uv_stride, 0, 24, 0, ch, vc0, tid, nthr);
            lf_tile_store<LF_CP, 4>(tv[b], gv, A.uv_stride, 0, 32, -lpy, 0, vc0, tid, nthr);
        }
    };
    auto put_left = [&](int c, int b, int tid, int nthr) { /* the left context of tile c = the right 8 columns of tile c-1 */
        uint8_t *gy = A.y + (size_t)y0 * A.y_stride + c * 64;
        lf_tile_store<LF_YP, 1>(ty[b], gy, A.y_stride, -8, 0, 0, lh, vy, tid, nthr);
        if (chroma) {
            uint8_t *gu = A.u + (size_t)cy0 * A.uv_stride + c * 32, *gv = A.v + (size_t)cy0 * A.uv_stride + c * 32;
            lf_tile_store<LF_CP, 1>(tu[b], gu, A.uv_stride, -8, 0, 0, ch, vc0, tid, nthr);
            lf_tile_store<LF_CP, 1>(tv[b], gv, A.uv_stride, -8, 0, 0, ch, vc0, tid, nthr);
        }
    };

    /* ---- prologue: tile 0 */
    if (t == 0) seen_s = above ? 0 : 1 << 20;
    for (int c = t; c < A.sbs_w; c += LF_THREADS) { /* every mask of the row, border-adjusted, once */
        msk[c] = A.lfm[c + r * A.lfm_stride];
        lf_adjust_mask(A.P.mi_rows, A.P.mi_cols, mi_row, c * 8, &msk[c]);
        for (int rr = 0; rr < 4; rr++) for (int cc = 0; cc < 4; cc++) lfl_uv[c][rr * 4 + cc] = msk[c].lfl_y[(rr * 2) * 8 + cc * 2];
    }
    __syncthreads();
    fetch(0, 0, false, t, LF_THREADS);

    for (int c = 0; c < A.sbs_w; c++) {
        const int cur = c & 1, oth = cur ^ 1;
        const LfSbGeom G = lf_geom(A, c);
        __syncthreads(); /* tile c (but for the strip above it) is in buffer cur; tile c-1 (filtered) is in buffer oth */
        unsigned long long *tr = A.trace ? A.trace + ((size_t)r * A.sbs_w + c) * 8 : nullptr;
        if (tr && t == 0) tr[0] = lf_clock();
        LfPlaneMasks MY, MC;
        uint2 pre[6];
        bool pre_valid = false;
        if (!mover) {
            const SvtB200LfMask &m = msk[c];
            MY.l16 = m.left_y[2]; MY.l8 = m.left_y[1]; MY.l4 = m.left_y[0]; MY.a16 = m.above_y[2]; MY.a8 = m.above_y[1]; MY.a4 = m.above_y[0];
            MY.i4v = m.int_4x4_y; MY.i4h = m.int_4x4_y; MY.lfl = m.lfl_y; MY.bw = 8; MY.rows = rows; MY.px_rows = lh; MY.top = mi_row == 0;
            MC.l16 = m.left_uv[2]; MC.l8 = m.left_uv[1]; MC.l4 = m.left_uv[0]; MC.a16 = m.above_uv[2]; MC.a8 = m.above_uv[1]; MC.a4 = m.above_uv[0];
            MC.i4v = m.int_4x4_uv; MC.i4h = m.int_4x4_uv;
            /* the internal 4x4 edge of the picture's last (half) chroma block row is not filtered horizontally */
            for (int rr = 0; rr < 8; rr += 2) if (mi_row + rr == A.P.mi_rows - 1) MC.i4h &= ~((uint64_t)0xf << ((rr >> 1) * 4));
            MC.lfl = lfl_uv[c]; MC.bw = 4; MC.rows = (rows + 1) >> 1; MC.px_rows = ch; MC.top = mi_row == 0;
            /* ---- vertical edges: one thread per sample row; the thread first pulls its row of the left context
             * (the previous tile's right 8 columns, final but for this SB's left edge) out of the other buffer */
            if (t < 64) {
                if (c > 0 && t < lh) {
                    const uint32_t *s = (const uint32_t *)(ty[oth] + (t + 8) * LF_YP + 8 + 56);
                    uint32_t *d = (uint32_t *)(ty[cur] + (t + 8) * LF_YP);
                    d[0] = s[0]; d[1] = s[1];
                }
                lf_vertical_row(ty[cur] + (t + 8) * LF_YP + 8, t, MY, thr);
            } else if (chroma) {
                const int q = (t - 64) & 31;
                uint8_t *tc = t < 96 ? tu[cur] : tv[cur];
                if (c > 0 && q < ch) {
                    const uint32_t *s = (const uint32_t *)((t < 96 ? tu[oth] : tv[oth]) + (q + 8) * LF_CP + 8 + 24);
                    uint32_t *d = (uint32_t *)(tc + (q + 8) * LF_CP);
                    d[0] = s[0]; d[1] = s[1];
                }
                lf_vertical_row(tc + (q + 8) * LF_CP + 8, q, MC, thr);
            }
        } else {
            /* ---- movers, first half of the step: issue the loads of tile c+1's own samples and - once the row above is known
             * to have published SB c - of the strip above tile c (only the horizontal edges need it); both land in registers
             * while tile c-1 is written back but for its right 8 columns (still the left context of tile c) */
            if (c + 1 < A.sbs_w && own_fast(c + 1)) { own_issue(c + 1, pre, mt); pre_valid = true; }
            const bool sf = strip_fast(c);
            const int need = c + 1 < A.sbs_w ? c + 1 : A.sbs_w;
            uint2 sa, sq;
            if (tr && mt == 0) tr[4] = lf_clock();
            if (seen_s < need) { /* not known yet from the look at the end of the previous step: wait for it */
                if (mt == 0) wait_above(need);
                lf_sync_movers();
            }
            if (tr && mt == 0) tr[5] = lf_clock();
            if (sf) strip_issue(c, sa, sq, mt);
            if (c > 0) put_body(c - 1, oth, mt, LF_HALF);
            if (sf) strip_commit(c, cur, sa, sq, mt);
            else fetch(c, cur, true, mt, LF_HALF);
            if (tr && mt == 0) tr[6] = lf_clock();
        }
        if (tr && t == 0) tr[1] = lf_clock();
        __syncthreads(); /* vertical edges done; buffer oth is free */
        if (tr && t == 0) tr[2] = lf_clock();
        if (!mover) {
            /* ---- horizontal edges: one thread per sample column */
            if (t < 64) { if (t < G.lw) lf_horizontal_col(ty[cur] + 8 * LF_YP + 8 + t, t, LF_YP, MY, thr); }
            else if (chroma) { const int q = (t - 64) & 31; if (q < G.cw) lf_horizontal_col((t < 96 ? tu[cur] : tv[cur]) + 8 * LF_CP + 8 + q, q, LF_CP, MC, thr); }
            if (tr && t == 0) tr[3] = lf_clock();
        } else {
            /* ---- movers, second half: the left context of tile c is final as soon as the vertical pass is over (horizontal
             * edges do not reach it): with it tile c-1 is complete in memory -> fence, publish at once; then tile c+1's own
             * samples go from the registers into the freed buffer */
            if (c > 0) put_left(c, cur, mt, LF_HALF);
            lf_sync_movers(); /* every mover's stores of SBs 0..c-1 are ordered before the one fence below */
            if (c + 1 < A.sbs_w) {
                if (pre_valid) own_commit(oth, pre, mt);
                else fetch(c + 1, oth, false, mt, LF_HALF);
            }
            if (mt == 0) {
                /* one fence serves both directions: behind the look at the row above (does it already cover the next step?
                 * then that step fetches its strip at once), and in front of this row's own progress word */
                const int f = above ? lf_flag_load(above) : 0;
                if (above || c > 0) lf_fence();
                if (c > 0) lf_flag_store(mine, (A.gen << 10) | c); /* SBs 0..c-1 are final in memory */
                if (above && (f >> 10) == A.gen && (f & 1023) > seen) seen = f & 1023;
                seen_s = seen;
            }
            if (tr && mt == 0) tr[7] = lf_clock();
        }
    }
    __syncthreads();
    /* ---- epilogue: the last tile, right columns included */
    {
        const LfSbGeom g = lf_geom(A, A.sbs_w - 1);
        put(A.sbs_w - 1, (A.sbs_w - 1) & 1, 0, g.lw, 0, lh, 0, g.cw, 0, ch, t, LF_THREADS);
        put(A.sbs_w - 1, (A.sbs_w - 1) & 1, 0, g.lw, -lpy, 0, 0, g.cw, -lpy, 0, t, LF_THREADS);
    }
    lf_fence();
    __syncthreads();
    if (t == 0) lf_flag_store(mine, (A.gen << 10) | A.sbs_w);
}
#endif
