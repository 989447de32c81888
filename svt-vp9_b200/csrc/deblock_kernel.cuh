/*
 * deblock_kernel.cuh -- S2: in-loop deblocking of one picture, replacing eb_vp9_loop_filter_frame()
 * (Source/Lib/VPX/vp9_loopfilter.c:1521 -> loop_filter_rows :1456 -> eb_vp9_adjust_mask :786,
 * eb_vp9_filter_block_plane_ss00/_ss11 :1238/:1342, eb_vp9_lpf_* Source/Lib/VPX/loopfilter.c:103-329).
 *
 * VP9 fixes the filtering order: super-blocks in raster order, and inside each the vertical edges of a plane
 * before its horizontal edges.  Filters reach 8 samples across super-block borders, so SB (r,c) depends on
 * (r,c-1) and (r-1,c+1): a wavefront.  One CTA per SB takes its SB index from an atomic ticket (so a CTA only
 * ever waits for CTAs that already hold a ticket, i.e. are running or done -- no deadlock).  The SB's own samples
 * are untouched until its turn, so they are fetched into the shared-memory tiles (72x72 luma, two 40x40 chroma:
 * SB + 8 samples of left/top context) BEFORE the CTA polls its two dependency flags; after the flags only the
 * 8-sample left and top strips remain to be read (8-byte L1-bypassing loads).  All three planes are filtered
 * concurrently (one thread per sample row for the vertical edges, one per sample column for the horizontal edges,
 * edges sequential along the thread's line); the tile is written back with 8-byte stores and the flag published
 * after a device-scope fence.
 *
 * The file also compiles as plain C++ under csrc/emu/cuda_emu.h for the CPU-side logic tests.
 */
#ifndef SVTB200_DEBLOCK_KERNEL_CUH
#define SVTB200_DEBLOCK_KERNEL_CUH
#include <stdint.h>
#include "../../include/svtb200.h"

#define LF_THREADS 128
#define LF_YP 76 /* luma tile pitch (odd number of words: conflict-free column access) */
#define LF_CP 44 /* chroma tile pitch */

struct LfKernelArgs {
    SvtB200LfParams P;
    uint8_t *y, *u, *v;
    int32_t  y_stride, uv_stride;
    const SvtB200LfMask *lfm;
    int32_t  lfm_stride;
    const SvtB200LfThresh *thr; /* [64] */
    int32_t *ticket;            /* zeroed before the launch */
    int32_t *done;              /* one flag per SB; a SB is finished when its flag == gen */
    int32_t  gen;
    int32_t  sbs_w, sbs_h;
};

#ifdef SVTB200_CPU_EMU
#define LF_SHARED16 alignas(16) static
static inline uint8_t lf_ldcg(const uint8_t *p) { return *p; }
static inline uint2 lf_ldcg8(const uint8_t *p) { uint2 v; memcpy(&v, p, 8); return v; }
static inline void lf_fence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void lf_sleep() {}
static inline int lf_flag_load(const int32_t *p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
static inline void lf_flag_store(int32_t *p, int v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
#else
#define LF_SHARED16 __shared__ __align__(16)
__device__ __forceinline__ uint8_t lf_ldcg(const uint8_t *p) { return __ldcg(p); }
__device__ __forceinline__ uint2 lf_ldcg8(const uint8_t *p) { return __ldcg((const uint2 *)p); }
__device__ __forceinline__ void lf_fence() { __threadfence(); }
__device__ __forceinline__ void lf_sleep() { __nanosleep(20); }
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

/* all vertical edges crossing sample row 'row' (0..8*bw-1) of the tile; t points at tile sample (0,row) of the SB */
__device__ static void lf_vertical_row(uint8_t *t, int row, const LfPlaneMasks &M, const SvtB200LfThresh *thr) {
    if (row >= M.px_rows) return;
    const int r = row >> 3, pr = r & ~1, other = r ^ 1;
    for (int c = 0; c < M.bw; c++) {
        const int bit = r * M.bw + c;
        uint8_t *s = t + c * 8;
        const SvtB200LfThresh own = thr[M.lfl[bit]];
        if (M.l16 >> bit & 1) {
            /* a 16-wide edge present in both rows of a row pair takes the thresholds of the pair's first row */
            const SvtB200LfThresh th = (M.l16 >> (other * M.bw + c) & 1) ? thr[M.lfl[pr * M.bw + c]] : own;
            lf_filter_across(s, 1, 16, th.mblim, th.lim, th.hev_thr);
        }
        if (M.l8 >> bit & 1) lf_filter_across(s, 1, 8, own.mblim, own.lim, own.hev_thr);
        if (M.l4 >> bit & 1) lf_filter_across(s, 1, 4, own.mblim, own.lim, own.hev_thr);
        if (M.i4v >> bit & 1) lf_filter_across(s + 4, 1, 4, own.mblim, own.lim, own.hev_thr);
    }
}

/* all horizontal edges crossing sample column 'col'; t points at tile sample (col,0) of the SB */
__device__ static void lf_horizontal_col(uint8_t *t, int col, int pitch, const LfPlaneMasks &M, const SvtB200LfThresh *thr) {
    const int c = col >> 3;
    for (int r = 0; r < M.rows; r++) {
        const int bit = r * M.bw + c;
        uint8_t *s = t + (r * 8) * pitch;
        const bool top = M.top && r == 0;
        const SvtB200LfThresh own = thr[M.lfl[bit]];
        if (!top && (M.a16 >> bit & 1)) {
            /* adjacent 16-wide edges are paired from the start of each run; a pair shares the first block's thresholds */
            int start = c;
            while (start > 0 && (M.a16 >> (r * M.bw + start - 1) & 1)) start--;
            const SvtB200LfThresh th = thr[M.lfl[r * M.bw + c - ((c - start) & 1)]];
            lf_filter_across(s, pitch, 16, th.mblim, th.lim, th.hev_thr);
        } else {
            if (!top && (M.a8 >> bit & 1)) lf_filter_across(s, pitch, 8, own.mblim, own.lim, own.hev_thr);
            else if (!top && (M.a4 >> bit & 1)) lf_filter_across(s, pitch, 4, own.mblim, own.lim, own.hev_thr);
            if ((M.i4h >> bit & 1) && r * 8 + 4 < M.px_rows) lf_filter_across(s + 4 * pitch, pitch, 4, own.mblim, own.lim, own.hev_thr);
        }
    }
}

/* Copy the samples [x_lo, x_hi) x [y_lo, y_hi) (SB-relative; the tile keeps SB sample (0,0) at (8,8)) between the
 * picture plane and the shared-memory tile.  vec: x bounds, plane pointer and stride are multiples of 8 -> 8-byte
 * global accesses (the tile pitch is only 4-byte aligned: two words on the shared side). */
template <int PITCH>
__device__ __forceinline__ void lf_tile_load(uint8_t *tile, const uint8_t *g, int gstride, int x_lo, int x_hi, int y_lo, int y_hi, bool vec) {
    if (x_hi <= x_lo || y_hi <= y_lo) return;
    if (vec) {
        const int nx = (x_hi - x_lo) >> 3, n = nx * (y_hi - y_lo);
        for (int i = threadIdx.x; i < n; i += LF_THREADS) {
            const int ry = i / nx, yy = y_lo + ry, xx = x_lo + (i - ry * nx) * 8;
            const uint2 v = lf_ldcg8(g + (ptrdiff_t)yy * gstride + xx);
            uint32_t *d = (uint32_t *)(tile + (yy + 8) * PITCH + xx + 8);
            d[0] = v.x; d[1] = v.y;
        }
    } else {
        const int nx = x_hi - x_lo, n = nx * (y_hi - y_lo);
        for (int i = threadIdx.x; i < n; i += LF_THREADS) {
            const int ry = i / nx, yy = y_lo + ry, xx = x_lo + (i - ry * nx);
            tile[(yy + 8) * PITCH + xx + 8] = lf_ldcg(g + (ptrdiff_t)yy * gstride + xx);
        }
    }
}
template <int PITCH>
__device__ __forceinline__ void lf_tile_store(const uint8_t *tile, uint8_t *g, int gstride, int x_lo, int x_hi, int y_lo, int y_hi, bool vec) {
    if (x_hi <= x_lo || y_hi <= y_lo) return;
    if (vec) {
        const int nx = (x_hi - x_lo) >> 3, n = nx * (y_hi - y_lo);
        for (int i = threadIdx.x; i < n; i += LF_THREADS) {
            const int ry = i / nx, yy = y_lo + ry, xx = x_lo + (i - ry * nx) * 8;
            const uint32_t *sp = (const uint32_t *)(tile + (yy + 8) * PITCH + xx + 8);
            uint2 v; v.x = sp[0]; v.y = sp[1];
            *(uint2 *)(g + (ptrdiff_t)yy * gstride + xx) = v;
        }
    } else {
        const int nx = x_hi - x_lo, n = nx * (y_hi - y_lo);
        for (int i = threadIdx.x; i < n; i += LF_THREADS) {
            const int ry = i / nx, yy = y_lo + ry, xx = x_lo + (i - ry * nx);
            g[(ptrdiff_t)yy * gstride + xx] = tile[(yy + 8) * PITCH + xx + 8];
        }
    }
}

__global__ void __launch_bounds__(LF_THREADS) lf_sb_kernel(const __grid_constant__ LfKernelArgs A) {
    LF_SHARED16 uint8_t ty[72 * LF_YP];
    LF_SHARED16 uint8_t tu[40 * LF_CP];
    LF_SHARED16 uint8_t tv[40 * LF_CP];
    __shared__ SvtB200LfMask m;
    __shared__ SvtB200LfThresh thr[64];
    __shared__ uint8_t lfl_uv[16];
    __shared__ int sb_s;
    const int t = threadIdx.x;
    if (t == 0) sb_s = atomicAdd(A.ticket, 1);
    if (t < 64) thr[t] = A.thr[t];
    __syncthreads();
    const int sb = sb_s;
    if (sb >= A.sbs_w * A.sbs_h) return;
    const int r = sb / A.sbs_w, c = sb - r * A.sbs_w, mi_row = r * 8, mi_col = c * 8;
    const int rows = A.P.mi_rows - mi_row < 8 ? A.P.mi_rows - mi_row : 8;
    const int cols = A.P.mi_cols - mi_col < 8 ? A.P.mi_cols - mi_col : 8;
    const int x0 = mi_col * 8, y0 = mi_row * 8, cx0 = mi_col * 4, cy0 = mi_row * 4;
    const int lw = cols * 8, lh = rows * 8, cw = cols * 4, ch = rows * 4;
    const int lpx = x0 >= 8 ? 8 : 0, lpy = y0 >= 8 ? 8 : 0; /* context available left / above */
    uint8_t *gy = A.y + (size_t)y0 * A.y_stride + x0, *gu = A.u + (size_t)cy0 * A.uv_stride + cx0, *gv = A.v + (size_t)cy0 * A.uv_stride + cx0;
    const bool vy = (((uintptr_t)A.y | (uintptr_t)A.y_stride) & 7) == 0;
    const bool vc = !A.P.y_only && ((((uintptr_t)A.u | (uintptr_t)A.v | (uintptr_t)A.uv_stride) & 7) == 0) && (cw & 7) == 0;
    /* ---- the SB's own samples: nobody has modified them yet, fetch them while the dependencies finish */
    lf_tile_load<LF_YP>(ty, gy, A.y_stride, 0, lw, 0, lh, vy);
    if (!A.P.y_only) {
        lf_tile_load<LF_CP>(tu, gu, A.uv_stride, 0, cw, 0, ch, vc);
        lf_tile_load<LF_CP>(tv, gv, A.uv_stride, 0, cw, 0, ch, vc);
    }
    if (t == 0) {
        m = A.lfm[c + r * A.lfm_stride];
        lf_adjust_mask(A.P.mi_rows, A.P.mi_cols, mi_row, mi_col, &m);
        for (int rr = 0; rr < 4; rr++) for (int cc = 0; cc < 4; cc++) lfl_uv[rr * 4 + cc] = m.lfl_y[(rr * 2) * 8 + cc * 2];
        /* wavefront: left neighbour and upper-right neighbour (upper neighbour in the last column) */
        if (c > 0) while (lf_flag_load(&A.done[sb - 1]) != A.gen) lf_sleep();
        if (r > 0) { const int dep = (r - 1) * A.sbs_w + (c + 1 < A.sbs_w ? c + 1 : c); while (lf_flag_load(&A.done[dep]) != A.gen) lf_sleep(); }
        lf_fence();
    }
    __syncthreads();
    /* ---- the 8-sample strips the neighbours have just finished: left of the SB, and above it (with the corner) */
    lf_tile_load<LF_YP>(ty, gy, A.y_stride, -lpx, 0, 0, lh, vy);
    lf_tile_load<LF_YP>(ty, gy, A.y_stride, -lpx, lw, -lpy, 0, vy);
    if (!A.P.y_only) {
        lf_tile_load<LF_CP>(tu, gu, A.uv_stride, -lpx, 0, 0, ch, vc);
        lf_tile_load<LF_CP>(tu, gu, A.uv_stride, -lpx, cw, -lpy, 0, vc);
        lf_tile_load<LF_CP>(tv, gv, A.uv_stride, -lpx, 0, 0, ch, vc);
        lf_tile_load<LF_CP>(tv, gv, A.uv_stride, -lpx, cw, -lpy, 0, vc);
    }
    __syncthreads();
    LfPlaneMasks MY, MC;
    MY.l16 = m.left_y[2]; MY.l8 = m.left_y[1]; MY.l4 = m.left_y[0]; MY.a16 = m.above_y[2]; MY.a8 = m.above_y[1]; MY.a4 = m.above_y[0];
    MY.i4v = m.int_4x4_y; MY.i4h = m.int_4x4_y; MY.lfl = m.lfl_y; MY.bw = 8; MY.rows = rows; MY.px_rows = lh; MY.top = mi_row == 0;
    MC.l16 = m.left_uv[2]; MC.l8 = m.left_uv[1]; MC.l4 = m.left_uv[0]; MC.a16 = m.above_uv[2]; MC.a8 = m.above_uv[1]; MC.a4 = m.above_uv[0];
    MC.i4v = m.int_4x4_uv; MC.i4h = m.int_4x4_uv;
    /* the internal 4x4 edge of the picture's last (half) chroma block row is not filtered horizontally */
    for (int rr = 0; rr < 8; rr += 2) if (mi_row + rr == A.P.mi_rows - 1) MC.i4h &= ~((uint64_t)0xf << ((rr >> 1) * 4));
    MC.lfl = lfl_uv; MC.bw = 4; MC.rows = (rows + 1) >> 1; MC.px_rows = ch; MC.top = mi_row == 0;
    /* ---- vertical edges: one thread per sample row */
    if (t < 64) lf_vertical_row(ty + (t + 8) * LF_YP + 8, t, MY, thr);
    else if (!A.P.y_only) { const int q = (t - 64) & 31; lf_vertical_row((t < 96 ? tu : tv) + (q + 8) * LF_CP + 8, q, MC, thr); }
    __syncthreads();
    /* ---- horizontal edges: one thread per sample column */
    if (t < 64) { if (t < lw) lf_horizontal_col(ty + 8 * LF_YP + 8 + t, t, LF_YP, MY, thr); }
    else if (!A.P.y_only) { const int q = (t - 64) & 31; if (q < cw) lf_horizontal_col((t < 96 ? tu : tv) + 8 * LF_CP + 8 + q, q, LF_CP, MC, thr); }
    __syncthreads();
    /* ---- write back everything this SB may have changed (its own samples + the 8-sample left / top strips; the
     * corner is never modified) */
    lf_tile_store<LF_YP>(ty, gy, A.y_stride, -lpx, lw, 0, lh, vy);
    lf_tile_store<LF_YP>(ty, gy, A.y_stride, 0, lw, -lpy, 0, vy);
    if (!A.P.y_only) {
        lf_tile_store<LF_CP>(tu, gu, A.uv_stride, -lpx, cw, 0, ch, vc);
        lf_tile_store<LF_CP>(tu, gu, A.uv_stride, 0, cw, -lpy, 0, vc);
        lf_tile_store<LF_CP>(tv, gv, A.uv_stride, -lpx, cw, 0, ch, vc);
        lf_tile_store<LF_CP>(tv, gv, A.uv_stride, 0, cw, -lpy, 0, vc);
    }
    lf_fence();
    __syncthreads();
    if (t == 0) lf_flag_store(&A.done[sb], A.gen);
}
#endif
