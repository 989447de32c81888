/* intra_device.cuh -- the VP9 intra predictors (Source/Lib/VPX/intrapred.c:22-218, 4x4 forms :253-417) in closed form:
 * sample (r, c) of a bs x bs block straight from the edge samples, so that every thread can produce any sample without the
 * row-to-row / column-to-column copies the sequential forms use.  e(j) walks the edge as one line: j >= 0 the above row
 * (running on into above-right), j = -1 the corner above[-1], j <= -2 the left column downwards (left[-2 - j]).
 * Included by pred.cu; tests/emu/pred_emu.cpp runs the same source on the CPU. */
#ifndef SVTB200_INTRA_DEVICE_CUH
#define SVTB200_INTRA_DEVICE_CUH
#include <stdint.h>
#include <stddef.h>

struct IntraEdge { const uint8_t *a, *l; };
__device__ __forceinline__ int intra_e(const IntraEdge &E, int j) { return j >= -1 ? E.a[j] : E.l[-2 - j]; }
__device__ __forceinline__ int intra_a2(int a, int b) { return (a + b + 1) >> 1; }
__device__ __forceinline__ int intra_a3(int a, int b, int c) { return (a + 2 * b + c + 2) >> 2; }

/* the value every sample of a DC_PRED block takes: dc_128 / dc_top / dc_left / dc by availability */
__device__ __forceinline__ int intra_dc_value(int bs, int have_left, int have_top, const IntraEdge &E) {
    int sum = 0;
    if (have_left) for (int i = 0; i < bs; i++) sum += E.l[i];
    if (have_top) for (int i = 0; i < bs; i++) sum += E.a[i];
    if (have_left && have_top) return (sum + bs) / (2 * bs);
    if (have_left || have_top) return (sum + (bs >> 1)) / bs;
    return 128;
}

__device__ __forceinline__ int intra_sample(int mode, int bs, int dc, const IntraEdge &E, int r, int c) {
    switch (mode) {
    case 0: return dc;
    case 1: return E.a[c];                                             /* V */
    case 2: return E.l[r];                                             /* H */
    case 9: { const int v = E.l[r] + E.a[c] - E.a[-1]; return v < 0 ? 0 : v > 255 ? 255 : v; } /* TM */
    case 3: {                                                          /* D45: anti-diagonals of the smoothed above row */
        const int k = r + c;
        if (bs == 4) return k < 6 ? intra_a3(E.a[k], E.a[k + 1], E.a[k + 2]) : E.a[7];
        return k < bs - 1 ? intra_a3(E.a[k], E.a[k + 1], E.a[k + 2]) : E.a[bs - 1];
    }
    case 8: {                                                          /* D63: two rows per step along the above row */
        const int h = r >> 1, k = h + c;
        if (bs != 4 && r >= 2 && c >= bs - 1 - h) return E.a[bs - 1];
        return (r & 1) ? intra_a3(E.a[k], E.a[k + 1], E.a[k + 2]) : intra_a2(E.a[k], E.a[k + 1]);
    }
    case 7: {                                                          /* D207: two columns per step down the left column, clamped at its end */
        const int k = r + (c >> 1), m = bs - 1;
        const int l0 = E.l[k < m ? k : m], l1 = E.l[k + 1 < m ? k + 1 : m], l2 = E.l[k + 2 < m ? k + 2 : m];
        return (c & 1) ? intra_a3(l0, l1, l2) : intra_a2(l0, l1);
    }
    case 5: {                                                          /* D117: (r, c) = (r - 2, c - 1) until the first two rows or the first column */
        const int s = (r >> 1) < c ? (r >> 1) : c, rr = r - 2 * s, cc = c - s;
        if (rr == 0) return intra_a2(intra_e(E, cc - 1), intra_e(E, cc));
        if (rr == 1) return intra_a3(intra_e(E, cc - 2), intra_e(E, cc - 1), intra_e(E, cc));
        return intra_a3(intra_e(E, 1 - rr), intra_e(E, -rr), intra_e(E, -1 - rr));
    }
    case 4: return intra_a3(intra_e(E, c - r - 2), intra_e(E, c - r - 1), intra_e(E, c - r)); /* D135: the smoothed edge line along the main diagonal */
    case 6: {                                                          /* D153: (r, c) = (r - 1, c - 2) until the first row or the first two columns */
        const int s = r < (c >> 1) ? r : (c >> 1), rr = r - s, cc = c - 2 * s;
        if (cc == 0) return intra_a2(intra_e(E, -1 - rr), intra_e(E, -2 - rr));
        if (cc == 1) return intra_a3(intra_e(E, -rr), intra_e(E, -1 - rr), intra_e(E, -2 - rr));
        return intra_a3(intra_e(E, cc - 3), intra_e(E, cc - 2), intra_e(E, cc - 1));
    }
    }
    return 0;
}
/* ---- generate_intra_reference_samples (Source/Lib/Codec/EbEncDecProcess.c:1128-1315) in closed form: byte k of a job's 128-byte edge
 * record straight from the plane (one thread per byte: every output sample is ONE plane sample or a constant, picked by the
 * availability / frame-edge rules; the sequential form's copy-then-replicate steps become an index clamp). ---- */
#ifdef SVTB200_CPU_EMU
#include "../../include/svtb200.h"
#endif
__device__ __forceinline__ uint8_t intra_edge_byte(const uint8_t *plane, int stride, int fw, int fh, const SvtB200IntraEdgeJob &b, int k) {
    const int bs = 4 << b.tx_size, x0 = b.x0, y0 = b.y0;
    if (k >= 80) { /* left column */
        const int i = k - 80;
        if (i >= bs) return 0;
        if (!b.have_left) return 129;
        const int s = (b.extend_bottom && y0 + bs > fh) ? min(i, fh - y0 - 1) : i; /* rows past the frame: the last row inside */
        return plane[(ptrdiff_t)(y0 + s) * stride + x0 - 1];
    }
    if (k < 15 || k >= 16 + 2 * bs) return 0;
    if (!b.have_top) return 127;
    const uint8_t *A = plane + (ptrdiff_t)(y0 - 1) * stride + x0;
    if (k == 15) return b.have_left ? A[-1] : 129;
    const int i = k - 16;
    const bool four_right = bs == 4 && b.have_right; /* the left 4x4 of an 8x8 block: its above-right samples exist */
    int last = four_right ? 2 * bs - 1 : bs - 1;       /* last sample taken from the plane; the rest replicates it */
    if (b.extend_right) {
        const int r = fw - x0;                         /* samples of the row inside the frame */
        if (x0 + 2 * bs <= fw) {}
        else if (x0 + bs <= fw) { if (four_right) last = r - 1; }
        else last = r - 1;
    }
    return A[min(i, last)];
}
#define INTRA_EDGE_THREADS 256
__global__ void __launch_bounds__(INTRA_EDGE_THREADS) intra_edges_kernel(const uint8_t *plane, int stride, int fw, int fh, const SvtB200IntraEdgeJob *jobs, int n,
                                                                         uint8_t *out) {
    const int j = blockIdx.x * 2 + (threadIdx.x >> 7), k = threadIdx.x & 127; /* two jobs per CTA, one thread per byte of the record */
    if (j >= n) return;
    out[(size_t)j * 128 + k] = intra_edge_byte(plane, stride, fw, fh, jobs[j], k);
}
#endif
