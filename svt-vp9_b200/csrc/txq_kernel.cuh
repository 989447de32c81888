/* txq_kernel.cuh -- device code of txq.cu (residual -> forward transform -> quantiser -> inverse transform + reconstruction, and the stages
 * alone); a header of its own so that tests/emu/txq_emu.cpp can run the same source on the CPU. */
#ifndef SVTB200_TXQ_KERNEL_CUH
#define SVTB200_TXQ_KERNEL_CUH
#include <stdint.h>
#include <stddef.h>
#include "txfm_device.cuh"

#define TXQ_THREADS 128
#ifndef TXQ_MIN_BLOCKS
#define TXQ_MIN_BLOCKS 6 /* resident CTAs per SM the register allocation aims at */
#endif

struct TxqArgs {
    /* mode 0: fused from planes */
    const SvtB200TxJob *jobs;
    SvtB200TxPlanes     pl;
    /* mode 1..3: contiguous blocks */
    const int16_t *in16;   /* residual (mode 1), coeff (mode 2), dqcoeff (mode 3) */
    const uint16_t *eob_in;
    uint8_t *dest;         /* mode 3: n blocks of N*N u8, contiguous */
    /* outputs */
    int16_t *coeff, *qcoeff, *dqcoeff;
    uint16_t *eob;
    /* quantiser */
    const SvtB200QuantRow *qrows;
    const uint16_t *qrow_index; /* modes 2: per-block row index or null */
    const int16_t *iscan[4];    /* per tx_type */
    int32_t n, tx_type, partial32;
};

template <int N> struct TxCfg;
template <> struct TxCfg<4> { static constexpr int TS = 0, SHIFT = 4; };
template <> struct TxCfg<8> { static constexpr int TS = 1, SHIFT = 5; };
template <> struct TxCfg<16> { static constexpr int TS = 2, SHIFT = 6; };
template <> struct TxCfg<32> { static constexpr int TS = 3, SHIFT = 6; };

/* transpose an NxN int tile between the N lanes of a block: lane l writes its vector as row l, reads column l */
template <int N>
__device__ __forceinline__ void tile_transpose(int (&v)[N], int *tile /* N*(N+1) ints, private to the block */, int lane, unsigned gmask) {
#pragma unroll
    for (int j = 0; j < N; j++) tile[j * (N + 1) + lane] = v[j];
    __syncwarp(gmask);
#pragma unroll
    for (int j = 0; j < N; j++) v[j] = tile[lane * (N + 1) + j];
    __syncwarp(gmask);
}

/* forward 2-D transform: in v[] = residual column 'lane'; out v[] = coefficient row 'lane' */
template <int N>
__device__ __forceinline__ void fwd2d(int (&v)[N], int tx_type, int *tile, int lane, unsigned gmask) {
    constexpr int TS = TxCfg<N>::TS;
    int t[N];
    if (TS == 3) {
#pragma unroll
        for (int j = 0; j < N; j++) v[j] *= 4;
        Tx1D<N>::fwd(false, v, t);
#pragma unroll
        for (int j = 0; j < N; j++) v[j] = (t[j] + 1 + (t[j] > 0)) >> 2;
        tile_transpose<N>(v, tile, lane, gmask);
        Tx1D<N>::fwd(false, v, t);
#pragma unroll
        for (int j = 0; j < N; j++) v[j] = w16((t[j] + 1 + (t[j] < 0)) >> 2);
        return;
    }
    const bool col_adst = tx_type & 1, row_adst = tx_type & 2;
#pragma unroll
    for (int j = 0; j < N; j++) v[j] = w16(v[j] * (TS == 0 ? 16 : 4));
    if (TS == 0 && lane == 0 && v[0]) v[0] += 1;
    Tx1D<N>::fwd(col_adst, v, t);
    if (tx_type == 0) { /* fused DCT of fwd_txfm.c */
#pragma unroll
        for (int j = 0; j < N; j++) v[j] = w16(t[j]);
        tile_transpose<N>(v, tile, lane, gmask);
        if (TS == 2) {
#pragma unroll
            for (int j = 0; j < N; j++) v[j] = (v[j] + 1) >> 2;
        }
        Tx1D<N>::fwd(false, v, t);
#pragma unroll
        for (int j = 0; j < N; j++) {
            const int o = w16(t[j]);
            v[j] = TS == 0 ? w16((o + 1) >> 2) : TS == 1 ? w16(o / 2) : o;
        }
    } else { /* hybrid transforms of vp9_dct.c */
#pragma unroll
        for (int j = 0; j < N; j++) { const int o = w16(t[j]); v[j] = TS == 2 ? w16((o + 1 + (o < 0)) >> 2) : o; }
        tile_transpose<N>(v, tile, lane, gmask);
        Tx1D<N>::fwd(row_adst, v, t);
#pragma unroll
        for (int j = 0; j < N; j++) {
            const int o = w16(t[j]);
            v[j] = TS == 0 ? w16((o + 1) >> 2) : TS == 1 ? w16((o + (o < 0)) >> 1) : o;
        }
    }
}

template <int N> __device__ __forceinline__ int group_max(int v) {
#pragma unroll
    for (int o = N / 2; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

/* store one int16 row of N values held as ints: 16-byte vector stores */
template <int N> __device__ __forceinline__ void store_row16(int16_t *dst, const int (&v)[N]) {
    if (N == 4) {
        uint2 p;
        p.x = (uint32_t)(uint16_t)v[0] | ((uint32_t)(uint16_t)v[1] << 16);
        p.y = (uint32_t)(uint16_t)v[2] | ((uint32_t)(uint16_t)v[3] << 16);
        *(uint2 *)dst = p;
    } else {
#pragma unroll
        for (int k = 0; k < N; k += 8) {
            uint4 p;
            p.x = (uint32_t)(uint16_t)v[k] | ((uint32_t)(uint16_t)v[k + 1] << 16);
            p.y = (uint32_t)(uint16_t)v[k + 2] | ((uint32_t)(uint16_t)v[k + 3] << 16);
            p.z = (uint32_t)(uint16_t)v[k + 4] | ((uint32_t)(uint16_t)v[k + 5] << 16);
            p.w = (uint32_t)(uint16_t)v[k + 6] | ((uint32_t)(uint16_t)v[k + 7] << 16);
            *(uint4 *)(dst + k) = p;
        }
    }
}
template <int N> __device__ __forceinline__ void load_row16(const int16_t *src, int (&v)[N]) {
    if (N == 4) {
        const uint2 p = *(const uint2 *)src;
        v[0] = (short)(p.x & 0xffff); v[1] = (short)(p.x >> 16); v[2] = (short)(p.y & 0xffff); v[3] = (short)(p.y >> 16);
    } else {
#pragma unroll
        for (int k = 0; k < N; k += 8) {
            const uint4 p = *(const uint4 *)(src + k);
            v[k] = (short)(p.x & 0xffff); v[k + 1] = (short)(p.x >> 16); v[k + 2] = (short)(p.y & 0xffff); v[k + 3] = (short)(p.y >> 16);
            v[k + 4] = (short)(p.z & 0xffff); v[k + 5] = (short)(p.z >> 16); v[k + 6] = (short)(p.w & 0xffff); v[k + 7] = (short)(p.w >> 16);
        }
    }
}

/* 4 * W consecutive bytes at ANY byte address as W little-endian words, without byte loads: one aligned vector load when the
 * address allows it, else the enclosing aligned words re-aligned with funnel shifts (reads at most the aligned word that holds the
 * last byte). */
template <int W> __device__ __forceinline__ void ld_bytes(const uint8_t *p, uint32_t (&w)[W]) {
    const uintptr_t a = (uintptr_t)p;
    if constexpr (W == 4) { if ((a & 15) == 0) { const uint4 v = __ldg((const uint4 *)p); w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w; return; } }
    if constexpr (W == 2) { if ((a & 7) == 0) { const uint2 v = __ldg((const uint2 *)p); w[0] = v.x; w[1] = v.y; return; } }
    const uint32_t *q = (const uint32_t *)(a & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(a & 3) * 8;
    uint32_t lo = __ldg(q);
    if (sh == 0) {
        w[0] = lo;
#pragma unroll
        for (int k = 1; k < W; k++) w[k] = __ldg(q + k);
    } else {
#pragma unroll
        for (int k = 0; k < W; k++) { const uint32_t hi = __ldg(q + k + 1); w[k] = __funnelshift_r(lo, hi, sh); lo = hi; }
    }
}
/* 4 * W bytes to a 4-byte aligned address (16-byte vector when the address allows it) */
template <int W> __device__ __forceinline__ void st_bytes(uint8_t *p, const uint32_t (&w)[W]) {
    const uintptr_t a = (uintptr_t)p;
    if constexpr (W == 4) { if ((a & 15) == 0) { *(uint4 *)p = make_uint4(w[0], w[1], w[2], w[3]); return; } }
    if constexpr (W == 2) { if ((a & 7) == 0) { *(uint2 *)p = make_uint2(w[0], w[1]); return; } }
#pragma unroll
    for (int k = 0; k < W; k++) ((uint32_t *)p)[k] = w[k];
}
/* The rows of an NxN block are moved by the block's N lanes in segments of SEG = min(N, 16) samples: lane -> (row, column of the
 * segment); N == 32 takes two passes of 16 rows x 2 segments. */
template <int N> struct SegMap {
    static constexpr int SEG = N < 16 ? N : 16, PER_ROW = N / SEG, ROWS = N / PER_ROW, PASSES = N / ROWS, W = SEG / 4;
};

/* residual = src - pred of the block at (s, p): every lane fetches whole segments (vector loads), the block is handed over through
 * the tile, lane c leaves with column c in v[] */
template <int N>
__device__ __forceinline__ void load_residual_columns(const uint8_t *s, int ss, const uint8_t *p, int ps, int (&v)[N], int *tile, int lane, unsigned gmask) {
    typedef SegMap<N> M;
#pragma unroll
    for (int i = 0; i < M::PASSES; i++) {
        const int r = i * M::ROWS + lane / M::PER_ROW, c0 = (lane % M::PER_ROW) * M::SEG;
        uint32_t a[M::W], b[M::W];
        ld_bytes<M::W>(s + (size_t)r * ss + c0, a);
        ld_bytes<M::W>(p + (size_t)r * ps + c0, b);
#pragma unroll
        for (int k = 0; k < M::SEG; k++)
            tile[r * (N + 1) + c0 + k] = (int)((a[k >> 2] >> (8 * (k & 3))) & 0xff) - (int)((b[k >> 2] >> (8 * (k & 3))) & 0xff);
    }
    __syncwarp(gmask);
#pragma unroll
    for (int j = 0; j < N; j++) v[j] = tile[j * (N + 1) + lane];
    __syncwarp(gmask);
}
/* reconstruction: lane c holds the inverse transform's column c (res[], already rounded and shifted); handed back through the tile,
 * every lane adds the prediction to whole segments and stores them as vectors */
template <int N>
__device__ __forceinline__ void store_recon_rows(const int (&res)[N], const uint8_t *p, int ps, uint8_t *d, int ds, int *tile, int lane, unsigned gmask, bool live) {
    typedef SegMap<N> M;
#pragma unroll
    for (int j = 0; j < N; j++) tile[j * (N + 1) + lane] = res[j];
    __syncwarp(gmask);
    if (live) {
#pragma unroll
        for (int i = 0; i < M::PASSES; i++) {
            const int r = i * M::ROWS + lane / M::PER_ROW, c0 = (lane % M::PER_ROW) * M::SEG;
            uint32_t b[M::W], o[M::W];
            ld_bytes<M::W>(p + (size_t)r * ps + c0, b);
#pragma unroll
            for (int k = 0; k < M::W; k++) o[k] = 0;
#pragma unroll
            for (int k = 0; k < M::SEG; k++) {
                const int x = (int)((b[k >> 2] >> (8 * (k & 3))) & 0xff) + tile[r * (N + 1) + c0 + k];
                o[k >> 2] |= (uint32_t)min(max(x, 0), 255) << (8 * (k & 3));
            }
            st_bytes<M::W>(d + (size_t)r * ds + c0, o);
        }
    }
    __syncwarp(gmask);
}

/* quantise coefficient row 'lane' (c[]): writes the q / dq rows (16-byte stores) when live, keeps dq[] for the
 * inverse transform; returns the block's eob on every lane */
template <int N>
__device__ __forceinline__ int quant_row(const int (&c)[N], int (&dq)[N], const SvtB200QuantRow &qr, const int16_t *iscan, int lane,
                                         int16_t *q_dst, int16_t *dq_dst, bool live) {
    constexpr int TS = TxCfg<N>::TS;
    /* this lane's row of the inverse scan, fetched with vector loads */
    int isc[N];
    load_row16<N>(iscan + lane * N, isc);
    /* AC constants; the DC entry differs only for coefficient (0,0), i.e. lane 0, k 0 */
    const int zb_ac = TS == 3 ? (qr.zbin[1] + 1) >> 1 : qr.zbin[1], rd_ac = TS == 3 ? (qr.round[1] + 1) >> 1 : qr.round[1];
    const int qu_ac = qr.quant[1], qs_ac = qr.quant_shift[1], dq_ac = qr.dequant[1];
    const bool dc = lane == 0;
    const int zb_0 = dc ? (TS == 3 ? (qr.zbin[0] + 1) >> 1 : qr.zbin[0]) : zb_ac, rd_0 = dc ? (TS == 3 ? (qr.round[0] + 1) >> 1 : qr.round[0]) : rd_ac;
    const int qu_0 = dc ? qr.quant[0] : qu_ac, qs_0 = dc ? qr.quant_shift[0] : qs_ac, dq_0 = dc ? qr.dequant[0] : dq_ac;
    int nzc = N * N;
    if (TS < 3) { /* trailing coefficients (scan order) inside the dead zone are dropped */
        int last = -1;
#pragma unroll
        for (int k = 0; k < N; k++) {
            const int zb = k ? zb_ac : zb_0;
            if (!(c[k] < zb && c[k] > -zb)) last = max(last, isc[k]);
        }
        nzc = group_max<N>(last) + 1;
    }
    int eob = -1;
    int q[N];
#pragma unroll
    for (int k = 0; k < N; k++) {
        const int zb = k ? zb_ac : zb_0, rd = k ? rd_ac : rd_0, qu = k ? qu_ac : qu_0, qs = k ? qs_ac : qs_0, dv = k ? dq_ac : dq_0;
        const int sign = c[k] >> 31, a = (c[k] ^ sign) - sign;
        /* branch-free: quantise every coefficient, keep the result only outside the dead zone and before the cut-off */
        int t = min(a + rd, 32767);
        t = ((((t * qu) >> 16) + t) * qs) >> (TS == 3 ? 15 : 16);
        const bool keep = a >= zb && isc[k] < nzc;
        t = keep ? t : 0;
        const int qv = w16((t ^ sign) - sign);
        q[k]  = qv;
        dq[k] = TS == 3 ? w16(qv * dv / 2) : w16(qv * dv);
        eob = t ? max(eob, isc[k]) : eob;
    }
    if (live) { store_row16<N>(q_dst, q); store_row16<N>(dq_dst, dq); }
    return group_max<N>(eob) + 1;
}

/* MODE 0 fused from planes, 1 forward only, 2 quantise only, 3 inverse + add only.
 * One pass = the G = TXQ_THREADS / N blocks [blk0, blk0 + G) of a batch, one block per group of N lanes. */
template <int N, int MODE>
__device__ __forceinline__ void txq_pass(const TxqArgs &A, int blk0, int *tiles) {
    constexpr int SHIFT = TxCfg<N>::SHIFT, TS = TxCfg<N>::TS;
    const int grp = threadIdx.x / N, lane = threadIdx.x % N;
    int *tile = tiles + grp * N * (N + 1);
    const unsigned gmask = N == 32 ? 0xffffffffu : (((1u << N) - 1u) << ((threadIdx.x & 31) - (threadIdx.x & 31) % N));
    /* all lanes of a warp work together (blocks past the end are processed on dummy data, stores masked) */
    const int blk = blk0 + grp;
    const bool live = blk < A.n;
    const int b = live ? blk : A.n - 1;
    int v[N], dq[N];
    int tx_type = A.tx_type, flags = 0, qrow = 0;
    const uint8_t *pred_p = nullptr;
    int pred_s = 0;
    uint32_t recon_off = 0;
    if (MODE == 0) {
        const uint4 jw = __ldg((const uint4 *)(A.jobs + b)); /* SvtB200TxJob as one 16-byte vector: src, pred, recon offsets, qrow | tx_type << 16 | flags << 24 */
        tx_type = TS == 3 ? 0 : (int)((jw.w >> 16) & 0xff); flags = (int)(jw.w >> 24); qrow = (int)(jw.w & 0xffff); recon_off = jw.z;
        pred_p = A.pl.pred + jw.y; pred_s = A.pl.pred_stride;
        load_residual_columns<N>(A.pl.src + jw.x, A.pl.src_stride, pred_p, pred_s, v, tile, lane, gmask);
    } else if (MODE == 1) {
        load_row16<N>(A.in16 + (size_t)b * N * N + lane * N, v); /* row 'lane' with vector loads, then lane c takes column c */
        tile_transpose<N>(v, tile, lane, gmask);
        flags = A.partial32 ? SVTB200_TXJ_PARTIAL32 : 0;
    } else {
        load_row16<N>(A.in16 + (size_t)b * N * N + lane * N, v);
        if (MODE == 2 && A.qrow_index) qrow = A.qrow_index[b];
    }
    if (MODE == 0 || MODE == 1) {
        fwd2d<N>(v, tx_type, tile, lane, gmask);
        if (TS == 3 && (flags & SVTB200_TXJ_PARTIAL32)) {
#pragma unroll
            for (int k = 0; k < N; k++) if (lane >= 16 || k >= 16) v[k] = 0;
        }
        if (A.coeff && live) store_row16<N>(A.coeff + (size_t)b * N * N + lane * N, v);
    }
    if (MODE == 0 || MODE == 2) {
        const SvtB200QuantRow qr = A.qrows[qrow];
        const int eob = quant_row<N>(v, dq, qr, A.iscan[TS == 3 ? 0 : tx_type], lane, A.qcoeff + (size_t)b * N * N + lane * N,
                                     A.dqcoeff + (size_t)b * N * N + lane * N, live);
        if (live && lane == 0) A.eob[b] = (uint16_t)eob;
    }
    if (MODE == 3 || (MODE == 0 && (flags & SVTB200_TXJ_RECON))) {
        int t[N];
        if (MODE == 3) {
#pragma unroll
            for (int k = 0; k < N; k++) dq[k] = v[k];
        }
        /* rows first, then columns (vp9_idct.c / inv_txfm.c 2-D drivers) */
        Tx1D<N>::inv(TS < 3 && (tx_type & 2), dq, t);
        tile_transpose<N>(t, tile, lane, gmask);
        Tx1D<N>::inv(TS < 3 && (tx_type & 1), t, dq);
#pragma unroll
        for (int j = 0; j < N; j++) dq[j] = w16((dq[j] + (1 << (SHIFT - 1))) >> SHIFT);
        if (MODE == 3) { uint8_t *d = A.dest + (size_t)b * N * N; store_recon_rows<N>(dq, d, N, d, N, tile, lane, gmask, live); }
        else store_recon_rows<N>(dq, pred_p, pred_s, A.pl.recon + recon_off, A.pl.recon_stride, tile, lane, gmask, live);
    }
}

template <int N, int MODE>
__global__ void __launch_bounds__(TXQ_THREADS, TXQ_MIN_BLOCKS) txq_kernel(const __grid_constant__ TxqArgs A) {
    constexpr int G = TXQ_THREADS / N; /* blocks per CTA pass */
    __shared__ int tiles[G * N * (N + 1)];
    for (int blk0 = blockIdx.x * G; blk0 < A.n; blk0 += gridDim.x * G) txq_pass<N, MODE>(A, blk0, tiles);
}

/* Several fused batches (different transform sizes, different planes) in ONE launch: a picture's luma 32x32, chroma 16x16 and
 * edge-row batches are a few hundred to a few thousand blocks each; launched one by one the small ones are latency-bound 5-7 us
 * kernels.  Work unit = one CTA pass of one batch; units of all batches form one index space. */
#define TXQ_MAX_BATCHES 8
struct TxqMultiArgs {
    TxqArgs batch[TXQ_MAX_BATCHES];
    int32_t tx_size[TXQ_MAX_BATCHES];
    int32_t unit_end[TXQ_MAX_BATCHES]; /* cumulative number of CTA passes */
    int32_t n_batches;
};
__global__ void __launch_bounds__(TXQ_THREADS, TXQ_MIN_BLOCKS) txq_multi_kernel(const __grid_constant__ TxqMultiArgs M) {
    __shared__ int tiles[4 * 32 * 33];
    const int total = M.unit_end[M.n_batches - 1];
    for (int u = blockIdx.x; u < total; u += gridDim.x) {
        int k = 0;
        while (u >= M.unit_end[k]) k++;
        const int local = u - (k ? M.unit_end[k - 1] : 0);
        const TxqArgs &A = M.batch[k];
        switch (M.tx_size[k]) { /* uniform over the CTA */
        case 0: txq_pass<4, 0>(A, local * (TXQ_THREADS / 4), tiles); break;
        case 1: txq_pass<8, 0>(A, local * (TXQ_THREADS / 8), tiles); break;
        case 2: txq_pass<16, 0>(A, local * (TXQ_THREADS / 16), tiles); break;
        default: txq_pass<32, 0>(A, local * (TXQ_THREADS / 32), tiles); break;
        }
        __syncthreads(); /* the next unit may belong to a batch of another size, which partitions the tile memory differently between the warps */
    }
}
#endif
