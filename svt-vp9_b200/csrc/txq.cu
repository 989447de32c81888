/*
 * txq.cu -- S3: batched forward transform, quantisation, inverse transform + reconstruction.
 *
 * Replaces, for a whole batch of transform blocks per launch, what perform_coding_loop() does per block
 * (Source/Lib/Codec/EbEncDecProcess.c:365-587): residual = src - pred; eb_vp9_fdct* / fht* or
 * eb_vpx_partial_fdct32x32; eb_vp9_quantize_b[_32x32]; and, for reconstruction, eb_vp9_idct*_add / iht*_add.
 *
 * Mapping: an NxN block is owned by N consecutive lanes of one warp (N = 4..32, so 8..1 blocks per warp).
 * Lane c holds column c in registers, runs the column transform, the block is transposed through shared
 * memory (padded pitch, warp-synchronous), lane r then holds row r for the row transform, the quantiser and
 * the inverse row transform; a second transpose feeds the inverse column transform and the reconstruction.
 * Coefficient rows are written with 16-byte stores; eob is a shuffle max-reduction over the block's lanes.
 * The data is touched once: the fused kernel moves 6 N^2 + 2 bytes per block (+ N^2 for the reconstruction),
 * which is what the HBM roofline of this stage is computed from.
 */
#include "common.cuh"
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

/* ------------------------------------------------------------------------------------------------------------ */
template <int MODE> static cudaError_t launch_mode(int tx_size, const TxqArgs &A, int sm_count, cudaStream_t st) {
    const int N = 4 << tx_size, G = TXQ_THREADS / N;
    int grid = (A.n + G - 1) / G;
    const int cap = sm_count * 16; /* a few waves, grid-stride beyond */
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    switch (tx_size) {
    case 0: txq_kernel<4, MODE><<<grid, TXQ_THREADS, 0, st>>>(A); break;
    case 1: txq_kernel<8, MODE><<<grid, TXQ_THREADS, 0, st>>>(A); break;
    case 2: txq_kernel<16, MODE><<<grid, TXQ_THREADS, 0, st>>>(A); break;
    default: txq_kernel<32, MODE><<<grid, TXQ_THREADS, 0, st>>>(A); break;
    }
    return cudaGetLastError();
}

static int fill_scans(SvtB200Ctx *c, int tx_size, TxqArgs *A) {
    for (int t = 0; t < 4; t++) {
        const ScanTable &s = c->scans[tx_size][tx_size == 3 ? 0 : t];
        A->iscan[t] = s.iscan;
    }
    if (!A->iscan[0]) { c->err = "scan tables not registered (svtb200_set_scan)"; return SVTB200_EINVAL; }
    return 0;
}

/* planes of a fused batch: reconstruction segments are stored as 32-bit words (16-byte vectors when aligned) */
static int check_planes(SvtB200Ctx *c, const SvtB200TxPlanes *pl) {
    if (pl->recon && ((pl->recon_stride & 3) || ((uintptr_t)pl->recon & 3))) { c->err = "txq: recon plane base and stride must be multiples of 4"; return SVTB200_EINVAL; }
    return 0;
}

static uint8_t *carve(uint8_t *&p, size_t bytes) { uint8_t *r = p; p += (bytes + 255) & ~(size_t)255; return r; }

extern "C" {

int svtb200_set_scan(SvtB200Ctx *c, int tx_size, int tx_type, const int16_t *scan, const int16_t *iscan) {
    if (!c || tx_size < 0 || tx_size > 3 || tx_type < 0 || tx_type > 3 || !scan || !iscan) return SVTB200_EINVAL;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    cudaSetDevice(c->device);
    const int n = (4 << tx_size) * (4 << tx_size);
    ScanTable &s = c->scans[tx_size][tx_type];
    if (!s.scan) {
        if (cudaMalloc((void **)&s.scan, n * 2) != cudaSuccess || cudaMalloc((void **)&s.iscan, n * 2) != cudaSuccess) { cudaGetLastError(); return SVTB200_ENOMEM; }
    }
    s.n = n;
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(s.scan, scan, n * 2, cudaMemcpyHostToDevice, c->stream));
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(s.iscan, iscan, n * 2, cudaMemcpyHostToDevice, c->stream));
    SVTB200_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    return SVTB200_OK;
}

int svtb200_fwd_txfm_batch_dev(SvtB200Ctx *c, int tx_size, int tx_type, int partial32, int n, const int16_t *in, int16_t *out) {
    if (!c || tx_size < 0 || tx_size > 3 || tx_type < 0 || tx_type > 3 || n < 0 || !in || !out) return SVTB200_EINVAL;
    if (n == 0) return SVTB200_OK;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    cudaSetDevice(c->device);
    TxqArgs A; memset(&A, 0, sizeof(A));
    A.in16 = in; A.coeff = out; A.n = n; A.tx_type = tx_size == 3 ? 0 : tx_type; A.partial32 = partial32 && tx_size == 3;
    SVTB200_CUDA_CHECK(c, launch_mode<1>(tx_size, A, c->sm_count, c->stream));
    c->launches++;
    return SVTB200_OK;
}

int svtb200_quantize_batch_dev(SvtB200Ctx *c, int tx_size, int tx_type, int n, const int16_t *coeff, const SvtB200QuantRow *qrows,
                               int n_qrows, const uint16_t *qrow_index, int16_t *qcoeff, int16_t *dqcoeff, uint16_t *eob) {
    if (!c || tx_size < 0 || tx_size > 3 || tx_type < 0 || tx_type > 3 || n < 0 || !coeff || !qrows || n_qrows < 1 || !qcoeff || !dqcoeff || !eob) return SVTB200_EINVAL;
    if (n == 0) return SVTB200_OK;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    cudaSetDevice(c->device);
    TxqArgs A; memset(&A, 0, sizeof(A));
    A.in16 = coeff; A.qrows = qrows; A.qrow_index = qrow_index; A.qcoeff = qcoeff; A.dqcoeff = dqcoeff; A.eob = eob;
    A.n = n; A.tx_type = tx_size == 3 ? 0 : tx_type;
    int rc = fill_scans(c, tx_size, &A);
    if (rc) return rc;
    SVTB200_CUDA_CHECK(c, launch_mode<2>(tx_size, A, c->sm_count, c->stream));
    c->launches++;
    return SVTB200_OK;
}

int svtb200_inv_txfm_add_batch_dev(SvtB200Ctx *c, int tx_size, int tx_type, int n, const int16_t *dq, const uint16_t *eob, uint8_t *dest) {
    if (!c || tx_size < 0 || tx_size > 3 || tx_type < 0 || tx_type > 3 || n < 0 || !dq || !dest) return SVTB200_EINVAL;
    if (n == 0) return SVTB200_OK;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    cudaSetDevice(c->device);
    TxqArgs A; memset(&A, 0, sizeof(A));
    A.in16 = dq; A.eob_in = eob; A.dest = dest; A.n = n; A.tx_type = tx_size == 3 ? 0 : tx_type;
    SVTB200_CUDA_CHECK(c, launch_mode<3>(tx_size, A, c->sm_count, c->stream));
    c->launches++;
    return SVTB200_OK;
}

int svtb200_txq_batch_dev(SvtB200Ctx *c, int tx_size, int n, const SvtB200TxJob *jobs, const SvtB200TxPlanes *pl,
                          const SvtB200QuantRow *qrows, int n_qrows, int16_t *coeff, int16_t *qcoeff, int16_t *dqcoeff, uint16_t *eob) {
    if (!c || tx_size < 0 || tx_size > 3 || n < 0 || !jobs || !pl || !qrows || n_qrows < 1 || !qcoeff || !dqcoeff || !eob) return SVTB200_EINVAL;
    if (n == 0) return SVTB200_OK;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    cudaSetDevice(c->device);
    TxqArgs A; memset(&A, 0, sizeof(A));
    A.jobs = jobs; A.pl = *pl; A.qrows = qrows; A.coeff = coeff; A.qcoeff = qcoeff; A.dqcoeff = dqcoeff; A.eob = eob; A.n = n;
    int rc = check_planes(c, pl);
    if (rc) return rc;
    if ((rc = fill_scans(c, tx_size, &A))) return rc;
    SVTB200_CUDA_CHECK(c, launch_mode<0>(tx_size, A, c->sm_count, c->stream));
    c->launches++;
    return SVTB200_OK;
}

int svtb200_txq_multi_dev(SvtB200Ctx *c, const SvtB200TxBatch *batches, int n_batches, const SvtB200QuantRow *qrows, int n_qrows) {
    if (!c || !batches || n_batches < 1 || n_batches > TXQ_MAX_BATCHES || !qrows || n_qrows < 1) return SVTB200_EINVAL;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    cudaSetDevice(c->device);
    TxqMultiArgs M;
    memset(&M, 0, sizeof(M));
    int units = 0, nb = 0;
    for (int k = 0; k < n_batches; k++) {
        const SvtB200TxBatch &b = batches[k];
        if (b.tx_size < 0 || b.tx_size > 3 || b.n < 0 || (b.n && (!b.jobs || !b.qcoeff || !b.dqcoeff || !b.eob || !b.planes.src || !b.planes.pred))) return SVTB200_EINVAL;
        if (!b.n) continue;
        int rc = check_planes(c, &b.planes);
        if (rc) return rc;
        TxqArgs &A = M.batch[nb];
        A.jobs = b.jobs; A.pl = b.planes; A.qrows = qrows; A.coeff = b.coeff; A.qcoeff = b.qcoeff; A.dqcoeff = b.dqcoeff; A.eob = b.eob; A.n = b.n;
        if ((rc = fill_scans(c, b.tx_size, &A))) return rc;
        const int G = TXQ_THREADS / (4 << b.tx_size);
        units += (b.n + G - 1) / G;
        M.tx_size[nb] = b.tx_size; M.unit_end[nb] = units;
        nb++;
    }
    if (!nb) return SVTB200_OK;
    M.n_batches = nb;
    int grid = units;
    const int cap = (c->sm_count > 0 ? c->sm_count : 148) * 16;
    if (grid > cap) grid = cap;
    txq_multi_kernel<<<grid, TXQ_THREADS, 0, c->stream>>>(M);
    c->launches++;
    SVTB200_CUDA_CHECK(c, cudaGetLastError());
    return SVTB200_OK;
}

/* ---- host-pointer variants: stage through device scratch, copies on the context stream ---- */
int svtb200_fwd_txfm_batch(SvtB200Ctx *c, int tx_size, int tx_type, int partial32, int n, const int16_t *in, int16_t *out) {
    if (!c || n < 0 || !in || !out || tx_size < 0 || tx_size > 3) return SVTB200_EINVAL;
    std::lock_guard<std::recursive_mutex> whole(c->mu); /* the staging buffer c->scratch is shared: one host call at a time */
    if (n == 0) return SVTB200_OK;
    const size_t bytes = (size_t)n * (4 << tx_size) * (4 << tx_size) * 2;
    {
        std::lock_guard<std::recursive_mutex> g(c->mu);
        cudaSetDevice(c->device);
        int rc = svtb200_grow(c, (void **)&c->scratch, &c->scratch_bytes, 2 * bytes + 1024);
        if (rc) return rc;
        SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(c->scratch, in, bytes, cudaMemcpyHostToDevice, c->stream));
    }
    uint8_t *p = c->scratch;
    int16_t *d_in = (int16_t *)carve(p, bytes), *d_out = (int16_t *)carve(p, bytes);
    int rc = svtb200_fwd_txfm_batch_dev(c, tx_size, tx_type, partial32, n, d_in, d_out);
    if (rc) return rc;
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(out, d_out, bytes, cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    return SVTB200_OK;
}

int svtb200_quantize_batch(SvtB200Ctx *c, int tx_size, int tx_type, int n, const int16_t *coeff, const SvtB200QuantRow *qrows,
                           int n_qrows, const uint16_t *qrow_index, int16_t *qcoeff, int16_t *dqcoeff, uint16_t *eob) {
    if (!c || n < 0 || !coeff || !qrows || n_qrows < 1 || !qcoeff || !dqcoeff || !eob || tx_size < 0 || tx_size > 3) return SVTB200_EINVAL;
    std::lock_guard<std::recursive_mutex> whole(c->mu); /* the staging buffer c->scratch is shared: one host call at a time */
    if (n == 0) return SVTB200_OK;
    const size_t bytes = (size_t)n * (4 << tx_size) * (4 << tx_size) * 2;
    uint8_t *p;
    {
        std::lock_guard<std::recursive_mutex> g(c->mu);
        cudaSetDevice(c->device);
        int rc = svtb200_grow(c, (void **)&c->scratch, &c->scratch_bytes, 3 * bytes + (size_t)n * 4 + (size_t)n_qrows * sizeof(SvtB200QuantRow) + 4096);
        if (rc) return rc;
        p = c->scratch;
    }
    int16_t *d_c = (int16_t *)carve(p, bytes), *d_q = (int16_t *)carve(p, bytes), *d_dq = (int16_t *)carve(p, bytes);
    uint16_t *d_eob = (uint16_t *)carve(p, (size_t)n * 2), *d_idx = (uint16_t *)carve(p, (size_t)n * 2);
    SvtB200QuantRow *d_rows = (SvtB200QuantRow *)carve(p, (size_t)n_qrows * sizeof(SvtB200QuantRow));
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(d_c, coeff, bytes, cudaMemcpyHostToDevice, c->stream));
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(d_rows, qrows, (size_t)n_qrows * sizeof(SvtB200QuantRow), cudaMemcpyHostToDevice, c->stream));
    if (qrow_index) SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(d_idx, qrow_index, (size_t)n * 2, cudaMemcpyHostToDevice, c->stream));
    int rc = svtb200_quantize_batch_dev(c, tx_size, tx_type, n, d_c, d_rows, n_qrows, qrow_index ? d_idx : nullptr, d_q, d_dq, d_eob);
    if (rc) return rc;
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(qcoeff, d_q, bytes, cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(dqcoeff, d_dq, bytes, cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(eob, d_eob, (size_t)n * 2, cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    return SVTB200_OK;
}

int svtb200_inv_txfm_add_batch(SvtB200Ctx *c, int tx_size, int tx_type, int n, const int16_t *dq, const uint16_t *eob, uint8_t *dest) {
    if (!c || n < 0 || !dq || !dest || tx_size < 0 || tx_size > 3) return SVTB200_EINVAL;
    std::lock_guard<std::recursive_mutex> whole(c->mu); /* the staging buffer c->scratch is shared: one host call at a time */
    if (n == 0) return SVTB200_OK;
    const size_t nn = (size_t)(4 << tx_size) * (4 << tx_size), bytes = (size_t)n * nn * 2;
    uint8_t *p;
    {
        std::lock_guard<std::recursive_mutex> g(c->mu);
        cudaSetDevice(c->device);
        int rc = svtb200_grow(c, (void **)&c->scratch, &c->scratch_bytes, bytes + (size_t)n * nn + 2048);
        if (rc) return rc;
        p = c->scratch;
    }
    int16_t *d_dq = (int16_t *)carve(p, bytes);
    uint8_t *d_dst = carve(p, (size_t)n * nn);
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(d_dq, dq, bytes, cudaMemcpyHostToDevice, c->stream));
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(d_dst, dest, (size_t)n * nn, cudaMemcpyHostToDevice, c->stream));
    int rc = svtb200_inv_txfm_add_batch_dev(c, tx_size, tx_type, n, d_dq, nullptr, d_dst);
    (void)eob;
    if (rc) return rc;
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(dest, d_dst, (size_t)n * nn, cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    return SVTB200_OK;
}

int svtb200_txq_batch(SvtB200Ctx *c, int tx_size, int n, const SvtB200TxJob *jobs, const SvtB200TxPlanes *pl,
                      const SvtB200QuantRow *qrows, int n_qrows, int16_t *coeff, int16_t *qcoeff, int16_t *dqcoeff, uint16_t *eob) {
    if (!c || n < 0 || !jobs || !pl || !pl->src || !pl->pred || !qrows || n_qrows < 1 || !qcoeff || !dqcoeff || !eob || tx_size < 0 || tx_size > 3) return SVTB200_EINVAL;
    std::lock_guard<std::recursive_mutex> whole(c->mu); /* the staging buffer c->scratch is shared: one host call at a time */
    if (n == 0) return SVTB200_OK;
    if (pl->recon) for (int i = 0; i < n; i++) if ((jobs[i].flags & SVTB200_TXJ_RECON) && (jobs[i].recon_offset & 3)) { c->err = "txq: recon_offset must be a multiple of 4"; return SVTB200_EINVAL; }
    const size_t nn = (size_t)(4 << tx_size) * (4 << tx_size), cb = (size_t)n * nn * 2;
    uint8_t *p;
    {
        std::lock_guard<std::recursive_mutex> g(c->mu);
        cudaSetDevice(c->device);
        const size_t need = pl->src_bytes + pl->pred_bytes + pl->recon_bytes + 3 * cb + (size_t)n * (sizeof(SvtB200TxJob) + 2) +
                            (size_t)n_qrows * sizeof(SvtB200QuantRow) + 8192;
        int rc = svtb200_grow(c, (void **)&c->scratch, &c->scratch_bytes, need);
        if (rc) return rc;
        p = c->scratch;
    }
    SvtB200TxPlanes d = *pl;
    uint8_t *d_src = carve(p, pl->src_bytes), *d_pred = carve(p, pl->pred_bytes);
    uint8_t *d_rec = pl->recon ? carve(p, pl->recon_bytes) : nullptr;
    int16_t *d_c = coeff ? (int16_t *)carve(p, cb) : nullptr, *d_q = (int16_t *)carve(p, cb), *d_dq = (int16_t *)carve(p, cb);
    uint16_t *d_eob = (uint16_t *)carve(p, (size_t)n * 2);
    SvtB200TxJob *d_jobs = (SvtB200TxJob *)carve(p, (size_t)n * sizeof(SvtB200TxJob));
    SvtB200QuantRow *d_rows = (SvtB200QuantRow *)carve(p, (size_t)n_qrows * sizeof(SvtB200QuantRow));
    d.src = d_src; d.pred = d_pred; d.recon = d_rec;
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(d_src, pl->src, pl->src_bytes, cudaMemcpyHostToDevice, c->stream));
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(d_pred, pl->pred, pl->pred_bytes, cudaMemcpyHostToDevice, c->stream));
    if (d_rec) SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(d_rec, pl->recon, pl->recon_bytes, cudaMemcpyHostToDevice, c->stream));
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(d_jobs, jobs, (size_t)n * sizeof(SvtB200TxJob), cudaMemcpyHostToDevice, c->stream));
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(d_rows, qrows, (size_t)n_qrows * sizeof(SvtB200QuantRow), cudaMemcpyHostToDevice, c->stream));
    int rc = svtb200_txq_batch_dev(c, tx_size, n, d_jobs, &d, d_rows, n_qrows, d_c, d_q, d_dq, d_eob);
    if (rc) return rc;
    if (coeff) SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(coeff, d_c, cb, cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(qcoeff, d_q, cb, cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(dqcoeff, d_dq, cb, cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(eob, d_eob, (size_t)n * 2, cudaMemcpyDeviceToHost, c->stream));
    if (d_rec) SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(pl->recon, d_rec, pl->recon_bytes, cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    return SVTB200_OK;
}

} /* extern "C" */
