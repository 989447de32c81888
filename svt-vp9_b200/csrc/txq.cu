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

/* MODE 0 fused from planes, 1 forward only, 2 quantise only, 3 inverse + add only */
template <int N, int MODE>
__global__ void __launch_bounds__(TXQ_THREADS, 4) txq_kernel(const __grid_constant__ TxqArgs A) {
    constexpr int G = TXQ_THREADS / N; /* blocks per CTA pass */
    constexpr int SHIFT = TxCfg<N>::SHIFT, TS = TxCfg<N>::TS;
    __shared__ int tiles[G * N * (N + 1)];
    const int grp = threadIdx.x / N, lane = threadIdx.x % N;
    int *tile = tiles + grp * N * (N + 1);
    const unsigned gmask = N == 32 ? 0xffffffffu : (((1u << N) - 1u) << ((threadIdx.x & 31) - (threadIdx.x & 31) % N));
    /* all lanes of a warp iterate together (blocks past the end are processed on dummy data, stores masked) */
    for (int blk0 = blockIdx.x * G; blk0 < A.n; blk0 += gridDim.x * G) {
        const int blk = blk0 + grp;
        const bool live = blk < A.n;
        const int b = live ? blk : A.n - 1;
        int v[N], dq[N];
        int tx_type = A.tx_type, flags = 0, qrow = 0;
        const uint8_t *pred_p = nullptr;
        int pred_s = 0;
        if (MODE == 0) {
            const SvtB200TxJob job = A.jobs[b];
            tx_type = TS == 3 ? 0 : job.tx_type; flags = job.flags; qrow = job.qrow;
            const uint8_t *s = A.pl.src + job.src_offset + lane;
            pred_p = A.pl.pred + job.pred_offset + lane; pred_s = A.pl.pred_stride;
#pragma unroll
            for (int j = 0; j < N; j++) v[j] = (int)__ldg(s + (size_t)j * A.pl.src_stride) - (int)__ldg(pred_p + (size_t)j * pred_s);
        } else if (MODE == 1) {
            const int16_t *s = A.in16 + (size_t)b * N * N + lane;
#pragma unroll
            for (int j = 0; j < N; j++) v[j] = __ldg(s + j * N);
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
            uint8_t *d;
            int ds;
            const uint8_t *p;
            int ps;
            if (MODE == 3) { d = A.dest + (size_t)b * N * N + lane; ds = N; p = d; ps = N; }
            else { d = A.pl.recon + A.jobs[b].recon_offset + lane; ds = A.pl.recon_stride; p = pred_p; ps = pred_s; }
            if (live) {
#pragma unroll
                for (int j = 0; j < N; j++) {
                    const int r = (int)p[(size_t)j * ps] + w16((dq[j] + (1 << (SHIFT - 1))) >> SHIFT);
                    d[(size_t)j * ds] = (uint8_t)min(max(r, 0), 255);
                }
            }
        }
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
    int rc = fill_scans(c, tx_size, &A);
    if (rc) return rc;
    SVTB200_CUDA_CHECK(c, launch_mode<0>(tx_size, A, c->sm_count, c->stream));
    c->launches++;
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
