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

#include "txq_kernel.cuh"

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
