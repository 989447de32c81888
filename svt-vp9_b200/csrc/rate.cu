/*
 * rate.cu -- coefficient rate estimation for a batch of transform blocks.
 *
 * Replaces, per block, coeff_rate_estimate() with use_fast_coef_costing == 0 -- the only way its caller invokes it
 * (Source/Lib/Codec/EbRateDistortionCost.c:55, call site Source/Lib/Codec/EbEncDecProcess.c:736; libvpx twin cost_coeffs,
 * Source/Lib/VPX/vp9_rdopt.c:334): the number of bits (in 1/256 bit units of the cost tables) the tokens of the block's quantised
 * coefficients take, walking the scan order with the neighbour-derived context of every position.
 *
 * The reference walks the scan sequentially because it keeps a running token_cache, band counter and "previous token" -- but
 * every one of those is a function of the coefficient VALUES alone: the token and its energy class depend only on the value at a
 * position (vp9_get_token_cost, Source/Lib/VPX/vp9_tokenize.h:118; eb_vp9_pt_energy_class, vp9_entropy.c:111), the band only on
 * the scan index (band_counts, EbRateDistortionCost.c:48), the context only on the two earlier neighbours' energy classes
 * (get_coef_context, vp9_scan.h:34) and the "skip EOB node" flag only on the previous coefficient in scan order.  The cost of a
 * block is therefore a sum of independent per-position terms, and the kernel is parallel over scan POSITIONS (rate_multi_kernel).
 * Integer arithmetic only: bit-exact by construction.
 */
#include "common.cuh"

#include "rate_kernel.cuh"

extern "C" {

int svtb200_set_scan_neighbors(SvtB200Ctx *c, int tx_size, int tx_type, const int16_t *neighbors) {
    if (!c || tx_size < 0 || tx_size > 3 || tx_type < 0 || tx_type > 3 || !neighbors) return SVTB200_EINVAL;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    cudaSetDevice(c->device);
    const size_t bytes = (size_t)2 * ((4 << tx_size) * (4 << tx_size) + 1) * sizeof(int16_t);
    int16_t *&d = c->scan_nb[tx_size][tx_type];
    if (!d && cudaMalloc((void **)&d, bytes) != cudaSuccess) { cudaGetLastError(); d = nullptr; return SVTB200_ENOMEM; }
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(d, neighbors, bytes, cudaMemcpyHostToDevice, c->stream));
    SVTB200_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    return SVTB200_OK;
}

int svtb200_set_rate_tables(SvtB200Ctx *c, const int16_t value_token[133], const int32_t value_cost[133], const int16_t cat6_low_cost[256],
                            const uint16_t cat6_high_cost[64]) {
    if (!c || !value_token || !value_cost || !cat6_low_cost || !cat6_high_cost) return SVTB200_EINVAL;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    cudaSetDevice(c->device);
    const size_t bytes = 133 * 2 + 2 + 133 * 4 + 256 * 2 + 64 * 2;
    if (!c->rate_tables && cudaMalloc((void **)&c->rate_tables, bytes) != cudaSuccess) { cudaGetLastError(); c->rate_tables = nullptr; return SVTB200_ENOMEM; }
    uint8_t *d = c->rate_tables;
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(d, value_cost, 133 * 4, cudaMemcpyHostToDevice, c->stream));
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(d + 133 * 4, value_token, 133 * 2, cudaMemcpyHostToDevice, c->stream));
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(d + 133 * 4 + 134 * 2, cat6_low_cost, 256 * 2, cudaMemcpyHostToDevice, c->stream));
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(d + 133 * 4 + 134 * 2 + 256 * 2, cat6_high_cost, 64 * 2, cudaMemcpyHostToDevice, c->stream));
    SVTB200_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    return SVTB200_OK;
}

static int rate_fill(SvtB200Ctx *c, RateArgs &A, int tx_size, int n, const int16_t *qcoeff, const uint16_t *eob, const SvtB200RateBlock *blocks,
                     const uint32_t *token_costs, int32_t *rate) {
    if (!c->rate_tables) { c->err = "rate tables not registered (svtb200_set_rate_tables)"; return SVTB200_EINVAL; }
    memset(&A, 0, sizeof(A));
    for (int t = 0; t < 4; t++) {
        const int tt = tx_size == 3 ? 0 : t;
        A.scan[t] = c->scans[tx_size][tt].scan; A.nb[t] = c->scan_nb[tx_size][tt];
        if (!A.scan[t] || !A.nb[t]) { c->err = "scan / neighbour tables not registered (svtb200_set_scan, svtb200_set_scan_neighbors)"; return SVTB200_EINVAL; }
    }
    A.qcoeff = qcoeff; A.eob = eob; A.blk = blocks; A.tc = token_costs; A.rate = rate; A.n = n; A.tx_size = tx_size;
    A.vcost = (const int32_t *)c->rate_tables; A.vtok = (const int16_t *)(c->rate_tables + 133 * 4);
    A.c6lo = (const int16_t *)(c->rate_tables + 133 * 4 + 134 * 2); A.c6hi = (const uint16_t *)(c->rate_tables + 133 * 4 + 134 * 2 + 256 * 2);
    return SVTB200_OK;
}

int svtb200_coeff_rate_multi_dev(SvtB200Ctx *c, const SvtB200RateBatch *batches, int n_batches) {
    if (!c || !batches || n_batches < 1 || n_batches > RATE_MAX_BATCHES) return SVTB200_EINVAL;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    cudaSetDevice(c->device);
    RateMultiArgs M;
    memset(&M, 0, sizeof(M));
    int total = 0;
    for (int k = 0; k < n_batches; k++) {
        const SvtB200RateBatch &b = batches[k];
        if (b.tx_size < 0 || b.tx_size > 3 || b.n < 0 || (b.n && (!b.qcoeff || !b.eob || !b.blocks || !b.token_costs || !b.rate))) return SVTB200_EINVAL;
        if (!b.n) continue;
        const int rc = rate_fill(c, M.b[M.nb], b.tx_size, b.n, b.qcoeff, b.eob, b.blocks, b.token_costs, b.rate);
        if (rc) return rc;
        total += b.n;
        M.end[M.nb++] = total;
    }
    if (!total) return SVTB200_OK;
    rate_multi_kernel<<<(total + RATE_CHUNK - 1) / RATE_CHUNK, RATE_THREADS, 0, c->stream>>>(M); /* the kernel's block assignment relies on this grid size */
    c->launches++;
    SVTB200_CUDA_CHECK(c, cudaGetLastError());
    return SVTB200_OK;
}

int svtb200_coeff_rate_batch_dev(SvtB200Ctx *c, int tx_size, int n, const int16_t *qcoeff, const uint16_t *eob, const SvtB200RateBlock *blocks,
                                 const uint32_t *token_costs, int32_t *rate) {
    if (!c || tx_size < 0 || tx_size > 3 || n < 0 || (n && (!qcoeff || !eob || !blocks || !token_costs || !rate))) return SVTB200_EINVAL;
    if (!n) return SVTB200_OK;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    cudaSetDevice(c->device);
    RateMultiArgs M;
    memset(&M, 0, sizeof(M));
    const int frc = rate_fill(c, M.b[0], tx_size, n, qcoeff, eob, blocks, token_costs, rate);
    if (frc) return frc;
    M.end[0] = n; M.nb = 1;
    rate_multi_kernel<<<(n + RATE_CHUNK - 1) / RATE_CHUNK, RATE_THREADS, 0, c->stream>>>(M);
    c->launches++;
    SVTB200_CUDA_CHECK(c, cudaGetLastError());
    return SVTB200_OK;
}

int svtb200_coeff_rate_batch(SvtB200Ctx *c, int tx_size, int n, const int16_t *qcoeff, const uint16_t *eob, const SvtB200RateBlock *blocks,
                             const uint32_t *token_costs, int32_t *rate) {
    if (!c || tx_size < 0 || tx_size > 3 || n < 0 || (n && (!qcoeff || !eob || !blocks || !token_costs || !rate))) return SVTB200_EINVAL;
    if (!n) return SVTB200_OK;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    cudaSetDevice(c->device);
    const size_t nn = (size_t)(4 << tx_size) * (4 << tx_size), qb = (size_t)n * nn * 2, tcb = (size_t)2 * 2 * 6 * 2 * 6 * RATE_TOKENS * 4;
    int rc = svtb200_grow(c, (void **)&c->scratch, &c->scratch_bytes, qb + (size_t)n * (2 + sizeof(SvtB200RateBlock) + 4) + tcb + 4096);
    if (rc) return rc;
    uint8_t *p = c->scratch;
    auto carve = [&](size_t b) { uint8_t *r = p; p += (b + 255) & ~(size_t)255; return r; };
    int16_t *d_q = (int16_t *)carve(qb);
    uint16_t *d_e = (uint16_t *)carve((size_t)n * 2);
    SvtB200RateBlock *d_b = (SvtB200RateBlock *)carve((size_t)n * sizeof(SvtB200RateBlock));
    uint32_t *d_tc = (uint32_t *)carve(tcb);
    int32_t *d_r = (int32_t *)carve((size_t)n * 4);
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(d_q, qcoeff, qb, cudaMemcpyHostToDevice, c->stream));
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(d_e, eob, (size_t)n * 2, cudaMemcpyHostToDevice, c->stream));
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(d_b, blocks, (size_t)n * sizeof(SvtB200RateBlock), cudaMemcpyHostToDevice, c->stream));
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(d_tc, token_costs, tcb, cudaMemcpyHostToDevice, c->stream));
    rc = svtb200_coeff_rate_batch_dev(c, tx_size, n, d_q, d_e, d_b, d_tc, d_r);
    if (rc) return rc;
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(rate, d_r, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    return SVTB200_OK;
}

} /* extern "C" */
