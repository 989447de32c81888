/*
 * pack.cu -- the quantised coefficients of a picture's transform blocks as the entropy coder reads them: for every block the
 * first eob coefficients in scan order, back to back.
 *
 * The reference's encode pass leaves N*N int16 per transform block (EbEncDecProcess.c:365-587) and its tokeniser walks
 * qcoeff[scan[c]] for c < eob only (tokenize_b, Source/Lib/VPX/vp9_tokenize.c:275, loop :316-336; everything past eob is zero by
 * construction of eb_vp9_quantize_b, VPX/quantize.c:112 / :206).  A dense 4K picture is 24.9 MB of int16, the scan-order prefixes
 * a few MB: this is what has to cross PCIe when the coefficients are produced on the device and consumed by the host's entropy
 * coder.  Output format (the consumer needs nothing else but the eob array it already gets):
 *   offsets[t]        first coefficient of block t in the stream, blocks numbered batch after batch in the order given;
 *                     every block's run is padded to an even count with a zero, so runs start 4-byte aligned
 *   offsets[n_total]  total length in coefficients
 *   stream[]          int16 coefficients
 * Two launches over chunks of 64 consecutive blocks: the chunk totals of the (even-padded) eob values, turned into chunk bases by
 * the last CTA to finish; then per chunk a CTA scan for the block offsets and one warp per block gathering qcoeff[scan[c]] and
 * storing 32-bit pairs (128 contiguous bytes per warp store).  stream / offsets may be device memory or
 * page-locked host memory mapped into the device's address space: in the second case the kernel's stores ARE the device-to-host
 * transfer (no staging copy whose size the host would first have to learn).
 */
#include "common.cuh"

#include "pack_kernel.cuh"

extern "C" int svtb200_coeff_pack_dev(SvtB200Ctx *c, const SvtB200TxBatch *batches, int n_batches, int16_t *stream, uint32_t capacity,
                                      uint32_t *offsets) {
    if (!c || !batches || n_batches < 1 || n_batches > PACK_MAX_BATCHES || !stream || !offsets || ((uintptr_t)stream & 3)) return SVTB200_EINVAL;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    cudaSetDevice(c->device);
    PackArgs A;
    memset(&A, 0, sizeof(A));
    int total = 0;
    for (int k = 0; k < n_batches; k++) {
        const SvtB200TxBatch &b = batches[k];
        if (b.tx_size < 0 || b.tx_size > 3 || b.n < 0 || (b.n && (!b.jobs || !b.qcoeff || !b.eob))) return SVTB200_EINVAL;
        if (!b.n) continue;
        PackBatch &P = A.b[A.nb++];
        P.jobs = b.jobs; P.qcoeff = b.qcoeff; P.eob = b.eob; P.first = total; P.n = b.n; P.nn = (4 << b.tx_size) * (4 << b.tx_size);
        for (int t = 0; t < 4; t++) {
            P.scan[t] = c->scans[b.tx_size][b.tx_size == 3 ? 0 : t].scan;
            if (!P.scan[t]) { c->err = "scan tables not registered (svtb200_set_scan)"; return SVTB200_EINVAL; }
        }
        total += b.n;
    }
    A.n_total = total; A.offsets = offsets; A.stream = stream; A.capacity = capacity;
    const int chunks = total ? (total + PACK_CHUNK - 1) / PACK_CHUNK : 1;
    /* chunk bases + the ticket word: per compute queue (launches on different queues may be in flight together) */
    const int qi = svtb200_compute_index(c);
    if (c->pack_chunks[qi] < chunks) {
        if (c->pack_scratch[qi]) cudaFree(c->pack_scratch[qi]);
        c->pack_scratch[qi] = nullptr; c->pack_chunks[qi] = 0;
        const int cap = chunks + chunks / 2 + 64;
        if (cudaMalloc((void **)&c->pack_scratch[qi], (size_t)(cap + 1) * 4) != cudaSuccess) { cudaGetLastError(); c->err = "cudaMalloc failed"; return SVTB200_ENOMEM; }
        SVTB200_CUDA_CHECK(c, cudaMemsetAsync(c->pack_scratch[qi], 0, (size_t)(cap + 1) * 4, c->stream));
        c->pack_chunks[qi] = cap;
    }
    uint32_t *chunk_base = c->pack_scratch[qi] + 1, *ticket = c->pack_scratch[qi];
    pack_sum_kernel<<<chunks, PACK_THREADS, 0, c->stream>>>(A, chunk_base, ticket);
    c->launches++;
    if (total) {
        pack_kernel<<<chunks, PACK_THREADS, 0, c->stream>>>(A, chunk_base);
        c->launches++;
    }
    SVTB200_CUDA_CHECK(c, cudaGetLastError());
    return SVTB200_OK;
}
