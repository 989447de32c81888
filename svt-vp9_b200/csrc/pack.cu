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
 * Two launches: an exclusive scan of the (even-padded) eob values on one CTA, then one warp per block gathering
 * qcoeff[scan[c]] and storing 32-bit pairs (128 contiguous bytes per warp store).  stream / offsets may be device memory or
 * page-locked host memory mapped into the device's address space: in the second case the kernel's stores ARE the device-to-host
 * transfer (no staging copy whose size the host would first have to learn).
 */
#include "common.cuh"

#define PACK_MAX_BATCHES 8
#define PACK_SCAN_THREADS 1024
#define PACK_THREADS 256

struct PackBatch {
    const SvtB200TxJob *jobs;
    const int16_t      *qcoeff;
    const uint16_t     *eob;
    const int16_t      *scan[4];
    int32_t             first, n, nn; /* index of the batch's first block, blocks, coefficients per block */
};
struct PackArgs {
    PackBatch b[PACK_MAX_BATCHES];
    int32_t   nb, n_total;
    uint32_t *offsets;
    int16_t  *stream;
    uint32_t  capacity;
};

__device__ __forceinline__ uint32_t pack_len(const PackArgs &A, int t) { /* even-padded eob of block t */
    int k = 0;
#pragma unroll
    for (int i = 1; i < PACK_MAX_BATCHES; i++) if (i < A.nb && t >= A.b[i].first) k = i;
    return ((uint32_t)A.b[k].eob[t - A.b[k].first] + 1u) & ~1u;
}

__global__ void __launch_bounds__(PACK_SCAN_THREADS) pack_scan_kernel(const __grid_constant__ PackArgs A) {
    __shared__ uint32_t wsum[PACK_SCAN_THREADS / 32];
    const int t = threadIdx.x, per = (A.n_total + PACK_SCAN_THREADS - 1) / PACK_SCAN_THREADS;
    const int t0 = min(t * per, A.n_total), t1 = min(t0 + per, A.n_total);
    uint32_t s = 0;
    for (int i = t0; i < t1; i++) s += pack_len(A, i);
    uint32_t inc = s; /* inclusive scan over the CTA */
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o); if ((t & 31) >= o) inc += v; }
    if ((t & 31) == 31) wsum[t >> 5] = inc;
    __syncthreads();
    if (t < 32) {
        uint32_t w = wsum[t], wi = w;
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, wi, o); if (t >= o) wi += v; }
        wsum[t] = wi - w;
    }
    __syncthreads();
    uint32_t run = wsum[t >> 5] + inc - s;
    for (int i = t0; i < t1; i++) { A.offsets[i] = run; run += pack_len(A, i); }
    if (t == PACK_SCAN_THREADS - 1) A.offsets[A.n_total] = run;
}

__global__ void __launch_bounds__(PACK_THREADS) pack_kernel(const __grid_constant__ PackArgs A) {
    const int lane = threadIdx.x & 31, wpb = PACK_THREADS / 32;
    for (int t = blockIdx.x * wpb + (threadIdx.x >> 5); t < A.n_total; t += gridDim.x * wpb) {
        int k = 0;
#pragma unroll
        for (int i = 1; i < PACK_MAX_BATCHES; i++) if (i < A.nb && t >= A.b[i].first) k = i;
        const PackBatch &B = A.b[k];
        const int j = t - B.first, eob = B.eob[j];
        if (!eob) continue;
        const uint32_t off = A.offsets[t];
        if (off + ((eob + 1) & ~1) > A.capacity) continue; /* the caller sees offsets[n_total] > capacity */
        const int16_t *q = B.qcoeff + (size_t)j * B.nn, *scan = B.scan[B.jobs[j].tx_type & 3];
        uint32_t *out = (uint32_t *)(A.stream + off);
        for (int c = 2 * lane; c < eob; c += 64) {
            const uint32_t sc = __ldg((const uint32_t *)(scan + c)); /* two scan entries */
            const uint32_t lo = (uint16_t)q[sc & 0xffff], hi = (c + 1 < eob) ? (uint16_t)q[sc >> 16] : 0u;
            out[c >> 1] = lo | (hi << 16);
        }
    }
}

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
    pack_scan_kernel<<<1, PACK_SCAN_THREADS, 0, c->stream>>>(A);
    c->launches++;
    if (total) {
        int grid = (total + PACK_THREADS / 32 - 1) / (PACK_THREADS / 32);
        const int cap = (c->sm_count > 0 ? c->sm_count : 148) * 8;
        if (grid > cap) grid = cap;
        pack_kernel<<<grid, PACK_THREADS, 0, c->stream>>>(A);
        c->launches++;
    }
    SVTB200_CUDA_CHECK(c, cudaGetLastError());
    return SVTB200_OK;
}
