/* pack_kernel.cuh -- device code of pack.cu (scan-order packing of the quantised coefficients); a header of its own so that
 * tests/emu/rate_emu.cpp can run the same source on the CPU. */
#ifndef SVTB200_PACK_KERNEL_CUH
#define SVTB200_PACK_KERNEL_CUH
#include <stdint.h>
#include <stddef.h>

#define PACK_MAX_BATCHES 8
#define PACK_THREADS 256
#ifndef PACK_CHUNK
#define PACK_CHUNK 32 /* blocks per CTA: 4 per warp, so that the dependent loads of a block (scan entry -> coefficient) overlap across many warps; 64 gave 389 CTAs at 4K: a third of the warp slots */
#endif

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

/* block t of the concatenated batches -> its batch */
__device__ __forceinline__ int pack_batch_of(const PackArgs &A, int t) {
    int k = 0;
#pragma unroll
    for (int i = 1; i < PACK_MAX_BATCHES; i++) if (i < A.nb && t >= A.b[i].first) k = i;
    return k;
}
__device__ __forceinline__ uint32_t pack_len(const PackArgs &A, int t) { /* even-padded eob of block t */
    const int k = pack_batch_of(A, t);
    return ((uint32_t)A.b[k].eob[t - A.b[k].first] + 1u) & ~1u;
}
/* inclusive scan over the CTA's PACK_THREADS values; every thread gets its own inclusive value, *total the CTA's sum */
__device__ __forceinline__ uint32_t pack_cta_scan(uint32_t v, uint32_t *wsum, uint32_t *total) {
    const int t = threadIdx.x;
    uint32_t inc = v;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, inc, o); if ((t & 31) >= o) inc += x; }
    __syncthreads(); /* wsum from the previous use */
    if ((t & 31) == 31) wsum[t >> 5] = inc;
    __syncthreads();
    uint32_t base = 0, all = 0;
#pragma unroll
    for (int w = 0; w < PACK_THREADS / 32; w++) { const uint32_t x = wsum[w]; if (w < (t >> 5)) base += x; all += x; }
    *total = all;
    return base + inc;
}

/* pass 1: one CTA per PACK_CHUNK consecutive blocks: the chunk's total length; the LAST CTA to finish (ticket counter) turns the
 * chunk totals into exclusive chunk bases and writes the stream length -- no second tiny launch, no serial walk over 24 000 blocks */
__global__ void __launch_bounds__(PACK_THREADS) pack_sum_kernel(const __grid_constant__ PackArgs A, uint32_t *chunk_base, uint32_t *ticket) {
    __shared__ uint32_t wsum[PACK_THREADS / 32];
    __shared__ uint32_t last;
    const int t = blockIdx.x * PACK_CHUNK + threadIdx.x;
    const bool mine = threadIdx.x < PACK_CHUNK && t < A.n_total;
    const uint32_t len = mine ? pack_len(A, t) : 0u;
    uint32_t total;
    const uint32_t inc = pack_cta_scan(len, wsum, &total);
    if (mine) A.offsets[t] = inc - len; /* relative to the chunk base for now; pass 2 makes it absolute */
    if (threadIdx.x == 0) {
        chunk_base[blockIdx.x] = total;
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    /* exclusive scan of the chunk totals, PACK_THREADS at a time */
    uint32_t run = 0;
    for (int c0 = 0; c0 < (int)gridDim.x; c0 += PACK_THREADS) {
        const int c = c0 + threadIdx.x;
        const uint32_t v = c < (int)gridDim.x ? ((volatile uint32_t *)chunk_base)[c] : 0u;
        uint32_t tot;
        const uint32_t inc = pack_cta_scan(v, wsum, &tot);
        if (c < (int)gridDim.x) chunk_base[c] = run + inc - v;
        run += tot;
    }
    if (threadIdx.x == 0) { A.offsets[A.n_total] = run; *ticket = 0; }
}

/* pass 2: the same chunks.  The first PACK_CHUNK threads fetch what is needed of their block (eob, coefficient and scan pointers)
 * and scan the lengths on top of the chunk base; then the chunk's coefficient PAIRS are one index space walked by all threads:
 * a thread finds the block of its pair by bisection in shared memory, gathers qcoeff[scan[c]], qcoeff[scan[c + 1]] and stores one
 * word -- consecutive threads store consecutive words, no lane idles on a short block, and every pair is an independent chain
 * of two loads */
__global__ void __launch_bounds__(PACK_THREADS) pack_kernel(const __grid_constant__ PackArgs A, const uint32_t *chunk_base) {
    __shared__ uint32_t wsum[PACK_THREADS / 32];
    __shared__ uint32_t off_s[PACK_CHUNK + 1]; /* relative to the chunk base, in coefficients */
    __shared__ const int16_t *q_s[PACK_CHUNK], *scan_s[PACK_CHUNK];
    __shared__ uint16_t eob_s[PACK_CHUNK];
    /* this CTA's blocks are t = blockIdx.x + k * gridDim.x (neighbouring blocks, which tend to be equally busy, go to different
     * CTAs); out_s[k]: where block k's run starts in the stream, off_s[]: running pair count of the CTA's own blocks */
    __shared__ uint32_t out_s[PACK_CHUNK];
    const int tid = threadIdx.x;
    {
        const int t = (int)blockIdx.x + tid * (int)gridDim.x;
        const bool mine = tid < PACK_CHUNK && t < A.n_total;
        uint32_t len = 0;
        if (mine) {
            const PackBatch &B = A.b[pack_batch_of(A, t)];
            const int j = t - B.first, eob = B.eob[j];
            len = ((uint32_t)eob + 1u) & ~1u;
            eob_s[tid] = (uint16_t)eob;
            q_s[tid] = B.qcoeff + (size_t)j * B.nn;
            scan_s[tid] = B.scan[B.jobs[j].tx_type & 3];
            const uint32_t off = chunk_base[t / PACK_CHUNK] + A.offsets[t];
            out_s[tid] = off;
            A.offsets[t] = off;
        }
        uint32_t total;
        const uint32_t inc = pack_cta_scan(len, wsum, &total);
        if (tid < PACK_CHUNK) off_s[tid] = inc - len;
        if (tid == 0) off_s[PACK_CHUNK] = total;
    }
    __syncthreads();
    const uint32_t pairs = off_s[PACK_CHUNK] >> 1;
    for (uint32_t p = tid; p < pairs; p += PACK_THREADS) {
        const uint32_t pos = 2 * p;
        int i = 0; /* the block with off_s[i] <= pos < off_s[i + 1] (empty blocks have equal bounds and are never selected) */
#pragma unroll
        for (int step = PACK_CHUNK / 2; step > 0; step >>= 1) if (off_s[i + step] <= pos) i += step;
        const uint32_t c = pos - off_s[i], eob = eob_s[i];
        if (out_s[i] + ((eob + 1u) & ~1u) > A.capacity) continue; /* the caller sees offsets[n_total] > capacity */
        const uint32_t sc = __ldg((const uint32_t *)(scan_s[i] + c)); /* two scan entries */
        const int16_t *q = q_s[i];
        const uint32_t lo = (uint16_t)q[sc & 0xffff], hi = (c + 1 < eob) ? (uint16_t)q[sc >> 16] : 0u;
        ((uint32_t *)A.stream)[(out_s[i] + c) >> 1] = lo | (hi << 16);
    }
}
#endif
