/* rate_kernel.cuh -- device code of rate.cu (coefficient rate estimation, position-parallel); a header of its own so that
 * tests/emu/rate_emu.cpp can run the same source on the CPU. */
#ifndef SVTB200_RATE_KERNEL_CUH
#define SVTB200_RATE_KERNEL_CUH
#include <stdint.h>
#include <stddef.h>

#define RATE_THREADS 256
#define RATE_TOKENS 12
#define RATE_EOB_TOKEN 11
#define RATE_CAT6_TOKEN 10
#define RATE_CAT6_MIN 67

struct RateArgs {
    const int16_t           *qcoeff;
    const uint16_t          *eob;
    const SvtB200RateBlock  *blk;
    const uint32_t          *tc;       /* token_costs[tx_size]: [plane_type 2][ref 2][band 6][skip-eob 2][ctx 6][token 12] */
    const int16_t           *scan[4], *nb[4];
    const int16_t           *vtok;     /* 133 entries: token of value v at [v + 66] */
    const int32_t           *vcost;    /* 133 entries: cost of value v's token extra bits + sign */
    const int16_t           *c6lo;     /* 256 */
    const uint16_t          *c6hi;     /* 64 */
    int32_t                 *rate;
    int32_t                  n, tx_size;
};

/* the value -> token / cost tables of the small values, copied into shared memory once per CTA (every term looks three values up) */
struct RateTables { int32_t vcost[133]; int16_t vtok[134]; };
__device__ __forceinline__ void rate_load_tables(const RateArgs &A, RateTables &T) {
    for (int i = threadIdx.x; i < 133; i += blockDim.x) { T.vcost[i] = __ldg(A.vcost + i); T.vtok[i] = __ldg(A.vtok + i); }
    __syncthreads();
}
__device__ __forceinline__ int rate_token(const RateArgs &A, const RateTables &T, int v, int *cost) {
    if (v >= RATE_CAT6_MIN || v <= -RATE_CAT6_MIN) {
        const int e = (v < 0 ? -v : v) - RATE_CAT6_MIN;
        *cost = (int)__ldg(A.c6lo + (e & 0xff)) + (int)__ldg(A.c6hi + (e >> 8));
        return RATE_CAT6_TOKEN;
    }
    *cost = T.vcost[v + 66];
    return T.vtok[v + 66];
}
/* token only (the neighbours' energy classes) */
__device__ __forceinline__ int rate_token_only(const RateTables &T, int v) {
    return (v >= RATE_CAT6_MIN || v <= -RATE_CAT6_MIN) ? RATE_CAT6_TOKEN : T.vtok[v + 66];
}
__device__ __forceinline__ int rate_energy(int tok) { return tok < 3 ? tok : tok < 5 ? 3 : tok < 7 ? 4 : 5; } /* eb_vp9_pt_energy_class */
__device__ __forceinline__ int rate_band(int c, int tx_size) { /* the band of scan position c (band_counts) */
    if (c < 10) return c == 0 ? 0 : c < 3 ? 1 : c < 6 ? 2 : 3;
    if (tx_size == 0) return c < 13 ? 4 : 5;
    return c < 21 ? 4 : 5;
}

/* The blocks of all batches form one index space, cut into chunks of RATE_CHUNK consecutive blocks, one CTA each.  The first
 * RATE_CHUNK threads fetch their block's descriptor (eob, coefficient / scan / neighbour / cost-table pointers, context) into shared
 * memory and add the end-of-block term; then the scan POSITIONS of the whole chunk are one index space walked by all threads (a
 * thread finds the block of its position by bisection over the running eob sums): no lane idles on a short block, and every
 * position is an independent chain of loads.  The terms of a block are summed in shared memory (one atomic per warp when the
 * warp's positions all belong to one block, which is the common case). */
#ifndef RATE_CHUNK
#define RATE_CHUNK 32 /* 64 left two thirds of the warp slots empty at 4K (389 CTAs on 148 SMs): the kernel is a chain of dependent loads per position */
#endif
#define RATE_MAX_BATCHES 8
struct RateMultiArgs {
    RateArgs b[RATE_MAX_BATCHES];
    int32_t  end[RATE_MAX_BATCHES]; /* running block count */
    int32_t  nb;
};
struct RateBlockS { /* one block's working set */
    const int16_t  *q, *scan, *nb;
    const uint32_t *tc;  /* cost table of the block's (plane type, reference type) */
    int32_t         eob, ctx, tx_size, nn;
};
__global__ void __launch_bounds__(RATE_THREADS) rate_multi_kernel(const __grid_constant__ RateMultiArgs M) {
    __shared__ RateTables T;
    __shared__ RateBlockS S[RATE_CHUNK];
    __shared__ uint32_t pos_s[RATE_CHUNK + 1]; /* exclusive running sum of eob over the chunk */
    __shared__ int32_t rate_s[RATE_CHUNK];
    __shared__ uint32_t wsum[RATE_THREADS / 32];
    /* block k of this CTA is g = blockIdx.x + k * gridDim.x: neighbouring blocks (which tend to be equally busy) go to different CTAs */
    const int tid = threadIdx.x, lane = tid & 31, total = M.end[M.nb - 1];
    rate_load_tables(M.b[0], T); /* the value tables are the context's: the same for every batch */
    const RateArgs &A0 = M.b[0];
    {
        const int g = (int)blockIdx.x + tid * (int)gridDim.x;
        const bool mine = tid < RATE_CHUNK && g < total;
        uint32_t len = 0;
        if (mine) {
            int k = 0;
#pragma unroll
            for (int i = 1; i < RATE_MAX_BATCHES; i++) if (i < M.nb && g >= M.end[i - 1]) k = i;
            const RateArgs &A = M.b[k];
            const int b = g - (k ? M.end[k - 1] : 0), nn = (4 << A.tx_size) * (4 << A.tx_size);
            const SvtB200RateBlock B = A.blk[b];
            const int tt = A.tx_size == 3 ? 0 : (B.tx_type & 3), eob = A.eob[b];
            RateBlockS s;
            s.q = A.qcoeff + (size_t)b * nn; s.scan = A.scan[tt]; s.nb = A.nb[tt];
            s.tc = A.tc + (size_t)((B.plane_type ? 1 : 0) * 2 + (B.is_inter ? 1 : 0)) * (6 * 2 * 6 * RATE_TOKENS);
            s.eob = eob; s.ctx = B.coeff_ctx; s.tx_size = A.tx_size; s.nn = nn;
            S[tid] = s;
            len = (uint32_t)eob;
            /* the end-of-block token: the only term of an empty block, absent when the last position of the block was coded */
            int r = 0;
            if (eob == 0) r = (int)__ldg(s.tc + (0 * 2 + 0) * 6 * RATE_TOKENS + B.coeff_ctx * RATE_TOKENS + RATE_EOB_TOKEN);
            else if (eob < nn) {
                const int e0 = rate_energy(rate_token_only(T, s.q[__ldg(s.nb + 2 * eob)]));
                const int e1 = rate_energy(rate_token_only(T, s.q[__ldg(s.nb + 2 * eob + 1)]));
                r = (int)__ldg(s.tc + ((rate_band(eob, A.tx_size) * 2 + 0) * 6 + ((1 + e0 + e1) >> 1)) * RATE_TOKENS + RATE_EOB_TOKEN);
            }
            rate_s[tid] = r;
        }
        /* exclusive scan of the eob values over the first RATE_CHUNK threads */
        uint32_t inc = len;
        for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += x; }
        if (lane == 31) wsum[tid >> 5] = inc;
        __syncthreads();
        if (tid < RATE_CHUNK) {
            uint32_t base = 0;
            for (int w = 0; w < (tid >> 5); w++) base += wsum[w];
            pos_s[tid] = base + inc - len;
            if (tid == RATE_CHUNK - 1) pos_s[RATE_CHUNK] = base + inc;
        }
    }
    __syncthreads();
    const uint32_t npos = pos_s[RATE_CHUNK];
    for (uint32_t p0 = 0; p0 < npos; p0 += RATE_THREADS) {
        const uint32_t p = p0 + tid;
        const bool act = p < npos;
        int i = 0, cost = 0, v = 0, cpos = 0;
        if (act) {
#pragma unroll
            for (int step = RATE_CHUNK / 2; step > 0; step >>= 1) if (pos_s[i + step] <= p) i += step;
            const RateBlockS &s = S[i];
            const int c = (int)(p - pos_s[i]);
            v = s.q[__ldg(s.scan + c)];
            cpos = c;
        }
        /* the previous coefficient in scan order sits in the lane below when that lane works on the same block */
        const int v_below = __shfl_up_sync(0xffffffffu, v, 1), i_below = __shfl_up_sync(0xffffffffu, i, 1);
        if (act) {
            const RateBlockS &s = S[i];
            const int c = cpos;
            int vc;
            const int tok = rate_token(A0, T, v, &vc);
            int ctx, skip;
            if (c == 0) { ctx = s.ctx; skip = 0; }
            else {
                const uint32_t nbp = __ldg((const uint32_t *)(s.nb + 2 * c)); /* both neighbour positions */
                const int e0 = rate_energy(rate_token_only(T, s.q[nbp & 0xffff]));
                const int e1 = rate_energy(rate_token_only(T, s.q[nbp >> 16]));
                ctx = (1 + e0 + e1) >> 1;
                const int prev = (lane > 0 && i_below == i) ? v_below : s.q[__ldg(s.scan + c - 1)];
                skip = prev == 0; /* previous token ZERO_TOKEN <=> previous value 0 */
            }
            cost = vc + (int)__ldg(s.tc + ((rate_band(c, s.tx_size) * 2 + skip) * 6 + ctx) * RATE_TOKENS + tok);
        }
        /* sum per block: one atomic per warp when all its (active) lanes work on the same block */
        const unsigned m = __ballot_sync(0xffffffffu, act);
        if (m) {
            const int lead = __ffs(m) - 1, i0 = __shfl_sync(0xffffffffu, i, lead);
            if (__all_sync(0xffffffffu, !act || i == i0)) {
                for (int o = 16; o > 0; o >>= 1) cost += __shfl_xor_sync(0xffffffffu, cost, o);
                if (lane == lead) atomicAdd(&rate_s[i0], cost);
            } else if (act) atomicAdd(&rate_s[i], cost);
        }
    }
    __syncthreads();
    {
        const int g = (int)blockIdx.x + tid * (int)gridDim.x;
        if (tid < RATE_CHUNK && g < total) {
            int k = 0;
#pragma unroll
            for (int i = 1; i < RATE_MAX_BATCHES; i++) if (i < M.nb && g >= M.end[i - 1]) k = i;
            M.b[k].rate[g - (k ? M.end[k - 1] : 0)] = rate_s[tid];
        }
    }
}
#endif
