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
 * block is therefore a sum of independent per-position terms: one warp per block, lanes stride over the scan positions below
 * eob, the block's coefficients staged once in shared memory (every term gathers four of them), shuffle reduction at the end.
 * Integer arithmetic only: bit-exact by construction.
 */
#include "common.cuh"

#define RATE_THREADS 128
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

__device__ __forceinline__ int rate_token(const RateArgs &A, int v, int *cost) {
    if (v >= RATE_CAT6_MIN || v <= -RATE_CAT6_MIN) {
        const int e = (v < 0 ? -v : v) - RATE_CAT6_MIN;
        *cost = (int)__ldg(A.c6lo + (e & 0xff)) + (int)__ldg(A.c6hi + (e >> 8));
        return RATE_CAT6_TOKEN;
    }
    *cost = __ldg(A.vcost + v + 66);
    return __ldg(A.vtok + v + 66);
}
__device__ __forceinline__ int rate_energy(int tok) { return tok < 3 ? tok : tok < 5 ? 3 : tok < 7 ? 4 : 5; } /* eb_vp9_pt_energy_class */
__device__ __forceinline__ int rate_band(int c, int tx_size) { /* the band of scan position c (band_counts) */
    if (c < 10) return c == 0 ? 0 : c < 3 ? 1 : c < 6 ? 2 : 3;
    if (tx_size == 0) return c < 13 ? 4 : 5;
    return c < 21 ? 4 : 5;
}

/* one block by one warp; qs: the warp's staging buffer (N * N coefficients) */
template <int N>
__device__ __forceinline__ void rate_block(const RateArgs &A, int b, int16_t *qs, int lane) {
    constexpr int NN = N * N;
    const int eob = A.eob[b];
    const SvtB200RateBlock B = A.blk[b];
    const uint32_t *tc = A.tc + (size_t)((B.plane_type ? 1 : 0) * 2 + (B.is_inter ? 1 : 0)) * (6 * 2 * 6 * RATE_TOKENS);
    int cost = 0;
    if (eob == 0) {
        if (lane == 0) A.rate[b] = (int)tc[(0 * 2 + 0) * 6 * RATE_TOKENS + B.coeff_ctx * RATE_TOKENS + RATE_EOB_TOKEN];
        return;
    }
    const int tt = N == 32 ? 0 : (B.tx_type & 3);
    const int16_t *scan = A.scan[tt], *nb = A.nb[tt];
    /* long runs: stage the block (vector loads; every term gathers four coefficients); short runs, the common case after motion
     * compensation, gather the few coefficients they touch straight from global memory */
    const int16_t *q = A.qcoeff + (size_t)b * NN;
    const bool staged = NN <= 64 || eob > 48;
    if (staged) {
        if (NN >= 256) { for (int i = lane; i < NN / 8; i += 32) ((uint4 *)qs)[i] = __ldg((const uint4 *)q + i); }
        else { for (int i = lane; i < NN / 2; i += 32) ((uint32_t *)qs)[i] = __ldg((const uint32_t *)q + i); }
        __syncwarp();
    }
    const int16_t *src = staged ? qs : q;
    for (int c = lane; c < eob; c += 32) {
        int vc;
        const int tok = rate_token(A, src[__ldg(scan + c)], &vc);
        int ctx, skip;
        if (c == 0) { ctx = B.coeff_ctx; skip = 0; }
        else {
            int dummy;
            const int e0 = rate_energy(rate_token(A, src[__ldg(nb + 2 * c)], &dummy));
            const int e1 = rate_energy(rate_token(A, src[__ldg(nb + 2 * c + 1)], &dummy));
            ctx = (1 + e0 + e1) >> 1;
            skip = src[__ldg(scan + c - 1)] == 0; /* previous token ZERO_TOKEN <=> previous value 0 */
        }
        cost += vc + (int)__ldg(tc + ((rate_band(c, A.tx_size) * 2 + skip) * 6 + ctx) * RATE_TOKENS + tok);
    }
    if (lane == 0 && eob < NN) { /* the end-of-block token, unless the last position of the block was coded */
        int dummy;
        const int e0 = rate_energy(rate_token(A, src[__ldg(nb + 2 * eob)], &dummy));
        const int e1 = rate_energy(rate_token(A, src[__ldg(nb + 2 * eob + 1)], &dummy));
        cost += (int)__ldg(tc + ((rate_band(eob, A.tx_size) * 2 + 0) * 6 + ((1 + e0 + e1) >> 1)) * RATE_TOKENS + RATE_EOB_TOKEN);
    }
    for (int o = 16; o > 0; o >>= 1) cost += __shfl_xor_sync(0xffffffffu, cost, o);
    if (lane == 0) A.rate[b] = cost;
    __syncwarp();
}

template <int N>
__global__ void __launch_bounds__(RATE_THREADS) rate_kernel(const __grid_constant__ RateArgs A) {
    constexpr int WARPS = RATE_THREADS / 32;
    __shared__ __align__(16) int16_t qs[WARPS][N * N];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int b = blockIdx.x * WARPS + warp; b < A.n; b += gridDim.x * WARPS) rate_block<N>(A, b, qs[warp], lane);
}

/* several batches in one launch: the blocks of all batches form one index space, a warp takes every (grid warps)-th */
#define RATE_MAX_BATCHES 8
struct RateMultiArgs {
    RateArgs b[RATE_MAX_BATCHES];
    int32_t  end[RATE_MAX_BATCHES]; /* running block count */
    int32_t  nb;
};
__global__ void __launch_bounds__(RATE_THREADS) rate_multi_kernel(const __grid_constant__ RateMultiArgs M) {
    constexpr int WARPS = RATE_THREADS / 32;
    __shared__ __align__(16) int16_t qs[WARPS][32 * 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, total = M.end[M.nb - 1];
    for (int g = blockIdx.x * WARPS + warp; g < total; g += gridDim.x * WARPS) {
        int k = 0;
#pragma unroll
        for (int i = 1; i < RATE_MAX_BATCHES; i++) if (i < M.nb && g >= M.end[i - 1]) k = i;
        const RateArgs &A = M.b[k];
        const int b = g - (k ? M.end[k - 1] : 0);
        switch (A.tx_size) { /* warp-uniform */
        case 0: rate_block<4>(A, b, qs[warp], lane); break;
        case 1: rate_block<8>(A, b, qs[warp], lane); break;
        case 2: rate_block<16>(A, b, qs[warp], lane); break;
        default: rate_block<32>(A, b, qs[warp], lane); break;
        }
    }
}

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
    const int warps = RATE_THREADS / 32;
    int grid = (total + warps - 1) / warps;
    const int cap = (c->sm_count > 0 ? c->sm_count : 148) * 8;
    if (grid > cap) grid = cap;
    rate_multi_kernel<<<grid, RATE_THREADS, 0, c->stream>>>(M);
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
    RateArgs A;
    const int frc = rate_fill(c, A, tx_size, n, qcoeff, eob, blocks, token_costs, rate);
    if (frc) return frc;
    const int warps = RATE_THREADS / 32;
    int grid = (n + warps - 1) / warps;
    const int cap = (c->sm_count > 0 ? c->sm_count : 148) * 16;
    if (grid > cap) grid = cap;
    switch (tx_size) {
    case 0: rate_kernel<4><<<grid, RATE_THREADS, 0, c->stream>>>(A); break;
    case 1: rate_kernel<8><<<grid, RATE_THREADS, 0, c->stream>>>(A); break;
    case 2: rate_kernel<16><<<grid, RATE_THREADS, 0, c->stream>>>(A); break;
    default: rate_kernel<32><<<grid, RATE_THREADS, 0, c->stream>>>(A); break;
    }
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
