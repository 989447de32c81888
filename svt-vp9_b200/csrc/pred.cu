/* pred.cu -- prediction of the encode pass and full-loop distortion (the calls either side of every transform call):
 *   svtb200_inter_predict_batch        inter_predictor() as build_inter_predictors() calls it (Source/Lib/VPX/vp9_reconinter.c:102-250,
 *                                      vp9_reconinter.h:23): the regular 8-tap kernel, unscaled, dispatched like scale_factors::predict
 *                                      (vp9_scale.c:78-128) on eb_vp9_convolve_copy / convolve8_horiz / _vert / convolve8 and their
 *                                      _avg twins (Source/Lib/VPX/vpx_convolve.c:101-215)
 *   svtb200_intra_predict_batch        pred[mode][tx_size] / dc_pred[..][..][tx_size] of build_intra_predictors (Source/Lib/VPX/vp9_reconintra.c:397-407)
 *                                      = the predictors of Source/Lib/VPX/intrapred.c, evaluated per sample (intra_device.cuh)
 *   svtb200_full_distortion_batch_dev  full_distortion_kernel32bit / _eob_zero32bit (Source/Lib/C_DEFAULT/EbPictureOperators_C.c:289,314)
 *                                      behind the table lookup of Source/Lib/Codec/EbEncDecProcess.c:723
 * One CTA per block: the (w + 7) x (h + 7) reference window is staged once in shared memory (only the rows / columns the
 * reference functions read are touched), the horizontal pass writes the 8-bit intermediate the reference keeps in 'temp', the
 * vertical pass reads it; a compound block runs both references through the same buffers and averages in registers. */
#include "common.cuh"
#include "intra_device.cuh"

#define PRED_THREADS 128
#define PRED_WP 80 /* window pitch: sample x of the block sits at byte x + 4, so that 4-sample groups are word aligned */
#define PRED_ROWS 73 /* 64 + 7 staged rows, + slack the 2-row groups of the vertical pass may touch */

/* the normative VP9 regular 8-tap filter (sub_pel_filters_8, Source/Lib/VPX/vp9_filter.c:32); phase 0 (a copy) never filters */
__constant__ __align__(8) int8_t c_pred_taps[16][8] = {
    {0, 0, 0, 127, 0, 0, 0, 0},
    {0, 1, -5, 126, 8, -3, 1, 0},      {-1, 3, -10, 122, 18, -6, 2, 0},   {-1, 4, -13, 118, 27, -9, 3, -1},
    {-1, 4, -16, 112, 37, -11, 4, -1}, {-1, 5, -18, 105, 48, -14, 4, -1}, {-1, 5, -19, 97, 58, -16, 5, -1},
    {-1, 6, -19, 88, 68, -18, 5, -1},  {-1, 6, -19, 78, 78, -19, 6, -1},  {-1, 5, -18, 68, 88, -19, 6, -1},
    {-1, 5, -16, 58, 97, -19, 5, -1},  {-1, 4, -14, 48, 105, -18, 5, -1}, {-1, 4, -11, 37, 112, -16, 4, -1},
    {-1, 3, -9, 27, 118, -13, 4, -1},  {0, 2, -6, 18, 122, -10, 3, -1},   {0, 1, -3, 8, 126, -5, 1, 0}};

__device__ __forceinline__ int pred_dp4a_us(uint32_t a /* 4 unsigned bytes */, uint32_t b /* 4 signed bytes */, int c) {
    int r;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
/* i / d for i < 5760, d <= 71 with rcp = ceil(2^32 / d) (exact over that range) */
__device__ __forceinline__ int pred_div(int i, uint32_t rcp) { return rcp ? (int)__umulhi((uint32_t)i, rcp) : i; }
__device__ __forceinline__ uint32_t pred_rcp(int d) { return d == 1 ? 0u : 0xffffffffu / (uint32_t)d + 1u; } /* 0: d == 1 */
__device__ __forceinline__ uint32_t pred_clip8(int v) { return (uint32_t)(v < 0 ? 0 : v > 255 ? 255 : v); }

/* one reference of the block -> res (pitch 64, shared).  Horizontal pass: four adjacent outputs per thread from three words
 * of the row, two 4-tap dot products each.  Vertical pass: a 4 x 2 tile per thread from nine words (4 columns each); even
 * and odd columns ride in the two 16-bit halves of a register (the sums are biased by 16384 so that a half never borrows from
 * its neighbour: 255 x the taps' negative / positive sums stays inside -10200 .. 42840). */
__device__ static void pred_one(const uint8_t *src /* at the block's integer position */, int stride, int sx, int sy, int w, int h,
                                uint8_t *win, uint8_t *tmp, uint8_t *res) {
    const int t = threadIdx.x;
    if (!sx && !sy) { /* eb_vp9_convolve_copy_c */
        const uint32_t rw = pred_rcp(w);
        for (int i = t; i < w * h; i += PRED_THREADS) {
            const int r = pred_div(i, rw), c = i - r * w;
            res[r * 64 + c] = __ldg(src + (ptrdiff_t)r * stride + c);
        }
        __syncthreads();
        return;
    }
    const int x_lo = sx ? -3 : 0, nx = sx ? w + 7 : w;
    const int y_lo = sy ? -3 : 0, ny = sy ? h + 7 : h;
    const uint32_t rnx = pred_rcp(nx);
    for (int i = t; i < nx * ny; i += PRED_THREADS) { /* only the samples the reference functions read */
        const int r = pred_div(i, rnx), c = i - r * nx;
        win[r * PRED_WP + c + x_lo + 4] = __ldg(src + (ptrdiff_t)(r + y_lo) * stride + c + x_lo);
    }
    __syncthreads();
    const int wq = (w + 3) >> 2;
    const uint32_t rwq = pred_rcp(wq);
    const uint8_t *v_in = win + 4;
    int v_pitch = PRED_WP;
    if (sx) { /* over every staged row: h, or h + 7 when a vertical pass follows (vpx_convolve.c:153-162) */
        const uint32_t K0 = ((const uint32_t *)c_pred_taps[sx])[0], K1 = ((const uint32_t *)c_pred_taps[sx])[1];
        uint8_t *out = sy ? tmp : res;
        for (int u = t; u < wq * ny; u += PRED_THREADS) {
            const int r = pred_div(u, rwq), c4 = (u - r * wq) * 4;
            const uint32_t *p = (const uint32_t *)(win + r * PRED_WP + c4); /* samples c4 - 4 .. c4 + 7 */
            const uint32_t W0 = p[0], W1 = p[1], W2 = p[2];
            const uint32_t v0 = pred_clip8(pred_dp4a_us(__funnelshift_r(W0, W1, 8), K0, pred_dp4a_us(__funnelshift_r(W1, W2, 8), K1, 64)) >> 7);
            const uint32_t v1 = pred_clip8(pred_dp4a_us(__funnelshift_r(W0, W1, 16), K0, pred_dp4a_us(__funnelshift_r(W1, W2, 16), K1, 64)) >> 7);
            const uint32_t v2 = pred_clip8(pred_dp4a_us(__funnelshift_r(W0, W1, 24), K0, pred_dp4a_us(__funnelshift_r(W1, W2, 24), K1, 64)) >> 7);
            const uint32_t v3 = pred_clip8(pred_dp4a_us(W1, K0, pred_dp4a_us(W2, K1, 64)) >> 7);
            *(uint32_t *)(out + r * 64 + c4) = v0 | v1 << 8 | v2 << 16 | v3 << 24;
        }
        __syncthreads();
        if (!sy) return;
        v_in = tmp; v_pitch = 64;
    }
    {
        int tap[8];
#pragma unroll
        for (int k = 0; k < 8; k++) tap[k] = c_pred_taps[sy][k];
        const int hq = (h + 1) >> 1;
        for (int u = t; u < wq * hq; u += PRED_THREADS) {
            const int rq = pred_div(u, rwq), c4 = (u - rq * wq) * 4, r0 = rq * 2;
            uint32_t E[9], O[9];
#pragma unroll
            for (int k = 0; k < 9; k++) {
                const uint32_t W = *(const uint32_t *)(v_in + (r0 + k) * v_pitch + c4);
                E[k] = W & 0x00ff00ffu; O[k] = (W >> 8) & 0x00ff00ffu;
            }
#pragma unroll
            for (int i = 0; i < 2; i++) {
                uint32_t aE = 0x40404040u, aO = 0x40404040u; /* per half: 16384 + the rounding 64 */
#pragma unroll
                for (int k = 0; k < 8; k++) { aE += E[i + k] * (uint32_t)tap[k]; aO += O[i + k] * (uint32_t)tap[k]; }
                /* per half: ((sum + 64) >> 7) + 128, clamped to 128 .. 383, minus 128 */
                uint32_t qE = (aE >> 7) & 0x01ff01ffu, qO = (aO >> 7) & 0x01ff01ffu;
                qE = __vminu2(__vmaxu2(qE, 0x00800080u), 0x017f017fu) - 0x00800080u;
                qO = __vminu2(__vmaxu2(qO, 0x00800080u), 0x017f017fu) - 0x00800080u;
                *(uint32_t *)(res + (r0 + i) * 64 + c4) = qE | qO << 8;
            }
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(PRED_THREADS) pred_kernel(const SvtB200PredBlock *blocks, int n) {
    __shared__ __align__(16) uint8_t win[PRED_ROWS * PRED_WP];
    __shared__ __align__(16) uint8_t tmp[PRED_ROWS * 64];
    __shared__ __align__(16) uint8_t res[2][68 * 64];
    for (int b = blockIdx.x; b < n; b += gridDim.x) {
        const SvtB200PredBlock B = blocks[b];
        pred_one(B.ref[0] + (ptrdiff_t)B.y[0] * B.ref_stride[0] + B.x[0], B.ref_stride[0], B.subpel_x[0], B.subpel_y[0], B.w, B.h, win, tmp, res[0]);
        if (B.compound)
            pred_one(B.ref[1] + (ptrdiff_t)B.y[1] * B.ref_stride[1] + B.x[1], B.ref_stride[1], B.subpel_x[1], B.subpel_y[1], B.w, B.h, win, tmp, res[1]);
        if (((uintptr_t)B.dst & 3) == 0 && (B.dst_stride & 3) == 0 && (B.w & 3) == 0) {
            const int wq = B.w >> 2;
            const uint32_t rwq = pred_rcp(wq);
            for (int i = threadIdx.x; i < wq * B.h; i += PRED_THREADS) {
                const int r = pred_div(i, rwq), c4 = (i - r * wq) * 4;
                uint32_t v = *(const uint32_t *)(res[0] + r * 64 + c4);
                if (B.compound) v = __vavgu4(v, *(const uint32_t *)(res[1] + r * 64 + c4)); /* (a + b + 1) >> 1: eb_vp9_convolve_avg_c, vpx_convolve.c:211 */
                *(uint32_t *)(B.dst + (ptrdiff_t)r * B.dst_stride + c4) = v;
            }
        } else {
            for (int i = threadIdx.x; i < B.w * B.h; i += PRED_THREADS) {
                const int r = i / B.w, c = i - r * B.w;
                int v = res[0][r * 64 + c];
                if (B.compound) v = (v + res[1][r * 64 + c] + 1) >> 1;
                B.dst[(ptrdiff_t)r * B.dst_stride + c] = (uint8_t)v;
            }
        }
        __syncthreads();
    }
}

/* Descriptors of a batch go to a per-context device buffer.  The previous batch's kernel (possibly on another compute queue)
 * may still be reading it: the copy is ordered behind pred_read, which the caller records after its launch. */
static void *pred_upload_desc(SvtB200Ctx *c, const void *host, size_t bytes) {
    if (c->pred_read) { if (cudaStreamWaitEvent(c->stream, c->pred_read, 0) != cudaSuccess) { c->err = "cudaStreamWaitEvent failed"; return nullptr; } }
    else if (cudaEventCreateWithFlags(&c->pred_read, cudaEventDisableTiming) != cudaSuccess) { c->err = "cudaEventCreate failed"; return nullptr; }
    if (c->pred_desc_bytes < bytes) cudaEventSynchronize(c->pred_read); /* about to free the buffer */
    c->err.clear();
    if (svtb200_grow(c, &c->pred_desc, &c->pred_desc_bytes, bytes)) return nullptr;
    if (cudaMemcpyAsync(c->pred_desc, host, bytes, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { c->err = "descriptor copy failed"; return nullptr; } /* pageable source: staged before it returns */
    return c->pred_desc;
}

extern "C" int svtb200_inter_predict_batch_dev(SvtB200Ctx *c, const SvtB200PredBlock *blocks_dev, int n) {
    if (!c || n < 0 || (n && !blocks_dev)) return SVTB200_EINVAL;
    if (!n) return SVTB200_OK;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    const int cap = (c->sm_count > 0 ? c->sm_count : 148) * 10; /* 20 KB of shared memory per CTA: ten resident per SM */
    pred_kernel<<<n < cap ? n : cap, PRED_THREADS, 0, c->stream>>>(blocks_dev, n);
    c->launches++;
    SVTB200_CUDA_CHECK(c, cudaGetLastError());
    return SVTB200_OK;
}

extern "C" int svtb200_inter_predict_batch(SvtB200Ctx *c, const SvtB200PredBlock *blocks, int n) {
    if (!c || n < 0 || (n && !blocks)) return SVTB200_EINVAL;
    if (!n) return SVTB200_OK;
    for (int i = 0; i < n; i++) {
        const SvtB200PredBlock &b = blocks[i];
        if (!b.ref[0] || !b.dst || b.w < 4 || b.w > 64 || b.h < 4 || b.h > 64 || b.subpel_x[0] > 15 || b.subpel_y[0] > 15 || b.compound > 1) return SVTB200_EINVAL;
        if (b.compound && (!b.ref[1] || b.subpel_x[1] > 15 || b.subpel_y[1] > 15)) return SVTB200_EINVAL;
    }
    std::lock_guard<std::recursive_mutex> g(c->mu);
    SvtB200PredBlock *d = (SvtB200PredBlock *)pred_upload_desc(c, blocks, (size_t)n * sizeof(SvtB200PredBlock));
    if (!d) return c->err.empty() ? SVTB200_ENOMEM : SVTB200_ECUDA;
    int rc = svtb200_inter_predict_batch_dev(c, d, n);
    if (rc) return rc;
    SVTB200_CUDA_CHECK(c, cudaEventRecord(c->pred_read, c->stream));
    return SVTB200_OK;
}

/* ---- intra prediction: one CTA per block, closed-form predictors (intra_device.cuh), four samples per thread ---- */
#define INTRA_THREADS 128
__global__ void __launch_bounds__(INTRA_THREADS) intra_kernel(const SvtB200IntraBlock *blocks, int n) {
    __shared__ __align__(16) uint8_t edge[128]; /* left[0..31] at 0, above[-1] at 47, above[0..63] at 48 */
    __shared__ int dc_s;
    const int t = threadIdx.x;
    for (int b = blockIdx.x; b < n; b += gridDim.x) {
        const SvtB200IntraBlock B = blocks[b];
        const int bs = 4 << B.tx_size;
        if (t < bs) edge[t] = __ldg(B.left + t);
        if (t >= 32 && t - 32 < 2 * bs + 1) edge[47 + t - 32] = __ldg(B.above - 1 + (t - 32));
        __syncthreads();
        IntraEdge E; E.a = edge + 48; E.l = edge;
        if (t == 0) dc_s = B.mode == 0 ? intra_dc_value(bs, B.have_left, B.have_top, E) : 0;
        __syncthreads();
        const int dc = dc_s;
        const bool words = (((uintptr_t)B.dst | (uintptr_t)B.dst_stride) & 3) == 0;
        for (int i = t; i < (bs * bs) >> 2; i += INTRA_THREADS) {
            const int r = i >> B.tx_size, c4 = (i & ((1 << B.tx_size) - 1)) * 4; /* bs / 4 = 1 << tx_size groups per row */
            const uint32_t v0 = (uint32_t)intra_sample(B.mode, bs, dc, E, r, c4), v1 = (uint32_t)intra_sample(B.mode, bs, dc, E, r, c4 + 1);
            const uint32_t v2 = (uint32_t)intra_sample(B.mode, bs, dc, E, r, c4 + 2), v3 = (uint32_t)intra_sample(B.mode, bs, dc, E, r, c4 + 3);
            uint8_t *d = B.dst + (ptrdiff_t)r * B.dst_stride + c4;
            if (words) *(uint32_t *)d = v0 | v1 << 8 | v2 << 16 | v3 << 24;
            else { d[0] = (uint8_t)v0; d[1] = (uint8_t)v1; d[2] = (uint8_t)v2; d[3] = (uint8_t)v3; }
        }
        __syncthreads();
    }
}

extern "C" int svtb200_intra_predict_batch(SvtB200Ctx *c, const SvtB200IntraBlock *blocks, int n) {
    if (!c || n < 0 || (n && !blocks)) return SVTB200_EINVAL;
    if (!n) return SVTB200_OK;
    for (int i = 0; i < n; i++)
        if (!blocks[i].above || !blocks[i].left || !blocks[i].dst || blocks[i].mode > 9 || blocks[i].tx_size > 3) return SVTB200_EINVAL;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    SvtB200IntraBlock *d = (SvtB200IntraBlock *)pred_upload_desc(c, blocks, (size_t)n * sizeof(SvtB200IntraBlock));
    if (!d) return c->err.empty() ? SVTB200_ENOMEM : SVTB200_ECUDA;
    const int cap = (c->sm_count > 0 ? c->sm_count : 148) * 16;
    intra_kernel<<<n < cap ? n : cap, INTRA_THREADS, 0, c->stream>>>(d, n);
    c->launches++;
    SVTB200_CUDA_CHECK(c, cudaGetLastError());
    SVTB200_CUDA_CHECK(c, cudaEventRecord(c->pred_read, c->stream));
    return SVTB200_OK;
}

/* ---- full-loop distortion: one warp per transform block ---- */
__global__ void __launch_bounds__(256) full_dist_kernel(int nn, int n, const int16_t *coeff, const int16_t *rc, const uint16_t *eob, uint64_t *dist) {
    const int k = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (k >= n) return;
    const int16_t *a = coeff + (size_t)k * nn, *b = rc + (size_t)k * nn;
    uint32_t res = 0, pred = 0;
    for (int i = lane; i < nn; i += 32) {
        const int c = a[i], d = (int16_t)(c - b[i]); /* sq_r16to32 takes an int16_t (EbPictureOperators_C.c:264) */
        res += (uint32_t)(d * d);
        pred += (uint32_t)(c * c);
    }
    for (int o = 16; o > 0; o >>= 1) { res += __shfl_xor_sync(0xffffffffu, res, o); pred += __shfl_xor_sync(0xffffffffu, pred, o); }
    if (lane == 0) { dist[2 * k] = eob[k] ? res : pred; dist[2 * k + 1] = pred; }
}

extern "C" int svtb200_full_distortion_batch_dev(SvtB200Ctx *c, int tx_size, int n, const int16_t *coeff, const int16_t *recon_coeff,
                                                 const uint16_t *eob, uint64_t *dist) {
    if (!c || tx_size < 0 || tx_size > 3 || n < 0 || (n && (!coeff || !recon_coeff || !eob || !dist))) return SVTB200_EINVAL;
    if (!n) return SVTB200_OK;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    const int nn = (4 << tx_size) * (4 << tx_size);
    full_dist_kernel<<<(n + 7) / 8, 256, 0, c->stream>>>(nn, n, coeff, recon_coeff, eob, dist);
    c->launches++;
    SVTB200_CUDA_CHECK(c, cudaGetLastError());
    return SVTB200_OK;
}
