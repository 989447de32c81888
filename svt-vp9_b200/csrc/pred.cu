/* pred.cu -- prediction of the encode pass and full-loop distortion (the calls either side of every transform call):
 *   svtb200_inter_predict_batch        inter_predictor() as build_inter_predictors() calls it (Source/Lib/VPX/vp9_reconinter.c:102-250,
 *                                      vp9_reconinter.h:23): the regular 8-tap kernel, unscaled, dispatched like scale_factors::predict
 *                                      (vp9_scale.c:78-128) on eb_vp9_convolve_copy / convolve8_horiz / _vert / convolve8 and their
 *                                      _avg twins (Source/Lib/VPX/vpx_convolve.c:101-215)
 *   svtb200_full_distortion_batch_dev  full_distortion_kernel32bit / _eob_zero32bit (Source/Lib/C_DEFAULT/EbPictureOperators_C.c:289,314)
 *                                      behind the table lookup of Source/Lib/Codec/EbEncDecProcess.c:723
 * One CTA per block: the (w + 7) x (h + 7) reference window is staged once in shared memory (only the rows / columns the
 * reference functions read are touched), the horizontal pass writes the 8-bit intermediate the reference keeps in 'temp', the
 * vertical pass reads it; a compound block runs both references through the same buffers and averages in registers. */
#include "common.cuh"

#define PRED_THREADS 128
#define PRED_WP 72 /* window pitch: 64 + 7 columns, word aligned */

/* the normative VP9 regular 8-tap filter (sub_pel_filters_8, Source/Lib/VPX/vp9_filter.c:32); phase 0 (a copy) never filters */
__constant__ int8_t c_pred_taps[16][8] = {
    {0, 0, 0, 127, 0, 0, 0, 0},
    {0, 1, -5, 126, 8, -3, 1, 0},      {-1, 3, -10, 122, 18, -6, 2, 0},   {-1, 4, -13, 118, 27, -9, 3, -1},
    {-1, 4, -16, 112, 37, -11, 4, -1}, {-1, 5, -18, 105, 48, -14, 4, -1}, {-1, 5, -19, 97, 58, -16, 5, -1},
    {-1, 6, -19, 88, 68, -18, 5, -1},  {-1, 6, -19, 78, 78, -19, 6, -1},  {-1, 5, -18, 68, 88, -19, 6, -1},
    {-1, 5, -16, 58, 97, -19, 5, -1},  {-1, 4, -14, 48, 105, -18, 5, -1}, {-1, 4, -11, 37, 112, -16, 4, -1},
    {-1, 3, -9, 27, 118, -13, 4, -1},  {0, 2, -6, 18, 122, -10, 3, -1},   {0, 1, -3, 8, 126, -5, 1, 0}};

__device__ __forceinline__ int pred_clip8(int v) { return v < 0 ? 0 : v > 255 ? 255 : v; }

/* one reference of the block -> res (pitch 64, shared) */
__device__ static void pred_one(const uint8_t *src /* at the block's integer position */, int stride, int sx, int sy, int w, int h,
                                uint8_t *win, uint8_t *tmp, uint8_t *res) {
    const int t = threadIdx.x;
    const int x_lo = sx ? -3 : 0, nx = sx ? w + 7 : w;
    const int y_lo = sy ? -3 : 0, ny = sy ? h + 7 : h;
    for (int i = t; i < nx * ny; i += PRED_THREADS) {
        const int r = i / nx, c = i - r * nx;
        win[r * PRED_WP + c] = __ldg(src + (ptrdiff_t)(r + y_lo) * stride + c + x_lo);
    }
    __syncthreads();
    /* horizontal pass over every staged row (h, or h + 7 when a vertical pass follows: vpx_convolve.c:153-162) */
    const uint8_t *v_in = win;
    int v_pitch = PRED_WP;
    if (sx) {
        int taps[8];
#pragma unroll
        for (int k = 0; k < 8; k++) taps[k] = c_pred_taps[sx][k];
        for (int i = t; i < w * ny; i += PRED_THREADS) {
            const int r = i / w, c = i - r * w;
            const uint8_t *p = win + r * PRED_WP + c;
            int sum = 64;
#pragma unroll
            for (int k = 0; k < 8; k++) sum += p[k] * taps[k];
            tmp[r * 64 + c] = (uint8_t)pred_clip8(sum >> 7);
        }
        __syncthreads();
        v_in = tmp; v_pitch = 64;
    }
    if (sy) {
        int taps[8];
#pragma unroll
        for (int k = 0; k < 8; k++) taps[k] = c_pred_taps[sy][k];
        for (int i = t; i < w * h; i += PRED_THREADS) {
            const int r = i / w, c = i - r * w;
            const uint8_t *p = v_in + r * v_pitch + c;
            int sum = 64;
#pragma unroll
            for (int k = 0; k < 8; k++) sum += p[k * v_pitch] * taps[k];
            res[r * 64 + c] = (uint8_t)pred_clip8(sum >> 7);
        }
    } else {
        for (int i = t; i < w * h; i += PRED_THREADS) {
            const int r = i / w, c = i - r * w;
            res[r * 64 + c] = v_in[r * v_pitch + c];
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(PRED_THREADS) pred_kernel(const SvtB200PredBlock *blocks, int n) {
    __shared__ __align__(16) uint8_t win[71 * PRED_WP];
    __shared__ __align__(16) uint8_t tmp[71 * 64];
    __shared__ __align__(16) uint8_t res[2][64 * 64];
    for (int b = blockIdx.x; b < n; b += gridDim.x) {
        const SvtB200PredBlock B = blocks[b];
        pred_one(B.ref[0] + (ptrdiff_t)B.y[0] * B.ref_stride[0] + B.x[0], B.ref_stride[0], B.subpel_x[0], B.subpel_y[0], B.w, B.h, win, tmp, res[0]);
        if (B.compound)
            pred_one(B.ref[1] + (ptrdiff_t)B.y[1] * B.ref_stride[1] + B.x[1], B.ref_stride[1], B.subpel_x[1], B.subpel_y[1], B.w, B.h, win, tmp, res[1]);
        for (int i = threadIdx.x; i < B.w * B.h; i += PRED_THREADS) {
            const int r = i / B.w, c = i - r * B.w;
            int v = res[0][r * 64 + c];
            if (B.compound) v = (v + res[1][r * 64 + c] + 1) >> 1; /* eb_vp9_convolve_avg_c, vpx_convolve.c:211 */
            B.dst[(ptrdiff_t)r * B.dst_stride + c] = (uint8_t)v;
        }
        __syncthreads();
    }
}

extern "C" int svtb200_inter_predict_batch(SvtB200Ctx *c, const SvtB200PredBlock *blocks, int n) {
    if (!c || n < 0 || (n && !blocks)) return SVTB200_EINVAL;
    if (!n) return SVTB200_OK;
    for (int i = 0; i < n; i++) {
        const SvtB200PredBlock &b = blocks[i];
        if (!b.ref[0] || !b.dst || b.w < 2 || b.w > 64 || b.h < 2 || b.h > 64 || b.subpel_x[0] > 15 || b.subpel_y[0] > 15 || b.compound > 1) return SVTB200_EINVAL;
        if (b.compound && (!b.ref[1] || b.subpel_x[1] > 15 || b.subpel_y[1] > 15)) return SVTB200_EINVAL;
    }
    std::lock_guard<std::recursive_mutex> g(c->mu);
    const size_t bytes = (size_t)n * sizeof(SvtB200PredBlock);
    /* the descriptors of the previous batch may still be read by its kernel, possibly on the other compute queue */
    if (c->pred_read) SVTB200_CUDA_CHECK(c, cudaStreamWaitEvent(c->stream, c->pred_read, 0));
    else SVTB200_CUDA_CHECK(c, cudaEventCreateWithFlags(&c->pred_read, cudaEventDisableTiming));
    if (c->pred_desc_bytes < bytes) SVTB200_CUDA_CHECK(c, cudaEventSynchronize(c->pred_read)); /* about to free the buffer */
    if (svtb200_grow(c, &c->pred_desc, &c->pred_desc_bytes, bytes)) return SVTB200_ENOMEM;
    SvtB200PredBlock *d = (SvtB200PredBlock *)c->pred_desc;
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(d, blocks, bytes, cudaMemcpyHostToDevice, c->stream)); /* pageable source: staged before it returns */
    const int grid = n < 148 * 8 ? n : 148 * 8;
    pred_kernel<<<grid, PRED_THREADS, 0, c->stream>>>(d, n);
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
