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

#include "pred_device.cuh"

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
    cudaSetDevice(c->device);
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
    cudaSetDevice(c->device);
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
    cudaSetDevice(c->device);
    SvtB200IntraBlock *d = (SvtB200IntraBlock *)pred_upload_desc(c, blocks, (size_t)n * sizeof(SvtB200IntraBlock));
    if (!d) return c->err.empty() ? SVTB200_ENOMEM : SVTB200_ECUDA;
    const int cap = (c->sm_count > 0 ? c->sm_count : 148) * 16;
    intra_kernel<<<n < cap ? n : cap, INTRA_THREADS, 0, c->stream>>>(d, n);
    c->launches++;
    SVTB200_CUDA_CHECK(c, cudaGetLastError());
    SVTB200_CUDA_CHECK(c, cudaEventRecord(c->pred_read, c->stream));
    return SVTB200_OK;
}

extern "C" int svtb200_intra_edges_dev(SvtB200Ctx *c, const uint8_t *plane, int32_t stride, int32_t fw, int32_t fh, const SvtB200IntraEdgeJob *jobs, int n,
                                       uint8_t *edges) {
    if (!c || n < 0 || (n && (!plane || !jobs || !edges)) || fw <= 0 || fh <= 0 || stride < fw) return SVTB200_EINVAL;
    if (!n) return SVTB200_OK;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    cudaSetDevice(c->device);
    intra_edges_kernel<<<(n + 1) / 2, INTRA_EDGE_THREADS, 0, c->stream>>>(plane, stride, fw, fh, jobs, n, edges);
    c->launches++;
    SVTB200_CUDA_CHECK(c, cudaGetLastError());
    return SVTB200_OK;
}

extern "C" int svtb200_full_distortion_batch_dev(SvtB200Ctx *c, int tx_size, int n, const int16_t *coeff, const int16_t *recon_coeff,
                                                 const uint16_t *eob, uint64_t *dist) {
    if (!c || tx_size < 0 || tx_size > 3 || n < 0 || (n && (!coeff || !recon_coeff || !eob || !dist))) return SVTB200_EINVAL;
    if (!n) return SVTB200_OK;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    cudaSetDevice(c->device);
    const int nn = (4 << tx_size) * (4 << tx_size);
    full_dist_kernel<<<(n + 7) / 8, 256, 0, c->stream>>>(nn, n, coeff, recon_coeff, eob, dist);
    c->launches++;
    SVTB200_CUDA_CHECK(c, cudaGetLastError());
    return SVTB200_OK;
}
