/*
 * blocks.cu -- the block walk in front of the inter predictors of the encode pass, on the device.
 *
 * inter_expand_kernel: per (block, plane) what build_inter_predictors() (Source/Lib/VPX/vp9_reconinter.c:102-250) does before it
 * calls inter_predictor(): pick the reference plane of the vector's list (:146-160), scale the vector to the plane's 1/16-sample
 * units and clamp it against the picture edges (eb_vp9_clamp_mv_to_umv_border_sb, :68-87; edges as EbEncDecProcess.c:2052-2066
 * sets them), round it to whole samples when use_subpel_flag is off (:168-180), split it into the integer displacement and the
 * filter phase (:162-166, :184).  Output: SvtB200PredBlock descriptors, consumed by pred_kernel (pred.cu) without a host round trip.
 *
 * me_best_blocks_kernel: an open-loop partition for the bench / tests -- 32x32 blocks (16x16 in a last strip of 16), every block
 * with the best candidate of the matching motion-estimation partition, vectors converted like EbModeDecision.c:222-227.
 *
 * Integer arithmetic only; the source also compiles under csrc/emu/cuda_emu.h (SVTB200_CPU_EMU is not needed: plain functions).
 */
#include "common.cuh"
#include "blocks_device.cuh"

__global__ void __launch_bounds__(256) inter_expand_kernel(const SvtB200InterBlock *__restrict__ blk, int n, const __grid_constant__ SvtB200PredPicture pic,
                                                           SvtB200PredBlock *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * n) return;
    const int b = i / 3, plane = i - 3 * b;
    out[i] = blocks_expand_one(blk[b], plane, pic);
}

__global__ void __launch_bounds__(256) me_best_blocks_kernel(const SvtB200MeCuResult *__restrict__ me, int width, int height, int n,
                                                             SvtB200InterBlock *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int x, y, size;
    blocks_tile_geometry(width, height, i, &x, &y, &size);
    const int sbs_w = (width + 63) >> 6, sb = (y >> 6) * sbs_w + (x >> 6);
    const int pu = size == 32 ? 1 + ((y & 32) >> 5) * 2 + ((x & 32) >> 5) : 5 + ((y & 48) >> 4) * 4 + ((x & 48) >> 4);
    out[i] = blocks_from_me(me[(size_t)sb * 85 + pu], x, y, size);
}

extern "C" int svtb200_inter_blocks_expand_dev(SvtB200Ctx *c, const SvtB200InterBlock *blocks, int n, const SvtB200PredPicture *pic, SvtB200PredBlock *out) {
    if (!c || n < 0 || !pic || (n && (!blocks || !out))) return SVTB200_EINVAL;
    if (pic->width < 8 || pic->height < 8) return SVTB200_EINVAL;
    for (int p = 0; p < 3; p++) if (!pic->ref[0][p] || !pic->dst[p]) return SVTB200_EINVAL;
    if (!n) return SVTB200_OK;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    cudaSetDevice(c->device);
    inter_expand_kernel<<<(3 * n + 255) / 256, 256, 0, c->stream>>>(blocks, n, *pic, out);
    c->launches++;
    SVTB200_CUDA_CHECK(c, cudaGetLastError());
    return SVTB200_OK;
}

extern "C" int svtb200_me_best_blocks_dev(SvtB200Ctx *c, const SvtB200MeCuResult *me, int32_t width, int32_t height, SvtB200InterBlock *blocks) {
    if (!c || !me || !blocks || width < 16 || height < 16 || ((width | height) & 15) || width > 65535 || height > 65535) return SVTB200_EINVAL;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    cudaSetDevice(c->device);
    const int n = blocks_tile_count(width, height);
    me_best_blocks_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(me, width, height, n, blocks);
    c->launches++;
    SVTB200_CUDA_CHECK(c, cudaGetLastError());
    return n;
}

extern "C" int svtb200_inter_chain_dev(SvtB200Ctx *c, const SvtB200InterChain *ch) {
    if (!c || !ch || ch->n_blocks < 0 || !ch->tx || ch->n_tx < 1 || ch->n_reads < 0 || ch->n_reads > 3) return SVTB200_EINVAL;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    cudaSetDevice(c->device);
    int rc;
    struct Restore { SvtB200Ctx *c; cudaStream_t s; ~Restore() { c->stream = s; } } restore{c, c->stream}; /* (the lock is held throughout) */
    if (ch->on_queue >= 0) {
        if (ch->on_queue != SVTB200_STREAM_COMPUTE && ch->on_queue != SVTB200_STREAM_COMPUTE2 && ch->on_queue != SVTB200_STREAM_COMPUTE3) return SVTB200_EINVAL;
        c->stream = svtb200_pick_stream(c, ch->on_queue);
    }
    if (ch->after_queue >= 0) {
        if (ch->after_queue > SVTB200_STREAM_COMPUTE3) return SVTB200_EINVAL;
        cudaStream_t from = svtb200_pick_stream(c, ch->after_queue);
        if (from != c->stream) {
            if (!c->chain_ev) SVTB200_CUDA_CHECK(c, cudaEventCreateWithFlags(&c->chain_ev, cudaEventDisableTiming));
            SVTB200_CUDA_CHECK(c, cudaEventRecord(c->chain_ev, from));
            SVTB200_CUDA_CHECK(c, cudaStreamWaitEvent(c->stream, c->chain_ev, 0));
        }
    }
    if (ch->me) {
        rc = svtb200_me_best_blocks_dev(c, ch->me, ch->pic.width, ch->pic.height, ch->blocks);
        if (rc < 0) return rc;
        if (rc != ch->n_blocks) { c->err = "inter chain: n_blocks is not the tiling's block count"; return SVTB200_EINVAL; }
    }
    if (ch->n_blocks) {
        if ((rc = svtb200_inter_blocks_expand_dev(c, ch->blocks, ch->n_blocks, &ch->pic, ch->desc))) return rc;
        if ((rc = svtb200_inter_predict_batch_dev(c, ch->desc, 3 * ch->n_blocks))) return rc;
    }
    if ((rc = svtb200_txq_multi_dev(c, ch->tx, ch->n_tx, ch->qrows, ch->n_qrows))) return rc;
    if (ch->stream && (rc = svtb200_coeff_pack_dev(c, ch->tx, ch->n_tx, ch->stream, ch->capacity, ch->offsets))) return rc;
    if (ch->n_rate && (rc = svtb200_coeff_rate_multi_dev(c, ch->rate, ch->n_rate))) return rc;
    for (int i = 0; i < ch->n_reads; i++) svtb200_picture_touch(c, ch->reads[i]);
    return SVTB200_OK;
}
