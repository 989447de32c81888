/* zz.cu -- compute_zz_sad (EbMotionEstimationProcess.c:431) for a whole picture: one warp per super-block. */
#include "common.cuh"

struct ZzArgs {
    const uint8_t *cur16; int32_t c_stride, c_ox, c_oy;   /* current picture, 1/16 plane */
    const uint8_t *prev;  int32_t p_stride, p_ox, p_oy;   /* previous picture, full plane */
    int32_t w, h, sbs_w, n_sb, th_shift;
    uint8_t *scores; uint32_t *sads;
};

__global__ void __launch_bounds__(256) zz_sad_kernel(const ZzArgs A) {
    const int sb = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (sb >= A.n_sb) return;
    const int sx = (sb % A.sbs_w) * 64, sy = (sb / A.sbs_w) * 64;
    const bool complete = sx + 64 <= A.w && sy + 64 <= A.h;
    uint32_t sad = 0xffffffffu;
    if (complete) {
        /* lane: row y of the 16x16 block, 8 samples from x0 */
        const int y = lane >> 1, x0 = (lane & 1) * 8;
        const uint8_t *c = A.cur16 + (size_t)(A.c_oy + (sy >> 2) + y) * A.c_stride + A.c_ox + (sx >> 2) + x0;
        const uint8_t *p = A.prev + (size_t)(A.p_oy + sy + 4 * y) * A.p_stride + A.p_ox + sx + 4 * x0;
        uint32_t acc = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int a = __ldg(c + k), b = __ldg(p + 4 * k);
            acc += (uint32_t)(a > b ? a - b : b - a);
        }
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        sad = acc;
    }
    if (lane == 0) {
        const uint32_t t0 = 512u >> A.th_shift, t1 = 1024u >> A.th_shift, t2 = 2048u >> A.th_shift;
        A.scores[sb] = sad < t0 ? 0 : sad < t1 ? 10 : sad < t2 ? 20 : 30;
        if (A.sads) A.sads[sb] = sad;
    }
}

static int zz_launch(SvtB200Ctx *c, uint64_t cur_id, uint64_t prev_id, int w, int h, int th_shift, uint8_t *scores, uint32_t *sads) {
    if (!scores || w <= 0 || h <= 0 || th_shift < 0 || th_shift > 11) return SVTB200_EINVAL;
    auto ic = c->pics.find(cur_id), ip = c->pics.find(prev_id);
    if (ic == c->pics.end() || ip == c->pics.end() || !ic->second.pl[2].buf || !ip->second.pl[0].buf) { c->err = "zz_sad: picture not resident"; return SVTB200_ENOENT; }
    ZzArgs A;
    const DevPlane &c16 = ic->second.pl[2], &pf = ip->second.pl[0];
    A.cur16 = c16.buf; A.c_stride = c16.stride; A.c_ox = c16.ox; A.c_oy = c16.oy;
    A.prev = pf.buf; A.p_stride = pf.stride; A.p_ox = pf.ox; A.p_oy = pf.oy;
    A.w = w; A.h = h; A.sbs_w = (w + 63) / 64; A.n_sb = A.sbs_w * ((h + 63) / 64); A.th_shift = th_shift;
    A.scores = scores; A.sads = sads;
    svtb200_pic_acquire(c, ic->second); svtb200_pic_acquire(c, ip->second);
    zz_sad_kernel<<<(A.n_sb + 7) / 8, 256, 0, c->stream>>>(A);
    svtb200_pic_release(c, ic->second); svtb200_pic_release(c, ip->second);
    c->launches++;
    SVTB200_CUDA_CHECK(c, cudaGetLastError());
    return SVTB200_OK;
}

extern "C" {

int svtb200_zz_sad_picture_dev(SvtB200Ctx *c, uint64_t cur_id, uint64_t prev_id, int32_t w, int32_t h, int32_t th_shift, uint8_t *scores, uint32_t *sads) {
    if (!c) return SVTB200_EINVAL;
    std::lock_guard<std::mutex> g(c->mu);
    cudaSetDevice(c->device);
    return zz_launch(c, cur_id, prev_id, w, h, th_shift, scores, sads);
}

int svtb200_zz_sad_picture(SvtB200Ctx *c, uint64_t cur_id, uint64_t prev_id, int32_t w, int32_t h, int32_t th_shift, uint8_t *scores, uint32_t *sads) {
    if (!c || !scores || w <= 0 || h <= 0) return SVTB200_EINVAL;
    std::lock_guard<std::mutex> g(c->mu);
    cudaSetDevice(c->device);
    const size_t n = (size_t)((w + 63) / 64) * ((h + 63) / 64);
    int rc = svtb200_grow(c, (void **)&c->scratch, &c->scratch_bytes, n * 5 + 512);
    if (rc) return rc;
    uint32_t *d_sad = (uint32_t *)c->scratch;
    uint8_t *d_sc = c->scratch + ((n * 4 + 255) & ~(size_t)255);
    rc = zz_launch(c, cur_id, prev_id, w, h, th_shift, d_sc, d_sad);
    if (rc) return rc;
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(scores, d_sc, n, cudaMemcpyDeviceToHost, c->stream));
    if (sads) SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(sads, d_sad, n * 4, cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    return SVTB200_OK;
}

} /* extern "C" */
