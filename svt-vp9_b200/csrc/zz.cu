/* zz.cu -- picture-analysis side kernels, one warp per super-block: compute_zz_sad (EbMotionEstimationProcess.c:431) and
 * compute_block_mean_compute_variance (EbPictureAnalysisProcess.c:2115). */
#include "common.cuh"

#include "zz_kernel.cuh"

static int stats_launch(SvtB200Ctx *c, uint64_t pic_id, int w, int h, uint8_t *y_mean, uint16_t *variance) {
    if (!y_mean || !variance || w <= 0 || h <= 0) return SVTB200_EINVAL;
    auto it = c->pics.find(pic_id);
    if (it == c->pics.end() || !it->second.pl[0].buf) { c->err = "block_stats: picture not resident"; return SVTB200_ENOENT; }
    const DevPlane &f = it->second.pl[0];
    StatArgs A;
    A.full = f.buf; A.stride = f.stride; A.ox = f.ox; A.oy = f.oy;
    A.sbs_w = (w + 63) / 64; A.n_sb = A.sbs_w * ((h + 63) / 64); A.y_mean = y_mean; A.variance = variance;
    svtb200_pic_acquire(c, it->second);
    block_stats_kernel<<<(A.n_sb + 7) / 8, 256, 0, c->stream>>>(A);
    svtb200_pic_release(c, it->second);
    c->launches++;
    SVTB200_CUDA_CHECK(c, cudaGetLastError());
    return SVTB200_OK;
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

int svtb200_luma_histogram_picture(SvtB200Ctx *c, uint64_t pic_id, int32_t rw, int32_t rh, uint32_t *hist, uint8_t *avg, uint64_t *sum16) {
    if (!c || !hist || !avg || !sum16 || rw <= 0 || rh <= 0 || rw > 16 || rh > 16) return SVTB200_EINVAL;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    cudaSetDevice(c->device);
    auto it = c->pics.find(pic_id);
    if (it == c->pics.end() || !it->second.pl[2].buf) { c->err = "luma_histogram: picture or its 1/16 plane not resident"; return SVTB200_ENOENT; }
    const DevPlane &p = it->second.pl[2];
    if (p.w < rw || p.h < rh) return SVTB200_EINVAL;
    const int nr = rw * rh;
    const size_t need = (size_t)nr * 256 * 4 * 2 + (size_t)nr * 8 + nr + 8 + 1024;
    int rc = svtb200_grow(c, (void **)&c->scratch, &c->scratch_bytes, need);
    if (rc) return rc;
    uint8_t *q = c->scratch;
    auto carve = [&](size_t b) { uint8_t *r = q; q += (b + 255) & ~(size_t)255; return r; };
    HistArgs A;
    A.p16 = p.buf; A.stride = p.stride; A.ox = p.ox; A.oy = p.oy; A.w = p.w; A.h = p.h; A.rw = rw; A.rh = rh; A.slices = 8;
    A.count = (uint32_t *)carve((size_t)nr * 1024); A.sum = (unsigned long long *)carve((size_t)nr * 8);
    uint32_t *d_hist = (uint32_t *)carve((size_t)nr * 1024);
    unsigned long long *d_s16 = (unsigned long long *)carve(8);
    uint8_t *d_avg = carve(nr);
    SVTB200_CUDA_CHECK(c, cudaMemsetAsync(c->scratch, 0, (size_t)(q - c->scratch), c->stream));
    svtb200_pic_acquire(c, it->second);
    hist_count_kernel<<<nr * A.slices, 256, 0, c->stream>>>(A);
    hist_finish_kernel<<<nr, 256, 0, c->stream>>>(A, d_hist, d_avg, d_s16);
    svtb200_pic_release(c, it->second);
    c->launches += 2;
    SVTB200_CUDA_CHECK(c, cudaGetLastError());
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(hist, d_hist, (size_t)nr * 1024, cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(avg, d_avg, nr, cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(sum16, d_s16, 8, cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    return SVTB200_OK;
}

int svtb200_block_stats_picture_dev(SvtB200Ctx *c, uint64_t pic_id, int32_t w, int32_t h, uint8_t *y_mean, uint16_t *variance) {
    if (!c) return SVTB200_EINVAL;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    cudaSetDevice(c->device);
    return stats_launch(c, pic_id, w, h, y_mean, variance);
}

int svtb200_block_stats_picture(SvtB200Ctx *c, uint64_t pic_id, int32_t w, int32_t h, uint8_t *y_mean, uint16_t *variance) {
    if (!c || !y_mean || !variance || w <= 0 || h <= 0) return SVTB200_EINVAL;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    cudaSetDevice(c->device);
    const size_t n = (size_t)((w + 63) / 64) * ((h + 63) / 64) * 85;
    int rc = svtb200_grow(c, (void **)&c->scratch, &c->scratch_bytes, n * 3 + 512);
    if (rc) return rc;
    uint16_t *d_v = (uint16_t *)c->scratch;
    uint8_t *d_m = c->scratch + ((n * 2 + 255) & ~(size_t)255);
    rc = stats_launch(c, pic_id, w, h, d_m, d_v);
    if (rc) return rc;
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(y_mean, d_m, n, cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(variance, d_v, n * 2, cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    return SVTB200_OK;
}

int svtb200_zz_sad_picture_dev(SvtB200Ctx *c, uint64_t cur_id, uint64_t prev_id, int32_t w, int32_t h, int32_t th_shift, uint8_t *scores, uint32_t *sads) {
    if (!c) return SVTB200_EINVAL;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    cudaSetDevice(c->device);
    return zz_launch(c, cur_id, prev_id, w, h, th_shift, scores, sads);
}

int svtb200_zz_sad_picture(SvtB200Ctx *c, uint64_t cur_id, uint64_t prev_id, int32_t w, int32_t h, int32_t th_shift, uint8_t *scores, uint32_t *sads) {
    if (!c || !scores || w <= 0 || h <= 0) return SVTB200_EINVAL;
    std::lock_guard<std::recursive_mutex> g(c->mu);
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
