/*
 * pa_chroma.cu -- the chroma statistics of the picture-analysis front end on device-resident planes.
 *
 *   chroma_mean_kernel  compute_chroma_block_mean / zero_out_chroma_block_mean (Source/Lib/Codec/EbPictureAnalysisProcess.c:1828,
 *                       :1760; driver :4810-4822): one warp per super-block, lane = (plane, 8x8 chroma block); the 8x8 mean is the
 *                       sum of rows 0, 2, 4, 6 << 3 (eb_vp9_compute_sub_mean8x8_sse2_intrin), larger blocks (sum of four) >> 2 by
 *                       shuffles -- including the reference's 64x64 formula, which adds the last 32x32 quadrant twice (:2011-2016)
 *   chroma_hist_*       sub_sample_chroma_generate_pixel_intensity_histogram_bins (:4313) + calculate_histogram (:127): every
 *                       4th sample of every 4th chroma row of a region; bins (1 + count) << 4, region average, sum << 4
 */
#include "common.cuh"

#include "pa_chroma_kernel.cuh"

extern "C" {

int svtb200_chroma_block_mean_dev(SvtB200Ctx *c, const uint8_t *cb, const uint8_t *cr, int32_t stride, int32_t luma_width, int32_t luma_height,
                                  uint8_t *cb_mean, uint8_t *cr_mean) {
    if (!c || !cb || !cr || !cb_mean || !cr_mean || luma_width <= 0 || luma_height <= 0 || (stride & 7) || (((uintptr_t)cb | (uintptr_t)cr) & 7)) return SVTB200_EINVAL;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    cudaSetDevice(c->device);
    ChromaMeanArgs A;
    A.pl[0] = cb; A.pl[1] = cr; A.stride = stride; A.w = luma_width; A.h = luma_height;
    A.sbs_w = (luma_width + 63) / 64; A.n_sb = A.sbs_w * ((luma_height + 63) / 64);
    A.mean[0] = cb_mean; A.mean[1] = cr_mean;
    chroma_mean_kernel<<<(A.n_sb + 7) / 8, 256, 0, c->stream>>>(A);
    c->launches++;
    SVTB200_CUDA_CHECK(c, cudaGetLastError());
    return SVTB200_OK;
}

int svtb200_chroma_histogram_dev(SvtB200Ctx *c, const uint8_t *cb, const uint8_t *cr, int32_t stride, int32_t luma_width, int32_t luma_height,
                                 int32_t rw, int32_t rh, uint32_t *hist, uint8_t *avg, uint64_t *sum16) {
    if (!c || !cb || !cr || !hist || !avg || !sum16 || rw <= 0 || rh <= 0 || rw > 16 || rh > 16 || luma_width < 8 * rw || luma_height < 8 * rh) return SVTB200_EINVAL;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    cudaSetDevice(c->device);
    const int nr = rw * rh * 2;
    const size_t need = (size_t)nr * 256 * 4 + (size_t)nr * 8 + 1024;
    int rc = svtb200_grow(c, (void **)&c->scratch, &c->scratch_bytes, need);
    if (rc) return rc;
    ChromaHistArgs A;
    A.pl[0] = cb; A.pl[1] = cr; A.stride = stride; A.w = luma_width; A.h = luma_height; A.rw = rw; A.rh = rh; A.slices = 4;
    A.count = (uint32_t *)c->scratch; A.sum = (unsigned long long *)(c->scratch + (size_t)nr * 1024);
    SVTB200_CUDA_CHECK(c, cudaMemsetAsync(c->scratch, 0, (size_t)nr * 1024 + (size_t)nr * 8, c->stream));
    SVTB200_CUDA_CHECK(c, cudaMemsetAsync(sum16, 0, 16, c->stream));
    chroma_hist_count_kernel<<<nr * A.slices, 256, 0, c->stream>>>(A);
    chroma_hist_finish_kernel<<<nr, 256, 0, c->stream>>>(A, hist, avg, (unsigned long long *)sum16);
    c->launches += 2;
    SVTB200_CUDA_CHECK(c, cudaGetLastError());
    return SVTB200_OK;
}

} /* extern "C" */
