/* pa_noise.cu -- rest of the picture-analysis front end: the noise filters and noise-detection statistics of
 * picture_pre_processing_operations (Source/Lib/Codec/EbPictureAnalysisProcess.c:1403-1760, :3512-3560) and the two edge maps of
 * compute_picture_spatial_statistics (edge_detection :4545, edge_detection_mean_luma_chroma16x16 :4434).  See include/svtb200.h. */
#include "common.cuh"

#include "pa_noise_kernel.cuh"

extern "C" {

int svtb200_noise_filter_dev(SvtB200Ctx *c, int kind, const uint8_t *in, int32_t in_stride, int32_t w, int32_t h, uint8_t *den, int32_t den_stride, uint8_t *noise) {
    if (!c || !in || !den || kind < 0 || kind > 3 || w < 3 || h < 3 || in_stride < w || den_stride < w || (noise && kind != SVTB200_NOISE_LUMA_WEAK)) return SVTB200_EINVAL;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    cudaSetDevice(c->device);
    const dim3 grid((w + 511) / 512, h), blk(128);
    if (kind == SVTB200_NOISE_LUMA_WEAK) noise_filter_kernel<0><<<grid, blk, 0, c->stream>>>(in, in_stride, w, h, den, den_stride, noise);
    else if (kind == SVTB200_NOISE_CHROMA_STRONG) noise_filter_kernel<6><<<grid, blk, 0, c->stream>>>(in, in_stride, w, h, den, den_stride, nullptr);
    else noise_filter_kernel<4><<<grid, blk, 0, c->stream>>>(in, in_stride, w, h, den, den_stride, nullptr); /* luma strong, chroma weak */
    c->launches++;
    SVTB200_CUDA_CHECK(c, cudaGetLastError());
    return SVTB200_OK;
}

static int noise_detect_launch(SvtB200Ctx *c, const uint8_t *in, int32_t in_stride, int32_t w, int32_t h, int32_t block, int32_t pw, int32_t n_sb, uint8_t *den,
                               int32_t den_stride, uint64_t noise_th, uint64_t den_th, uint64_t *noise_var, uint64_t *den_var, uint8_t *flat, uint64_t *totals) {
    if (!c || !in || !den || !noise_var || !den_var || !flat || !totals || w < 3 || h < 3 || in_stride < w || den_stride < w) return SVTB200_EINVAL;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    cudaSetDevice(c->device);
    NoiseDetArgs A;
    A.in = in; A.is = in_stride; A.w = w; A.h = h; A.den = den; A.ds = den_stride;
    A.block = block; A.pw = pw; A.n_sb = n_sb;
    A.noise_th = noise_th; A.den_th = den_th;
    A.noise_var = (unsigned long long *)noise_var; A.den_var = (unsigned long long *)den_var; A.totals = (unsigned long long *)totals; A.flat = flat;
    SVTB200_CUDA_CHECK(c, cudaMemsetAsync(totals, 0, 16, c->stream));
    int areas;
    if (block == 64) { A.sbs_w = (w + 63) / 64; areas = A.sbs_w * ((h + 63) / 64); }
    else { /* the reference walks whole 64x64 areas of the decimated plane only; super-blocks it never visits keep zeros */
        A.sbs_w = w / 64; areas = A.sbs_w * (h / 64);
        SVTB200_CUDA_CHECK(c, cudaMemsetAsync(noise_var, 0, (size_t)n_sb * 8, c->stream));
        SVTB200_CUDA_CHECK(c, cudaMemsetAsync(den_var, 0, (size_t)n_sb * 8, c->stream));
        SVTB200_CUDA_CHECK(c, cudaMemsetAsync(flat, 0, (size_t)n_sb, c->stream));
    }
    if (areas > 0) {
        noise_detect_kernel<<<areas, 256, 0, c->stream>>>(A);
        c->launches++;
    }
    SVTB200_CUDA_CHECK(c, cudaGetLastError());
    return SVTB200_OK;
}

int svtb200_noise_detect_dev(SvtB200Ctx *c, const uint8_t *in, int32_t in_stride, int32_t w, int32_t h, uint8_t *den, int32_t den_stride, uint64_t noise_th,
                             uint64_t den_th, uint64_t *noise_var, uint64_t *den_var, uint8_t *flat, uint64_t *totals) {
    return noise_detect_launch(c, in, in_stride, w, h, 64, (w + 63) / 64, ((w + 63) / 64) * ((h + 63) / 64), den, den_stride, noise_th, den_th, noise_var, den_var, flat, totals);
}

int svtb200_noise_detect_decimated_dev(SvtB200Ctx *c, const uint8_t *in, int32_t in_stride, int32_t w, int32_t h, int32_t block, int32_t picture_width_in_sb, int32_t n_sb,
                                       uint8_t *den, int32_t den_stride, uint64_t noise_th, uint64_t den_th, uint64_t *noise_var, uint64_t *den_var, uint8_t *flat,
                                       uint64_t *totals) {
    if ((block != 32 && block != 16) || picture_width_in_sb <= 0 || n_sb <= 0) return SVTB200_EINVAL;
    return noise_detect_launch(c, in, in_stride, w, h, block, picture_width_in_sb, n_sb, den, den_stride, noise_th, den_th, noise_var, den_var, flat, totals);
}

int svtb200_edge_detection_dev(SvtB200Ctx *c, const uint8_t *y_mean, const uint16_t *variance, int32_t w, int32_t h, uint8_t *edge_block_num, uint8_t *isolated,
                               uint8_t *sharp_edge, uint16_t *pic_avg_variance) {
    if (!c || !y_mean || !variance || !edge_block_num || !isolated || !sharp_edge || !pic_avg_variance || w <= 0 || h <= 0) return SVTB200_EINVAL;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    cudaSetDevice(c->device);
    const int sbs_w = (w + 63) / 64, sbs_h = (h + 63) / 64, n = sbs_w * sbs_h;
    pa_avg_variance_kernel<<<1, 1024, 0, c->stream>>>(variance, n, pic_avg_variance);
    edge_detection_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(y_mean, variance, sbs_w, sbs_h, pic_avg_variance, edge_block_num, isolated, sharp_edge);
    c->launches += 2;
    SVTB200_CUDA_CHECK(c, cudaGetLastError());
    return SVTB200_OK;
}

int svtb200_homogeneous_regions_dev(SvtB200Ctx *c, const uint16_t *variance, int32_t w, int32_t h, uint32_t low_var_th, uint64_t detail_th, uint64_t *var_of_var32,
                                    uint8_t *homogeneous, uint64_t *counts) {
    if (!c || !variance || !var_of_var32 || !homogeneous || !counts || w <= 0 || h <= 0) return SVTB200_EINVAL;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    cudaSetDevice(c->device);
    const int sbs_w = (w + 63) / 64, n = sbs_w * ((h + 63) / 64);
    SVTB200_CUDA_CHECK(c, cudaMemsetAsync(counts, 0, 24, c->stream));
    homogeneous_regions_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(variance, w, h, sbs_w, n, low_var_th, detail_th, (unsigned long long *)var_of_var32, homogeneous,
                                                                    (unsigned long long *)counts);
    c->launches++;
    SVTB200_CUDA_CHECK(c, cudaGetLastError());
    return SVTB200_OK;
}

int svtb200_edge_cu_dev(SvtB200Ctx *c, const uint8_t *y_mean, const uint8_t *cb_mean, const uint8_t *cr_mean, const uint8_t *logo_flags, int32_t w, int32_t h,
                        uint64_t picture_number, uint8_t *edge_cu, uint32_t *max_grad) {
    if (!c || !y_mean || !cb_mean || !cr_mean || !logo_flags || !edge_cu || !max_grad || w <= 0 || h <= 0) return SVTB200_EINVAL;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    cudaSetDevice(c->device);
    const int sbs_w = (w + 63) / 64, n = sbs_w * ((h + 63) / 64);
    if (picture_number & 3) { /* "the values are calculated for every 4th frame" (:4441); the others get zeros */
        SVTB200_CUDA_CHECK(c, cudaMemsetAsync(edge_cu, 0, (size_t)n * 16, c->stream));
        return SVTB200_OK;
    }
    SVTB200_CUDA_CHECK(c, cudaMemsetAsync(max_grad, 0, 4, c->stream));
    const int blocks = (n * 16 + 255) / 256;
    edge_cu_kernel<0><<<blocks, 256, 0, c->stream>>>(y_mean, cb_mean, cr_mean, logo_flags, w, h, sbs_w, n, max_grad, edge_cu);
    edge_cu_kernel<1><<<blocks, 256, 0, c->stream>>>(y_mean, cb_mean, cr_mean, logo_flags, w, h, sbs_w, n, max_grad, edge_cu);
    c->launches += 2;
    SVTB200_CUDA_CHECK(c, cudaGetLastError());
    return SVTB200_OK;
}

} /* extern "C" */
