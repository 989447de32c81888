/* lfmask.cu -- host side of svtb200_build_masks[_dev] (see include/svtb200.h). */
#include "common.cuh"
#include "lfmask_kernel.cuh"

static int lfm_launch(SvtB200Ctx *c, int mi_rows, int mi_cols, const SvtB200LfModeInfo *mi, int mi_stride, SvtB200LfMask *lfm, int lfm_stride) {
    if (!mi || !lfm || mi_rows <= 0 || mi_cols <= 0 || mi_stride < mi_cols) return SVTB200_EINVAL;
    LfmArgs A;
    A.mi = mi; A.mi_rows = mi_rows; A.mi_cols = mi_cols; A.mi_stride = mi_stride; A.lfm = lfm; A.lfm_stride = lfm_stride;
    A.sbs_w = (mi_cols + 7) >> 3; A.n_sb = A.sbs_w * ((mi_rows + 7) >> 3);
    if (lfm_stride < A.sbs_w) return SVTB200_EINVAL;
    lf_build_masks_kernel<<<(A.n_sb + 63) / 64, 64, 0, c->stream>>>(A);
    c->launches++;
    SVTB200_CUDA_CHECK(c, cudaGetLastError());
    return SVTB200_OK;
}

extern "C" {

int svtb200_pick_filter_level(int32_t ac_q, int32_t is_key_frame) {
    int guess = (ac_q * 20723 + 1015158 + (1 << 17)) >> 18;
    if (is_key_frame) guess -= 4;
    return guess < 0 ? 0 : guess > 63 ? 63 : guess;
}

int svtb200_build_masks_dev(SvtB200Ctx *c, int32_t mi_rows, int32_t mi_cols, const SvtB200LfModeInfo *mi, int32_t mi_stride,
                            SvtB200LfMask *lfm, int32_t lfm_stride) {
    if (!c) return SVTB200_EINVAL;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    cudaSetDevice(c->device);
    return lfm_launch(c, mi_rows, mi_cols, mi, mi_stride, lfm, lfm_stride);
}

int svtb200_build_masks(SvtB200Ctx *c, int32_t mi_rows, int32_t mi_cols, const SvtB200LfModeInfo *mi, int32_t mi_stride,
                        SvtB200LfMask *lfm, int32_t lfm_stride) {
    if (!c || !mi || !lfm || mi_rows <= 0 || mi_cols <= 0) return SVTB200_EINVAL;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    cudaSetDevice(c->device);
    const int sbs_h = (mi_rows + 7) >> 3;
    const size_t mib = (size_t)mi_rows * mi_stride * sizeof(SvtB200LfModeInfo), lb = (size_t)sbs_h * lfm_stride * sizeof(SvtB200LfMask);
    int rc = svtb200_grow(c, (void **)&c->scratch, &c->scratch_bytes, mib + lb + 1024);
    if (rc) return rc;
    SvtB200LfModeInfo *dmi = (SvtB200LfModeInfo *)c->scratch;
    SvtB200LfMask *dl = (SvtB200LfMask *)(c->scratch + ((mib + 255) & ~(size_t)255));
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(dmi, mi, mib, cudaMemcpyHostToDevice, c->stream));
    SVTB200_CUDA_CHECK(c, cudaMemsetAsync(dl, 0, lb, c->stream));
    rc = lfm_launch(c, mi_rows, mi_cols, dmi, mi_stride, dl, lfm_stride);
    if (rc) return rc;
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(lfm, dl, lb, cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    return SVTB200_OK;
}

} /* extern "C" */
