// CPU execution of the deblocking kernel source under csrc/emu/cuda_emu.h (test infrastructure only).
#define SVTB200_CPU_EMU 1
#include "../../svt-vp9_b200/csrc/emu/cuda_emu.h"
#include "../../svt-vp9_b200/csrc/deblock_kernel.cuh"
#include "../../svt-vp9_b200/csrc/lfmask_kernel.cuh"

extern "C" int lf_emu_run(const SvtB200LfParams *p, uint8_t *y, int ys, uint8_t *u, uint8_t *v, int cs, const SvtB200LfMask *lfm,
                          int lfm_stride, const SvtB200LfThresh *thr) {
    LfKernelArgs A;
    memset(&A, 0, sizeof(A));
    A.P = *p; A.y = y; A.u = u; A.v = v; A.y_stride = ys; A.uv_stride = cs; A.lfm = lfm; A.lfm_stride = lfm_stride; A.thr = thr;
    A.sbs_w = (p->mi_cols + 7) >> 3; A.sbs_h = (p->mi_rows + 7) >> 3;
    std::vector<int32_t> done((size_t)A.sbs_h * LF_FLAG_PAD, 0);
    int32_t ticket = 0;
    A.ticket = &ticket; A.done = done.data(); A.gen = 1;
    emu_launch(dim3(A.sbs_h), dim3(LF_THREADS), (size_t)A.sbs_w * (sizeof(SvtB200LfMask) + 16), [&] { lf_row_kernel(A); });
    return 0;
}

extern "C" int lfmask_emu_run(int mi_rows, int mi_cols, const SvtB200LfModeInfo *mi, int stride, SvtB200LfMask *lfm, int lfm_stride) {
    LfmArgs A;
    A.mi = mi; A.mi_rows = mi_rows; A.mi_cols = mi_cols; A.mi_stride = stride; A.lfm = lfm; A.lfm_stride = lfm_stride;
    A.sbs_w = (mi_cols + 7) >> 3; A.n_sb = A.sbs_w * ((mi_rows + 7) >> 3);
    emu_launch(dim3((A.n_sb + 63) / 64), dim3(64), 0, [&] { lf_build_masks_kernel(A); });
    return 0;
}
