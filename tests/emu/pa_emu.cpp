// CPU execution of the picture-analysis noise / edge kernels' device code (test infrastructure only).
#define SVTB200_CPU_EMU 1
#include "../../svt-vp9_b200/csrc/emu/cuda_emu.h"
#include "../../svt-vp9_b200/csrc/pa_noise_kernel.cuh"

extern "C" int pa_emu_noise_filter(int kind, const uint8_t *in, int is, int w, int h, uint8_t *den, int ds, uint8_t *noise) {
    const dim3 grid((unsigned)((w + 511) / 512), (unsigned)h), blk(128);
    if (kind == 0) emu_launch(grid, blk, 0, [&] { noise_filter_kernel<0>(in, is, w, h, den, ds, noise); });
    else if (kind == 3) emu_launch(grid, blk, 0, [&] { noise_filter_kernel<6>(in, is, w, h, den, ds, nullptr); });
    else emu_launch(grid, blk, 0, [&] { noise_filter_kernel<4>(in, is, w, h, den, ds, nullptr); });
    return 0;
}

extern "C" int pa_emu_noise_detect(const uint8_t *in, int is, int w, int h, int block, int pw, int n_sb, uint8_t *den, int ds, uint64_t noise_th, uint64_t den_th,
                                   uint64_t *noise_var, uint64_t *den_var, uint8_t *flat, uint64_t *totals) {
    NoiseDetArgs A;
    A.in = in; A.is = is; A.w = w; A.h = h; A.den = den; A.ds = ds; A.block = block; A.pw = pw; A.n_sb = n_sb;
    A.noise_th = noise_th; A.den_th = den_th;
    A.noise_var = (unsigned long long *)noise_var; A.den_var = (unsigned long long *)den_var; A.totals = (unsigned long long *)totals; A.flat = flat;
    totals[0] = totals[1] = 0;
    int areas;
    if (block == 64) { A.sbs_w = (w + 63) / 64; areas = A.sbs_w * ((h + 63) / 64); }
    else { A.sbs_w = w / 64; areas = A.sbs_w * (h / 64); for (int i = 0; i < n_sb; i++) { noise_var[i] = den_var[i] = 0; flat[i] = 0; } }
    if (areas > 0) emu_launch(dim3((unsigned)areas), dim3(256), 0, [&] { noise_detect_kernel(A); });
    return 0;
}

extern "C" int pa_emu_edge_detection(const uint8_t *y_mean, const uint16_t *variance, int w, int h, uint8_t *ebn, uint8_t *iso, uint8_t *sharp, uint16_t *avg) {
    const int sbs_w = (w + 63) / 64, sbs_h = (h + 63) / 64, n = sbs_w * sbs_h;
    emu_launch(dim3(1), dim3(1024), 0, [&] { pa_avg_variance_kernel(variance, n, avg); });
    emu_launch(dim3((unsigned)((n + 127) / 128)), dim3(128), 0, [&] { edge_detection_kernel(y_mean, variance, sbs_w, sbs_h, avg, ebn, iso, sharp); });
    return 0;
}

extern "C" int pa_emu_edge_cu(const uint8_t *y_mean, const uint8_t *cb, const uint8_t *cr, const uint8_t *logo, int w, int h, uint8_t *edge_cu, uint32_t *max_grad) {
    const int sbs_w = (w + 63) / 64, n = sbs_w * ((h + 63) / 64), blocks = (n * 16 + 255) / 256;
    *max_grad = 0;
    emu_launch(dim3((unsigned)blocks), dim3(256), 0, [&] { edge_cu_kernel<0>(y_mean, cb, cr, logo, w, h, sbs_w, n, max_grad, edge_cu); });
    emu_launch(dim3((unsigned)blocks), dim3(256), 0, [&] { edge_cu_kernel<1>(y_mean, cb, cr, logo, w, h, sbs_w, n, max_grad, edge_cu); });
    return 0;
}

extern "C" int pa_emu_homogeneous(const uint16_t *variance, int w, int h, uint32_t low_var_th, uint64_t detail_th, uint64_t *vv, uint8_t *hom, uint64_t *counts) {
    const int sbs_w = (w + 63) / 64, n = sbs_w * ((h + 63) / 64);
    counts[0] = counts[1] = counts[2] = 0;
    emu_launch(dim3((unsigned)((n + 127) / 128)), dim3(128), 0, [&] { homogeneous_regions_kernel(variance, w, h, sbs_w, n, low_var_th, detail_th, (unsigned long long *)vv, hom, (unsigned long long *)counts); });
    return 0;
}
