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

// ---- zz.cu / pa_chroma.cu kernels
#include "../../include/svtb200.h"
#include "../../svt-vp9_b200/csrc/zz_kernel.cuh"
#include "../../svt-vp9_b200/csrc/pa_chroma_kernel.cuh"

extern "C" int pa_emu_zz_sad(const SvtB200Plane *cur16, const SvtB200Plane *prev, int w, int h, int th_shift, uint8_t *scores, uint32_t *sads) {
    ZzArgs A;
    A.cur16 = cur16->buf; A.c_stride = cur16->stride; A.c_ox = cur16->origin_x; A.c_oy = cur16->origin_y;
    A.prev = prev->buf; A.p_stride = prev->stride; A.p_ox = prev->origin_x; A.p_oy = prev->origin_y;
    A.w = w; A.h = h; A.sbs_w = (w + 63) / 64; A.n_sb = A.sbs_w * ((h + 63) / 64); A.th_shift = th_shift; A.scores = scores; A.sads = sads;
    emu_launch(dim3((unsigned)((A.n_sb + 7) / 8)), dim3(256), 0, [&] { zz_sad_kernel(A); });
    return 0;
}
extern "C" int pa_emu_block_stats(const SvtB200Plane *full, int w, int h, uint8_t *y_mean, uint16_t *variance) {
    StatArgs A;
    A.full = full->buf; A.stride = full->stride; A.ox = full->origin_x; A.oy = full->origin_y;
    A.sbs_w = (w + 63) / 64; A.n_sb = A.sbs_w * ((h + 63) / 64); A.y_mean = y_mean; A.variance = variance;
    emu_launch(dim3((unsigned)((A.n_sb + 7) / 8)), dim3(256), 0, [&] { block_stats_kernel(A); });
    return 0;
}
extern "C" int pa_emu_luma_histogram(const SvtB200Plane *p16, int rw, int rh, uint32_t *hist, uint8_t *avg, uint64_t *sum16) {
    const int nr = rw * rh;
    std::vector<uint32_t> count((size_t)nr * 256, 0u);
    std::vector<unsigned long long> sum((size_t)nr, 0ull);
    HistArgs A;
    A.p16 = p16->buf; A.stride = p16->stride; A.ox = p16->origin_x; A.oy = p16->origin_y; A.w = p16->width; A.h = p16->height; A.rw = rw; A.rh = rh; A.slices = 8;
    A.count = count.data(); A.sum = sum.data();
    *sum16 = 0;
    emu_launch(dim3((unsigned)(nr * A.slices)), dim3(256), 0, [&] { hist_count_kernel(A); });
    emu_launch(dim3((unsigned)nr), dim3(256), 0, [&] { hist_finish_kernel(A, hist, avg, (unsigned long long *)sum16); });
    return 0;
}
extern "C" int pa_emu_chroma_mean(const uint8_t *cb, const uint8_t *cr, int stride, int w, int h, uint8_t *cb_mean, uint8_t *cr_mean) {
    ChromaMeanArgs A;
    A.pl[0] = cb; A.pl[1] = cr; A.stride = stride; A.w = w; A.h = h; A.sbs_w = (w + 63) / 64; A.n_sb = A.sbs_w * ((h + 63) / 64); A.mean[0] = cb_mean; A.mean[1] = cr_mean;
    emu_launch(dim3((unsigned)((A.n_sb + 7) / 8)), dim3(256), 0, [&] { chroma_mean_kernel(A); });
    return 0;
}
extern "C" int pa_emu_chroma_histogram(const uint8_t *cb, const uint8_t *cr, int stride, int w, int h, int rw, int rh, uint32_t *hist, uint8_t *avg, uint64_t *sum16) {
    const int nr = rw * rh * 2;
    std::vector<uint32_t> count((size_t)nr * 256, 0u);
    std::vector<unsigned long long> sum((size_t)nr, 0ull);
    ChromaHistArgs A;
    A.pl[0] = cb; A.pl[1] = cr; A.stride = stride; A.w = w; A.h = h; A.rw = rw; A.rh = rh; A.slices = 4; A.count = count.data(); A.sum = sum.data();
    sum16[0] = sum16[1] = 0;
    emu_launch(dim3((unsigned)(nr * A.slices)), dim3(256), 0, [&] { chroma_hist_count_kernel(A); });
    emu_launch(dim3((unsigned)nr), dim3(256), 0, [&] { chroma_hist_finish_kernel(A, hist, avg, (unsigned long long *)sum16); });
    return 0;
}

// ---- the front end of svtb200_picture_upload_luma (ctx.cu): device geometry of the full plane, decimated planes
#include "../../svt-vp9_b200/csrc/pa_front_kernel.cuh"
// full: the HOST padded plane (stride = w + 2 * host_pad); dev: a (h + 2 * host_pad) x dpitch buffer that receives the device layout (picture dev_ox columns in)
extern "C" int pa_emu_front(const uint8_t *full, int host_stride, int host_pad, int w, int h, uint8_t *dev, int dpitch, int dev_ox, uint8_t *quarter, int qstride, int qpad,
                            uint8_t *sixteenth, int sstride, int spad) {
    const int rows = h + 2 * host_pad;
    for (int y = 0; y < h; y++) memcpy(dev + (size_t)(host_pad + y) * dpitch + dev_ox, full + (size_t)(host_pad + y) * host_stride + host_pad, (size_t)w); // the 2-D copy
    emu_launch(dim3((unsigned)((dpitch / 4 + 127) / 128), (unsigned)rows), dim3(128), 0, [&] { pa_pad_full_kernel(dev, dpitch, dev_ox, host_pad, w, h, rows, dpitch, host_pad); });
    for (int lvl = 1; lvl <= 2; lvl++) {
        uint8_t *d = lvl == 1 ? quarter : sixteenth;
        if (!d) continue;
        const int step = lvl == 1 ? 2 : 4, pad = lvl == 1 ? qpad : spad, stride = lvl == 1 ? qstride : sstride;
        const int dw = (w + step - 1) / step, dh = (h + step - 1) / step, drows = dh + 2 * pad;
        emu_launch(dim3((unsigned)((stride / 4 + 127) / 128), (unsigned)drows), dim3(128), 0, [&] { pa_decimate_kernel(dev, dpitch, dev_ox, host_pad, d, stride, pad, pad, dw, dh, drows, step); });
    }
    return 0;
}
