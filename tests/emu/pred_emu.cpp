// CPU execution of the prediction kernels' device code: the closed-form intra predictors and the inter-prediction kernel
// (test infrastructure only).
#define SVTB200_CPU_EMU 1
#include "../../svt-vp9_b200/csrc/emu/cuda_emu.h"
#include "../../svt-vp9_b200/csrc/intra_device.cuh"
#include "../../svt-vp9_b200/csrc/pred_device.cuh"
#include "../../include/svtb200.h"

extern "C" int intra_emu_run(const SvtB200IntraBlock *b, int n) {
    for (int i = 0; i < n; i++, b++) {
        const int bs = 4 << b->tx_size;
        IntraEdge E; E.a = b->above; E.l = b->left;
        const int dc = b->mode == 0 ? intra_dc_value(bs, b->have_left, b->have_top, E) : 0;
        for (int r = 0; r < bs; r++)
            for (int c = 0; c < bs; c++) b->dst[(ptrdiff_t)r * b->dst_stride + c] = (uint8_t)intra_sample(b->mode, bs, dc, E, r, c);
    }
    return 0;
}

extern "C" int pred_emu_run(const SvtB200PredBlock *blocks, int n, int grid) {
    emu_launch(dim3((unsigned)(n < grid ? n : grid)), dim3(PRED_THREADS), 0, [&] { pred_kernel(blocks, n); });
    return 0;
}

extern "C" int intra_edges_emu_run(const uint8_t *plane, int stride, int fw, int fh, const SvtB200IntraEdgeJob *jobs, int n, uint8_t *out) {
    emu_launch(dim3((unsigned)((n + 1) / 2)), dim3(INTRA_EDGE_THREADS), 0, [&] { intra_edges_kernel(plane, stride, fw, fh, jobs, n, out); });
    return 0;
}

extern "C" int full_dist_emu_run(int tx_size, int n, const int16_t *coeff, const int16_t *rc, const uint16_t *eob, uint64_t *dist) {
    const int nn = (4 << tx_size) * (4 << tx_size);
    emu_launch(dim3((unsigned)((n + 7) / 8)), dim3(256), 0, [&] { full_dist_kernel(nn, n, coeff, rc, eob, dist); });
    return 0;
}
