// CPU execution of the ME kernel source under csrc/emu/cuda_emu.h -- used by the "not gpu" tests to
// check the kernel's control logic against the oracle without a GPU.  Test infrastructure only.
#define SVTB200_CPU_EMU 1
#include "../../svt-vp9_b200/csrc/emu/cuda_emu.h"
SVTB200_EMU_DEFINE_STATE
#include "../../svt-vp9_b200/csrc/me_kernel.cuh"

static void to_dev(const SvtB200Plane &d, MePlaneDev *o) { o->buf = d.buf; o->stride = d.stride; o->ox = d.origin_x; o->oy = d.origin_y; o->w = d.width; o->h = d.height; }

extern "C" int me_emu_run(const SvtB200MeParams *P, const SvtB200Plane cur[3], const SvtB200Plane ref0[3], const SvtB200Plane ref1[3],
                          SvtB200MeCuResult *out, uint32_t *rcme, MeTraceDev *trace, int sb_begin, int sb_end, int win_bytes,
                          int force_global_interp) {
    MeKernelArgs A;
    memset(&A, 0, sizeof(A));
    A.P = *P;
    for (int i = 0; i < 3; i++) { to_dev(cur[i], &A.cur[i]); to_dev(ref0[i], &A.ref[0][i]); to_dev(ref1 ? ref1[i] : ref0[i], &A.ref[1][i]); }
    A.sbs_w = (P->luma_width + 63) / 64;
    A.n_sb  = A.sbs_w * ((P->luma_height + 63) / 64);
    A.out = out; A.rcme = rcme; A.trace = trace;
    A.win_bytes = win_bytes > 0 ? win_bytes : me_window_bytes(*P); /* 0: the size the host code would choose */
    A.interp_plane_bytes = me_interp_plane_bytes(*P);
    A.interp_in_smem = !force_global_interp;
    std::vector<uint8_t> scratch;
    if (force_global_interp) { scratch.resize((size_t)A.n_sb * 6 * me_align16(A.interp_plane_bytes)); A.interp_scratch = scratch.data(); }
    A.sb_first = sb_begin;
    A.n_work = sb_end - sb_begin;
    /* as the host code launches it: a persistent grid pulling super-blocks from a counter (two CTAs here), so that state carried
     * from one super-block of a CTA to its next is exercised too */
    uint32_t counter = 0;
    A.sb_counter = &counter;
    MeSmemLayout L = me_smem_layout(A.win_bytes, A.interp_plane_bytes, A.interp_in_smem, me_bip0_bytes(*P));
    emu_launch(dim3(A.n_work < 2 ? A.n_work : 2), dim3(ME_THREADS), L.total, [&] { me_sb_kernel(A); });
    return L.total;
}
extern "C" int me_emu_check_tables() {
    for (int i = 0; i < 64; i++) if (me_itab8[me_tab8[i]] != i) return 1;
    return 0;
}
