// CPU execution of the per-block arithmetic of blocks.cu (csrc/blocks_device.cuh: plain host + device functions), thread by thread as the two
// kernels index it (test infrastructure only).
#include "../../svt-vp9_b200/csrc/blocks_device.cuh"

extern "C" int blocks_emu_expand(const SvtB200InterBlock *blk, int n, const SvtB200PredPicture *pic, SvtB200PredBlock *out) {
    for (int i = 0; i < 3 * n; i++) { const int b = i / 3, plane = i - 3 * b; out[i] = blocks_expand_one(blk[b], plane, *pic); } // inter_expand_kernel
    return 0;
}

extern "C" int blocks_emu_me_best(const SvtB200MeCuResult *me, int width, int height, SvtB200InterBlock *out) {
    const int n = blocks_tile_count(width, height);
    for (int i = 0; i < n; i++) { // me_best_blocks_kernel
        int x, y, size;
        blocks_tile_geometry(width, height, i, &x, &y, &size);
        const int sbs_w = (width + 63) >> 6, sb = (y >> 6) * sbs_w + (x >> 6);
        const int pu = size == 32 ? 1 + ((y & 32) >> 5) * 2 + ((x & 32) >> 5) : 5 + ((y & 48) >> 4) * 4 + ((x & 48) >> 4);
        out[i] = blocks_from_me(me[(size_t)sb * 85 + pu], x, y, size);
    }
    return n;
}
