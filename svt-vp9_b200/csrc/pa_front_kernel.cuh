/* pa_front_kernel.cuh -- device code of the picture-analysis front end in ctx.cu (edge replication of the full-resolution plane in the device's own
 * geometry, point decimation + padding of the 1/4 and 1/16 planes); a header of its own so that tests/emu/pa_emu.cpp can run the same source on the CPU. */
#ifndef SVTB200_PA_FRONT_KERNEL_CUH
#define SVTB200_PA_FRONT_KERNEL_CUH
#include <stdint.h>
#include <stddef.h>

/* Edge replication of the padded source plane (eb_vp9_generate_padding, EbPictureOperators.c) and the point decimations
 * of eb_vp9_decimation_2d (EbPictureAnalysisProcess.c:102) with their own padding, one thread per 4 output bytes. */
__global__ void pa_pad_full_kernel(uint8_t *buf, int pitch, int ox, int oy, int w, int h, int rows, int cols /* bytes of a row to write */,
                                   int host_pad /* padding of the HOST row on each side */) {
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y;
    if (x4 >= cols || y >= rows) return;
    if (y >= oy && y < oy + h && x4 >= ox && x4 + 4 <= ox + w) return; /* interior word */
    /* Columns further out than the host row's own padding hold what a read past the end of the HOST row finds: the padded
     * planes of the reference are contiguous (pitch w + 2 * host_pad), so column w + host_pad + k of row y is column -host_pad + k
     * of row y + 1, and column -host_pad - k is the tail of row y - 1.  The reference's search windows do run a few samples over
     * (EbMotionEstimation.c:855-856 matches two 16x16 blocks sixteen samples to the right); a device pitch of its own must
     * reproduce those bytes. */
    const int hp = w + 2 * host_pad;
    uint32_t v = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        int xs = x4 + k - ox, yy = y;      /* picture-relative column, padded-plane row */
        if (xs < -host_pad || xs >= w + host_pad) {
            const int a = xs + host_pad;     /* offset from the start of host row yy */
            const int dr = a >= 0 ? a / hp : -((-a + hp - 1) / hp);
            yy += dr; xs = a - dr * hp - host_pad;
        }
        const int sy = min(max(yy - oy, 0), h - 1);
        v |= (uint32_t)buf[(size_t)(oy + sy) * pitch + ox + min(max(xs, 0), w - 1)] << (8 * k);
    }
    if (x4 + 4 <= cols) *(uint32_t *)(buf + (size_t)y * pitch + x4) = v;
    else for (int k = 0; x4 + k < cols; k++) buf[(size_t)y * pitch + x4 + k] = (uint8_t)(v >> (8 * k));
}
__global__ void pa_decimate_kernel(const uint8_t *full, int fstride, int fox, int foy, uint8_t *dst, int stride, int ox, int oy,
                                   int w, int h, int rows, int step) {
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y;
    if (x4 >= stride || y >= rows) return;
    const uint8_t *row = full + (size_t)(foy + min(max(y - oy, 0), h - 1) * step) * fstride + fox;
    uint32_t v = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) v |= (uint32_t)row[min(max(x4 + k - ox, 0), w - 1) * step] << (8 * k);
    if (x4 + 4 <= stride) *(uint32_t *)(dst + (size_t)y * stride + x4) = v;
    else for (int k = 0; x4 + k < stride; k++) dst[(size_t)y * stride + x4 + k] = (uint8_t)(v >> (8 * k));
}
#endif
