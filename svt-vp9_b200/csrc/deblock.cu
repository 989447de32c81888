/* deblock.cu -- host side of S2: svtb200_deblock_picture[_dev] (see include/svtb200.h). */
#include "common.cuh"
#include "deblock_kernel.cuh"

/* wavefront scratch (row ticket counter + per-row progress words): one set per context AND compute queue, so that two pictures
 * being filtered on different queues of a context never share a counter; launches on one queue are ordered by the queue itself */
struct LfScratch { int32_t *ticket = nullptr, *done = nullptr; size_t done_n = 0; int32_t gen = 0; };
struct LfScratchSet { LfScratch q[3]; };
static std::map<SvtB200Ctx *, LfScratchSet> g_lf;
static std::mutex g_lf_mu;
static unsigned long long *g_lf_trace; /* tools only */
extern "C" void svtb200_lf_set_trace(void *dev) { g_lf_trace = (unsigned long long *)dev; }

/* called by svtb200_destroy */
void svtb200_lf_release(SvtB200Ctx *c) {
    std::lock_guard<std::mutex> g(g_lf_mu);
    auto it = g_lf.find(c);
    if (it == g_lf.end()) return;
    for (int q = 0; q < 3; q++) {
        if (it->second.q[q].ticket) cudaFree(it->second.q[q].ticket);
        if (it->second.q[q].done) cudaFree(it->second.q[q].done);
    }
    g_lf.erase(it);
}

static int lf_launch(SvtB200Ctx *c, const SvtB200LfParams *p, uint8_t *y, int ys, uint8_t *u, uint8_t *v, int cs,
                     const SvtB200LfMask *lfm, int lfm_stride, const SvtB200LfThresh *thr) {
    if (!p || !y || !lfm || !thr || p->mi_rows <= 0 || p->mi_cols <= 0 || (!p->y_only && (!u || !v))) return SVTB200_EINVAL;
    if (!p->filter_level) return SVTB200_OK;
    LfKernelArgs A;
    memset(&A, 0, sizeof(A));
    A.P = *p; A.y = y; A.u = u; A.v = v; A.y_stride = ys; A.uv_stride = cs; A.lfm = lfm; A.lfm_stride = lfm_stride; A.thr = thr;
    A.sbs_w = (p->mi_cols + 7) >> 3; A.sbs_h = (p->mi_rows + 7) >> 3;
    if (A.sbs_w > 1023) return SVTB200_EINVAL;
    const size_t n = (size_t)A.sbs_h; /* one progress word (in a 128-byte line of its own) per super-block row */
    LfScratch *S;
    { std::lock_guard<std::mutex> g(g_lf_mu); S = &g_lf[c].q[svtb200_compute_index(c)]; }
    if (S->done_n < n) {
        if (S->done) cudaFree(S->done);
        if (!S->ticket && cudaMalloc((void **)&S->ticket, 4) != cudaSuccess) { cudaGetLastError(); return SVTB200_ENOMEM; }
        if (cudaMalloc((void **)&S->done, n * 4 * LF_FLAG_PAD) != cudaSuccess) { cudaGetLastError(); S->done = nullptr; S->done_n = 0; return SVTB200_ENOMEM; }
        S->done_n = n; S->gen = 0;
        SVTB200_CUDA_CHECK(c, cudaMemsetAsync(S->done, 0, n * 4 * LF_FLAG_PAD, c->stream));
    }
    if (++S->gen >= (1 << 20)) { /* the progress words carry the generation in their upper bits: restart well before it overflows */
        S->gen = 1;
        SVTB200_CUDA_CHECK(c, cudaMemsetAsync(S->done, 0, S->done_n * 4 * LF_FLAG_PAD, c->stream));
    }
    A.ticket = S->ticket; A.done = S->done; A.gen = S->gen; A.trace = g_lf_trace;
    SVTB200_CUDA_CHECK(c, cudaMemsetAsync(S->ticket, 0, 4, c->stream));
    const size_t dyn = (size_t)A.sbs_w * (sizeof(SvtB200LfMask) + 16);
    if (dyn > 48 * 1024) SVTB200_CUDA_CHECK(c, cudaFuncSetAttribute(lf_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    lf_row_kernel<<<(unsigned)n, LF_THREADS, dyn, c->stream>>>(A);
    c->launches++;
    SVTB200_CUDA_CHECK(c, cudaGetLastError());
    return SVTB200_OK;
}

extern "C" {

int svtb200_deblock_picture_dev(SvtB200Ctx *c, const SvtB200LfParams *p, uint8_t *y, int32_t ys, uint8_t *u, uint8_t *v, int32_t cs,
                                const SvtB200LfMask *lfm, int32_t lfm_stride, const SvtB200LfThresh *thr) {
    if (!c) return SVTB200_EINVAL;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    cudaSetDevice(c->device);
    return lf_launch(c, p, y, ys, u, v, cs, lfm, lfm_stride, thr);
}

int svtb200_deblock_picture(SvtB200Ctx *c, const SvtB200LfParams *p, uint8_t *y, int32_t ys, uint8_t *u, uint8_t *v, int32_t cs,
                            const SvtB200LfMask *lfm, int32_t lfm_stride, const SvtB200LfThresh thr[64]) {
    if (!c || !p || !y || !lfm || !thr) return SVTB200_EINVAL;
    if (!p->filter_level) return SVTB200_OK;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    cudaSetDevice(c->device);
    const int H = p->mi_rows * 8, W = p->mi_cols * 8, sbs_h = (p->mi_rows + 7) >> 3;
    const size_t yb = (size_t)H * W, cb = p->y_only ? 0 : (size_t)(H / 2) * (W / 2), mb = (size_t)lfm_stride * sbs_h * sizeof(SvtB200LfMask);
    int rc = svtb200_grow(c, (void **)&c->scratch, &c->scratch_bytes, yb + 2 * cb + mb + 64 * sizeof(SvtB200LfThresh) + 4096);
    if (rc) return rc;
    uint8_t *q = c->scratch;
    auto carve = [&](size_t b) { uint8_t *r = q; q += (b + 255) & ~(size_t)255; return r; };
    uint8_t *dy = carve(yb), *du = carve(cb), *dv = carve(cb);
    SvtB200LfMask *dm = (SvtB200LfMask *)carve(mb);
    SvtB200LfThresh *dt = (SvtB200LfThresh *)carve(64 * sizeof(SvtB200LfThresh));
    /* picture area only (the filter never touches samples outside it), packed pitch on the device */
    SVTB200_CUDA_CHECK(c, cudaMemcpy2DAsync(dy, W, y, ys, W, H, cudaMemcpyHostToDevice, c->stream));
    if (!p->y_only) {
        SVTB200_CUDA_CHECK(c, cudaMemcpy2DAsync(du, W / 2, u, cs, W / 2, H / 2, cudaMemcpyHostToDevice, c->stream));
        SVTB200_CUDA_CHECK(c, cudaMemcpy2DAsync(dv, W / 2, v, cs, W / 2, H / 2, cudaMemcpyHostToDevice, c->stream));
    }
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(dm, lfm, mb, cudaMemcpyHostToDevice, c->stream));
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(dt, thr, 64 * sizeof(SvtB200LfThresh), cudaMemcpyHostToDevice, c->stream));
    rc = lf_launch(c, p, dy, W, du, dv, W / 2, dm, lfm_stride, dt);
    if (rc) return rc;
    /* from here on filtered samples land in the caller's planes: a failure can no longer be answered with "filter it on the CPU" */
#define LF_WRITEBACK_CHECK(call)                                                                                   \
    do {                                                                                                           \
        cudaError_t e_ = (call);                                                                                   \
        if (e_ != cudaSuccess) { c->err = std::string("deblock write-back: ") + cudaGetErrorString(e_); return SVTB200_EPARTIAL; } \
    } while (0)
    LF_WRITEBACK_CHECK(cudaMemcpy2DAsync(y, ys, dy, W, W, H, cudaMemcpyDeviceToHost, c->stream));
    if (!p->y_only) {
        LF_WRITEBACK_CHECK(cudaMemcpy2DAsync(u, cs, du, W / 2, W / 2, H / 2, cudaMemcpyDeviceToHost, c->stream));
        LF_WRITEBACK_CHECK(cudaMemcpy2DAsync(v, cs, dv, W / 2, W / 2, H / 2, cudaMemcpyDeviceToHost, c->stream));
    }
    LF_WRITEBACK_CHECK(cudaStreamSynchronize(c->stream));
#undef LF_WRITEBACK_CHECK
    return SVTB200_OK;
}

} /* extern "C" */
