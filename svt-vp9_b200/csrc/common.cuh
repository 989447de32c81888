/* common.cuh -- context object and error plumbing shared by the C-ABI translation units. */
#ifndef SVTB200_COMMON_CUH
#define SVTB200_COMMON_CUH
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include "../../include/svtb200.h"

#define SVTB200_MARKER_RING 1024
#define SVTB200_DEV_PAD_L 80   /* svtb200_picture_upload_luma: device columns left of the picture (multiple of 16, >= the reference's 68) */
#define SVTB200_DEV_PAD_R 96   /* ... and right of it: the reference's 68 + the samples its windows read past the row end */

struct DevTensorMap { /* a cached CUtensorMap of one plane for one box shape (me.cu) */
    alignas(64) uint64_t opaque[16];
    const uint8_t *buf = nullptr;
    int32_t stride = 0, rows = 0, box_w = 0, box_h = 0;
};
struct DevPlane {
    uint8_t *buf = nullptr;
    int32_t  stride = 0, ox = 0, oy = 0, w = 0, h = 0, rows = 0;
    DevTensorMap tm[2]; /* two box shapes per plane at most (source super-block / search window) */
};
struct DevPicture {
    DevPlane    pl[3];            /* full, quarter, sixteenth */
    cudaEvent_t ready = nullptr;  /* recorded on the upload stream after the picture's copies / front-end kernels */
    cudaEvent_t used[3] = {nullptr, nullptr, nullptr}; /* one per compute queue: recorded after the last kernel on that queue that reads the picture */
};

#define SVTB200_STAGE_SLOTS 8
struct StageSlot { /* svtb200_picture_upload_luma_staged: one page-locked picture area */
    uint8_t    *buf = nullptr;
    size_t      cap = 0;
    cudaEvent_t done = nullptr; /* the transfer out of the slot has finished */
    std::mutex  mu;
};

struct ScanTable {
    int16_t *scan = nullptr, *iscan = nullptr;
    int      n = 0;
};

struct SvtB200Ctx {
    int          device = 0;
    cudaStream_t stream = nullptr;              /* the compute queue kernels are launched on (one of compute[]) */
    cudaStream_t compute[3] = {nullptr, nullptr, nullptr};
    cudaStream_t up = nullptr, down = nullptr;  /* host->device and device->host copy streams */
    cudaEvent_t  markers[SVTB200_MARKER_RING] = {};
    uint64_t     next_ticket = 1;
    std::recursive_mutex mu; /* recursive: the host-pointer entry points hold it across their copy -> _dev call -> copy sequence */
    std::string  err;
    uint64_t     launches = 0;
    int          smem_optin = 0, sm_count = 0;
    std::map<uint64_t, DevPicture> pics;
    uint32_t *me_counter = nullptr; int me_counter_next = 0; /* ring of super-block counters of the persistent ME grid, one per launch in flight */
    /* ME scratch, grown on demand */
    uint8_t *interp_scratch = nullptr; size_t interp_scratch_bytes = 0;
    void    *me_out = nullptr;  size_t me_out_bytes = 0;
    uint32_t *me_rcme = nullptr; size_t me_rcme_bytes = 0;
    void    *me_trace = nullptr; size_t me_trace_bytes = 0; int me_trace_enabled = 0;
    /* pinned staging */
    uint8_t *pin = nullptr; size_t pin_bytes = 0;
    StageSlot stage[SVTB200_STAGE_SLOTS]; std::atomic<unsigned> stage_next{0};
    /* transform / quant */
    ScanTable scans[4][4];
    int16_t  *scan_nb[4][4] = {};  /* neighbour tables of the scans (rate estimation) */
    uint8_t  *rate_tables = nullptr; /* value -> token / cost tables (rate.cu) */
    cudaEvent_t chain_ev = nullptr;  /* svtb200_inter_chain_dev: "behind the other queue" */
    uint32_t *pack_scratch[3] = {nullptr, nullptr, nullptr}; int pack_chunks[3] = {0, 0, 0}; /* pack.cu: ticket + chunk bases per compute queue */
    /* generic device scratch for the host-pointer entry points */
    uint8_t *scratch = nullptr; size_t scratch_bytes = 0;
    /* prediction descriptors of the batch in flight (pred.cu); pred_read: the kernel that reads them has finished */
    void *pred_desc = nullptr; size_t pred_desc_bytes = 0; cudaEvent_t pred_read = nullptr;
};

#define SVTB200_CUDA_CHECK(ctx, call)                                                                 \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            char b_[512];                                                                             \
            snprintf(b_, sizeof(b_), "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            (ctx)->err = b_;                                                                          \
            return SVTB200_ECUDA;                                                                     \
        }                                                                                             \
    } while (0)

static inline int svtb200_grow(SvtB200Ctx *ctx, void **p, size_t *cap, size_t need) {
    if (*cap >= need) return 0;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    size_t n = need + need / 4 + 256;
    if (cudaMalloc(p, n) != cudaSuccess) { ctx->err = "cudaMalloc failed"; cudaGetLastError(); return SVTB200_ENOMEM; }
    *cap = n;
    return 0;
}
static inline int svtb200_grow_pinned(SvtB200Ctx *ctx, size_t need) {
    if (ctx->pin_bytes >= need) return 0;
    if (ctx->pin) cudaFreeHost(ctx->pin);
    ctx->pin = nullptr; ctx->pin_bytes = 0;
    size_t n = need + need / 4 + 256;
    if (cudaMallocHost((void **)&ctx->pin, n) != cudaSuccess) { ctx->err = "cudaMallocHost failed"; cudaGetLastError(); return SVTB200_ENOMEM; }
    ctx->pin_bytes = n;
    return 0;
}
void svtb200_lf_release(SvtB200Ctx *c); /* deblock.cu: per-context wavefront scratch */

static inline cudaStream_t svtb200_pick_stream(SvtB200Ctx *c, int which) {
    return which == SVTB200_STREAM_UPLOAD ? c->up : which == SVTB200_STREAM_DOWNLOAD ? c->down : which == SVTB200_STREAM_COMPUTE2 ? c->compute[1] : which == SVTB200_STREAM_COMPUTE3 ? c->compute[2] : c->compute[0];
}
/* a kernel about to read picture p on the compute stream: wait for its upload; afterwards note the use */
static inline void svtb200_pic_acquire(SvtB200Ctx *c, DevPicture &p) { if (p.ready) cudaStreamWaitEvent(c->stream, p.ready, 0); }
static inline int svtb200_compute_index(const SvtB200Ctx *c) { return c->stream == c->compute[1] ? 1 : c->stream == c->compute[2] ? 2 : 0; }
static inline void svtb200_pic_release(SvtB200Ctx *c, DevPicture &p) {
    cudaEvent_t &e = p.used[svtb200_compute_index(c)];
    if (!e) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    cudaEventRecord(e, c->stream);
}
#endif
