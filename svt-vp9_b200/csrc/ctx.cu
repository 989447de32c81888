/* ctx.cu -- context life cycle, memory helpers and the picture cache of the C ABI. */
#include <stdlib.h>
#include "common.cuh"

extern "C" {

int svtb200_create(int device, SvtB200Ctx **out) {
    if (!out) return SVTB200_EINVAL;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) { cudaGetLastError(); return SVTB200_ENODEV; }
    /* SVTB200_BLOCKING_SYNC=1: host threads sleep in the synchronising calls instead of spinning (hosts whose cores are all
     * busy with the encoder's own threads); must be chosen before the device's context exists */
    const char *bs = getenv("SVTB200_BLOCKING_SYNC");
    if (bs && bs[0] == '1') { cudaSetDeviceFlags(cudaDeviceScheduleBlockingSync); cudaGetLastError(); }
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return SVTB200_ENODEV; }
    SvtB200Ctx *c = new SvtB200Ctx();
    c->device = device;
    /* the upload queue also runs the (tiny) padding / decimation kernels: give it priority over the picture-sized compute kernels */
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    /* SVTB200_COMPUTE2_PRIORITY=1: CTAs of the second compute queue are dispatched ahead of the first queue's whenever an SM has room */
    const char *p2 = getenv("SVTB200_COMPUTE2_PRIORITY");
    if (cudaStreamCreateWithFlags(&c->compute[0], cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithPriority(&c->compute[1], cudaStreamNonBlocking, (p2 && p2[0] == '1') ? prio_hi : prio_lo) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->compute[2], cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithPriority(&c->up, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaStreamCreateWithPriority(&c->down, cudaStreamNonBlocking, prio_hi) != cudaSuccess) { delete c; cudaGetLastError(); return SVTB200_ENODEV; }
    c->stream = c->compute[0];
    cudaDeviceGetAttribute(&c->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    *out = c;
    return SVTB200_OK;
}

void svtb200_destroy(SvtB200Ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->up); cudaStreamSynchronize(c->compute[0]); cudaStreamSynchronize(c->compute[1]); cudaStreamSynchronize(c->compute[2]); cudaStreamSynchronize(c->down);
    for (auto &kv : c->pics) {
        for (int i = 0; i < 3; i++) if (kv.second.pl[i].buf) cudaFree(kv.second.pl[i].buf);
        if (kv.second.ready) cudaEventDestroy(kv.second.ready);
        for (int q = 0; q < 3; q++) if (kv.second.used[q]) cudaEventDestroy(kv.second.used[q]);
    }
    for (int i = 0; i < SVTB200_MARKER_RING; i++) if (c->markers[i]) cudaEventDestroy(c->markers[i]);
    svtb200_lf_release(c);
    if (c->interp_scratch) cudaFree(c->interp_scratch);
    if (c->me_out) cudaFree(c->me_out);
    if (c->me_rcme) cudaFree(c->me_rcme);
    if (c->me_counter) cudaFree(c->me_counter);
    if (c->me_trace) cudaFree(c->me_trace);
    if (c->scratch) cudaFree(c->scratch);
    if (c->pred_desc) cudaFree(c->pred_desc);
    if (c->pred_read) cudaEventDestroy(c->pred_read);
    if (c->pin) cudaFreeHost(c->pin);
    for (StageSlot &st : c->stage) { if (st.buf) cudaFreeHost(st.buf); if (st.done) cudaEventDestroy(st.done); }
    for (int a = 0; a < 4; a++) for (int b = 0; b < 4; b++) { if (c->scans[a][b].scan) cudaFree(c->scans[a][b].scan); if (c->scans[a][b].iscan) cudaFree(c->scans[a][b].iscan); if (c->scan_nb[a][b]) cudaFree(c->scan_nb[a][b]); }
    if (c->rate_tables) cudaFree(c->rate_tables);
    if (c->chain_ev) cudaEventDestroy(c->chain_ev);
    for (int q = 0; q < 3; q++) if (c->pack_scratch[q]) cudaFree(c->pack_scratch[q]);
    cudaStreamDestroy(c->compute[0]); cudaStreamDestroy(c->compute[1]); cudaStreamDestroy(c->compute[2]); cudaStreamDestroy(c->up); cudaStreamDestroy(c->down);
    delete c;
}

const char *svtb200_last_error(const SvtB200Ctx *c) { return c ? c->err.c_str() : "null context"; }
uint64_t svtb200_launch_count(const SvtB200Ctx *c) { return c ? c->launches : 0; }
void *svtb200_stream(SvtB200Ctx *c) { return c ? (void *)c->stream : nullptr; }

int svtb200_sync(SvtB200Ctx *c) {
    if (!c) return SVTB200_EINVAL;
    SVTB200_CUDA_CHECK(c, cudaStreamSynchronize(c->up));
    SVTB200_CUDA_CHECK(c, cudaStreamSynchronize(c->compute[0]));
    SVTB200_CUDA_CHECK(c, cudaStreamSynchronize(c->compute[1]));
    SVTB200_CUDA_CHECK(c, cudaStreamSynchronize(c->compute[2]));
    SVTB200_CUDA_CHECK(c, cudaStreamSynchronize(c->down));
    return SVTB200_OK;
}
int svtb200_select_compute(SvtB200Ctx *c, int which) {
    if (!c || (which != SVTB200_STREAM_COMPUTE && which != SVTB200_STREAM_COMPUTE2 && which != SVTB200_STREAM_COMPUTE3)) return SVTB200_EINVAL;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    c->stream = svtb200_pick_stream(c, which);
    return SVTB200_OK;
}

/* ---- the three in-order queues of a context and the markers that order work between them ---- */
int svtb200_memcpy_h2d_async(SvtB200Ctx *c, void *d, const void *h, size_t n) {
    if (!c) return SVTB200_EINVAL;
    cudaSetDevice(c->device);
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, c->up));
    return SVTB200_OK;
}
int svtb200_memcpy_h2d_2d_async(SvtB200Ctx *c, void *d, size_t d_pitch, const void *h, size_t h_pitch, size_t width_bytes, size_t rows) {
    if (!c || !d || !h || width_bytes > d_pitch || width_bytes > h_pitch) return SVTB200_EINVAL;
    cudaSetDevice(c->device);
    SVTB200_CUDA_CHECK(c, cudaMemcpy2DAsync(d, d_pitch, h, h_pitch, width_bytes, rows, cudaMemcpyHostToDevice, c->up));
    return SVTB200_OK;
}
int svtb200_memcpy_d2h_async(SvtB200Ctx *c, void *h, const void *d, size_t n) {
    if (!c) return SVTB200_EINVAL;
    cudaSetDevice(c->device);
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, c->down));
    return SVTB200_OK;
}
int svtb200_marker(SvtB200Ctx *c, int stream, uint64_t *ticket) {
    if (!c || !ticket || stream < 0 || stream > SVTB200_STREAM_COMPUTE3) return SVTB200_EINVAL;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    const uint64_t t = c->next_ticket++;
    cudaEvent_t &e = c->markers[t % SVTB200_MARKER_RING];
    if (!e) SVTB200_CUDA_CHECK(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    SVTB200_CUDA_CHECK(c, cudaEventRecord(e, svtb200_pick_stream(c, stream)));
    *ticket = t;
    return SVTB200_OK;
}
static int marker_event(SvtB200Ctx *c, uint64_t ticket, cudaEvent_t *e) {
    if (ticket == 0 || ticket >= c->next_ticket) return SVTB200_EINVAL;
    if (c->next_ticket - ticket > SVTB200_MARKER_RING) { /* its event has been reused for a newer marker: no ordering can be promised any more */
        c->err = "marker ticket too old (more than SVTB200_MARKER_RING markers ago)";
        return SVTB200_EINVAL;
    }
    *e = c->markers[ticket % SVTB200_MARKER_RING];
    return SVTB200_OK;
}
int svtb200_stream_wait(SvtB200Ctx *c, int stream, uint64_t ticket) {
    if (!c || stream < 0 || stream > SVTB200_STREAM_COMPUTE3) return SVTB200_EINVAL;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    cudaEvent_t e;
    int rc = marker_event(c, ticket, &e);
    if (rc) return rc;
    if (e) SVTB200_CUDA_CHECK(c, cudaStreamWaitEvent(svtb200_pick_stream(c, stream), e, 0));
    return SVTB200_OK;
}
int svtb200_fence(SvtB200Ctx *c, int from, int to) {
    uint64_t t;
    int rc = svtb200_marker(c, from, &t);
    return rc ? rc : svtb200_stream_wait(c, to, t);
}
int svtb200_wait(SvtB200Ctx *c, uint64_t ticket) {
    if (!c) return SVTB200_EINVAL;
    cudaEvent_t e;
    {
        std::lock_guard<std::recursive_mutex> g(c->mu);
        int rc = marker_event(c, ticket, &e);
        if (rc) return rc;
    }
    if (e) SVTB200_CUDA_CHECK(c, cudaEventSynchronize(e));
    return SVTB200_OK;
}

void *svtb200_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void svtb200_host_free(void *p) { if (p) cudaFreeHost(p); }
int svtb200_host_register(void *p, size_t bytes) {
    if (!p || !bytes) return SVTB200_EINVAL;
    const cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
    if (e == cudaSuccess) return SVTB200_OK;
    cudaGetLastError();
    return e == cudaErrorHostMemoryAlreadyRegistered ? SVTB200_OK : SVTB200_ECUDA;
}
int svtb200_host_unregister(void *p) {
    if (!p) return SVTB200_EINVAL;
    if (cudaHostUnregister(p) != cudaSuccess) { cudaGetLastError(); return SVTB200_ECUDA; }
    return SVTB200_OK;
}

void *svtb200_dev_alloc(SvtB200Ctx *c, size_t bytes) {
    if (!c) return nullptr;
    cudaSetDevice(c->device);
    void *p = nullptr;
    if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); c->err = "cudaMalloc failed"; return nullptr; }
    return p;
}
void svtb200_dev_free(SvtB200Ctx *c, void *p) { if (c && p) { cudaSetDevice(c->device); cudaFree(p); } }

int svtb200_memcpy_h2d(SvtB200Ctx *c, void *d, const void *h, size_t n) {
    if (!c) return SVTB200_EINVAL;
    cudaStream_t s;
    { std::lock_guard<std::recursive_mutex> g(c->mu); cudaSetDevice(c->device); s = c->stream; }
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s));
    SVTB200_CUDA_CHECK(c, cudaStreamSynchronize(s));
    return SVTB200_OK;
}
int svtb200_memcpy_d2h(SvtB200Ctx *c, void *h, const void *d, size_t n) {
    if (!c) return SVTB200_EINVAL;
    cudaStream_t s;
    { std::lock_guard<std::recursive_mutex> g(c->mu); cudaSetDevice(c->device); s = c->stream; }
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s));
    SVTB200_CUDA_CHECK(c, cudaStreamSynchronize(s));
    return SVTB200_OK;
}

static int upload_plane(SvtB200Ctx *c, const SvtB200Plane *src, DevPlane *dst) {
    if (!src || !src->buf) { if (dst->buf) cudaFree(dst->buf); *dst = DevPlane(); return 0; }
    if (src->stride <= 0 || src->padded_height <= 0) return SVTB200_EINVAL;
    const size_t bytes = (size_t)src->stride * src->padded_height;
    if (!dst->buf || (size_t)dst->stride * dst->rows < bytes) {
        if (dst->buf) cudaFree(dst->buf);
        dst->buf = nullptr;
        /* slack: window staging may read a few bytes past the last padded row */
        if (cudaMalloc((void **)&dst->buf, bytes + 256) != cudaSuccess) { cudaGetLastError(); c->err = "cudaMalloc(plane) failed"; return SVTB200_ENOMEM; }
        cudaMemsetAsync(dst->buf + bytes, 0, 256, c->up);
    }
    dst->stride = src->stride; dst->ox = src->origin_x; dst->oy = src->origin_y; dst->w = src->width; dst->h = src->height;
    dst->rows = src->padded_height;
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(dst->buf, src->buf, bytes, cudaMemcpyHostToDevice, c->up));
    return 0;
}
/* uploads run on the upload queue: after every kernel that still reads the picture's previous contents, and
 * followed by the event the compute queue waits on before it touches the new contents */
static void upload_begin(SvtB200Ctx *c, DevPicture &p) { for (int q = 0; q < 3; q++) if (p.used[q]) cudaStreamWaitEvent(c->up, p.used[q], 0); }
static int upload_end(SvtB200Ctx *c, DevPicture &p) {
    if (!p.ready) SVTB200_CUDA_CHECK(c, cudaEventCreateWithFlags(&p.ready, cudaEventDisableTiming));
    SVTB200_CUDA_CHECK(c, cudaEventRecord(p.ready, c->up));
    return SVTB200_OK;
}

int svtb200_picture_upload(SvtB200Ctx *c, uint64_t pic_id, const SvtB200Plane *full, const SvtB200Plane *quarter,
                           const SvtB200Plane *sixteenth) {
    if (!c || !full || !full->buf) return SVTB200_EINVAL;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    cudaSetDevice(c->device);
    DevPicture &p = c->pics[pic_id];
    int rc;
    upload_begin(c, p);
    if ((rc = upload_plane(c, full, &p.pl[0]))) return rc;
    if ((rc = upload_plane(c, quarter, &p.pl[1]))) return rc;
    if ((rc = upload_plane(c, sixteenth, &p.pl[2]))) return rc;
    return upload_end(c, p);
}

#include "pa_front_kernel.cuh"

static int alloc_plane(SvtB200Ctx *c, DevPlane *dst, int stride, int rows) {
    const size_t bytes = (size_t)stride * rows;
    if (!dst->buf || (size_t)dst->stride * dst->rows < bytes) {
        if (dst->buf) cudaFree(dst->buf);
        dst->buf = nullptr;
        if (cudaMalloc((void **)&dst->buf, bytes + 256) != cudaSuccess) { cudaGetLastError(); c->err = "cudaMalloc(plane) failed"; return SVTB200_ENOMEM; }
        cudaMemsetAsync(dst->buf + bytes, 0, 256, c->up);
    }
    dst->stride = stride; dst->rows = rows;
    return 0;
}

static int upload_luma_args_ok(const SvtB200Ctx *c, const SvtB200Plane *full, int32_t quarter_pad, int32_t sixteenth_pad) {
    if (!c || !full || !full->buf || full->stride <= 0 || full->padded_height <= 0 || quarter_pad < 0 || sixteenth_pad < 0) return 0;
    return !((full->stride & 3) || full->width <= 0 || full->height <= 0);
}
/* area / area_pitch: where the picture area's first sample is read from (the caller's plane itself, or a page-locked staging copy of it);
 * the geometry of the device plane always follows 'full'.  Called with c->mu held. */
static int upload_luma_locked(SvtB200Ctx *c, uint64_t pic_id, const SvtB200Plane *full, const uint8_t *area, size_t area_pitch, int32_t quarter_pad,
                              int32_t sixteenth_pad, int make_quarter) {
    cudaSetDevice(c->device);
    DevPicture &p = c->pics[pic_id];
    int rc;
    upload_begin(c, p);
    DevPlane *f = &p.pl[0];
    /* Device geometry of the full-resolution plane: the picture starts SVTB200_DEV_PAD_L columns into a row whose pitch is a multiple
     * of 16, so that every 16-sample group of the picture is a 16-byte aligned vector (the transform kernels' loads) and a row is a
     * legal TMA row (the motion-estimation windows); SVTB200_DEV_PAD_R columns follow the picture.  Columns beyond the host row's own
     * padding carry the bytes a read past the host row would find (pa_pad_full_kernel). */
    const int dpitch = (full->width + SVTB200_DEV_PAD_L + SVTB200_DEV_PAD_R + 15) & ~15;
    if (full->origin_x > SVTB200_DEV_PAD_L) { c->err = "upload_luma: host padding wider than the device layout"; return SVTB200_EINVAL; }
    if ((rc = alloc_plane(c, f, dpitch, full->padded_height))) return rc;
    f->ox = SVTB200_DEV_PAD_L; f->oy = full->origin_y; f->w = full->width; f->h = full->height;
    /* only the picture area crosses PCIe; the padding is regenerated on the device */
    SVTB200_CUDA_CHECK(c, cudaMemcpy2DAsync(f->buf + (size_t)f->oy * f->stride + f->ox, f->stride, area, area_pitch, f->w, f->h, cudaMemcpyHostToDevice, c->up));
    dim3 blk(128);
    pa_pad_full_kernel<<<dim3((f->stride / 4 + 127) / 128, f->rows), blk, 0, c->up>>>(f->buf, f->stride, f->ox, f->oy, f->w, f->h, f->rows, f->stride,
                                                                                     full->origin_x);
    c->launches++;
    for (int lvl = 1; lvl <= 2; lvl++) {
        DevPlane *d = &p.pl[lvl];
        if (lvl == 1 && !make_quarter) { if (d->buf) cudaFree(d->buf); *d = DevPlane(); continue; }
        const int step = lvl == 1 ? 2 : 4, pad = lvl == 1 ? quarter_pad : sixteenth_pad;
        const int w = (f->w + step - 1) / step, h = (f->h + step - 1) / step;
        const int stride = (w + 2 * pad + 15) & ~15; /* multiple of 16: word stores here, legal TMA rows for the HME windows */
        if ((rc = alloc_plane(c, d, stride, h + 2 * pad))) return rc;
        d->ox = pad; d->oy = pad; d->w = w; d->h = h;
        pa_decimate_kernel<<<dim3((stride / 4 + 127) / 128, d->rows), blk, 0, c->up>>>(f->buf, f->stride, f->ox, f->oy, d->buf, stride, pad, pad, w, h,
                                                                                     d->rows, step);
        c->launches++;
    }
    SVTB200_CUDA_CHECK(c, cudaGetLastError());
    return upload_end(c, p);
}

int svtb200_picture_upload_luma(SvtB200Ctx *c, uint64_t pic_id, const SvtB200Plane *full, int32_t quarter_pad, int32_t sixteenth_pad,
                                int make_quarter) {
    if (!upload_luma_args_ok(c, full, quarter_pad, sixteenth_pad)) return SVTB200_EINVAL;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    return upload_luma_locked(c, pic_id, full, full->buf + (size_t)full->origin_y * full->stride + full->origin_x, (size_t)full->stride, quarter_pad,
                              sixteenth_pad, make_quarter);
}

/* The same from PAGEABLE host memory (the reference's picture pool as it is): the picture area is first copied, by the calling thread, into a
 * slot of a small ring of page-locked buffers the context owns, and the DMA transfer runs from there.  Page-locking the caller's buffers
 * instead (svtb200_host_register) costs ~1 ms per MB the first time each pool object is seen -- at 3840x2160 that was 10 ms per picture on
 * clips as short as the pool is deep -- while the staging copy is a plain memcpy of the picture area (no padding) whenever a picture is
 * uploaded.  Several threads may stage different pictures at the same time (one slot each); the context lock is taken only to queue the work. */
int svtb200_picture_upload_luma_staged(SvtB200Ctx *c, uint64_t pic_id, const SvtB200Plane *full, int32_t quarter_pad, int32_t sixteenth_pad,
                                       int make_quarter) {
    if (!upload_luma_args_ok(c, full, quarter_pad, sixteenth_pad)) return SVTB200_EINVAL;
    StageSlot &s = c->stage[c->stage_next.fetch_add(1) % SVTB200_STAGE_SLOTS];
    std::lock_guard<std::mutex> sl(s.mu);
    cudaSetDevice(c->device);
    if (s.done) SVTB200_CUDA_CHECK(c, cudaEventSynchronize(s.done)); /* the transfer that last read this slot */
    const size_t w = (size_t)full->width, h = (size_t)full->height;
    if (s.cap < w * h) {
        if (s.buf) cudaFreeHost(s.buf);
        s.buf = nullptr; s.cap = 0;
        if (cudaMallocHost((void **)&s.buf, w * h) != cudaSuccess) { cudaGetLastError(); c->err = "cudaMallocHost(staging slot) failed"; return SVTB200_ENOMEM; }
        s.cap = w * h;
    }
    const uint8_t *src = full->buf + (size_t)full->origin_y * full->stride + full->origin_x;
    for (size_t y = 0; y < h; y++) memcpy(s.buf + y * w, src + y * (size_t)full->stride, w);
    std::lock_guard<std::recursive_mutex> g(c->mu);
    const int rc = upload_luma_locked(c, pic_id, full, s.buf, w, quarter_pad, sixteenth_pad, make_quarter);
    if (!s.done) SVTB200_CUDA_CHECK(c, cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming | cudaEventBlockingSync));
    SVTB200_CUDA_CHECK(c, cudaEventRecord(s.done, c->up));
    return rc;
}

int svtb200_picture_plane_dev(SvtB200Ctx *c, uint64_t pic_id, int which, SvtB200Plane *out) {
    if (!c || !out || which < 0 || which > 2) return SVTB200_EINVAL;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    auto it = c->pics.find(pic_id);
    if (it == c->pics.end() || !it->second.pl[which].buf) return SVTB200_ENOENT;
    svtb200_pic_acquire(c, it->second); /* work queued on the compute stream after this call sees the uploaded contents */
    const DevPlane &d = it->second.pl[which];
    out->buf = d.buf; out->stride = d.stride; out->origin_x = d.ox; out->origin_y = d.oy; out->width = d.w; out->height = d.h;
    out->padded_height = d.rows;
    return SVTB200_OK;
}

/* eb_vp9_generate_padding (Source/Lib/Codec/EbMcp.c:17) on a device plane: horizontal then vertical edge replication */
int svtb200_pad_plane_dev(SvtB200Ctx *c, uint8_t *buf, int32_t stride, int32_t width, int32_t height, int32_t pad_x, int32_t pad_y) {
    if (!c || !buf || stride <= 0 || width <= 0 || height <= 0 || pad_x < 0 || pad_y < 0 || width + 2 * pad_x > stride) return SVTB200_EINVAL;
    if ((stride & 3) || ((uintptr_t)buf & 3) || ((width + 2 * pad_x) & 3)) { c->err = "pad_plane: stride, base and padded width must be multiples of 4"; return SVTB200_EINVAL; }
    std::lock_guard<std::recursive_mutex> g(c->mu);
    cudaSetDevice(c->device);
    const int cols = width + 2 * pad_x, rows = height + 2 * pad_y;
    /* the kernel treats 'cols' as the row length to write and 'stride' as the pitch */
    pa_pad_full_kernel<<<dim3((cols / 4 + 127) / 128, rows), dim3(128), 0, c->stream>>>(buf, stride, pad_x, pad_y, width, height, rows, cols, pad_x);
    c->launches++;
    SVTB200_CUDA_CHECK(c, cudaGetLastError());
    return SVTB200_OK;
}

int svtb200_picture_touch(SvtB200Ctx *c, uint64_t pic_id) {
    if (!c) return SVTB200_EINVAL;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    auto it = c->pics.find(pic_id);
    if (it == c->pics.end()) return SVTB200_ENOENT;
    cudaSetDevice(c->device);
    svtb200_pic_release(c, it->second);
    return SVTB200_OK;
}

int svtb200_picture_release(SvtB200Ctx *c, uint64_t pic_id) {
    if (!c) return SVTB200_EINVAL;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    auto it = c->pics.find(pic_id);
    if (it == c->pics.end()) return SVTB200_ENOENT;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->up); cudaStreamSynchronize(c->compute[0]); cudaStreamSynchronize(c->compute[1]); cudaStreamSynchronize(c->compute[2]);
    for (int i = 0; i < 3; i++) if (it->second.pl[i].buf) cudaFree(it->second.pl[i].buf);
    if (it->second.ready) cudaEventDestroy(it->second.ready);
    for (int q = 0; q < 3; q++) if (it->second.used[q]) cudaEventDestroy(it->second.used[q]);
    c->pics.erase(it);
    return SVTB200_OK;
}

} /* extern "C" */
