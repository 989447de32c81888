/* ctx.cu -- context life cycle, memory helpers and the picture cache of the C ABI. */
#include "common.cuh"

extern "C" {

int svtb200_create(int device, SvtB200Ctx **out) {
    if (!out) return SVTB200_EINVAL;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) { cudaGetLastError(); return SVTB200_ENODEV; }
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return SVTB200_ENODEV; }
    SvtB200Ctx *c = new SvtB200Ctx();
    c->device = device;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; cudaGetLastError(); return SVTB200_ENODEV; }
    cudaDeviceGetAttribute(&c->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    *out = c;
    return SVTB200_OK;
}

void svtb200_destroy(SvtB200Ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (auto &kv : c->pics) for (int i = 0; i < 3; i++) if (kv.second.pl[i].buf) cudaFree(kv.second.pl[i].buf);
    if (c->interp_scratch) cudaFree(c->interp_scratch);
    if (c->me_out) cudaFree(c->me_out);
    if (c->me_rcme) cudaFree(c->me_rcme);
    if (c->me_trace) cudaFree(c->me_trace);
    if (c->scratch) cudaFree(c->scratch);
    if (c->pin) cudaFreeHost(c->pin);
    for (int a = 0; a < 4; a++) for (int b = 0; b < 4; b++) { if (c->scans[a][b].scan) cudaFree(c->scans[a][b].scan); if (c->scans[a][b].iscan) cudaFree(c->scans[a][b].iscan); }
    cudaStreamDestroy(c->stream);
    delete c;
}

const char *svtb200_last_error(const SvtB200Ctx *c) { return c ? c->err.c_str() : "null context"; }
uint64_t svtb200_launch_count(const SvtB200Ctx *c) { return c ? c->launches : 0; }
void *svtb200_stream(SvtB200Ctx *c) { return c ? (void *)c->stream : nullptr; }

int svtb200_sync(SvtB200Ctx *c) {
    if (!c) return SVTB200_EINVAL;
    SVTB200_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    return SVTB200_OK;
}

void *svtb200_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void svtb200_host_free(void *p) { if (p) cudaFreeHost(p); }

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
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, c->stream));
    SVTB200_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    return SVTB200_OK;
}
int svtb200_memcpy_d2h(SvtB200Ctx *c, void *h, const void *d, size_t n) {
    if (!c) return SVTB200_EINVAL;
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
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
        cudaMemsetAsync(dst->buf + bytes, 0, 256, c->stream);
    }
    dst->stride = src->stride; dst->ox = src->origin_x; dst->oy = src->origin_y; dst->w = src->width; dst->h = src->height;
    dst->rows = src->padded_height;
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(dst->buf, src->buf, bytes, cudaMemcpyHostToDevice, c->stream));
    return 0;
}

int svtb200_picture_upload(SvtB200Ctx *c, uint64_t pic_id, const SvtB200Plane *full, const SvtB200Plane *quarter,
                           const SvtB200Plane *sixteenth) {
    if (!c || !full || !full->buf) return SVTB200_EINVAL;
    std::lock_guard<std::mutex> g(c->mu);
    cudaSetDevice(c->device);
    DevPicture &p = c->pics[pic_id];
    int rc;
    if ((rc = upload_plane(c, full, &p.pl[0]))) return rc;
    if ((rc = upload_plane(c, quarter, &p.pl[1]))) return rc;
    if ((rc = upload_plane(c, sixteenth, &p.pl[2]))) return rc;
    return SVTB200_OK;
}

int svtb200_picture_plane_dev(SvtB200Ctx *c, uint64_t pic_id, int which, SvtB200Plane *out) {
    if (!c || !out || which < 0 || which > 2) return SVTB200_EINVAL;
    std::lock_guard<std::mutex> g(c->mu);
    auto it = c->pics.find(pic_id);
    if (it == c->pics.end() || !it->second.pl[which].buf) return SVTB200_ENOENT;
    const DevPlane &d = it->second.pl[which];
    out->buf = d.buf; out->stride = d.stride; out->origin_x = d.ox; out->origin_y = d.oy; out->width = d.w; out->height = d.h;
    out->padded_height = d.rows;
    return SVTB200_OK;
}

int svtb200_picture_release(SvtB200Ctx *c, uint64_t pic_id) {
    if (!c) return SVTB200_EINVAL;
    std::lock_guard<std::mutex> g(c->mu);
    auto it = c->pics.find(pic_id);
    if (it == c->pics.end()) return SVTB200_ENOENT;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (int i = 0; i < 3; i++) if (it->second.pl[i].buf) cudaFree(it->second.pl[i].buf);
    c->pics.erase(it);
    return SVTB200_OK;
}

} /* extern "C" */
