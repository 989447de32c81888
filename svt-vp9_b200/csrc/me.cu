/* me.cu -- host side of S1: svtb200_me_picture[_dev] (see include/svtb200.h). */
#include <cuda.h>
#include <algorithm>
#include <stdlib.h>
#include "common.cuh"
#include "me_kernel.cuh"

/* cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda) */
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
        else cudaGetLastError();
    }
    return fn;
}
static_assert(sizeof(CUtensorMap) == sizeof(MeTensorMap), "tensor map size");

/* 2-D byte tensor over a whole device plane (pitch x rows) with a box_w x box_h box; cached in the plane.  Fails (-> software
 * staging) when the plane is not TMA-legal: base and pitch multiples of 16, box sides <= 256, box width a multiple of 16. */
static bool plane_tensor_map(DevPlane &pl, int box_w, int box_h, MeTensorMap *out) {
    if (!pl.buf || ((uintptr_t)pl.buf & 15) || (pl.stride & 15) || box_w <= 0 || box_h <= 0 || box_w > 256 || box_h > 256 || (box_w & 15)) return false;
    for (int i = 0; i < 2; i++) {
        DevTensorMap &t = pl.tm[i];
        if (t.buf == pl.buf && t.stride == pl.stride && t.rows == pl.rows && t.box_w == box_w && t.box_h == box_h) { memcpy(out, t.opaque, sizeof(*out)); return true; }
    }
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return false;
    CUtensorMap m;
    const cuuint64_t dims[2] = {(cuuint64_t)pl.stride, (cuuint64_t)pl.rows}, strides[1] = {(cuuint64_t)pl.stride};
    const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h}, estr[2] = {1, 1};
    if (enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, pl.buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return false;
    DevTensorMap &t = (pl.tm[0].buf == pl.buf && pl.tm[0].stride == pl.stride && pl.tm[0].rows == pl.rows) ? pl.tm[1] : pl.tm[0]; /* keep the other shape of the same buffer */
    memcpy(t.opaque, &m, sizeof(m));
    t.buf = pl.buf; t.stride = pl.stride; t.rows = pl.rows; t.box_w = box_w; t.box_h = box_h;
    memcpy(out, &m, sizeof(*out));
    return true;
}

static void to_dev(const DevPlane &d, MePlaneDev *o) { o->buf = d.buf; o->stride = d.stride; o->ox = d.ox; o->oy = d.oy; o->w = d.w; o->h = d.h; }

static int me_launch(SvtB200Ctx *c, const SvtB200MeParams *P, uint64_t cur_id, const uint64_t ref_id[2],
                     SvtB200MeCuResult *out_dev, uint32_t *rcme_dev) {
    if (!c || !P || !ref_id || !out_dev) return SVTB200_EINVAL;
    if (P->num_lists < 1 || P->num_lists > 2 || P->luma_width <= 0 || P->luma_height <= 0) return SVTB200_EINVAL;
    if (P->number_hme_search_region_in_width > 2 || P->number_hme_search_region_in_height > 2) return SVTB200_EINVAL;
    MeKernelArgs A;
    memset(&A, 0, sizeof(A));
    A.P = *P;
    auto itc = c->pics.find(cur_id);
    if (itc == c->pics.end()) { c->err = "unknown current picture id"; return SVTB200_ENOENT; }
    for (int i = 0; i < 3; i++) to_dev(itc->second.pl[i], &A.cur[i]);
    for (int l = 0; l < P->num_lists; l++) {
        auto itr = c->pics.find(ref_id[l]);
        if (itr == c->pics.end()) { c->err = "unknown reference picture id"; return SVTB200_ENOENT; }
        for (int i = 0; i < 3; i++) to_dev(itr->second.pl[i], &A.ref[l][i]);
    }
    if (P->num_lists == 1) for (int i = 0; i < 3; i++) A.ref[1][i] = A.ref[0][i];
    if (P->enable_hme_flag && ((P->enable_hme_level0 && !A.cur[2].buf) || (P->enable_hme_level1 && !A.cur[1].buf))) {
        c->err = "HME level enabled but its decimated plane was not uploaded"; return SVTB200_EINVAL;
    }
    A.sbs_w = (P->luma_width + ME_SB - 1) / ME_SB;
    A.n_sb  = A.sbs_w * ((P->luma_height + ME_SB - 1) / ME_SB);
    A.out = out_dev; A.rcme = rcme_dev;
    A.n_work = A.n_sb;
    A.win_bytes = me_window_bytes(*P);
    A.interp_plane_bytes = me_interp_plane_bytes(*P);
    A.interp_in_smem = 1;
    /* ---- TMA staging of the source super-block where the plane allows it (svtb200_picture_upload_luma's device geometry does) */
    static const bool no_tma = getenv("SVTB200_ME_NO_TMA") != nullptr;
    if (!no_tma) {
        DevPicture &pc = itc->second;
        if ((pc.pl[0].ox & 15) == 0 && plane_tensor_map(pc.pl[0], ME_SB, ME_SB, &A.tm_cur)) A.tma |= ME_TMA_SRC; /* box start = ox + 64 k: 16-byte aligned */
    }
    MeSmemLayout L = me_smem_layout(A.win_bytes, A.interp_plane_bytes, 1, me_bip0_bytes(*P));
    if (L.total > c->smem_optin) {
        A.interp_in_smem = 0;
        L = me_smem_layout(A.win_bytes, A.interp_plane_bytes, 0, me_bip0_bytes(*P));
        if (L.total > c->smem_optin) { c->err = "search window does not fit in shared memory"; return SVTB200_EINVAL; }
        const size_t need = (size_t)A.n_sb * 2 * 3 * me_align16(A.interp_plane_bytes);
        int rc = svtb200_grow(c, (void **)&c->interp_scratch, &c->interp_scratch_bytes, need);
        if (rc) return rc;
        A.interp_scratch = c->interp_scratch;
    }
    if (c->me_trace_enabled) {
        int rc = svtb200_grow(c, &c->me_trace, &c->me_trace_bytes, (size_t)A.n_sb * sizeof(MeTraceDev));
        if (rc) return rc;
        A.trace = (MeTraceDev *)c->me_trace;
    }
    SVTB200_CUDA_CHECK(c, cudaFuncSetAttribute(me_sb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    /* persistent grid: as many CTAs as stay resident (4 per SM by registers, fewer if the windows are large), pulling super-blocks
     * from a counter; 64 counters in a ring so that launches in flight on different queues never share one */
    int grid = A.n_sb;
    static const bool no_persist = getenv("SVTB200_ME_NO_PERSISTENT") != nullptr;
    if (!no_persist) {
        if (!c->me_counter && cudaMalloc((void **)&c->me_counter, 64 * sizeof(uint32_t)) != cudaSuccess) { cudaGetLastError(); c->me_counter = nullptr; }
        if (c->me_counter) {
            int per_sm = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, me_sb_kernel, ME_THREADS, (size_t)L.total) != cudaSuccess || per_sm < 1) { cudaGetLastError(); per_sm = 1; }
            /* SVTB200_ME_CTAS_PER_SM=n: leave room on every SM (registers, shared memory, warp slots) for the kernels of the OTHER compute queues --
             * a persistent grid that fills the SMs keeps everything else out until it ends */
            static const int cap_per_sm = getenv("SVTB200_ME_CTAS_PER_SM") ? atoi(getenv("SVTB200_ME_CTAS_PER_SM")) : 0;
            if (cap_per_sm > 0 && cap_per_sm < per_sm) per_sm = cap_per_sm;
            A.sb_counter = c->me_counter + (c->me_counter_next++ & 63);
            SVTB200_CUDA_CHECK(c, cudaMemsetAsync(A.sb_counter, 0, sizeof(uint32_t), c->stream));
            grid = std::min(A.n_sb, per_sm * (c->sm_count > 0 ? c->sm_count : 148));
        }
    }
    svtb200_pic_acquire(c, itc->second);
    for (int l = 0; l < P->num_lists; l++) svtb200_pic_acquire(c, c->pics.find(ref_id[l])->second);
    me_sb_kernel<<<grid, ME_THREADS, L.total, c->stream>>>(A);
    svtb200_pic_release(c, itc->second);
    for (int l = 0; l < P->num_lists; l++) svtb200_pic_release(c, c->pics.find(ref_id[l])->second);
    c->launches++;
    SVTB200_CUDA_CHECK(c, cudaGetLastError());
    return SVTB200_OK;
}

extern "C" {

int svtb200_me_picture_dev(SvtB200Ctx *c, const SvtB200MeParams *P, uint64_t cur_id, const uint64_t ref_id[2],
                           SvtB200MeCuResult *out_dev, uint32_t *rcme_dev) {
    if (!c) return SVTB200_EINVAL;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    cudaSetDevice(c->device);
    return me_launch(c, P, cur_id, ref_id, out_dev, rcme_dev);
}

int svtb200_me_picture(SvtB200Ctx *c, const SvtB200MeParams *P, uint64_t cur_id, const uint64_t ref_id[2],
                       SvtB200MeCuResult *out, uint32_t *out_rcme) {
    if (!c || !P || !out) return SVTB200_EINVAL;
    std::lock_guard<std::recursive_mutex> g(c->mu);
    cudaSetDevice(c->device);
    const int n_sb = ((P->luma_width + ME_SB - 1) / ME_SB) * ((P->luma_height + ME_SB - 1) / ME_SB);
    const size_t bytes = (size_t)n_sb * 85 * sizeof(SvtB200MeCuResult);
    int rc = svtb200_grow(c, &c->me_out, &c->me_out_bytes, bytes);
    if (rc) return rc;
    rc = svtb200_grow(c, (void **)&c->me_rcme, &c->me_rcme_bytes, (size_t)n_sb * 4);
    if (rc) return rc;
    rc = me_launch(c, P, cur_id, ref_id, (SvtB200MeCuResult *)c->me_out, c->me_rcme);
    if (rc) return rc;
    SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(out, c->me_out, bytes, cudaMemcpyDeviceToHost, c->stream));
    if (out_rcme) SVTB200_CUDA_CHECK(c, cudaMemcpyAsync(out_rcme, c->me_rcme, (size_t)n_sb * 4, cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    return SVTB200_OK;
}

/* debugging aid for the parity tests: per-SB stage outputs of the next me_picture calls */
int svtb200_me_enable_trace(SvtB200Ctx *c, int on) { if (!c) return SVTB200_EINVAL; c->me_trace_enabled = on; return SVTB200_OK; }
int svtb200_me_read_trace(SvtB200Ctx *c, void *host, size_t bytes) {
    if (!c || !c->me_trace || bytes > c->me_trace_bytes) return SVTB200_EINVAL;
    return svtb200_memcpy_d2h(c, host, c->me_trace, bytes);
}

} /* extern "C" */
