/*
 * svtb200_wrap.c -- the reference-side binding of libsvtvp9_b200.so, as INTEGRATION.md describes it: link-time wrappers
 * around the two non-static seam functions of the UNMODIFIED reference objects
 *
 *     gcc -shared -o libSvtVp9Enc.so <reference objects> svtb200_wrap.o -lsvtvp9_b200 \
 *         -Wl,--wrap=motion_estimate_sb -Wl,--wrap=eb_vp9_loop_filter_frame
 *
 *   S1  motion_estimate_sb()        Source/Lib/Codec/EbMotionEstimation.c:4524 (call site EbMotionEstimationProcess.c:1037)
 *   S2  eb_vp9_loop_filter_frame()  Source/Lib/VPX/vp9_loopfilter.c:1521       (call site EbEncDecProcess.c:5678)
 *
 * Compiled only against the reference's headers and include/svtb200.h; no reference code is copied.  The reference calls
 * motion_estimate_sb() once per super-block from up to 20 ME threads; the GPU works per picture: the first thread that
 * reaches a picture uploads what is not resident yet, runs svtb200_me_picture() for all super-blocks and publishes the
 * results, every call then copies its own super-block's 85 entries into pcs->me_results.  If no GPU context can be created
 * the wrappers fall through to the reference functions (__real_*), so the library still works on a CPU-only box.
 * Environment: SVT_VP9_GPU=<device index> (default 0), SVT_VP9_GPU=off disables the GPU path.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "svtb200_ref_glue.h"

EbErrorType __real_motion_estimate_sb(PictureParentControlSet *pcs, uint32_t sb_index, uint32_t sb_origin_x, uint32_t sb_origin_y,
                                      MeContext *me, EbPictureBufferDesc *input);
void __real_eb_vp9_loop_filter_frame(VP9_COMMON *cm, MACROBLOCKD *xd, int frame_filter_level, int y_only, int partial_frame);

static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t  g_cv = PTHREAD_COND_INITIALIZER;
static SvtB200Ctx     *g_ctx;    /* picture residency + motion estimation */
static SvtB200Ctx     *g_ctx_lf; /* deblocking: its own queues and lock, so that a picture being filtered never waits behind (or blocks) the
                                    motion estimation of the pictures after it -- they are different pipeline stages in the reference too */
static int             g_state; /* 0 untried, 1 ready, -1 unavailable */
static unsigned long   g_me_pictures, g_me_sbs, g_lf_pictures, g_fallbacks;
static double          g_t_resident, g_t_me, g_t_lf; /* seconds spent in uploads / svtb200_me_picture / svtb200_deblock_picture */
static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

static void report(void) {
    const char *p = getenv("SVT_VP9_GPU_REPORT");
    if (!p) return;
    FILE *f = fopen(p, "w");
    if (!f) return;
    fprintf(f, "state %d\nme_pictures %lu\nme_sbs %lu\nlf_pictures %lu\nfallbacks %lu\nlaunches %llu\n", g_state, g_me_pictures, g_me_sbs,
            g_lf_pictures, g_fallbacks, g_ctx ? (unsigned long long)(svtb200_launch_count(g_ctx) + svtb200_launch_count(g_ctx_lf)) : 0ull);
    fprintf(f, "ms_upload %d\nms_me %d\nms_lf %d\n", (int)(1e3 * g_t_resident), (int)(1e3 * g_t_me), (int)(1e3 * g_t_lf));
    fclose(f);
}

/* called with g_mu held */
static int ctx_ready(void) {
    if (g_state == 0) {
        const char *e = getenv("SVT_VP9_GPU");
        g_state = -1;
        if (!(e && !strcmp(e, "off")) && svtb200_create(e ? atoi(e) : 0, &g_ctx) == SVTB200_OK) {
            if (svtb200_create(e ? atoi(e) : 0, &g_ctx_lf) == SVTB200_OK) g_state = 1;
            else { svtb200_destroy(g_ctx); g_ctx = NULL; }
        }
        if (g_state < 0 && getenv("SVT_VP9_GPU_VERBOSE")) fprintf(stderr, "svtb200: no GPU context, using the CPU kernels\n");
    }
    return g_state == 1;
}

/* The CUDA context (~0.5 s) is created while the encoder allocates its own pools, not inside the first picture. */
static void *warm_up(void *arg) {
    (void)arg;
    pthread_mutex_lock(&g_mu);
    ctx_ready();
    pthread_mutex_unlock(&g_mu);
    return NULL;
}
__attribute__((constructor)) static void svtb200_wrap_init(void) {
    pthread_t th;
    setenv("SVTB200_BLOCKING_SYNC", "1", 0); /* the encoder keeps every core busy: do not spin while the GPU works */
    if (pthread_create(&th, NULL, warm_up, NULL) == 0) pthread_detach(th);
}

/* ------------------------------------------------------------------------------------------------------------ S1 */
#define MAX_IN_FLIGHT 64
typedef struct {
    PictureParentControlSet *pcs;
    uint64_t                 picture_number;
    int                      state; /* 0 free, 1 running on the GPU, 2 results ready, 3 failed */
    int                      n_sb, served;
    SvtB200MeParams          P;
    SvtB200MeCuResult       *res;   /* pinned, n_sb * 85 */
    uint32_t                *rcme;  /* pinned, n_sb */
    size_t                   cap_sb;
} MeSlot;
static MeSlot g_slots[MAX_IN_FLIGHT];

/* device residency of the picture-analysis reference objects: id = object address, tag = display number of its contents */
#define MAX_RESIDENT 256
static struct { const void *obj; uint64_t poc; int valid; } g_res[MAX_RESIDENT];

static int ensure_resident(const EbPaReferenceObject *o, uint64_t poc, int want_quarter) {
    int free_i = -1;
    for (int i = 0; i < MAX_RESIDENT; i++) {
        if (g_res[i].valid && g_res[i].obj == (const void *)o) {
            if (g_res[i].poc == poc) return SVTB200_OK;
            free_i = i; break; /* same buffer, new picture: upload over it */
        }
        if (!g_res[i].valid && free_i < 0) free_i = i;
    }
    if (free_i < 0) return SVTB200_ENOMEM;
    SvtB200Plane pl[3];
    svtb200_glue_pa_planes(o, pl);
    static int pagelock = -1; /* SVT_VP9_GPU_PAGELOCK=0: leave the reference's buffers pageable (staged copies, no page-locking cost) */
    if (pagelock < 0) { const char *e = getenv("SVT_VP9_GPU_PAGELOCK"); pagelock = !(e && e[0] == '0'); }
    if (pagelock && (!g_res[free_i].valid || g_res[free_i].obj != (const void *)o)) /* a buffer of the reference's pool, first time seen: page-lock it */
        for (int k = 0; k < 3; k++)
            if (pl[k].buf && pl[k].stride > 0 && pl[k].padded_height > 0) svtb200_host_register((void *)pl[k].buf, (size_t)pl[k].stride * pl[k].padded_height);
    int rc = svtb200_picture_upload(g_ctx, (uint64_t)(uintptr_t)o, &pl[0], want_quarter ? &pl[1] : NULL, &pl[2]);
    if (rc != SVTB200_OK) return rc;
    g_res[free_i].obj = o; g_res[free_i].poc = poc; g_res[free_i].valid = 1;
    return SVTB200_OK;
}

static void abi_to_ref(const SvtB200MeCuResult *r, MeCuResults *o) {
    o->x_mv_l0 = r->x_mv_l0; o->y_mv_l0 = r->y_mv_l0; o->x_mv_l1 = r->x_mv_l1; o->y_mv_l1 = r->y_mv_l1;
    o->total_me_candidate_index = r->total_me_candidate_index;
    for (int i = 0; i < 3; i++) { o->distortion_direction[i].distortion = r->cand[i].distortion; o->distortion_direction[i].direction = r->cand[i].direction; }
}

/* runs the whole picture on the GPU; called WITHOUT g_mu (the slot is marked "running") */
static int run_picture(MeSlot *s, MeContext *me) {
    PictureParentControlSet *pcs = s->pcs;
    const int want_q = s->P.enable_hme_flag && s->P.enable_hme_level1;
    EbPaReferenceObject *cur = (EbPaReferenceObject *)pcs->pareference_picture_wrapper_ptr->object_ptr;
    EbPaReferenceObject *r0 = (EbPaReferenceObject *)pcs->ref_pa_pic_ptr_array[0]->object_ptr;
    EbPaReferenceObject *r1 = s->P.num_lists == 2 ? (EbPaReferenceObject *)pcs->ref_pa_pic_ptr_array[1]->object_ptr : r0;
    int rc;
    pthread_mutex_lock(&g_mu); /* residency table + uploads are serialised; the kernels of different pictures still overlap the uploads */
    const double t0 = now_s();
    rc = ensure_resident(cur, pcs->picture_number, want_q);
    if (rc == SVTB200_OK) rc = ensure_resident(r0, pcs->ref_pic_poc_array[0], want_q);
    if (rc == SVTB200_OK && s->P.num_lists == 2) rc = ensure_resident(r1, pcs->ref_pic_poc_array[1], want_q);
    g_t_resident += now_s() - t0;
    pthread_mutex_unlock(&g_mu);
    if (rc != SVTB200_OK) return rc;
    const uint64_t ids[2] = {(uint64_t)(uintptr_t)r0, (uint64_t)(uintptr_t)r1};
    (void)me;
    const double t1 = now_s();
    rc = svtb200_me_picture(g_ctx, &s->P, (uint64_t)(uintptr_t)cur, ids, s->res, s->P.rate_control_mode ? s->rcme : NULL);
    pthread_mutex_lock(&g_mu);
    g_t_me += now_s() - t1;
    pthread_mutex_unlock(&g_mu);
    return rc;
}

/* One call per super-block from many threads.  Only the first call for a picture (which runs it on the GPU) and the calls that
 * arrive while it runs take the lock; afterwards each thread remembers the picture's slot and copies its super-block's results
 * out of the pinned buffer without any lock.  A slot is recycled when its last super-block has been handed out (atomic count),
 * which cannot happen while a call that has not yet been counted is still reading it. */
EbErrorType __wrap_motion_estimate_sb(PictureParentControlSet *pcs, uint32_t sb_index, uint32_t sb_origin_x, uint32_t sb_origin_y,
                                      MeContext *me, EbPictureBufferDesc *input) {
    static __thread MeSlot *tl_slot;
    MeSlot *s = tl_slot;
    if (!(s && __atomic_load_n(&s->state, __ATOMIC_ACQUIRE) >= 2 && s->pcs == pcs && s->picture_number == pcs->picture_number)) {
        pthread_mutex_lock(&g_mu);
        if (!ctx_ready()) {
            pthread_mutex_unlock(&g_mu);
            return __real_motion_estimate_sb(pcs, sb_index, sb_origin_x, sb_origin_y, me, input);
        }
        MeSlot *fr = NULL;
        s = NULL;
        for (int i = 0; i < MAX_IN_FLIGHT; i++) {
            if (g_slots[i].state && g_slots[i].pcs == pcs && g_slots[i].picture_number == pcs->picture_number) { s = &g_slots[i]; break; }
            if (!g_slots[i].state && !fr) fr = &g_slots[i];
        }
        if (!s) {
            if (!fr) { g_fallbacks++; pthread_mutex_unlock(&g_mu); return __real_motion_estimate_sb(pcs, sb_index, sb_origin_x, sb_origin_y, me, input); }
            s = fr;
            s->pcs = pcs; s->picture_number = pcs->picture_number; s->served = 0;
            __atomic_store_n(&s->state, 1, __ATOMIC_RELEASE);
            svtb200_glue_me_params(pcs, me, &s->P);
            s->n_sb = ((s->P.luma_width + 63) / 64) * ((s->P.luma_height + 63) / 64);
            if (s->cap_sb < (size_t)s->n_sb) {
                if (s->res) svtb200_host_free(s->res);
                if (s->rcme) svtb200_host_free(s->rcme);
                s->res = (SvtB200MeCuResult *)svtb200_host_alloc((size_t)s->n_sb * 85 * sizeof(SvtB200MeCuResult));
                s->rcme = (uint32_t *)svtb200_host_alloc((size_t)s->n_sb * sizeof(uint32_t));
                s->cap_sb = (s->res && s->rcme) ? (size_t)s->n_sb : 0;
            }
            g_me_pictures++;
            pthread_mutex_unlock(&g_mu);
            const int rc = s->cap_sb ? run_picture(s, me) : SVTB200_ENOMEM;
            pthread_mutex_lock(&g_mu);
            __atomic_store_n(&s->state, rc == SVTB200_OK ? 2 : 3, __ATOMIC_RELEASE);
            if (rc != SVTB200_OK && getenv("SVT_VP9_GPU_VERBOSE")) fprintf(stderr, "svtb200: ME on the GPU failed (%d: %s)\n", rc, svtb200_last_error(g_ctx));
            pthread_cond_broadcast(&g_cv);
        } else {
            while (s->state == 1) pthread_cond_wait(&g_cv, &g_mu);
        }
        pthread_mutex_unlock(&g_mu);
        tl_slot = s;
    }
    const int ok = __atomic_load_n(&s->state, __ATOMIC_ACQUIRE) == 2;
    const int rc_mode = s->P.rate_control_mode;
    if (ok) {
        const SvtB200MeCuResult *mine = s->res + (size_t)sb_index * 85;
        for (int i = 0; i < 85; i++) abi_to_ref(&mine[i], &pcs->me_results[sb_index][i]);
        if (rc_mode) pcs->rcme_distortion[sb_index] = s->rcme[sb_index];
    }
    __atomic_add_fetch(&g_me_sbs, 1, __ATOMIC_RELAXED);
    if (!ok) __atomic_add_fetch(&g_fallbacks, 1, __ATOMIC_RELAXED);
    if (__atomic_add_fetch(&s->served, 1, __ATOMIC_ACQ_REL) == s->n_sb) { /* every super-block of the picture has been handed out */
        tl_slot = NULL;
        pthread_mutex_lock(&g_mu);
        __atomic_store_n(&s->state, 0, __ATOMIC_RELEASE);
        pthread_mutex_unlock(&g_mu);
    }
    if (!ok) return __real_motion_estimate_sb(pcs, sb_index, sb_origin_x, sb_origin_y, me, input);
    return EB_ErrorNone;
}

/* ------------------------------------------------------------------------------------------------------------ S2 */
void __wrap_eb_vp9_loop_filter_frame(VP9_COMMON *cm, MACROBLOCKD *xd, int frame_filter_level, int y_only, int partial_frame) {
    pthread_mutex_lock(&g_mu);
    const int ready = ctx_ready();
    pthread_mutex_unlock(&g_mu);
    if (!ready || partial_frame || !frame_filter_level) {
        __real_eb_vp9_loop_filter_frame(cm, xd, frame_filter_level, y_only, partial_frame);
        return;
    }
    SvtB200LfParams P;
    SvtB200LfThresh thr[64];
    svtb200_glue_lf(cm, frame_filter_level, y_only, &P, thr);
    {   /* the reconstruction buffers come from a small pool: page-lock each plane the first time it is filtered */
        static const void *seen[64];
        static int n_seen;
        pthread_mutex_lock(&g_mu);
        int known = 0;
        for (int i = 0; i < n_seen; i++) known |= seen[i] == (const void *)xd->plane[0].dst.buf;
        if (!known && n_seen < 64) {
            seen[n_seen++] = xd->plane[0].dst.buf;
            svtb200_host_register(xd->plane[0].dst.buf, (size_t)xd->plane[0].dst.stride * (P.height - 1) + P.width);
            if (!y_only) {
                svtb200_host_register(xd->plane[1].dst.buf, (size_t)xd->plane[1].dst.stride * (P.height / 2 - 1) + P.width / 2);
                svtb200_host_register(xd->plane[2].dst.buf, (size_t)xd->plane[2].dst.stride * (P.height / 2 - 1) + P.width / 2);
            }
        }
        pthread_mutex_unlock(&g_mu);
    }
    const double t2 = now_s();
    const int rc = svtb200_deblock_picture(g_ctx_lf, &P, xd->plane[0].dst.buf, xd->plane[0].dst.stride, xd->plane[1].dst.buf, xd->plane[2].dst.buf,
                                           xd->plane[1].dst.stride, (const SvtB200LfMask *)cm->lf.lfm, cm->lf.lfm_stride, thr);
    pthread_mutex_lock(&g_mu);
    g_t_lf += now_s() - t2;
    if (rc == SVTB200_OK) g_lf_pictures++; else g_fallbacks++;
    pthread_mutex_unlock(&g_mu);
    if (rc != SVTB200_OK) __real_eb_vp9_loop_filter_frame(cm, xd, frame_filter_level, y_only, partial_frame);
}

__attribute__((destructor)) static void svtb200_wrap_fini(void) { report(); }
