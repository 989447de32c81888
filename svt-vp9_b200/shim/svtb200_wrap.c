/*
 * svtb200_wrap.c -- the reference-side binding of libsvtvp9_b200.so, as INTEGRATION.md describes it: link-time wrappers
 * around the two non-static seam functions of the UNMODIFIED reference objects
 *
 *     gcc -shared -o libSvtVp9Enc.so.1 <reference objects> svtb200_wrap.o -lsvtvp9_b200 \
 *         -Wl,--wrap=motion_estimate_sb -Wl,--wrap=eb_vp9_loop_filter_frame
 *
 *   S1  motion_estimate_sb()        Source/Lib/Codec/EbMotionEstimation.c:4524 (call site EbMotionEstimationProcess.c:1037)
 *   S2  eb_vp9_loop_filter_frame()  Source/Lib/VPX/vp9_loopfilter.c:1521       (call site EbEncDecProcess.c:5678)
 *   prefetch: eb_vp9_post_full_object() (Source/Lib/Codec/EbSystemResourceManager.c:482) is wrapped only to SEE the picture-analysis results go
 *       by (EbPictureAnalysisProcess.c, end of the kernel loop): that is the moment a picture's padded / decimated planes are final, a whole
 *       mini-GOP of look-ahead before motion estimation asks for them, so the upload is queued there instead of inside the first
 *       motion_estimate_sb() call of the picture.  The results objects are recognised by their system resource, learnt from the wrapped
 *       eb_vp9_picture_analysis_context_ctor() (EbPictureAnalysisProcess.c:46).  Off by default: SVT_VP9_GPU_PREFETCH=lock page-locks the picture's buffers there (background thread), =1 also uploads there; on
 *       the 128-thread box neither beat the on-demand upload inside the picture's first ME call (the short clips of the bench are
 *       dominated by first-touch costs: cudaMalloc + page-locking per pool object).
 *   teardown: eb_vp9_deinit_encoder() (Source/Lib/Codec/EbEncHandle.c:1711) is exported from here; the reference's own
 *       definition is kept under the name svtb200_real_eb_vp9_deinit_encoder (objcopy --redefine-sym on a COPY of the
 *       compiled object, oracle/build_ref.sh) and called after the GPU state of that encoder instance has been dropped.
 *
 * Compiled only against the reference's headers and include/svtb200.h; no reference code is copied.  The reference calls
 * motion_estimate_sb() once per super-block from up to 20 ME threads; the GPU works per picture: the first thread that
 * reaches a picture uploads what is not resident yet (per-picture locks: uploads of different pictures do not wait for each
 * other), queues svtb200_me_picture_dev() and the download of the results on the context's queues WITHOUT blocking inside the
 * library, then waits for that picture's marker; calls for other pictures queue their own work behind it meanwhile.  Every call
 * then copies its own super-block's 85 entries into pcs->me_results.  If no GPU context can be created the wrappers fall
 * through to the reference functions (__real_*), so the library still works on a CPU-only box.
 *
 * Environment: SVT_VP9_GPU=<device index> (default 0) or a comma-separated list of devices ("0,1,2,3": ONE encoder, its pictures
 * dealt to the devices mini-GOP by mini-GOP -- motion estimation is open loop, a picture needs only its own and its references' SOURCE
 * planes, which are uploaded to the device that runs it; filtered pictures are dealt round-robin; no exchange between devices),
 * SVT_VP9_GPU=off disables the GPU path; SVT_VP9_GPU_REPORT=<file> counters at exit.  Measurement aids (tools / bench.py): SVT_VP9_ME_RECORD=<file> writes every picture's ME results to a table at exit;
 * SVT_VP9_ME_REPLAY=<file> serves motion estimation from such a table (no ME work at all: the speed the pipeline around the
 * seam would reach with a zero-cost motion estimation).
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "svtb200_ref_glue.h"
#include "EbEncHandle.h"
#include "EbPictureAnalysisProcess.h"
#include "EbPictureAnalysisResults.h"
#include "EbSystemResourceManager.h"

EbErrorType __real_motion_estimate_sb(PictureParentControlSet *pcs, uint32_t sb_index, uint32_t sb_origin_x, uint32_t sb_origin_y,
                                      MeContext *me, EbPictureBufferDesc *input);
void __real_eb_vp9_loop_filter_frame(VP9_COMMON *cm, MACROBLOCKD *xd, int frame_filter_level, int y_only, int partial_frame);

static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;    /* slot table, residency table, counters */
static pthread_mutex_t g_lf_mu = PTHREAD_MUTEX_INITIALIZER; /* deblocking: page-lock table */
static pthread_cond_t  g_cv = PTHREAD_COND_INITIALIZER;
#define MAX_DEV 8
static int             g_ndev;              /* devices in use (SVT_VP9_GPU list) */
static SvtB200Ctx     *g_ctxs[MAX_DEV];     /* per device: picture residency + motion estimation */
static SvtB200Ctx     *g_ctxs_lf[MAX_DEV];  /* per device: deblocking on its own queues and lock, so that a picture being filtered never waits behind
                                               (or blocks) the motion estimation of the pictures after it -- different pipeline stages in the reference too */
static pthread_mutex_t g_queue_mus[MAX_DEV], g_lf_mus[MAX_DEV];
static unsigned long   g_dev_pictures[MAX_DEV];
static unsigned        g_lf_next;
#define g_ctx g_ctxs[0]
#define g_ctx_lf g_ctxs_lf[0]
static int             g_state; /* 0 untried, 1 ready, -1 unavailable */
static unsigned long   g_me_pictures, g_me_sbs, g_lf_pictures, g_fallbacks, g_replayed, g_prefetched;
static unsigned long   g_uploads;
static double          g_t_register, g_t_upcall; /* of g_t_resident: page-locking / inside the upload calls */
static double          g_t_resident, g_t_me, g_t_lf, g_t_prefetch; /* seconds spent in uploads / waiting for ME results / svtb200_deblock_picture */
static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

/* ------------------------------------------------------------------------------------------------ record / replay */
typedef struct { uint64_t picture_number; uint32_t n_sb, rc_mode; SvtB200MeCuResult *res; uint32_t *rcme; } TabEntry;
#define TAB_MAX 65536
static TabEntry *g_tab;
static size_t    g_tab_n, g_tab_cap;
static int       g_replay; /* table loaded: serve ME from it */
static const char *g_record_path;

static void table_put(uint64_t picture_number, int n_sb, int rc_mode, const SvtB200MeCuResult *res, const uint32_t *rcme) { /* g_mu held */
    if (!g_tab) { g_tab_cap = TAB_MAX; g_tab = (TabEntry *)calloc(g_tab_cap, sizeof(TabEntry)); } /* never moved: readers hold pointers into it */
    if (g_tab_n == g_tab_cap) return;
    TabEntry *e = &g_tab[g_tab_n++];
    e->picture_number = picture_number; e->n_sb = (uint32_t)n_sb; e->rc_mode = (uint32_t)rc_mode;
    e->res = (SvtB200MeCuResult *)malloc((size_t)n_sb * 85 * sizeof(SvtB200MeCuResult));
    e->rcme = (uint32_t *)malloc((size_t)n_sb * sizeof(uint32_t));
    memcpy(e->res, res, (size_t)n_sb * 85 * sizeof(SvtB200MeCuResult));
    memcpy(e->rcme, rcme, (size_t)n_sb * sizeof(uint32_t));
}
/* CPU-path recording (no GPU): one super-block of the reference's own results */
static void table_put_sb(PictureParentControlSet *pcs, uint32_t sb_index) {
    SequenceControlSet *scs = (SequenceControlSet *)pcs->sequence_control_set_wrapper_ptr->object_ptr;
    const uint32_t n_sb = (uint32_t)(((scs->luma_width + 63) / 64) * ((scs->luma_height + 63) / 64));
    pthread_mutex_lock(&g_mu);
    TabEntry *e = NULL;
    for (size_t i = g_tab_n; i-- > 0;) if (g_tab[i].picture_number == pcs->picture_number) { e = &g_tab[i]; break; }
    if (!e) {
        if (!g_tab) { g_tab_cap = TAB_MAX; g_tab = (TabEntry *)calloc(g_tab_cap, sizeof(TabEntry)); }
        if (g_tab_n == g_tab_cap) { pthread_mutex_unlock(&g_mu); return; }
        e = &g_tab[g_tab_n++];
        e->picture_number = pcs->picture_number; e->n_sb = n_sb; e->rc_mode = scs->static_config.rate_control_mode;
        e->res = (SvtB200MeCuResult *)calloc((size_t)n_sb * 85, sizeof(SvtB200MeCuResult));
        e->rcme = (uint32_t *)calloc(n_sb, sizeof(uint32_t));
    }
    pthread_mutex_unlock(&g_mu);
    for (int i = 0; i < 85; i++) {
        const MeCuResults *m = &pcs->me_results[sb_index][i];
        SvtB200MeCuResult *o = &e->res[(size_t)sb_index * 85 + i];
        o->x_mv_l0 = m->x_mv_l0; o->y_mv_l0 = m->y_mv_l0; o->x_mv_l1 = m->x_mv_l1; o->y_mv_l1 = m->y_mv_l1;
        o->total_me_candidate_index = m->total_me_candidate_index;
        for (int k = 0; k < 3; k++) { o->cand[k].distortion = m->distortion_direction[k].distortion; o->cand[k].direction = m->distortion_direction[k].direction; }
    }
    if (e->rc_mode) e->rcme[sb_index] = pcs->rcme_distortion[sb_index];
}
static void table_write(void) {
    if (!g_record_path || !g_tab_n) return;
    FILE *f = fopen(g_record_path, "wb");
    if (!f) return;
    uint64_t n = g_tab_n;
    fwrite(&n, 8, 1, f);
    for (size_t i = 0; i < g_tab_n; i++) {
        const TabEntry *e = &g_tab[i];
        fwrite(&e->picture_number, 8, 1, f); fwrite(&e->n_sb, 4, 1, f); fwrite(&e->rc_mode, 4, 1, f);
        fwrite(e->res, sizeof(SvtB200MeCuResult), (size_t)e->n_sb * 85, f);
        fwrite(e->rcme, 4, e->n_sb, f);
    }
    fclose(f);
}
static void table_read(const char *path) {
    FILE *f = fopen(path, "rb");
    uint64_t n = 0;
    if (!f) return;
    if (fread(&n, 8, 1, f) == 1 && n < (1u << 24)) {
        const uint64_t cap = 2 * n + 1024; /* indexed by picture number: pictures without ME (intra) leave holes */
        g_tab = (TabEntry *)calloc((size_t)cap, sizeof(TabEntry));
        for (uint64_t i = 0; i < n; i++) {
            TabEntry e;
            if (fread(&e.picture_number, 8, 1, f) != 1 || fread(&e.n_sb, 4, 1, f) != 1 || fread(&e.rc_mode, 4, 1, f) != 1) break;
            e.res = (SvtB200MeCuResult *)malloc((size_t)e.n_sb * 85 * sizeof(SvtB200MeCuResult));
            e.rcme = (uint32_t *)malloc((size_t)e.n_sb * sizeof(uint32_t));
            if (fread(e.res, sizeof(SvtB200MeCuResult), (size_t)e.n_sb * 85, f) != (size_t)e.n_sb * 85 || fread(e.rcme, 4, e.n_sb, f) != e.n_sb) break;
            if (e.picture_number < cap) g_tab[e.picture_number] = e; /* indexed by picture number for the lock-free lookup */
            g_tab_n++;
        }
        g_replay = g_tab_n > 0;
        g_tab_cap = (size_t)cap;
    }
    fclose(f);
}

/* svtb200_wrap_tx.c: seam S3 bound call by call (SVT_VP9_GPU_ENCODE_PASS=1) */
void svtb200_tx_install(int device);
void svtb200_tx_uninstall(void);
void svtb200_tx_report(FILE *f);

static unsigned long long launches_total(void) {
    unsigned long long n = 0;
    for (int d = 0; d < g_ndev; d++) n += svtb200_launch_count(g_ctxs[d]) + svtb200_launch_count(g_ctxs_lf[d]);
    return n;
}
static void report(void) {
    const char *p = getenv("SVT_VP9_GPU_REPORT");
    table_write();
    if (!p) return;
    FILE *f = fopen(p, "w");
    if (!f) return;
    fprintf(f, "state %d\nme_pictures %lu\nme_sbs %lu\nlf_pictures %lu\nfallbacks %lu\nreplayed_sbs %lu\nlaunches %llu\n", g_state, g_me_pictures, g_me_sbs,
            g_lf_pictures, g_fallbacks, g_replayed, launches_total());
    fprintf(f, "devices %d\n", g_ndev);
    for (int d = 0; d < g_ndev; d++) fprintf(f, "me_pictures_dev%d %lu\n", d, g_dev_pictures[d]);
    fprintf(f, "ms_upload %d\nms_me %d\nms_lf %d\nprefetched %lu\nms_prefetch %d\n", (int)(1e3 * g_t_resident), (int)(1e3 * g_t_me), (int)(1e3 * g_t_lf),
            g_prefetched, (int)(1e3 * g_t_prefetch));
    fprintf(f, "uploads %lu\nms_pagelock %d\nms_upload_calls %d\n", g_uploads, (int)(1e3 * g_t_register), (int)(1e3 * g_t_upcall));
    svtb200_tx_report(f);
    fclose(f);
}

/* called with g_mu held */
static int ctx_ready(void) {
    if (g_state == 0) {
        const char *e = getenv("SVT_VP9_GPU");
        g_state = -1;
        if (!(e && !strcmp(e, "off"))) {
            int devs[MAX_DEV], n = 0;
            for (const char *q = e ? e : "0"; *q && n < MAX_DEV;) { /* "2" or "0,1,2,3" */
                devs[n++] = atoi(q);
                while (*q && *q != ',') q++;
                if (*q == ',') q++;
            }
            int ok = n > 0;
            for (int d = 0; d < n && ok; d++) {
                ok = svtb200_create(devs[d], &g_ctxs[d]) == SVTB200_OK && svtb200_create(devs[d], &g_ctxs_lf[d]) == SVTB200_OK;
                pthread_mutex_init(&g_queue_mus[d], NULL);
                pthread_mutex_init(&g_lf_mus[d], NULL);
            }
            if (ok) { g_ndev = n; g_state = 1; }
            else for (int d = 0; d < n; d++) { if (g_ctxs[d]) svtb200_destroy(g_ctxs[d]); if (g_ctxs_lf[d]) svtb200_destroy(g_ctxs_lf[d]); g_ctxs[d] = g_ctxs_lf[d] = NULL; }
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
    const char *rp = getenv("SVT_VP9_ME_REPLAY");
    g_record_path = getenv("SVT_VP9_ME_RECORD");
    if (rp) table_read(rp);
    setenv("SVTB200_BLOCKING_SYNC", "1", 0); /* the encoder keeps every core busy: do not spin while the GPU works */
    if (!g_replay && pthread_create(&th, NULL, warm_up, NULL) == 0) pthread_detach(th);
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
    int                      dev;   /* the device (index into g_ctxs) this picture runs on */
    SvtB200MeCuResult       *d_res_dev[MAX_DEV]; /* device twins, one set per device the slot has been used with */
    uint32_t                *d_rcme_dev[MAX_DEV];
    size_t                   cap_sb, cap_dev[MAX_DEV];
} MeSlot;
static MeSlot g_slots[MAX_IN_FLIGHT];

/* device residency of the picture-analysis reference objects: id = object address, tag = display number of its contents, inst = the
 * encoder instance (its EncodeContext) the object belongs to -- eb_vp9_deinit_encoder drops that instance's entries, so that a later
 * encoder in the same process whose buffers land on the same addresses never matches a stale picture */
#define MAX_RESIDENT 256
typedef struct { const void *obj, *inst; uint64_t poc[MAX_DEV]; int valid[MAX_DEV], has_q[MAX_DEV], locked; void *reg[3]; pthread_mutex_t mu; } ResEntry;
static ResEntry g_res[MAX_RESIDENT];
static pthread_once_t g_res_once = PTHREAD_ONCE_INIT;
static void res_init(void) { for (int i = 0; i < MAX_RESIDENT; i++) pthread_mutex_init(&g_res[i].mu, NULL); }

/* dev < 0: page-lock the object's buffers only (the background pass at the end of picture analysis) */
static int ensure_resident(int dev, const EbPaReferenceObject *o, uint64_t poc, const void *inst, int want_quarter) {
    int idx = -1, fresh = 0;
    pthread_once(&g_res_once, res_init);
    pthread_mutex_lock(&g_mu);
    for (int i = 0; i < MAX_RESIDENT; i++) {
        if (g_res[i].obj == (const void *)o && g_res[i].inst == inst) { idx = i; break; }
        if (!g_res[i].obj && idx < 0) idx = i;
    }
    if (idx >= 0 && g_res[idx].obj != (const void *)o) { g_res[idx].obj = o; g_res[idx].inst = inst; memset(g_res[idx].valid, 0, sizeof(g_res[idx].valid)); fresh = 1; }
    pthread_mutex_unlock(&g_mu);
    if (idx < 0) return SVTB200_ENOMEM;
    ResEntry *e = &g_res[idx];
    pthread_mutex_lock(&e->mu); /* only callers that need THIS picture wait here */
    int rc = SVTB200_OK;
    /* how a picture reaches the device (SVT_VP9_GPU_UPLOAD): "pagelock" = the three planes as they are, each pool object page-locked the
     * first time it is seen (asynchronous DMA uploads afterwards; the first touch costs ~1 ms per MB and more on a cold host: up to 10 ms per
     * 3840x2160 picture on clips as short as the pool is deep); "pageable" = the same planes through the driver's own staging of pageable memory
     * (no first-touch cost, ~1 ms per 1080p picture more on every upload, with the context lock held meanwhile); "staged" = only the picture
     * area, copied by this thread into a ring of page-locked slots the library owns, padding and decimations made on the device
     * (svtb200_picture_upload_luma_staged).  Default "auto": page-lock objects of up to SVT_VP9_GPU_PAGELOCK_MAX_MB (6) megabytes, i.e. up to
     * 1080p, pageable above -- what measured best on 300-frame 1080p and 100-frame 4K encodes (tools/encoder_fps.py). */
    static int upmode = -1;
    static size_t lock_max_bytes;
    if (upmode < 0) {
        const char *ev = getenv("SVT_VP9_GPU_UPLOAD"), *pl0 = getenv("SVT_VP9_GPU_PAGELOCK"), *mb = getenv("SVT_VP9_GPU_PAGELOCK_MAX_MB");
        lock_max_bytes = (size_t)(mb ? atoi(mb) : 6) << 20;
        upmode = (ev && !strcmp(ev, "pagelock")) ? 1 : (ev && !strcmp(ev, "pageable")) ? 2 : (ev && !strcmp(ev, "staged")) ? 0 : (pl0 && pl0[0] == '1') ? 1 : (pl0 && pl0[0] == '0') ? 2 : 3;
    }
    if (dev < 0 ? (upmode == 1 && !e->locked) : !(e->valid[dev] && e->poc[dev] == poc && (e->has_q[dev] || !want_quarter))) {
        SvtB200Plane pl[3];
        svtb200_glue_pa_planes(o, pl);
        const double tr0 = now_s();
        const int lock_it = upmode == 1 || (upmode == 3 && (size_t)pl[0].stride * (size_t)pl[0].padded_height <= lock_max_bytes);
        if (lock_it && (fresh || !e->locked)) { /* a buffer of the reference's pool, first time seen: page-lock it */
            for (int k = 0; k < 3; k++)
                if (pl[k].buf && pl[k].stride > 0 && pl[k].padded_height > 0 &&
                    svtb200_host_register((void *)pl[k].buf, (size_t)pl[k].stride * pl[k].padded_height) == SVTB200_OK) e->reg[k] = (void *)pl[k].buf;
            e->locked = 1;
        }
        const double tr1 = now_s();
        if (dev >= 0) {
            rc = SVTB200_EINVAL;
            if (upmode == 0) rc = svtb200_picture_upload_luma_staged(g_ctxs[dev], (uint64_t)(uintptr_t)o, &pl[0], pl[1].origin_x, pl[2].origin_x, want_quarter);
            if (rc == SVTB200_EINVAL) rc = svtb200_picture_upload(g_ctxs[dev], (uint64_t)(uintptr_t)o, &pl[0], want_quarter ? &pl[1] : NULL, &pl[2]);
            if (rc == SVTB200_OK) { e->poc[dev] = poc; e->valid[dev] = 1; e->has_q[dev] = want_quarter; } else e->valid[dev] = 0;
        }
        pthread_mutex_lock(&g_mu);
        g_t_register += tr1 - tr0;
        g_t_upcall += now_s() - tr1;
        g_uploads += dev >= 0;
        pthread_mutex_unlock(&g_mu);
    }
    pthread_mutex_unlock(&e->mu);
    return rc;
}

static void abi_to_ref(const SvtB200MeCuResult *r, MeCuResults *o) {
    o->x_mv_l0 = r->x_mv_l0; o->y_mv_l0 = r->y_mv_l0; o->x_mv_l1 = r->x_mv_l1; o->y_mv_l1 = r->y_mv_l1;
    o->total_me_candidate_index = r->total_me_candidate_index;
    for (int i = 0; i < 3; i++) { o->distortion_direction[i].distortion = r->cand[i].distortion; o->distortion_direction[i].direction = r->cand[i].direction; }
}

/* runs the whole picture on the GPU; called WITHOUT g_mu (the slot is marked "running") */
static int run_picture(MeSlot *s) {
    PictureParentControlSet *pcs = s->pcs;
    SequenceControlSet *scs = (SequenceControlSet *)pcs->sequence_control_set_wrapper_ptr->object_ptr;
    const void *inst = (const void *)scs->encode_context_ptr;
    const int want_q = s->P.enable_hme_flag && s->P.enable_hme_level1;
    EbPaReferenceObject *cur = (EbPaReferenceObject *)pcs->pareference_picture_wrapper_ptr->object_ptr;
    EbPaReferenceObject *r0 = (EbPaReferenceObject *)pcs->ref_pa_pic_ptr_array[0]->object_ptr;
    EbPaReferenceObject *r1 = s->P.num_lists == 2 ? (EbPaReferenceObject *)pcs->ref_pa_pic_ptr_array[1]->object_ptr : r0;
    const size_t rb = (size_t)s->n_sb * 85 * sizeof(SvtB200MeCuResult);
    const double t0 = now_s();
    const int dev = s->dev;
    SvtB200Ctx *ctx = g_ctxs[dev];
    SvtB200MeCuResult *d_res = s->d_res_dev[dev];
    uint32_t *d_rcme = s->d_rcme_dev[dev];
    int rc = ensure_resident(dev, cur, pcs->picture_number, inst, want_q);
    if (rc == SVTB200_OK) rc = ensure_resident(dev, r0, pcs->ref_pic_poc_array[0], inst, want_q);
    if (rc == SVTB200_OK && s->P.num_lists == 2) rc = ensure_resident(dev, r1, pcs->ref_pic_poc_array[1], inst, want_q);
    const double t1 = now_s();
    uint64_t ticket = 0;
    if (rc == SVTB200_OK) {
        const uint64_t ids[2] = {(uint64_t)(uintptr_t)r0, (uint64_t)(uintptr_t)r1};
        pthread_mutex_lock(&g_queue_mus[dev]); /* kernel, fence, download, marker: queued as one unit; nothing here waits for the GPU */
        rc = svtb200_me_picture_dev(ctx, &s->P, (uint64_t)(uintptr_t)cur, ids, d_res, s->P.rate_control_mode ? d_rcme : NULL);
        if (rc == SVTB200_OK) rc = svtb200_fence(ctx, SVTB200_STREAM_COMPUTE, SVTB200_STREAM_DOWNLOAD);
        if (rc == SVTB200_OK) rc = svtb200_memcpy_d2h_async(ctx, s->res, d_res, rb);
        if (rc == SVTB200_OK && s->P.rate_control_mode) rc = svtb200_memcpy_d2h_async(ctx, s->rcme, d_rcme, (size_t)s->n_sb * 4);
        if (rc == SVTB200_OK) rc = svtb200_marker(ctx, SVTB200_STREAM_DOWNLOAD, &ticket);
        pthread_mutex_unlock(&g_queue_mus[dev]);
    }
    if (rc == SVTB200_OK) rc = svtb200_wait(ctx, ticket);
    else svtb200_sync(ctx); /* an upload of this picture may still be reading the encoder's buffer: let it finish before the CPU path takes over */
    pthread_mutex_lock(&g_mu);
    g_t_resident += t1 - t0;
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
    if (g_replay) { /* measurement aid: zero-cost motion estimation from a recorded table */
        const uint64_t pn = pcs->picture_number;
        if (pn < g_tab_cap && g_tab[pn].res && sb_index < g_tab[pn].n_sb) {
            const SvtB200MeCuResult *mine = g_tab[pn].res + (size_t)sb_index * 85;
            for (int i = 0; i < 85; i++) abi_to_ref(&mine[i], &pcs->me_results[sb_index][i]);
            if (g_tab[pn].rc_mode) pcs->rcme_distortion[sb_index] = g_tab[pn].rcme[sb_index];
            __atomic_add_fetch(&g_replayed, 1, __ATOMIC_RELAXED);
            return EB_ErrorNone;
        }
        return __real_motion_estimate_sb(pcs, sb_index, sb_origin_x, sb_origin_y, me, input);
    }
    MeSlot *s = tl_slot;
    if (!(s && __atomic_load_n(&s->state, __ATOMIC_ACQUIRE) >= 2 && s->pcs == pcs && s->picture_number == pcs->picture_number)) {
        pthread_mutex_lock(&g_mu);
        if (!ctx_ready()) {
            pthread_mutex_unlock(&g_mu);
            const EbErrorType er = __real_motion_estimate_sb(pcs, sb_index, sb_origin_x, sb_origin_y, me, input);
            if (g_record_path) table_put_sb(pcs, sb_index);
            return er;
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
            /* pictures dealt to the devices mini-GOP by mini-GOP (display order): the pictures of a mini-GOP reference each other, so a
             * device mostly finds its references already resident (dealing single pictures made every device upload ~3 pictures per
             * picture it ran) */
            s->dev = (int)(((pcs->picture_number + ((uint64_t)1 << pcs->hierarchical_levels) - 1) >> pcs->hierarchical_levels) % (uint64_t)g_ndev);
            if (s->cap_sb < (size_t)s->n_sb) {
                if (s->res) svtb200_host_free(s->res);
                if (s->rcme) svtb200_host_free(s->rcme);
                s->res = (SvtB200MeCuResult *)svtb200_host_alloc((size_t)s->n_sb * 85 * sizeof(SvtB200MeCuResult));
                s->rcme = (uint32_t *)svtb200_host_alloc((size_t)s->n_sb * sizeof(uint32_t));
                s->cap_sb = (s->res && s->rcme) ? (size_t)s->n_sb : 0;
            }
            if (s->cap_dev[s->dev] < (size_t)s->n_sb) {
                const int d = s->dev;
                if (s->d_res_dev[d]) svtb200_dev_free(g_ctxs[d], s->d_res_dev[d]);
                if (s->d_rcme_dev[d]) svtb200_dev_free(g_ctxs[d], s->d_rcme_dev[d]);
                s->d_res_dev[d] = (SvtB200MeCuResult *)svtb200_dev_alloc(g_ctxs[d], (size_t)s->n_sb * 85 * sizeof(SvtB200MeCuResult));
                s->d_rcme_dev[d] = (uint32_t *)svtb200_dev_alloc(g_ctxs[d], (size_t)s->n_sb * sizeof(uint32_t));
                s->cap_dev[d] = (s->d_res_dev[d] && s->d_rcme_dev[d]) ? (size_t)s->n_sb : 0;
            }
            g_dev_pictures[s->dev]++;
            g_me_pictures++;
            pthread_mutex_unlock(&g_mu);
            const int rc = (s->cap_sb && s->cap_dev[s->dev]) ? run_picture(s) : SVTB200_ENOMEM;
            pthread_mutex_lock(&g_mu);
            if (rc == SVTB200_OK && g_record_path) table_put(s->picture_number, s->n_sb, s->P.rate_control_mode, s->res, s->rcme);
            __atomic_store_n(&s->state, rc == SVTB200_OK ? 2 : 3, __ATOMIC_RELEASE);
            if (rc != SVTB200_OK && getenv("SVT_VP9_GPU_VERBOSE")) fprintf(stderr, "svtb200: ME on the GPU failed (%d: %s)\n", rc, svtb200_last_error(g_ctxs[s->dev]));
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

/* ------------------------------------------------------------------------------------------------------------ prefetch */
EbErrorType __real_eb_vp9_picture_analysis_context_ctor(EbPictureBufferDescInitData *init, EB_BOOL denoise_flag, PictureAnalysisContext **ctx,
                                                        EbFifo *in_fifo, EbFifo *out_fifo, uint16_t sb_total_count);
EbErrorType __real_eb_vp9_post_full_object(EbObjectWrapper *w);
#define MAX_PA_QUEUES 16
static const void *g_pa_queues[MAX_PA_QUEUES]; /* empty queues of the picture-analysis results resources of the live encoders */

EbErrorType __wrap_eb_vp9_picture_analysis_context_ctor(EbPictureBufferDescInitData *init, EB_BOOL denoise_flag, PictureAnalysisContext **ctx,
                                                        EbFifo *in_fifo, EbFifo *out_fifo, uint16_t sb_total_count) {
    if (out_fifo && out_fifo->queue_ptr) {
        pthread_mutex_lock(&g_mu);
        int known = 0, fr = -1;
        for (int i = 0; i < MAX_PA_QUEUES; i++) { known |= g_pa_queues[i] == (const void *)out_fifo->queue_ptr; if (!g_pa_queues[i] && fr < 0) fr = i; }
        if (!known && fr >= 0) g_pa_queues[fr] = (const void *)out_fifo->queue_ptr;
        pthread_mutex_unlock(&g_mu);
    }
    {   /* this constructor runs inside eb_vp9_init_encoder after the run-time dispatch tables were filled: the place to bind seam S3 */
        const char *e = getenv("SVT_VP9_GPU");
        if (!(e && !strcmp(e, "off"))) svtb200_tx_install(e ? atoi(e) : 0);
    }
    return __real_eb_vp9_picture_analysis_context_ctor(init, denoise_flag, ctx, in_fifo, out_fifo, sb_total_count);
}

/* the uploads run on a thread of their own: the picture-analysis thread only drops a note and goes on */
typedef struct { const EbPaReferenceObject *obj; uint64_t poc; const void *inst; int want_q; } PrefetchNote;
#define PREFETCH_RING 256
static PrefetchNote    g_pf_ring[PREFETCH_RING];
static unsigned        g_pf_head, g_pf_tail; /* g_mu */
static pthread_cond_t  g_pf_cv = PTHREAD_COND_INITIALIZER;
static int             g_pf_thread_started, g_pf_mode = -1;

static void *prefetch_thread(void *arg) {
    (void)arg;
    for (;;) {
        pthread_mutex_lock(&g_mu);
        while (g_pf_head == g_pf_tail) pthread_cond_wait(&g_pf_cv, &g_mu);
        const PrefetchNote n = g_pf_ring[g_pf_tail % PREFETCH_RING];
        g_pf_tail++;
        const int live = g_state == 1;
        pthread_mutex_unlock(&g_mu);
        if (!live) continue;
        const double t0 = now_s();
        /* to the device that will run this picture's motion estimation; as a reference on other devices it is uploaded on demand */
        if (ensure_resident(g_pf_mode == 2 ? (int)(n.poc % (uint64_t)g_ndev) : -1, n.obj, n.poc, n.inst, n.want_q) == SVTB200_OK) __atomic_add_fetch(&g_prefetched, 1, __ATOMIC_RELAXED);
        pthread_mutex_lock(&g_mu);
        g_t_prefetch += now_s() - t0;
        pthread_mutex_unlock(&g_mu);
    }
    return NULL;
}

EbErrorType __wrap_eb_vp9_post_full_object(EbObjectWrapper *w) {
    if (g_pf_mode < 0) { /* default "0": off; "lock": page-lock the picture's buffers at the end of picture analysis (background thread); "1": upload there */
        const char *e = getenv("SVT_VP9_GPU_PREFETCH");
        g_pf_mode = (e && e[0] == '1') ? 2 : (e && e[0] == 'l') ? 1 : 0;
    }
    if (g_pf_mode && !g_replay && w && w->system_resource_ptr && w->object_ptr) {
        const void *q = (const void *)w->system_resource_ptr->empty_queue;
        int is_pa = 0;
        for (int i = 0; i < MAX_PA_QUEUES; i++) is_pa |= g_pa_queues[i] && g_pa_queues[i] == q; /* written once at start-up, before any picture */
        if (is_pa) {
            PictureAnalysisResults *r = (PictureAnalysisResults *)w->object_ptr;
            PictureParentControlSet *pcs = r->picture_control_set_wrapper_ptr ? (PictureParentControlSet *)r->picture_control_set_wrapper_ptr->object_ptr : NULL;
            if (pcs && pcs->pareference_picture_wrapper_ptr && pcs->sequence_control_set_wrapper_ptr) {
                SequenceControlSet *scs = (SequenceControlSet *)pcs->sequence_control_set_wrapper_ptr->object_ptr;
                pthread_mutex_lock(&g_mu);
                if (ctx_ready() && g_pf_head - g_pf_tail < PREFETCH_RING) {
                    PrefetchNote *n = &g_pf_ring[g_pf_head % PREFETCH_RING];
                    n->obj = (const EbPaReferenceObject *)pcs->pareference_picture_wrapper_ptr->object_ptr; n->poc = pcs->picture_number;
                    n->inst = (const void *)scs->encode_context_ptr; n->want_q = pcs->enable_hme_flag && pcs->enable_hme_level_1_flag;
                    g_pf_head++;
                    if (!g_pf_thread_started) { pthread_t th; if (pthread_create(&th, NULL, prefetch_thread, NULL) == 0) { pthread_detach(th); g_pf_thread_started = 1; } }
                    pthread_cond_signal(&g_pf_cv);
                }
                pthread_mutex_unlock(&g_mu);
            }
        }
    }
    return __real_eb_vp9_post_full_object(w);
}

/* ------------------------------------------------------------------------------------------------------------ S2 */
static void *g_lf_reg[192];
static int   g_lf_reg_n;

void __wrap_eb_vp9_loop_filter_frame(VP9_COMMON *cm, MACROBLOCKD *xd, int frame_filter_level, int y_only, int partial_frame) {
    int ready = 0;
    if (!g_replay || g_ctx) {
        pthread_mutex_lock(&g_mu);
        ready = ctx_ready();
        pthread_mutex_unlock(&g_mu);
    }
    if (!ready || partial_frame || !frame_filter_level) {
        __real_eb_vp9_loop_filter_frame(cm, xd, frame_filter_level, y_only, partial_frame);
        return;
    }
    SvtB200LfParams P;
    SvtB200LfThresh thr[64];
    svtb200_glue_lf(cm, frame_filter_level, y_only, &P, thr);
    pthread_mutex_lock(&g_lf_mu);
    {   /* the reconstruction buffers come from a small pool: page-lock each plane the first time it is filtered */
        int known = 0;
        for (int i = 0; i < g_lf_reg_n; i++) known |= g_lf_reg[i] == (void *)xd->plane[0].dst.buf;
        /* (page-locked whatever their size: the pool is a handful of buffers, and a pageable 4K reconstruction measured 24 ms per filtered
         * picture against 2 ms) */
        if (!known && g_lf_reg_n + 3 <= 192) {
            if (svtb200_host_register(xd->plane[0].dst.buf, (size_t)xd->plane[0].dst.stride * (P.height - 1) + P.width) == SVTB200_OK)
                g_lf_reg[g_lf_reg_n++] = xd->plane[0].dst.buf;
            if (!y_only) {
                if (svtb200_host_register(xd->plane[1].dst.buf, (size_t)xd->plane[1].dst.stride * (P.height / 2 - 1) + P.width / 2) == SVTB200_OK)
                    g_lf_reg[g_lf_reg_n++] = xd->plane[1].dst.buf;
                if (svtb200_host_register(xd->plane[2].dst.buf, (size_t)xd->plane[2].dst.stride * (P.height / 2 - 1) + P.width / 2) == SVTB200_OK)
                    g_lf_reg[g_lf_reg_n++] = xd->plane[2].dst.buf;
            }
        }
    }
    const int ldev = (int)(g_lf_next++ % (unsigned)g_ndev); /* filtered pictures are dealt round-robin */
    pthread_mutex_unlock(&g_lf_mu);
    const double t2 = now_s();
    pthread_mutex_lock(&g_lf_mus[ldev]);
    const int rc = svtb200_deblock_picture(g_ctxs_lf[ldev], &P, xd->plane[0].dst.buf, xd->plane[0].dst.stride, xd->plane[1].dst.buf, xd->plane[2].dst.buf,
                                           xd->plane[1].dst.stride, (const SvtB200LfMask *)cm->lf.lfm, cm->lf.lfm_stride, thr);
    pthread_mutex_unlock(&g_lf_mus[ldev]);
    pthread_mutex_lock(&g_mu);
    g_t_lf += now_s() - t2;
    if (rc == SVTB200_OK) g_lf_pictures++; else g_fallbacks++;
    pthread_mutex_unlock(&g_mu);
    if (rc == SVTB200_EPARTIAL) {
        /* the device failed AFTER filtered samples had started to land in the reconstruction: neither filtering again nor skipping the
         * filter gives the reference's picture any more.  A CUDA error at that point is a lost device; stop instead of writing a
         * bitstream that silently differs. */
        fprintf(stderr, "svtb200: GPU failure while the filtered picture was being written back (%s); aborting\n", svtb200_last_error(g_ctxs_lf[ldev]));
        abort();
    }
    if (rc != SVTB200_OK) __real_eb_vp9_loop_filter_frame(cm, xd, frame_filter_level, y_only, partial_frame);
}

/* ------------------------------------------------------------------------------------------------------------ teardown */
EbErrorType svtb200_real_eb_vp9_deinit_encoder(EbComponentType *svt_enc_component); /* the reference's definition, renamed at link time */

/* Drops every piece of GPU state that refers to buffers of this encoder instance BEFORE the reference frees them: resident pictures,
 * page-locked ranges, result slots of its pictures.  The worker threads are idle by now (the application has drained the packets). */
#if defined(__linux__)
__attribute__((visibility("default")))
#endif
EbErrorType eb_vp9_deinit_encoder(EbComponentType *svt_enc_component) {
    const void *inst = NULL;
    if (svt_enc_component && svt_enc_component->p_component_private) {
        EbEncHandle *h = (EbEncHandle *)svt_enc_component->p_component_private;
        if (h->sequence_control_set_instance_array && h->sequence_control_set_instance_array[0])
            inst = (const void *)h->sequence_control_set_instance_array[0]->encode_context_ptr;
    }
    pthread_mutex_lock(&g_mu);
    const int live = g_state == 1;
    pthread_mutex_unlock(&g_mu);
    if (live) {
        int cs;
        pthread_setcancelstate(PTHREAD_CANCEL_DISABLE, &cs);
        pthread_mutex_lock(&g_mu);
        g_pf_tail = g_pf_head; /* notes not yet served refer to buffers that are about to be freed (one in flight holds its entry's lock, taken below) */
        pthread_mutex_unlock(&g_mu);
        for (int d = 0; d < g_ndev; d++) { svtb200_sync(g_ctxs[d]); svtb200_sync(g_ctxs_lf[d]); }
        pthread_once(&g_res_once, res_init);
        for (int i = 0; i < MAX_RESIDENT; i++) {
            ResEntry *e = &g_res[i];
            pthread_mutex_lock(&g_mu);
            const int mine = e->obj && (e->inst == inst || !inst);
            pthread_mutex_unlock(&g_mu);
            if (!mine) continue;
            pthread_mutex_lock(&e->mu);
            for (int d = 0; d < g_ndev; d++) svtb200_picture_release(g_ctxs[d], (uint64_t)(uintptr_t)e->obj);
            for (int k = 0; k < 3; k++) if (e->reg[k]) { svtb200_host_unregister(e->reg[k]); e->reg[k] = NULL; }
            pthread_mutex_lock(&g_mu);
            e->obj = NULL; e->inst = NULL; memset(e->valid, 0, sizeof(e->valid)); e->locked = 0;
            pthread_mutex_unlock(&g_mu);
            pthread_mutex_unlock(&e->mu);
        }
        pthread_mutex_lock(&g_lf_mu); /* reconstruction buffers carry no instance tag: another live encoder simply registers its planes again */
        for (int i = 0; i < g_lf_reg_n; i++) svtb200_host_unregister(g_lf_reg[i]);
        g_lf_reg_n = 0;
        pthread_mutex_unlock(&g_lf_mu);
        pthread_mutex_lock(&g_mu);
        for (int i = 0; i < MAX_IN_FLIGHT; i++) if (g_slots[i].state != 1) g_slots[i].state = 0; /* no picture of a finished encode is mid-flight */
        pthread_mutex_unlock(&g_mu);
        pthread_setcancelstate(cs, NULL);
    }
    svtb200_tx_uninstall();
    report(); /* counters / recorded table now: the reference's own teardown aborts at 3840x2160, which skips destructors */
    return svtb200_real_eb_vp9_deinit_encoder(svt_enc_component);
}

__attribute__((destructor)) static void svtb200_wrap_fini(void) { report(); }
