/*
 * svtb200_wrap_tx.c -- seam S3 bound into the UNMODIFIED reference encoder, call by call (SVT_VP9_GPU_ENCODE_PASS=1).
 *
 * The transform kernels of perform_coding_loop() / perform_inv_trans_add() (Source/Lib/Codec/EbEncDecProcess.c:365-587, :589-655; the
 * same file's mode-decision full loop uses them too) are reached through two mechanisms in the reference, and both are used here:
 *   - run-time dispatch pointers filled by setup_rtcd_internal() / setup_rtcd_internal_vp9() (Source/Lib/VPX/vpx_dsp_rtcd.h:1123,
 *     vp9_rtcd.h:138; called from eb_vp9_init_encoder, Source/Lib/Codec/EbEncHandle.c:806-807): eb_vp9_fdct32x32, eb_vpx_partial_fdct32x32,
 *     eb_vp9_fht16x16 / 8x8 / 4x4, eb_vp9_quantize_b, eb_vp9_quantize_b_32x32, eb_vp9_iht4x4_16_add.  svtb200_tx_install() -- called from the
 *     wrapped eb_vp9_picture_analysis_context_ctor, i.e. inside eb_vp9_init_encoder AFTER the dispatch tables were filled -- replaces them;
 *   - plain functions of Source/Lib/VPX/vp9_idct.c (:111-190) called across objects: eb_vp9_idct32x32_add, eb_vp9_iht16x16_add,
 *     eb_vp9_iht8x8_add, eb_vp9_idct4x4_add -> link-time wrappers (-Wl,--wrap=...).
 * Every call becomes one svtb200_fwd_txfm_batch / svtb200_quantize_batch / svtb200_inv_txfm_add_batch of ONE block on a context of its own:
 * a host->device->host round trip per call, which is what binding S3 at the reference's own granularity costs -- this
 * mode exists to prove that the device transforms are a drop-in for every size / type / eob the encoder really produces (identical
 * bitstream, tests/test_dropin_encoder.py), not as a fast path; the fast path is the batched per-picture chain (svtb200_inter_chain_dev).
 * Any call the device cannot serve (no context, an unknown scan table, skip_block) falls through to the reference's own kernel.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "EbDefinitions.h"
#include "vpx_dsp_rtcd.h"
#include "vp9_rtcd.h"
#include "vp9_scan.h"
#include "vp9_idct.h"
#include "svtb200.h"

static SvtB200Ctx     *g_tx;                 /* its own context: never queues behind motion estimation or deblocking */
static int             g_tx_on;              /* dispatch pointers replaced */
static unsigned long   g_n_fwd, g_n_quant, g_n_inv, g_n_fall;
static pthread_mutex_t g_tx_mu = PTHREAD_MUTEX_INITIALIZER;

/* the reference's own kernels, as its dispatch tables held them */
static void (*r_fdct32x32)(const int16_t *, tran_low_t *, int);
static void (*r_partial_fdct32x32)(const int16_t *, tran_low_t *, int);
static void (*r_fht16x16)(const int16_t *, tran_low_t *, int, int);
static void (*r_fht8x8)(const int16_t *, tran_low_t *, int, int);
static void (*r_fht4x4)(const int16_t *, tran_low_t *, int, int);
static void (*r_iht4x4_16_add)(const tran_low_t *, uint8_t *, int, int);
typedef void (*QuantFn)(const tran_low_t *, intptr_t, int, const int16_t *, const int16_t *, const int16_t *, const int16_t *, tran_low_t *,
                        tran_low_t *, const int16_t *, uint16_t *, const int16_t *, const int16_t *);
static QuantFn r_quantize_b, r_quantize_b_32x32;

void __real_eb_vp9_idct32x32_add(const tran_low_t *input, uint8_t *dest, int stride, int eob);
void __real_eb_vp9_iht16x16_add(TX_TYPE tx_type, const tran_low_t *input, uint8_t *dest, int stride, int eob);
void __real_eb_vp9_iht8x8_add(TX_TYPE tx_type, const tran_low_t *input, uint8_t *dest, int stride, int eob);
void __real_eb_vp9_idct4x4_add(const tran_low_t *input, uint8_t *dest, int stride, int eob);

static void count(unsigned long *c) { __atomic_add_fetch(c, 1, __ATOMIC_RELAXED); }

/* ---- forward ---- */
static int fwd(int ts, int tx_type, int partial, const int16_t *in, tran_low_t *out, int stride) {
    const int n = 4 << ts;
    int16_t blk[32 * 32];
    for (int r = 0; r < n; r++) memcpy(blk + r * n, in + (size_t)r * stride, (size_t)n * 2);
    if (svtb200_fwd_txfm_batch(g_tx, ts, tx_type, partial, 1, blk, out) != SVTB200_OK) { count(&g_n_fall); return 0; }
    count(&g_n_fwd);
    return 1;
}
static void g_fdct32x32(const int16_t *in, tran_low_t *out, int stride) { if (!fwd(3, 0, 0, in, out, stride)) r_fdct32x32(in, out, stride); }
static void g_partial_fdct32x32(const int16_t *in, tran_low_t *out, int stride) { if (!fwd(3, 0, 1, in, out, stride)) r_partial_fdct32x32(in, out, stride); }
static void g_fht16x16(const int16_t *in, tran_low_t *out, int stride, int t) { if (!fwd(2, t, 0, in, out, stride)) r_fht16x16(in, out, stride, t); }
static void g_fht8x8(const int16_t *in, tran_low_t *out, int stride, int t) { if (!fwd(1, t, 0, in, out, stride)) r_fht8x8(in, out, stride, t); }
static void g_fht4x4(const int16_t *in, tran_low_t *out, int stride, int t) { if (!fwd(0, t, 0, in, out, stride)) r_fht4x4(in, out, stride, t); }

/* ---- quantise: the scan table of the call tells the transform type (the device holds the same tables, svtb200_set_scan) ---- */
static int quant(int ts, const tran_low_t *coeff, int skip_block, const int16_t *zbin, const int16_t *round, const int16_t *quant_, const int16_t *shift,
                 tran_low_t *q, tran_low_t *dq, const int16_t *deq, uint16_t *eob, const int16_t *scan) {
    int tt = -1;
    if (skip_block) return 0;
    if (ts == 3) tt = scan == eb_vp9_default_scan_orders[TX_32X32].scan ? 0 : -1;
    else for (int t = 0; t < 4; t++) if (scan == eb_vp9_scan_orders[ts][t].scan) { tt = t; break; }
    if (tt < 0) { count(&g_n_fall); return 0; }
    SvtB200QuantRow row;
    for (int k = 0; k < 2; k++) { row.zbin[k] = zbin[k]; row.round[k] = round[k]; row.quant[k] = quant_[k]; row.quant_shift[k] = shift[k]; row.dequant[k] = deq[k]; }
    if (svtb200_quantize_batch(g_tx, ts, tt, 1, coeff, &row, 1, NULL, q, dq, eob) != SVTB200_OK) { count(&g_n_fall); return 0; }
    count(&g_n_quant);
    return 1;
}
static void g_quantize_b(const tran_low_t *c, intptr_t n, int skip, const int16_t *zb, const int16_t *rd, const int16_t *qt, const int16_t *sh, tran_low_t *q,
                         tran_low_t *dq, const int16_t *deq, uint16_t *eob, const int16_t *scan, const int16_t *iscan) {
    const int ts = n == 16 ? 0 : n == 64 ? 1 : n == 256 ? 2 : -1;
    if (ts < 0 || !quant(ts, c, skip, zb, rd, qt, sh, q, dq, deq, eob, scan)) r_quantize_b(c, n, skip, zb, rd, qt, sh, q, dq, deq, eob, scan, iscan);
}
static void g_quantize_b_32x32(const tran_low_t *c, intptr_t n, int skip, const int16_t *zb, const int16_t *rd, const int16_t *qt, const int16_t *sh,
                               tran_low_t *q, tran_low_t *dq, const int16_t *deq, uint16_t *eob, const int16_t *scan, const int16_t *iscan) {
    if (n != 1024 || !quant(3, c, skip, zb, rd, qt, sh, q, dq, deq, eob, scan)) r_quantize_b_32x32(c, n, skip, zb, rd, qt, sh, q, dq, deq, eob, scan, iscan);
}

/* ---- inverse + reconstruction ---- */
static int inv(int ts, int tx_type, const tran_low_t *input, uint8_t *dest, int stride, int eob) {
    if (!g_tx_on) return 0;
    const int n = 4 << ts;
    uint8_t blk[32 * 32];
    uint16_t e = (uint16_t)eob;
    for (int r = 0; r < n; r++) memcpy(blk + r * n, dest + (size_t)r * stride, (size_t)n);
    if (svtb200_inv_txfm_add_batch(g_tx, ts, tx_type, 1, input, &e, blk) != SVTB200_OK) { count(&g_n_fall); return 0; }
    for (int r = 0; r < n; r++) memcpy(dest + (size_t)r * stride, blk + r * n, (size_t)n);
    count(&g_n_inv);
    return 1;
}
void __wrap_eb_vp9_idct32x32_add(const tran_low_t *input, uint8_t *dest, int stride, int eob) {
    if (!inv(3, 0, input, dest, stride, eob)) __real_eb_vp9_idct32x32_add(input, dest, stride, eob);
}
void __wrap_eb_vp9_iht16x16_add(TX_TYPE t, const tran_low_t *input, uint8_t *dest, int stride, int eob) {
    if (!inv(2, (int)t, input, dest, stride, eob)) __real_eb_vp9_iht16x16_add(t, input, dest, stride, eob);
}
void __wrap_eb_vp9_iht8x8_add(TX_TYPE t, const tran_low_t *input, uint8_t *dest, int stride, int eob) {
    if (!inv(1, (int)t, input, dest, stride, eob)) __real_eb_vp9_iht8x8_add(t, input, dest, stride, eob);
}
void __wrap_eb_vp9_idct4x4_add(const tran_low_t *input, uint8_t *dest, int stride, int eob) {
    if (!inv(0, 0, input, dest, stride, eob)) __real_eb_vp9_idct4x4_add(input, dest, stride, eob);
}
static void g_iht4x4_16_add(const tran_low_t *input, uint8_t *dest, int stride, int t) { if (!inv(0, t, input, dest, stride, 16)) r_iht4x4_16_add(input, dest, stride, t); }

/* Called inside eb_vp9_init_encoder, after its setup_rtcd_internal*() calls.  device: the first entry of SVT_VP9_GPU. */
void svtb200_tx_install(int device) {
    const char *e = getenv("SVT_VP9_GPU_ENCODE_PASS");
    if (!(e && e[0] == '1')) return;
    pthread_mutex_lock(&g_tx_mu);
    if (!g_tx_on && !g_tx && svtb200_create(device, &g_tx) == SVTB200_OK) {
        int ok = svtb200_set_scan(g_tx, 3, 0, eb_vp9_default_scan_orders[TX_32X32].scan, eb_vp9_default_scan_orders[TX_32X32].iscan) == SVTB200_OK;
        for (int ts = 0; ts < 3 && ok; ts++)
            for (int t = 0; t < 4 && ok; t++) ok = svtb200_set_scan(g_tx, ts, t, eb_vp9_scan_orders[ts][t].scan, eb_vp9_scan_orders[ts][t].iscan) == SVTB200_OK;
        if (ok) {
            r_fdct32x32 = eb_vp9_fdct32x32; r_partial_fdct32x32 = eb_vpx_partial_fdct32x32;
            r_fht16x16 = eb_vp9_fht16x16; r_fht8x8 = eb_vp9_fht8x8; r_fht4x4 = eb_vp9_fht4x4;
            r_quantize_b = eb_vp9_quantize_b; r_quantize_b_32x32 = eb_vp9_quantize_b_32x32; r_iht4x4_16_add = eb_vp9_iht4x4_16_add;
            eb_vp9_fdct32x32 = g_fdct32x32; eb_vpx_partial_fdct32x32 = g_partial_fdct32x32;
            eb_vp9_fht16x16 = g_fht16x16; eb_vp9_fht8x8 = g_fht8x8; eb_vp9_fht4x4 = g_fht4x4;
            eb_vp9_quantize_b = g_quantize_b; eb_vp9_quantize_b_32x32 = g_quantize_b_32x32; eb_vp9_iht4x4_16_add = g_iht4x4_16_add;
            __atomic_store_n(&g_tx_on, 1, __ATOMIC_RELEASE);
        } else {
            svtb200_destroy(g_tx);
            g_tx = NULL;
        }
    }
    pthread_mutex_unlock(&g_tx_mu);
}

/* The reference's kernels back into its dispatch tables (eb_vp9_deinit_encoder: worker threads are gone). */
void svtb200_tx_uninstall(void) {
    pthread_mutex_lock(&g_tx_mu);
    if (g_tx_on) {
        eb_vp9_fdct32x32 = r_fdct32x32; eb_vpx_partial_fdct32x32 = r_partial_fdct32x32;
        eb_vp9_fht16x16 = r_fht16x16; eb_vp9_fht8x8 = r_fht8x8; eb_vp9_fht4x4 = r_fht4x4;
        eb_vp9_quantize_b = r_quantize_b; eb_vp9_quantize_b_32x32 = r_quantize_b_32x32; eb_vp9_iht4x4_16_add = r_iht4x4_16_add;
        __atomic_store_n(&g_tx_on, 0, __ATOMIC_RELEASE);
        svtb200_destroy(g_tx);
        g_tx = NULL;
    }
    pthread_mutex_unlock(&g_tx_mu);
}

void svtb200_tx_report(FILE *f) {
    fprintf(f, "tx_forward %lu\ntx_quantize %lu\ntx_inverse %lu\ntx_fallbacks %lu\n", g_n_fwd, g_n_quant, g_n_inv, g_n_fall);
}
