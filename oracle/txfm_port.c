/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle ("port") for the transform / quantisation hot path.
 *
 * Scalar restatement of the reference's integer transforms and quantiser, organised around the
 * recursive even/odd structure of the butterfly networks (an N-point transform = the N/2-point
 * transform of the folded input + an N-specific rotation network).  Rounding, truncation and wrap
 * points follow the reference exactly; they are what makes these maps non-linear.
 *
 * Parity pinning: tests/test_txfm_oracle_vs_ref.py calls the reference's own _c functions in
 * oracle/_ref/libSvtVp9Enc_ref.so on random and structured inputs and requires identical output;
 * the fixtures under tests/golden/txfm_*.npz were generated the same way (oracle/make_golden_txfm.py).
 *
 * Reference functions followed (Source/Lib/VPX):
 *   eb_vp9_fdct4x4_c / 8x8_c / 16x16_c / 32x32_c, eb_vp9_fdct32, eb_vpx_partial_fdct32x32_c
 *                                              fwd_txfm.c:15, :90, :183, :708, :385, :1051
 *   eb_vp9_fht4x4_c / 8x8_c / 16x16_c, fdct4/8/16, fadst4/8/16
 *                                              vp9_dct.c:525, :653, :734, :18-228, :230-523
 *   eb_vp9_quantize_b_c, eb_vp9_quantize_b_32x32_c
 *                                              quantize.c:112, :206
 *   eb_vp9_idct4_c/8_c/16_c/32_c, eb_vp9_iadst4_c/8_c/16_c, *_add_c 2-D drivers
 *                                              inv_txfm.c:93-1256
 *   eb_vp9_idct*_add / eb_vp9_iht*_add dispatch  vp9_idct.c:18-190
 *   constants: cospi_k_64 = round(16384 cos(k pi/64)), sinpi_k_9 = round(16384 sqrt2 sin(k pi/9) 2/3)
 *                                              txfm_common.h:17-64
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

static const int C64[32] = {16384, 16364, 16305, 16207, 16069, 15893, 15679, 15426, 15137, 14811, 14449,
                            14053, 13623, 13160, 12665, 12140, 11585, 11003, 10394, 9760,  9102,  8423,
                            7723,  7005,  6270,  5520,  4756,  3981,  3196,  2404,  1606,  804};
static const int S9[5] = {0, 5283, 9929, 13377, 15212};

static inline int rnd14(int v) { return (v + (1 << 13)) >> 14; }
static inline int w16(int v) { return (int16_t)v; }

/* =============================== forward DCT ================================================= */
/* All forward kernels work on int arrays; the caller applies the int16 stores of the reference. */

static void fdct4_1d(const int *in, int *out) {
    const int s0 = in[0] + in[3], s1 = in[1] + in[2], s2 = in[1] - in[2], s3 = in[0] - in[3];
    out[0] = rnd14((s0 + s1) * C64[16]);
    out[2] = rnd14((s0 - s1) * C64[16]);
    out[1] = rnd14(s2 * C64[24] + s3 * C64[8]);
    out[3] = rnd14(-s2 * C64[8] + s3 * C64[24]);
}

/* odd half of the 8-point DCT: d[i] = x[3-i] - x[4+i] (i.e. s4..s7 of the reference) */
static void fdct8_odd(const int *d, int *o1, int *o3, int *o5, int *o7) {
    const int t2 = rnd14((d[2] - d[1]) * C64[16]), t3 = rnd14((d[2] + d[1]) * C64[16]);
    const int x0 = d[0] + t2, x1 = d[0] - t2, x2 = d[3] - t3, x3 = d[3] + t3;
    *o1 = rnd14(x0 * C64[28] + x3 * C64[4]);
    *o5 = rnd14(x1 * C64[12] + x2 * C64[20]);
    *o3 = rnd14(x2 * C64[12] + x1 * -C64[20]);
    *o7 = rnd14(x3 * C64[28] + x0 * -C64[4]);
}

static void fdct8_1d(const int *in, int *out) {
    int e[4], d[4], ev[4];
    for (int i = 0; i < 4; i++) { e[i] = in[i] + in[7 - i]; }
    d[0] = in[3] - in[4]; d[1] = in[2] - in[5]; d[2] = in[1] - in[6]; d[3] = in[0] - in[7];
    fdct4_1d(e, ev);
    out[0] = ev[0]; out[2] = ev[1]; out[4] = ev[2]; out[6] = ev[3];
    fdct8_odd(d, &out[1], &out[3], &out[5], &out[7]);
}

/* odd half of the 16-point DCT: s[i] = x[7-i] - x[8+i] */
static void fdct16_odd(const int *s, int *out /* stride 2, starting at out[1] */) {
    int a2 = rnd14((s[5] - s[2]) * C64[16]), a3 = rnd14((s[4] - s[3]) * C64[16]);
    int a4 = rnd14((s[4] + s[3]) * C64[16]), a5 = rnd14((s[5] + s[2]) * C64[16]);
    int b0 = s[0] + a3, b1 = s[1] + a2, b2 = s[1] - a2, b3 = s[0] - a3;
    int b4 = s[7] - a4, b5 = s[6] - a5, b6 = s[6] + a5, b7 = s[7] + a4;
    int c1 = rnd14(b1 * -C64[8] + b6 * C64[24]), c2 = rnd14(b2 * C64[24] + b5 * C64[8]);
    int c5 = rnd14(b2 * C64[8] - b5 * C64[24]), c6 = rnd14(b1 * C64[24] + b6 * C64[8]);
    int d0 = b0 + c1, d1 = b0 - c1, d2 = b3 + c2, d3 = b3 - c2;
    int d4 = b4 - c5, d5 = b4 + c5, d6 = b7 - c6, d7 = b7 + c6;
    out[1]  = rnd14(d0 * C64[30] + d7 * C64[2]);
    out[9]  = rnd14(d1 * C64[14] + d6 * C64[18]);
    out[5]  = rnd14(d2 * C64[22] + d5 * C64[10]);
    out[13] = rnd14(d3 * C64[6] + d4 * C64[26]);
    out[3]  = rnd14(d3 * -C64[26] + d4 * C64[6]);
    out[11] = rnd14(d2 * -C64[10] + d5 * C64[22]);
    out[7]  = rnd14(d1 * -C64[18] + d6 * C64[14]);
    out[15] = rnd14(d0 * -C64[2] + d7 * C64[30]);
}

static void fdct16_1d(const int *in, int *out) {
    int e[8], s[8], ev[8];
    for (int i = 0; i < 8; i++) { e[i] = in[i] + in[15 - i]; s[i] = in[7 - i] - in[8 + i]; }
    fdct8_1d(e, ev);
    for (int i = 0; i < 8; i++) out[2 * i] = ev[i];
    fdct16_odd(s, out);
}

/* 32-point DCT (eb_vp9_fdct32 with round == 0); every multiply is rounded, adds are exact */
static void fdct32_1d(const int *in, int *out) {
    int a[32], b[32];
    for (int i = 0; i < 16; i++) { a[i] = in[i] + in[31 - i]; a[16 + i] = in[15 - i] - in[16 + i]; }
    /* stage 2 */
    for (int i = 0; i < 8; i++) { b[i] = a[i] + a[15 - i]; b[8 + i] = a[7 - i] - a[8 + i]; }
    b[16] = a[16]; b[17] = a[17]; b[18] = a[18]; b[19] = a[19];
    for (int i = 0; i < 4; i++) { b[20 + i] = rnd14((a[27 - i] - a[20 + i]) * C64[16]); b[27 - i] = rnd14((a[27 - i] + a[20 + i]) * C64[16]); }
    b[28] = a[28]; b[29] = a[29]; b[30] = a[30]; b[31] = a[31];
    /* stage 3 */
    for (int i = 0; i < 4; i++) { a[i] = b[i] + b[7 - i]; a[4 + i] = b[3 - i] - b[4 + i]; }
    a[8] = b[8]; a[9] = b[9];
    a[10] = rnd14((b[13] - b[10]) * C64[16]); a[13] = rnd14((b[13] + b[10]) * C64[16]);
    a[11] = rnd14((b[12] - b[11]) * C64[16]); a[12] = rnd14((b[12] + b[11]) * C64[16]);
    a[14] = b[14]; a[15] = b[15];
    for (int i = 0; i < 4; i++) {
        a[16 + i] = b[16 + i] + b[23 - i]; a[20 + i] = b[19 - i] - b[20 + i];
        a[24 + i] = b[31 - i] - b[24 + i]; a[28 + i] = b[28 + i] + b[27 - i];
    }
    /* stage 4 */
    b[0] = a[0] + a[3]; b[1] = a[1] + a[2]; b[2] = a[1] - a[2]; b[3] = a[0] - a[3];
    b[4] = a[4]; b[5] = rnd14((a[6] - a[5]) * C64[16]); b[6] = rnd14((a[6] + a[5]) * C64[16]); b[7] = a[7];
    b[8] = a[8] + a[11]; b[9] = a[9] + a[10]; b[10] = a[9] - a[10]; b[11] = a[8] - a[11];
    b[12] = a[15] - a[12]; b[13] = a[14] - a[13]; b[14] = a[14] + a[13]; b[15] = a[15] + a[12];
    b[16] = a[16]; b[17] = a[17];
    b[18] = rnd14(a[18] * -C64[8] + a[29] * C64[24]); b[19] = rnd14(a[19] * -C64[8] + a[28] * C64[24]);
    b[20] = rnd14(a[20] * -C64[24] + a[27] * -C64[8]); b[21] = rnd14(a[21] * -C64[24] + a[26] * -C64[8]);
    b[22] = a[22]; b[23] = a[23]; b[24] = a[24]; b[25] = a[25];
    b[26] = rnd14(a[26] * C64[24] + a[21] * -C64[8]); b[27] = rnd14(a[27] * C64[24] + a[20] * -C64[8]);
    b[28] = rnd14(a[28] * C64[8] + a[19] * C64[24]); b[29] = rnd14(a[29] * C64[8] + a[18] * C64[24]);
    b[30] = a[30]; b[31] = a[31];
    /* stage 5 */
    a[0] = rnd14((b[0] + b[1]) * C64[16]); a[1] = rnd14((b[0] - b[1]) * C64[16]);
    a[2] = rnd14(b[2] * C64[24] + b[3] * C64[8]); a[3] = rnd14(b[3] * C64[24] - b[2] * C64[8]);
    a[4] = b[4] + b[5]; a[5] = b[4] - b[5]; a[6] = b[7] - b[6]; a[7] = b[7] + b[6];
    a[8] = b[8];
    a[9] = rnd14(b[9] * -C64[8] + b[14] * C64[24]); a[10] = rnd14(b[10] * -C64[24] + b[13] * -C64[8]);
    a[11] = b[11]; a[12] = b[12];
    a[13] = rnd14(b[13] * C64[24] + b[10] * -C64[8]); a[14] = rnd14(b[14] * C64[8] + b[9] * C64[24]);
    a[15] = b[15];
    a[16] = b[16] + b[19]; a[17] = b[17] + b[18]; a[18] = b[17] - b[18]; a[19] = b[16] - b[19];
    a[20] = b[23] - b[20]; a[21] = b[22] - b[21]; a[22] = b[22] + b[21]; a[23] = b[23] + b[20];
    a[24] = b[24] + b[27]; a[25] = b[25] + b[26]; a[26] = b[25] - b[26]; a[27] = b[24] - b[27];
    a[28] = b[31] - b[28]; a[29] = b[30] - b[29]; a[30] = b[30] + b[29]; a[31] = b[31] + b[28];
    /* stage 6 */
    b[0] = a[0]; b[1] = a[1]; b[2] = a[2]; b[3] = a[3];
    b[4] = rnd14(a[4] * C64[28] + a[7] * C64[4]); b[5] = rnd14(a[5] * C64[12] + a[6] * C64[20]);
    b[6] = rnd14(a[6] * C64[12] + a[5] * -C64[20]); b[7] = rnd14(a[7] * C64[28] + a[4] * -C64[4]);
    b[8] = a[8] + a[9]; b[9] = a[8] - a[9]; b[10] = a[11] - a[10]; b[11] = a[11] + a[10];
    b[12] = a[12] + a[13]; b[13] = a[12] - a[13]; b[14] = a[15] - a[14]; b[15] = a[15] + a[14];
    b[16] = a[16];
    b[17] = rnd14(a[17] * -C64[4] + a[30] * C64[28]); b[18] = rnd14(a[18] * -C64[28] + a[29] * -C64[4]);
    b[19] = a[19]; b[20] = a[20];
    b[21] = rnd14(a[21] * -C64[20] + a[26] * C64[12]); b[22] = rnd14(a[22] * -C64[12] + a[25] * -C64[20]);
    b[23] = a[23]; b[24] = a[24];
    b[25] = rnd14(a[25] * C64[12] + a[22] * -C64[20]); b[26] = rnd14(a[26] * C64[20] + a[21] * C64[12]);
    b[27] = a[27]; b[28] = a[28];
    b[29] = rnd14(a[29] * C64[28] + a[18] * -C64[4]); b[30] = rnd14(a[30] * C64[4] + a[17] * C64[28]);
    b[31] = a[31];
    /* stage 7 */
    for (int i = 0; i < 8; i++) a[i] = b[i];
    a[8]  = rnd14(b[8] * C64[30] + b[15] * C64[2]);   a[9]  = rnd14(b[9] * C64[14] + b[14] * C64[18]);
    a[10] = rnd14(b[10] * C64[22] + b[13] * C64[10]); a[11] = rnd14(b[11] * C64[6] + b[12] * C64[26]);
    a[12] = rnd14(b[12] * C64[6] + b[11] * -C64[26]); a[13] = rnd14(b[13] * C64[22] + b[10] * -C64[10]);
    a[14] = rnd14(b[14] * C64[14] + b[9] * -C64[18]); a[15] = rnd14(b[15] * C64[30] + b[8] * -C64[2]);
    a[16] = b[16] + b[17]; a[17] = b[16] - b[17]; a[18] = b[19] - b[18]; a[19] = b[19] + b[18];
    a[20] = b[20] + b[21]; a[21] = b[20] - b[21]; a[22] = b[23] - b[22]; a[23] = b[23] + b[22];
    a[24] = b[24] + b[25]; a[25] = b[24] - b[25]; a[26] = b[27] - b[26]; a[27] = b[27] + b[26];
    a[28] = b[28] + b[29]; a[29] = b[28] - b[29]; a[30] = b[31] - b[30]; a[31] = b[31] + b[30];
    /* output permutation (bit reversal) and the last rotations */
    out[0] = a[0]; out[16] = a[1]; out[8] = a[2]; out[24] = a[3]; out[4] = a[4]; out[20] = a[5]; out[12] = a[6]; out[28] = a[7];
    out[2] = a[8]; out[18] = a[9]; out[10] = a[10]; out[26] = a[11]; out[6] = a[12]; out[22] = a[13]; out[14] = a[14]; out[30] = a[15];
    out[1]  = rnd14(a[16] * C64[31] + a[31] * C64[1]);   out[17] = rnd14(a[17] * C64[15] + a[30] * C64[17]);
    out[9]  = rnd14(a[18] * C64[23] + a[29] * C64[9]);   out[25] = rnd14(a[19] * C64[7] + a[28] * C64[25]);
    out[5]  = rnd14(a[20] * C64[27] + a[27] * C64[5]);   out[21] = rnd14(a[21] * C64[11] + a[26] * C64[21]);
    out[13] = rnd14(a[22] * C64[19] + a[25] * C64[13]);  out[29] = rnd14(a[23] * C64[3] + a[24] * C64[29]);
    out[3]  = rnd14(a[24] * C64[3] + a[23] * -C64[29]);  out[19] = rnd14(a[25] * C64[19] + a[22] * -C64[13]);
    out[11] = rnd14(a[26] * C64[11] + a[21] * -C64[21]); out[27] = rnd14(a[27] * C64[27] + a[20] * -C64[5]);
    out[7]  = rnd14(a[28] * C64[7] + a[19] * -C64[25]);  out[23] = rnd14(a[29] * C64[23] + a[18] * -C64[9]);
    out[15] = rnd14(a[30] * C64[15] + a[17] * -C64[17]); out[31] = rnd14(a[31] * C64[31] + a[16] * -C64[1]);
}

/* =============================== forward ADST ================================================ */
static void fadst4_1d(const int *in, int *out) {
    int x0 = in[0], x1 = in[1], x2 = in[2], x3 = in[3];
    if (!(x0 | x1 | x2 | x3)) { out[0] = out[1] = out[2] = out[3] = 0; return; }
    int s0 = S9[1] * x0, s1 = S9[4] * x0, s2 = S9[2] * x1, s3 = S9[1] * x1, s4 = S9[3] * x2, s5 = S9[4] * x3, s6 = S9[2] * x3;
    int s7 = x0 + x1 - x3;
    x0 = s0 + s2 + s5; x1 = S9[3] * s7; x2 = s1 - s3 + s6; x3 = s4;
    out[0] = rnd14(x0 + x3); out[1] = rnd14(x1); out[2] = rnd14(x2 - x3); out[3] = rnd14(x2 - x0 + x3);
}

static void fadst8_1d(const int *in, int *out) {
    int x0 = in[7], x1 = in[0], x2 = in[5], x3 = in[2], x4 = in[3], x5 = in[4], x6 = in[1], x7 = in[6];
    int s0 = C64[2] * x0 + C64[30] * x1, s1 = C64[30] * x0 - C64[2] * x1;
    int s2 = C64[10] * x2 + C64[22] * x3, s3 = C64[22] * x2 - C64[10] * x3;
    int s4 = C64[18] * x4 + C64[14] * x5, s5 = C64[14] * x4 - C64[18] * x5;
    int s6 = C64[26] * x6 + C64[6] * x7, s7 = C64[6] * x6 - C64[26] * x7;
    x0 = rnd14(s0 + s4); x1 = rnd14(s1 + s5); x2 = rnd14(s2 + s6); x3 = rnd14(s3 + s7);
    x4 = rnd14(s0 - s4); x5 = rnd14(s1 - s5); x6 = rnd14(s2 - s6); x7 = rnd14(s3 - s7);
    s0 = x0; s1 = x1; s2 = x2; s3 = x3;
    s4 = C64[8] * x4 + C64[24] * x5; s5 = C64[24] * x4 - C64[8] * x5;
    s6 = -C64[24] * x6 + C64[8] * x7; s7 = C64[8] * x6 + C64[24] * x7;
    x0 = s0 + s2; x1 = s1 + s3; x2 = s0 - s2; x3 = s1 - s3;
    x4 = rnd14(s4 + s6); x5 = rnd14(s5 + s7); x6 = rnd14(s4 - s6); x7 = rnd14(s5 - s7);
    s2 = C64[16] * (x2 + x3); s3 = C64[16] * (x2 - x3); s6 = C64[16] * (x6 + x7); s7 = C64[16] * (x6 - x7);
    x2 = rnd14(s2); x3 = rnd14(s3); x6 = rnd14(s6); x7 = rnd14(s7);
    out[0] = x0; out[1] = -x4; out[2] = x6; out[3] = -x2; out[4] = x3; out[5] = -x7; out[6] = x5; out[7] = -x1;
}

static void fadst16_1d(const int *in, int *out) {
    int x[16], s[16];
    static const uint8_t perm[16] = {15, 0, 13, 2, 11, 4, 9, 6, 7, 8, 5, 10, 3, 12, 1, 14};
    static const uint8_t ck[8] = {1, 5, 9, 13, 17, 21, 25, 29};
    for (int i = 0; i < 16; i++) x[i] = in[perm[i]];
    for (int i = 0; i < 8; i++) { /* stage 1 rotations by odd multiples of pi/64 */
        s[2 * i]     = x[2 * i] * C64[ck[i]] + x[2 * i + 1] * C64[32 - ck[i]];
        s[2 * i + 1] = x[2 * i] * C64[32 - ck[i]] - x[2 * i + 1] * C64[ck[i]];
    }
    for (int i = 0; i < 8; i++) { x[i] = rnd14(s[i] + s[i + 8]); x[i + 8] = rnd14(s[i] - s[i + 8]); }
    /* stage 2 */
    for (int i = 0; i < 8; i++) s[i] = x[i];
    s[8] = x[8] * C64[4] + x[9] * C64[28];    s[9] = x[8] * C64[28] - x[9] * C64[4];
    s[10] = x[10] * C64[20] + x[11] * C64[12]; s[11] = x[10] * C64[12] - x[11] * C64[20];
    s[12] = -x[12] * C64[28] + x[13] * C64[4]; s[13] = x[12] * C64[4] + x[13] * C64[28];
    s[14] = -x[14] * C64[12] + x[15] * C64[20]; s[15] = x[14] * C64[20] + x[15] * C64[12];
    for (int i = 0; i < 4; i++) { x[i] = s[i] + s[i + 4]; x[i + 4] = s[i] - s[i + 4]; }
    for (int i = 8; i < 12; i++) { x[i] = rnd14(s[i] + s[i + 4]); x[i + 4] = rnd14(s[i] - s[i + 4]); }
    /* stage 3 */
    for (int h = 0; h < 16; h += 8) {
        s[h] = x[h]; s[h + 1] = x[h + 1]; s[h + 2] = x[h + 2]; s[h + 3] = x[h + 3];
        s[h + 4] = x[h + 4] * C64[8] + x[h + 5] * C64[24];  s[h + 5] = x[h + 4] * C64[24] - x[h + 5] * C64[8];
        s[h + 6] = -x[h + 6] * C64[24] + x[h + 7] * C64[8]; s[h + 7] = x[h + 6] * C64[8] + x[h + 7] * C64[24];
        x[h] = s[h] + s[h + 2]; x[h + 1] = s[h + 1] + s[h + 3]; x[h + 2] = s[h] - s[h + 2]; x[h + 3] = s[h + 1] - s[h + 3];
        x[h + 4] = rnd14(s[h + 4] + s[h + 6]); x[h + 5] = rnd14(s[h + 5] + s[h + 7]);
        x[h + 6] = rnd14(s[h + 4] - s[h + 6]); x[h + 7] = rnd14(s[h + 5] - s[h + 7]);
    }
    /* stage 4 */
    s[2] = -C64[16] * (x[2] + x[3]);  s[3] = C64[16] * (x[2] - x[3]);
    s[6] = C64[16] * (x[6] + x[7]);   s[7] = C64[16] * (-x[6] + x[7]);
    s[10] = C64[16] * (x[10] + x[11]); s[11] = C64[16] * (-x[10] + x[11]);
    s[14] = -C64[16] * (x[14] + x[15]); s[15] = C64[16] * (x[14] - x[15]);
    x[2] = rnd14(s[2]); x[3] = rnd14(s[3]); x[6] = rnd14(s[6]); x[7] = rnd14(s[7]);
    x[10] = rnd14(s[10]); x[11] = rnd14(s[11]); x[14] = rnd14(s[14]); x[15] = rnd14(s[15]);
    out[0] = x[0]; out[1] = -x[8]; out[2] = x[12]; out[3] = -x[4]; out[4] = x[6]; out[5] = x[14]; out[6] = x[10]; out[7] = x[2];
    out[8] = x[3]; out[9] = x[11]; out[10] = x[15]; out[11] = x[7]; out[12] = x[5]; out[13] = -x[13]; out[14] = x[9]; out[15] = -x[1];
}

typedef void (*fwd1d)(const int *, int *);

/* the 1-D kernels of vp9_dct.c read and write tran_low_t (int16) arrays */
static void run1d_i16(fwd1d f, int n, const int16_t *in, int16_t *out) {
    int a[32], b[32];
    for (int i = 0; i < n; i++) a[i] = in[i];
    f(a, b);
    for (int i = 0; i < n; i++) out[i] = (int16_t)b[i];
}

/* tx_size 0..3 (4,8,16,32), tx_type 0..3 (bit0: ADST on columns, bit1: ADST on rows), in: n*n residuals (stride), out n*n */
void svtb200_oracle_fwd_txfm(int tx_size, int tx_type, int partial32, const int16_t *in, int stride, int16_t *out) {
    const int n = 4 << tx_size;
    if (tx_size == 3) { /* always DCT; 32-bit intermediate, per-pass rounding rules */
        int mid[32 * 32], ti[32], to[32];
        const int keep = partial32 ? 16 : 32;
        for (int c = 0; c < 32; c++) {
            for (int r = 0; r < 32; r++) ti[r] = in[r * stride + c] * 4;
            fdct32_1d(ti, to);
            for (int r = 0; r < keep; r++) mid[r * 32 + c] = (to[r] + 1 + (to[r] > 0)) >> 2;
        }
        if (partial32) memset(out, 0, 32 * 32 * sizeof(int16_t));
        for (int r = 0; r < keep; r++) {
            fdct32_1d(mid + r * 32, to);
            for (int c = 0; c < keep; c++) out[r * 32 + c] = (int16_t)((to[c] + 1 + (to[c] < 0)) >> 2);
        }
        return;
    }
    if (tx_type == 0) { /* pure DCT: the fused 2-D routines of fwd_txfm.c */
        int16_t mid[16 * 16];
        int ti[16], to[16];
        fwd1d f = tx_size == 0 ? fdct4_1d : tx_size == 1 ? fdct8_1d : fdct16_1d;
        for (int c = 0; c < n; c++) { /* columns -> transposed intermediate */
            for (int r = 0; r < n; r++) ti[r] = in[r * stride + c] * (tx_size == 0 ? 16 : 4);
            if (tx_size == 0 && c == 0 && ti[0]) ti[0]++;
            f(ti, to);
            for (int r = 0; r < n; r++) mid[c * n + r] = (int16_t)to[r];
        }
        for (int c = 0; c < n; c++) { /* "columns" of the transposed intermediate = rows */
            for (int r = 0; r < n; r++) ti[r] = tx_size == 2 ? ((mid[r * n + c] + 1) >> 2) : mid[r * n + c];
            f(ti, to);
            for (int r = 0; r < n; r++) out[c * n + r] = (int16_t)to[r];
        }
        if (tx_size == 0) for (int i = 0; i < 16; i++) out[i] = (int16_t)((out[i] + 1) >> 2);
        if (tx_size == 1) for (int i = 0; i < 64; i++) out[i] = (int16_t)(out[i] / 2);
        return;
    }
    { /* hybrid transforms of vp9_dct.c: int16 1-D kernels, tx_type bit0 -> ADST on columns, bit1 -> rows */
        int16_t mid[16 * 16], ti[16], to[16];
        fwd1d dct = tx_size == 0 ? fdct4_1d : tx_size == 1 ? fdct8_1d : fdct16_1d;
        fwd1d adst = tx_size == 0 ? fadst4_1d : tx_size == 1 ? fadst8_1d : fadst16_1d;
        fwd1d fc = (tx_type & 1) ? adst : dct, fr = (tx_type & 2) ? adst : dct;
        for (int c = 0; c < n; c++) {
            for (int r = 0; r < n; r++) ti[r] = (int16_t)(in[r * stride + c] * (tx_size == 0 ? 16 : 4));
            if (tx_size == 0 && c == 0 && ti[0]) ti[0] += 1;
            run1d_i16(fc, n, ti, to);
            for (int r = 0; r < n; r++) mid[r * n + c] = tx_size == 2 ? (int16_t)((to[r] + 1 + (to[r] < 0)) >> 2) : to[r];
        }
        for (int r = 0; r < n; r++) {
            run1d_i16(fr, n, mid + r * n, to);
            for (int c = 0; c < n; c++)
                out[r * n + c] = tx_size == 0 ? (int16_t)((to[c] + 1) >> 2) : tx_size == 1 ? (int16_t)((to[c] + (to[c] < 0)) >> 1) : to[c];
        }
    }
}

/* =============================== quantiser =================================================== */
void svtb200_oracle_quantize(int tx_size, const int16_t *coeff, const SvtB200QuantRow *q, const int16_t *scan,
                             int16_t *qcoeff, int16_t *dqcoeff, uint16_t *eob_out) {
    const int n = (4 << tx_size) * (4 << tx_size);
    int eob = -1;
    memset(qcoeff, 0, n * sizeof(int16_t));
    memset(dqcoeff, 0, n * sizeof(int16_t));
    if (tx_size < 3) {
        int nz = n;
        for (int i = n - 1; i >= 0; i--) { /* trailing coefficients inside the dead zone are dropped */
            const int rc = scan[i], c = coeff[rc], zb = q->zbin[rc != 0];
            if (c < zb && c > -zb) nz--; else break;
        }
        for (int i = 0; i < nz; i++) {
            const int rc = scan[i], c = coeff[rc], sign = c >> 31, ac = (c ^ sign) - sign, k = rc != 0;
            if (ac >= q->zbin[k]) {
                int t = ac + q->round[k];
                t = t < -32768 ? -32768 : t > 32767 ? 32767 : t;
                t = ((((t * q->quant[k]) >> 16) + t) * q->quant_shift[k]) >> 16;
                qcoeff[rc]  = (int16_t)((t ^ sign) - sign);
                dqcoeff[rc] = (int16_t)(qcoeff[rc] * q->dequant[k]);
                if (t) eob = i;
            }
        }
    } else {
        for (int i = 0; i < n; i++) {
            const int rc = scan[i], c = coeff[rc], k = rc != 0;
            const int zb = (q->zbin[k] + 1) >> 1;
            if (c >= zb || c <= -zb) {
                const int sign = c >> 31;
                int ac = (c ^ sign) - sign;
                ac += (q->round[k] + 1) >> 1;
                ac = ac < -32768 ? -32768 : ac > 32767 ? 32767 : ac;
                const int t = ((((ac * q->quant[k]) >> 16) + ac) * q->quant_shift[k]) >> 15;
                qcoeff[rc]  = (int16_t)((t ^ sign) - sign);
                dqcoeff[rc] = (int16_t)(qcoeff[rc] * q->dequant[k] / 2);
                if (t) eob = i;
            }
        }
    }
    *eob_out = (uint16_t)(eob + 1);
}

/* eb_vp9_init_quantizer for one luma qindex given the DC/AC step sizes (dc_quant / ac_quant lookups are
 * bitstream constants supplied by the caller) -- vp9_quantize.c:206 with sharpness 0, 8-bit */
void svtb200_oracle_quant_row(int qindex, int dc_q, int ac_q, SvtB200QuantRow *row) {
    for (int i = 0; i < 2; i++) {
        const int quant = i == 0 ? dc_q : ac_q;
        const int qzbin_factor = qindex == 0 ? 64 : (quant < 148 ? 84 : 80), qround = qindex == 0 ? 64 : 48;
        /* invert_quant: quant = (1<<16)/d split into a 16-bit mantissa and a shift */
        unsigned t = (unsigned)quant;
        int l = 0;
        while (t > 1) { t >>= 1; l++; }
        const int m = 1 + (1 << (16 + l)) / quant;
        row->quant[i]       = (int16_t)(m - (1 << 16));
        row->quant_shift[i] = (int16_t)(1 << (16 - l));
        row->zbin[i]        = (int16_t)((qzbin_factor * quant + 64) >> 7);
        row->round[i]       = (int16_t)((qround * quant) >> 7);
        row->dequant[i]     = (int16_t)quant;
    }
}

/* =============================== inverse transforms ========================================== */
#define RW(v) w16(rnd14(v))
/* rotation pair used all over the inverse networks: lo = R(a*c1 - b*c2), hi = R(a*c2 + b*c1) */
#define ROT(a, b, c1, c2, lo, hi) do { const int ra_ = (a), rb_ = (b); (lo) = RW(ra_ * (c1) - rb_ * (c2)); (hi) = RW(ra_ * (c2) + rb_ * (c1)); } while (0)

static void idct4_1d(const int16_t *in, int16_t *out) {
    const int s0 = RW((in[0] + in[2]) * C64[16]), s1 = RW((in[0] - in[2]) * C64[16]);
    int s2, s3;
    ROT(in[1], in[3], C64[24], C64[8], s2, s3);
    out[0] = (int16_t)(s0 + s3); out[1] = (int16_t)(s1 + s2); out[2] = (int16_t)(s1 - s2); out[3] = (int16_t)(s0 - s3);
}

static void idct8_odd(int i1, int i3, int i5, int i7, int *o /* [4]: o4..o7 */) {
    int a4, a7, a5, a6;
    ROT(i1, i7, C64[28], C64[4], a4, a7);
    ROT(i5, i3, C64[12], C64[20], a5, a6);
    const int b4 = w16(a4 + a5), b5 = w16(a4 - a5), b6 = w16(-a6 + a7), b7 = w16(a6 + a7);
    o[0] = b4; o[1] = RW((b6 - b5) * C64[16]); o[2] = RW((b5 + b6) * C64[16]); o[3] = b7;
}

static void idct8_1d(const int16_t *in, int16_t *out) {
    int16_t ein[4] = {in[0], in[2], in[4], in[6]}, e[4];
    int o[4];
    idct4_1d(ein, e);
    idct8_odd(in[1], in[3], in[5], in[7], o);
    for (int i = 0; i < 4; i++) { out[i] = (int16_t)(e[i] + o[3 - i]); out[7 - i] = (int16_t)(e[i] - o[3 - i]); }
}

static void idct16_odd(const int16_t *in /* full 16-point input */, int *o /* [8]: o8..o15 */) {
    int s[16];
    ROT(in[1], in[15], C64[30], C64[2], s[8], s[15]);
    ROT(in[9], in[7], C64[14], C64[18], s[9], s[14]);
    ROT(in[5], in[11], C64[22], C64[10], s[10], s[13]);
    ROT(in[13], in[3], C64[6], C64[26], s[11], s[12]);
    int t[16];
    t[8] = w16(s[8] + s[9]); t[9] = w16(s[8] - s[9]); t[10] = w16(-s[10] + s[11]); t[11] = w16(s[10] + s[11]);
    t[12] = w16(s[12] + s[13]); t[13] = w16(s[12] - s[13]); t[14] = w16(-s[14] + s[15]); t[15] = w16(s[14] + s[15]);
    s[8] = t[8]; s[15] = t[15]; s[11] = t[11]; s[12] = t[12];
    s[9] = RW(-t[9] * C64[8] + t[14] * C64[24]);   s[14] = RW(t[9] * C64[24] + t[14] * C64[8]);
    s[10] = RW(-t[10] * C64[24] - t[13] * C64[8]); s[13] = RW(-t[10] * C64[8] + t[13] * C64[24]);
    t[8] = w16(s[8] + s[11]); t[9] = w16(s[9] + s[10]); t[10] = w16(s[9] - s[10]); t[11] = w16(s[8] - s[11]);
    t[12] = w16(-s[12] + s[15]); t[13] = w16(-s[13] + s[14]); t[14] = w16(s[13] + s[14]); t[15] = w16(s[12] + s[15]);
    o[0] = t[8]; o[1] = t[9]; o[6] = t[14]; o[7] = t[15];
    o[2] = RW((-t[10] + t[13]) * C64[16]); o[5] = RW((t[10] + t[13]) * C64[16]);
    o[3] = RW((-t[11] + t[12]) * C64[16]); o[4] = RW((t[11] + t[12]) * C64[16]);
}

static void idct16_1d(const int16_t *in, int16_t *out) {
    int16_t ein[8], e[8];
    int o[8];
    for (int i = 0; i < 8; i++) ein[i] = in[2 * i];
    idct8_1d(ein, e);
    idct16_odd(in, o);
    for (int i = 0; i < 8; i++) { out[i] = (int16_t)(e[i] + o[7 - i]); out[15 - i] = (int16_t)(e[i] - o[7 - i]); }
}

static void idct32_odd(const int16_t *in, int *o /* [16]: o16..o31 */) {
    int s[32], t[32];
    ROT(in[1], in[31], C64[31], C64[1], s[16], s[31]);
    ROT(in[17], in[15], C64[15], C64[17], s[17], s[30]);
    ROT(in[9], in[23], C64[23], C64[9], s[18], s[29]);
    ROT(in[25], in[7], C64[7], C64[25], s[19], s[28]);
    ROT(in[5], in[27], C64[27], C64[5], s[20], s[27]);
    ROT(in[21], in[11], C64[11], C64[21], s[21], s[26]);
    ROT(in[13], in[19], C64[19], C64[13], s[22], s[25]);
    ROT(in[29], in[3], C64[3], C64[29], s[23], s[24]);
    for (int k = 16; k < 32; k += 4) { /* stage 2 */
        t[k] = w16(s[k] + s[k + 1]); t[k + 1] = w16(s[k] - s[k + 1]); t[k + 2] = w16(-s[k + 2] + s[k + 3]); t[k + 3] = w16(s[k + 2] + s[k + 3]);
    }
    /* stage 3 */
    for (int k = 16; k < 32; k++) s[k] = t[k];
    s[17] = RW(-t[17] * C64[4] + t[30] * C64[28]);  s[30] = RW(t[17] * C64[28] + t[30] * C64[4]);
    s[18] = RW(-t[18] * C64[28] - t[29] * C64[4]);  s[29] = RW(-t[18] * C64[4] + t[29] * C64[28]);
    s[21] = RW(-t[21] * C64[20] + t[26] * C64[12]); s[26] = RW(t[21] * C64[12] + t[26] * C64[20]);
    s[22] = RW(-t[22] * C64[12] - t[25] * C64[20]); s[25] = RW(-t[22] * C64[20] + t[25] * C64[12]);
    /* stage 4 */
    t[16] = w16(s[16] + s[19]); t[17] = w16(s[17] + s[18]); t[18] = w16(s[17] - s[18]); t[19] = w16(s[16] - s[19]);
    t[20] = w16(-s[20] + s[23]); t[21] = w16(-s[21] + s[22]); t[22] = w16(s[21] + s[22]); t[23] = w16(s[20] + s[23]);
    t[24] = w16(s[24] + s[27]); t[25] = w16(s[25] + s[26]); t[26] = w16(s[25] - s[26]); t[27] = w16(s[24] - s[27]);
    t[28] = w16(-s[28] + s[31]); t[29] = w16(-s[29] + s[30]); t[30] = w16(s[29] + s[30]); t[31] = w16(s[28] + s[31]);
    /* stage 5 */
    for (int k = 16; k < 32; k++) s[k] = t[k];
    s[18] = RW(-t[18] * C64[8] + t[29] * C64[24]);  s[29] = RW(t[18] * C64[24] + t[29] * C64[8]);
    s[19] = RW(-t[19] * C64[8] + t[28] * C64[24]);  s[28] = RW(t[19] * C64[24] + t[28] * C64[8]);
    s[20] = RW(-t[20] * C64[24] - t[27] * C64[8]);  s[27] = RW(-t[20] * C64[8] + t[27] * C64[24]);
    s[21] = RW(-t[21] * C64[24] - t[26] * C64[8]);  s[26] = RW(-t[21] * C64[8] + t[26] * C64[24]);
    /* stage 6 */
    for (int i = 0; i < 4; i++) {
        t[16 + i] = w16(s[16 + i] + s[23 - i]); t[23 - i] = w16(s[16 + i] - s[23 - i]);
        t[24 + i] = w16(-s[24 + i] + s[31 - i]); t[31 - i] = w16(s[24 + i] + s[31 - i]);
    }
    /* stage 7 */
    for (int k = 16; k < 32; k++) o[k - 16] = t[k];
    for (int i = 0; i < 4; i++) {
        o[4 + i]  = RW((-t[20 + i] + t[27 - i]) * C64[16]);
        o[11 - i] = RW((t[20 + i] + t[27 - i]) * C64[16]);
    }
}

static void idct32_1d(const int16_t *in, int16_t *out) {
    int16_t ein[16], e[16];
    int o[16];
    for (int i = 0; i < 16; i++) ein[i] = in[2 * i];
    idct16_1d(ein, e);
    idct32_odd(in, o);
    for (int i = 0; i < 16; i++) { out[i] = (int16_t)(e[i] + o[15 - i]); out[31 - i] = (int16_t)(e[i] - o[15 - i]); }
}

/* inverse ADSTs keep 32-bit intermediates (no int16 wrap inside), only the outputs are truncated */
static void iadst4_1d(const int16_t *in, int16_t *out) {
    const int x0 = in[0], x1 = in[1], x2 = in[2], x3 = in[3];
    if (!(x0 | x1 | x2 | x3)) { memset(out, 0, 4 * sizeof(*out)); return; }
    int s0 = S9[1] * x0, s1 = S9[2] * x0, s2 = S9[3] * x1, s3 = S9[4] * x2, s4 = S9[1] * x2, s5 = S9[2] * x3, s6 = S9[4] * x3;
    const int s7 = x0 - x2 + x3;
    s0 = s0 + s3 + s5; s1 = s1 - s4 - s6; s3 = s2; s2 = S9[3] * s7;
    out[0] = (int16_t)rnd14(s0 + s3); out[1] = (int16_t)rnd14(s1 + s3); out[2] = (int16_t)rnd14(s2); out[3] = (int16_t)rnd14(s0 + s1 - s3);
}

static void iadst8_1d(const int16_t *in, int16_t *out) {
    int x0 = in[7], x1 = in[0], x2 = in[5], x3 = in[2], x4 = in[3], x5 = in[4], x6 = in[1], x7 = in[6];
    if (!(x0 | x1 | x2 | x3 | x4 | x5 | x6 | x7)) { memset(out, 0, 8 * sizeof(*out)); return; }
    int s0 = C64[2] * x0 + C64[30] * x1, s1 = C64[30] * x0 - C64[2] * x1;
    int s2 = C64[10] * x2 + C64[22] * x3, s3 = C64[22] * x2 - C64[10] * x3;
    int s4 = C64[18] * x4 + C64[14] * x5, s5 = C64[14] * x4 - C64[18] * x5;
    int s6 = C64[26] * x6 + C64[6] * x7, s7 = C64[6] * x6 - C64[26] * x7;
    x0 = rnd14(s0 + s4); x1 = rnd14(s1 + s5); x2 = rnd14(s2 + s6); x3 = rnd14(s3 + s7);
    x4 = rnd14(s0 - s4); x5 = rnd14(s1 - s5); x6 = rnd14(s2 - s6); x7 = rnd14(s3 - s7);
    s0 = x0; s1 = x1; s2 = x2; s3 = x3;
    s4 = C64[8] * x4 + C64[24] * x5; s5 = C64[24] * x4 - C64[8] * x5;
    s6 = -C64[24] * x6 + C64[8] * x7; s7 = C64[8] * x6 + C64[24] * x7;
    x0 = s0 + s2; x1 = s1 + s3; x2 = s0 - s2; x3 = s1 - s3;
    x4 = rnd14(s4 + s6); x5 = rnd14(s5 + s7); x6 = rnd14(s4 - s6); x7 = rnd14(s5 - s7);
    s2 = C64[16] * (x2 + x3); s3 = C64[16] * (x2 - x3); s6 = C64[16] * (x6 + x7); s7 = C64[16] * (x6 - x7);
    x2 = rnd14(s2); x3 = rnd14(s3); x6 = rnd14(s6); x7 = rnd14(s7);
    out[0] = (int16_t)x0; out[1] = (int16_t)-x4; out[2] = (int16_t)x6; out[3] = (int16_t)-x2;
    out[4] = (int16_t)x3; out[5] = (int16_t)-x7; out[6] = (int16_t)x5; out[7] = (int16_t)-x1;
}

static void iadst16_1d(const int16_t *in, int16_t *out) {
    int a[16], b[16], any = 0;
    for (int i = 0; i < 16; i++) { a[i] = in[i]; any |= a[i]; }
    if (!any) { memset(out, 0, 16 * sizeof(*out)); return; }
    /* the network is the forward ADST16's (same rotations, same rounding points) on 32-bit values */
    fadst16_1d(a, b);
    for (int i = 0; i < 16; i++) out[i] = (int16_t)b[i];
}

typedef void (*inv1d)(const int16_t *, int16_t *);

static inline uint8_t clip_add(uint8_t d, int v) { v = d + (int16_t)v; return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }

/* dest (stride) += inverse(coeffs); rows first, then columns; eob only selects the reference's
 * reduced variants, which equal the full transform whenever eob is consistent with the data */
void svtb200_oracle_inv_txfm_add(int tx_size, int tx_type, const int16_t *in, int eob, uint8_t *dest, int stride) {
    const int n = 4 << tx_size, shift = tx_size == 0 ? 4 : tx_size == 1 ? 5 : 6;
    int16_t mid[32 * 32], ti[32], to[32];
    (void)eob;
    inv1d dct = tx_size == 0 ? idct4_1d : tx_size == 1 ? idct8_1d : tx_size == 2 ? idct16_1d : idct32_1d;
    inv1d adst = tx_size == 0 ? iadst4_1d : tx_size == 1 ? iadst8_1d : iadst16_1d;
    /* IHT tables: tx_type 1 (ADST_DCT) = ADST on columns, DCT on rows; 2 = DCT cols, ADST rows */
    inv1d frow = (tx_size < 3 && (tx_type & 2)) ? adst : dct, fcol = (tx_size < 3 && (tx_type & 1)) ? adst : dct;
    for (int r = 0; r < n; r++) {
        if (tx_size == 3) { /* the 32x32 row pass skips all-zero rows (output is zero either way) */
            int nz = 0;
            for (int c = 0; c < 32; c++) nz |= in[r * 32 + c];
            if (!nz) { memset(mid + r * 32, 0, 32 * sizeof(int16_t)); continue; }
        }
        frow(in + r * n, mid + r * n);
    }
    for (int c = 0; c < n; c++) {
        for (int r = 0; r < n; r++) ti[r] = mid[r * n + c];
        fcol(ti, to);
        for (int r = 0; r < n; r++) dest[r * stride + c] = clip_add(dest[r * stride + c], (to[r] + (1 << (shift - 1))) >> shift);
    }
}
