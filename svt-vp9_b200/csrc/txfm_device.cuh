/*
 * txfm_device.cuh -- register-resident 1-D integer transforms of the VP9 transform set, as used by the reference's
 * eb_vp9_fdct* / fht* / idct* / iht* kernels (Source/Lib/VPX/fwd_txfm.c, vp9_dct.c, inv_txfm.c).
 *
 * Each routine transforms one length-N vector held in registers by ONE thread (fully unrolled, compile-time
 * indices only); the 2-D kernels in txq.cu give every thread one column, transpose through shared memory and
 * give every thread one row.  The networks are organised recursively (N-point = N/2-point of the folded input +
 * an N-specific rotation network); rounding (+8192 >> 14 after every multiply stage), the int16 wrap of the
 * inverse DCT and the 32-bit arithmetic of the ADSTs follow the reference bit for bit -- they are what makes
 * these transforms non-linear and rule out a GEMM formulation.
 */
#ifndef SVTB200_TXFM_DEVICE_CUH
#define SVTB200_TXFM_DEVICE_CUH
#include <stdint.h>

#define TXD __device__ __forceinline__

/* cospi_k_64 = round(16384 cos(k pi / 64)) as compile-time constants (txfm_common.h:17-64) */
template <int K> struct Cos64;
#define SVTB200_C64(k, v) template <> struct Cos64<k> { static constexpr int v_ = v; };
SVTB200_C64(1, 16364) SVTB200_C64(2, 16305) SVTB200_C64(3, 16207) SVTB200_C64(4, 16069) SVTB200_C64(5, 15893)
SVTB200_C64(6, 15679) SVTB200_C64(7, 15426) SVTB200_C64(8, 15137) SVTB200_C64(9, 14811) SVTB200_C64(10, 14449)
SVTB200_C64(11, 14053) SVTB200_C64(12, 13623) SVTB200_C64(13, 13160) SVTB200_C64(14, 12665) SVTB200_C64(15, 12140)
SVTB200_C64(16, 11585) SVTB200_C64(17, 11003) SVTB200_C64(18, 10394) SVTB200_C64(19, 9760) SVTB200_C64(20, 9102)
SVTB200_C64(21, 8423) SVTB200_C64(22, 7723) SVTB200_C64(23, 7005) SVTB200_C64(24, 6270) SVTB200_C64(25, 5520)
SVTB200_C64(26, 4756) SVTB200_C64(27, 3981) SVTB200_C64(28, 3196) SVTB200_C64(29, 2404) SVTB200_C64(30, 1606)
SVTB200_C64(31, 804)
#define C(k) (Cos64<k>::v_)
#define S9_1 5283
#define S9_2 9929
#define S9_3 13377
#define S9_4 15212

TXD int rnd14(int v) { return (v + 8192) >> 14; }
TXD int w16(int v) { return (int)(short)v; }

/* ================================ forward DCT ================================ */
TXD void fdct4_1d(const int (&in)[4], int (&out)[4]) {
    const int s0 = in[0] + in[3], s1 = in[1] + in[2], s2 = in[1] - in[2], s3 = in[0] - in[3];
    out[0] = rnd14((s0 + s1) * C(16));
    out[2] = rnd14((s0 - s1) * C(16));
    out[1] = rnd14(s2 * C(24) + s3 * C(8));
    out[3] = rnd14(-s2 * C(8) + s3 * C(24));
}

TXD void fdct8_1d(const int (&in)[8], int (&out)[8]) {
    int e[4], ev[4];
#pragma unroll
    for (int i = 0; i < 4; i++) e[i] = in[i] + in[7 - i];
    const int d0 = in[3] - in[4], d1 = in[2] - in[5], d2 = in[1] - in[6], d3 = in[0] - in[7];
    fdct4_1d(e, ev);
    out[0] = ev[0]; out[2] = ev[1]; out[4] = ev[2]; out[6] = ev[3];
    const int t2 = rnd14((d2 - d1) * C(16)), t3 = rnd14((d2 + d1) * C(16));
    const int x0 = d0 + t2, x1 = d0 - t2, x2 = d3 - t3, x3 = d3 + t3;
    out[1] = rnd14(x0 * C(28) + x3 * C(4));
    out[5] = rnd14(x1 * C(12) + x2 * C(20));
    out[3] = rnd14(x2 * C(12) + x1 * -C(20));
    out[7] = rnd14(x3 * C(28) + x0 * -C(4));
}

TXD void fdct16_1d(const int (&in)[16], int (&out)[16]) {
    int e[8], s[8], ev[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { e[i] = in[i] + in[15 - i]; s[i] = in[7 - i] - in[8 + i]; }
    fdct8_1d(e, ev);
#pragma unroll
    for (int i = 0; i < 8; i++) out[2 * i] = ev[i];
    const int a2 = rnd14((s[5] - s[2]) * C(16)), a3 = rnd14((s[4] - s[3]) * C(16));
    const int a4 = rnd14((s[4] + s[3]) * C(16)), a5 = rnd14((s[5] + s[2]) * C(16));
    const int b0 = s[0] + a3, b1 = s[1] + a2, b2 = s[1] - a2, b3 = s[0] - a3;
    const int b4 = s[7] - a4, b5 = s[6] - a5, b6 = s[6] + a5, b7 = s[7] + a4;
    const int c1 = rnd14(b1 * -C(8) + b6 * C(24)), c2 = rnd14(b2 * C(24) + b5 * C(8));
    const int c5 = rnd14(b2 * C(8) - b5 * C(24)), c6 = rnd14(b1 * C(24) + b6 * C(8));
    const int d0 = b0 + c1, d1 = b0 - c1, d2 = b3 + c2, d3 = b3 - c2;
    const int d4 = b4 - c5, d5 = b4 + c5, d6 = b7 - c6, d7 = b7 + c6;
    out[1]  = rnd14(d0 * C(30) + d7 * C(2));
    out[9]  = rnd14(d1 * C(14) + d6 * C(18));
    out[5]  = rnd14(d2 * C(22) + d5 * C(10));
    out[13] = rnd14(d3 * C(6) + d4 * C(26));
    out[3]  = rnd14(d3 * -C(26) + d4 * C(6));
    out[11] = rnd14(d2 * -C(10) + d5 * C(22));
    out[7]  = rnd14(d1 * -C(18) + d6 * C(14));
    out[15] = rnd14(d0 * -C(2) + d7 * C(30));
}

/* eb_vp9_fdct32 (round = 0): every multiply stage rounded, additions exact */
TXD void fdct32_1d(const int (&in)[32], int (&out)[32]) {
    int a[32], b[32];
#pragma unroll
    for (int i = 0; i < 16; i++) { a[i] = in[i] + in[31 - i]; a[16 + i] = in[15 - i] - in[16 + i]; }
#pragma unroll
    for (int i = 0; i < 8; i++) { b[i] = a[i] + a[15 - i]; b[8 + i] = a[7 - i] - a[8 + i]; }
    b[16] = a[16]; b[17] = a[17]; b[18] = a[18]; b[19] = a[19];
#pragma unroll
    for (int i = 0; i < 4; i++) { b[20 + i] = rnd14((a[27 - i] - a[20 + i]) * C(16)); b[27 - i] = rnd14((a[27 - i] + a[20 + i]) * C(16)); }
    b[28] = a[28]; b[29] = a[29]; b[30] = a[30]; b[31] = a[31];
#pragma unroll
    for (int i = 0; i < 4; i++) { a[i] = b[i] + b[7 - i]; a[4 + i] = b[3 - i] - b[4 + i]; }
    a[8] = b[8]; a[9] = b[9];
    a[10] = rnd14((b[13] - b[10]) * C(16)); a[13] = rnd14((b[13] + b[10]) * C(16));
    a[11] = rnd14((b[12] - b[11]) * C(16)); a[12] = rnd14((b[12] + b[11]) * C(16));
    a[14] = b[14]; a[15] = b[15];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        a[16 + i] = b[16 + i] + b[23 - i]; a[20 + i] = b[19 - i] - b[20 + i];
        a[24 + i] = b[31 - i] - b[24 + i]; a[28 + i] = b[28 + i] + b[27 - i];
    }
    b[0] = a[0] + a[3]; b[1] = a[1] + a[2]; b[2] = a[1] - a[2]; b[3] = a[0] - a[3];
    b[4] = a[4]; b[5] = rnd14((a[6] - a[5]) * C(16)); b[6] = rnd14((a[6] + a[5]) * C(16)); b[7] = a[7];
    b[8] = a[8] + a[11]; b[9] = a[9] + a[10]; b[10] = a[9] - a[10]; b[11] = a[8] - a[11];
    b[12] = a[15] - a[12]; b[13] = a[14] - a[13]; b[14] = a[14] + a[13]; b[15] = a[15] + a[12];
    b[16] = a[16]; b[17] = a[17];
    b[18] = rnd14(a[18] * -C(8) + a[29] * C(24)); b[19] = rnd14(a[19] * -C(8) + a[28] * C(24));
    b[20] = rnd14(a[20] * -C(24) + a[27] * -C(8)); b[21] = rnd14(a[21] * -C(24) + a[26] * -C(8));
    b[22] = a[22]; b[23] = a[23]; b[24] = a[24]; b[25] = a[25];
    b[26] = rnd14(a[26] * C(24) + a[21] * -C(8)); b[27] = rnd14(a[27] * C(24) + a[20] * -C(8));
    b[28] = rnd14(a[28] * C(8) + a[19] * C(24)); b[29] = rnd14(a[29] * C(8) + a[18] * C(24));
    b[30] = a[30]; b[31] = a[31];
    a[0] = rnd14((b[0] + b[1]) * C(16)); a[1] = rnd14((b[0] - b[1]) * C(16));
    a[2] = rnd14(b[2] * C(24) + b[3] * C(8)); a[3] = rnd14(b[3] * C(24) - b[2] * C(8));
    a[4] = b[4] + b[5]; a[5] = b[4] - b[5]; a[6] = b[7] - b[6]; a[7] = b[7] + b[6];
    a[8] = b[8];
    a[9] = rnd14(b[9] * -C(8) + b[14] * C(24)); a[10] = rnd14(b[10] * -C(24) + b[13] * -C(8));
    a[11] = b[11]; a[12] = b[12];
    a[13] = rnd14(b[13] * C(24) + b[10] * -C(8)); a[14] = rnd14(b[14] * C(8) + b[9] * C(24));
    a[15] = b[15];
    a[16] = b[16] + b[19]; a[17] = b[17] + b[18]; a[18] = b[17] - b[18]; a[19] = b[16] - b[19];
    a[20] = b[23] - b[20]; a[21] = b[22] - b[21]; a[22] = b[22] + b[21]; a[23] = b[23] + b[20];
    a[24] = b[24] + b[27]; a[25] = b[25] + b[26]; a[26] = b[25] - b[26]; a[27] = b[24] - b[27];
    a[28] = b[31] - b[28]; a[29] = b[30] - b[29]; a[30] = b[30] + b[29]; a[31] = b[31] + b[28];
    b[0] = a[0]; b[1] = a[1]; b[2] = a[2]; b[3] = a[3];
    b[4] = rnd14(a[4] * C(28) + a[7] * C(4)); b[5] = rnd14(a[5] * C(12) + a[6] * C(20));
    b[6] = rnd14(a[6] * C(12) + a[5] * -C(20)); b[7] = rnd14(a[7] * C(28) + a[4] * -C(4));
    b[8] = a[8] + a[9]; b[9] = a[8] - a[9]; b[10] = a[11] - a[10]; b[11] = a[11] + a[10];
    b[12] = a[12] + a[13]; b[13] = a[12] - a[13]; b[14] = a[15] - a[14]; b[15] = a[15] + a[14];
    b[16] = a[16];
    b[17] = rnd14(a[17] * -C(4) + a[30] * C(28)); b[18] = rnd14(a[18] * -C(28) + a[29] * -C(4));
    b[19] = a[19]; b[20] = a[20];
    b[21] = rnd14(a[21] * -C(20) + a[26] * C(12)); b[22] = rnd14(a[22] * -C(12) + a[25] * -C(20));
    b[23] = a[23]; b[24] = a[24];
    b[25] = rnd14(a[25] * C(12) + a[22] * -C(20)); b[26] = rnd14(a[26] * C(20) + a[21] * C(12));
    b[27] = a[27]; b[28] = a[28];
    b[29] = rnd14(a[29] * C(28) + a[18] * -C(4)); b[30] = rnd14(a[30] * C(4) + a[17] * C(28));
    b[31] = a[31];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = b[i];
    a[8]  = rnd14(b[8] * C(30) + b[15] * C(2));   a[9]  = rnd14(b[9] * C(14) + b[14] * C(18));
    a[10] = rnd14(b[10] * C(22) + b[13] * C(10)); a[11] = rnd14(b[11] * C(6) + b[12] * C(26));
    a[12] = rnd14(b[12] * C(6) + b[11] * -C(26)); a[13] = rnd14(b[13] * C(22) + b[10] * -C(10));
    a[14] = rnd14(b[14] * C(14) + b[9] * -C(18)); a[15] = rnd14(b[15] * C(30) + b[8] * -C(2));
    a[16] = b[16] + b[17]; a[17] = b[16] - b[17]; a[18] = b[19] - b[18]; a[19] = b[19] + b[18];
    a[20] = b[20] + b[21]; a[21] = b[20] - b[21]; a[22] = b[23] - b[22]; a[23] = b[23] + b[22];
    a[24] = b[24] + b[25]; a[25] = b[24] - b[25]; a[26] = b[27] - b[26]; a[27] = b[27] + b[26];
    a[28] = b[28] + b[29]; a[29] = b[28] - b[29]; a[30] = b[31] - b[30]; a[31] = b[31] + b[30];
    out[0] = a[0]; out[16] = a[1]; out[8] = a[2]; out[24] = a[3]; out[4] = a[4]; out[20] = a[5]; out[12] = a[6]; out[28] = a[7];
    out[2] = a[8]; out[18] = a[9]; out[10] = a[10]; out[26] = a[11]; out[6] = a[12]; out[22] = a[13]; out[14] = a[14]; out[30] = a[15];
    out[1]  = rnd14(a[16] * C(31) + a[31] * C(1));   out[17] = rnd14(a[17] * C(15) + a[30] * C(17));
    out[9]  = rnd14(a[18] * C(23) + a[29] * C(9));   out[25] = rnd14(a[19] * C(7) + a[28] * C(25));
    out[5]  = rnd14(a[20] * C(27) + a[27] * C(5));   out[21] = rnd14(a[21] * C(11) + a[26] * C(21));
    out[13] = rnd14(a[22] * C(19) + a[25] * C(13));  out[29] = rnd14(a[23] * C(3) + a[24] * C(29));
    out[3]  = rnd14(a[24] * C(3) + a[23] * -C(29));  out[19] = rnd14(a[25] * C(19) + a[22] * -C(13));
    out[11] = rnd14(a[26] * C(11) + a[21] * -C(21)); out[27] = rnd14(a[27] * C(27) + a[20] * -C(5));
    out[7]  = rnd14(a[28] * C(7) + a[19] * -C(25));  out[23] = rnd14(a[29] * C(23) + a[18] * -C(9));
    out[15] = rnd14(a[30] * C(15) + a[17] * -C(17)); out[31] = rnd14(a[31] * C(31) + a[16] * -C(1));
}

/* ================================ ADST (forward networks; the inverse 16-point shares it) ==== */
TXD void fadst4_1d(const int (&in)[4], int (&out)[4]) {
    int x0 = in[0], x1 = in[1], x2 = in[2], x3 = in[3];
    if (!(x0 | x1 | x2 | x3)) { out[0] = out[1] = out[2] = out[3] = 0; return; }
    const int s0 = S9_1 * x0, s1 = S9_4 * x0, s2 = S9_2 * x1, s3 = S9_1 * x1, s4 = S9_3 * x2, s5 = S9_4 * x3, s6 = S9_2 * x3;
    const int s7 = x0 + x1 - x3;
    x0 = s0 + s2 + s5; x1 = S9_3 * s7; x2 = s1 - s3 + s6; x3 = s4;
    out[0] = rnd14(x0 + x3); out[1] = rnd14(x1); out[2] = rnd14(x2 - x3); out[3] = rnd14(x2 - x0 + x3);
}

TXD void adst8_core(const int (&in)[8], int (&out)[8]) {
    int x0 = in[7], x1 = in[0], x2 = in[5], x3 = in[2], x4 = in[3], x5 = in[4], x6 = in[1], x7 = in[6];
    int s0 = C(2) * x0 + C(30) * x1, s1 = C(30) * x0 - C(2) * x1;
    int s2 = C(10) * x2 + C(22) * x3, s3 = C(22) * x2 - C(10) * x3;
    int s4 = C(18) * x4 + C(14) * x5, s5 = C(14) * x4 - C(18) * x5;
    int s6 = C(26) * x6 + C(6) * x7, s7 = C(6) * x6 - C(26) * x7;
    x0 = rnd14(s0 + s4); x1 = rnd14(s1 + s5); x2 = rnd14(s2 + s6); x3 = rnd14(s3 + s7);
    x4 = rnd14(s0 - s4); x5 = rnd14(s1 - s5); x6 = rnd14(s2 - s6); x7 = rnd14(s3 - s7);
    s0 = x0; s1 = x1; s2 = x2; s3 = x3;
    s4 = C(8) * x4 + C(24) * x5; s5 = C(24) * x4 - C(8) * x5;
    s6 = -C(24) * x6 + C(8) * x7; s7 = C(8) * x6 + C(24) * x7;
    x0 = s0 + s2; x1 = s1 + s3; x2 = s0 - s2; x3 = s1 - s3;
    x4 = rnd14(s4 + s6); x5 = rnd14(s5 + s7); x6 = rnd14(s4 - s6); x7 = rnd14(s5 - s7);
    s2 = C(16) * (x2 + x3); s3 = C(16) * (x2 - x3); s6 = C(16) * (x6 + x7); s7 = C(16) * (x6 - x7);
    x2 = rnd14(s2); x3 = rnd14(s3); x6 = rnd14(s6); x7 = rnd14(s7);
    out[0] = x0; out[1] = -x4; out[2] = x6; out[3] = -x2; out[4] = x3; out[5] = -x7; out[6] = x5; out[7] = -x1;
}

TXD void adst16_core(const int (&in)[16], int (&out)[16]) {
    int x[16], s[16];
    x[0] = in[15]; x[1] = in[0]; x[2] = in[13]; x[3] = in[2]; x[4] = in[11]; x[5] = in[4]; x[6] = in[9]; x[7] = in[6];
    x[8] = in[7]; x[9] = in[8]; x[10] = in[5]; x[11] = in[10]; x[12] = in[3]; x[13] = in[12]; x[14] = in[1]; x[15] = in[14];
    s[0] = x[0] * C(1) + x[1] * C(31);    s[1] = x[0] * C(31) - x[1] * C(1);
    s[2] = x[2] * C(5) + x[3] * C(27);    s[3] = x[2] * C(27) - x[3] * C(5);
    s[4] = x[4] * C(9) + x[5] * C(23);    s[5] = x[4] * C(23) - x[5] * C(9);
    s[6] = x[6] * C(13) + x[7] * C(19);   s[7] = x[6] * C(19) - x[7] * C(13);
    s[8] = x[8] * C(17) + x[9] * C(15);   s[9] = x[8] * C(15) - x[9] * C(17);
    s[10] = x[10] * C(21) + x[11] * C(11); s[11] = x[10] * C(11) - x[11] * C(21);
    s[12] = x[12] * C(25) + x[13] * C(7);  s[13] = x[12] * C(7) - x[13] * C(25);
    s[14] = x[14] * C(29) + x[15] * C(3);  s[15] = x[14] * C(3) - x[15] * C(29);
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = rnd14(s[i] + s[i + 8]); x[i + 8] = rnd14(s[i] - s[i + 8]); }
#pragma unroll
    for (int i = 0; i < 8; i++) s[i] = x[i];
    s[8] = x[8] * C(4) + x[9] * C(28);     s[9] = x[8] * C(28) - x[9] * C(4);
    s[10] = x[10] * C(20) + x[11] * C(12);  s[11] = x[10] * C(12) - x[11] * C(20);
    s[12] = -x[12] * C(28) + x[13] * C(4);  s[13] = x[12] * C(4) + x[13] * C(28);
    s[14] = -x[14] * C(12) + x[15] * C(20); s[15] = x[14] * C(20) + x[15] * C(12);
#pragma unroll
    for (int i = 0; i < 4; i++) { x[i] = s[i] + s[i + 4]; x[i + 4] = s[i] - s[i + 4]; }
#pragma unroll
    for (int i = 8; i < 12; i++) { x[i] = rnd14(s[i] + s[i + 4]); x[i + 4] = rnd14(s[i] - s[i + 4]); }
#pragma unroll
    for (int h = 0; h < 16; h += 8) {
        s[h] = x[h]; s[h + 1] = x[h + 1]; s[h + 2] = x[h + 2]; s[h + 3] = x[h + 3];
        s[h + 4] = x[h + 4] * C(8) + x[h + 5] * C(24);  s[h + 5] = x[h + 4] * C(24) - x[h + 5] * C(8);
        s[h + 6] = -x[h + 6] * C(24) + x[h + 7] * C(8); s[h + 7] = x[h + 6] * C(8) + x[h + 7] * C(24);
        x[h] = s[h] + s[h + 2]; x[h + 1] = s[h + 1] + s[h + 3]; x[h + 2] = s[h] - s[h + 2]; x[h + 3] = s[h + 1] - s[h + 3];
        x[h + 4] = rnd14(s[h + 4] + s[h + 6]); x[h + 5] = rnd14(s[h + 5] + s[h + 7]);
        x[h + 6] = rnd14(s[h + 4] - s[h + 6]); x[h + 7] = rnd14(s[h + 5] - s[h + 7]);
    }
    s[2] = -C(16) * (x[2] + x[3]);   s[3] = C(16) * (x[2] - x[3]);
    s[6] = C(16) * (x[6] + x[7]);    s[7] = C(16) * (-x[6] + x[7]);
    s[10] = C(16) * (x[10] + x[11]); s[11] = C(16) * (-x[10] + x[11]);
    s[14] = -C(16) * (x[14] + x[15]); s[15] = C(16) * (x[14] - x[15]);
    x[2] = rnd14(s[2]); x[3] = rnd14(s[3]); x[6] = rnd14(s[6]); x[7] = rnd14(s[7]);
    x[10] = rnd14(s[10]); x[11] = rnd14(s[11]); x[14] = rnd14(s[14]); x[15] = rnd14(s[15]);
    out[0] = x[0]; out[1] = -x[8]; out[2] = x[12]; out[3] = -x[4]; out[4] = x[6]; out[5] = x[14]; out[6] = x[10]; out[7] = x[2];
    out[8] = x[3]; out[9] = x[11]; out[10] = x[15]; out[11] = x[7]; out[12] = x[5]; out[13] = -x[13]; out[14] = x[9]; out[15] = -x[1];
}

/* ================================ inverse DCT (int16 wrap after every operation) ============= */
#define RW(v) w16(rnd14(v))
#define IROT(a, b, c1, c2, lo, hi) { const int ra_ = (a), rb_ = (b); (lo) = RW(ra_ * (c1) - rb_ * (c2)); (hi) = RW(ra_ * (c2) + rb_ * (c1)); }

TXD void idct4_1d(const int (&in)[4], int (&out)[4]) {
    const int s0 = RW((in[0] + in[2]) * C(16)), s1 = RW((in[0] - in[2]) * C(16));
    int s2, s3;
    IROT(in[1], in[3], C(24), C(8), s2, s3)
    out[0] = w16(s0 + s3); out[1] = w16(s1 + s2); out[2] = w16(s1 - s2); out[3] = w16(s0 - s3);
}

TXD void idct8_1d(const int (&in)[8], int (&out)[8]) {
    const int ein[4] = {in[0], in[2], in[4], in[6]};
    int e[4], a4, a7, a5, a6;
    idct4_1d(ein, e);
    IROT(in[1], in[7], C(28), C(4), a4, a7)
    IROT(in[5], in[3], C(12), C(20), a5, a6)
    const int b4 = w16(a4 + a5), b5 = w16(a4 - a5), b6 = w16(-a6 + a7), b7 = w16(a6 + a7);
    const int o[4] = {b4, RW((b6 - b5) * C(16)), RW((b5 + b6) * C(16)), b7};
#pragma unroll
    for (int i = 0; i < 4; i++) { out[i] = w16(e[i] + o[3 - i]); out[7 - i] = w16(e[i] - o[3 - i]); }
}

TXD void idct16_1d(const int (&in)[16], int (&out)[16]) {
    int ein[8], e[8], s[16], t[16], o[8];
#pragma unroll
    for (int i = 0; i < 8; i++) ein[i] = in[2 * i];
    idct8_1d(ein, e);
    IROT(in[1], in[15], C(30), C(2), s[8], s[15])
    IROT(in[9], in[7], C(14), C(18), s[9], s[14])
    IROT(in[5], in[11], C(22), C(10), s[10], s[13])
    IROT(in[13], in[3], C(6), C(26), s[11], s[12])
    t[8] = w16(s[8] + s[9]); t[9] = w16(s[8] - s[9]); t[10] = w16(-s[10] + s[11]); t[11] = w16(s[10] + s[11]);
    t[12] = w16(s[12] + s[13]); t[13] = w16(s[12] - s[13]); t[14] = w16(-s[14] + s[15]); t[15] = w16(s[14] + s[15]);
    s[8] = t[8]; s[15] = t[15]; s[11] = t[11]; s[12] = t[12];
    s[9] = RW(-t[9] * C(8) + t[14] * C(24));   s[14] = RW(t[9] * C(24) + t[14] * C(8));
    s[10] = RW(-t[10] * C(24) - t[13] * C(8)); s[13] = RW(-t[10] * C(8) + t[13] * C(24));
    t[8] = w16(s[8] + s[11]); t[9] = w16(s[9] + s[10]); t[10] = w16(s[9] - s[10]); t[11] = w16(s[8] - s[11]);
    t[12] = w16(-s[12] + s[15]); t[13] = w16(-s[13] + s[14]); t[14] = w16(s[13] + s[14]); t[15] = w16(s[12] + s[15]);
    o[0] = t[8]; o[1] = t[9]; o[6] = t[14]; o[7] = t[15];
    o[2] = RW((-t[10] + t[13]) * C(16)); o[5] = RW((t[10] + t[13]) * C(16));
    o[3] = RW((-t[11] + t[12]) * C(16)); o[4] = RW((t[11] + t[12]) * C(16));
#pragma unroll
    for (int i = 0; i < 8; i++) { out[i] = w16(e[i] + o[7 - i]); out[15 - i] = w16(e[i] - o[7 - i]); }
}

TXD void idct32_1d(const int (&in)[32], int (&out)[32]) {
    int ein[16], e[16], s[32], t[32], o[16];
#pragma unroll
    for (int i = 0; i < 16; i++) ein[i] = in[2 * i];
    idct16_1d(ein, e);
    IROT(in[1], in[31], C(31), C(1), s[16], s[31])
    IROT(in[17], in[15], C(15), C(17), s[17], s[30])
    IROT(in[9], in[23], C(23), C(9), s[18], s[29])
    IROT(in[25], in[7], C(7), C(25), s[19], s[28])
    IROT(in[5], in[27], C(27), C(5), s[20], s[27])
    IROT(in[21], in[11], C(11), C(21), s[21], s[26])
    IROT(in[13], in[19], C(19), C(13), s[22], s[25])
    IROT(in[29], in[3], C(3), C(29), s[23], s[24])
#pragma unroll
    for (int k = 16; k < 32; k += 4) {
        t[k] = w16(s[k] + s[k + 1]); t[k + 1] = w16(s[k] - s[k + 1]); t[k + 2] = w16(-s[k + 2] + s[k + 3]); t[k + 3] = w16(s[k + 2] + s[k + 3]);
    }
#pragma unroll
    for (int k = 16; k < 32; k++) s[k] = t[k];
    s[17] = RW(-t[17] * C(4) + t[30] * C(28));  s[30] = RW(t[17] * C(28) + t[30] * C(4));
    s[18] = RW(-t[18] * C(28) - t[29] * C(4));  s[29] = RW(-t[18] * C(4) + t[29] * C(28));
    s[21] = RW(-t[21] * C(20) + t[26] * C(12)); s[26] = RW(t[21] * C(12) + t[26] * C(20));
    s[22] = RW(-t[22] * C(12) - t[25] * C(20)); s[25] = RW(-t[22] * C(20) + t[25] * C(12));
    t[16] = w16(s[16] + s[19]); t[17] = w16(s[17] + s[18]); t[18] = w16(s[17] - s[18]); t[19] = w16(s[16] - s[19]);
    t[20] = w16(-s[20] + s[23]); t[21] = w16(-s[21] + s[22]); t[22] = w16(s[21] + s[22]); t[23] = w16(s[20] + s[23]);
    t[24] = w16(s[24] + s[27]); t[25] = w16(s[25] + s[26]); t[26] = w16(s[25] - s[26]); t[27] = w16(s[24] - s[27]);
    t[28] = w16(-s[28] + s[31]); t[29] = w16(-s[29] + s[30]); t[30] = w16(s[29] + s[30]); t[31] = w16(s[28] + s[31]);
#pragma unroll
    for (int k = 16; k < 32; k++) s[k] = t[k];
    s[18] = RW(-t[18] * C(8) + t[29] * C(24));  s[29] = RW(t[18] * C(24) + t[29] * C(8));
    s[19] = RW(-t[19] * C(8) + t[28] * C(24));  s[28] = RW(t[19] * C(24) + t[28] * C(8));
    s[20] = RW(-t[20] * C(24) - t[27] * C(8));  s[27] = RW(-t[20] * C(8) + t[27] * C(24));
    s[21] = RW(-t[21] * C(24) - t[26] * C(8));  s[26] = RW(-t[21] * C(8) + t[26] * C(24));
#pragma unroll
    for (int i = 0; i < 4; i++) {
        t[16 + i] = w16(s[16 + i] + s[23 - i]); t[23 - i] = w16(s[16 + i] - s[23 - i]);
        t[24 + i] = w16(-s[24 + i] + s[31 - i]); t[31 - i] = w16(s[24 + i] + s[31 - i]);
    }
#pragma unroll
    for (int k = 16; k < 32; k++) o[k - 16] = t[k];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        o[4 + i]  = RW((-t[20 + i] + t[27 - i]) * C(16));
        o[11 - i] = RW((t[20 + i] + t[27 - i]) * C(16));
    }
#pragma unroll
    for (int i = 0; i < 16; i++) { out[i] = w16(e[i] + o[15 - i]); out[31 - i] = w16(e[i] - o[15 - i]); }
}

/* inverse ADSTs: 32-bit intermediates, outputs truncated to int16 by the caller */
TXD void iadst4_1d(const int (&in)[4], int (&out)[4]) {
    const int x0 = in[0], x1 = in[1], x2 = in[2], x3 = in[3];
    if (!(x0 | x1 | x2 | x3)) { out[0] = out[1] = out[2] = out[3] = 0; return; }
    int s0 = S9_1 * x0, s1 = S9_2 * x0, s2 = S9_3 * x1, s3 = S9_4 * x2;
    const int s4 = S9_1 * x2, s5 = S9_2 * x3, s6 = S9_4 * x3, s7 = x0 - x2 + x3;
    s0 = s0 + s3 + s5; s1 = s1 - s4 - s6; s3 = s2; s2 = S9_3 * s7;
    out[0] = w16(rnd14(s0 + s3)); out[1] = w16(rnd14(s1 + s3)); out[2] = w16(rnd14(s2)); out[3] = w16(rnd14(s0 + s1 - s3));
}

/* ---- dispatch helpers: N-point forward / inverse 1-D transform of the requested kind, int16 stores applied ---- */
template <int N> struct Tx1D;
template <> struct Tx1D<4> {
    TXD static void fwd(bool adst, const int (&in)[4], int (&out)[4]) { if (adst) fadst4_1d(in, out); else fdct4_1d(in, out); }
    TXD static void inv(bool adst, const int (&in)[4], int (&out)[4]) { if (adst) iadst4_1d(in, out); else idct4_1d(in, out); }
};
template <> struct Tx1D<8> {
    TXD static void fwd(bool adst, const int (&in)[8], int (&out)[8]) { if (adst) adst8_core(in, out); else fdct8_1d(in, out); }
    TXD static void inv(bool adst, const int (&in)[8], int (&out)[8]) {
        if (adst) {
            int any = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) any |= in[i];
            if (!any) {
#pragma unroll
                for (int i = 0; i < 8; i++) out[i] = 0;
                return;
            }
            adst8_core(in, out);
#pragma unroll
            for (int i = 0; i < 8; i++) out[i] = w16(out[i]);
        } else idct8_1d(in, out);
    }
};
template <> struct Tx1D<16> {
    TXD static void fwd(bool adst, const int (&in)[16], int (&out)[16]) { if (adst) adst16_core(in, out); else fdct16_1d(in, out); }
    TXD static void inv(bool adst, const int (&in)[16], int (&out)[16]) {
        if (adst) {
            adst16_core(in, out);
#pragma unroll
            for (int i = 0; i < 16; i++) out[i] = w16(out[i]);
        } else idct16_1d(in, out);
    }
};
template <> struct Tx1D<32> {
    TXD static void fwd(bool, const int (&in)[32], int (&out)[32]) { fdct32_1d(in, out); }
    TXD static void inv(bool, const int (&in)[32], int (&out)[32]) { idct32_1d(in, out); }
};

#endif
