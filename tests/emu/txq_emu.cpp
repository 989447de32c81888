// CPU execution of the transform / quantiser kernels' device code (test infrastructure only).
#define SVTB200_CPU_EMU 1
#include "../../svt-vp9_b200/csrc/emu/cuda_emu.h"
#include "../../include/svtb200.h"
#include "../../svt-vp9_b200/csrc/txq_kernel.cuh"

template <int MODE> static void run_mode(int tx_size, const TxqArgs &A) {
    const int N = 4 << tx_size, G = TXQ_THREADS / N;
    int grid = (A.n + G - 1) / G;
    if (grid > 6) grid = 6; /* grid-stride beyond, as on the device */
    if (grid < 1) grid = 1;
    switch (tx_size) {
    case 0: emu_launch(dim3((unsigned)grid), dim3(TXQ_THREADS), 0, [&] { txq_kernel<4, MODE>(A); }); break;
    case 1: emu_launch(dim3((unsigned)grid), dim3(TXQ_THREADS), 0, [&] { txq_kernel<8, MODE>(A); }); break;
    case 2: emu_launch(dim3((unsigned)grid), dim3(TXQ_THREADS), 0, [&] { txq_kernel<16, MODE>(A); }); break;
    default: emu_launch(dim3((unsigned)grid), dim3(TXQ_THREADS), 0, [&] { txq_kernel<32, MODE>(A); }); break;
    }
}

extern "C" int txq_emu_fwd(int tx_size, int tx_type, int partial32, int n, const int16_t *in, int16_t *out) {
    TxqArgs A; memset(&A, 0, sizeof(A));
    A.in16 = in; A.coeff = out; A.n = n; A.tx_type = tx_size == 3 ? 0 : tx_type; A.partial32 = partial32 && tx_size == 3;
    run_mode<1>(tx_size, A);
    return 0;
}
extern "C" int txq_emu_quant(int tx_size, int tx_type, int n, const int16_t *coeff, const SvtB200QuantRow *qrows, const uint16_t *qrow_index, const int16_t *const iscan[4],
                             int16_t *qcoeff, int16_t *dqcoeff, uint16_t *eob) {
    TxqArgs A; memset(&A, 0, sizeof(A));
    A.in16 = coeff; A.qrows = qrows; A.qrow_index = qrow_index; A.qcoeff = qcoeff; A.dqcoeff = dqcoeff; A.eob = eob; A.n = n; A.tx_type = tx_size == 3 ? 0 : tx_type;
    for (int t = 0; t < 4; t++) A.iscan[t] = iscan[t];
    run_mode<2>(tx_size, A);
    return 0;
}
extern "C" int txq_emu_inv(int tx_size, int tx_type, int n, const int16_t *dq, const uint16_t *eob, uint8_t *dest) {
    TxqArgs A; memset(&A, 0, sizeof(A));
    A.in16 = dq; A.eob_in = eob; A.dest = dest; A.n = n; A.tx_type = tx_size == 3 ? 0 : tx_type;
    run_mode<3>(tx_size, A);
    return 0;
}
struct TxqEmuBatch { int32_t tx_size, n; const SvtB200TxJob *jobs; SvtB200TxPlanes planes; const int16_t *iscan[4]; int16_t *coeff, *qcoeff, *dqcoeff; uint16_t *eob; };
/* the fused kernel: one batch through txq_kernel<N, 0>, or several through txq_multi_kernel */
extern "C" int txq_emu_fused(const TxqEmuBatch *b, int n_batches, const SvtB200QuantRow *qrows, int multi) {
    TxqMultiArgs M; memset(&M, 0, sizeof(M));
    int units = 0, nb = 0;
    for (int k = 0; k < n_batches; k++) {
        if (!b[k].n) continue;
        TxqArgs &A = M.batch[nb];
        A.jobs = b[k].jobs; A.pl = b[k].planes; A.qrows = qrows; A.coeff = b[k].coeff; A.qcoeff = b[k].qcoeff; A.dqcoeff = b[k].dqcoeff; A.eob = b[k].eob; A.n = b[k].n;
        for (int t = 0; t < 4; t++) A.iscan[t] = b[k].iscan[t];
        const int G = TXQ_THREADS / (4 << b[k].tx_size);
        units += (b[k].n + G - 1) / G;
        M.tx_size[nb] = b[k].tx_size; M.unit_end[nb] = units;
        nb++;
    }
    if (!nb) return 0;
    M.n_batches = nb;
    if (!multi) { for (int k = 0; k < nb; k++) run_mode<0>(M.tx_size[k], M.batch[k]); return 0; }
    emu_launch(dim3((unsigned)(units < 5 ? units : 5)), dim3(TXQ_THREADS), 0, [&] { txq_multi_kernel(M); });
    return 0;
}
