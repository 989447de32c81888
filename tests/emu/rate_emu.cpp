// CPU execution of the coefficient rate / packing kernels' device code (test infrastructure only).
#define SVTB200_CPU_EMU 1
#include "../../svt-vp9_b200/csrc/emu/cuda_emu.h"
#include "../../include/svtb200.h"
#include "../../svt-vp9_b200/csrc/rate_kernel.cuh"
#include "../../svt-vp9_b200/csrc/pack_kernel.cuh"

struct RateEmuBatch { // host pointers
    int32_t tx_size, n;
    const int16_t *qcoeff; const uint16_t *eob; const SvtB200RateBlock *blocks; const uint32_t *token_costs;
    const int16_t *scan[4], *nb[4];
    int32_t *rate;
};
extern "C" int rate_emu_run(const RateEmuBatch *b, int n_batches, const int16_t *vtok, const int32_t *vcost, const int16_t *c6lo, const uint16_t *c6hi) {
    RateMultiArgs M;
    memset(&M, 0, sizeof(M));
    int total = 0;
    for (int k = 0; k < n_batches; k++) {
        if (!b[k].n) continue;
        RateArgs &A = M.b[M.nb];
        A.qcoeff = b[k].qcoeff; A.eob = b[k].eob; A.blk = b[k].blocks; A.tc = b[k].token_costs; A.rate = b[k].rate; A.n = b[k].n; A.tx_size = b[k].tx_size;
        for (int t = 0; t < 4; t++) { A.scan[t] = b[k].scan[t]; A.nb[t] = b[k].nb[t]; }
        A.vtok = vtok; A.vcost = vcost; A.c6lo = c6lo; A.c6hi = c6hi;
        total += b[k].n;
        M.end[M.nb++] = total;
    }
    if (!total) return 0;
    emu_launch(dim3((unsigned)((total + RATE_CHUNK - 1) / RATE_CHUNK)), dim3(RATE_THREADS), 0, [&] { rate_multi_kernel(M); });
    return 0;
}

struct PackEmuBatch { int32_t tx_size, n; const SvtB200TxJob *jobs; const int16_t *qcoeff; const uint16_t *eob; const int16_t *scan[4]; };
extern "C" int pack_emu_run(const PackEmuBatch *b, int n_batches, int16_t *stream, uint32_t capacity, uint32_t *offsets) {
    PackArgs A;
    memset(&A, 0, sizeof(A));
    int total = 0;
    for (int k = 0; k < n_batches; k++) {
        if (!b[k].n) continue;
        PackBatch &P = A.b[A.nb++];
        P.jobs = b[k].jobs; P.qcoeff = b[k].qcoeff; P.eob = b[k].eob; P.first = total; P.n = b[k].n; P.nn = (4 << b[k].tx_size) * (4 << b[k].tx_size);
        for (int t = 0; t < 4; t++) P.scan[t] = b[k].scan[t];
        total += b[k].n;
    }
    A.n_total = total; A.offsets = offsets; A.stream = stream; A.capacity = capacity;
    const int chunks = total ? (total + PACK_CHUNK - 1) / PACK_CHUNK : 1;
    std::vector<uint32_t> chunk_base((size_t)chunks + 64, 0u);
    uint32_t ticket = 0;
    emu_launch(dim3((unsigned)chunks), dim3(PACK_THREADS), 0, [&] { pack_sum_kernel(A, chunk_base.data(), &ticket); });
    if (total) emu_launch(dim3((unsigned)chunks), dim3(PACK_THREADS), 0, [&] { pack_kernel(A, chunk_base.data()); });
    return 0;
}
