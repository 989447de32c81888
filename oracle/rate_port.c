/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's coefficient rate estimation for one batch of blocks:
 * coeff_rate_estimate() with use_fast_coef_costing == 0 (Source/Lib/Codec/EbRateDistortionCost.c:55-167), i.e. the sequential
 * walk over the scan order with token_cache / band counter / previous-token state, exactly in the reference's order (the CUDA
 * kernel evaluates the same terms position by position in parallel).  Helpers restated: vp9_get_token_cost
 * (Source/Lib/VPX/vp9_tokenize.h:118), get_coef_context (vp9_scan.h:34), eb_vp9_pt_energy_class (vp9_entropy.c:111), band_counts
 * (EbRateDistortionCost.c:48).  Pinned against the reference's own function through oracle/ref_path_driver.c
 * (tests/test_rate.py).  Tables (scan, neighbours, value tokens / costs, token costs) are inputs, as in the C ABI.
 */
#include <stdint.h>
#include <stdlib.h>
#include "oracle.h"

static const int16_t k_band_counts[4][8] = {
    {1, 2, 3, 4, 3, 16 - 13, 0}, {1, 2, 3, 4, 11, 64 - 21, 0}, {1, 2, 3, 4, 11, 256 - 21, 0}, {1, 2, 3, 4, 11, 1024 - 21, 0}};
static const uint8_t k_energy[12] = {0, 1, 2, 3, 3, 4, 4, 5, 5, 5, 5, 5};

static int token_cost(int v, int *tok, const int16_t *vtok, const int32_t *vcost, const int16_t *c6lo, const uint16_t *c6hi) {
    if (v >= 67 || v <= -67) {
        const int e = abs(v) - 67;
        *tok = 10;
        return c6lo[e & 0xff] + c6hi[e >> 8];
    }
    *tok = vtok[v + 66];
    return vcost[v + 66];
}

/* token_costs: [plane type 2][ref 2][band 6][2][ctx 6][token 12] of this transform size; scan / neighbors of the block's (tx_size, tx_type) */
void svtb200_oracle_coeff_rate(int tx_size, int n, const int16_t *qcoeff, const uint16_t *eob, const SvtB200RateBlock *blocks,
                               const uint32_t *token_costs, const int16_t *const scans[4], const int16_t *const neighbors[4],
                               const int16_t *vtok, const int32_t *vcost, const int16_t *c6lo, const uint16_t *c6hi, int32_t *rate) {
    const int nn = (4 << tx_size) * (4 << tx_size);
    uint8_t *cache = (uint8_t *)malloc((size_t)nn);
    for (int b = 0; b < n; b++) {
        const int16_t *q = qcoeff + (size_t)b * nn;
        const SvtB200RateBlock *B = &blocks[b];
        const int tt = tx_size == 3 ? 0 : (B->tx_type & 3);
        const int16_t *scan = scans[tt], *nb = neighbors[tt];
        const uint32_t(*tc)[2][6][12] = (const uint32_t(*)[2][6][12])(token_costs + (size_t)((B->plane_type ? 1 : 0) * 2 + (B->is_inter ? 1 : 0)) * (6 * 2 * 6 * 12));
        const int16_t *band_count = &k_band_counts[tx_size][1];
        int pt = B->coeff_ctx, cost, tok, c;
        if (eob[b] == 0) { rate[b] = (int32_t)tc[0][0][pt][11]; continue; }
        int band_left = *band_count++;
        cost = token_cost(q[0], &tok, vtok, vcost, c6lo, c6hi);
        cost += (int)(*tc)[0][pt][tok];
        cache[0] = k_energy[tok];
        ++tc;
        const uint32_t(*tok_cost)[6][12] = &((*tc)[!tok]);
        for (c = 1; c < eob[b]; c++) {
            const int rc = scan[c];
            cost += token_cost(q[rc], &tok, vtok, vcost, c6lo, c6hi);
            pt = (1 + cache[nb[2 * c]] + cache[nb[2 * c + 1]]) >> 1;
            cost += (int)(*tok_cost)[pt][tok];
            cache[rc] = k_energy[tok];
            if (!--band_left) { band_left = *band_count++; ++tc; }
            tok_cost = &((*tc)[!tok]);
        }
        if (band_left) {
            pt = (1 + cache[nb[2 * c]] + cache[nb[2 * c + 1]]) >> 1;
            cost += (int)(*tc)[0][pt][11];
        }
        rate[b] = cost;
    }
    free(cache);
}
