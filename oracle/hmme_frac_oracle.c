/*
 * hmme_frac_oracle.c -- CPU oracle of the fractional-pel refinement that follows the integer search (SURVEY.md section 8,
 * row f1).  TEST INFRASTRUCTURE ONLY, same rules as hmme_oracle.h: only tests/, smoke() and bench.py's CPU legs load it.
 *
 * Plain-C restatement (never a copy) of
 *   TEncSearch::xPatternSearchFracDIF .... /root/reference/source/Lib/TLibEncoder/TEncSearch.cpp:4294-4331
 *   TEncSearch::xPatternRefinement ....... TEncSearch.cpp:816-872   (candidate order tables :51-75, strict '<')
 *   xExtDIFUpSamplingH / Q ............... TEncSearch.cpp:5386-5600 (which sample positions the 16 planes hold)
 *   TComInterpolationFilter::filter ...... TLibCommon/TComInterpolationFilter.cpp:57-63,155-250 (8-tap luma, 14-bit
 *                                          intermediate, horizontal pass first, rounding of the last pass)
 *   TComRdCost::xGetHADs / xCalcHADs8x8 / xCalcHADs4x4 ... TLibCommon/TComRdCost.cpp:1343-1600
 *   TComRdCost::getCost / getBits ........ TLibCommon/TComRdCost.h:166-185, TComRdCost.cpp:278-292
 *
 * The reference fills planes m_filteredBlock[fy][fx] and addresses them with +1 / +stride fix-ups; every sample it reads
 * is the HEVC luma prediction sample at quarter-pel position 4*(integer MV) + (dx, dy), computed horizontally first:
 *   h(r, c)   = sum_k C[fx][k] * ref[r][c + ix + k - 3]                (64*ref when fx == 0; the -8192 offset cancels)
 *   fy == 0 : clip255((h + 32) >> 6)          fy != 0 : clip255((sum_k C[fy][k] * h(r + iy + k - 3, c) + 2048) >> 12)
 * with ix = dx >> 2, fx = dx & 3 (same for y).  Pinned against records logged from the reference encoder itself
 * (oracle/gen_frac_golden.py, tests/golden/frac_records.npz).
 */
#include "hmme_oracle.h"

#include <stdlib.h>
#include <string.h>

static const int kLuma[4][8] = {
    {0, 0, 0, 64, 0, 0, 0, 0}, {-1, 4, -10, 58, 17, -5, 1, 0}, {-1, 4, -11, 40, 40, -11, 4, -1}, {0, 1, -5, 17, 58, -10, 4, -1}};

static int clip255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

/* prediction block of size w x h whose top-left integer sample is ref0, displaced by (dx, dy) quarter-pels */
static void predict(const int16_t* ref0, int stride, int dx, int dy, int w, int h, int16_t* out /* w*h */) {
    const int ix = dx >> 2, fx = dx & 3, iy = dy >> 2, fy = dy & 3;
    int* hp = (int*)malloc(sizeof(int) * (size_t)w * (size_t)(h + 8));
    for (int r = -3; r < h + 5; ++r)                                  /* rows iy-3 .. iy+h+4 relative to the block */
        for (int c = 0; c < w; ++c) {
            const int16_t* s = ref0 + (long)(r + iy) * stride + c + ix;
            int acc = 0;
            for (int k = 0; k < 8; ++k) acc += kLuma[fx][k] * s[k - 3];
            hp[(r + 3) * w + c] = acc;
        }
    for (int r = 0; r < h; ++r)
        for (int c = 0; c < w; ++c) {
            int v;
            if (fy == 0) v = (hp[(r + 3) * w + c] + 32) >> 6;
            else {
                int acc = 0;
                for (int k = 0; k < 8; ++k) acc += kLuma[fy][k] * hp[(r + k) * w + c];
                v = (acc + 2048) >> 12;
            }
            out[r * w + c] = (int16_t)clip255(v);
        }
    free(hp);
}

/* sum |H D H| over an n x n block (n = 4 or 8); the sum of magnitudes does not depend on the row order of H */
static uint32_t hadamard_abs_sum(const int* d, int n) {
    int a[64], b[64];
    memcpy(a, d, sizeof(int) * (size_t)(n * n));
    for (int pass = 0; pass < 2; ++pass) {                            /* rows, then columns (via transpose) */
        for (int r = 0; r < n; ++r) {
            int* v = a + r * n;
            for (int len = 1; len < n; len <<= 1)
                for (int i = 0; i < n; i += 2 * len)
                    for (int j = i; j < i + len; ++j) { const int p = v[j], q = v[j + len]; v[j] = p + q; v[j + len] = p - q; }
        }
        for (int r = 0; r < n; ++r) for (int c = 0; c < n; ++c) b[c * n + r] = a[r * n + c];
        memcpy(a, b, sizeof(int) * (size_t)(n * n));
    }
    uint32_t s = 0;
    for (int i = 0; i < n * n; ++i) s += (uint32_t)abs(a[i]);
    return s;
}

static uint32_t distortion(const int16_t* cur, int curStride, const int16_t* pred, int w, int h, int useHad) {
    uint32_t sum = 0;
    if (!useHad) {
        for (int r = 0; r < h; ++r) for (int c = 0; c < w; ++c) sum += (uint32_t)abs(cur[r * curStride + c] - pred[r * w + c]);
        return sum;
    }
    const int n = (w % 8 == 0 && h % 8 == 0) ? 8 : 4;                /* xGetHADs: 8x8 transforms when both sides allow */
    int d[64];
    for (int y = 0; y < h; y += n)
        for (int x = 0; x < w; x += n) {
            for (int r = 0; r < n; ++r) for (int c = 0; c < n; ++c) d[r * n + c] = cur[(y + r) * curStride + x + c] - pred[(y + r) * w + x + c];
            const uint32_t s = hadamard_abs_sum(d, n);
            sum += (n == 8) ? ((s + 2) >> 2) : ((s + 1) >> 1);
        }
    return sum;
}

static uint32_t mv_cost(uint32_t lambda, int x, int y, int scale, int predx, int predy) {
    const uint32_t bits = hmme_oracle_mv_bits((int)((unsigned)x << scale) - predx) + hmme_oracle_mv_bits((int)((unsigned)y << scale) - predy);
    return (uint32_t)(lambda * bits) >> 16;
}

static const int kHalf[9][2] = {{0, 0}, {0, -1}, {0, 1}, {-1, 0}, {1, 0}, {-1, -1}, {1, -1}, {-1, 1}, {1, 1}};
static const int kQter[9][2] = {{0, 0}, {0, -1}, {0, 1}, {-1, -1}, {1, -1}, {-1, 0}, {1, 0}, {-1, 1}, {1, 1}};

int hmme_oracle_refine_frac(const int16_t* curOrigin, int curStride, const int16_t* refOrigin, int refStride,
                            const hmme_oracle_pu* pus, int npus, uint32_t lambda, int useHad,
                            int32_t* mvq, int32_t* half, int32_t* qter, uint32_t* cost, uint32_t* dist, uint32_t* cand) {
    for (int n = 0; n < npus; ++n) {
        const hmme_oracle_pu* P = &pus[n];
        if (P->w <= 0 || P->h <= 0 || P->w > 64 || P->h > 64 || (P->w & 3) || (P->h & 3)) return -1;
        const int16_t* cur = curOrigin + (long)P->y * curStride + P->x;
        const int16_t* ref0 = refOrigin + (long)(P->y + P->mvy) * refStride + P->x + P->mvx;
        int16_t pred[64 * 64];
        /* half-pel stage: cost scale 1, candidates around the integer MV in half-pel units */
        uint32_t best = 0xFFFFFFFFu; int bi = 0; uint32_t bestMvc = 0;
        for (int i = 0; i < 9; ++i) {
            predict(ref0, refStride, 2 * kHalf[i][0], 2 * kHalf[i][1], P->w, P->h, pred);
            const uint32_t mvc = mv_cost(lambda, 2 * P->mvx + kHalf[i][0], 2 * P->mvy + kHalf[i][1], 1, P->predx, P->predy);
            const uint32_t c = distortion(cur, curStride, pred, P->w, P->h, useHad) + mvc;
            if (cand) cand[18 * n + i] = c;
            if (c < best) { best = c; bi = i; }
        }
        const int hx = kHalf[bi][0], hy = kHalf[bi][1];
        /* quarter-pel stage: cost scale 0, candidates around the half-pel winner in quarter-pel units */
        best = 0xFFFFFFFFu; int bq = 0;
        for (int i = 0; i < 9; ++i) {
            const int dx = 2 * hx + kQter[i][0], dy = 2 * hy + kQter[i][1];
            predict(ref0, refStride, dx, dy, P->w, P->h, pred);
            const uint32_t mvc = mv_cost(lambda, 4 * P->mvx + dx, 4 * P->mvy + dy, 0, P->predx, P->predy);
            const uint32_t c = distortion(cur, curStride, pred, P->w, P->h, useHad) + mvc;
            if (cand) cand[18 * n + 9 + i] = c;
            if (c < best) { best = c; bq = i; bestMvc = mvc; }
        }
        half[2 * n] = hx; half[2 * n + 1] = hy;
        qter[2 * n] = kQter[bq][0]; qter[2 * n + 1] = kQter[bq][1];
        mvq[2 * n] = 4 * P->mvx + 2 * hx + kQter[bq][0];
        mvq[2 * n + 1] = 4 * P->mvy + 2 * hy + kQter[bq][1];
        cost[n] = best; dist[n] = best - bestMvc;
    }
    return 0;
}

/* Distortion of the motion-compensated uni-prediction at a quarter-pel MV:
 * TEncSearch::xGetTemplateCost (TEncSearch.cpp:3634-3674) = xPredInterBlk (TComPrediction.cpp, uni-directional: the same two-pass
 * 8-tap interpolation with final rounding) + TComRdCost::getDistPart(DF_SAD) (TComRdCost.cpp:433-455); with useHad the distortion
 * of xGetInterPredictionError (TEncSearch.cpp:2814-2836).  The MV is taken as given (the caller has applied clipMv). */
int hmme_oracle_mc_cost(const int16_t* curOrigin, int curStride, const int16_t* refOrigin, int refStride,
                        const hmme_oracle_mc_pu* pus, int npus, int useHad, uint32_t* dist) {
    for (int n = 0; n < npus; ++n) {
        const hmme_oracle_mc_pu* P = &pus[n];
        if (P->w <= 0 || P->h <= 0 || P->w > 64 || P->h > 64 || (P->w & 3) || (P->h & 3)) return -1;
        int16_t pred[64 * 64];
        predict(refOrigin + (long)P->y * refStride + P->x, refStride, P->mvqx, P->mvqy, P->w, P->h, pred);
        dist[n] = distortion(curOrigin + (long)P->y * curStride + P->x, curStride, pred, P->w, P->h, useHad);
    }
    return 0;
}

/* 14-bit bi-prediction intermediate of one list, plus 8192: xPredInterBlk with bi = true (TComPrediction.cpp:669-707) stores
 * h - 8192 (fy == 0), sum_k C[fy][k]*ref - 8192 (fx == 0) or ((sum_k C[fy][k]*(h - 8192)) >> 6) (both fractional,
 * TComInterpolationFilter.cpp:203-224) -- in every case (sum_k C[fy][k] * h(r + k - 3) >> 6) - 8192 with h the horizontal sums. */
static void predict_raw(const int16_t* ref0, int stride, int dx, int dy, int w, int h, int* out /* w*h */) {
    const int ix = dx >> 2, fx = dx & 3, iy = dy >> 2, fy = dy & 3;
    int* hp = (int*)malloc(sizeof(int) * (size_t)w * (size_t)(h + 8));
    for (int r = -3; r < h + 5; ++r)
        for (int c = 0; c < w; ++c) {
            const int16_t* s = ref0 + (long)(r + iy) * stride + c + ix;
            int acc = 0;
            for (int k = 0; k < 8; ++k) acc += kLuma[fx][k] * s[k - 3];
            hp[(r + 3) * w + c] = acc;
        }
    for (int r = 0; r < h; ++r)
        for (int c = 0; c < w; ++c) {
            int acc = 0;
            for (int k = 0; k < 8; ++k) acc += kLuma[fy][k] * hp[(r + k) * w + c];
            out[r * w + c] = acc >> 6;
        }
    free(hp);
}

/* Bi-directional PUs: TComPrediction::xPredInterBi (TComPrediction.cpp:603-651) + TComYuv::addAvg (TComYuv.cpp:352-410):
 * pred = clip255((p0 + p1 + 64 + 2*8192) >> 7) with p = the intermediates above; the 8192 offsets cancel. */
int hmme_oracle_mc_cost_bi(const int16_t* curOrigin, int curStride, const int16_t* ref0Origin, int ref0Stride,
                           const int16_t* ref1Origin, int ref1Stride, const hmme_oracle_mc_bi_pu* pus, int npus, int useHad, uint32_t* dist) {
    for (int n = 0; n < npus; ++n) {
        const hmme_oracle_mc_bi_pu* P = &pus[n];
        if (P->w <= 0 || P->h <= 0 || P->w > 64 || P->h > 64 || (P->w & 3) || (P->h & 3)) return -1;
        static __thread int q0[64 * 64], q1[64 * 64];
        int16_t pred[64 * 64];
        predict_raw(ref0Origin + (long)P->y * ref0Stride + P->x, ref0Stride, P->mv0x, P->mv0y, P->w, P->h, q0);
        predict_raw(ref1Origin + (long)P->y * ref1Stride + P->x, ref1Stride, P->mv1x, P->mv1y, P->w, P->h, q1);
        for (int i = 0; i < P->w * P->h; ++i) pred[i] = (int16_t)clip255((q0[i] + q1[i] + 64) >> 7);
        dist[n] = distortion(curOrigin + (long)P->y * curStride + P->x, curStride, pred, P->w, P->h, useHad);
    }
    return 0;
}
