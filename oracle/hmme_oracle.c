/*
 * hmme_oracle.c -- CPU oracle (TEST INFRASTRUCTURE, see hmme_oracle.h for the rules and citations).
 *
 * Two independent derivations of the same semantics live here on purpose:
 *   hmme_oracle_search_ctu   : per candidate, 256 4x4 SADs -> 17x17 integral image -> every
 *                              partition SAD is a rectangle sum read from the layout table.
 *   hmme_oracle_search_frame : per candidate, the sums are built level by level from HEVC geometry
 *                              (8x4/4x8 -> 8x8 -> 16x8/8x16 -> ... -> 64x64) with the group bases
 *                              written out explicitly; threaded over jobs.
 * tests/test_oracle.py checks them against each other, against the lock-step run of the
 * reference's own kernels (oracle/_ref) and against the golden vectors.
 */
#include "hmme_oracle.h"

#include <limits.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define NPARTS HMME_ORACLE_NUM_PARTS

/* --- layout (TComDataCU.cpp:4676-6461 as decoded in SURVEY.md A.3): 36 contiguous groups ------- */
typedef struct { int base, n, w, h, nx, sx, sy, xo, yo; } group_t;
static const group_t kGroups[36] = {
    {0, 128, 8, 4, 8, 8, 4, 0, 0},      {128, 128, 4, 8, 16, 4, 8, 0, 0},
    {256, 16, 16, 4, 4, 16, 16, 0, 0},  {272, 16, 16, 4, 4, 16, 16, 0, 12},
    {288, 16, 16, 12, 4, 16, 16, 0, 0}, {304, 16, 16, 12, 4, 16, 16, 0, 4},
    {320, 16, 4, 16, 4, 16, 16, 0, 0},  {336, 16, 4, 16, 4, 16, 16, 12, 0},
    {352, 16, 12, 16, 4, 16, 16, 0, 0}, {368, 16, 12, 16, 4, 16, 16, 4, 0},
    {384, 64, 8, 8, 8, 8, 8, 0, 0},     {448, 32, 16, 8, 4, 16, 8, 0, 0},
    {480, 32, 8, 16, 8, 8, 16, 0, 0},   {512, 4, 32, 8, 2, 32, 32, 0, 0},
    {516, 4, 32, 8, 2, 32, 32, 0, 24},  {520, 4, 32, 24, 2, 32, 32, 0, 0},
    {524, 4, 32, 24, 2, 32, 32, 0, 8},  {528, 4, 8, 32, 2, 32, 32, 0, 0},
    {532, 4, 8, 32, 2, 32, 32, 24, 0},  {536, 4, 24, 32, 2, 32, 32, 0, 0},
    {540, 4, 24, 32, 2, 32, 32, 8, 0},  {544, 16, 16, 16, 4, 16, 16, 0, 0},
    {560, 8, 32, 16, 2, 32, 16, 0, 0},  {568, 8, 16, 32, 4, 16, 32, 0, 0},
    {576, 1, 64, 16, 1, 64, 64, 0, 0},  {577, 1, 64, 16, 1, 64, 64, 0, 48},
    {578, 1, 64, 48, 1, 64, 64, 0, 0},  {579, 1, 64, 48, 1, 64, 64, 0, 16},
    {580, 1, 16, 64, 1, 64, 64, 0, 0},  {581, 1, 16, 64, 1, 64, 64, 48, 0},
    {582, 1, 48, 64, 1, 64, 64, 0, 0},  {583, 1, 48, 64, 1, 64, 64, 16, 0},
    {584, 4, 32, 32, 2, 32, 32, 0, 0},  {588, 2, 64, 32, 1, 64, 32, 0, 0},
    {590, 2, 32, 64, 2, 32, 64, 0, 0},  {592, 1, 64, 64, 1, 64, 64, 0, 0},
};

void hmme_oracle_partition_table(hmme_oracle_rect out[NPARTS]) {
    for (int g = 0; g < 36; ++g) {
        const group_t* G = &kGroups[g];
        for (int k = 0; k < G->n; ++k) {
            hmme_oracle_rect* r = &out[G->base + k];
            r->x = G->xo + (k % G->nx) * G->sx;
            r->y = G->yo + (k / G->nx) * G->sy;
            r->w = G->w;
            r->h = G->h;
        }
    }
}

/* sad.cl:377-396: len=1; t = v<=0 ? -2v+1 : 2v; while (t != 1) { t >>= 1; len += 2; } */
uint32_t hmme_oracle_mv_bits(int v) {
    uint32_t len = 1;
    uint32_t t = (v <= 0) ? (uint32_t)(-v * 2) + 1u : (uint32_t)(v * 2);
    while (t != 1u) { t >>= 1; len += 2; }
    return len;
}

uint32_t hmme_oracle_lambda_q16(double lambda) { return (uint32_t)floor(65536.0 * sqrt(lambda)); }

/* sad.cl:398: tempSad + uiCost*(bitsX+bitsY)/65536 in 32-bit unsigned arithmetic (wraps). */
static inline uint32_t mv_cost(uint32_t lambda, int mvx, int mvy) {
    uint32_t bits = hmme_oracle_mv_bits(mvx * 4) + hmme_oracle_mv_bits(mvy * 4);
    return (uint32_t)(lambda * bits) / 65536u;
}

/* sad.cl:171-186: abs_diff(short, short) is exact (returns ushort), accumulated in unsigned int. */
static void base_sads(const int16_t* cur, int curStride, const int16_t* ref, long refStride,
                      uint32_t s4[16][16]) {
    for (int j = 0; j < 16; ++j)
        for (int i = 0; i < 16; ++i) {
            uint32_t s = 0;
            for (int r = 0; r < 4; ++r)
                for (int c = 0; c < 4; ++c) {
                    int a = cur[(4 * j + r) * curStride + 4 * i + c];
                    int b = ref[(long)(4 * j + r) * refStride + 4 * i + c];
                    s += (uint32_t)(a > b ? a - b : b - a);
                }
            s4[j][i] = s;
        }
}

int hmme_oracle_search_ctu(const int16_t* cur, int curStride, const int16_t* refAtCtu, int refStride,
                           int range, int ltx, int lty, uint32_t lambda,
                           int32_t* X, int32_t* Y, uint32_t* sad, uint32_t* cost) {
    if (!cur || !refAtCtu || range < 0 || !X || !Y || !sad || !cost) return -1;
    hmme_oracle_rect rect[NPARTS];
    hmme_oracle_partition_table(rect);
    for (int p = 0; p < NPARTS; ++p) { X[p] = 0; Y[p] = 0; sad[p] = 0; cost[p] = UINT_MAX; }
    /* TEncOpenCL.cpp:251: window origin; :312-313: y outer, x inner, both 0..2R inclusive */
    const int16_t* win = refAtCtu + (long)refStride * lty + ltx;
    for (int y = 0; y <= 2 * range; ++y)
        for (int x = 0; x <= 2 * range; ++x) {
            uint32_t s4[16][16], ii[17][17];
            base_sads(cur, curStride, win + (long)y * refStride + x, refStride, s4);
            for (int i = 0; i <= 16; ++i) ii[0][i] = 0;
            for (int j = 1; j <= 16; ++j) {
                ii[j][0] = 0;
                for (int i = 1; i <= 16; ++i)
                    ii[j][i] = s4[j - 1][i - 1] + ii[j - 1][i] + ii[j][i - 1] - ii[j - 1][i - 1];
            }
            const int mvx = x + ltx, mvy = y + lty;          /* TEncOpenCL.cpp:323-324 */
            const uint32_t mvc = mv_cost(lambda, mvx, mvy);
            for (int p = 0; p < NPARTS; ++p) {
                int x0 = rect[p].x >> 2, y0 = rect[p].y >> 2;
                int x1 = x0 + (rect[p].w >> 2), y1 = y0 + (rect[p].h >> 2);
                uint32_t s = ii[y1][x1] - ii[y0][x1] - ii[y1][x0] + ii[y0][x0];
                uint32_t c = s + mvc;                        /* 32-bit wrap like the kernel */
                if (c < cost[p]) { cost[p] = c; sad[p] = s; X[p] = mvx; Y[p] = mvy; }   /* sad.cl:400-406 */
            }
        }
    return 0;
}

/* ------------------------- explicit level-by-level hierarchy (second derivation) -------------- */
static void hierarchy(const uint32_t s[16][16], uint32_t out[NPARTS]) {
    uint32_t h[16][8], v[8][16], e[8][8], r16x4[16][4], c4x16[4][16];
    uint32_t p16x8[8][4], p8x16[4][8], b16[4][4], r32x8[8][2], c8x32[2][8];
    uint32_t p32x16[4][2], p16x32[2][4], b32[2][2], r64x16[4], c16x64[4];
    for (int j = 0; j < 16; ++j) for (int i = 0; i < 8; ++i) out[0 + j * 8 + i] = h[j][i] = s[j][2 * i] + s[j][2 * i + 1];
    for (int j = 0; j < 8; ++j) for (int i = 0; i < 16; ++i) out[128 + j * 16 + i] = v[j][i] = s[2 * j][i] + s[2 * j + 1][i];
    for (int j = 0; j < 8; ++j) for (int i = 0; i < 8; ++i) out[384 + j * 8 + i] = e[j][i] = h[2 * j][i] + h[2 * j + 1][i];
    for (int j = 0; j < 16; ++j) for (int k = 0; k < 4; ++k) r16x4[j][k] = h[j][2 * k] + h[j][2 * k + 1];
    for (int m = 0; m < 4; ++m) for (int i = 0; i < 16; ++i) c4x16[m][i] = v[2 * m][i] + v[2 * m + 1][i];
    for (int m = 0; m < 4; ++m) for (int k = 0; k < 4; ++k) {
        int o = m * 4 + k;
        out[256 + o] = r16x4[4 * m][k];
        out[272 + o] = r16x4[4 * m + 3][k];
        out[288 + o] = r16x4[4 * m][k] + r16x4[4 * m + 1][k] + r16x4[4 * m + 2][k];
        out[304 + o] = r16x4[4 * m + 1][k] + r16x4[4 * m + 2][k] + r16x4[4 * m + 3][k];
        out[320 + o] = c4x16[m][4 * k];
        out[336 + o] = c4x16[m][4 * k + 3];
        out[352 + o] = c4x16[m][4 * k] + c4x16[m][4 * k + 1] + c4x16[m][4 * k + 2];
        out[368 + o] = c4x16[m][4 * k + 1] + c4x16[m][4 * k + 2] + c4x16[m][4 * k + 3];
    }
    for (int j = 0; j < 8; ++j) for (int k = 0; k < 4; ++k) out[448 + j * 4 + k] = p16x8[j][k] = e[j][2 * k] + e[j][2 * k + 1];
    for (int m = 0; m < 4; ++m) for (int i = 0; i < 8; ++i) out[480 + m * 8 + i] = p8x16[m][i] = e[2 * m][i] + e[2 * m + 1][i];
    for (int m = 0; m < 4; ++m) for (int k = 0; k < 4; ++k) out[544 + m * 4 + k] = b16[m][k] = p16x8[2 * m][k] + p16x8[2 * m + 1][k];
    for (int j = 0; j < 8; ++j) for (int K = 0; K < 2; ++K) r32x8[j][K] = p16x8[j][2 * K] + p16x8[j][2 * K + 1];
    for (int M = 0; M < 2; ++M) for (int i = 0; i < 8; ++i) c8x32[M][i] = p8x16[2 * M][i] + p8x16[2 * M + 1][i];
    for (int M = 0; M < 2; ++M) for (int K = 0; K < 2; ++K) {
        int o = M * 2 + K;
        out[512 + o] = r32x8[4 * M][K];
        out[516 + o] = r32x8[4 * M + 3][K];
        out[520 + o] = r32x8[4 * M][K] + r32x8[4 * M + 1][K] + r32x8[4 * M + 2][K];
        out[524 + o] = r32x8[4 * M + 1][K] + r32x8[4 * M + 2][K] + r32x8[4 * M + 3][K];
        out[528 + o] = c8x32[M][4 * K];
        out[532 + o] = c8x32[M][4 * K + 3];
        out[536 + o] = c8x32[M][4 * K] + c8x32[M][4 * K + 1] + c8x32[M][4 * K + 2];
        out[540 + o] = c8x32[M][4 * K + 1] + c8x32[M][4 * K + 2] + c8x32[M][4 * K + 3];
    }
    for (int m = 0; m < 4; ++m) for (int K = 0; K < 2; ++K) out[560 + m * 2 + K] = p32x16[m][K] = b16[m][2 * K] + b16[m][2 * K + 1];
    for (int M = 0; M < 2; ++M) for (int k = 0; k < 4; ++k) out[568 + M * 4 + k] = p16x32[M][k] = b16[2 * M][k] + b16[2 * M + 1][k];
    for (int M = 0; M < 2; ++M) for (int K = 0; K < 2; ++K) out[584 + M * 2 + K] = b32[M][K] = p32x16[2 * M][K] + p32x16[2 * M + 1][K];
    for (int m = 0; m < 4; ++m) r64x16[m] = p32x16[m][0] + p32x16[m][1];
    for (int k = 0; k < 4; ++k) c16x64[k] = p16x32[0][k] + p16x32[1][k];
    out[576] = r64x16[0];
    out[577] = r64x16[3];
    out[578] = r64x16[0] + r64x16[1] + r64x16[2];
    out[579] = r64x16[1] + r64x16[2] + r64x16[3];
    out[580] = c16x64[0];
    out[581] = c16x64[3];
    out[582] = c16x64[0] + c16x64[1] + c16x64[2];
    out[583] = c16x64[1] + c16x64[2] + c16x64[3];
    out[588] = b32[0][0] + b32[0][1];
    out[589] = b32[1][0] + b32[1][1];
    out[590] = b32[0][0] + b32[1][0];
    out[591] = b32[0][1] + b32[1][1];
    out[592] = out[588] + out[589];
}

static void search_job_hier(const int16_t* cur, int curStride, const int16_t* refAtCtu, int refStride,
                            int range, int ltx, int lty, uint32_t lambda,
                            int32_t* X, int32_t* Y, uint32_t* sad, uint32_t* cost) {
    int16_t c[64][64];
    for (int r = 0; r < 64; ++r) memcpy(c[r], cur + (long)r * curStride, 64 * sizeof(int16_t));
    for (int p = 0; p < NPARTS; ++p) { X[p] = 0; Y[p] = 0; sad[p] = 0; cost[p] = UINT_MAX; }
    const int n = 2 * range + 1;
    uint32_t* bx = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)n);
    for (int x = 0; x < n; ++x) bx[x] = hmme_oracle_mv_bits((x + ltx) * 4);
    const int16_t* win = refAtCtu + (long)refStride * lty + ltx;
    for (int y = 0; y < n; ++y) {
        const uint32_t by = hmme_oracle_mv_bits((y + lty) * 4);
        for (int x = 0; x < n; ++x) {
            const int16_t* rp = win + (long)y * refStride + x;
            uint32_t s4[16][16], ps[NPARTS];
            for (int j = 0; j < 16; ++j) {
                uint32_t acc[16] = {0};
                for (int r = 0; r < 4; ++r) {
                    const int16_t* cr = c[4 * j + r];
                    const int16_t* rr = rp + (long)(4 * j + r) * refStride;
                    for (int i = 0; i < 16; ++i) {
                        uint32_t t = 0;
                        for (int k = 0; k < 4; ++k) {
                            int d = (int)cr[4 * i + k] - (int)rr[4 * i + k];
                            t += (uint32_t)(d < 0 ? -d : d);
                        }
                        acc[i] += t;
                    }
                }
                for (int i = 0; i < 16; ++i) s4[j][i] = acc[i];
            }
            hierarchy(s4, ps);
            const uint32_t mvc = (uint32_t)(lambda * (bx[x] + by)) / 65536u;
            const int mvx = x + ltx, mvy = y + lty;
            for (int p = 0; p < NPARTS; ++p) {
                uint32_t cc = ps[p] + mvc;
                if (cc < cost[p]) { cost[p] = cc; sad[p] = ps[p]; X[p] = mvx; Y[p] = mvy; }
            }
        }
    }
    free(bx);
}

typedef struct {
    const int16_t* cur; int curStride; const int16_t* ref; int refStride;
    const int32_t* jobs; int njobs; int range; uint32_t lambda;
    int32_t* X; int32_t* Y; uint32_t* sad; uint32_t* cost;
    int next; pthread_mutex_t mu;
} frame_work_t;

static void* frame_worker(void* arg) {
    frame_work_t* w = (frame_work_t*)arg;
    for (;;) {
        pthread_mutex_lock(&w->mu);
        int j = w->next++;
        pthread_mutex_unlock(&w->mu);
        if (j >= w->njobs) break;
        const int32_t* jb = w->jobs + 4 * j;
        const int16_t* cur = w->cur + (long)jb[1] * w->curStride + jb[0];
        const int16_t* ref = w->ref + (long)jb[1] * w->refStride + jb[0];
        search_job_hier(cur, w->curStride, ref, w->refStride, w->range, jb[2], jb[3], w->lambda,
                        w->X + (size_t)j * NPARTS, w->Y + (size_t)j * NPARTS,
                        w->sad + (size_t)j * NPARTS, w->cost + (size_t)j * NPARTS);
    }
    return NULL;
}

int hmme_oracle_search_frame(const int16_t* curOrigin, int curStride, const int16_t* refOrigin, int refStride,
                             const int32_t* jobs, int njobs, int range, uint32_t lambda, int nthreads,
                             int32_t* X, int32_t* Y, uint32_t* sad, uint32_t* cost) {
    if (!curOrigin || !refOrigin || !jobs || njobs < 0 || range < 0) return -1;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    frame_work_t w = {curOrigin, curStride, refOrigin, refStride, jobs, njobs, range, lambda, X, Y, sad, cost, 0,
                      PTHREAD_MUTEX_INITIALIZER};
    pthread_t th[256];
    int started = 0;
    for (int t = 1; t < nthreads; ++t)
        if (pthread_create(&th[started], NULL, frame_worker, &w) == 0) ++started;
    frame_worker(&w);
    for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
    pthread_mutex_destroy(&w.mu);
    return 0;
}
