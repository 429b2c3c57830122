// me_frac_kernel.cuh -- fractional-pel refinement of prediction units after the integer search.
//
// Semantics: TEncSearch::xPatternSearchFracDIF (/root/reference/source/Lib/TLibEncoder/TEncSearch.cpp:4294-4331) =
//   half-pel stage : 9 candidates around the integer MV (table s_acMvRefineH, :51-62), cost scale 1
//   quarter stage  : 9 candidates around the half-pel winner (table s_acMvRefineQ, :64-75), cost scale 0
//   each candidate : HEVC 8-tap luma interpolation, horizontal pass first with a 14-bit intermediate
//                    (TComInterpolationFilter.cpp:57-63,155-250; which positions the reference's planes hold: TEncSearch.cpp:5386-5600),
//                    distortion = Hadamard SATD in 8x8 blocks when both PU sides are multiples of 8, else 4x4 blocks
//                    (TComRdCost::xGetHADs, TComRdCost.cpp:1537-1600), or plain SAD when HadamardME is off,
//                    plus lambda * bits(mv - predictor) >> 16 (TComRdCost.h:166-185); strict '<' in table order (:816-872).
// Nothing here is copied from the reference: every candidate sample is computed from the closed form
//   h(r, c) = sum_k C[fx][k] * ref[r][c + ix + k - 3],   pred = clip255((sum_k C[fy][k] * h(r + iy + k - 3, c) + 2048) >> 12)
// (ix = dx >> 2, fx = dx & 3; C[0] = {0,0,0,64,0,0,0,0} makes the integer cases fall out of the same arithmetic), which is what the
// reference's plane bookkeeping evaluates (the tests pin this against the 18 candidate costs logged from the reference encoder).
//
// Mapping: one warp per PU, looping over its 8x8 tiles; both stages are 3x3 grids (dx in 3 values) x (dy in 3 values):
//   H step : lane = (patch row 0..15, column half) -> 3 horizontal planes, two dp4a per sample (u8 samples x s8 taps)
//   V step : lane = (dx index, column) -> the column of its plane in registers, 3 dy x 8 rows of 8-tap sums, difference against
//            the current block and the vertical Hadamard pass, all in registers
//   SATD   : transposed through warp-private shared memory, lane = (candidate, coefficient row) does the horizontal pass
// No CTA-wide synchronisation; warps are independent and take PUs round-robin (callers order PUs large to small).
#pragma once
#include "me_common.cuh"

namespace hmme {

constexpr int kFracWarps = 4;
constexpr int kFracThreads = kFracWarps * 32;
constexpr int kFracCoopTiles = 8;      // PUs of this many 8x8 tiles or more are shared by the 4 warps of a CTA

struct FracPu { int x, y, w, h, mvx, mvy, predx, predy; };   // == hmme_pu (include/hmme_b200.h)

struct FracParams {
    const void* cur;            // picture sample (0,0); u8 or s16 (bi-prediction target 2*org - pred)
    const uint8_t* ref;         // picture sample (0,0); 8-bit
    long long curPitch, refPitch;
    int curBytes;
    const FracPu* pus;
    const int* slots;           // optional: result index of PU n (whole-frame path: job * 593 + partition)
    int npus;
    int nBig;                   // PUs [0, nBig) have at least kFracCoopTiles tiles and get a whole CTA each (one per CTA, 4 warps share the tiles)
    uint32_t lambda;
    int useHad;
    int4* out;                  // {mv x, mv y (quarter pel), cost, distortion}
    uint32_t* cand;             // optional [npus][18]: cost of every candidate in the reference's table order
};
// group form (me_frac_group_kernel): the list is cut into five segments -- PUs of >= 4 tiles (one PU per warp), 2..3 tiles with the
// 8x8 / the 4x4 Hadamard (two PUs per warp), one tile with the 8x8 / the 4x4 Hadamard (four PUs per warp)
struct FracGroupParams : FracParams {
    int segPu[6];               // first PU of segment q (segPu[5] = npus)
    int segGrp[6];              // first group of segment q (segGrp[5] = number of groups)
};
constexpr int kFracSegPus[5] = {1, 2, 2, 4, 4};    // PUs per group (= per warp at a time) in each segment
// segment of a PU: 0 = four tiles or more; 1 / 2 = two or three tiles, 8x8 / 4x4 Hadamard; 3 / 4 = one tile, 8x8 / 4x4 Hadamard
__host__ __device__ inline int frac_segment(int w, int h) {
    const int nT = ((w + 7) >> 3) * ((h + 7) >> 3);
    if (nT >= 4) return 0;
    return (nT >= 2 ? 1 : 3) + (((w | h) & 7) ? 1 : 0);
}

// HEVC luma taps (TComInterpolationFilter.cpp:57-63), as ints and as packed signed bytes {k0..k3}, {k4..k7}
__constant__ int kLumaTap[4][8] = {
    {0, 0, 0, 64, 0, 0, 0, 0}, {-1, 4, -10, 58, 17, -5, 1, 0}, {-1, 4, -11, 40, 40, -11, 4, -1}, {0, 1, -5, 17, 58, -10, 4, -1}};
__constant__ uint32_t kLumaPack[4][2] = {
    {0x40000000u, 0x00000000u}, {0x3AF604FFu, 0x0001FB11u}, {0x28F504FFu, 0xFF04F528u}, {0x11FB0100u, 0xFF04F63Au}};

struct __align__(16) FracScratch {
    uint32_t ref[16][4];        // 16 x 16 reference samples: rows/cols -4..11 of the tile, displaced by the integer MV
    int16_t cur[8][8];          // current tile, [column][row], zero outside the PU
    int16_t h[3][8][24];        // horizontal planes [dx index][column][patch row], 16 used (+8: conflict-free 128-bit column loads)
    int16_t t[9][72];           // vertically transformed differences [candidate][column * 8 + coefficient row] (+8 pad)
};

__device__ __forceinline__ int dp4a_us(uint32_t samples, uint32_t taps, int acc) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(samples), "r"(taps), "r"(acc));
    return d;
}

__device__ __forceinline__ int dp2a_lo(uint32_t pair, uint32_t taps, int acc) {   // acc + pair.lo * taps.b0 + pair.hi * taps.b1
    int d;
    asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(pair), "r"(taps), "r"(acc));
    return d;
}
__device__ __forceinline__ int dp2a_hi(uint32_t pair, uint32_t taps, int acc) {   // acc + pair.lo * taps.b2 + pair.hi * taps.b3
    int d;
    asm("dp2a.hi.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(pair), "r"(taps), "r"(acc));
    return d;
}
__device__ __forceinline__ int frac_round_clip(int a) { return __vimin_s32_relu(a >> 12, 255); }   // one VIMNMX.RELU: clip to [0, 255]

// Vertical 8-tap pass over one column held as packed pairs of 16-bit rows (wv[q] = rows 2q, 2q+1): output r starts at row
// r + S0.  An even start takes the taps as packed {k0..k3}, {k4..k7}; an odd start uses the same bytes moved up by one.
template <int S0, int NOUT, bool RAW = false>
__device__ __forceinline__ void frac_vfilter(const uint32_t (&wv)[8], const uint32_t ca, const uint32_t cb, int (&pr)[NOUT]) {
    const uint32_t o1 = ca << 8, o2 = (ca >> 24) | (cb << 8), o3 = cb >> 24;
#pragma unroll
    for (int r = 0; r < NOUT; ++r) {
        const int s = r + S0, m = s >> 1;
        int a = RAW ? 0 : 2048;
        if (s & 1) {
            a = dp2a_lo(wv[m], o1, a); a = dp2a_hi(wv[m + 1], o1, a); a = dp2a_lo(wv[m + 2], o2, a); a = dp2a_hi(wv[m + 3], o2, a);
            a = dp2a_lo(wv[m + 4], o3, a);
        } else {
            a = dp2a_lo(wv[m], ca, a); a = dp2a_hi(wv[m + 1], ca, a); a = dp2a_lo(wv[m + 2], cb, a); a = dp2a_hi(wv[m + 3], cb, a);
        }
        pr[r] = RAW ? (a >> 6) : frac_round_clip(a);          // RAW: the 14-bit bi-prediction intermediate (+8192), TComInterpolationFilter.cpp:203-224
    }
}

template <int N>
__device__ __forceinline__ void hadamard_inplace(int* d) {
#pragma unroll
    for (int len = 1; len < N; len <<= 1)
#pragma unroll
        for (int i = 0; i < N; i += 2 * len)
#pragma unroll
            for (int j = i; j < i + len; ++j) { const int a = d[j], b = d[j + len]; d[j] = a + b; d[j + len] = a - b; }
}

constexpr int kFracSad = 0, kFracHad4 = 1, kFracHad8 = 2;

// Distortion of the candidates of one PU and one stage, summed over its tiles.
//   HALF  : dx, dy in {-2, 0, 2} (all phases known at compile time): the two outer rows (columns) of candidates read the same
//           half-sample plane one row (column) apart, the centre row is a rounding copy;
//   !HALF : dx = cx + {-1, 0, 1}, dy = cy + {-1, 0, 1} around the half-pel winner (cx, cy); the centre candidate is that winner,
//           whose distortion is already known, so the Hadamard pass skips it.
// Result layout (per lane, 3 registers):
//   Hadamard, half stage    : candidate g = j*3+i in acc[g >> 2] on the lanes with (lane >> 3) == (g & 3)
//   Hadamard, quarter stage : same with k = g - (g > 4) in place of g (g = 4 is not computed)
//   SAD mode                : candidate g in acc[j] on the lanes with (lane >> 3) == i
template <int MODE, bool HALF, bool COOP>
__device__ __forceinline__ void frac_eval(const FracParams& p, const FracPu& P, FracScratch& S, const int lane, const int cx, const int cy,
                                          uint32_t (&acc)[3]) {
    const int t0 = COOP ? (int)(threadIdx.x >> 5) : 0, tstep = COOP ? kFracWarps : 1;
    const int step = HALF ? 2 : 1;
    int hOff[3], vOff[3];                     // first tap of output 0 inside the patch row / column: 0 or 1 (= 1 + (d >> 2))
    uint32_t cLo[3], cHi[3], vLo[3], vHi[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int dx = cx + (i - 1) * step, dy = cy + (i - 1) * step;
        hOff[i] = 1 + (dx >> 2); vOff[i] = 1 + (dy >> 2);
        cLo[i] = kLumaPack[dx & 3][0]; cHi[i] = kLumaPack[dx & 3][1];
        vLo[i] = kLumaPack[dy & 3][0]; vHi[i] = kLumaPack[dy & 3][1];
    }
    acc[0] = acc[1] = acc[2] = 0;
    const uint8_t* refPu = p.ref + (long long)(P.y + P.mvy - 4) * p.refPitch + (P.x + P.mvx - 4);
    const int rowL = lane >> 1, halfL = lane & 1;            // patch loader / H step role
    const int curR = lane >> 2, curC = (lane & 3) * 2;       // current-tile loader role

    // next tile's samples travel in registers while the current tile is being worked on
    // (a lane's 8 patch bytes, any alignment: three aligned 32-bit words + two funnel shifts instead of eight byte loads and their
    // packing; the planes carry slack bytes behind the last row, see hmme_plane)
    uint32_t nw0, nw1; int ncur[2];
    auto fetch = [&](const int tx, const int ty) {
        const uint8_t* g = refPu + (long long)(ty + rowL) * p.refPitch + tx + halfL * 8;
        {
            const uint32_t sh = 8u * (uint32_t)((uintptr_t)g & 3);
            const uint32_t* qa = reinterpret_cast<const uint32_t*>((uintptr_t)g & ~(uintptr_t)3);
            const uint32_t a0 = __ldg(qa), a1 = __ldg(qa + 1), a2 = __ldg(qa + 2);
            nw0 = __funnelshift_r(a0, a1, sh);
            nw1 = __funnelshift_r(a1, a2, sh);
        }
        const int tw = min(8, P.w - tx), th = min(8, P.h - ty);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            ncur[q] = 0;
            if (curR < th && curC + q < tw) {
                const long long o = (long long)(P.y + ty + curR) * p.curPitch + P.x + tx + curC + q;
                ncur[q] = p.curBytes == 1 ? (int)static_cast<const uint8_t*>(p.cur)[o] : (int)static_cast<const int16_t*>(p.cur)[o];
            }
        }
    };
    // this warp's tiles: t0, t0 + tstep, ... in raster order over the PU
    const int ntxT = (P.w + 7) >> 3, nT = ntxT * ((P.h + 7) >> 3);
    const int rcp = (65536 + ntxT - 1) / ntxT;               // t / ntxT == (t * rcp) >> 16 for t < 128, ntxT <= 8
    int t = t0, ty = ((t * rcp) >> 16) * 8, tx = t * 8 - ty * ntxT;
    fetch(tx, ty);
    while (true) {
        const int tw = min(8, P.w - tx), th = min(8, P.h - ty);
        {
            *reinterpret_cast<uint2*>(&S.ref[rowL][halfL * 2]) = make_uint2(nw0, nw1);
            S.cur[curC][curR] = (int16_t)ncur[0];
            S.cur[curC + 1][curR] = (int16_t)ncur[1];
        }
        __syncwarp();
        int tn = t + tstep, ntx, nty;
        bool more;
        if (COOP) { nty = ((tn * rcp) >> 16) * 8; ntx = tn * 8 - nty * ntxT; more = tn < nT; }
        else { ntx = tx + 8; nty = ty; if (ntx >= P.w) { ntx = 0; nty += 8; } more = nty < P.h; }   // raster walk, no index arithmetic
        if (more) fetch(ntx, nty);
        {   // H step: 4 columns x 3 planes per lane; columns outside the PU become 0 (so does their prediction, and cur is 0 there)
            const bool outside = MODE != kFracHad8 && halfL && tw < 8;       // zero samples -> zero planes
            const uint32_t W0 = outside ? 0u : S.ref[rowL][halfL], W1 = outside ? 0u : S.ref[rowL][halfL + 1], W2 = outside ? 0u : S.ref[rowL][halfL + 2];
            if (HALF) {
                // dx = -2 and dx = +2 are the same half-sample row one column apart (5 windows serve both); dx = 0 is 64 * sample
                int f[5];
#pragma unroll
                for (int o = 0; o < 5; ++o) f[o] = dp4a_us(__funnelshift_rc(W1, W2, 8 * o), cHi[0], dp4a_us(__funnelshift_rc(W0, W1, 8 * o), cLo[0], 0));
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    S.h[0][halfL * 4 + j][rowL] = (int16_t)f[j];
                    S.h[2][halfL * 4 + j][rowL] = (int16_t)f[j + 1];
                    S.h[1][halfL * 4 + j][rowL] = (int16_t)(((W1 >> (8 * j)) & 0xFFu) << 6);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    int out[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int o8 = 8 * (j + hOff[i]);                   // byte window start (x8): clamp mode makes 32 mean "next word"
                        out[j] = dp4a_us(__funnelshift_rc(W1, W2, o8), cHi[i], dp4a_us(__funnelshift_rc(W0, W1, o8), cLo[i], 0));
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) S.h[i][halfL * 4 + j][rowL] = (int16_t)out[j];
                }
            }
        }
        __syncwarp();
        {   // V step: lane = (dx index, column); lanes 24..31 compute on plane 0 and their results are never read
            const int di = lane >> 3, c = lane & 7;
            const bool live = lane < 24;
            int cu[8];
            const int dii = live ? di : 0;
            const uint4 ha = *reinterpret_cast<const uint4*>(&S.h[dii][c][0]), hb = *reinterpret_cast<const uint4*>(&S.h[dii][c][8]);
            const uint32_t wv[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
            {
                const uint4 cc = *reinterpret_cast<const uint4*>(&S.cur[c][0]);
                const uint32_t wc[4] = {cc.x, cc.y, cc.z, cc.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) { cu[2 * q] = (int)(int16_t)(wc[q] & 0xFFFFu); cu[2 * q + 1] = (int)wc[q] >> 16; }
            }
            // prediction column d[0..7] of candidate row j -> difference, vertical transform, hand-over
            auto emit = [&](const int j, int (&d)[8]) {
#pragma unroll
                for (int r = 0; r < 8; ++r) d[r] = cu[r] - d[r];
                if (MODE != kFracHad8 && th < 8) { d[4] = 0; d[5] = 0; d[6] = 0; d[7] = 0; }
                if (MODE == kFracSad) {
                    uint32_t s = 0;
#pragma unroll
                    for (int r = 0; r < 8; ++r) s = __sad(d[r], 0, s);
                    s += __shfl_xor_sync(0xFFFFFFFFu, s, 1); s += __shfl_xor_sync(0xFFFFFFFFu, s, 2); s += __shfl_xor_sync(0xFFFFFFFFu, s, 4);
                    acc[j] += s;
                } else {
                    if (MODE == kFracHad8) hadamard_inplace<8>(d);
                    else { hadamard_inplace<4>(d); hadamard_inplace<4>(d + 4); }
                    uint4 pk;
                    pk.x = __byte_perm(d[0], d[1], 0x5410); pk.y = __byte_perm(d[2], d[3], 0x5410);
                    pk.z = __byte_perm(d[4], d[5], 0x5410); pk.w = __byte_perm(d[6], d[7], 0x5410);
                    if (live) *reinterpret_cast<uint4*>(&S.t[j * 3 + di][c * 8]) = pk;
                }
            };
            int d[8];
            // integer vertical phase: rows 4..11 of the column (patch rows 0..7 of the tile), rounded from the 14-bit intermediate
            auto copy_rows = [&]() {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    d[2 * q] = __vimin_s32_relu(((int)(int16_t)(wv[q + 2] & 0xFFFFu) + 32) >> 6, 255);
                    d[2 * q + 1] = __vimin_s32_relu((((int)wv[q + 2] >> 16) + 32) >> 6, 255);
                }
            };
            if (HALF) {
                int o9[9];                                   // half-sample rows -1..7: dy = -2 reads 0..7 of them, dy = +2 reads 1..8
                frac_vfilter<0, 9>(wv, vLo[0], vHi[0], o9);
#pragma unroll
                for (int r = 0; r < 8; ++r) d[r] = o9[r];
                emit(0, d);
#pragma unroll
                for (int r = 0; r < 8; ++r) d[r] = o9[r + 1];
                emit(2, d);
                copy_rows();
                emit(1, d);
            } else {
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    if (j == 1 && cy == 0) copy_rows();      // dy = 0: the only integer phase a quarter-pel row can have
                    else {
                        uint32_t ws[8];                      // column moved up by vOff rows, so that output r starts at row r
#pragma unroll
                        for (int q = 0; q < 7; ++q) ws[q] = __funnelshift_r(wv[q], wv[q + 1], 16 * vOff[j]);
                        ws[7] = wv[7] >> (16 * vOff[j]);
                        frac_vfilter<0, 8>(ws, vLo[j], vHi[j], d);
                    }
                    emit(j, d);
                }
            }
        }
        if (MODE != kFracSad) {
            __syncwarp();
            const int i = lane & 7;
#pragma unroll
            for (int pass = 0; pass < (HALF ? 3 : 2); ++pass) {
                const int k = pass * 4 + (lane >> 3);
                const int g = HALF ? k : k + (k >= 4);         // quarter stage: the centre (4) is not evaluated
                int e[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) e[c] = g < 9 ? (int)S.t[g][c * 8 + i] : 0;
                uint32_t tot;
                if (MODE == kFracHad8) {
                    hadamard_inplace<8>(e);
                    uint32_t s = 0;
#pragma unroll
                    for (int c = 0; c < 8; ++c) s = __sad(e[c], 0, s);
                    s += __shfl_xor_sync(0xFFFFFFFFu, s, 1); s += __shfl_xor_sync(0xFFFFFFFFu, s, 2); s += __shfl_xor_sync(0xFFFFFFFFu, s, 4);
                    tot = (s + 2) >> 2;                                     // xCalcHADs8x8 rounding
                } else {
                    hadamard_inplace<4>(e); hadamard_inplace<4>(e + 4);
                    uint32_t sa = 0, sb = 0;
#pragma unroll
                    for (int c = 0; c < 4; ++c) { sa = __sad(e[c], 0, sa); sb = __sad(e[c + 4], 0, sb); }
                    sa += __shfl_xor_sync(0xFFFFFFFFu, sa, 1); sa += __shfl_xor_sync(0xFFFFFFFFu, sa, 2);
                    sb += __shfl_xor_sync(0xFFFFFFFFu, sb, 1); sb += __shfl_xor_sync(0xFFFFFFFFu, sb, 2);
                    uint32_t bl = ((sa + 1) >> 1) + ((sb + 1) >> 1);       // xCalcHADs4x4 rounding, per 4x4 block
                    bl += __shfl_xor_sync(0xFFFFFFFFu, bl, 4);
                    tot = bl;
                }
                acc[pass] += tot;
            }
        }
        __syncwarp();
        if (!more) break;
        t = tn; tx = ntx; ty = nty;
    }
}

// cost of the MV at quarter-pel position (qx, qy) against the predictor: TComRdCost::getCost(x, y) with any cost scale
__device__ __forceinline__ uint32_t frac_mv_cost(uint32_t lambda, int qx, int qy, int predx, int predy) {
    return (uint32_t)(lambda * (mv_bits(qx - predx) + mv_bits(qy - predy))) >> 16;
}

template <int MODE, bool COOP>
__device__ __forceinline__ void frac_refine_pu(const FracParams& p, const FracPu& P, FracScratch& S, const int lane, const int n, uint32_t* red) {
    constexpr bool coop = COOP;
    const int warp = threadIdx.x >> 5;
    const int slot = p.slots ? p.slots[n] : n;
    // candidate of this lane in the 3x3 grid (lanes 0..8): row-major over (dy index, dx index)
    const int gi = lane % 3, gj = (lane / 3) % 3;
    // grid position -> index in the reference's candidate tables (TEncSearch.cpp:51-75); 4 bits each, grid position 0 lowest
    const unsigned long long lutHalf = 0x827403615ull, lutQter = 0x827605413ull;
    int cx = 0, cy = 0;                                   // stage centre, quarter-pel offset from the integer MV
    uint32_t bestCost = 0, bestMvc = 0;
#pragma unroll 1
    for (int stage = 0; stage < 2; ++stage) {
        const int step = stage == 0 ? 2 : 1;
        uint32_t acc[3];
        if (stage == 0) frac_eval<MODE, true, COOP>(p, P, S, lane, 0, 0, acc);
        else frac_eval<MODE, false, COOP>(p, P, S, lane, cx, cy, acc);
        if (coop) {                                        // CTA-uniform: the four warps hold partial sums over their tiles
            uint32_t* r = red + stage * (kFracWarps * 3 * 32);
#pragma unroll
            for (int k = 0; k < 3; ++k) r[(warp * 3 + k) * 32 + lane] = acc[k];
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 3; ++k) acc[k] = r[k * 32 + lane] + r[(3 + k) * 32 + lane] + r[(6 + k) * 32 + lane] + r[(9 + k) * 32 + lane];
        }
        // where frac_eval left the distortion of this lane's candidate (see its header)
        const int gq = (lane < 9 ? lane : 0), kq = (MODE != kFracSad && stage == 1) ? gq - (gq > 4) : gq;
        const int srcLane = (MODE == kFracSad ? gi : (kq & 3)) * 8, srcReg = MODE == kFracSad ? gj : (kq >> 2);
        const uint32_t v0 = __shfl_sync(0xFFFFFFFFu, acc[0], srcLane & 31), v1 = __shfl_sync(0xFFFFFFFFu, acc[1], srcLane & 31),
                       v2 = __shfl_sync(0xFFFFFFFFu, acc[2], srcLane & 31);
        uint32_t dist = srcReg == 0 ? v0 : (srcReg == 1 ? v1 : v2);
        if (MODE != kFracSad && stage == 1 && lane == 4) dist = bestCost - bestMvc;     // the half-pel winner, evaluated in stage 0
        const int qx = 4 * P.mvx + cx + (gi - 1) * step, qy = 4 * P.mvy + cy + (gj - 1) * step;
        const uint32_t mvc = frac_mv_cost(p.lambda, qx, qy, P.predx, P.predy);
        const uint32_t cost = dist + mvc;
        const uint32_t tIdx = (uint32_t)((stage == 0 ? lutHalf : lutQter) >> (4 * (lane < 9 ? lane : 0))) & 15u;
        unsigned long long key = lane < 9 ? (((unsigned long long)cost << 8) | tIdx) : ~0ull;
        if (p.cand && lane < 9 && (!coop || warp == 0)) p.cand[(size_t)slot * 18 + stage * 9 + tIdx] = cost;
        unsigned long long m = key;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(0xFFFFFFFFu, m, o); m = t < m ? t : m; }
        const int win = __ffs(__ballot_sync(0xFFFFFFFFu, key == m)) - 1;
        cx += ((win % 3) - 1) * step; cy += ((win / 3) - 1) * step;
        bestCost = __shfl_sync(0xFFFFFFFFu, cost, win);
        bestMvc = __shfl_sync(0xFFFFFFFFu, mvc, win);
    }
    if (lane == 0 && (!coop || warp == 0)) p.out[slot] = make_int4(4 * P.mvx + cx, 4 * P.mvy + cy, (int)bestCost, (int)(bestCost - bestMvc));
}

template <bool COOP>
__device__ __forceinline__ void frac_dispatch(const FracParams& p, FracScratch& S, const int lane, const int n, uint32_t* red) {
    const FracPu P = p.pus[n];
    if (!p.useHad) frac_refine_pu<kFracSad, COOP>(p, P, S, lane, n, red);
    else if (((P.w | P.h) & 7) == 0) frac_refine_pu<kFracHad8, COOP>(p, P, S, lane, n, red);
    else frac_refine_pu<kFracHad4, COOP>(p, P, S, lane, n, red);
}

// Throughput form: every warp refines PUs on its own (nBig is 0).
__global__ void __launch_bounds__(kFracThreads) me_frac_kernel(const FracParams p) {
    __shared__ FracScratch scratch[kFracWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int stride = (int)gridDim.x * kFracWarps;
    for (int n = (int)blockIdx.x * kFracWarps + warp; n < p.npus; n += stride) frac_dispatch<false>(p, scratch[warp], lane, n, nullptr);
}

// Latency form for small batches (a band of a frame on one of several GPUs, a single PU from the encoder): the first nBig PUs
// (kFracCoopTiles tiles or more) get a CTA each, whose four warps share the tiles and add up their partial sums per stage.
__global__ void __launch_bounds__(kFracThreads) me_frac_coop_kernel(const FracParams p) {
    __shared__ FracScratch scratch[kFracWarps];
    __shared__ uint32_t red[2 * kFracWarps * 3 * 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if ((int)blockIdx.x < p.nBig) {
        frac_dispatch<true>(p, scratch[warp], lane, blockIdx.x, red);
        return;
    }
    const int stride = ((int)gridDim.x - p.nBig) * kFracWarps;
    for (int n = p.nBig + ((int)blockIdx.x - p.nBig) * kFracWarps + warp; n < p.npus; n += stride) frac_dispatch<false>(p, scratch[warp], lane, n, nullptr);
}

// PU list of a whole frame from the winners of the preceding integer search: PU n = (job, partition), ordered by partition
// area, large to small (order[] from the host), so that the round-robin over warps stays balanced.
__global__ void me_frac_build_kernel(const int4* jobs, const int32_t* X, const int32_t* Y, const int2* preds, const int* order, int njobs,
                                     FracPu* pus, int* slots) {   // order[] is by tile count, so the cooperative PUs come first
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= njobs * HMME_NPARTS) return;
    const int part = order[n / njobs], job = n - (n / njobs) * njobs;
    const int4 jb = jobs[job];
    const PartRect r = part_rect(part);
    const int slot = job * HMME_NPARTS + part;
    const int2 pr = preds ? preds[job] : make_int2(0, 0);
    pus[n] = FracPu{jb.x + r.x, jb.y + r.y, r.w, r.h, X[slot], Y[slot], pr.x, pr.y};
    slots[n] = slot;
}

// ---------------------------------------------------------------------------------------------------------------------------
// Group form of the refinement (the throughput kernel): a warp is four independent TILE PROCESSORS of eight lanes ("octets").
// An octet works on one 8x8 tile at a time, all roles inside the octet:
//   fetch / H step : lane r holds patch rows 2r, 2r+1 (16 bytes each, in registers -- the patch never goes through shared memory)
//                    and filters them into the three horizontal planes, stored as packed row pairs
//   V step         : lane c = column c, one plane per pass (3 passes), all 32 lanes busy
//   Hadamard pass  : lane i = coefficient row i, one candidate per pass (9 or 8 passes), all 32 lanes busy
// What the four octets work on depends on the PU size: a PU of four tiles or more has all four (tiles t = octet, octet + 4, ...),
// PUs of two or three tiles two octets each (two PUs per warp), one-tile PUs one octet each (four PUs per warp): small PUs no longer
// leave lanes idle, and their set-up and their decision logic run four at a time.  Everything a lane needs about "its" PU is lane
// state, so the three cases are one code path; the PUs of a group share the distortion mode (the host orders the list that way).
// With one tile per warp and iteration (frac_eval) the V step has 24 busy lanes, the half-pel Hadamard pass 72 tasks on 96 lanes.
struct __align__(16) FracScratch4 {
    int16_t cur[4][72];         // [octet][row * 8 + column], zero outside the PU (+8: octets 36 words apart)
    uint32_t h[3][4][104];      // [dx index][octet][column * 12 + row pair]: rows 2q | 2q+1 << 16 (12: conflict-free 128-bit column loads, +8 per octet)
    int16_t t[9][4][72];        // [candidate][octet][column * 8 + coefficient row] (+8)
};
// ceil(65536 / n): t / n == (t * rcp) >> 16 for t < 128, n <= 8
__constant__ int kTileRcp[9] = {0, 65536, 32768, 21846, 16384, 13108, 10923, 9363, 8192};

struct FracOct {                // lane state of the group form: the PU this lane's octet works on
    const uint8_t* refPu;       // patch origin: PU position + integer MV - (4, 4)
    long long curOff;           // element offset of the PU's sample (0, 0) in the current plane
    int w, h, ntxT, nT, rcp;
    int t0, tstep;              // this octet's tiles: t0, t0 + tstep, ...
    bool valid;                 // the octet has a PU at all (ragged last group of a segment)
};

template <int MODE, bool HALF>
__device__ __forceinline__ void frac_eval_oct(const FracParams& p, const FracOct& O, FracScratch4& S, const int lane, const int iters,
                                              const int cx, const int cy, uint32_t (&acc)[9]) {
    constexpr int step = HALF ? 2 : 1;
    const int s = lane >> 3, r = lane & 7;
    int hOff[3], vOff[3];
    uint32_t cLo[3], cHi[3], vLo[3], vHi[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int dx = cx + (i - 1) * step, dy = cy + (i - 1) * step;
        hOff[i] = 1 + (dx >> 2); vOff[i] = 1 + (dy >> 2);
        cLo[i] = kLumaPack[dx & 3][0]; cHi[i] = kLumaPack[dx & 3][1];
        vLo[i] = kLumaPack[dy & 3][0]; vHi[i] = kLumaPack[dy & 3][1];
    }
#pragma unroll
    for (int g = 0; g < 9; ++g) acc[g] = 0;
    const bool allInt = __all_sync(0xFFFFFFFFu, cy == 0);    // every octet's middle candidate row is a rounding copy (always in the half-pel stage)

    // the next tile's samples travel in registers while the current one is worked on; an octet without a tile gets zeros, which
    // flow through as zero planes, zero differences and zero sums: no masks further down
    uint32_t nW[2][4], nC[4];
    int ntw = 8, nth = 8;
    auto fetch = [&](const int t) {
#pragma unroll
        for (int k = 0; k < 2; ++k)
#pragma unroll
            for (int q = 0; q < 4; ++q) nW[k][q] = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) nC[q] = 0;
        if (O.valid && t < O.nT) {
            const int ty = ((t * O.rcp) >> 16) * 8, tx = t * 8 - ty * O.ntxT;
            ntw = min(8, O.w - tx); nth = min(8, O.h - ty);
#pragma unroll
            for (int k = 0; k < 2; ++k) {                    // 16 patch bytes of a row, any alignment: five aligned words, four funnel shifts
                const uint8_t* g = O.refPu + (long long)(ty + 2 * r + k) * p.refPitch + tx;
                const uint32_t sh = 8u * (uint32_t)((uintptr_t)g & 3);
                const uint32_t* qa = reinterpret_cast<const uint32_t*>((uintptr_t)g & ~(uintptr_t)3);
                const uint32_t a0 = __ldg(qa), a1 = __ldg(qa + 1), a2 = __ldg(qa + 2), a3 = __ldg(qa + 3), a4 = __ldg(qa + 4);
                nW[k][0] = __funnelshift_r(a0, a1, sh); nW[k][1] = __funnelshift_r(a1, a2, sh);
                nW[k][2] = __funnelshift_r(a2, a3, sh); nW[k][3] = __funnelshift_r(a3, a4, sh);
            }
            if (r < nth) {                                   // row r of the current tile as four pairs of 16-bit samples
                const long long o = O.curOff + (long long)(ty + r) * p.curPitch + tx;
                if (p.curBytes == 1) {
                    const uint8_t* g = static_cast<const uint8_t*>(p.cur) + o;
                    const uint32_t sh = 8u * (uint32_t)((uintptr_t)g & 3);
                    const uint32_t* qa = reinterpret_cast<const uint32_t*>((uintptr_t)g & ~(uintptr_t)3);
                    const uint32_t a0 = __ldg(qa), a1 = __ldg(qa + 1), a2 = __ldg(qa + 2);
                    const uint32_t b0 = __funnelshift_r(a0, a1, sh), b1 = __funnelshift_r(a1, a2, sh);
                    nC[0] = __byte_perm(b0, 0, 0x4140); nC[1] = __byte_perm(b0, 0, 0x4342);
                    nC[2] = __byte_perm(b1, 0, 0x4140); nC[3] = __byte_perm(b1, 0, 0x4342);
                } else {
                    const int16_t* g = static_cast<const int16_t*>(p.cur) + o;
#pragma unroll
                    for (int q = 0; q < 4; ++q) nC[q] = ((uint32_t)(uint16_t)g[2 * q]) | ((uint32_t)(uint16_t)g[2 * q + 1] << 16);
                }
                if (ntw < 8) { nC[2] = 0; nC[3] = 0; }       // tile widths are 4 or 8
            }
        }
    };
    int t = O.t0;
    fetch(t);
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        uint32_t W[2][4];
#pragma unroll
        for (int k = 0; k < 2; ++k)
#pragma unroll
            for (int q = 0; q < 4; ++q) W[k][q] = nW[k][q];
        *reinterpret_cast<uint4*>(&S.cur[s][r * 8]) = make_uint4(nC[0], nC[1], nC[2], nC[3]);
        const int tw = ntw, th = nth;
        t += O.tstep;
        if (it + 1 < iters) fetch(t);
        {   // H step: rows 2r and 2r+1, eight columns, three planes; a row pair leaves as one 32-bit word per column and plane
            if (HALF) {
                // dx = -2 and dx = +2 are the same half-sample row one column apart (nine windows serve both); dx = 0 is 64 * sample
                int f[2][9];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    uint32_t L[13];                          // L[o] = bytes o .. o+3 of the row
#pragma unroll
                    for (int o = 0; o < 13; ++o) L[o] = (o & 3) ? __funnelshift_r(W[k][o >> 2], W[k][(o >> 2) + 1], 8 * (o & 3)) : W[k][o >> 2];
#pragma unroll
                    for (int o = 0; o < 9; ++o) f[k][o] = dp4a_us(L[o + 4], cHi[0], dp4a_us(L[o], cLo[0], 0));
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    S.h[0][s][j * 12 + r] = __byte_perm(f[0][j], f[1][j], 0x5410);
                    S.h[2][s][j * 12 + r] = __byte_perm(f[0][j + 1], f[1][j + 1], 0x5410);
                    // sample j + 4 of both rows, times 64: bytes {row 2r, 0, row 2r+1, 0} << 6
                    const uint32_t lo = __byte_perm(W[0][1 + (j >> 2)], 0, 0x4440 | (j & 3)), hi = __byte_perm(W[1][1 + (j >> 2)], 0, 0x4044 | ((j & 3) << 8));
                    S.h[1][s][j * 12 + r] = (lo | hi) << 6;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    int out[2][8];
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        // the row moved left by hOff bytes (0 or 1; the patch has 16 bytes, windows reach byte 7 + 1 + 7), then fixed windows
                        const uint32_t sh = 8u * (uint32_t)hOff[i];
                        uint32_t V[4];
                        V[0] = __funnelshift_r(W[k][0], W[k][1], sh); V[1] = __funnelshift_r(W[k][1], W[k][2], sh);
                        V[2] = __funnelshift_r(W[k][2], W[k][3], sh); V[3] = W[k][3] >> sh;
                        uint32_t L[12];
#pragma unroll
                        for (int o = 0; o < 12; ++o) L[o] = (o & 3) ? __funnelshift_r(V[o >> 2], V[min((o >> 2) + 1, 3)], 8 * (o & 3)) : V[o >> 2];
#pragma unroll
                        for (int j = 0; j < 8; ++j) out[k][j] = dp4a_us(L[j + 4], cHi[i], dp4a_us(L[j], cLo[i], 0));
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) S.h[i][s][j * 12 + r] = __byte_perm(out[0][j], out[1][j], 0x5410);
                }
            }
        }
        __syncwarp();
        {   // V step: lane = column r of the octet's tile, one plane per pass
            int cu[8];
#pragma unroll
            for (int rr = 0; rr < 8; ++rr) cu[rr] = (int)S.cur[s][rr * 8 + r];
            const bool colDead = MODE != kFracHad8 && r >= tw;                // columns right of a 4-wide tile: zero plane -> zero prediction
#pragma unroll 1
            for (int di = 0; di < 3; ++di) {
                const uint4 ha = *reinterpret_cast<const uint4*>(&S.h[di][s][r * 12]), hb = *reinterpret_cast<const uint4*>(&S.h[di][s][r * 12 + 4]);
                uint32_t wv[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
                if (MODE != kFracHad8) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) wv[q] = colDead ? 0u : wv[q];
                }
                auto emit = [&](const int j, int (&d)[8]) {
#pragma unroll
                    for (int rr = 0; rr < 8; ++rr) d[rr] = cu[rr] - d[rr];
                    if (MODE != kFracHad8 && th < 8) { d[4] = 0; d[5] = 0; d[6] = 0; d[7] = 0; }
                    if (MODE == kFracSad) {
                        uint32_t sm = 0;
#pragma unroll
                        for (int rr = 0; rr < 8; ++rr) sm = __sad(d[rr], 0, sm);
                        acc[j * 3] += sm;                     // the three sums of a candidate row rotate with di (below); lanes are added at the end
                    } else {
                        if (MODE == kFracHad8) hadamard_inplace<8>(d);
                        else { hadamard_inplace<4>(d); hadamard_inplace<4>(d + 4); }
                        uint4 pk;
                        pk.x = __byte_perm(d[0], d[1], 0x5410); pk.y = __byte_perm(d[2], d[3], 0x5410);
                        pk.z = __byte_perm(d[4], d[5], 0x5410); pk.w = __byte_perm(d[6], d[7], 0x5410);
                        *reinterpret_cast<uint4*>(&S.t[j * 3 + di][s][r * 8]) = pk;
                    }
                };
                int d[8];
                auto copy_rows = [&]() {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        d[2 * q] = __vimin_s32_relu(((int)(int16_t)(wv[q + 2] & 0xFFFFu) + 32) >> 6, 255);
                        d[2 * q + 1] = __vimin_s32_relu((((int)wv[q + 2] >> 16) + 32) >> 6, 255);
                    }
                };
                if (HALF) {
                    int o9[9];                               // half-sample rows -1..7: dy = -2 reads 0..7 of them, dy = +2 reads 1..8
                    frac_vfilter<0, 9>(wv, vLo[0], vHi[0], o9);
#pragma unroll
                    for (int rr = 0; rr < 8; ++rr) d[rr] = o9[rr];
                    emit(0, d);
#pragma unroll
                    for (int rr = 0; rr < 8; ++rr) d[rr] = o9[rr + 1];
                    emit(2, d);
                    copy_rows();
                    emit(1, d);
                } else {
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        if (MODE != kFracSad && di == 1 && j == 1) continue;   // the half-pel winner: its distortion is known
                        if (j == 1 && allInt) copy_rows();   // dy = 0 in every octet (else the general filter does the same with the taps of phase 0)
                        else {
                            uint32_t ws[8];                  // column moved up by vOff rows, so that output r starts at row r
#pragma unroll
                            for (int q = 0; q < 7; ++q) ws[q] = __funnelshift_r(wv[q], wv[q + 1], 16 * vOff[j]);
                            ws[7] = wv[7] >> (16 * vOff[j]);
                            frac_vfilter<0, 8>(ws, vLo[j], vHi[j], d);
                        }
                        emit(j, d);
                    }
                }
                if (MODE == kFracSad) {                       // acc[3j + i] <- acc[3j + i + 1]: after the three planes every sum is back in place
#pragma unroll
                    for (int j = 0; j < 3; ++j) { const uint32_t a0 = acc[3 * j]; acc[3 * j] = acc[3 * j + 1]; acc[3 * j + 1] = acc[3 * j + 2]; acc[3 * j + 2] = a0; }
                }
            }
        }
        if (MODE != kFracSad) {
            __syncwarp();
            // Hadamard pass: lane = coefficient row r, one candidate per pass.  (This loop, the plane loop above and the iteration loop
            // stay rolled: unrolled, the bodies of the two stages exceed the 32 KB instruction cache next to the SM -- measured on the
            // first version of this form: 28 % of the stall samples "no instruction".)
#pragma unroll 1
            for (int j = 0; j < 3; ++j) {
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    if (!HALF && i == 1 && j == 1) continue;
                    int e[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) e[c] = (int)S.t[j * 3 + i][s][c * 8 + r];
                    uint32_t tot;
                    if (MODE == kFracHad8) {
                        hadamard_inplace<8>(e);
                        uint32_t sm = 0;
#pragma unroll
                        for (int c = 0; c < 8; ++c) sm = __sad(e[c], 0, sm);
                        sm += __shfl_xor_sync(0xFFFFFFFFu, sm, 1); sm += __shfl_xor_sync(0xFFFFFFFFu, sm, 2); sm += __shfl_xor_sync(0xFFFFFFFFu, sm, 4);
                        tot = (sm + 2) >> 2;                                    // xCalcHADs8x8 rounding
                    } else {
                        hadamard_inplace<4>(e); hadamard_inplace<4>(e + 4);
                        uint32_t sa = 0, sb = 0;
#pragma unroll
                        for (int c = 0; c < 4; ++c) { sa = __sad(e[c], 0, sa); sb = __sad(e[c + 4], 0, sb); }
                        sa += __shfl_xor_sync(0xFFFFFFFFu, sa, 1); sa += __shfl_xor_sync(0xFFFFFFFFu, sa, 2);
                        sb += __shfl_xor_sync(0xFFFFFFFFu, sb, 1); sb += __shfl_xor_sync(0xFFFFFFFFu, sb, 2);
                        uint32_t bl = ((sa + 1) >> 1) + ((sb + 1) >> 1);       // xCalcHADs4x4 rounding, per 4x4 block
                        bl += __shfl_xor_sync(0xFFFFFFFFu, bl, 4);
                        tot = bl;
                    }
                    acc[i] += tot;                            // identical on the eight lanes of the octet
                }
                // acc[g] <- acc[g + 3]: the sums of candidate row j sit in acc[0..2] during its pass; after three passes all are back in place
                const uint32_t a0 = acc[0], a1 = acc[1], a2 = acc[2];
#pragma unroll
                for (int g = 0; g < 6; ++g) acc[g] = acc[g + 3];
                acc[6] = a0; acc[7] = a1; acc[8] = a2;
            }
        }
        __syncwarp();
    }
    if (MODE == kFracSad) {                                   // column sums -> tile sums
#pragma unroll
        for (int g = 0; g < 9; ++g) {
            acc[g] += __shfl_xor_sync(0xFFFFFFFFu, acc[g], 1); acc[g] += __shfl_xor_sync(0xFFFFFFFFu, acc[g], 2); acc[g] += __shfl_xor_sync(0xFFFFFFFFu, acc[g], 4);
        }
    }
}

// One group: `nIn` PUs (first PU `first`), each on `m` = 4, 2 or 1 octets.
template <int MODE>
__device__ __forceinline__ void frac_group(const FracParams& p, FracScratch4& S, const int lane, const int first, const int nIn, const int m) {
    const int s = lane >> 3, k = lane & 7;
    const int puLocal = m == 4 ? 0 : (m == 2 ? s >> 1 : s);
    FracOct O;
    O.valid = puLocal < nIn;
    const int n = first + (O.valid ? puLocal : 0);
    const FracPu P = p.pus[n];
    const int slot = p.slots ? p.slots[n] : n;
    O.refPu = p.ref + (long long)(P.y + P.mvy - 4) * p.refPitch + (P.x + P.mvx - 4);
    O.curOff = (long long)P.y * p.curPitch + P.x;
    O.w = P.w; O.h = P.h;
    O.ntxT = (P.w + 7) >> 3; O.nT = O.ntxT * ((P.h + 7) >> 3); O.rcp = kTileRcp[O.ntxT];
    O.t0 = s & (m - 1); O.tstep = m;
    const int iters = __reduce_max_sync(0xFFFFFFFFu, O.valid ? (O.nT + m - 1) >> (m >> 1) : 0);   // ceil(nT / m), m = 1, 2, 4
    const bool writer = O.valid && O.t0 == 0;               // the PU's first octet reports
    // grid position g = 3 * (dy index) + (dx index) -> index in the reference's candidate tables (TEncSearch.cpp:51-75), 4 bits each
    const unsigned long long lutHalf = 0x827403615ull, lutQter = 0x827605413ull;
    int cx = 0, cy = 0;                                      // stage centre, quarter-pel offset from the integer MV
    uint32_t bestCost = 0, bestMvc = 0;
#pragma unroll 1
    for (int stage = 0; stage < 2; ++stage) {
        const int step = stage == 0 ? 2 : 1;
        uint32_t acc[9];
        if (stage == 0) frac_eval_oct<MODE, true>(p, O, S, lane, iters, 0, 0, acc);
        else frac_eval_oct<MODE, false>(p, O, S, lane, iters, cx, cy, acc);
        if (m >= 2) {                                        // warp-uniform: add up the octets of a PU
#pragma unroll
            for (int g = 0; g < 9; ++g) acc[g] += __shfl_xor_sync(0xFFFFFFFFu, acc[g], 8);
        }
        if (m == 4) {
#pragma unroll
            for (int g = 0; g < 9; ++g) acc[g] += __shfl_xor_sync(0xFFFFFFFFu, acc[g], 16);
        }
        // decision per PU: lane k of an octet costs candidate k, every lane candidate 8; key = cost | table index | grid position
        const unsigned long long lut = stage == 0 ? lutHalf : lutQter;
        uint32_t distK = acc[0];
#pragma unroll
        for (int g = 1; g < 8; ++g) distK = k == g ? acc[g] : distK;
        if (MODE != kFracSad && stage == 1 && k == 4) distK = bestCost - bestMvc;      // the half-pel winner, evaluated in stage 0
        const int gi = k % 3, gj = k / 3;
        const int bx = 4 * P.mvx + cx, by = 4 * P.mvy + cy;
        const uint32_t costK = distK + frac_mv_cost(p.lambda, bx + (gi - 1) * step, by + (gj - 1) * step, P.predx, P.predy);
        const uint32_t cost8 = acc[8] + frac_mv_cost(p.lambda, bx + step, by + step, P.predx, P.predy);
        const uint32_t tK = (uint32_t)(lut >> (4 * k)) & 15u, t8 = (uint32_t)(lut >> 32) & 15u;
        if (p.cand && writer) {
            p.cand[(size_t)slot * 18 + stage * 9 + tK] = costK;
            if (k == 0) p.cand[(size_t)slot * 18 + stage * 9 + t8] = cost8;
        }
        unsigned long long key = ((unsigned long long)costK << 8) | (tK << 4) | (uint32_t)k;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) { const unsigned long long q = __shfl_xor_sync(0xFFFFFFFFu, key, o); key = q < key ? q : key; }
        const unsigned long long key8 = ((unsigned long long)cost8 << 8) | (t8 << 4) | 8u;
        key = key8 < key ? key8 : key;
        const int win = (int)(key & 15u);
        cx += ((win % 3) - 1) * step; cy += ((win / 3) - 1) * step;
        bestCost = (uint32_t)(key >> 8);
        bestMvc = frac_mv_cost(p.lambda, 4 * P.mvx + cx, 4 * P.mvy + cy, P.predx, P.predy);
    }
    if (writer && k == 0) p.out[slot] = make_int4(4 * P.mvx + cx, 4 * P.mvy + cy, (int)bestCost, (int)(bestCost - bestMvc));
}

// Throughput form: every warp takes groups of the list (see FracParams::segPu), large PUs first.
__global__ void __launch_bounds__(kFracThreads) me_frac_group_kernel(const FracGroupParams p) {
    __shared__ FracScratch4 scratch[kFracWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int stride = (int)gridDim.x * kFracWarps, nGroups = p.segGrp[5];
    for (int G = (int)blockIdx.x * kFracWarps + warp; G < nGroups; G += stride) {
        int seg = 0;
#pragma unroll
        for (int q = 1; q < 5; ++q) seg += G >= p.segGrp[q];
        const int per = seg == 0 ? 1 : (seg <= 2 ? 2 : 4);
        const int first = p.segPu[seg] + (G - p.segGrp[seg]) * per;
        const int nIn = min(per, p.segPu[seg + 1] - first);
        bool had8 = seg == 1 || seg == 3;
        if (seg == 0) had8 = ((p.pus[first].w | p.pus[first].h) & 7) == 0;
        if (!p.useHad) frac_group<kFracSad>(p, scratch[warp], lane, first, nIn, 4 / per);
        else if (had8) frac_group<kFracHad8>(p, scratch[warp], lane, first, nIn, 4 / per);
        else frac_group<kFracHad4>(p, scratch[warp], lane, first, nIn, 4 / per);
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Distortion of the motion-compensated uni-prediction of a PU at a given QUARTER-PEL MV: the arithmetic
// of TEncSearch::xGetTemplateCost (TEncSearch.cpp:3634-3674: xPredInterBlk + SAD, the AMVP candidate check) and of the uni-directional
// candidates of xMergeEstimation / xGetInterPredictionError (:2814-2836, Hadamard).  Same interpolation and distortion code as the
// refinement above with a single candidate, so one warp works on four 8x8 tiles at a time: lane = (tile, column) in the V step and
// (tile, coefficient row) in the Hadamard pass.
struct McParams {
    const void* cur; const uint8_t* ref; const uint8_t* ref1;     // ref1: second reference plane of bi-directional PUs
    long long curPitch, refPitch, ref1Pitch;
    int curBytes;
    const int* pus;            // uni: {x, y, w, h, mvx, mvy}; bi: {x, y, w, h, mv0x, mv0y, mv1x, mv1y} (quarter pel, clipped)
    int npus;
    int useHad;
    uint32_t* out;
};

template <bool BI>
struct __align__(16) McScratch {
    uint32_t ref[BI ? 2 : 1][4][16][4];
    int16_t cur[4][8][8];
    int16_t h[BI ? 2 : 1][4][8][24];
    int16_t t[4][72];
};

// Bi-directional PUs (TComPrediction::xPredInterBi, TComPrediction.cpp:603-651): each list gives the 14-bit intermediate
// q = (sum_k C[fy][k] * h(r + k - 3)) >> 6 (xPredInterBlk with bi = true; its -8192 offsets cancel against addAvg's), the prediction
// is clip255((q0 + q1 + 64) >> 7) (TComYuv::addAvg, TComYuv.cpp:352-410).
template <int MODE, bool BI>
__device__ __forceinline__ uint32_t mc_cost_pu(const McParams& p, const int n, McScratch<BI>& S, const int lane) {
    constexpr int NL = BI ? 2 : 1;
    const int* pu = p.pus + (size_t)n * (BI ? 8 : 6);
    const int Px = pu[0], Py = pu[1], Pw = pu[2], Ph = pu[3];
    uint32_t cLo[NL], cHi[NL], vLo[NL], vHi[NL];
    int fyL[NL];
    const uint8_t* refPu[NL];
    long long pitchL[NL];
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        const int mx = pu[4 + 2 * l], my = pu[5 + 2 * l];
        cLo[l] = kLumaPack[mx & 3][0]; cHi[l] = kLumaPack[mx & 3][1];
        vLo[l] = kLumaPack[my & 3][0]; vHi[l] = kLumaPack[my & 3][1];
        fyL[l] = my & 3;
        pitchL[l] = l ? p.ref1Pitch : p.refPitch;
        // patch origin: the integer part of the MV is folded in, so taps of output 0 always start at patch index 1
        refPu[l] = (l ? p.ref1 : p.ref) + (long long)(Py + (my >> 2) - 4) * pitchL[l] + (Px + (mx >> 2) - 4);
    }
    const int ntxT = (Pw + 7) >> 3, nT = ntxT * ((Ph + 7) >> 3);
    const int rcp = (65536 + ntxT - 1) / ntxT;
    const int rowL = lane >> 1, halfL = lane & 1, curR = lane >> 2, curC = (lane & 3) * 2;
    uint32_t acc = 0;
    // Patches and current tiles of a group of up to four tiles travel global -> registers -> shared memory; the registers of group
    // g + 4 are loaded right after group g has been handed to shared memory, so their latency hides behind the filtering of group g.
    // A lane's 8 patch bytes (any alignment) come from three aligned 32-bit words and two funnel shifts (the planes carry slack bytes
    // behind the last row, see hmme_plane); its two current samples the same way from two aligned words.  (Measured and dropped: carrying
    // the pipeline across PUs -- next PU's record and first group fetched under the current PU -- costs 80 registers instead of 64
    // and is slower, 0.27 vs 0.23 ms for the frame's 284 640 PUs: occupancy hides the inter-PU latency better.)
    uint32_t pw0[NL][4], pw1[NL][4], pcur[4];
    auto fetch = [&](const int g0) {
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) {
            const int t = g0 + s4;
#pragma unroll
            for (int l = 0; l < NL; ++l) { pw0[l][s4] = 0; pw1[l][s4] = 0; }
            pcur[s4] = 0;
            if (t >= nT) continue;                             // warp-uniform: tiles past the last one are masked in the V step
            const int ty = ((t * rcp) >> 16) * 8, tx = t * 8 - ty * ntxT;
            const int tw = min(8, Pw - tx), th = min(8, Ph - ty);
#pragma unroll
            for (int l = 0; l < NL; ++l) {
                const uint8_t* q = refPu[l] + (long long)(ty + rowL) * pitchL[l] + tx + halfL * 8;
                const uint32_t sh = 8u * (uint32_t)((uintptr_t)q & 3);
                const uint32_t* qa = reinterpret_cast<const uint32_t*>((uintptr_t)q & ~(uintptr_t)3);
                const uint32_t a0 = __ldg(qa), a1 = __ldg(qa + 1), a2 = __ldg(qa + 2);
                pw0[l][s4] = __funnelshift_r(a0, a1, sh);
                pw1[l][s4] = __funnelshift_r(a1, a2, sh);
            }
            if (curR < th && curC < tw) {                      // two neighbouring samples, any alignment (PU x is arbitrary here)
                const long long o = (long long)(Py + ty + curR) * p.curPitch + Px + tx + curC;
                const uintptr_t b = (uintptr_t)p.cur + (uintptr_t)(o * p.curBytes);
                const uint32_t* qa = reinterpret_cast<const uint32_t*>(b & ~(uintptr_t)3);
                const uint32_t v = __funnelshift_r(__ldg(qa), __ldg(qa + 1), 8u * (uint32_t)(b & 3));
                pcur[s4] = p.curBytes == 1 ? (v & 0xFFu) | ((v & 0xFF00u) << 8) : v;
            }
        }
    };
    fetch(0);
    for (int g = 0; g < nT; g += 4) {
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) {
            if (g + s4 >= nT) break;
#pragma unroll
            for (int l = 0; l < NL; ++l) *reinterpret_cast<uint2*>(&S.ref[l][s4][rowL][halfL * 2]) = make_uint2(pw0[l][s4], pw1[l][s4]);
            S.cur[s4][curC][curR] = (int16_t)(pcur[s4] & 0xFFFFu);
            S.cur[s4][curC + 1][curR] = (int16_t)(pcur[s4] >> 16);
        }
        if (g + 4 < nT) fetch(g + 4);
        __syncwarp();
#pragma unroll
        for (int l = 0; l < NL; ++l)
#pragma unroll
            for (int s4 = 0; s4 < 4; ++s4) {                  // H step, one plane per tile and list
                if (g + s4 >= nT) break;
                const uint32_t W0 = S.ref[l][s4][rowL][halfL], W1 = S.ref[l][s4][rowL][halfL + 1], W2 = S.ref[l][s4][rowL][halfL + 2];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int o8 = 8 * (j + 1);
                    const int out = dp4a_us(__funnelshift_rc(W1, W2, o8), cHi[l], dp4a_us(__funnelshift_rc(W0, W1, o8), cLo[l], 0));
                    S.h[l][s4][halfL * 4 + j][rowL] = (int16_t)out;
                }
            }
        __syncwarp();
        {   // V step: lane = (tile, column)
            const int s4 = lane >> 3, c = lane & 7;
            const int t = g + s4;
            const int ty = ((t * rcp) >> 16) * 8, tx = t * 8 - ty * ntxT;
            const int tw = min(8, Pw - tx), th = min(8, Ph - ty);
            int d[8];
            if (!BI) {
                const uint4 ha = *reinterpret_cast<const uint4*>(&S.h[0][s4][c][0]), hb = *reinterpret_cast<const uint4*>(&S.h[0][s4][c][8]);
                const uint32_t wv[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
                if (fyL[0] == 0) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        d[2 * q] = __vimin_s32_relu(((int)(int16_t)(wv[q + 2] & 0xFFFFu) + 32) >> 6, 255);
                        d[2 * q + 1] = __vimin_s32_relu((((int)wv[q + 2] >> 16) + 32) >> 6, 255);
                    }
                } else frac_vfilter<1, 8>(wv, vLo[0], vHi[0], d);
            } else {
                int qs[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) qs[r] = 64;                     // addAvg rounding
#pragma unroll
                for (int l = 0; l < NL; ++l) {
                    const uint4 ha = *reinterpret_cast<const uint4*>(&S.h[l][s4][c][0]), hb = *reinterpret_cast<const uint4*>(&S.h[l][s4][c][8]);
                    const uint32_t wv[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
                    int q8[8];
                    if (fyL[l] == 0) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) { q8[2 * q] = (int)(int16_t)(wv[q + 2] & 0xFFFFu); q8[2 * q + 1] = (int)wv[q + 2] >> 16; }
                    } else frac_vfilter<1, 8, true>(wv, vLo[l], vHi[l], q8);
#pragma unroll
                    for (int r = 0; r < 8; ++r) qs[r] += q8[r];
                }
#pragma unroll
                for (int r = 0; r < 8; ++r) d[r] = __vimin_s32_relu(qs[r] >> 7, 255);
            }
            const uint4 cc = *reinterpret_cast<const uint4*>(&S.cur[s4][c][0]);
            const uint32_t wc[4] = {cc.x, cc.y, cc.z, cc.w};
            const bool inside = t < nT && c < tw;              // columns right of a 4-wide tile and tiles past the last one contribute nothing
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int cu0 = (int)(int16_t)(wc[q] & 0xFFFFu), cu1 = (int)wc[q] >> 16;
                d[2 * q] = inside ? cu0 - d[2 * q] : 0;
                d[2 * q + 1] = inside ? cu1 - d[2 * q + 1] : 0;
            }
            if (MODE != kFracHad8 && th < 8) { d[4] = 0; d[5] = 0; d[6] = 0; d[7] = 0; }
            if (MODE == kFracSad) {
                uint32_t sm = 0;
#pragma unroll
                for (int r = 0; r < 8; ++r) sm = __sad(d[r], 0, sm);
                acc += sm;                                       // summed over all lanes at the end
            } else {
                if (MODE == kFracHad8) hadamard_inplace<8>(d);
                else { hadamard_inplace<4>(d); hadamard_inplace<4>(d + 4); }
                uint4 pk;
                pk.x = __byte_perm(d[0], d[1], 0x5410); pk.y = __byte_perm(d[2], d[3], 0x5410);
                pk.z = __byte_perm(d[4], d[5], 0x5410); pk.w = __byte_perm(d[6], d[7], 0x5410);
                *reinterpret_cast<uint4*>(&S.t[s4][c * 8]) = pk;
            }
        }
        if (MODE != kFracSad) {
            __syncwarp();
            const int s4 = lane >> 3, i = lane & 7;
            int e[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) e[c] = (int)S.t[s4][c * 8 + i];
            if (MODE == kFracHad8) {
                hadamard_inplace<8>(e);
                uint32_t sm = 0;
#pragma unroll
                for (int c = 0; c < 8; ++c) sm = __sad(e[c], 0, sm);
                sm += __shfl_xor_sync(0xFFFFFFFFu, sm, 1); sm += __shfl_xor_sync(0xFFFFFFFFu, sm, 2); sm += __shfl_xor_sync(0xFFFFFFFFu, sm, 4);
                if (i == 0) acc += (sm + 2) >> 2;
            } else {
                hadamard_inplace<4>(e); hadamard_inplace<4>(e + 4);
                uint32_t sa = 0, sb = 0;
#pragma unroll
                for (int c = 0; c < 4; ++c) { sa = __sad(e[c], 0, sa); sb = __sad(e[c + 4], 0, sb); }
                sa += __shfl_xor_sync(0xFFFFFFFFu, sa, 1); sa += __shfl_xor_sync(0xFFFFFFFFu, sa, 2);
                sb += __shfl_xor_sync(0xFFFFFFFFu, sb, 1); sb += __shfl_xor_sync(0xFFFFFFFFu, sb, 2);
                if ((i & 3) == 0) acc += ((sa + 1) >> 1) + ((sb + 1) >> 1);
            }
        }
        __syncwarp();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    return acc;
}

template <bool BI>
__global__ void __launch_bounds__(kFracThreads) me_mc_cost_kernel(const McParams p) {
    __shared__ McScratch<BI> scratch[kFracWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    McScratch<BI>& S = scratch[warp];
    const int stride = (int)gridDim.x * kFracWarps;
    for (int n = (int)blockIdx.x * kFracWarps + warp; n < p.npus; n += stride) {
        const int w = p.pus[(size_t)n * (BI ? 8 : 6) + 2], h = p.pus[(size_t)n * (BI ? 8 : 6) + 3];
        uint32_t d;
        if (!p.useHad) d = mc_cost_pu<kFracSad, BI>(p, n, S, lane);
        else if (((w | h) & 7) == 0) d = mc_cost_pu<kFracHad8, BI>(p, n, S, lane);
        else d = mc_cost_pu<kFracHad4, BI>(p, n, S, lane);
        if (lane == 0) p.out[n] = d;
    }
}


// ---------------------------------------------------------------------------------------------------------------------------
// Group form of the distortion kernel: the same four tile processors per warp as me_frac_group_kernel, one candidate per PU.  With one PU
// per warp (me_mc_cost_kernel) the 320 one-tile and 128 two-tile partitions of a CTU leave three quarters / half of the warp idle in the
// V step and the Hadamard pass; here they run four / two at a time.  Filter phases are lane state (every octet has its own PU), so the
// integer phases go through the general filter with the taps {0,0,0,64,0,0,0,0} unless the whole warp has them.
struct McGroupParams : McParams {
    const int* slots;           // result index of PU n of the (segment-ordered) list
    int segPu[6], segGrp[6];    // as FracGroupParams
};

template <bool BI>
struct __align__(16) McScratch4 {
    int16_t cur[4][72];                  // [octet][row * 8 + column]
    uint32_t h[BI ? 2 : 1][4][104];      // [list][octet][column * 12 + row pair]
    int16_t t[4][72];                    // [octet][column * 8 + coefficient row]
};

template <int MODE, bool BI>
__device__ __forceinline__ void mc_group(const McGroupParams& p, McScratch4<BI>& S, const int lane, const int first, const int nIn, const int m) {
    constexpr int NL = BI ? 2 : 1;
    const int s = lane >> 3, r = lane & 7;
    const int puLocal = m == 4 ? 0 : (m == 2 ? s >> 1 : s);
    const bool valid = puLocal < nIn;
    const int n = first + (valid ? puLocal : 0);
    const int* pu = p.pus + (size_t)n * (BI ? 8 : 6);
    const int Px = pu[0], Py = pu[1], Pw = pu[2], Ph = pu[3];
    const int slot = p.slots ? p.slots[n] : n;
    uint32_t cLo[NL], cHi[NL], vLo[NL], vHi[NL];
    bool allInt[NL];
    const uint8_t* refPu[NL];
    long long pitchL[NL];
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        const int mx = pu[4 + 2 * l], my = pu[5 + 2 * l];
        cLo[l] = kLumaPack[mx & 3][0]; cHi[l] = kLumaPack[mx & 3][1];
        vLo[l] = kLumaPack[my & 3][0]; vHi[l] = kLumaPack[my & 3][1];
        allInt[l] = __all_sync(0xFFFFFFFFu, (my & 3) == 0);
        pitchL[l] = l ? p.ref1Pitch : p.refPitch;
        // patch origin: the integer part of the MV is folded in, so the taps of output 0 always start at patch index 1
        refPu[l] = (l ? p.ref1 : p.ref) + (long long)(Py + (my >> 2) - 4) * pitchL[l] + (Px + (mx >> 2) - 4);
    }
    const long long curOff = (long long)Py * p.curPitch + Px;
    const int ntxT = (Pw + 7) >> 3, nT = ntxT * ((Ph + 7) >> 3), rcp = kTileRcp[ntxT];
    const int t0 = s & (m - 1);
    const int iters = __reduce_max_sync(0xFFFFFFFFu, valid ? (nT + m - 1) >> (m >> 1) : 0);
    uint32_t acc = 0;

    uint32_t nW[NL][2][4], nC[4];
    int ntw = 8, nth = 8;
    auto fetch = [&](const int t) {
#pragma unroll
        for (int l = 0; l < NL; ++l)
#pragma unroll
            for (int k = 0; k < 2; ++k)
#pragma unroll
                for (int q = 0; q < 4; ++q) nW[l][k][q] = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) nC[q] = 0;
        if (valid && t < nT) {
            const int ty = ((t * rcp) >> 16) * 8, tx = t * 8 - ty * ntxT;
            ntw = min(8, Pw - tx); nth = min(8, Ph - ty);
#pragma unroll
            for (int l = 0; l < NL; ++l)
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const uint8_t* g = refPu[l] + (long long)(ty + 2 * r + k) * pitchL[l] + tx;
                    const uint32_t sh = 8u * (uint32_t)((uintptr_t)g & 3);
                    const uint32_t* qa = reinterpret_cast<const uint32_t*>((uintptr_t)g & ~(uintptr_t)3);
                    const uint32_t a0 = __ldg(qa), a1 = __ldg(qa + 1), a2 = __ldg(qa + 2), a3 = __ldg(qa + 3), a4 = __ldg(qa + 4);
                    nW[l][k][0] = __funnelshift_r(a0, a1, sh); nW[l][k][1] = __funnelshift_r(a1, a2, sh);
                    nW[l][k][2] = __funnelshift_r(a2, a3, sh); nW[l][k][3] = __funnelshift_r(a3, a4, sh);
                }
            if (r < nth) {
                const long long o = curOff + (long long)(ty + r) * p.curPitch + tx;
                if (p.curBytes == 1) {
                    const uint8_t* g = static_cast<const uint8_t*>(p.cur) + o;
                    const uint32_t sh = 8u * (uint32_t)((uintptr_t)g & 3);
                    const uint32_t* qa = reinterpret_cast<const uint32_t*>((uintptr_t)g & ~(uintptr_t)3);
                    const uint32_t a0 = __ldg(qa), a1 = __ldg(qa + 1), a2 = __ldg(qa + 2);
                    const uint32_t b0 = __funnelshift_r(a0, a1, sh), b1 = __funnelshift_r(a1, a2, sh);
                    nC[0] = __byte_perm(b0, 0, 0x4140); nC[1] = __byte_perm(b0, 0, 0x4342);
                    nC[2] = __byte_perm(b1, 0, 0x4140); nC[3] = __byte_perm(b1, 0, 0x4342);
                } else {
                    const int16_t* g = static_cast<const int16_t*>(p.cur) + o;
#pragma unroll
                    for (int q = 0; q < 4; ++q) nC[q] = ((uint32_t)(uint16_t)g[2 * q]) | ((uint32_t)(uint16_t)g[2 * q + 1] << 16);
                }
                if (ntw < 8) { nC[2] = 0; nC[3] = 0; }
            }
        }
    };
    int t = t0;
    fetch(t);
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        uint32_t W[NL][2][4];
#pragma unroll
        for (int l = 0; l < NL; ++l)
#pragma unroll
            for (int k = 0; k < 2; ++k)
#pragma unroll
                for (int q = 0; q < 4; ++q) W[l][k][q] = nW[l][k][q];
        *reinterpret_cast<uint4*>(&S.cur[s][r * 8]) = make_uint4(nC[0], nC[1], nC[2], nC[3]);
        const int tw = ntw, th = nth;
        t += m;
        if (it + 1 < iters) fetch(t);
#pragma unroll
        for (int l = 0; l < NL; ++l) {                       // H step: rows 2r, 2r+1 of the list's patch, output j from bytes j+1 .. j+8
            int out[2][8];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                uint32_t L[13];
#pragma unroll
                for (int o = 1; o < 13; ++o) L[o] = (o & 3) ? __funnelshift_r(W[l][k][o >> 2], W[l][k][(o >> 2) + 1], 8 * (o & 3)) : W[l][k][o >> 2];
#pragma unroll
                for (int j = 0; j < 8; ++j) out[k][j] = dp4a_us(L[j + 5], cHi[l], dp4a_us(L[j + 1], cLo[l], 0));
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) S.h[l][s][j * 12 + r] = __byte_perm(out[0][j], out[1][j], 0x5410);
        }
        __syncwarp();
        {   // V step: lane = column r of the octet's tile
            int d[8];
            const bool colDead = MODE != kFracHad8 && r >= tw;
            if (!BI) {
                const uint4 ha = *reinterpret_cast<const uint4*>(&S.h[0][s][r * 12]), hb = *reinterpret_cast<const uint4*>(&S.h[0][s][r * 12 + 4]);
                uint32_t wv[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
                if (MODE != kFracHad8) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) wv[q] = colDead ? 0u : wv[q];
                }
                if (allInt[0]) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        d[2 * q] = __vimin_s32_relu(((int)(int16_t)(wv[q + 2] & 0xFFFFu) + 32) >> 6, 255);
                        d[2 * q + 1] = __vimin_s32_relu((((int)wv[q + 2] >> 16) + 32) >> 6, 255);
                    }
                } else frac_vfilter<1, 8>(wv, vLo[0], vHi[0], d);
            } else {
                int qs[8];
#pragma unroll
                for (int rr = 0; rr < 8; ++rr) qs[rr] = 64;                 // addAvg rounding
#pragma unroll
                for (int l = 0; l < NL; ++l) {
                    const uint4 ha = *reinterpret_cast<const uint4*>(&S.h[l][s][r * 12]), hb = *reinterpret_cast<const uint4*>(&S.h[l][s][r * 12 + 4]);
                    uint32_t wv[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
                    if (MODE != kFracHad8) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) wv[q] = colDead ? 0u : wv[q];
                    }
                    int q8[8];
                    if (allInt[l]) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) { q8[2 * q] = (int)(int16_t)(wv[q + 2] & 0xFFFFu); q8[2 * q + 1] = (int)wv[q + 2] >> 16; }
                    } else frac_vfilter<1, 8, true>(wv, vLo[l], vHi[l], q8);
#pragma unroll
                    for (int rr = 0; rr < 8; ++rr) qs[rr] += q8[rr];
                }
#pragma unroll
                for (int rr = 0; rr < 8; ++rr) d[rr] = __vimin_s32_relu(qs[rr] >> 7, 255);
            }
#pragma unroll
            for (int rr = 0; rr < 8; ++rr) d[rr] = (int)S.cur[s][rr * 8 + r] - d[rr];
            if (MODE != kFracHad8 && th < 8) { d[4] = 0; d[5] = 0; d[6] = 0; d[7] = 0; }
            if (MODE == kFracSad) {
                uint32_t sm = 0;
#pragma unroll
                for (int rr = 0; rr < 8; ++rr) sm = __sad(d[rr], 0, sm);
                acc += sm;                                       // column sums: added over the octet at the end
            } else {
                if (MODE == kFracHad8) hadamard_inplace<8>(d);
                else { hadamard_inplace<4>(d); hadamard_inplace<4>(d + 4); }
                uint4 pk;
                pk.x = __byte_perm(d[0], d[1], 0x5410); pk.y = __byte_perm(d[2], d[3], 0x5410);
                pk.z = __byte_perm(d[4], d[5], 0x5410); pk.w = __byte_perm(d[6], d[7], 0x5410);
                *reinterpret_cast<uint4*>(&S.t[s][r * 8]) = pk;
            }
        }
        if (MODE != kFracSad) {
            __syncwarp();
            int e[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) e[c] = (int)S.t[s][c * 8 + r];
            if (MODE == kFracHad8) {
                hadamard_inplace<8>(e);
                uint32_t sm = 0;
#pragma unroll
                for (int c = 0; c < 8; ++c) sm = __sad(e[c], 0, sm);
                sm += __shfl_xor_sync(0xFFFFFFFFu, sm, 1); sm += __shfl_xor_sync(0xFFFFFFFFu, sm, 2); sm += __shfl_xor_sync(0xFFFFFFFFu, sm, 4);
                acc += (sm + 2) >> 2;                            // identical on the eight lanes of the octet
            } else {
                hadamard_inplace<4>(e); hadamard_inplace<4>(e + 4);
                uint32_t sa = 0, sb = 0;
#pragma unroll
                for (int c = 0; c < 4; ++c) { sa = __sad(e[c], 0, sa); sb = __sad(e[c + 4], 0, sb); }
                sa += __shfl_xor_sync(0xFFFFFFFFu, sa, 1); sa += __shfl_xor_sync(0xFFFFFFFFu, sa, 2);
                sb += __shfl_xor_sync(0xFFFFFFFFu, sb, 1); sb += __shfl_xor_sync(0xFFFFFFFFu, sb, 2);
                uint32_t bl = ((sa + 1) >> 1) + ((sb + 1) >> 1);
                bl += __shfl_xor_sync(0xFFFFFFFFu, bl, 4);
                acc += bl;
            }
        }
        __syncwarp();
    }
    if (MODE == kFracSad) { acc += __shfl_xor_sync(0xFFFFFFFFu, acc, 1); acc += __shfl_xor_sync(0xFFFFFFFFu, acc, 2); acc += __shfl_xor_sync(0xFFFFFFFFu, acc, 4); }
    if (m >= 2) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, 8);
    if (m == 4) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, 16);
    if (valid && t0 == 0 && r == 0) p.out[slot] = acc;
}

template <bool BI>
__global__ void __launch_bounds__(kFracThreads) me_mc_group_kernel(const McGroupParams p) {
    __shared__ McScratch4<BI> scratch[kFracWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int stride = (int)gridDim.x * kFracWarps, nGroups = p.segGrp[5];
    for (int G = (int)blockIdx.x * kFracWarps + warp; G < nGroups; G += stride) {
        int seg = 0;
#pragma unroll
        for (int q = 1; q < 5; ++q) seg += G >= p.segGrp[q];
        const int per = seg == 0 ? 1 : (seg <= 2 ? 2 : 4);
        const int first = p.segPu[seg] + (G - p.segGrp[seg]) * per;
        const int nIn = min(per, p.segPu[seg + 1] - first);
        bool had8 = seg == 1 || seg == 3;
        if (seg == 0) { const int* pu = p.pus + (size_t)first * (BI ? 8 : 6); had8 = ((pu[2] | pu[3]) & 7) == 0; }
        if (!p.useHad) mc_group<kFracSad, BI>(p, scratch[warp], lane, first, nIn, 4 / per);
        else if (had8) mc_group<kFracHad8, BI>(p, scratch[warp], lane, first, nIn, 4 / per);
        else mc_group<kFracHad4, BI>(p, scratch[warp], lane, first, nIn, 4 / per);
    }
}

}  // namespace hmme
