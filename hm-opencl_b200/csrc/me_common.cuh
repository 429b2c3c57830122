// me_common.cuh -- shared device/host helpers of the B200 motion-estimation kernels.
//
// Semantics follow HM-OpenCL's GPU path (normative pseudo-code in SURVEY.md App. A.2):
//   bit cost   : /root/reference/cl/sad.cl:374-398   (predictor-free, quarter-pel scale)
//   layout     : /root/reference/source/Lib/TLibCommon/TComDataCU.cpp:4676-6461 (593 partitions)
// Nothing here is copied from the reference; the layout is generated from HEVC geometry.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define HMME_NPARTS 593

namespace hmme {

// 64-bit arg-min key: (cost << 32) | (y*(2R+1)+x).  Unsigned min == the reference's strict-'<' update
// in y-outer/x-inner scan order (TEncOpenCL.cpp:312-313, sad.cl:400).
constexpr unsigned long long kNoWinner = 0xFFFFFFFFFFFFFFFFull;

struct PartRect { int x, y, w, h; };

// Partition index -> rectangle, derived from the quadtree: the 36 index groups are
//   CU8 : 2NxN(8x4) @0, Nx2N(4x8) @128, 2Nx2N @384
//   CU16: 2NxnU/nD (16x4,16x12) @256..319, nLx2N/nRx2N (4x16,12x16) @320..383, 2NxN @448, Nx2N @480, 2Nx2N @544
//   CU32: AMP @512..543, 2NxN @560, Nx2N @568, 2Nx2N @584
//   CU64: AMP @576..583, 2NxN @588, Nx2N @590, 2Nx2N @592
__host__ __device__ inline PartRect part_rect(int p) {
    PartRect r{0, 0, 0, 0};
    if (p < 128)      { r = {(p & 7) * 8, (p >> 3) * 4, 8, 4}; }
    else if (p < 256) { int k = p - 128; r = {(k & 15) * 4, (k >> 4) * 8, 4, 8}; }
    else if (p < 384) {                           // CU16 asymmetric: 8 groups of 16 (raster over 4x4 CUs)
        int g = (p - 256) >> 4, k = (p - 256) & 15, cx = (k & 3) * 16, cy = (k >> 2) * 16;
        switch (g) {
            case 0: r = {cx, cy, 16, 4}; break;        // 2NxnU part 0
            case 1: r = {cx, cy + 12, 16, 4}; break;   // 2NxnD part 1
            case 2: r = {cx, cy, 16, 12}; break;       // 2NxnD part 0
            case 3: r = {cx, cy + 4, 16, 12}; break;   // 2NxnU part 1
            case 4: r = {cx, cy, 4, 16}; break;        // nLx2N part 0
            case 5: r = {cx + 12, cy, 4, 16}; break;   // nRx2N part 1
            case 6: r = {cx, cy, 12, 16}; break;       // nRx2N part 0
            default: r = {cx + 4, cy, 12, 16}; break;  // nLx2N part 1
        }
    }
    else if (p < 448) { int k = p - 384; r = {(k & 7) * 8, (k >> 3) * 8, 8, 8}; }
    else if (p < 480) { int k = p - 448; r = {(k & 3) * 16, (k >> 2) * 8, 16, 8}; }
    else if (p < 512) { int k = p - 480; r = {(k & 7) * 8, (k >> 3) * 16, 8, 16}; }
    else if (p < 544) {                           // CU32 asymmetric: 8 groups of 4
        int g = (p - 512) >> 2, k = (p - 512) & 3, cx = (k & 1) * 32, cy = (k >> 1) * 32;
        switch (g) {
            case 0: r = {cx, cy, 32, 8}; break;
            case 1: r = {cx, cy + 24, 32, 8}; break;
            case 2: r = {cx, cy, 32, 24}; break;
            case 3: r = {cx, cy + 8, 32, 24}; break;
            case 4: r = {cx, cy, 8, 32}; break;
            case 5: r = {cx + 24, cy, 8, 32}; break;
            case 6: r = {cx, cy, 24, 32}; break;
            default: r = {cx + 8, cy, 24, 32}; break;
        }
    }
    else if (p < 560) { int k = p - 544; r = {(k & 3) * 16, (k >> 2) * 16, 16, 16}; }
    else if (p < 568) { int k = p - 560; r = {(k & 1) * 32, (k >> 1) * 16, 32, 16}; }
    else if (p < 576) { int k = p - 568; r = {(k & 3) * 16, (k >> 2) * 32, 16, 32}; }
    else if (p < 584) {
        switch (p - 576) {
            case 0: r = {0, 0, 64, 16}; break;
            case 1: r = {0, 48, 64, 16}; break;
            case 2: r = {0, 0, 64, 48}; break;
            case 3: r = {0, 16, 64, 48}; break;
            case 4: r = {0, 0, 16, 64}; break;
            case 5: r = {48, 0, 16, 64}; break;
            case 6: r = {0, 0, 48, 64}; break;
            default: r = {16, 0, 48, 64}; break;
        }
    }
    else if (p < 588) { int k = p - 584; r = {(k & 1) * 32, (k >> 1) * 32, 32, 32}; }
    else if (p < 590) { r = {0, (p - 588) * 32, 64, 32}; }
    else if (p < 592) { r = {(p - 590) * 32, 0, 32, 64}; }
    else              { r = {0, 0, 64, 64}; }
    return r;
}

// Closed form of TComDataCU::getIndexBlock (TComDataCU.cpp:3379-6464, 593-entry variant): (PartSize, depth, partIdx,
// z-order index of the CU in 4x4 units, CU width, CU height) -> partition index, or -1 for every combination the
// reference's 1786-line switch does not list.  PartSize: 0 2Nx2N, 1 2NxN, 2 Nx2N, 4 2NxnU, 5 2NxnD, 6 nLx2N, 7 nRx2N.
__host__ __device__ inline int index_block(int partSize, int depth, int partIdx, int zIdx, int cuW, int cuH) {
    if (depth < 0 || depth > 3 || partIdx < 0 || partIdx > 1 || zIdx < 0 || zIdx > 255) return -1;
    const int S = 64 >> depth;
    if (cuW != S || cuH != S) return -1;
    if (partSize < 0 || partSize > 7 || partSize == 3) return -1;           // NxN is never searched (TEncCu.cpp:500)
    if (partSize == 0 && partIdx != 0) return -1;
    if (partSize >= 4 && depth == 3) return -1;                             // no AMP for 8x8 CUs
    const int unitsPerCu = (S / 4) * (S / 4);
    if (zIdx % unitsPerCu) return -1;                                       // CU origin must be aligned to its size
    int x4 = 0, y4 = 0;                                                     // de-interleave the z-order index
    for (int b = 0; b < 4; ++b) { x4 |= ((zIdx >> (2 * b)) & 1) << b; y4 |= ((zIdx >> (2 * b + 1)) & 1) << b; }
    const int cx = 4 * x4, cy = 4 * y4, N = S / 2, q = S / 4;
    int x = cx, y = cy, w = S, h = S;
    switch (partSize) {
        case 1: h = N; y += partIdx * N; break;
        case 2: w = N; x += partIdx * N; break;
        case 4: h = partIdx ? S - q : q; y += partIdx ? q : 0; break;
        case 5: h = partIdx ? q : S - q; y += partIdx ? S - q : 0; break;
        case 6: w = partIdx ? S - q : q; x += partIdx ? q : 0; break;
        case 7: w = partIdx ? q : S - q; x += partIdx ? S - q : 0; break;
        default: break;
    }
    // locate (x, y, w, h) in the layout: each (w, h) pair belongs to at most two groups, told apart by the offset in the CU
    for (int p = 0; p < HMME_NPARTS; ++p) {
        const PartRect r = part_rect(p);
        if (r.x == x && r.y == y && r.w == w && r.h == h) return p;
    }
    return -1;
}

// Search-window placement of TEncSearch::xSetSearchRange (TEncSearch.cpp:3814-3830) with TComDataCU::clipMv
// (TComDataCU.cpp:2907-2920): the quarter-pel MV the window is centred on (the AMVP predictor, or the current MV for the
// bi-prediction refinement) is clipped to the picture + 8 samples (+ one CTU on the low side), +-range is applied in
// quarter-pel units, both corners are clipped again and brought to integer-pel with an arithmetic shift.
// The GPU path then searches [lt, lt + 2*range] in both axes and ignores rb.
__host__ __device__ inline void search_window(int predHorQpel, int predVerQpel, int range, int cuX, int cuY, int picW, int picH,
                                              int maxCuW, int maxCuH, int* ltx, int* lty, int* rbx, int* rby) {
    const int sh = 2, off = 8;
    const int horMax = (picW + off - cuX - 1) * 4, horMin = (-maxCuW - off - cuX + 1) * 4;
    const int verMax = (picH + off - cuY - 1) * 4, verMin = (-maxCuH - off - cuY + 1) * 4;
    auto clipH = [&](int v) { return v < horMin ? horMin : (v > horMax ? horMax : v); };
    auto clipV = [&](int v) { return v < verMin ? verMin : (v > verMax ? verMax : v); };
    auto s16 = [](int v) { return (int)(short)v; };                      // TComMv stores Short (TComMv.h:54-55)
    const int ph = clipH(s16(predHorQpel)), pv = clipV(s16(predVerQpel));
    *ltx = clipH(s16(ph - (range << sh))) >> sh;
    *lty = clipV(s16(pv - (range << sh))) >> sh;
    *rbx = clipH(s16(ph + (range << sh))) >> sh;
    *rby = clipV(s16(pv + (range << sh))) >> sh;
}

// bits(v) = 2*floor(log2 t) + 1 with t = (v <= 0) ? -2v+1 : 2v   (sad.cl:377-396)
__host__ __device__ inline uint32_t mv_bits(int v) {
    uint32_t t = (v <= 0) ? (uint32_t)(-2 * v) + 1u : (uint32_t)(2 * v);
#ifdef __CUDA_ARCH__
    return 2u * (31u - (uint32_t)__clz((int)t)) + 1u;
#else
    uint32_t len = 1;
    while (t != 1u) { t >>= 1; len += 2; }
    return len;
#endif
}

// (lambda * (bitsX + bitsY)) / 65536 in 32-bit unsigned arithmetic (sad.cl:398)
__host__ __device__ inline uint32_t mv_cost(uint32_t lambda, int mvx, int mvy) {
    return (uint32_t)(lambda * (mv_bits(mvx * 4) + mv_bits(mvy * 4))) >> 16;
}

}  // namespace hmme
