// me_u8_kernel.cuh -- the hot kernel: whole-CTU integer-pel full search on 8-bit content, sm_100a.
//
// What it replaces: the reference's (2R+1)^2 x {calcSAD_AMP, compareSAD} launch pairs per CTU
// (/root/reference/source/Lib/TLibEncoder/TEncOpenCL.cpp:312-333, /root/reference/cl/sad.cl:141-408).
// Here ONE launch covers every (CTU, reference) job of a frame; nothing like tempSad touches HBM.
//
// Decomposition
//   CTA (512 threads)      = one job x one candidate tile (tw x th candidates, tw*th <= 2048)
//   warp w                 = 16x16 luma block b = w of the CTU (bx = 16*(w&3), by = 16*(w>>2))
//   lane, per round        = one "unit": candidate column x, YB consecutive candidate rows
// Block phase (per round, per thread): YB candidates x 64 VABSDIFF4.U8.ACC build the 16 4x4 SADs of
// the block; every reference word fetched from shared memory serves up to YB candidates (vertical
// reuse in registers).  The 33 partitions that live inside a 16x16 block (8x4, 4x8, 8x8, 16x8, 8x16,
// 16x16 and the 8 AMP shapes) are built hierarchically, strip by strip as the SADs complete, with 37
// additions and one key formation per partition on the FMA pipe (IMAD):
//   key = sad * 2^11 + ((mvcost << 11) | candidateIndexInTile)
// and folded into 33 thread-private running keys on the ALU pipe -- one VIMNMX3 for two candidates at
// a time.  The low bits of the key reproduce the reference's first-in-scan-order tie-break.
// Upper phase (per round, lagging one round behind): the four 8x8 sums and the 16x16 sum of every
// (candidate, block) go through a shared-memory record ring guarded by mbarriers; 5 x (32*YB) threads
// (4 quadrant roles + one 64x64 role) build the 65 partitions of the 32x32 and 64x64 levels.
// Tile end: warp-wide CREDUX.MIN per key, conversion to the global 64-bit key (cost<<32 | y*(2R+1)+x)
// and one atomicMin per (warp, partition) into best[job][593]; a tiny finalize kernel decodes MVs.
//
// Reference window staging: rows arrive by TMA bulk copies (cp.async.bulk + mbarrier) and are expanded in shared memory
// into SLIDING ENTRIES -- entry x of a row holds bytes x..x+7 (64 bit; round 1: x..x+3) -- so lane l (candidate column x0+l) gets the 16
// reference bytes of a row with two LDS.64 (entries x and x+8) and lanes with consecutive candidate x never conflict, whatever the
// alignment of x.  The row pitch is a
// compile-time constant, which turns every address in the unrolled inner loop into an immediate offset.
// Addressing of the source plane is linear (row*pitch + col), which is exactly the reference's
// pelSearchArray[j + i*iRefStride] including its row-wrap quirk.
#pragma once
#include <cuda.h>   // CUtensorMap (type only: the encode function is fetched through cudaGetDriverEntryPoint, no link against libcuda)

#include "me_common.cuh"

namespace hmme {

constexpr int kFastThreads = 512;
constexpr int kIdxBits = 11;                  // candidates per tile <= 2048
constexpr int kMaxTileCands = 1 << kIdxBits;
#ifndef HMME_WIN32
#define HMME_WIN64 1                          // default since round 2 (measured: 1.270 -> 1.253 ms per 1080p +-64 frame); -DHMME_WIN32 builds the 32-bit layout
#endif
#ifdef HMME_WIN64
// Sliding 64-bit entries: entry x of a row = bytes x..x+7, so the 16 reference bytes a lane needs per row are TWO LDS.64 (entries x and
// x + 8) instead of four LDS.32 -- half the shared-memory load instructions for twice the window footprint (117 KB instead of 63 KB).
constexpr int kWinPitch64 = 187;              // entries per row: (tw-1) + 48 + 8 + 1 <= 185 for tw <= 129, padded so that YB * pitch == 1 (mod 16):
                                              // a half-warp straddling two row groups of a 129-wide tile still covers 16 distinct 8-byte bank pairs
constexpr int kWinPitch = 2 * kWinPitch64;    // words per window row
#else
constexpr int kWinPitch = 203;                // sliding-word entries per window row: (tw-1) + 4*15 + 1 <= 189 for tw <= 129, padded so that
                                              // YB * pitch == 129 (mod 32): a warp whose 32 units straddle two row groups of a 129-wide
                                              // tile (+-64) still reads 32 distinct banks
#endif
constexpr int kMaxTileW = 129;
constexpr int kKbPitch = kMaxTileW;           // words per row of the per-tile key-base table (constant: LDS offsets of a unit's YB rows are immediates)
constexpr int kDensePitch = 224;              // bytes per row of the TMA landing buffer: 15 (alignment) + 129 + 63 + 3, rounded up to 16
constexpr int kRecWords = 49;                 // upper-phase words per candidate slot: 32 (8x8 pairs, u16) + 16 (16x16<<11) + 1 (key base),
                                              // stored slot-minor ([word][slot]) so that every access is lane-contiguous
constexpr uint32_t kInvalidBlockKeyBase = 0xF0000000u;   // block sums << 11 stay below 2^27: no wrap, never wins
constexpr uint32_t kInvalidSlot = 0xFFFFFFFFu;
#ifndef HMME_SCHED_FENCE
#define HMME_SCHED_FENCE 0x3CF3C
#endif
constexpr unsigned kSchedFence = HMME_SCHED_FENCE;
constexpr int kLag = 1;                       // the upper phase of round k runs after the block phase of round k + kLag
constexpr int kRing = 2 * kLag + 2;           // record buffers: a slot may be rewritten only after every warp consumed it (ring >= 2*lag + 2)
static_assert((kRing & (kRing - 1)) == 0, "ring slots wrap with a mask");

struct alignas(64) FastParams {
    // Tensor maps (2-D tiled TMA): the whole reference window of a tile arrives with ONE cp.async.bulk.tensor instead of one bulk copy
    // per row (the copy engine takes about 20 cycles per request: 81 + 64 requests were 3000 cycles per tile), the 64x64 block with another.
    // Used per CTA when the window lies inside one plane row (no row wrap) and inside the rows the map covers; otherwise that CTA falls
    // back to the per-row copies, whose linear addressing keeps the reference's row-wrap behaviour.
    CUtensorMap refMap;        // u8 [refMapRows][refPitch] from refLo, box = kDensePitch x (tileRows + 63)
    CUtensorMap curMap;        // u8 rows of curPitch bytes from `cur`, box = 64 x 64
    int refMapOk, curMapOk;
    int refCol0, refRow0;      // (column, row) of picture sample (0,0) in refMap
    int refMapRows;
    const uint8_t* cur;        // picture sample (0,0) of the current plane
    const uint8_t* curBlocks;  // non-NULL: job j's 64x64 block is the dense 4 KiB record curBlocks + 4096*j instead (bi-prediction path)
    const uint8_t* ref;        // picture sample (0,0) of the reference plane
    const uint8_t* refLo;      // first addressable byte of the reference allocation (16B aligned)
    const uint8_t* refHi;      // one past the last addressable byte
    long long curPitch, refPitch;
    const int4* jobs;          // {ctuX, ctuY, ltx, lty}
    unsigned long long* best;  // [njobs][593] arg-min keys, all "no winner" between searches
    uint32_t lambda;
    int W;                     // 2R+1 candidates per axis
    int tw;                    // nominal width of a column stripe in candidates (<= kMaxTileW)
    int upt;                   // units per tile: a tile is `upt` consecutive units of its stripe's unit sequence (row group major)
    int tileRows;              // candidate rows (a multiple of YB) the row groups of any tile span at most: sizes the window
    int nTx, nTy;              // stripes per job, tiles per stripe
    uint32_t magicTiles, magicNTx, magicTw, magicTwLast;   // floor(2^32 / d) + 1 for d = nTx * nTy, nTx, tw, width of the last stripe: n / d = umulhi(n, magic)
    int stagger;               // SM cycles by which warps 8..15 start their first round late (0 = off), see the kernel
};

__host__ __device__ inline int fast_win_rows(int tileRows) { return tileRows + 63; }
// n / d with magic = floor(2^32 / d) + 1 (d >= 2; exact while n * d < 2^32) or 0 for d = 1
__device__ __forceinline__ int fast_div(uint32_t n, uint32_t magic) { return magic ? (int)__umulhi(n, magic) : (int)n; }

// A tile = `upt` consecutive units of one column stripe (unit = candidate column x one row group of YB candidate rows; units are
// numbered row group by row group).  With upt a multiple of 32 every round of every tile but a stripe's last has 32 busy lanes
// -- rectangular tiles of 129 columns waste most of one round per tile (129 = 4 * 32 + 1).  The first and the last row group of a tile
// are therefore partial in x: columns [xs, twA) of row group 0, [0, xe) of row group nRG - 1 ([xs, xe) if the tile has one row group).
struct TileGeo {
    int twA;                   // columns of the stripe
    int xs, xe, nRG;           // see above
    int w0;                    // columns of the first row group that belong to the tile
    int x0, y0;                // candidate (column, row) of the stripe's column 0 / the tile's row 0 in the job's window
};
// rank of candidate (row y of the tile, column x) in the tile's scan order (row major over the tile's candidates): the low key bits
__device__ __forceinline__ uint32_t tile_rank(const TileGeo& t, int y, int x, int yb) {
    const int rg = y / yb;
    if (rg == 0) return (uint32_t)(y * t.w0 + (x - t.xs));
    if (rg < t.nRG - 1) return (uint32_t)(yb * t.w0 + (y - yb) * t.twA + x);
    return (uint32_t)(yb * t.w0 + (t.nRG - 2) * yb * t.twA + (y - (t.nRG - 1) * yb) * t.xe + x);
}

// one VABSDIFF4.U8.ACC: acc += sum of |a.b[k] - b.b[k]| over the four bytes.  Written as PTX so that the accumulate operand
// stays chained (the compiler otherwise splits it into VABSDIFF4(...,RZ) + an extra add per packed SAD).
__device__ __forceinline__ uint32_t sad4_acc(uint32_t a, uint32_t b, uint32_t acc) {
    uint32_t d;
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(acc));
    return d;
}

// ---- mbarrier (shared::cta) helpers: the record ring between the block phase (producers: all 16 warps) and the upper phase
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
#ifdef HMME_POLL_SLEEP
    asm volatile(
        "{ .reg .pred p;\n"
        "W_%=: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra D_%=;\n"
        "nanosleep.u32 %2;\n"
        "bra W_%=;\n"
        "D_%=:\n}" ::"r"(a), "r"(parity), "n"(HMME_POLL_SLEEP) : "memory");
    return;
#endif
    asm volatile(
        "{ .reg .pred p;\n"
        "W_%=: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra W_%=;\n}" ::"r"(a), "r"(parity) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
// TMA bulk copy global -> shared (cp.async.bulk, SASS UBLKCP): 16-byte aligned source/destination, size a multiple of 16;
// completion is signalled as transaction bytes on the mbarrier.
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(dst)),
                 "l"(src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar))
                 : "memory");
}

// 2-D tiled TMA (cp.async.bulk.tensor, SASS UTMALDG): box of the tensor map at element coordinates (x, y); x * element size must be a
// multiple of 16 bytes, out-of-range elements arrive as zeros, the destination must be 128-byte aligned.
__device__ __forceinline__ void tma_tile_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(dst)),
                 "l"(map), "r"(x), "r"(y), "r"((uint32_t)__cvta_generic_to_shared(bar))
                 : "memory");
}

// partition index of block-level key k (0..32) for block b
__device__ __forceinline__ int block_part_index(int b, int k) {
    const int bxi = b & 3, byi = b >> 2, o = byi * 4 + bxi;
    if (k < 8)  return (4 * byi + (k >> 1)) * 8 + 2 * bxi + (k & 1);                   // 8x4
    if (k < 16) return 128 + (2 * byi + ((k - 8) >> 2)) * 16 + 4 * bxi + ((k - 8) & 3);  // 4x8
    if (k < 20) return 384 + (2 * byi + ((k - 16) >> 1)) * 8 + 2 * bxi + ((k - 16) & 1); // 8x8
    if (k < 28) return 256 + (k - 20) * 16 + o;                                        // 16x4 U/D, 16x12 U/D, 4x16 L/R, 12x16 L/R
    if (k < 30) return 448 + (2 * byi + (k - 28)) * 4 + bxi;                           // 16x8
    if (k < 32) return 480 + byi * 8 + 2 * bxi + (k - 30);                             // 8x16
    return 544 + o;                                                                    // 16x16
}
// partition index of quadrant-level key k (0..12) for quadrant q
__device__ __forceinline__ int quad_part_index(int q, int k) {
    const int qx = q & 1, qy = q >> 1;
    if (k < 8)  return 512 + k * 4 + q;                   // 32x8 U/D, 32x24 U/D, 8x32 L/R, 24x32 L/R
    if (k < 10) return 560 + (2 * qy + (k - 8)) * 2 + qx; // 32x16
    if (k < 12) return 568 + qy * 4 + 2 * qx + (k - 10);  // 16x32
    return 584 + q;                                       // 32x32
}
// partition index of CTU-level key k (0..12)
__device__ __forceinline__ int ctu_part_index(int k) {
    return k < 8 ? 576 + k : 588 + (k - 8);               // AMP 576..583, 64x32 588/589, 32x64 590/591, 64x64 592
}

// a + b on the FMA pipe (IMAD a*1+b).  The ALU pipe is the kernel's bottleneck (packed SADs + add-mins live there), so the
// hierarchy's plain additions are kept off it; written as PTX because the compiler otherwise folds them into 3-input ALU adds.
__device__ __forceinline__ uint32_t fadd(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, 1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
// (a << 11) + b, likewise one IMAD
__device__ __forceinline__ uint32_t fshladd(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, 2048, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
// Per-candidate state carried across the row loop.  The packed-SAD chains run over EIGHT rows (they restart at rows 0 and 8), so
// after strip 0 / 2 an accumulator holds a 4x4 sum and after strip 1 / 3 the 4x8 column sum -- the 4x8 level costs no addition.
// Sums are folded strip by strip as soon as they complete, so that the add/key work (FMA pipe) and the min updates interleave with
// the packed SADs (ALU pipe) of the candidates that are still being accumulated.  All sums are plain (unshifted) SADs; a key is
// formed by one IMAD: key = sum * 2^11 + ((mvcost << 11) | idxInTile), and a key of a difference or a sum of a keyed and an
// unkeyed part by one IMAD on the existing key (exact mod 2^32): 17 additions + 33 key IMADs per candidate for the 33 partitions.
// Eight-row SAD chains (the 4x8 column sums then cost no addition: 17 instead of 25 additions per candidate) are the default; -DHMME_CHAIN4
// builds the four-row form.  Without the scheduling fences of round_body the eight-row form was the slower one (1.22 vs 1.19 ms per 1080p +-64
// frame: ptxas sank all folding behind the SADs); with them it is the faster (1.140 vs 1.179 ms).
#ifndef HMME_CHAIN4
#define HMME_CHAIN8 1
#endif
struct BlockState {
#ifndef HMME_CHAIN8
    uint32_t sp[4];     // 4x4 sums of the previous even strip
#endif
    uint32_t ha[2];     // 8x4 sums of strip 0 (later: of strip 2)
    uint32_t q0, q2;    // 16x4 sums of strips 0 and 2
    uint32_t vl, vr;    // 4x8 sums of the top half, leftmost and rightmost column
    uint32_t e0[2];     // 8x8 sums of the top half
    uint32_t top, kTop; // 16x8 top and its key
};

// key of (whole - part) from the key of the whole: one IMAD, no subtraction of sums (keys are exact mod 2^32)
__device__ __forceinline__ uint32_t fsubkey(uint32_t part, uint32_t keyWhole) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, 0xFFFFF800, %2;" : "=r"(d) : "r"(part), "r"(keyWhole));
    return d;
}

// best = min(best, key[c] for the NC candidates handled together).  Two candidates cost ONE ALU instruction (VIMNMX3) instead of two
// ALU add-mins: the ALU pipe is the kernel's bottleneck, key formation lives on the FMA pipe.
template <int NC>
__device__ __forceinline__ void updk(uint32_t& best, const uint32_t (&key)[NC]) {
    if constexpr (NC == 1) best = min(best, key[0]);
    else best = min(min(best, key[0]), key[1]);
}
template <int NC>
__device__ __forceinline__ void upd(uint32_t& best, const uint32_t (&sum)[NC], const uint32_t (&kb)[NC]) {
    uint32_t k[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) k[c] = fshladd(sum[c], kb[c]);
    updk<NC>(best, k);
}

// Strip T (rows 4T..4T+3 of the 16x16 block) of NC candidates is complete: a[c][i] = the four accumulators of candidate c -- the 4x4
// sums of the strip for T = 0, 2, the 4x8 column sums of the half for T = 1, 3.
// rec[c] points at candidate c's slot of the upper-level record (word w lives at rec[c][w * slots]).
template <int T, int NC>
__device__ __forceinline__ void emit_strip(const uint32_t (&ain)[NC][4], BlockState (&st)[NC], const uint32_t (&kb)[NC], uint32_t (&best)[33],
                                           uint32_t* const (&rec)[NC], int b, bool writeBase, const uint32_t (&recBase)[NC], int slots) {
    uint32_t k[NC];
#ifdef HMME_CHAIN8
    const uint32_t (&a)[NC][4] = ain;
#else
    uint32_t a[NC][4];                                             // four-row chains: the 4x8 column sums are formed here
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (T == 0 || T == 2) { a[c][i] = ain[c][i]; st[c].sp[i] = ain[c][i]; }
            else a[c][i] = fadd(st[c].sp[i], ain[c][i]);
        }
#endif
    if constexpr (T == 0 || T == 2) {
        uint32_t h0[NC], h1[NC], q[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            h0[c] = st[c].ha[0] = fadd(a[c][0], a[c][1]);
            h1[c] = st[c].ha[1] = fadd(a[c][2], a[c][3]);
            q[c] = fadd(h0[c], h1[c]);
            if (T == 0) st[c].q0 = q[c]; else st[c].q2 = q[c];
        }
        upd<NC>(best[2 * T], h0, kb);                              // 8x4
        upd<NC>(best[2 * T + 1], h1, kb);
        if constexpr (T == 0) upd<NC>(best[20], q, kb);            // 16x4  (2NxnU part 0)
    } else {
        uint32_t e0[NC], e1[NC], hf[NC], kE0[NC], kE1[NC], kH[NC];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint32_t v[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) v[c] = a[c][i];
            upd<NC>(best[(T == 1 ? 8 : 12) + i], v, kb);           // 4x8: the chain itself
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            e0[c] = fadd(a[c][0], a[c][1]); e1[c] = fadd(a[c][2], a[c][3]);   // 8x8
            hf[c] = fadd(e0[c], e1[c]);                                        // 16x8
            kE0[c] = fshladd(e0[c], kb[c]); kE1[c] = fshladd(e1[c], kb[c]); kH[c] = fshladd(hf[c], kb[c]);
        }
        updk<NC>(best[T == 1 ? 16 : 18], kE0);                     // 8x8
        updk<NC>(best[T == 1 ? 17 : 19], kE1);
        updk<NC>(best[T == 1 ? 28 : 29], kH);                      // 16x8
#pragma unroll
        for (int c = 0; c < NC; ++c) k[c] = fsubkey(st[c].ha[0], kE0[c]);
        updk<NC>(best[2 * T], k);                                  // 8x4 of this strip = 8x8 - 8x4 of the strip above
#pragma unroll
        for (int c = 0; c < NC; ++c) k[c] = fsubkey(st[c].ha[1], kE1[c]);
        updk<NC>(best[2 * T + 1], k);
        if constexpr (T == 1) {
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                st[c].vl = a[c][0]; st[c].vr = a[c][3];
                st[c].e0[0] = e0[c]; st[c].e0[1] = e1[c];
                st[c].top = hf[c]; st[c].kTop = kH[c];
            }
        } else {
            uint32_t cl[NC], cr[NC], left[NC], right[NC], all[NC], kAll[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                cl[c] = fadd(st[c].vl, a[c][0]); cr[c] = fadd(st[c].vr, a[c][3]);
                left[c] = fadd(st[c].e0[0], e0[c]); right[c] = fadd(st[c].e0[1], e1[c]);
                all[c] = fadd(st[c].top, hf[c]);
                kAll[c] = fshladd(all[c], kb[c]);
            }
#pragma unroll
            for (int c = 0; c < NC; ++c) k[c] = fsubkey(st[c].q2, kH[c]);
            updk<NC>(best[21], k);                                 // 16x4  (2NxnD part 1) = bottom half - strip 2
            upd<NC>(best[24], cl, kb);                             // 4x16  (nLx2N part 0)
            upd<NC>(best[25], cr, kb);                             // 4x16  (nRx2N part 1)
            upd<NC>(best[30], left, kb);                           // 8x16
            upd<NC>(best[31], right, kb);
            updk<NC>(best[32], kAll);                              // 16x16
#pragma unroll
            for (int c = 0; c < NC; ++c) k[c] = fshladd(st[c].q2, st[c].kTop);
            updk<NC>(best[22], k);                                 // 16x12 rows 0..11 = top half + strip 2
#pragma unroll
            for (int c = 0; c < NC; ++c) k[c] = fsubkey(st[c].q0, kAll[c]);
            updk<NC>(best[23], k);                                 // 16x12 rows 4..15 = all - strip 0
#pragma unroll
            for (int c = 0; c < NC; ++c) k[c] = fsubkey(cr[c], kAll[c]);
            updk<NC>(best[26], k);                                 // 12x16 cols 0..11 = all - right column
#pragma unroll
            for (int c = 0; c < NC; ++c) k[c] = fsubkey(cl[c], kAll[c]);
            updk<NC>(best[27], k);                                 // 12x16 cols 4..15 = all - left column
            // upper-level hand-over: 8x8 sums as u16 pairs (<= 16320 each), 16x16 sum (<= 65280)
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                uint32_t pk0, pk1;
                asm("mad.lo.u32 %0, %1, 65536, %2;" : "=r"(pk0) : "r"(st[c].e0[1]), "r"(st[c].e0[0]));
                asm("mad.lo.u32 %0, %1, 65536, %2;" : "=r"(pk1) : "r"(e1[c]), "r"(e0[c]));
                rec[c][(2 * b) * slots] = pk0;
                rec[c][(2 * b + 1) * slots] = pk1;
                rec[c][(32 + b) * slots] = all[c];
                if (writeBase) rec[c][48 * slots] = recBase[c];
            }
        }
    }
}

// low + high half of a packed u16 pair (each sum <= 65535): (r * 0x10001) >> 16, one IMAD + one shift
__device__ __forceinline__ uint32_t lohi(uint32_t r) {
    uint32_t t;
    asm("mad.lo.u32 %0, %1, 65537, 0;" : "=r"(t) : "r"(r));
    return t >> 16;
}
__device__ __forceinline__ void upd1(uint32_t& best, uint32_t sum, uint32_t kb) { best = min(best, fshladd(sum, kb)); }

__device__ __forceinline__ void emit_quadrant(const uint32_t* rec, int q, uint32_t (&ub)[13], int slots) {
    const uint32_t kb = rec[48 * slots];
    if (kb == kInvalidSlot) return;
    const int b0 = 8 * (q >> 1) + 2 * (q & 1);
    uint4 T, B;                                                             // blocks b0, b0+1 / b0+4, b0+5 : {e00|e01, e10|e11} each
    T.x = rec[(2 * b0) * slots]; T.y = rec[(2 * b0 + 1) * slots]; T.z = rec[(2 * b0 + 2) * slots]; T.w = rec[(2 * b0 + 3) * slots];
    B.x = rec[(2 * b0 + 8) * slots]; B.y = rec[(2 * b0 + 9) * slots]; B.z = rec[(2 * b0 + 10) * slots]; B.w = rec[(2 * b0 + 11) * slots];
    // rows of 8x8 sums across the 32-wide quadrant, still packed (each half <= 32640), then low + high
    const uint32_t R0 = lohi(fadd(T.x, T.z)), R1 = lohi(fadd(T.y, T.w)), R2 = lohi(fadd(B.x, B.z)), R3 = lohi(fadd(B.y, B.w));
    // columns: packed sums over the four 8-row strips (each half <= 65280)
    const uint32_t cl = fadd(fadd(T.x, T.y), fadd(B.x, B.y)), cr = fadd(fadd(T.z, T.w), fadd(B.z, B.w));
    const uint32_t C0 = cl & 0xFFFFu, C1 = cl >> 16, C2 = cr & 0xFFFFu, C3 = cr >> 16;
    const uint32_t top = fadd(R0, R1), bot = fadd(R2, R3), left = fadd(C0, C1), right = fadd(C2, C3);
    const uint32_t kAll = fshladd(fadd(top, bot), kb);
    upd1(ub[0], R0, kb);                           // 32x8  (2NxnU part 0)
    upd1(ub[1], R3, kb);                           // 32x8  (2NxnD part 1)
    ub[2] = min(ub[2], fsubkey(R3, kAll));         // 32x24 rows 0..23 = all - last strip
    ub[3] = min(ub[3], fsubkey(R0, kAll));         // 32x24 rows 8..31 = all - first strip
    upd1(ub[4], C0, kb);                           // 8x32
    upd1(ub[5], C3, kb);
    ub[6] = min(ub[6], fsubkey(C3, kAll));         // 24x32 cols 0..23
    ub[7] = min(ub[7], fsubkey(C0, kAll));         // 24x32 cols 8..31
    upd1(ub[8], top, kb);                          // 32x16
    upd1(ub[9], bot, kb);
    upd1(ub[10], left, kb);                        // 16x32
    upd1(ub[11], right, kb);
    ub[12] = min(ub[12], kAll);                    // 32x32
}

__device__ __forceinline__ void emit_ctu(const uint32_t* rec, uint32_t (&ub)[13], int slots) {
    const uint32_t kb = rec[48 * slots];
    if (kb == kInvalidSlot) return;
    uint4 m[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {                                           // 16x16 sums, block row r
        m[r].x = rec[(32 + 4 * r) * slots]; m[r].y = rec[(33 + 4 * r) * slots];
        m[r].z = rec[(34 + 4 * r) * slots]; m[r].w = rec[(35 + 4 * r) * slots];
    }
    uint32_t R[4], C[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) R[r] = fadd(fadd(m[r].x, m[r].y), fadd(m[r].z, m[r].w));
    C[0] = fadd(fadd(m[0].x, m[1].x), fadd(m[2].x, m[3].x));
    C[1] = fadd(fadd(m[0].y, m[1].y), fadd(m[2].y, m[3].y));
    C[2] = fadd(fadd(m[0].z, m[1].z), fadd(m[2].z, m[3].z));
    C[3] = fadd(fadd(m[0].w, m[1].w), fadd(m[2].w, m[3].w));
    const uint32_t top = fadd(R[0], R[1]), bot = fadd(R[2], R[3]), left = fadd(C[0], C[1]), right = fadd(C[2], C[3]);
    const uint32_t kAll = fshladd(fadd(top, bot), kb);
    upd1(ub[0], R[0], kb);                         // 64x16 (2NxnU part 0)
    upd1(ub[1], R[3], kb);                         // 64x16 (2NxnD part 1)
    ub[2] = min(ub[2], fsubkey(R[3], kAll));       // 64x48 rows 0..47 = all - last strip
    ub[3] = min(ub[3], fsubkey(R[0], kAll));       // 64x48 rows 16..63
    upd1(ub[4], C[0], kb);                         // 16x64
    upd1(ub[5], C[3], kb);
    ub[6] = min(ub[6], fsubkey(C[3], kAll));       // 48x64 cols 0..47
    ub[7] = min(ub[7], fsubkey(C[0], kAll));       // 48x64 cols 16..63
    upd1(ub[8], top, kb);                          // 64x32
    upd1(ub[9], bot, kb);
    upd1(ub[10], left, kb);                        // 32x64
    upd1(ub[11], right, kb);
    ub[12] = min(ub[12], kAll);                    // 64x64
}

// tile key -> global key, one atomicMin into best[job][part]
template <int YB>
__device__ __forceinline__ void publish(unsigned long long* bestJob, int part, uint32_t key, const TileGeo& t, int W) {
    if (key == 0xFFFFFFFFu) return;
    const uint32_t cost = key >> kIdxBits;
    int r = (int)(key & (kMaxTileCands - 1)), ty, tx;
    const int nFirst = YB * t.w0, nMid = (t.nRG - 2) * YB * t.twA;
    if (r < nFirst) { ty = r / t.w0; tx = t.xs + r - ty * t.w0; }
    else if (r - nFirst < nMid) { r -= nFirst; ty = r / t.twA; tx = r - ty * t.twA; ty += YB; }
    else { r -= nFirst + nMid; ty = r / t.xe; tx = r - ty * t.xe; ty += (t.nRG - 1) * YB; }
    const uint32_t gidx = (uint32_t)(t.y0 + ty) * (uint32_t)W + (uint32_t)(t.x0 + tx);
    atomicMin(bestJob + part, ((unsigned long long)cost << 32) | gidx);
}

// One round of one warp: lane = unit (candidate column ux, rows rg*YB .. rg*YB+YB-1), block (bx, by).
// CHECKED = false is the steady state (all lanes and rows valid); CHECKED = true handles the ragged last rounds.
template <int YB, bool CHECKED>
__device__ __forceinline__ void round_body(const uint32_t* sWin, const uint32_t* cp, const uint32_t* sKb,
                                           uint32_t* recBuf, uint32_t (&best)[33], int rg, int ux, int by, int bx, int b, int lane,
                                           bool unitValid, bool never) {
    constexpr int SLOTS = 32 * YB;
    const bool uvalid = !CHECKED || unitValid;
    const int rgc = (CHECKED && !uvalid) ? 0 : rg, uxc = (CHECKED && !uvalid) ? 0 : ux;
#ifdef HMME_WIN64
    const uint2* wp = reinterpret_cast<const uint2*>(sWin) + (rgc * YB + by) * kWinPitch64 + uxc + bx;
#else
    const uint32_t* wp = sWin + (rgc * YB + by) * kWinPitch + uxc + bx;
#endif
    uint32_t kb[YB], kbRec[YB];
    const uint32_t* kp = sKb + rgc * (YB * kKbPitch) + uxc;
#pragma unroll
    for (int j = 0; j < YB; ++j) {
        const uint32_t k = kp[j * kKbPitch];             // (mvcost << 11) | rank in tile, tabulated while the window was in flight
        const bool valid = !CHECKED || (uvalid && k != kInvalidBlockKeyBase);   // rows past the window's last one are marked in the table
        kb[j] = valid ? k : kInvalidBlockKeyBase;
        kbRec[j] = valid ? k : kInvalidSlot;
    }
    // Candidates 0 and 1 are folded together (their strips complete one reference row apart), the odd one out alone.
    uint32_t acc[YB][4], held[4];
    BlockState stP[2], stS[1];
    const uint32_t kbP[2] = {kb[0], kb[1]}, kbRecP[2] = {kbRec[0], kbRec[1]};
    uint32_t* const recP[2] = {recBuf + lane, recBuf + 32 + lane};
    const uint32_t kbS[1] = {kb[YB - 1]}, kbRecS[1] = {kbRec[YB - 1]};
    uint32_t* const recS[1] = {recBuf + (YB - 1) * 32 + lane};
#pragma unroll
    for (int j = 0; j < YB; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[j][i] = 0;
    uint4 cw[YB];
#pragma unroll
    for (int rho = 0; rho < 16 + YB - 1; ++rho) {
#ifdef HMME_WIN64
        const uint2 ra = wp[rho * kWinPitch64], rb = wp[rho * kWinPitch64 + 8];
        const uint32_t r0 = ra.x, r1 = ra.y, r2 = rb.x, r3 = rb.y;
#else
        const uint32_t r0 = wp[rho * kWinPitch + 0], r1 = wp[rho * kWinPitch + 4], r2 = wp[rho * kWinPitch + 8], r3 = wp[rho * kWinPitch + 12];
#endif
        if (rho < 16) cw[rho % YB] = *reinterpret_cast<const uint4*>(cp + rho * 16);
#pragma unroll
        for (int j = 0; j < YB; ++j) {
            const int r = rho - j;                           // row of the block this reference row meets for candidate j
            if (r >= 0 && r < 16) {
                const uint4 c = cw[r % YB];
                acc[j][0] = sad4_acc(c.x, r0, acc[j][0]);
                acc[j][1] = sad4_acc(c.y, r1, acc[j][1]);
                acc[j][2] = sad4_acc(c.z, r2, acc[j][2]);
                acc[j][3] = sad4_acc(c.w, r3, acc[j][3]);
                if ((r & 3) == 3) {
                    if (j == 0) {                            // hold candidate 0's strip until candidate 1's arrives with the next row
#pragma unroll
                        for (int i = 0; i < 4; ++i) held[i] = acc[0][i];
                    } else if (j == 1) {
                        const uint32_t a2[2][4] = {{held[0], held[1], held[2], held[3]}, {acc[1][0], acc[1][1], acc[1][2], acc[1][3]}};
                        if (r == 3) emit_strip<0, 2>(a2, stP, kbP, best, recP, b, b == 0, kbRecP, SLOTS);
                        if (r == 7) emit_strip<1, 2>(a2, stP, kbP, best, recP, b, b == 0, kbRecP, SLOTS);
                        if (r == 11) emit_strip<2, 2>(a2, stP, kbP, best, recP, b, b == 0, kbRecP, SLOTS);
                        if (r == 15) emit_strip<3, 2>(a2, stP, kbP, best, recP, b, b == 0, kbRecP, SLOTS);
                    } else {
                        const uint32_t a1[1][4] = {{acc[j][0], acc[j][1], acc[j][2], acc[j][3]}};
                        if (r == 3) emit_strip<0, 1>(a1, stS, kbS, best, recS, b, b == 0, kbRecS, SLOTS);
                        if (r == 7) emit_strip<1, 1>(a1, stS, kbS, best, recS, b, b == 0, kbRecS, SLOTS);
                        if (r == 11) emit_strip<2, 1>(a1, stS, kbS, best, recS, b, b == 0, kbRecS, SLOTS);
                        if (r == 15) emit_strip<3, 1>(a1, stS, kbS, best, recS, b, b == 0, kbRecS, SLOTS);
                    }
#ifdef HMME_CHAIN8
                    if ((r & 7) == 7)                            // the chains restart at rows 0 and 8
#endif
                    {
#pragma unroll
                        for (int i = 0; i < 4; ++i) acc[j][i] = 0;
                    }
                }
            }
        }
        // A never-executed, predicated `trap` that ptxas does not move code across: without such fences it hoists all packed SADs of the
        // round to the front and sinks the folding work (FMA pipe) behind them, so a warp alternates between long single-pipe stretches.
        // kSchedFence = bit mask of the rows fenced (default 0x3CF3C); the masks tried span 1.13 .. 1.16 ms, no fences 1.159, every row 1.151.
        if ((kSchedFence >> rho) & 1)
            if (never) asm volatile("trap;");
    }
}

template <int YB>
__global__ void __launch_bounds__(kFastThreads, 1) me_u8_tile_kernel(const __grid_constant__ FastParams p) {
    extern __shared__ __align__(128) uint32_t smem[];
    constexpr int SLOTS = 32 * YB;
    const int winRows = fast_win_rows(p.tileRows);
    uint32_t* sWin = smem;                                  // winRows x kWinPitch sliding words
    uint32_t* sCur = sWin + ((winRows * kWinPitch + 31) & ~31); // 64 rows x 16 words, 128-byte aligned (LDS.128, TMA destination)
    uint32_t* sUp = sCur + 1024;                            // kRing x kRecWords x SLOTS
    uint32_t* sKb = sUp + kRing * SLOTS * kRecWords;        // tileRows x kKbPitch key bases
    uint32_t* sBitsX = sKb + p.tileRows * kKbPitch;         // kMaxTileW MV-bit counts of the tile's columns, then tileRows of its rows
    uint32_t* sBitsY = sBitsX + kMaxTileW;
    uint32_t* sRowBase = sBitsY + p.tileRows;               // per row of the tile: rank of its column 0 in the tile's scan order (tile_rank(y, 0))

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#ifndef HMME_NO_MAP_PREFETCH
    // the copy engine reads the 128-byte tensor maps from the parameter space: fetch them while the job record is on its way
    if (tid == 0 && p.refMapOk) asm volatile("prefetch.tensormap [%0];" ::"l"(&p.refMap) : "memory");
    if (tid == 32 && p.curMapOk) asm volatile("prefetch.tensormap [%0];" ::"l"(&p.curMap) : "memory");
#endif
#ifdef HMME_DIAG_TIMES
    long long dt[6];
    dt[0] = clock64();
#define HMME_DIAG_T(i) dt[i] = clock64()
#else
#define HMME_DIAG_T(i)
#endif
    // divisions by launch constants as multiplications (exact for these magnitudes: n * d < 2^32): the tile geometry is on the critical path
    // to the job record and the copies
    const int tilesPerJob = p.nTx * p.nTy;
    const int job = fast_div(blockIdx.x, p.magicTiles), tile = blockIdx.x - job * tilesPerJob;
    const int tiy = fast_div((uint32_t)tile, p.magicNTx), tix = tile - tiy * p.nTx;
    const int x0 = tix * p.tw;
    const int twA = min(p.tw, p.W - x0);
    const int u0 = tiy * p.upt;                              // first unit of the tile in its stripe's unit sequence
    const int nUnits = min(p.upt, ((p.W + YB - 1) / YB) * twA - u0);
    if (nUnits <= 0) return;                                 // a narrower last stripe has fewer units than the nominal one
    const uint32_t magicW = twA == p.tw ? p.magicTw : p.magicTwLast;
    const int rg0 = fast_div((uint32_t)u0, magicW), rgLast = fast_div((uint32_t)(u0 + nUnits - 1), magicW);
    TileGeo tg;
    tg.twA = twA; tg.nRG = rgLast - rg0 + 1;
    tg.xs = u0 - rg0 * twA; tg.xe = u0 + nUnits - rgLast * twA;
    tg.w0 = (tg.nRG == 1 ? tg.xe : twA) - tg.xs;
    tg.x0 = x0; tg.y0 = rg0 * YB;
    const int nRG = tg.nRG, y0 = tg.y0;
    const int thA = min(nRG * YB, p.W - y0);                 // candidate rows of the tile's row groups that exist in the window
    const int4 jb = p.jobs[job];
    __shared__ uint64_t fullBar[kRing];
    __shared__ uint64_t winBar;
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < kRing; ++k) mbar_init(&fullBar[k], kFastThreads / 32);
    }

    // ---- stage the reference window, the CTU and the MV-bit tables.
    // TMA bulk copies (one per window row, 16-byte aligned superset of the row's bytes; linear source addressing, so the
    // reference's row-wrap behaviour is kept) land in a dense buffer that aliases the not-yet-used record ring; the CTU rows
    // go the same way when they are 16-byte aligned.  All threads then expand the dense rows into sliding words.
    {
        // rows the unrolled loop addresses (nRG * YB + 63) vs rows that exist in the job's window: when thA is not a multiple of YB
        // the last one or two belong to masked candidates only and may lie past the plane's last row, so they are not fetched
        const int rows = nRG * YB + 63, rowsReal = thA + 63, nPos = twA + 60;      // entries 0 .. (twA-1) + 4*15
        const uint8_t* wbase = p.ref + (long long)(jb.y + jb.w + y0) * p.refPitch + (jb.x + jb.z + x0);
        const long long curPitch = p.curBlocks ? 64 : p.curPitch;
        const uint8_t* cbase = p.curBlocks ? p.curBlocks + (size_t)job * 4096 : p.cur + (long long)jb.y * p.curPitch + jb.x;
        const bool curTma = (((uintptr_t)cbase | (uintptr_t)curPitch) & 15) == 0;
        // how the 64x64 block arrives: 0 = thread loads, 1 = 64 row copies, 2 = one 4 KiB copy (dense record), 3 = one 2-D tile
        const int curMode = p.curBlocks ? 2 : (p.curMapOk ? 3 : (curTma ? 1 : 0));
        // the window as one 2-D tile: its columns must lie inside one plane row (the per-row path keeps the reference's row wrap) and its rows
        // inside the map; rows past the map arrive as zeros and belong to masked candidates only
        const int wc = p.refCol0 + jb.x + jb.z + x0, wr = p.refRow0 + jb.y + jb.w + y0;
        const bool ref2d = p.refMapOk && wc >= 0 && (long long)wc + nPos + 3 <= p.refPitch && wr >= 0 && wr + rowsReal <= p.refMapRows;
        uint8_t* dense = reinterpret_cast<uint8_t*>(sUp);
        if (tid == 0) {
            mbar_init(&winBar, (uint32_t)((ref2d ? 1 : rowsReal) + (curMode == 1 ? 64 : (curMode ? 1 : 0))));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            // the one-copy forms leave from here, before the block-wide barrier: issuing a tile copy holds the thread for ~900 cycles (the copy
            // engine fetches the tensor map), which now overlaps the MV-bit tables of the other warps and the barrier itself
            if (ref2d) {
                mbar_arrive_expect_tx(&winBar, (uint32_t)(winRows * kDensePitch));
                tma_tile_2d(dense, &p.refMap, wc & ~15, wr, &winBar);
            }
            if (curMode >= 2) {
                mbar_arrive_expect_tx(&winBar, 4096);
                if (curMode == 2) tma_bulk_g2s(sCur, cbase, 4096, &winBar);
                else tma_tile_2d(sCur, &p.curMap, jb.x, jb.y, &winBar);
            }
        }
        if (tid >= 256) {                                     // MV-bit counts per column and per row of the tile (the warps that issue no copies)
            for (int x = tid - 256; x < twA; x += 256) sBitsX[x] = mv_bits(4 * (jb.z + x0 + x));
            for (int y = tid - 256; y < nRG * YB; y += 256) { sBitsY[y] = mv_bits(4 * (jb.w + y0 + y)); sRowBase[y] = tile_rank(tg, y, 0, YB); }
        }
        __syncthreads();
        HMME_DIAG_T(1);
        if (!ref2d && tid < rowsReal) {
            const uintptr_t g = (uintptr_t)(wbase + (long long)tid * p.refPitch);
            const uintptr_t g0 = g & ~(uintptr_t)15;
            uint32_t bytes = (uint32_t)(((g - g0) + (uintptr_t)(nPos + 3) + 15) & ~(uintptr_t)15);
            const uintptr_t room = g0 < (uintptr_t)p.refHi ? ((uintptr_t)p.refHi - g0) & ~(uintptr_t)15 : 0;   // never read past the allocation (+slack)
            bytes = (uint32_t)min((uintptr_t)bytes, room);
            if (bytes) {
                mbar_arrive_expect_tx(&winBar, bytes);
                tma_bulk_g2s(dense + tid * kDensePitch, reinterpret_cast<const void*>(g0), bytes, &winBar);
            } else
                mbar_arrive(&winBar);
        }
        if (curMode == 1 && tid >= kFastThreads - 64) {
            const int r = tid - (kFastThreads - 64);
            mbar_arrive_expect_tx(&winBar, 64);
            tma_bulk_g2s(sCur + r * 16, cbase + (long long)r * curPitch, 64, &winBar);
        }
        if (curMode == 0) {
            for (int idx = tid; idx < 1024; idx += kFastThreads) {
                const uint8_t* c = cbase + (long long)(idx >> 4) * curPitch + 4 * (idx & 15);
                sCur[idx] = (uint32_t)c[0] | ((uint32_t)c[1] << 8) | ((uint32_t)c[2] << 16) | ((uint32_t)c[3] << 24);
            }
        }
        // key base of every candidate of the tile: (lambda * (bits(mvx) + bits(mvy)) >> 16) << 11 | scan-order index in the tile
        // (warp 0 issues the window copy, which holds it up for about as long as the copy takes: the other fifteen fill the table)
        for (int y = warp - 1; y >= 0 && y < nRG * YB; y += kFastThreads / 32 - 1) {
            const uint32_t bitsY = sBitsY[y], rowBase = sRowBase[y];    // the rank is linear in x inside a row
            for (int x = lane; x < twA; x += 32)
                sKb[y * kKbPitch + x] = y < thA ? ((((uint32_t)(p.lambda * (sBitsX[x] + bitsY)) >> 16) << kIdxBits) | ((rowBase + (uint32_t)x) & (kMaxTileCands - 1)))
                                                : kInvalidBlockKeyBase;
        }
        HMME_DIAG_T(2);
        mbar_wait(&winBar, 0);
        HMME_DIAG_T(3);
        for (int row = kFastThreads / 32 - 1 - warp; row < rows; row += kFastThreads / 32) {   // warp 0 issued the copies and starts last: it gets the fewest rows
            const uint32_t off = (uint32_t)((uintptr_t)(wbase + (long long)row * p.refPitch) & 15);
            const uint32_t* d = reinterpret_cast<const uint32_t*>(dense + row * kDensePitch);
#ifdef HMME_WIN64
            // entries 0 .. (twA-1) + 48 + 8, two per lane and step: the pair shares its three source words and leaves with one 128-bit store
            // when the row starts on a 16-byte boundary of the window (even rows: the pitch is an odd number of entries)
            uint2* dst = reinterpret_cast<uint2*>(sWin) + row * kWinPitch64;
            const int odd = row & 1;                          // entry index of the first 16-byte aligned entry of this row
            if (odd && lane == 0) {
                const uint32_t sh = 8 * (off & 3);
                const uint32_t w0 = d[off >> 2], w1 = d[(off >> 2) + 1], w2 = d[(off >> 2) + 2];
                dst[0] = make_uint2(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh));
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {                       // x < nPos - 4 <= 185: three steps of 64 entries, unrolled so that their loads overlap
                const int x = odd + 2 * lane + 64 * k;
                if (x >= nPos - 4) break;
                const uint32_t q = off + (uint32_t)x, sh = 8 * (q & 3);
                const uint32_t w0 = d[q >> 2], w1 = d[(q >> 2) + 1], w2 = d[(q >> 2) + 2];
                const bool cross = (q & 3) == 3;               // the second entry starts on the next word
                const uint4 e = make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), cross ? w1 : __funnelshift_r(w0, w1, sh + 8),
                                           cross ? w2 : __funnelshift_r(w1, w2, sh + 8));
                *reinterpret_cast<uint4*>(dst + x) = e;         // x + 1 <= twA + 56 < kWinPitch64: the spare entry stays inside the row
            }
#else
            uint32_t* dst = sWin + row * kWinPitch;
            for (int x = lane; x < nPos; x += 32) {           // lane-contiguous stores, 4-lane broadcast loads: conflict-free
                const uint32_t q = off + (uint32_t)x;
                dst[x] = __funnelshift_r(d[q >> 2], d[(q >> 2) + 1], 8 * (q & 3));
            }
#endif
        }
    }
    __syncthreads();

    HMME_DIAG_T(4);
    const int b = warp, bx = (b & 3) * 16, by = (b >> 2) * 16;
    uint32_t best[33], ub[13];
#pragma unroll
    for (int k = 0; k < 33; ++k) best[k] = 0xFFFFFFFFu;
#pragma unroll
    for (int k = 0; k < 13; ++k) ub[k] = 0xFFFFFFFFu;
    const int role = tid / SLOTS, slot = tid - role * SLOTS;   // upper phase: role 0..3 quadrant, 4 CTU level, >=5 idle
    const uint32_t* cp = sCur + by * 16 + (bx >> 2);

#ifdef HMME_DIAG_NOROUNDS
    const int nRounds = 0;                                  // diagnostic build: staging and tile end only
#else
    const int nRounds = (nUnits + 31) >> 5;
#endif
    // rounds in which every lane has a unit and every candidate row exists (rows are missing only in a window's last, partial row group)
    const int nFull = min(nUnits, thA < nRG * YB ? max(0, (nRG - 1) * twA - tg.xs) : nUnits) >> 5;
    int rg = fast_div((uint32_t)(tg.xs + lane), magicW), ux = tg.xs + lane - rg * twA;   // this lane's unit of round 0 (row group within the tile, column), advanced incrementally
    int unit = lane;
    const bool never = p.W < 0;                             // false, but not to the compiler (see HMME_SCHED_FENCE)
    // Warps are NOT barrier-locked per round: records travel through a ring of kRing buffers guarded by mbarriers, and the
    // upper phase of round k runs after the block phase of round k+kLag, by which time every producer has long arrived.  That
    // lets the two warps of each scheduler that start `stagger` cycles late stay half a round out of phase, so the
    // SAD-heavy (ALU pipe) and sum-heavy (FMA pipe) stretches of different warps overlap.
    if (p.stagger > 0 && (warp & 8)) {
        const long long t0 = clock64();
        while (clock64() - t0 < p.stagger) {}
    }
#ifndef HMME_LOOP_PLAIN
    // loop control in down-counters (loop-carried, so the compiler cannot re-derive the round counts from the tile geometry in every
    // iteration); a round advances every lane by 32 units = stepRg row groups + stepUx columns (one predicated wrap, no loop)
    const int stepRg = fast_div(32u, magicW), stepUx = 32 - stepRg * twA;
    int blockLeft = nRounds, fullLeft = nFull, upperLeft = nRounds, lagLeft = kLag;
    int wSlot = 0, rSlot = 0;
    uint32_t rPhase = 0;
    while (upperLeft > 0) {
        if (blockLeft > 0) {
            uint32_t* recBuf = sUp + wSlot * (SLOTS * kRecWords);
            if (fullLeft > 0)
                round_body<YB, false>(sWin, cp, sKb, recBuf, best, rg, ux, by, bx, b, lane, true, never);
            else
                round_body<YB, true>(sWin, cp, sKb, recBuf, best, rg, ux, by, bx, b, lane, unit < nUnits, never);
            --blockLeft; --fullLeft;
            unit += 32;
            ux += stepUx; rg += stepRg;
            if (ux >= twA) { ux -= twA; ++rg; }
            __syncwarp();
            if (lane == 0) mbar_arrive(&fullBar[wSlot]);
            wSlot = (wSlot + 1) & (kRing - 1);
        }
        if (lagLeft > 0) --lagLeft;
        else {
            mbar_wait(&fullBar[rSlot], rPhase);
            const uint32_t* recBuf = sUp + rSlot * (SLOTS * kRecWords);
            if (role < 4) emit_quadrant(recBuf + slot, role, ub, SLOTS);
            else if (role == 4) emit_ctu(recBuf + slot, ub, SLOTS);
            rSlot = (rSlot + 1) & (kRing - 1);
            if (rSlot == 0) rPhase ^= 1u;
            --upperLeft;
        }
    }
#else
    int wSlot = 0, rSlot = 0;
    uint32_t rPhase = 0;
    for (int round = 0; round < nRounds + kLag; ++round) {
        if (round < nRounds) {
            uint32_t* recBuf = sUp + wSlot * (SLOTS * kRecWords);
            if (round < nFull)
                round_body<YB, false>(sWin, cp, sKb, recBuf, best, rg, ux, by, bx, b, lane, true, never);
            else
                round_body<YB, true>(sWin, cp, sKb, recBuf, best, rg, ux, by, bx, b, lane, unit < nUnits, never);
            unit += 32;
            ux += 32;
            while (ux >= twA) { ux -= twA; ++rg; }
            __syncwarp();
            if (lane == 0) mbar_arrive(&fullBar[wSlot]);
            wSlot = (wSlot + 1 == kRing) ? 0 : wSlot + 1;
        }
        if (round >= kLag) {
            mbar_wait(&fullBar[rSlot], rPhase);
            const uint32_t* recBuf = sUp + rSlot * (SLOTS * kRecWords);
            if (role < 4) emit_quadrant(recBuf + slot, role, ub, SLOTS);
            else if (role == 4) emit_ctu(recBuf + slot, ub, SLOTS);
            if (++rSlot == kRing) { rSlot = 0; rPhase ^= 1u; }
        }
    }

#endif
    HMME_DIAG_T(5);
#ifdef HMME_DIAG_TIMES
    if (tid == 0 && (blockIdx.x % 997) == 5)
        printf("cta %d: init %lld, issue+tables %lld, tma wait %lld, expand %lld, rounds %lld cycles\n", (int)blockIdx.x, dt[1] - dt[0], dt[2] - dt[1], dt[3] - dt[2],
               dt[4] - dt[3], dt[5] - dt[4]);
#endif
    // ---- tile end: warp arg-min per key (CREDUX.MIN), lane k keeps key k, then all lanes publish in parallel
    unsigned long long* bestJob = p.best + (size_t)job * HMME_NPARTS;
    uint32_t mine = 0xFFFFFFFFu, last = 0xFFFFFFFFu;
#pragma unroll
    for (int k = 0; k < 33; ++k) {
        const uint32_t m = __reduce_min_sync(0xFFFFFFFFu, best[k]);
        if (k < 32) mine = (lane == k) ? m : mine;
        else last = m;
    }
    publish<YB>(bestJob, block_part_index(b, lane), mine, tg, p.W);
    if (lane == 0) publish<YB>(bestJob, block_part_index(b, 32), last, tg, p.W);
    if (role <= 4) {                                          // warp-uniform: SLOTS is a multiple of 32
        mine = 0xFFFFFFFFu;
#pragma unroll
        for (int k = 0; k < 13; ++k) {
            const uint32_t m = __reduce_min_sync(0xFFFFFFFFu, ub[k]);
            mine = (lane == k) ? m : mine;
        }
        if (lane < 13) publish<YB>(bestJob, role < 4 ? quad_part_index(role, lane) : ctu_part_index(lane), mine, tg, p.W);
    }
}

}  // namespace hmme
