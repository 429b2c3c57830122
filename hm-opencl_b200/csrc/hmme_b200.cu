// hmme_b200.cu -- host side of libhmme_b200.so (C ABI in include/hmme_b200.h).
//
// Replaces the host orchestration of /root/reference/source/Lib/TLibEncoder/TEncOpenCL.cpp:
//   findDevice/compileKernelSource/createBuffers (:69-238)  -> hmme_create
//   calcMotionVectors (:240-362)                             -> hmme_search_ctu (sync) / hmme_search_frame (batched)
//   xFillSADBuffer/xResetArrays (:366-392)                   -> arg-min scratch kept in its reset state by the kernels
// and, one step further along the encoder's path,
//   TEncSearch::xPatternSearchFracDIF (TEncSearch.cpp:4294-4331) -> hmme_refine_frac (PU list) / hmme_refine_frame (all partitions)
// No OpenCL, no runtime compilation, no CPU fallback: every failure is an error code + message.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/hmme_b200.h"
#include "hmme_internal.cuh"
#include "me_common.cuh"
#include "me_frac_kernel.cuh"
#include "me_generic_kernel.cuh"
#include "me_u8_kernel.cuh"

#ifndef HMME_FAST_YB
#define HMME_FAST_YB 3
#endif

using namespace hmme;

namespace {

std::mutex g_errMu;
std::string g_createErr;

struct FastGeom { int tw, upt, tileRows, nTx, nTy; size_t smemBytes; };

// cuTensorMapEncodeTiled through the runtime's driver entry point query: no link-time dependency on libcuda
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
    static const EncodeTiledFn fn = [] {
        if (std::getenv("HMME_NO_TMA2D")) return (EncodeTiledFn) nullptr;      // experiments: per-row copies only
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return (EncodeTiledFn) nullptr;
        return (EncodeTiledFn)f;
    }();
    return fn;
}
// u8 rows of `pitch` bytes from `base`, box = boxW x boxH bytes; false when the layout does not meet the copy engine's alignment rules
bool make_u8_map(CUtensorMap* m, const void* base, long long pitch, long long rows, int boxW, int boxH) {
    const EncodeTiledFn fn = encode_tiled_fn();
    if (!fn || ((uintptr_t)base & 15) || (pitch & 15) || pitch < boxW || rows < 1 || boxW > 256 || boxH > 256) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)rows}, strides[1] = {(cuuint64_t)pitch};
    const cuuint32_t box[2] = {(cuuint32_t)boxW, (cuuint32_t)boxH}, es[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

constexpr size_t kSmemBudget = 224 * 1024;   // of the 227 KB a CTA may opt in to (1 CTA per SM anyway: 128 registers x 512 threads)

size_t fast_smem_bytes(int tileRows, int yb) {
    const size_t words = (((size_t)fast_win_rows(tileRows) * kWinPitch + 31) & ~(size_t)31) + 1024 + kRing * (size_t)(32 * yb) * kRecWords +
                         (size_t)tileRows * (kKbPitch + 2) + kMaxTileW + 4;
    return words * 4;
}

// row groups a tile of `upt` consecutive units can touch in a stripe `w` columns wide (worst case over the tile's start column)
int span_rgs(int upt, int w) { return upt % w == 0 ? upt / w : (upt + 2 * w - 2) / w; }

// Tile geometry of the packed kernel for a (2R+1)^2 window and njobs jobs on `sms` SMs (1 CTA per SM).
// x: column stripes of at most 129 candidates.  Within a stripe the units (column x row group of YB candidate rows) are numbered row
// group by row group and cut into tiles of `upt` consecutive units, upt * YB <= 2048 (11-bit rank in the key) and the window of the
// row groups a tile touches within the shared-memory budget.
//  * Frame-sized launches (at least one CTA per SM with the largest tiles): upt is a multiple of 32 -- every round of every tile but the
//    stripe's last is full -- chosen so that the tiles of a stripe are as equal as possible (+-64: 5547 units = 8 x 640 + 427, 174 rounds
//    per job and warp instead of the 181 that 129 x 15 rectangles need).
//  * Small launches -- the per-CTU call of the encoder is ONE job -- use rectangular tiles (upt = columns x k row groups) and also split
//    the columns, with the k of the smallest predicted makespan, which spreads a single job over up to 129 SMs:
//      waves(k) * (rounds(k) + c),  waves = ceil(CTAs / sms),  rounds = ceil(tw * k / 32),  c ~ 4 rounds of per-tile overhead
FastGeom fast_geometry(int W, int yb, int njobs, int sms, int forceRG) {
    FastGeom g{};
    const int nTxMin = (W + kMaxTileW - 1) / kMaxTileW;
    const int nRGjob = (W + yb - 1) / yb;
    auto tile_rows = [&](int upt, int nx, int tw) {
        const int twLast = W - (nx - 1) * tw;
        return yb * std::min(nRGjob, std::max(span_rgs(upt, tw), span_rgs(upt, twLast)));
    };
    auto fits = [&](int upt, int nx, int tw) { return upt * yb <= kMaxTileCands && fast_smem_bytes(tile_rows(upt, nx, tw), yb) <= kSmemBudget; };
    auto max_rg = [&](int nx, int tw) {
        int m = std::max(1, std::min(kMaxTileCands / (tw * yb), nRGjob));
        while (m > 1 && !fits(m * tw, nx, tw)) --m;
        return m;
    };
    int bestNx = nTxMin, bestTw = (W + nTxMin - 1) / nTxMin;
    int bestK = max_rg(bestNx, bestTw), upt = 0;
    const bool fullWave = (long long)njobs * nTxMin * ((nRGjob + bestK - 1) / bestK) >= sms;
    if (forceRG >= 1 && forceRG <= bestK) upt = forceRG * bestTw;
    else if (fullWave) {
        const int total = nRGjob * bestTw;
        int uMax = (kMaxTileCands / yb) & ~31;
        while (uMax > 32 && !fits(uMax, bestNx, bestTw)) uMax -= 32;
        const int nT = (total + uMax - 1) / uMax;
        upt = std::min(uMax, (((total + nT - 1) / nT) + 31) & ~31);
        if (total <= 32 || !fits(upt, bestNx, bestTw)) upt = bestK * bestTw;     // tiny windows: one rectangular tile
    } else {
        // Small launch (the encoder's per-CTU call is ONE job): split columns as well as rows so that the job reaches every SM.
        // cost = waves * (rounds + c): rounds = ceil(units / 32) with units = tw * k (lane = candidate column x row group),
        // c ~ 4 rounds of per-tile staging.  +-64, one job: 129 x 3 tiles (43 CTAs, 5 rounds) -> 43 x 3 tiles (129 CTAs, 2 rounds).
        double bestCost = 1e300;
        for (int nx = nTxMin; nx <= 4 * nTxMin && nx <= W; ++nx) {
            const int tw = (W + nx - 1) / nx;
            if ((W + tw - 1) / tw != nx) continue;                        // this column count does not change the tile width
            const int maxRG = max_rg(nx, tw);
            for (int k = maxRG; k >= 1; --k) {
                const long long ctas = (long long)njobs * nx * ((nRGjob + k - 1) / k);
                const double cost = (double)((ctas + sms - 1) / sms) * ((tw * k + 31) / 32 + 4.0);
                if (cost < bestCost * 0.995) { bestCost = cost; bestK = k; bestNx = nx; bestTw = tw; }   // prefer the larger tile unless clearly worse
            }
        }
        upt = bestK * bestTw;
    }
    g.nTx = bestNx; g.tw = bestTw; g.upt = upt;
    g.tileRows = tile_rows(upt, bestNx, bestTw);
    g.nTy = (nRGjob * bestTw + upt - 1) / upt;
    g.smemBytes = fast_smem_bytes(g.tileRows, yb);
    return g;
}

}  // namespace

int hmme_fail(hmme_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg;
    else { std::lock_guard<std::mutex> l(g_errMu); g_createErr = msg; }
    return code;
}

namespace {

inline int fail(hmme_ctx* c, int code, const std::string& msg) { return hmme_fail(c, code, msg); }

int ensure_jobs(hmme_ctx* c, size_t njobs) {
    if (njobs <= c->jobCap) return HMME_OK;
    size_t cap = std::max<size_t>(njobs, std::max<size_t>(64, c->jobCap * 2));
    for (cudaStream_t s : {c->ioStream[0], c->ioStream[1], c->stream}) CU_TRY(c, cudaStreamSynchronize(s));
    if (c->dJobs) cudaFree(c->dJobs);
    if (c->dBest) cudaFree(c->dBest);
    if (c->dRes) cudaFree(c->dRes);
    if (c->hJobs) cudaFreeHost(c->hJobs);
    c->dJobs = nullptr; c->dBest = nullptr; c->dRes = nullptr; c->hJobs = nullptr; c->jobCap = 0;
    CU_TRY(c, cudaMalloc(&c->dJobs, cap * sizeof(int4)));
    CU_TRY(c, cudaMalloc(&c->dBest, cap * HMME_NPARTS * sizeof(unsigned long long)));
    me_init_kernel<<<(unsigned)((cap * HMME_NPARTS + 255) / 256), 256, 0, c->stream>>>(c->dBest, cap * HMME_NPARTS);
    CU_TRY(c, cudaMalloc(&c->dRes, 4 * cap * HMME_NPARTS * sizeof(int32_t)));
    CU_TRY(c, cudaMallocHost(&c->hJobs, cap * sizeof(hmme_job)));
    c->jobCap = cap;
    ++c->bufGen;                                            // graphs recorded against the old buffers must not be replayed
    return HMME_OK;
}

size_t plane_elems(const hmme_plane* p) { return (size_t)p->pitch * (size_t)(p->height + 2 * p->marginY); }

int check_plane(hmme_ctx* c, const hmme_plane* p, const char* what) {
    if (!p || !p->base || (p->elemBytes != 1 && p->elemBytes != 2) || p->width <= 0 || p->height <= 0 || p->marginX < 0 ||
        p->marginY < 0 || p->pitch < p->width + 2 * p->marginX)
        return fail(c, HMME_ERR_ARG, std::string("invalid plane descriptor: ") + what);
    if ((reinterpret_cast<uintptr_t>(p->base) & 15) != 0) return fail(c, HMME_ERR_ARG, std::string("plane base must be 16-byte aligned: ") + what);
    return HMME_OK;
}

// Every job's 64x64 block and (2R+64)^2 window must lie inside the allocations under LINEAR addressing
// (the reference reads pelSearchArray[j + i*iRefStride] regardless of row ends, App. B4).
int check_jobs(hmme_ctx* c, const hmme_plane* cur, const hmme_plane* ref, const hmme_job* jobs, int njobs, int R) {
    const long long curN = (long long)plane_elems(cur), refN = (long long)plane_elems(ref);
    for (int j = 0; j < njobs; ++j) {
        const hmme_job& b = jobs[j];
        const long long c0 = (long long)(cur->marginY + b.ctuY) * cur->pitch + cur->marginX + b.ctuX;
        const long long c1 = c0 + 63LL * cur->pitch + 63;
        const long long r0 = (long long)(ref->marginY + b.ctuY + b.lty) * ref->pitch + ref->marginX + b.ctuX + b.ltx;
        const long long r1 = r0 + (long long)(2 * R + 63) * ref->pitch + 2 * R + 63;
        if (c0 < 0 || c1 >= curN || b.ctuX < -cur->marginX || b.ctuX + 64 > cur->width + cur->marginX)
            return fail(c, HMME_ERR_BOUNDS, "job " + std::to_string(j) + ": CTU outside the current plane");
        if (r0 < 0 || r1 >= refN)
            return fail(c, HMME_ERR_BOUNDS, "job " + std::to_string(j) + ": search window leaves the reference allocation");
    }
    return HMME_OK;
}

const char* origin_ptr(const hmme_plane* p) {
    return static_cast<const char*>(p->base) + ((size_t)p->marginY * p->pitch + p->marginX) * p->elemBytes;
}

int ensure_pus(hmme_ctx* c, size_t npus) {
    if (!c->evF0) { CU_TRY(c, cudaEventCreate(&c->evF0)); CU_TRY(c, cudaEventCreate(&c->evF1)); }
    if (!c->dOrder) {
        // partitions by area, large to small (stable), so that consecutive PUs of the whole-frame list cost about the same
        std::vector<int> order(HMME_NPARTS);
        for (int i = 0; i < HMME_NPARTS; ++i) order[i] = i;
        auto tiles = [](int q) { const PartRect r = part_rect(q); return ((r.w + 7) / 8) * ((r.h + 7) / 8); };
        auto seg = [](int q) { const PartRect r = part_rect(q); return frac_segment(r.w, r.h); };
        // by segment of the group kernel (which starts with the PUs of four tiles or more), inside a segment by tile count
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return seg(a) != seg(b) ? seg(a) < seg(b) : tiles(a) > tiles(b); });
        c->bigParts = 0; c->tilesPerCtu = 0;
        for (int q = 0; q < 5; ++q) c->segParts[q] = 0;
        for (int q : order) { c->bigParts += tiles(q) >= kFracCoopTiles; c->tilesPerCtu += tiles(q); c->segParts[seg(q)] += 1; }
        CU_TRY(c, cudaMalloc(&c->dOrder, HMME_NPARTS * sizeof(int)));
        CU_TRY(c, cudaMemcpy(c->dOrder, order.data(), HMME_NPARTS * sizeof(int), cudaMemcpyHostToDevice));
    }
    if (npus <= c->puCap) return HMME_OK;
    const size_t cap = std::max<size_t>(npus, std::max<size_t>(1024, c->puCap * 2));
    for (cudaStream_t s : {c->ioStream[0], c->ioStream[1], c->stream}) CU_TRY(c, cudaStreamSynchronize(s));
    cudaFree(c->dPus); cudaFree(c->dSlots); cudaFree(c->dFrac); cudaFree(c->dCand);
    c->dPus = nullptr; c->dSlots = nullptr; c->dFrac = nullptr; c->dCand = nullptr; c->puCap = 0;
    CU_TRY(c, cudaMalloc(&c->dPus, cap * sizeof(FracPu)));
    CU_TRY(c, cudaMalloc(&c->dSlots, cap * sizeof(int)));
    CU_TRY(c, cudaMalloc(&c->dFrac, cap * sizeof(int4)));
    c->puCap = cap;
    ++c->bufGen;
    return HMME_OK;
}

// the 8-tap filters read 4 samples around the MV-displaced PU; the kernel fetches whole 16x16 patches per 8x8 tile
int check_pus(hmme_ctx* c, const hmme_plane* cur, const hmme_plane* ref, const hmme_pu* pus, int npus) {
    for (int n = 0; n < npus; ++n) {
        const hmme_pu& u = pus[n];
        if (u.w <= 0 || u.h <= 0 || u.w > 64 || u.h > 64 || (u.w & 3) || (u.h & 3))
            return fail(c, HMME_ERR_ARG, "PU " + std::to_string(n) + ": width/height must be multiples of 4 in [4,64]");
        if (u.x < -cur->marginX || u.y < -cur->marginY || u.x + u.w > cur->width + cur->marginX || u.y + u.h > cur->height + cur->marginY)
            return fail(c, HMME_ERR_BOUNDS, "PU " + std::to_string(n) + ": outside the current plane");
        const int w8 = (u.w + 7) & ~7, h8 = (u.h + 7) & ~7;
        if (u.x + u.mvx - 4 < -ref->marginX || u.y + u.mvy - 4 < -ref->marginY || u.x + u.mvx + w8 + 4 > ref->width + ref->marginX ||
            u.y + u.mvy + h8 + 4 > ref->height + ref->marginY)
            return fail(c, HMME_ERR_BOUNDS, "PU " + std::to_string(n) + ": interpolation apron leaves the reference plane (needs 4 samples + tile padding)");
    }
    return HMME_OK;
}

// "last reader" marker: uploads enqueued later wait for everything the compute stream holds up to here (searches, refinements
// and distortion kernels all read planes an upload may overwrite)
int mark_compute(hmme_ctx* c) {
    if (!c->capturing) CU_TRY(c, cudaEventRecord(c->evCompute, c->stream));
    return HMME_OK;
}

// Order of a PU list for the group kernels: by segment (hmme::frac_segment), the PUs of four tiles or more by tile count, large to small;
// stable inside equal keys.  Counting sort: the lists have a few hundred thousand entries.  wh(n, w, h) reports PU n's size.
template <typename WH>
void segment_order(int npus, WH wh, std::vector<int>& idx, long long (&segCount)[5], int& nBig, long long& totalTiles) {
    auto key = [&](int n, int& tiles) {
        int w, h;
        wh(n, w, h);
        tiles = ((w + 7) / 8) * ((h + 7) / 8);
        const int sg = frac_segment(w, h);
        return sg == 0 ? 64 - std::min(tiles, 64) : 60 + sg;   // 0..60 for 64..4 tiles, 61..64 for segments 1..4
    };
    std::vector<int> start(66, 0);
    nBig = 0; totalTiles = 0;
    for (int n = 0, t; n < npus; ++n) { start[key(n, t) + 1] += 1; nBig += t >= kFracCoopTiles; totalTiles += t; }
    for (int q = 0; q < 5; ++q) segCount[q] = 0;
    for (int k = 0; k < 65; ++k) segCount[k <= 60 ? 0 : k - 60] += start[k + 1];
    for (int k = 0; k < 65; ++k) start[k + 1] += start[k];
    idx.resize(npus);
    for (int n = 0, t; n < npus; ++n) idx[start[key(n, t)]++] = n;
}

// segment boundaries of the group kernel from the number of PUs in each segment (the list is ordered by segment)
void frac_segments(const long long (&cnt)[5], FracGroupParams& fp) {
    fp.segPu[0] = 0; fp.segGrp[0] = 0;
    for (int q = 0; q < 5; ++q) {
        fp.segPu[q + 1] = fp.segPu[q] + (int)cnt[q];
        fp.segGrp[q + 1] = fp.segGrp[q] + (int)((cnt[q] + kFracSegPus[q] - 1) / kFracSegPus[q]);
    }
}

int enqueue_frac(hmme_ctx* c, const hmme_plane* cur, const hmme_plane* ref, int npus, int nBig, long long totalTiles, bool slots, int useHad, bool wantCand,
                 const long long (&segCount)[5]) {
    if (wantCand && !c->dCand) CU_TRY(c, cudaMalloc(&c->dCand, c->puCap * 18 * sizeof(uint32_t)));
    FracGroupParams fp{};
    fp.cur = origin_ptr(cur); fp.ref = reinterpret_cast<const uint8_t*>(origin_ptr(ref));
    fp.curPitch = cur->pitch; fp.refPitch = ref->pitch; fp.curBytes = cur->elemBytes;
    // Large PUs get a CTA each only when the batch is too small for longest-first scheduling to hide a 64-tile PU running on one
    // warp (measured at 1080p: 8 / 30 CTU jobs 0.094 -> 0.035 ms / 0.115 -> 0.086 ms, but 60 jobs 0.156 -> 0.160 and 480 jobs 1.14 -> 1.21)
    const char* coopStr = std::getenv("HMME_FRAC_COOP");   // experiments and tests: 0 never, 1 always (read per call, so a test can switch forms)
    const int coopEnv = coopStr ? std::atoi(coopStr) : -1;
    const bool coop = nBig > 0 && (coopEnv >= 0 ? coopEnv != 0 : totalTiles < 32LL * c->prop.multiProcessorCount * 16);
    if (!coop) nBig = 0;
    fp.pus = c->dPus; fp.slots = slots ? c->dSlots : nullptr; fp.npus = npus; fp.nBig = nBig;
    fp.lambda = c->lambda; fp.useHad = useHad ? 1 : 0;
    fp.out = c->dFrac; fp.cand = wantCand ? c->dCand : nullptr;
    // about one PU per warp: CTAs start in index order, so with the list ordered large to small the hardware's CTA scheduler does
    // longest-first load balancing (measured 1080p: 16 CTAs per SM 1.25 ms, 32: 1.20, 128 and more: 1.14; one resident wave: 1.30)
    static const int perSm = std::getenv("HMME_FRAC_CTAS_PER_SM") ? std::max(1, std::atoi(std::getenv("HMME_FRAC_CTAS_PER_SM"))) : 256;
    frac_segments(segCount, fp);
    const char* formStr = std::getenv("HMME_FRAC_FORM");   // experiments and tests: 1 = one PU per warp (round 1's kernel)
    const int formEnv = formStr ? std::atoi(formStr) : 2;
    const bool group = !coop && formEnv != 1;
    const int units = group ? fp.segGrp[5] : npus - nBig;   // one group / one PU per warp
    const int ctas = std::max(1, nBig + std::min((units + kFracWarps - 1) / kFracWarps, c->prop.multiProcessorCount * perSm));
    if (!c->capturing) CU_TRY(c, cudaEventRecord(c->evF0, c->stream));
    if (coop) me_frac_coop_kernel<<<ctas, kFracThreads, 0, c->stream>>>(fp);
    else if (group) me_frac_group_kernel<<<ctas, kFracThreads, 0, c->stream>>>(fp);
    else me_frac_kernel<<<ctas, kFracThreads, 0, c->stream>>>(fp);
    if (!c->capturing) CU_TRY(c, cudaEventRecord(c->evF1, c->stream));
    c->evFracValid = !c->capturing;
    c->launches += 1;
    CU_TRY(c, cudaGetLastError());
    return mark_compute(c);
}

int check_frac_planes(hmme_ctx* c, const hmme_plane* cur, const hmme_plane* ref) {
    int rc = check_plane(c, cur, "current");
    if (rc == HMME_OK) rc = check_plane(c, ref, "reference");
    if (rc != HMME_OK) return rc;
    if (ref->elemBytes != 1)
        return fail(c, HMME_ERR_ARG, "fractional refinement needs an 8-bit reference plane (the path is defined for 8-bit video only)");
    return HMME_OK;
}

template <typename TC, typename TR>
void launch_generic(hmme_ctx* c, const GenericParams& gp, int njobs) {
    me_generic_kernel<TC, TR><<<njobs * gp.nChunks, kGenThreads, 0, c->stream>>>(gp);
}

// Where a search reads its jobs and leaves its results.  Default: the context's own buffers ([4][jobCap][593]); a device-resident
// table slot (hmme_table) or the compact per-CTU buffer of hmme_search_ctu otherwise.
struct SearchIO {
    const int4* jobs; int32_t* X; int32_t* Y; uint32_t* S; uint32_t* Cst;
    bool finalizeInline;       // finalisation on the compute stream itself (latency path: no cross-stream events)
};

SearchIO default_io(hmme_ctx* c) {
    SearchIO io{};
    io.jobs = c->dJobs;
    io.X = c->dRes; io.Y = io.X + c->jobCap * HMME_NPARTS;
    io.S = reinterpret_cast<uint32_t*>(io.Y + c->jobCap * HMME_NPARTS); io.Cst = io.S + c->jobCap * HMME_NPARTS;
    io.finalizeInline = false;
    return io;
}

// Enqueue search -> finalize for njobs jobs already present at io.jobs.  The arg-min scratch is in its reset state before and
// after (me_init_kernel once per allocation, me_finalize_kernel restores it).
int enqueue_search(hmme_ctx* c, const SearchIO& io, const void* curOrigin, int curElem, long long curPitch, const void* refOrigin, int refElem,
                   long long refPitch, const void* refLo, const void* refHi, int njobs, int R) {
    const int W = 2 * R + 1;
    const size_t nres = (size_t)njobs * HMME_NPARTS;
    // A 16-bit block against an 8-bit picture (the bi-prediction refinement): clamped block + per-partition constants, then the
    // packed 8-bit kernel (me_bipred_prep_kernel explains why that is exact)
    const bool bipred = curElem == 2 && refElem == 1;
    if (!c->capturing) CU_TRY(c, cudaEventRecord(c->ev0, c->stream));   // timing events stay out of captured graphs
    if (bipred) {
        if ((size_t)njobs > c->biCap) {
            if (c->capturing) return fail(c, HMME_ERR_ARG, "bi-prediction buffers would grow while capturing (run the step once first)");
            const size_t cap = std::max<size_t>((size_t)njobs, std::max<size_t>(64, c->biCap * 2));
            for (cudaStream_t s : {c->ioStream[0], c->ioStream[1], c->stream}) CU_TRY(c, cudaStreamSynchronize(s));
            cudaFree(c->dBiBlocks); cudaFree(c->dBiOffsets);
            c->dBiBlocks = nullptr; c->dBiOffsets = nullptr; c->biCap = 0;
            CU_TRY(c, cudaMalloc(&c->dBiBlocks, cap * 4096 + 64));
            CU_TRY(c, cudaMalloc(&c->dBiOffsets, cap * HMME_NPARTS * sizeof(uint32_t)));
            c->biCap = cap;
            ++c->bufGen;
        }
        me_bipred_prep_kernel<<<njobs, 256, 0, c->stream>>>(static_cast<const int16_t*>(curOrigin), curPitch, io.jobs, c->dBiBlocks, c->dBiOffsets);
        c->launches += 1;
    }
    if ((curElem == 1 || bipred) && refElem == 1) {
        constexpr int yb = HMME_FAST_YB;             // candidate rows per thread (3: measured best; 2 is 17 % slower, 4 does not fit)
        const FastGeom g = fast_geometry(W, yb, njobs, c->prop.multiProcessorCount, c->forceRG);
        FastParams fp{};
        fp.cur = static_cast<const uint8_t*>(curOrigin);
        fp.curBlocks = bipred ? c->dBiBlocks : nullptr;
        fp.ref = static_cast<const uint8_t*>(refOrigin);
        fp.refLo = static_cast<const uint8_t*>(refLo);
        fp.refHi = static_cast<const uint8_t*>(refHi);
        fp.curPitch = curPitch; fp.refPitch = refPitch;
        fp.jobs = io.jobs; fp.best = c->dBest; fp.lambda = c->lambda; fp.W = W;
        fp.tw = g.tw; fp.upt = g.upt; fp.tileRows = g.tileRows; fp.nTx = g.nTx; fp.nTy = g.nTy; fp.stagger = c->stagger;
        auto magic = [](int d) { return d <= 1 ? 0u : (uint32_t)(0x100000000ull / (unsigned)d) + 1u; };   // see fast_div
        fp.magicTiles = magic(g.nTx * g.nTy); fp.magicNTx = magic(g.nTx); fp.magicTw = magic(g.tw); fp.magicTwLast = magic(W - (g.nTx - 1) * g.tw);
        // one 2-D copy per window / block where the planes allow it (the kernel decides per tile, see FastParams)
        const long long refOff = fp.ref - fp.refLo;
        if (refOff >= 0 && refPitch > 0 &&
            make_u8_map(&fp.refMap, fp.refLo, refPitch, (fp.refHi - fp.refLo) / refPitch, kDensePitch, fast_win_rows(g.tileRows))) {
            fp.refMapOk = 1;
            fp.refRow0 = (int)(refOff / refPitch); fp.refCol0 = (int)(refOff % refPitch);
            fp.refMapRows = (int)std::min<long long>((fp.refHi - fp.refLo) / refPitch, 1 << 30);
        }
        if (!bipred && curPitch > 0 && make_u8_map(&fp.curMap, fp.cur, curPitch, 1 << 20, 64, 64)) fp.curMapOk = 1;
        const unsigned grid = (unsigned)(njobs * g.nTx * g.nTy);
        if (g.smemBytes > c->fastSmemSet) {           // the opt-in only ever has to grow
            CU_TRY(c, cudaFuncSetAttribute(me_u8_tile_kernel<HMME_FAST_YB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget));
            c->fastSmemSet = kSmemBudget;
        }
        me_u8_tile_kernel<HMME_FAST_YB><<<grid, kFastThreads, g.smemBytes, c->stream>>>(fp);
    } else {
        GenericParams gp{};
        gp.cur = curOrigin; gp.ref = refOrigin; gp.curPitch = curPitch; gp.refPitch = refPitch;
        gp.jobs = io.jobs; gp.best = c->dBest; gp.lambda = c->lambda; gp.W = W;
        const int nCand = W * W;
        int nChunks = std::max(1, std::min((nCand + 63) / 64, (4 * c->prop.multiProcessorCount + njobs - 1) / njobs));
        gp.chunk = (nCand + nChunks - 1) / nChunks;
        gp.chunk = (gp.chunk + kGenBatch - 1) / kGenBatch * kGenBatch;
        gp.nChunks = (nCand + gp.chunk - 1) / gp.chunk;
        if (curElem == 1) launch_generic<uint8_t, int16_t>(c, gp, njobs);     // non-8-bit reference planes: outside the reference's defined
        else launch_generic<int16_t, int16_t>(c, gp, njobs);                  // domain (App. A.2), kept exact and slow
    }
    CU_TRY(c, cudaEventRecord(c->ev1, c->stream));
    c->evValid = !c->capturing;
    c->launches += 2;
    if (io.finalizeInline) {
        me_finalize_kernel<<<(unsigned)((nres + 255) / 256), 256, 0, c->stream>>>(c->dBest, io.jobs, njobs, W, c->lambda, io.X, io.Y, io.S, io.Cst, bipred ? c->dBiOffsets : nullptr);
        CU_TRY(c, cudaGetLastError());
        return mark_compute(c);
    }
    // Finalisation (keys -> X, Y, sad, cost; keys handed back in their reset state) runs on a high-priority stream: a tiny
    // kernel on the compute stream would queue behind every pending CTA of another context's search, delaying the result
    // copy -- and with it the host -- by a whole search.
    CU_TRY(c, cudaStreamWaitEvent(c->ioStream[0], c->ev1, 0));
    me_finalize_kernel<<<(unsigned)((nres + 255) / 256), 256, 0, c->ioStream[0]>>>(c->dBest, io.jobs, njobs, W, c->lambda, io.X, io.Y, io.S, io.Cst, bipred ? c->dBiOffsets : nullptr);
    CU_TRY(c, cudaEventRecord(c->evFinal, c->ioStream[0]));
    CU_TRY(c, cudaStreamWaitEvent(c->stream, c->evFinal, 0));
    CU_TRY(c, cudaGetLastError());
    return mark_compute(c);
}

// stream sync + deferred content check of asynchronous 8-bit uploads
int sync_ctx(hmme_ctx* c) {
    CU_TRY(c, cudaStreamSynchronize(c->ioStream[0]));
    CU_TRY(c, cudaStreamSynchronize(c->ioStream[1]));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    if (c->contentCheckPending) {
        c->contentCheckPending = false;
        CU_TRY(c, cudaMemcpy(c->hFlag, c->dFlag, sizeof(int), cudaMemcpyDeviceToHost));
        if (*c->hFlag) {
            CU_TRY(c, cudaMemset(c->dFlag, 0, sizeof(int)));
            return fail(c, HMME_ERR_CONTENT, "plane declared 8-bit holds samples outside [0,255]; allocate it with elemBytes=2");
        }
    }
    return HMME_OK;
}

int fetch_async(hmme_ctx* c, int njobs, int32_t* X, int32_t* Y, uint32_t* sad, uint32_t* cost) {
    // straight into the caller's arrays: a true DMA when they are page-locked (bench.py), a driver-staged copy otherwise
    const size_t n = (size_t)njobs * HMME_NPARTS, cap = c->jobCap * HMME_NPARTS;
    void* outs[4] = {X, Y, sad, cost};
    for (int k = 0; k < 4; ++k)
        if (outs[k]) CU_TRY(c, cudaMemcpyAsync(outs[k], c->dRes + k * cap, n * 4, cudaMemcpyDeviceToHost, c->stream));
    return HMME_OK;
}

int fetch(hmme_ctx* c, int njobs, int32_t* X, int32_t* Y, uint32_t* sad, uint32_t* cost) {
    const int rc = fetch_async(c, njobs, X, Y, sad, cost);
    return rc != HMME_OK ? rc : sync_ctx(c);
}

}  // namespace

extern "C" {

const char* hmme_version(void) { return "hmme_b200 0.1 (sm_100a)"; }

int hmme_device_count(int* count) {
    if (!count) return HMME_ERR_ARG;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { *count = 0; return fail(nullptr, HMME_ERR_NO_DEVICE, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e)); }
    *count = n;
    return HMME_OK;
}

int hmme_create(hmme_ctx** out, int device, int maxCtuW, int maxCtuH, int maxSearchRange) {
    if (!out) return HMME_ERR_ARG;
    *out = nullptr;
    if (maxCtuW != HMME_CTU_SIZE || maxCtuH != HMME_CTU_SIZE)
        return fail(nullptr, HMME_ERR_RANGE, "only 64x64 CTUs are defined for this path (TEncSearch.cpp:3745)");
    if (maxSearchRange < 0 || maxSearchRange > 1024) return fail(nullptr, HMME_ERR_RANGE, "search range out of [0,1024]");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(nullptr, HMME_ERR_NO_DEVICE, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "count is 0"));
    if (device < 0 || device >= n) return fail(nullptr, HMME_ERR_NO_DEVICE, "device index out of range");
    hmme_ctx* c = new hmme_ctx;
    c->device = device;
    auto bail = [&](const std::string& m, int code) { g_createErr = m; hmme_destroy(c); return code; };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail(std::string("cudaSetDevice: ") + cudaGetErrorString(e), HMME_ERR_CUDA);
    if ((e = cudaGetDeviceProperties(&c->prop, device)) != cudaSuccess) return bail(std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e), HMME_ERR_CUDA);
    if (c->prop.major != 10)
        return bail(std::string("device '") + c->prop.name + "' is sm_" + std::to_string(c->prop.major) + std::to_string(c->prop.minor) +
                    "; this library carries sm_100a code only and has no fallback", HMME_ERR_NO_DEVICE);
    int prLo = 0, prHi = 0;
    cudaDeviceGetStreamPriorityRange(&prLo, &prHi);
    if ((e = cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prLo)) != cudaSuccess ||
        (e = cudaStreamCreateWithPriority(&c->ioStream[0], cudaStreamNonBlocking, prHi)) != cudaSuccess ||
        (e = cudaStreamCreateWithPriority(&c->ioStream[1], cudaStreamNonBlocking, prHi)) != cudaSuccess)
        return bail(std::string("cudaStreamCreate: ") + cudaGetErrorString(e), HMME_ERR_CUDA);
    if ((e = cudaEventCreate(&c->ev0)) != cudaSuccess || (e = cudaEventCreate(&c->ev1)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&c->evUpload[0], cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&c->evUpload[1], cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&c->evCompute, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&c->evFinal, cudaEventDisableTiming)) != cudaSuccess)
        return bail(std::string("cudaEventCreate: ") + cudaGetErrorString(e), HMME_ERR_CUDA);
    c->maxRange = maxSearchRange;
    if (const char* e = std::getenv("HMME_FAST_RG")) c->forceRG = std::atoi(e);
    if (const char* e = std::getenv("HMME_STAGGER")) c->stagger = std::max(0, std::atoi(e));
    const size_t side = (size_t)2 * maxSearchRange + 64 + 16;
    c->winElems = std::max(side * side, (size_t)96 * 80) + 4128;   // + {job | 64x64 block} of hmme_search_ctu; also stages the one or two 96-pitch reference patches of the per-PU calls
    if ((e = cudaMallocHost(&c->hWin, c->winElems * 2)) != cudaSuccess || (e = cudaMalloc(&c->dWin, c->winElems * 2 + 64)) != cudaSuccess ||
        (e = cudaMallocHost(&c->hCtuRes, 4 * HMME_NPARTS * sizeof(int32_t))) != cudaSuccess || (e = cudaMalloc(&c->dCtuRes, 4 * HMME_NPARTS * sizeof(int32_t))) != cudaSuccess ||
        (e = cudaMallocHost(&c->hCurBlk, 4096 * 2)) != cudaSuccess || (e = cudaMalloc(&c->dCurBlk, 4096 * 2)) != cudaSuccess ||
        (e = cudaMalloc(&c->dFlag, sizeof(int))) != cudaSuccess || (e = cudaMallocHost(&c->hFlag, sizeof(int))) != cudaSuccess)
        return bail(std::string("buffer allocation: ") + cudaGetErrorString(e), HMME_ERR_CUDA);
    if ((e = cudaMemset(c->dFlag, 0, sizeof(int))) != cudaSuccess) return bail(std::string("cudaMemset: ") + cudaGetErrorString(e), HMME_ERR_CUDA);
    int rc = ensure_jobs(c, 64);
    if (rc != HMME_OK) return bail(c->err, rc);
    *out = c;
    return HMME_OK;
}

void hmme_destroy(hmme_ctx* c) {
    if (!c) return;
    if (c->device >= 0) cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    cudaFree(c->dJobs); cudaFree(c->dBest); cudaFree(c->dRes); cudaFreeHost(c->hJobs);
    cudaFreeHost(c->hWin); cudaFree(c->dWin); cudaFreeHost(c->hCurBlk); cudaFree(c->dCurBlk); cudaFreeHost(c->hCtuRes); cudaFree(c->dCtuRes);
    cudaFree(c->dBiBlocks); cudaFree(c->dBiOffsets);
    cudaFree(c->dStage[0]); cudaFree(c->dStage[1]); cudaFree(c->dFlag); cudaFreeHost(c->hFlag);
    cudaFree(c->dPus); cudaFree(c->dSlots); cudaFree(c->dFrac); cudaFree(c->dCand); cudaFree(c->dOrder); cudaFree(c->dPreds);
    if (c->evFork) cudaEventDestroy(c->evFork);
    for (int k = 0; k < 2; ++k) if (c->evJoin[k]) cudaEventDestroy(c->evJoin[k]);
    if (c->evF0) cudaEventDestroy(c->evF0);
    if (c->evF1) cudaEventDestroy(c->evF1);
    for (int k = 0; k < 2; ++k) {
        if (c->evUpload[k]) cudaEventDestroy(c->evUpload[k]);
        if (c->ioStream[k]) { cudaStreamSynchronize(c->ioStream[k]); cudaStreamDestroy(c->ioStream[k]); }
    }
    if (c->evCompute) cudaEventDestroy(c->evCompute);
    if (c->evFinal) cudaEventDestroy(c->evFinal);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

// page-locked host memory for callers that do not link the CUDA runtime themselves (asynchronous copies need it to be truly asynchronous)
void* hmme_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void hmme_host_free(void* p) { if (p) cudaFreeHost(p); }

const char* hmme_device_name(hmme_ctx* c) { return c ? c->prop.name : ""; }
const char* hmme_last_error(hmme_ctx* c) {
    if (c) return c->err.c_str();
    static thread_local std::string copy;
    std::lock_guard<std::mutex> l(g_errMu);
    copy = g_createErr;
    return copy.c_str();
}
void* hmme_stream(hmme_ctx* c) { return c ? (void*)c->stream : nullptr; }

int hmme_set_lambda(hmme_ctx* c, double lambda) {
    if (!c || !(lambda >= 0.0)) return fail(c, HMME_ERR_ARG, "lambda must be >= 0");
    c->lambda = (uint32_t)std::floor(65536.0 * std::sqrt(lambda));   // TEncOpenCL.h:121
    return HMME_OK;
}
int hmme_set_lambda_q16(hmme_ctx* c, uint32_t v) { if (!c) return HMME_ERR_ARG; c->lambda = v; return HMME_OK; }
uint32_t hmme_get_lambda_q16(hmme_ctx* c) { return c ? c->lambda : 0; }

// classify + narrow one row of int16 samples (vectorises: distinct source and destination)
static inline uint32_t narrow_row(const int16_t* __restrict__ src, uint8_t* __restrict__ dst, int n) {
    uint32_t o = 0;
    for (int q = 0; q < n; ++q) { o |= (uint16_t)src[q]; dst[q] = (uint8_t)src[q]; }
    return o;
}

int hmme_search_ctu(hmme_ctx* c, const int16_t* cur, int curStride, const int16_t* refAtCtu, int refStride, int range, int ltx,
                    int lty, int32_t* X, int32_t* Y, uint32_t* sad, uint32_t* cost) {
    if (!c) return HMME_ERR_ARG;
    if (!cur || !refAtCtu || !X || !Y || !sad || curStride < 64) return fail(c, HMME_ERR_ARG, "hmme_search_ctu: null pointer or stride < 64");
    if (range < 0 || range > c->maxRange) return fail(c, HMME_ERR_RANGE, "search range " + std::to_string(range) + " exceeds the context's " + std::to_string(c->maxRange));
    if (c->capturing) return fail(c, HMME_ERR_ARG, "hmme_search_ctu is synchronous: not allowed while capturing a graph");
    CU_TRY(c, cudaSetDevice(c->device));
    const int side = 2 * range + 64;                 // TEncOpenCL.cpp:256 areaStride
    const int wp = (side + 15) & ~15;
    // One page-locked staging block {job | 64x64 block | window rows}, gathered with the reference's linear addressing
    // (TEncOpenCL.cpp:251,275-277), narrowed to 8 bit on the way when the content allows (every encoder call except the
    // bi-prediction refinement, whose block is 2*org - pred), moved with ONE copy.
    char* hs = static_cast<char*>(c->hWin);
    char* ds = static_cast<char*>(c->dWin);
    constexpr size_t kCurOff = 64, kWinOff = 64 + 8192;
    const int16_t* w0 = refAtCtu + (ptrdiff_t)refStride * lty + ltx;
    uint32_t orCur = 0, orWin = 0;
    {
        uint8_t* hc8 = reinterpret_cast<uint8_t*>(hs + kCurOff);
        for (int r = 0; r < 64; ++r) orCur |= narrow_row(cur + (size_t)r * curStride, hc8 + r * 64, 64);
        uint8_t* hw8 = reinterpret_cast<uint8_t*>(hs + kWinOff);
        for (int r = 0; r < side; ++r) orWin |= narrow_row(w0 + (ptrdiff_t)r * refStride, hw8 + (size_t)r * wp, side);
    }
    const int curElem = (orCur & 0xFF00u) ? 2 : 1, elem = (orWin & 0xFF00u) ? 2 : 1;
    if (curElem == 2) {                                   // e.g. the bi-prediction block 2*org - pred
        int16_t* hc = reinterpret_cast<int16_t*>(hs + kCurOff);
        for (int r = 0; r < 64; ++r) std::memcpy(hc + r * 64, cur + (size_t)r * curStride, 128);
    }
    if (elem == 2) {                                      // not 8-bit video: outside the reference's defined domain, kept exact
        int16_t* hw = reinterpret_cast<int16_t*>(hs + kWinOff);
        for (int r = 0; r < side; ++r) std::memcpy(hw + (size_t)r * wp, w0 + (ptrdiff_t)r * refStride, (size_t)side * 2);
    }
    *reinterpret_cast<hmme_job*>(hs) = hmme_job{0, 0, ltx, lty};
    const size_t winBytes = (size_t)side * wp * elem;
    CU_TRY(c, cudaMemcpyAsync(ds, hs, kWinOff + winBytes, cudaMemcpyHostToDevice, c->stream));
    // virtual origin: sample (ctu + lt) is element 0 of the staged window
    const char* dWin = ds + kWinOff;
    const char* refOrigin = dWin - ((ptrdiff_t)lty * wp + ltx) * elem;
    SearchIO io{};
    io.jobs = reinterpret_cast<const int4*>(ds);
    io.X = c->dCtuRes; io.Y = io.X + HMME_NPARTS;
    io.S = reinterpret_cast<uint32_t*>(io.Y + HMME_NPARTS); io.Cst = io.S + HMME_NPARTS;
    io.finalizeInline = true;
    int rc = enqueue_search(c, io, ds + kCurOff, curElem, 64, refOrigin, elem, wp, dWin, dWin + winBytes + 64, 1, range);
    if (rc != HMME_OK) return rc;
    CU_TRY(c, cudaMemcpyAsync(c->hCtuRes, c->dCtuRes, 4 * HMME_NPARTS * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    const int32_t* hr = c->hCtuRes;
    std::memcpy(X, hr, HMME_NPARTS * 4); std::memcpy(Y, hr + HMME_NPARTS, HMME_NPARTS * 4);
    std::memcpy(sad, hr + 2 * HMME_NPARTS, HMME_NPARTS * 4);
    if (cost) std::memcpy(cost, hr + 3 * HMME_NPARTS, HMME_NPARTS * 4);
    return HMME_OK;
}

int hmme_plane_alloc(hmme_ctx* c, hmme_plane* out, int elemBytes, int width, int height, int marginX, int marginY) {
    if (!c || !out || (elemBytes != 1 && elemBytes != 2) || width <= 0 || height <= 0 || marginX < 0 || marginY < 0)
        return fail(c, HMME_ERR_ARG, "hmme_plane_alloc: bad argument");
    CU_TRY(c, cudaSetDevice(c->device));
    hmme_plane p{};
    p.elemBytes = elemBytes; p.width = width; p.height = height; p.marginX = marginX; p.marginY = marginY;
    p.pitch = (width + 2 * marginX + 15) & ~15;       // 16-element pitch keeps every row 16-byte aligned (TMA-ready)
    CU_TRY(c, cudaMalloc(&p.base, plane_elems(&p) * elemBytes + 64));
    // pitch padding and slack read as defined samples.  On the context's own stream and waited for: a plain cudaMemset runs on the
    // legacy stream, which the (non-blocking) upload streams do not order against -- it could land after the first upload
    CU_TRY(c, cudaMemsetAsync(p.base, 0, plane_elems(&p) * elemBytes + 64, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    *out = p;
    return HMME_OK;
}

int hmme_plane_free(hmme_ctx* c, hmme_plane* p) {
    if (!c || !p) return HMME_ERR_ARG;
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    if (p->base) CU_TRY(c, cudaFree(p->base));
    p->base = nullptr;
    return HMME_OK;
}

// Host rectangle -> device plane.  [x0, x1) x [y0, y1) are picture coordinates (negative / beyond the picture = margin samples,
// which must exist on the host side as they do in TComPicYuv).  Equal sample sizes: one strided DMA straight into the plane.
// int16 host -> 8-bit plane: DMA into a dense staging buffer, then me_narrow_rect_kernel (content check deferred to the next
// synchronising call).  Everything runs on one of the two high-priority io streams, ordered
//   after  every compute call already enqueued on this context (it may still read the plane: evCompute),
//   after  the previous upload on the other io stream (two uploads to the same plane keep their order),
//   before every compute call enqueued later (the compute stream waits on evUpload).
static int upload_rect(hmme_ctx* c, const hmme_plane* p, const void* hostOrigin, int hostStride, int hostElem, int x0, int y0, int x1, int y1,
                       const char* who) {
    if (!c || !hostOrigin) return fail(c, HMME_ERR_ARG, std::string(who) + ": null pointer");
    int rc = check_plane(c, p, "upload target");
    if (rc != HMME_OK) return rc;
    if (hostElem != 1 && hostElem != 2) return fail(c, HMME_ERR_ARG, std::string(who) + ": host samples must be 1 or 2 bytes");
    if (hostElem == 1 && p->elemBytes == 2) return fail(c, HMME_ERR_ARG, std::string(who) + ": an int16 plane takes int16 host samples");
    if (x0 >= x1 || y0 >= y1 || x0 < -p->marginX || y0 < -p->marginY || x1 > p->width + p->marginX || y1 > p->height + p->marginY)
        return fail(c, HMME_ERR_ARG, std::string(who) + ": rectangle empty or outside the padded plane");
    if (hostStride < x1 - x0) return fail(c, HMME_ERR_ARG, std::string(who) + ": host stride smaller than the rectangle");
    CU_TRY(c, cudaSetDevice(c->device));
    const int rows = y1 - y0, cols = x1 - x0;
    const char* src = static_cast<const char*>(hostOrigin) + ((ptrdiff_t)y0 * hostStride + x0) * hostElem;
    char* dst = static_cast<char*>(p->base) + ((size_t)(p->marginY + y0) * p->pitch + (size_t)(p->marginX + x0)) * p->elemBytes;
    const int sb = c->stageNext;
    c->stageNext ^= 1;
    cudaStream_t io = c->ioStream[sb];
    if (!c->capturing) CU_TRY(c, cudaStreamWaitEvent(io, c->evCompute, 0));   // inside a graph the whole previous launch has completed (stream order)
    if (c->uploadValid[sb ^ 1]) CU_TRY(c, cudaStreamWaitEvent(io, c->evUpload[sb ^ 1], 0));
    if (hostElem == p->elemBytes) {
        CU_TRY(c, cudaMemcpy2DAsync(dst, (size_t)p->pitch * hostElem, src, (size_t)hostStride * hostElem, (size_t)cols * hostElem, rows,
                                    cudaMemcpyHostToDevice, io));
    } else {
        const int spitch = (cols + 7) & ~7;
        const size_t n = (size_t)spitch * rows;
        if (c->stageElems[sb] < n) {
            if (c->capturing) return fail(c, HMME_ERR_ARG, std::string(who) + ": staging buffer would grow while capturing (run the step once first)");
            CU_TRY(c, cudaStreamSynchronize(io));
            if (c->dStage[sb]) cudaFree(c->dStage[sb]);
            c->dStage[sb] = nullptr; c->stageElems[sb] = 0;
            CU_TRY(c, cudaMalloc(&c->dStage[sb], n * 2));
            c->stageElems[sb] = n;
            ++c->bufGen;
        }
        CU_TRY(c, cudaMemcpy2DAsync(c->dStage[sb], (size_t)spitch * 2, src, (size_t)hostStride * 2, (size_t)cols * 2, rows, cudaMemcpyHostToDevice, io));
        const long long threads = (long long)((cols + 7) >> 3) * rows;
        me_narrow_rect_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, io>>>(c->dStage[sb], spitch, reinterpret_cast<uint8_t*>(dst), p->pitch, cols,
                                                                               rows, c->dFlag);
        c->launches += 1;
        CU_TRY(c, cudaGetLastError());
        c->contentCheckPending = true;                      // sticky device flag, read back at the next synchronising call
    }
    CU_TRY(c, cudaEventRecord(c->evUpload[sb], io));
    c->uploadValid[sb] = true;
    CU_TRY(c, cudaStreamWaitEvent(c->stream, c->evUpload[sb], 0));
    return HMME_OK;
}

int hmme_plane_upload_rect_async(hmme_ctx* c, const hmme_plane* p, const void* hostOrigin, int hostStride, int hostElemBytes, int x0, int y0,
                                 int x1, int y1) {
    return upload_rect(c, p, hostOrigin, hostStride, hostElemBytes, x0, y0, x1, y1, "hmme_plane_upload_rect");
}

int hmme_plane_upload_s16_async(hmme_ctx* c, const hmme_plane* p, const int16_t* hostOrigin, int hostStride) {
    if (!p) return fail(c, HMME_ERR_ARG, "hmme_plane_upload_s16: null plane");
    return upload_rect(c, p, hostOrigin, hostStride, 2, -p->marginX, -p->marginY, p->width + p->marginX, p->height + p->marginY, "hmme_plane_upload_s16");
}

int hmme_plane_upload_s16(hmme_ctx* c, const hmme_plane* p, const int16_t* hostOrigin, int hostStride) {
    const int rc = hmme_plane_upload_s16_async(c, p, hostOrigin, hostStride);
    return rc != HMME_OK ? rc : sync_ctx(c);
}

int hmme_plane_upload_u8_async(hmme_ctx* c, const hmme_plane* p, const uint8_t* hostOrigin, int hostStride) {
    if (!p) return fail(c, HMME_ERR_ARG, "hmme_plane_upload_u8: null plane");
    if (p->elemBytes != 1) return fail(c, HMME_ERR_ARG, "hmme_plane_upload_u8 needs an 8-bit plane");
    return upload_rect(c, p, hostOrigin, hostStride, 1, -p->marginX, -p->marginY, p->width + p->marginX, p->height + p->marginY, "hmme_plane_upload_u8");
}

int hmme_plane_upload_u8(hmme_ctx* c, const hmme_plane* p, const uint8_t* hostOrigin, int hostStride) {
    const int rc = hmme_plane_upload_u8_async(c, p, hostOrigin, hostStride);
    return rc != HMME_OK ? rc : sync_ctx(c);
}

int hmme_search_launch_size(hmme_ctx* c, int njobs, int range, int* ctas, int* ctasPerWave) {
    if (!c || njobs <= 0 || range < 0 || range > 1024) return fail(c, HMME_ERR_ARG, "hmme_search_launch_size: bad argument");
    const FastGeom g = fast_geometry(2 * range + 1, HMME_FAST_YB, njobs, c->prop.multiProcessorCount, c->forceRG);
    if (ctas) *ctas = njobs * g.nTx * g.nTy;
    if (ctasPerWave) *ctasPerWave = c->prop.multiProcessorCount;          // one 512-thread block per SM
    return HMME_OK;
}

int hmme_search_frame_async(hmme_ctx* c, const hmme_plane* cur, const hmme_plane* ref, const hmme_job* jobs, int njobs, int range) {
    if (!c) return HMME_ERR_ARG;
    if (!jobs || njobs <= 0) return fail(c, HMME_ERR_ARG, "hmme_search_frame: no jobs");
    if (range < 0 || range > 1024) return fail(c, HMME_ERR_RANGE, "search range out of [0,1024]");
    int rc = check_plane(c, cur, "current");
    if (rc == HMME_OK) rc = check_plane(c, ref, "reference");
    if (rc == HMME_OK) rc = check_jobs(c, cur, ref, jobs, njobs, range);
    if (rc != HMME_OK) return rc;
    CU_TRY(c, cudaSetDevice(c->device));
    rc = ensure_jobs(c, (size_t)njobs);
    if (rc != HMME_OK) return rc;
    // straight from the caller's (pageable) array: the runtime stages it before returning, and stream order protects dJobs,
    // so consecutive frames can be enqueued without a host synchronisation in between
    if (c->capturing) {                                     // a graph replays the copy: it reads a page-locked copy that the graph owns
        void* hj = nullptr;
        CU_TRY(c, cudaMallocHost(&hj, (size_t)njobs * sizeof(hmme_job)));
        c->captureBufs.push_back(hj);
        std::memcpy(hj, jobs, (size_t)njobs * sizeof(hmme_job));
        CU_TRY(c, cudaMemcpyAsync(c->dJobs, hj, (size_t)njobs * sizeof(hmme_job), cudaMemcpyHostToDevice, c->stream));
    } else
        CU_TRY(c, cudaMemcpyAsync(c->dJobs, jobs, (size_t)njobs * sizeof(hmme_job), cudaMemcpyHostToDevice, c->stream));
    c->lastSearchJobs = njobs;
    c->lastBox[0] = c->lastBox[1] = INT32_MAX; c->lastBox[2] = c->lastBox[3] = INT32_MIN;
    for (int j = 0; j < njobs; ++j) {
        c->lastBox[0] = std::min(c->lastBox[0], jobs[j].ctuX + jobs[j].ltx); c->lastBox[1] = std::min(c->lastBox[1], jobs[j].ctuY + jobs[j].lty);
        c->lastBox[2] = std::max(c->lastBox[2], jobs[j].ctuX + jobs[j].ltx + 2 * range + 64); c->lastBox[3] = std::max(c->lastBox[3], jobs[j].ctuY + jobs[j].lty + 2 * range + 64);
    }
    const char* refLo = static_cast<const char*>(ref->base);
    // planes carry 64 bytes of slack after the last row (hmme_plane_alloc adds it; required of external memory): the
    // 16-byte granular TMA row copies may run a few bytes past the window's last sample
    return enqueue_search(c, default_io(c), origin_ptr(cur), cur->elemBytes, cur->pitch, origin_ptr(ref), ref->elemBytes, ref->pitch, refLo,
                          refLo + plane_elems(ref) * ref->elemBytes + 64, njobs, range);
}

// ---- device-resident result tables: the reference keeps allMotionVectors[2][33][593] / allRuiCost for ONE
// CTU on the host (TEncSearch.h:114-115); a table holds the same 593-entry records for every CTU of a picture and every (list, reference
// index) slot in HBM, so that a B picture's two lists (and several window hypotheses per list) coexist and sub-CU lookups or the
// fractional refinement can read them without a host round trip.
struct hmme_table {
    hmme_ctx* owner = nullptr;
    int slots = 0, jobsPerSlot = 0;
    int32_t* dRes = nullptr;       // [slots][4][jobsPerSlot][593]: X, Y, sad, cost
    int4* dJobs = nullptr;         // [slots][jobsPerSlot]: the jobs each slot was searched with
    std::vector<int> njobs, range; // per slot, 0 = never searched
};

int hmme_table_create(hmme_ctx* c, hmme_table** out, int slots, int jobsPerSlot) {
    if (!c || !out || slots <= 0 || jobsPerSlot <= 0) return fail(c, HMME_ERR_ARG, "hmme_table_create: bad argument");
    *out = nullptr;
    CU_TRY(c, cudaSetDevice(c->device));
    int rc = ensure_jobs(c, (size_t)jobsPerSlot);          // the arg-min scratch must cover a whole slot
    if (rc != HMME_OK) return rc;
    hmme_table* t = new hmme_table;
    t->owner = c; t->slots = slots; t->jobsPerSlot = jobsPerSlot; t->njobs.assign(slots, 0); t->range.assign(slots, 0);
    const size_t n = (size_t)slots * jobsPerSlot;
    if (cudaMalloc(&t->dRes, n * 4 * HMME_NPARTS * sizeof(int32_t)) != cudaSuccess || cudaMalloc(&t->dJobs, n * sizeof(int4)) != cudaSuccess) {
        cudaFree(t->dRes); delete t; cudaGetLastError();
        return fail(c, HMME_ERR_CUDA, "hmme_table_create: out of device memory");
    }
    *out = t;
    return HMME_OK;
}

void hmme_table_destroy(hmme_table* t) {
    if (!t) return;
    if (t->owner) { cudaSetDevice(t->owner->device); cudaStreamSynchronize(t->owner->stream); cudaStreamSynchronize(t->owner->ioStream[0]); }
    cudaFree(t->dRes); cudaFree(t->dJobs);
    delete t;
}

int hmme_search_frame_table_async(hmme_ctx* c, const hmme_plane* cur, const hmme_plane* ref, const hmme_job* jobs, int njobs, int range,
                                  hmme_table* t, int slot) {
    if (!c) return HMME_ERR_ARG;
    if (!t || t->owner != c || slot < 0 || slot >= t->slots) return fail(c, HMME_ERR_ARG, "hmme_search_frame_table: table does not belong to this context / bad slot");
    if (!jobs || njobs <= 0 || njobs > t->jobsPerSlot) return fail(c, HMME_ERR_ARG, "hmme_search_frame_table: job count exceeds the table's jobs per slot");
    if (range < 0 || range > 1024) return fail(c, HMME_ERR_RANGE, "search range out of [0,1024]");
    if (c->capturing) return fail(c, HMME_ERR_ARG, "hmme_search_frame_table: not available while capturing a graph");
    int rc = check_plane(c, cur, "current");
    if (rc == HMME_OK) rc = check_plane(c, ref, "reference");
    if (rc == HMME_OK) rc = check_jobs(c, cur, ref, jobs, njobs, range);
    if (rc != HMME_OK) return rc;
    CU_TRY(c, cudaSetDevice(c->device));
    int4* dj = t->dJobs + (size_t)slot * t->jobsPerSlot;
    CU_TRY(c, cudaMemcpyAsync(dj, jobs, (size_t)njobs * sizeof(hmme_job), cudaMemcpyHostToDevice, c->stream));
    SearchIO io{};
    const size_t plane = (size_t)t->jobsPerSlot * HMME_NPARTS;
    io.jobs = dj;
    io.X = t->dRes + (size_t)slot * 4 * plane; io.Y = io.X + plane;
    io.S = reinterpret_cast<uint32_t*>(io.Y + plane); io.Cst = io.S + plane;
    io.finalizeInline = false;
    t->njobs[slot] = njobs; t->range[slot] = range;
    const char* refLo = static_cast<const char*>(ref->base);
    return enqueue_search(c, io, origin_ptr(cur), cur->elemBytes, cur->pitch, origin_ptr(ref), ref->elemBytes, ref->pitch, refLo,
                          refLo + plane_elems(ref) * ref->elemBytes + 64, njobs, range);
}

int hmme_table_fetch_async(hmme_ctx* c, hmme_table* t, int slot, int firstJob, int njobs, int32_t* X, int32_t* Y, uint32_t* sad, uint32_t* cost) {
    if (!c) return HMME_ERR_ARG;
    if (!t || t->owner != c || slot < 0 || slot >= t->slots || firstJob < 0 || njobs <= 0 || firstJob + njobs > t->njobs[slot])
        return fail(c, HMME_ERR_ARG, "hmme_table_fetch: bad slot or job range (the slot holds " + std::to_string(t && slot >= 0 && slot < t->slots ? t->njobs[slot] : 0) + " jobs)");
    CU_TRY(c, cudaSetDevice(c->device));
    const size_t plane = (size_t)t->jobsPerSlot * HMME_NPARTS, n = (size_t)njobs * HMME_NPARTS;
    const int32_t* base = t->dRes + (size_t)slot * 4 * plane + (size_t)firstJob * HMME_NPARTS;
    void* outs[4] = {X, Y, sad, cost};
    for (int k = 0; k < 4; ++k)
        if (outs[k]) CU_TRY(c, cudaMemcpyAsync(outs[k], base + k * plane, n * 4, cudaMemcpyDeviceToHost, c->stream));
    return HMME_OK;
}

const void* hmme_table_device_ptr(hmme_table* t, int slot, int array) {
    if (!t || slot < 0 || slot >= t->slots || array < 0 || array > 3) return nullptr;
    const size_t plane = (size_t)t->jobsPerSlot * HMME_NPARTS;
    return t->dRes + (size_t)slot * 4 * plane + (size_t)array * plane;
}

int hmme_fetch_results(hmme_ctx* c, int njobs, int32_t* X, int32_t* Y, uint32_t* sad, uint32_t* cost) {
    if (!c || njobs <= 0 || (size_t)njobs > c->jobCap) return fail(c, HMME_ERR_ARG, "hmme_fetch_results: bad job count");
    CU_TRY(c, cudaSetDevice(c->device));
    return fetch(c, njobs, X, Y, sad, cost);
}

int hmme_search_frame(hmme_ctx* c, const hmme_plane* cur, const hmme_plane* ref, const hmme_job* jobs, int njobs, int range,
                      int32_t* X, int32_t* Y, uint32_t* sad, uint32_t* cost) {
    if (!c) return HMME_ERR_ARG;
    if (!X || !Y || !sad) return fail(c, HMME_ERR_ARG, "hmme_search_frame: null output");
    int rc = hmme_search_frame_async(c, cur, ref, jobs, njobs, range);
    if (rc != HMME_OK) return rc;
    return fetch(c, njobs, X, Y, sad, cost);
}

int hmme_fetch_results_async(hmme_ctx* c, int njobs, int32_t* X, int32_t* Y, uint32_t* sad, uint32_t* cost) {
    if (!c || njobs <= 0 || (size_t)njobs > c->jobCap) return fail(c, HMME_ERR_ARG, "hmme_fetch_results: bad job count");
    CU_TRY(c, cudaSetDevice(c->device));
    return fetch_async(c, njobs, X, Y, sad, cost);
}

// ---- fractional-pel refinement (TEncSearch::xPatternSearchFracDIF, TEncSearch.cpp:4294-4331) ------------------------------
int hmme_refine_frac(hmme_ctx* c, const hmme_plane* cur, const hmme_plane* ref, const hmme_pu* pus, int npus, int useHad,
                     hmme_frac_result* results, uint32_t* candCosts) {
    if (!c) return HMME_ERR_ARG;
    if (!pus || npus <= 0 || !results) return fail(c, HMME_ERR_ARG, "hmme_refine_frac: no PUs / null output");
    int rc = check_frac_planes(c, cur, ref);
    if (rc == HMME_OK) rc = check_pus(c, cur, ref, pus, npus);
    if (rc != HMME_OK) return rc;
    CU_TRY(c, cudaSetDevice(c->device));
    rc = ensure_pus(c, (size_t)npus);
    if (rc != HMME_OK) return rc;
    static_assert(sizeof(hmme_pu) == sizeof(FracPu) && sizeof(hmme_frac_result) == sizeof(int4), "ABI structs mirror the kernel's");
    // PUs by 8x8-tile count, large to small (longest-first scheduling; the large ones get a CTA each); results go back to list order
    std::vector<int> idx;
    int nBig = 0;
    long long totalTiles = 0, segCount[5];
    segment_order(npus, [&](int n, int& w, int& h) { w = pus[n].w; h = pus[n].h; }, idx, segCount, nBig, totalTiles);
    std::vector<hmme_pu> sorted(npus);
    for (int i = 0; i < npus; ++i) sorted[i] = pus[idx[i]];
    CU_TRY(c, cudaMemcpyAsync(c->dPus, sorted.data(), (size_t)npus * sizeof(hmme_pu), cudaMemcpyHostToDevice, c->stream));
    CU_TRY(c, cudaMemcpyAsync(c->dSlots, idx.data(), (size_t)npus * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));           // the two vectors are pageable and go out of scope
    rc = enqueue_frac(c, cur, ref, npus, nBig, totalTiles, true, useHad, candCosts != nullptr, segCount);
    if (rc != HMME_OK) return rc;
    CU_TRY(c, cudaMemcpyAsync(results, c->dFrac, (size_t)npus * sizeof(int4), cudaMemcpyDeviceToHost, c->stream));
    if (candCosts) CU_TRY(c, cudaMemcpyAsync(candCosts, c->dCand, (size_t)npus * 18 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    return sync_ctx(c);
}

// One PU with HOST pointers, synchronous: the body of TEncSearch::xPatternSearchFracDIF as the encoder calls it
// (pattern key block, piRefY at the PU origin, integer MV).  Stages the block and the MV-displaced reference patch
// (4-sample apron) in the context's per-call buffers, like hmme_search_ctu stages its window.
int hmme_refine_pu(hmme_ctx* c, const int16_t* cur, int curStride, const int16_t* refAtPu, int refStride, int w, int h, int mvx, int mvy,
                   int predx, int predy, int useHad, int32_t* mvQpelX, int32_t* mvQpelY, uint32_t* cost, uint32_t* dist) {
    if (!c) return HMME_ERR_ARG;
    if (!cur || !refAtPu || !mvQpelX || !mvQpelY || !cost) return fail(c, HMME_ERR_ARG, "hmme_refine_pu: null pointer");
    if (w <= 0 || h <= 0 || w > 64 || h > 64 || (w & 3) || (h & 3)) return fail(c, HMME_ERR_ARG, "hmme_refine_pu: width/height must be multiples of 4 in [4,64]");
    CU_TRY(c, cudaSetDevice(c->device));
    int rc = ensure_pus(c, 1);
    if (rc != HMME_OK) return rc;
    constexpr int kPatchPitch = 96;                           // >= 64 + 8 (apron) + 8 (tile padding), multiple of 16
    const int pw = w + 8, ph = h + 8, w8 = (w + 7) & ~7, h8 = (h + 7) & ~7;
    uint8_t* hp = static_cast<uint8_t*>(c->hWin);
    int16_t* hc = static_cast<int16_t*>(c->hCurBlk);
    std::memset(hp, 0, (size_t)kPatchPitch * (h8 + 8));
    const int16_t* r0 = refAtPu + (long long)(mvy - 4) * refStride + (mvx - 4);
    for (int y = 0; y < ph; ++y)
        for (int x = 0; x < pw; ++x) {
            const int v = r0[(long long)y * refStride + x];
            if (v < 0 || v > 255) return fail(c, HMME_ERR_CONTENT, "hmme_refine_pu: reference samples outside [0,255] (8-bit video only)");
            hp[y * kPatchPitch + x] = (uint8_t)v;
        }
    for (int y = 0; y < h; ++y) std::memcpy(hc + y * 64, cur + (long long)y * curStride, (size_t)w * sizeof(int16_t));
    const hmme_pu pu{0, 0, w, h, 0, 0, predx - 4 * mvx, predy - 4 * mvy};      // the MV is folded into the patch origin; costs only see differences
    CU_TRY(c, cudaMemcpyAsync(c->dWin, hp, (size_t)kPatchPitch * (h8 + 8), cudaMemcpyHostToDevice, c->stream));
    CU_TRY(c, cudaMemcpyAsync(c->dCurBlk, hc, 64 * 64 * sizeof(int16_t), cudaMemcpyHostToDevice, c->stream));
    CU_TRY(c, cudaMemcpyAsync(c->dPus, &pu, sizeof(pu), cudaMemcpyHostToDevice, c->stream));
    FracParams fp{};
    fp.cur = c->dCurBlk; fp.ref = static_cast<const uint8_t*>(c->dWin) + 4 * kPatchPitch + 4;
    fp.curPitch = 64; fp.refPitch = kPatchPitch; fp.curBytes = 2;
    fp.pus = c->dPus; fp.slots = nullptr; fp.npus = 1; fp.nBig = (w8 / 8) * (h8 / 8) >= kFracCoopTiles ? 1 : 0;
    fp.lambda = c->lambda; fp.useHad = useHad ? 1 : 0;
    fp.out = c->dFrac; fp.cand = nullptr;
    if (fp.nBig) me_frac_coop_kernel<<<1, kFracThreads, 0, c->stream>>>(fp);
    else me_frac_kernel<<<1, kFracThreads, 0, c->stream>>>(fp);
    c->launches += 1;
    CU_TRY(c, cudaGetLastError());
    int4 res;
    CU_TRY(c, cudaMemcpyAsync(&res, c->dFrac, sizeof(res), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    *mvQpelX = res.x + 4 * mvx; *mvQpelY = res.y + 4 * mvy; *cost = (uint32_t)res.z;
    if (dist) *dist = (uint32_t)res.w;
    return HMME_OK;
}

// ---- distortion of motion-compensated prediction at quarter-pel MVs (xGetTemplateCost / merge candidates / xGetInterPredictionError)
namespace {

// pus: npus records of 6 (uni) or 8 (bi) ints; ref1 is the second reference plane of bi-directional PUs
int mc_cost_planes(hmme_ctx* c, const hmme_plane* cur, const hmme_plane* ref, const hmme_plane* ref1, const int32_t* pus, int npus, int useHad,
                   uint32_t* dist, const char* who) {
    const bool bi = ref1 != nullptr;
    const int stride = bi ? 8 : 6;
    if (!pus || npus <= 0 || !dist) return fail(c, HMME_ERR_ARG, std::string(who) + ": no PUs / null output");
    int rc = check_frac_planes(c, cur, ref);
    if (rc == HMME_OK && bi) rc = check_frac_planes(c, cur, ref1);
    if (rc != HMME_OK) return rc;
    for (int n = 0; n < npus; ++n)                          // same geometry rules as the refinement, around the integer part of each MV
        for (int l = 0; l < (bi ? 2 : 1); ++l) {
            const int32_t* u = pus + (size_t)n * stride;
            const hmme_pu g{u[0], u[1], u[2], u[3], u[4 + 2 * l] >> 2, u[5 + 2 * l] >> 2, 0, 0};
            rc = check_pus(c, cur, l ? ref1 : ref, &g, 1);
            if (rc != HMME_OK) { c->err = "PU " + std::to_string(n) + " of " + who + ": " + c->err; return rc; }
        }
    CU_TRY(c, cudaSetDevice(c->device));
    rc = ensure_pus(c, (size_t)npus);                       // dPus (32 B per entry) holds the 24- or 32-byte records, dFrac the results
    if (rc != HMME_OK) return rc;
    static_assert(sizeof(hmme_mc_pu) == 24 && sizeof(hmme_mc_bi_pu) == 32 && sizeof(FracPu) == 32, "record sizes");
    // The list goes to the device ordered by segment of the group kernel (counting sort; the PUs of four tiles or more by tile count, large
    // to small), results come back in list order through the slot table.
    const char* formStr = std::getenv("HMME_MC_FORM");     // experiments and tests: 1 = one PU per warp in list order, 2 = group form whatever the list length
    const int formEnv = formStr ? std::atoi(formStr) : 0;
    const bool group = formEnv == 2 || (formEnv != 1 && npus >= 64);
    McGroupParams mp{};
    if (group) {
        std::vector<int> idx;
        int nBig = 0;
        long long totalTiles = 0, segCount[5];
        segment_order(npus, [&](int n, int& w, int& h) { w = pus[(size_t)n * stride + 2]; h = pus[(size_t)n * stride + 3]; }, idx, segCount, nBig, totalTiles);
        std::vector<int32_t> sorted((size_t)npus * stride);
        for (int i = 0; i < npus; ++i) std::memcpy(&sorted[(size_t)i * stride], pus + (size_t)idx[i] * stride, stride * sizeof(int32_t));
        CU_TRY(c, cudaMemcpyAsync(c->dPus, sorted.data(), (size_t)npus * stride * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        CU_TRY(c, cudaMemcpyAsync(c->dSlots, idx.data(), (size_t)npus * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        CU_TRY(c, cudaStreamSynchronize(c->stream));       // the two vectors are pageable and go out of scope
        mp.slots = c->dSlots;
        mp.segPu[0] = 0; mp.segGrp[0] = 0;
        for (int q = 0; q < 5; ++q) {
            mp.segPu[q + 1] = mp.segPu[q] + (int)segCount[q];
            mp.segGrp[q + 1] = mp.segGrp[q] + (int)((segCount[q] + kFracSegPus[q] - 1) / kFracSegPus[q]);
        }
    } else
        CU_TRY(c, cudaMemcpyAsync(c->dPus, pus, (size_t)npus * stride * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    mp.cur = origin_ptr(cur); mp.ref = reinterpret_cast<const uint8_t*>(origin_ptr(ref));
    mp.ref1 = bi ? reinterpret_cast<const uint8_t*>(origin_ptr(ref1)) : nullptr;
    mp.curPitch = cur->pitch; mp.refPitch = ref->pitch; mp.ref1Pitch = bi ? ref1->pitch : 0; mp.curBytes = cur->elemBytes;
    mp.pus = reinterpret_cast<const int*>(c->dPus); mp.npus = npus; mp.useHad = useHad ? 1 : 0;
    mp.out = reinterpret_cast<uint32_t*>(c->dFrac);
    const int units = group ? mp.segGrp[5] : npus;         // one group / one PU per warp
    const int ctas = std::max(1, std::min((units + kFracWarps - 1) / kFracWarps, c->prop.multiProcessorCount * 512));
    CU_TRY(c, cudaEventRecord(c->evF0, c->stream));
    if (group) {
        if (bi) me_mc_group_kernel<true><<<ctas, kFracThreads, 0, c->stream>>>(mp);
        else me_mc_group_kernel<false><<<ctas, kFracThreads, 0, c->stream>>>(mp);
    } else {
        if (bi) me_mc_cost_kernel<true><<<ctas, kFracThreads, 0, c->stream>>>(mp);
        else me_mc_cost_kernel<false><<<ctas, kFracThreads, 0, c->stream>>>(mp);
    }
    CU_TRY(c, cudaEventRecord(c->evF1, c->stream));
    c->evFracValid = true;
    c->launches += 1;
    CU_TRY(c, cudaGetLastError());
    { const int rcm = mark_compute(c); if (rcm != HMME_OK) return rcm; }
    CU_TRY(c, cudaMemcpyAsync(dist, c->dFrac, (size_t)npus * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    return sync_ctx(c);
}

// One PU with HOST pointers: stages the block and one or two MV-displaced reference patches (4-sample apron) in the context's per-call
// buffers, like hmme_refine_pu.  The integer part of each MV is folded into where its patch is cut.
int mc_cost_host(hmme_ctx* c, const int16_t* cur, int curStride, const int16_t* const refAtPu[2], const int refStride[2], const int mv[2][2],
                 int nLists, int w, int h, int useHad, uint32_t* dist, const char* who) {
    if (!cur || !refAtPu[0] || (nLists == 2 && !refAtPu[1]) || !dist) return fail(c, HMME_ERR_ARG, std::string(who) + ": null pointer");
    if (w <= 0 || h <= 0 || w > 64 || h > 64 || (w & 3) || (h & 3)) return fail(c, HMME_ERR_ARG, std::string(who) + ": width/height must be multiples of 4 in [4,64]");
    CU_TRY(c, cudaSetDevice(c->device));
    int rc = ensure_pus(c, 1);
    if (rc != HMME_OK) return rc;
    constexpr int kPatchPitch = 96;
    const int pw = w + 8, ph = h + 8, h8 = (h + 7) & ~7;
    const size_t patchBytes = (size_t)kPatchPitch * (h8 + 8);   // <= 96 * 80; the staging window holds at least 80 * 80 * 2 bytes
    uint8_t* hp = static_cast<uint8_t*>(c->hWin);
    int16_t* hc = static_cast<int16_t*>(c->hCurBlk);
    if (2 * patchBytes > c->winElems * 2) return fail(c, HMME_ERR_RANGE, std::string(who) + ": staging window too small");
    std::memset(hp, 0, nLists * patchBytes);
    for (int l = 0; l < nLists; ++l) {
        const int16_t* r0 = refAtPu[l] + (long long)((mv[l][1] >> 2) - 4) * refStride[l] + ((mv[l][0] >> 2) - 4);
        for (int y = 0; y < ph; ++y)
            for (int x = 0; x < pw; ++x) {
                const int v = r0[(long long)y * refStride[l] + x];
                if (v < 0 || v > 255) return fail(c, HMME_ERR_CONTENT, std::string(who) + ": reference samples outside [0,255] (8-bit video only)");
                hp[l * patchBytes + y * kPatchPitch + x] = (uint8_t)v;
            }
    }
    for (int y = 0; y < h; ++y) std::memcpy(hc + y * 64, cur + (long long)y * curStride, (size_t)w * sizeof(int16_t));
    const int32_t pu[8] = {0, 0, w, h, mv[0][0] & 3, mv[0][1] & 3, nLists == 2 ? (mv[1][0] & 3) : 0, nLists == 2 ? (mv[1][1] & 3) : 0};
    CU_TRY(c, cudaMemcpyAsync(c->dWin, hp, nLists * patchBytes, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(c, cudaMemcpyAsync(c->dCurBlk, hc, 64 * 64 * sizeof(int16_t), cudaMemcpyHostToDevice, c->stream));
    CU_TRY(c, cudaMemcpyAsync(c->dPus, pu, sizeof(pu), cudaMemcpyHostToDevice, c->stream));
    McParams mp{};
    mp.cur = c->dCurBlk; mp.ref = static_cast<const uint8_t*>(c->dWin) + 4 * kPatchPitch + 4;
    mp.ref1 = static_cast<const uint8_t*>(c->dWin) + patchBytes + 4 * kPatchPitch + 4;
    mp.curPitch = 64; mp.refPitch = kPatchPitch; mp.ref1Pitch = kPatchPitch; mp.curBytes = 2;
    mp.pus = reinterpret_cast<const int*>(c->dPus); mp.npus = 1; mp.useHad = useHad ? 1 : 0;
    mp.out = reinterpret_cast<uint32_t*>(c->dFrac);
    if (nLists == 2) me_mc_cost_kernel<true><<<1, kFracThreads, 0, c->stream>>>(mp);
    else me_mc_cost_kernel<false><<<1, kFracThreads, 0, c->stream>>>(mp);
    c->launches += 1;
    CU_TRY(c, cudaGetLastError());
    CU_TRY(c, cudaMemcpyAsync(dist, c->dFrac, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return HMME_OK;
}

}  // namespace

int hmme_mc_cost(hmme_ctx* c, const hmme_plane* cur, const hmme_plane* ref, const hmme_mc_pu* pus, int npus, int useHad, uint32_t* dist) {
    if (!c) return HMME_ERR_ARG;
    return mc_cost_planes(c, cur, ref, nullptr, reinterpret_cast<const int32_t*>(pus), npus, useHad, dist, "hmme_mc_cost");
}

int hmme_mc_cost_bi(hmme_ctx* c, const hmme_plane* cur, const hmme_plane* ref0, const hmme_plane* ref1, const hmme_mc_bi_pu* pus, int npus,
                    int useHad, uint32_t* dist) {
    if (!c) return HMME_ERR_ARG;
    if (!ref1) return fail(c, HMME_ERR_ARG, "hmme_mc_cost_bi: null second reference plane");
    return mc_cost_planes(c, cur, ref0, ref1, reinterpret_cast<const int32_t*>(pus), npus, useHad, dist, "hmme_mc_cost_bi");
}

int hmme_mc_cost_pu(hmme_ctx* c, const int16_t* cur, int curStride, const int16_t* refAtPu, int refStride, int w, int h, int mvQpelX,
                    int mvQpelY, int useHad, uint32_t* dist) {
    if (!c) return HMME_ERR_ARG;
    const int16_t* const refs[2] = {refAtPu, nullptr};
    const int strides[2] = {refStride, 0};
    const int mv[2][2] = {{mvQpelX, mvQpelY}, {0, 0}};
    return mc_cost_host(c, cur, curStride, refs, strides, mv, 1, w, h, useHad, dist, "hmme_mc_cost_pu");
}

int hmme_mc_cost_bi_pu(hmme_ctx* c, const int16_t* cur, int curStride, const int16_t* ref0AtPu, int ref0Stride, int mv0QpelX, int mv0QpelY,
                       const int16_t* ref1AtPu, int ref1Stride, int mv1QpelX, int mv1QpelY, int w, int h, int useHad, uint32_t* dist) {
    if (!c) return HMME_ERR_ARG;
    const int16_t* const refs[2] = {ref0AtPu, ref1AtPu};
    const int strides[2] = {ref0Stride, ref1Stride};
    const int mv[2][2] = {{mv0QpelX, mv0QpelY}, {mv1QpelX, mv1QpelY}};
    return mc_cost_host(c, cur, curStride, refs, strides, mv, 2, w, h, useHad, dist, "hmme_mc_cost_bi_pu");
}

int hmme_refine_frame_async(hmme_ctx* c, const hmme_plane* cur, const hmme_plane* ref, int njobs, const int32_t* predsQpel, int useHad) {
    if (!c) return HMME_ERR_ARG;
    if (njobs <= 0 || njobs != c->lastSearchJobs)
        return fail(c, HMME_ERR_ARG, "hmme_refine_frame: job count must match the preceding hmme_search_frame on this context");
    int rc = check_frac_planes(c, cur, ref);
    if (rc != HMME_OK) return rc;
    // every MV the search can have returned keeps its PU inside lastBox; the filters add 4 samples and the kernel up to 4 of tile padding
    if (c->lastBox[0] - 4 < -ref->marginX || c->lastBox[1] - 4 < -ref->marginY || c->lastBox[2] + 8 > ref->width + ref->marginX ||
        c->lastBox[3] + 8 > ref->height + ref->marginY)
        return fail(c, HMME_ERR_BOUNDS, "hmme_refine_frame: the search windows leave less than 8 samples of margin in the reference plane for the interpolation apron");
    CU_TRY(c, cudaSetDevice(c->device));
    const size_t npus = (size_t)njobs * HMME_NPARTS;
    rc = ensure_pus(c, npus);
    if (rc != HMME_OK) return rc;
    if (predsQpel) {
        if ((size_t)njobs > c->predCap) {
            CU_TRY(c, cudaStreamSynchronize(c->stream));
            cudaFree(c->dPreds); c->dPreds = nullptr; c->predCap = 0;
            CU_TRY(c, cudaMalloc(&c->dPreds, c->jobCap * sizeof(int2)));
            c->predCap = c->jobCap;
        }
        CU_TRY(c, cudaMemcpyAsync(c->dPreds, predsQpel, (size_t)njobs * sizeof(int2), cudaMemcpyHostToDevice, c->stream));
    }
    // winners of the integer search, still on the device: X, Y of [jobCap][593]
    const int32_t* X = c->dRes;
    const int32_t* Y = X + c->jobCap * HMME_NPARTS;
    me_frac_build_kernel<<<(unsigned)((npus + 255) / 256), 256, 0, c->stream>>>(c->dJobs, X, Y, predsQpel ? c->dPreds : nullptr, c->dOrder, njobs,
                                                                                c->dPus, c->dSlots);
    c->launches += 1;
    long long segCount[5];
    for (int q = 0; q < 5; ++q) segCount[q] = (long long)c->segParts[q] * njobs;
    return enqueue_frac(c, cur, ref, (int)npus, njobs * c->bigParts, (long long)njobs * c->tilesPerCtu, true, useHad, false, segCount);
}

int hmme_fetch_frac_async(hmme_ctx* c, int njobs, hmme_frac_result* results) {
    if (!c || !results || njobs <= 0 || (size_t)njobs * HMME_NPARTS > c->puCap) return fail(c, HMME_ERR_ARG, "hmme_fetch_frac: bad job count / null output");
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, cudaMemcpyAsync(results, c->dFrac, (size_t)njobs * HMME_NPARTS * sizeof(int4), cudaMemcpyDeviceToHost, c->stream));
    return HMME_OK;
}

int hmme_refine_frame(hmme_ctx* c, const hmme_plane* cur, const hmme_plane* ref, int njobs, const int32_t* predsQpel, int useHad,
                      hmme_frac_result* results) {
    int rc = hmme_refine_frame_async(c, cur, ref, njobs, predsQpel, useHad);
    if (rc == HMME_OK) rc = hmme_fetch_frac_async(c, njobs, results);
    return rc != HMME_OK ? rc : sync_ctx(c);
}

int hmme_last_frac_ms(hmme_ctx* c, float* ms) {
    if (!c || !ms) return HMME_ERR_ARG;
    if (!c->evFracValid) return fail(c, HMME_ERR_ARG, "no refinement has been enqueued yet");
    CU_TRY(c, cudaEventSynchronize(c->evF1));
    CU_TRY(c, cudaEventElapsedTime(ms, c->evF0, c->evF1));
    return HMME_OK;
}

// ---- CUDA graphs: a launch-bound step (narrow bands on many GPUs: ~25 runtime calls for 0.16 ms of kernels) recorded once, replayed
// with one call.  Between begin and end the asynchronous calls of this context are captured instead of executed; their host
// buffers must be page-locked and stay where they are, device buffers must already have their final size (run the step once first).
int hmme_graph_begin(hmme_ctx* c) {
    if (!c) return HMME_ERR_ARG;
    if (c->capturing) return fail(c, HMME_ERR_ARG, "hmme_graph_begin: already capturing");
    CU_TRY(c, cudaSetDevice(c->device));
    if (!c->evFork) {
        CU_TRY(c, cudaEventCreateWithFlags(&c->evFork, cudaEventDisableTiming));
        for (int k = 0; k < 2; ++k) CU_TRY(c, cudaEventCreateWithFlags(&c->evJoin[k], cudaEventDisableTiming));
    }
    int rc = sync_ctx(c);
    if (rc != HMME_OK) return rc;
    c->uploadValid[0] = c->uploadValid[1] = false;          // events recorded before the capture are not part of it
    CU_TRY(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed));
    c->capturing = true;
    // the io streams join the capture by depending on the origin stream
    CU_TRY(c, cudaEventRecord(c->evFork, c->stream));
    for (int k = 0; k < 2; ++k) CU_TRY(c, cudaStreamWaitEvent(c->ioStream[k], c->evFork, 0));
    return HMME_OK;
}

int hmme_graph_end(hmme_ctx* c, hmme_graph** out) {
    if (!c || !out) return HMME_ERR_ARG;
    *out = nullptr;
    if (!c->capturing) return fail(c, HMME_ERR_ARG, "hmme_graph_end: not capturing");
    c->capturing = false;
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaSuccess;
    for (int k = 0; k < 2 && e == cudaSuccess; ++k) {       // every forked stream has to flow back into the origin stream
        e = cudaEventRecord(c->evJoin[k], c->ioStream[k]);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(c->stream, c->evJoin[k], 0);
    }
    const cudaError_t e2 = cudaStreamEndCapture(c->stream, &g);
    // events recorded while capturing belong to the graph: nothing outside may wait on or time them
    c->evValid = false; c->evFracValid = false; c->uploadValid[0] = c->uploadValid[1] = false;
    std::vector<void*> bufs;
    bufs.swap(c->captureBufs);
    if (e != cudaSuccess || e2 != cudaSuccess || !g) {
        if (g) cudaGraphDestroy(g);
        for (void* b : bufs) cudaFreeHost(b);
        cudaGetLastError();
        return fail(c, HMME_ERR_CUDA, std::string("graph capture failed: ") + cudaGetErrorString(e != cudaSuccess ? e : e2));
    }
    // Captured kernel nodes do not keep the priority of the io streams they were recorded on; give the small kernels that must
    // slip in between another context's search CTAs (narrowing, finalisation, PU list) their high priority back.
    {
        size_t nn = 0;
        cudaGraphGetNodes(g, nullptr, &nn);
        std::vector<cudaGraphNode_t> nodes(nn);
        if (nn) cudaGraphGetNodes(g, nodes.data(), &nn);
        int prLo = 0, prHi = 0;
        cudaDeviceGetStreamPriorityRange(&prLo, &prHi);
        for (cudaGraphNode_t nd : nodes) {
            cudaGraphNodeType ty;
            if (cudaGraphNodeGetType(nd, &ty) != cudaSuccess || ty != cudaGraphNodeTypeKernel) continue;
            cudaKernelNodeParams kp{};
            if (cudaGraphKernelNodeGetParams(nd, &kp) != cudaSuccess) continue;
            const bool small = kp.func == (void*)me_narrow_kernel || kp.func == (void*)me_finalize_kernel || kp.func == (void*)me_frac_build_kernel;
            cudaLaunchAttributeValue v{};
            v.priority = small ? prHi : prLo;
            cudaGraphKernelNodeSetAttribute(nd, cudaLaunchAttributePriority, &v);
        }
        cudaGetLastError();
    }
    hmme_graph* h = new hmme_graph;
    h->graph = g; h->owner = c; h->bufGen = c->bufGen; h->pinned.swap(bufs);
    e = cudaGraphInstantiateWithFlags(&h->exec, g, cudaGraphInstantiateFlagUseNodePriority);
    if (e != cudaSuccess) { hmme_graph_destroy(h); return fail(c, HMME_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e)); }
    CU_TRY(c, cudaEventRecord(c->evCompute, c->stream));    // outside the capture again: later uploads have a valid event to wait on
    *out = h;
    return HMME_OK;
}

int hmme_graph_launch(hmme_ctx* c, hmme_graph* g) {
    if (!c || !g || g->owner != c || !g->exec) return fail(c, HMME_ERR_ARG, "hmme_graph_launch: graph does not belong to this context");
    if (c->capturing) return fail(c, HMME_ERR_ARG, "hmme_graph_launch: capturing");
    if (g->bufGen != c->bufGen)
        return fail(c, HMME_ERR_ARG, "hmme_graph_launch: a device buffer of this context was reallocated after the graph was recorded; record it again");
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, cudaGraphLaunch(g->exec, c->stream));
    CU_TRY(c, cudaEventRecord(c->evCompute, c->stream));     // later uploads outside graphs order themselves after this launch
    c->contentCheckPending = true;
    c->evValid = false; c->evFracValid = false;             // the timing events were recorded inside the graph: not readable
    return HMME_OK;
}

void hmme_graph_destroy(hmme_graph* g) {
    if (!g) return;
    if (g->owner && g->owner->device >= 0) cudaSetDevice(g->owner->device);
    if (g->exec) cudaGraphExecDestroy(g->exec);
    if (g->graph) cudaGraphDestroy(g->graph);
    for (void* b : g->pinned) cudaFreeHost(b);
    delete g;
}

int hmme_sync(hmme_ctx* c) {
    if (!c) return HMME_ERR_ARG;
    CU_TRY(c, cudaSetDevice(c->device));
    return sync_ctx(c);
}

int hmme_last_kernel_ms(hmme_ctx* c, float* ms) {
    if (!c || !ms) return HMME_ERR_ARG;
    if (!c->evValid) return fail(c, HMME_ERR_ARG, "no search has been enqueued yet");
    CU_TRY(c, cudaEventSynchronize(c->ev1));
    CU_TRY(c, cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return HMME_OK;
}

uint64_t hmme_kernel_launches(hmme_ctx* c) { return c ? c->launches : 0; }

int hmme_measure_int_alu_peak(hmme_ctx* c, double* laneOpsPerSec, double* lanesPerClkPerSm, double* smMhz) {
    if (!c) return HMME_ERR_ARG;
    CU_TRY(c, cudaSetDevice(c->device));
    const int sms = c->prop.multiProcessorCount, threads = 1024, iters = 4000;
    uint32_t* dOut = nullptr; unsigned long long* dCyc = nullptr;
    CU_TRY(c, cudaMalloc(&dOut, (size_t)sms * threads * 4));
    CU_TRY(c, cudaMalloc(&dCyc, (size_t)sms * 8));
    me_alu_probe_kernel<<<sms, threads, 0, c->stream>>>(dOut, dCyc, 64, 1u);          // warm-up
    float best = 1e30f; double cycMed = 0;
    for (int rep = 0; rep < 3; ++rep) {
        CU_TRY(c, cudaEventRecord(c->ev0, c->stream));
        me_alu_probe_kernel<<<sms, threads, 0, c->stream>>>(dOut, dCyc, iters, 3u + rep);
        CU_TRY(c, cudaEventRecord(c->ev1, c->stream));
        CU_TRY(c, cudaEventSynchronize(c->ev1));
        float ms = 0; CU_TRY(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        if (ms < best) {
            best = ms;
            std::vector<unsigned long long> cyc(sms);
            CU_TRY(c, cudaMemcpy(cyc.data(), dCyc, (size_t)sms * 8, cudaMemcpyDeviceToHost));
            std::sort(cyc.begin(), cyc.end());
            cycMed = (double)cyc[sms / 2];
        }
    }
    c->launches += 4; c->evValid = false;
    cudaFree(dOut); cudaFree(dCyc);
    const double opsPerCta = (double)threads * 8 * 32 * iters;
    if (laneOpsPerSec) *laneOpsPerSec = opsPerCta * sms / (best * 1e-3);
    if (lanesPerClkPerSm) *lanesPerClkPerSm = opsPerCta / cycMed;
    if (smMhz) *smMhz = cycMed / (best * 1e-3) / 1e6;
    return HMME_OK;
}

int hmme_search_window(int predHorQpel, int predVerQpel, int range, int cuX, int cuY, int picWidth, int picHeight, int* ltx, int* lty,
                       int* rbx, int* rby) {
    if (!ltx || !lty || range < 0) return HMME_ERR_ARG;
    int rx = 0, ry = 0;
    search_window(predHorQpel, predVerQpel, range, cuX, cuY, picWidth, picHeight, HMME_CTU_SIZE, HMME_CTU_SIZE, ltx, lty, &rx, &ry);
    if (rbx) *rbx = rx;
    if (rby) *rby = ry;
    return HMME_OK;
}

int hmme_index_block(int partSize, int depth, int partIdx, int absZIdxInCtu, int cuWidth, int cuHeight) {
    return index_block(partSize, depth, partIdx, absZIdxInCtu, cuWidth, cuHeight);
}

int hmme_partition_rect(int index, int* x, int* y, int* w, int* h) {
    if (index < 0 || index >= HMME_NPARTS) return HMME_ERR_ARG;
    const PartRect r = part_rect(index);
    if (x) *x = r.x;
    if (y) *y = r.y;
    if (w) *w = r.w;
    if (h) *h = r.h;
    return HMME_OK;
}

}  // extern "C"
