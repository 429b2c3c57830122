// hmme_internal.cuh -- state shared by the translation units of libhmme_b200.so (not part of the C ABI).
#pragma once
#include <cstdint>
#include <mutex>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/hmme_b200.h"

namespace hmme { struct FracPu; }   // me_frac_kernel.cuh (kernel definitions live in hmme_b200.cu only)

struct hmme_ctx {
    int device = -1;
    cudaStream_t stream = nullptr;      // compute: search kernels and result copies
    cudaStream_t ioStream[2] = {nullptr, nullptr};   // high priority, one per staging buffer: a frame's two plane uploads copy back to back
                                                      // instead of the second copy queueing behind the first plane's narrowing kernel;
                                                      // ioStream[0] also runs the finalize kernel
    cudaEvent_t evUpload[2] = {nullptr, nullptr}, evCompute = nullptr, evFinal = nullptr;   // io -> compute and compute -> io ordering
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool evValid = false;
    bool uploadValid[2] = {false, false};   // evUpload[k] has been recorded (since the last capture began)
    cudaDeviceProp prop{};
    std::string err;
    uint32_t lambda = 0;
    int maxRange = 0;
    uint64_t launches = 0;
    int stagger = 1050;          // cycles of start skew between the two warp groups of the packed kernel (HMME_STAGGER env overrides)
    int forceRG = 0;             // HMME_FAST_RG env: force the number of row groups per tile (experiments)
    // job / result buffers (grown on demand)
    size_t jobCap = 0;
    int4* dJobs = nullptr;
    unsigned long long* dBest = nullptr;   // arg-min scratch, kept all "no winner" between searches
    int32_t* dRes = nullptr;      // [4][jobCap][593]: X, Y, sad, cost
    hmme_job* hJobs = nullptr;    // pinned
    // per-CTU synchronous path staging
    size_t winElems = 0;          // (2*maxRange+64+16)^2
    void* hWin = nullptr;         // pinned, int16-sized
    void* dWin = nullptr;
    int32_t* hCtuRes = nullptr;   // pinned [4][593]: X, Y, sad, cost of the per-CTU call, fetched with one copy
    int32_t* dCtuRes = nullptr;
    size_t fastSmemSet = 0;       // dynamic shared memory the packed kernel has been opted in for
    void* hCurBlk = nullptr;      // pinned 64x64 int16
    void* dCurBlk = nullptr;
    // bi-prediction blocks (16-bit current samples against an 8-bit picture): clamped 64x64 records + per-partition SAD constants
    uint8_t* dBiBlocks = nullptr; uint32_t* dBiOffsets = nullptr; size_t biCap = 0;
    // upload staging
    int16_t* dStage[2] = {nullptr, nullptr}; size_t stageElems[2] = {0, 0}; int stageNext = 0;   // two staging buffers: a frame's reference
                                                                                                  // and current plane copy back to back
    int* dFlag = nullptr; int* hFlag = nullptr;
    bool contentCheckPending = false;   // an _async 8-bit upload has not had its range flag read back yet
    // fractional-pel refinement (grown on demand)
    size_t puCap = 0;
    hmme::FracPu* dPus = nullptr; int* dSlots = nullptr; int4* dFrac = nullptr; uint32_t* dCand = nullptr;
    int* dOrder = nullptr;        // 593 partition indices, by 8x8-tile count, large to small
    int bigParts = 0;             // how many of them get a whole CTA in the small-batch form (kFracCoopTiles tiles or more)
    int tilesPerCtu = 0;          // 8x8 tiles of all 593 partitions (1792)
    int segParts[5] = {0, 0, 0, 0, 0};   // partitions per segment of the group kernel (hmme::frac_segment)
    int2* dPreds = nullptr; size_t predCap = 0;
    cudaEvent_t evF0 = nullptr, evF1 = nullptr; bool evFracValid = false;
    uint64_t bufGen = 0;          // bumped whenever a device buffer a graph may reference is reallocated
    std::vector<void*> captureBufs;   // page-locked buffers allocated while capturing; handed to the hmme_graph at hmme_graph_end
    bool capturing = false;       // between hmme_graph_begin and hmme_graph_end: the asynchronous calls are recorded, not run
    cudaEvent_t evFork = nullptr, evJoin[2] = {nullptr, nullptr};
    int lastSearchJobs = 0;       // job count of the most recent frame search (its winners feed hmme_refine_frame)
    int lastBox[4] = {0, 0, 0, 0};   // picture-coordinate bounding box [x0, y0, x1, y1) of every sample that search could point a PU at
};

struct hmme_graph {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    hmme_ctx* owner = nullptr;
    uint64_t bufGen = 0;              // owner->bufGen when the graph was recorded
    std::vector<void*> pinned;        // page-locked job lists the graph's copy nodes read
};


int hmme_fail(hmme_ctx* c, int code, const std::string& msg);

#define CU_TRY(c, expr)                                                                                      \
    do {                                                                                                     \
        cudaError_t e_ = (expr);                                                                             \
        if (e_ != cudaSuccess)                                                                               \
            return hmme_fail((c), HMME_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));       \
    } while (0)
