// hmme_group.cu -- one frame over several B200s behind the C ABI (include/hmme_b200.h "multi-GPU").
//
// The reference has no multi-device code at all (one queue on one device, /root/reference/source/Lib/TLibEncoder/TEncOpenCL.cpp:129,185);
// north_star adds it: the jobs of a frame are cut into contiguous bands (CTU rows, cut at CTU granularity so every GPU gets the
// same number of jobs +-1), every GPU searches its band against the reference picture and the results land in ONE host table.
// Two ways to get the reference picture onto the GPUs, both here and both measured by bench.py:
//   HMME_REF_BAND_HALO : every GPU fetches, over its own PCIe link, only the rows its band can read (band + search-window halo);
//                        no collective at all;
//   HMME_REF_BROADCAST : global rank 0 uploads the whole plane and ncclBroadcast moves it over NVLink / NVSwitch.
// Two process models, same code: hmme_group_create drives n GPUs from one process (one enqueue thread per GPU, ncclCommInitAll),
// hmme_group_create_rank is one rank of a one-process-per-GPU job (torchrun; ncclCommInitRank with an id the launcher distributes).
// NCCL is bound at run time (dlopen) and only when a broadcast is asked for: a single-GPU encoder never needs it.
#include <algorithm>
#include <climits>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <thread>

#include <dlfcn.h>
#include <nccl.h>

#include "hmme_internal.cuh"

namespace {

constexpr int kDeepPipelineWaves = 10;      // hmme_group_pipeline_depth: bands shorter than this many waves take the third slot
constexpr int kSlots = HMME_GROUP_SLOTS;   // frames in flight per GPU: the copies of one overlap the kernels of the others.  Three, because with two a
                                           // narrow band (8 GPUs: 0.14 ms of kernel) leaves the GPU with ONE queued search while the host waits for a
                                           // slot's results and refills it, so the last, partly filled wave of that search runs alone: measured on a
                                           // 60-job band 0.156 ms per frame with two slots, 0.138 with three (= the resident rate), 0.176 with four

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
};

NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        // the copy a host application already loaded (e.g. the one bundled with torch) wins; otherwise the system library
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) { api.err = std::string("dlopen(libnccl.so.2): ") + dlerror(); return; }
        auto sym = [&](const char* n) { void* p = dlsym(api.lib, n); if (!p && api.err.empty()) api.err = std::string("NCCL symbol missing: ") + n; return p; };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
        api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
        api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    });
    return &api;
}

// One enqueue thread per GPU of a single-process group: a frame is ~10 runtime calls per GPU, which one host thread would
// serialise (8 GPUs x ~40 us against a 0.16 ms step).
class Worker {
public:
    Worker() : th_([this] { loop(); }) {}
    ~Worker() {
        { std::lock_guard<std::mutex> l(mu_); quit_ = true; }
        cv_.notify_all();
        th_.join();
    }
    void post(std::function<void()> f) {
        { std::lock_guard<std::mutex> l(mu_); task_ = std::move(f); busy_ = true; }
        cv_.notify_all();
    }
    void wait() {
        std::unique_lock<std::mutex> l(mu_);
        done_.wait(l, [this] { return !busy_; });
    }
private:
    void loop() {
        for (;;) {
            std::function<void()> f;
            {
                std::unique_lock<std::mutex> l(mu_);
                cv_.wait(l, [this] { return quit_ || busy_; });
                if (quit_ && !busy_) return;
                f = std::move(task_);
            }
            f();
            { std::lock_guard<std::mutex> l(mu_); busy_ = false; }
            done_.notify_all();
        }
    }
    std::mutex mu_;
    std::condition_variable cv_, done_;
    std::function<void()> task_;
    bool busy_ = false, quit_ = false;
    std::thread th_;
};

struct GroupDev {
    int device = -1, globalRank = 0;
    hmme_ctx* ctx[kSlots] = {};
    hmme_plane cur[kSlots] = {}, ref[kSlots] = {};
    ncclComm_t comm = nullptr;
    cudaStream_t ncclStream = nullptr;             // high priority: the broadcast slips in between the other slot's search CTAs
    cudaEvent_t evPre = nullptr, evPost = nullptr;
    Worker* worker = nullptr;
    int rc = 0; std::string err;                   // outcome of the most recent per-device enqueue
    int first[kSlots] = {}, count[kSlots] = {};    // band of the frame in flight in each slot
};

}  // namespace

struct hmme_group {
    int world = 1, nlocal = 1;
    std::vector<GroupDev> devs;
    int maxRange = 0;
    int width = 0, height = 0, marginX = 0, marginY = 0, refDist = HMME_REF_BAND_HALO;
    bool commReady = false;
    ncclUniqueId uid{}; bool haveUid = false;
    std::string err;
};

namespace {

std::mutex g_groupErrMu;
std::string g_groupCreateErr;

int gfail(hmme_group* g, int code, const std::string& msg) {
    if (g) g->err = msg;
    else { std::lock_guard<std::mutex> l(g_groupErrMu); g_groupCreateErr = msg; }
    return code;
}

// run fn(dev index) for every local GPU: inline for one, on the per-GPU enqueue threads otherwise; first failure wins
int for_devices(hmme_group* g, const std::function<int(int)>& fn) {
    if (g->nlocal == 1) {
        g->devs[0].rc = fn(0);
    } else {
        for (int i = 0; i < g->nlocal; ++i) g->devs[i].worker->post([g, i, &fn] { g->devs[i].rc = fn(i); });
        for (int i = 0; i < g->nlocal; ++i) g->devs[i].worker->wait();
    }
    for (int i = 0; i < g->nlocal; ++i)
        if (g->devs[i].rc != HMME_OK) {
            const char* m = g->devs[i].err.empty() ? "" : g->devs[i].err.c_str();
            return gfail(g, g->devs[i].rc, "GPU " + std::to_string(g->devs[i].device) + " (rank " + std::to_string(g->devs[i].globalRank) + "): " + m);
        }
    return HMME_OK;
}

int dev_fail(GroupDev& d, hmme_ctx* c, int rc) { d.err = c ? hmme_last_error(c) : "no context"; return rc; }

int ensure_comm(hmme_group* g) {
    if (g->commReady) return HMME_OK;
    NcclApi* n = nccl_api();
    if (!n->lib || !n->err.empty()) return gfail(g, HMME_ERR_NCCL, "NCCL is not available: " + n->err);
    ncclResult_t r = ncclSuccess;
    if (g->nlocal == g->world) {                       // one process, all GPUs
        std::vector<int> ids(g->nlocal);
        std::vector<ncclComm_t> comms(g->nlocal);
        for (int i = 0; i < g->nlocal; ++i) ids[i] = g->devs[i].device;
        r = n->CommInitAll(comms.data(), g->nlocal, ids.data());
        if (r == ncclSuccess) for (int i = 0; i < g->nlocal; ++i) g->devs[i].comm = comms[i];
    } else {                                           // one rank of a one-process-per-GPU job
        if (!g->haveUid) return gfail(g, HMME_ERR_ARG, "hmme_group_create_rank was given no NCCL unique id: HMME_REF_BROADCAST is not available");
        if (cudaSetDevice(g->devs[0].device) != cudaSuccess) return gfail(g, HMME_ERR_CUDA, "cudaSetDevice failed");
        r = n->CommInitRank(&g->devs[0].comm, g->world, g->uid, g->devs[0].globalRank);
    }
    if (r != ncclSuccess) return gfail(g, HMME_ERR_NCCL, std::string("NCCL communicator: ") + n->GetErrorString(r));
    int prLo = 0, prHi = 0;
    for (GroupDev& d : g->devs) {
        if (cudaSetDevice(d.device) != cudaSuccess) return gfail(g, HMME_ERR_CUDA, "cudaSetDevice failed");
        cudaDeviceGetStreamPriorityRange(&prLo, &prHi);
        if (cudaStreamCreateWithPriority(&d.ncclStream, cudaStreamNonBlocking, prHi) != cudaSuccess ||
            cudaEventCreateWithFlags(&d.evPre, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&d.evPost, cudaEventDisableTiming) != cudaSuccess)
            return gfail(g, HMME_ERR_CUDA, "stream/event creation for the NCCL broadcast failed");
    }
    g->commReady = true;
    return HMME_OK;
}

int group_alloc(hmme_group** out, const int* devices, int nlocal, int world, int rank0, int maxRange) {
    if (!out) return HMME_ERR_ARG;
    *out = nullptr;
    if (!devices || nlocal <= 0 || world < nlocal || rank0 < 0 || rank0 + nlocal > world) return gfail(nullptr, HMME_ERR_ARG, "hmme_group_create: bad device list / rank");
    for (int i = 0; i < nlocal; ++i)
        for (int j = 0; j < i; ++j)
            if (devices[i] == devices[j]) return gfail(nullptr, HMME_ERR_ARG, "hmme_group_create: a device is listed twice");
    hmme_group* g = new hmme_group;
    g->world = world; g->nlocal = nlocal; g->maxRange = maxRange;
    g->devs.resize(nlocal);
    for (int i = 0; i < nlocal; ++i) {
        GroupDev& d = g->devs[i];
        d.device = devices[i]; d.globalRank = rank0 + i;
        for (int s = 0; s < kSlots; ++s) {
            const int rc = hmme_create(&d.ctx[s], devices[i], HMME_CTU_SIZE, HMME_CTU_SIZE, maxRange);
            if (rc != HMME_OK) {
                const std::string m = hmme_last_error(nullptr);
                hmme_group_destroy(g);
                return gfail(nullptr, rc, "hmme_group_create: GPU " + std::to_string(devices[i]) + ": " + m);
            }
        }
        if (nlocal > 1) d.worker = new Worker;
    }
    *out = g;
    return HMME_OK;
}

}  // namespace

extern "C" {

// ---- pure host logic (exported so that it can be tested without a GPU) -------------------------------------------------------
int hmme_band_split(int njobs, int world, int rank, int* first, int* count) {
    if (njobs < 0 || world <= 0 || rank < 0 || rank >= world || !first || !count) return HMME_ERR_ARG;
    const int base = njobs / world, extra = njobs % world;
    *first = rank * base + std::min(rank, extra);
    *count = base + (rank < extra ? 1 : 0);
    return HMME_OK;
}

int hmme_band_extent(const hmme_job* jobs, int njobs, int range, int* curRect, int* refRect) {
    if (!jobs || njobs <= 0 || range < 0 || !curRect || !refRect) return HMME_ERR_ARG;
    int c[4] = {INT_MAX, INT_MAX, INT_MIN, INT_MIN}, r[4] = {INT_MAX, INT_MAX, INT_MIN, INT_MIN};
    const int side = 2 * range + HMME_CTU_SIZE;
    for (int j = 0; j < njobs; ++j) {
        const hmme_job& b = jobs[j];
        c[0] = std::min(c[0], b.ctuX); c[1] = std::min(c[1], b.ctuY);
        c[2] = std::max(c[2], b.ctuX + HMME_CTU_SIZE); c[3] = std::max(c[3], b.ctuY + HMME_CTU_SIZE);
        r[0] = std::min(r[0], b.ctuX + b.ltx); r[1] = std::min(r[1], b.ctuY + b.lty);
        r[2] = std::max(r[2], b.ctuX + b.ltx + side); r[3] = std::max(r[3], b.ctuY + b.lty + side);
    }
    std::memcpy(curRect, c, sizeof(c)); std::memcpy(refRect, r, sizeof(r));
    return HMME_OK;
}

// ---- lifetime ------------------------------------------------------------------------------------------------------------------
int hmme_group_create(hmme_group** out, const int* devices, int ndev, int maxSearchRange) {
    return group_alloc(out, devices, ndev, ndev, 0, maxSearchRange);
}

int hmme_group_unique_id(void* id, size_t bytes) {
    if (!id || bytes < NCCL_UNIQUE_ID_BYTES) return gfail(nullptr, HMME_ERR_ARG, "hmme_group_unique_id: buffer smaller than 128 bytes");
    NcclApi* n = nccl_api();
    if (!n->lib || !n->err.empty()) return gfail(nullptr, HMME_ERR_NCCL, "NCCL is not available: " + n->err);
    ncclUniqueId u;
    const ncclResult_t r = n->GetUniqueId(&u);
    if (r != ncclSuccess) return gfail(nullptr, HMME_ERR_NCCL, std::string("ncclGetUniqueId: ") + n->GetErrorString(r));
    std::memcpy(id, &u, NCCL_UNIQUE_ID_BYTES);
    return HMME_OK;
}

int hmme_group_create_rank(hmme_group** out, int device, int rank, int nranks, const void* uniqueId, int maxSearchRange) {
    const int rc = group_alloc(out, &device, 1, nranks, rank, maxSearchRange);
    if (rc != HMME_OK) return rc;
    if (uniqueId) { std::memcpy(&(*out)->uid, uniqueId, NCCL_UNIQUE_ID_BYTES); (*out)->haveUid = true; }
    return HMME_OK;
}

void hmme_group_destroy(hmme_group* g) {
    if (!g) return;
    for (GroupDev& d : g->devs) {
        delete d.worker;
        if (d.device >= 0) cudaSetDevice(d.device);
        for (int s = 0; s < kSlots; ++s) {
            if (!d.ctx[s]) continue;
            hmme_sync(d.ctx[s]);
            if (d.cur[s].base) hmme_plane_free(d.ctx[s], &d.cur[s]);
            if (d.ref[s].base) hmme_plane_free(d.ctx[s], &d.ref[s]);
        }
        if (d.ncclStream) { cudaStreamSynchronize(d.ncclStream); cudaStreamDestroy(d.ncclStream); }
        if (d.evPre) cudaEventDestroy(d.evPre);
        if (d.evPost) cudaEventDestroy(d.evPost);
        if (d.comm && nccl_api()->CommDestroy) nccl_api()->CommDestroy(d.comm);
        for (int s = 0; s < kSlots; ++s) hmme_destroy(d.ctx[s]);
    }
    delete g;
}

const char* hmme_group_last_error(hmme_group* g) {
    if (g) return g->err.c_str();
    static thread_local std::string copy;
    std::lock_guard<std::mutex> l(g_groupErrMu);
    copy = g_groupCreateErr;
    return copy.c_str();
}

int hmme_group_size(hmme_group* g, int* world, int* nlocal) {
    if (!g) return HMME_ERR_ARG;
    if (world) *world = g->world;
    if (nlocal) *nlocal = g->nlocal;
    return HMME_OK;
}

int hmme_group_set_lambda_q16(hmme_group* g, uint32_t lambdaQ16) {
    if (!g) return HMME_ERR_ARG;
    for (GroupDev& d : g->devs)
        for (int s = 0; s < kSlots; ++s) hmme_set_lambda_q16(d.ctx[s], lambdaQ16);
    return HMME_OK;
}

int hmme_group_configure(hmme_group* g, int width, int height, int marginX, int marginY, int refDist) {
    if (!g) return HMME_ERR_ARG;
    if (width < HMME_CTU_SIZE || height < HMME_CTU_SIZE || marginX < 0 || marginY < 0) return gfail(g, HMME_ERR_ARG, "hmme_group_configure: bad picture geometry");
    if (refDist != HMME_REF_BAND_HALO && refDist != HMME_REF_BROADCAST) return gfail(g, HMME_ERR_ARG, "hmme_group_configure: unknown reference distribution");
    if (refDist == HMME_REF_BROADCAST && g->world > 1) {
        const int rc = ensure_comm(g);
        if (rc != HMME_OK) return rc;
    }
    const bool same = width == g->width && height == g->height && marginX == g->marginX && marginY == g->marginY;
    g->refDist = refDist;
    if (same) return HMME_OK;
    const int rc = for_devices(g, [&](int i) -> int {
        GroupDev& d = g->devs[i];
        for (int s = 0; s < kSlots; ++s) {
            hmme_ctx* c = d.ctx[s];
            int r = hmme_sync(c);
            if (r != HMME_OK) return dev_fail(d, c, r);
            if (d.cur[s].base) hmme_plane_free(c, &d.cur[s]);
            if (d.ref[s].base) hmme_plane_free(c, &d.ref[s]);
            // the current frame is only ever read inside its CTUs: no margin; every GPU holds full-size planes and fills the
            // rows of its band (180 GB of HBM per GPU: the two 8-bit planes of a 4K picture are 18 MB)
            if ((r = hmme_plane_alloc(c, &d.cur[s], 1, width, height, 0, 0)) != HMME_OK) return dev_fail(d, c, r);
            if ((r = hmme_plane_alloc(c, &d.ref[s], 1, width, height, marginX, marginY)) != HMME_OK) return dev_fail(d, c, r);
        }
        return HMME_OK;
    });
    if (rc != HMME_OK) return rc;
    g->width = width; g->height = height; g->marginX = marginX; g->marginY = marginY;
    return HMME_OK;
}

// ---- one frame ------------------------------------------------------------------------------------------------------------------
int hmme_group_search_frame_async(hmme_group* g, int slot, const void* curHostOrigin, int curHostStride, const void* refHostOrigin,
                                  int refHostStride, int hostElemBytes, const hmme_job* jobs, int njobs, int range, int32_t* X, int32_t* Y,
                                  uint32_t* sad, uint32_t* cost) {
    if (!g) return HMME_ERR_ARG;
    if (slot < 0 || slot >= kSlots) return gfail(g, HMME_ERR_ARG, "hmme_group_search_frame: slot must be in [0, HMME_GROUP_SLOTS)");
    if (!g->width) return gfail(g, HMME_ERR_ARG, "hmme_group_search_frame: call hmme_group_configure first");
    if (!curHostOrigin || !refHostOrigin || !jobs || njobs <= 0 || !X || !Y || !sad) return gfail(g, HMME_ERR_ARG, "hmme_group_search_frame: null pointer / no jobs");
    if (hostElemBytes != 1 && hostElemBytes != 2) return gfail(g, HMME_ERR_ARG, "hmme_group_search_frame: host samples must be uint8 or int16");
    if (range < 0 || range > g->maxRange) return gfail(g, HMME_ERR_RANGE, "hmme_group_search_frame: search range beyond what the group was created for");
    // every window inside the padded reference picture, every CTU inside the picture: bands are uploaded as rectangles, so the
    // reference's row-wrap addressing (App. B4) is not available here (the per-CTU call keeps it)
    for (int j = 0; j < njobs; ++j) {
        const hmme_job& b = jobs[j];
        if (b.ctuX < 0 || b.ctuY < 0 || b.ctuX + HMME_CTU_SIZE > g->width || b.ctuY + HMME_CTU_SIZE > g->height)
            return gfail(g, HMME_ERR_BOUNDS, "job " + std::to_string(j) + ": CTU outside the picture");
        if (b.ctuX + b.ltx < -g->marginX || b.ctuY + b.lty < -g->marginY || b.ctuX + b.ltx + 2 * range + HMME_CTU_SIZE > g->width + g->marginX ||
            b.ctuY + b.lty + 2 * range + HMME_CTU_SIZE > g->height + g->marginY)
            return gfail(g, HMME_ERR_BOUNDS, "job " + std::to_string(j) + ": search window leaves the padded reference picture");
    }
    const bool bcast = g->refDist == HMME_REF_BROADCAST && g->world > 1;

    // phase 1 (per GPU, in parallel): the band's rows of the current frame, and its share of the reference picture
    int rc = for_devices(g, [&](int i) -> int {
        GroupDev& d = g->devs[i];
        hmme_ctx* c = d.ctx[slot];
        hmme_band_split(njobs, g->world, d.globalRank, &d.first[slot], &d.count[slot]);
        int r = HMME_OK;
        if (bcast && d.globalRank == 0) {
            r = hmme_plane_upload_rect_async(c, &d.ref[slot], refHostOrigin, refHostStride, hostElemBytes, -g->marginX, -g->marginY,
                                             g->width + g->marginX, g->height + g->marginY);
            if (r != HMME_OK) return dev_fail(d, c, r);
        }
        if (!d.count[slot]) return HMME_OK;
        int cr[4], rr[4];
        hmme_band_extent(jobs + d.first[slot], d.count[slot], range, cr, rr);
        if (!bcast) {
            r = hmme_plane_upload_rect_async(c, &d.ref[slot], refHostOrigin, refHostStride, hostElemBytes, rr[0], rr[1], rr[2], rr[3]);
            if (r != HMME_OK) return dev_fail(d, c, r);
        }
        r = hmme_plane_upload_rect_async(c, &d.cur[slot], curHostOrigin, curHostStride, hostElemBytes, cr[0], cr[1], cr[2], cr[3]);
        return r != HMME_OK ? dev_fail(d, c, r) : HMME_OK;
    });
    if (rc != HMME_OK) return rc;

    // phase 2 (calling thread): the reference picture travels GPU to GPU; the collective sits on a high-priority stream between
    // "rank 0's upload has landed / nobody reads the old picture any more" and "the searches may start"
    if (bcast) {
        NcclApi* n = nccl_api();
        for (GroupDev& d : g->devs) {
            hmme_ctx* c = d.ctx[slot];
            if (cudaSetDevice(d.device) != cudaSuccess || cudaEventRecord(d.evPre, c->stream) != cudaSuccess ||
                cudaStreamWaitEvent(d.ncclStream, d.evPre, 0) != cudaSuccess)
                return gfail(g, HMME_ERR_CUDA, "event ordering before the NCCL broadcast failed");
        }
        ncclResult_t r = g->nlocal > 1 ? n->GroupStart() : ncclSuccess;
        for (GroupDev& d : g->devs) {
            if (r != ncclSuccess) break;
            const hmme_plane& p = d.ref[slot];
            const size_t bytes = (size_t)p.pitch * (size_t)(p.height + 2 * p.marginY);
            r = n->Broadcast(p.base, p.base, bytes, ncclUint8, 0, d.comm, d.ncclStream);
        }
        if (g->nlocal > 1) { const ncclResult_t r2 = n->GroupEnd(); if (r == ncclSuccess) r = r2; }
        if (r != ncclSuccess) return gfail(g, HMME_ERR_NCCL, std::string("ncclBroadcast: ") + n->GetErrorString(r));
        for (GroupDev& d : g->devs) {
            hmme_ctx* c = d.ctx[slot];
            if (cudaSetDevice(d.device) != cudaSuccess || cudaEventRecord(d.evPost, d.ncclStream) != cudaSuccess ||
                cudaStreamWaitEvent(c->stream, d.evPost, 0) != cudaSuccess)
                return gfail(g, HMME_ERR_CUDA, "event ordering after the NCCL broadcast failed");
        }
    }

    // phase 3 (per GPU, in parallel): search the band, results straight into the band's rows of the caller's tables
    return for_devices(g, [&](int i) -> int {
        GroupDev& d = g->devs[i];
        if (!d.count[slot]) return HMME_OK;
        hmme_ctx* c = d.ctx[slot];
        const int f = d.first[slot], n = d.count[slot];
        int r = hmme_search_frame_async(c, &d.cur[slot], &d.ref[slot], jobs + f, n, range);
        if (r != HMME_OK) return dev_fail(d, c, r);
        const size_t o = (size_t)f * HMME_NUM_CTU_PARTS;
        r = hmme_fetch_results_async(c, n, X + o, Y + o, sad + o, cost ? cost + o : nullptr);
        return r != HMME_OK ? dev_fail(d, c, r) : HMME_OK;
    });
}

int hmme_group_sync(hmme_group* g, int slot) {
    if (!g) return HMME_ERR_ARG;
    if (slot < -1 || slot >= kSlots) return gfail(g, HMME_ERR_ARG, "hmme_group_sync: slot must be in [0, HMME_GROUP_SLOTS) or -1 (all)");
    for (GroupDev& d : g->devs)
        for (int s = 0; s < kSlots; ++s) {
            if (slot >= 0 && s != slot) continue;
            if (d.ncclStream) { cudaSetDevice(d.device); cudaStreamSynchronize(d.ncclStream); }
            const int rc = hmme_sync(d.ctx[s]);
            if (rc != HMME_OK) return gfail(g, rc, "GPU " + std::to_string(d.device) + ": " + hmme_last_error(d.ctx[s]));
        }
    return HMME_OK;
}

int hmme_group_search_frame(hmme_group* g, const void* curHostOrigin, int curHostStride, const void* refHostOrigin, int refHostStride,
                            int hostElemBytes, const hmme_job* jobs, int njobs, int range, int32_t* X, int32_t* Y, uint32_t* sad, uint32_t* cost) {
    const int rc = hmme_group_search_frame_async(g, 0, curHostOrigin, curHostStride, refHostOrigin, refHostStride, hostElemBytes, jobs, njobs, range,
                                                 X, Y, sad, cost);
    return rc != HMME_OK ? rc : hmme_group_sync(g, 0);
}

int hmme_group_pipeline_depth(hmme_group* g, int njobs, int range) {
    if (!g || njobs <= 0) return 2;
    // Frames worth keeping in flight: two while a GPU's band is many waves of thread blocks (the copies of one frame hide behind the other's
    // kernel), three when it is only a few -- then the last, partly filled wave of a search is a sizeable part of it and must overlap the
    // next search, which with two slots is not queued yet while the host waits for a slot's results and refills it.  Measured, 1080p +-64:
    // 29 waves (1 GPU) 1.126 / 1.132 ms per frame with 2 / 3 slots, 3.6 waves (8 GPUs) 0.156 / 0.138.
    // the largest band of the frame (bands differ by at most one job), so that every rank of a one-process-per-GPU job gets the same answer
    const int count = (njobs + g->world - 1) / g->world;
    int ctas = 0, wave = 1;
    if (hmme_search_launch_size(g->devs[0].ctx[0], count, range, &ctas, &wave) == HMME_OK && ctas < kDeepPipelineWaves * wave) return 3;
    return 2;
}

int hmme_group_band(hmme_group* g, int njobs, int localIndex, int* first, int* count) {
    if (!g || localIndex < 0 || localIndex >= g->nlocal) return HMME_ERR_ARG;
    return hmme_band_split(njobs, g->world, g->devs[localIndex].globalRank, first, count);
}

hmme_ctx* hmme_group_context(hmme_group* g, int localIndex, int slot) {
    if (!g || localIndex < 0 || localIndex >= g->nlocal || slot < 0 || slot >= kSlots) return nullptr;
    return g->devs[localIndex].ctx[slot];
}

int hmme_group_last_kernel_ms(hmme_group* g, int slot, float* ms) {
    if (!g || !ms || slot < 0 || slot >= kSlots) return HMME_ERR_ARG;
    float worst = 0.f;
    for (GroupDev& d : g->devs) {
        if (!d.count[slot]) continue;
        float m = 0.f;
        const int rc = hmme_last_kernel_ms(d.ctx[slot], &m);
        if (rc != HMME_OK) return gfail(g, rc, hmme_last_error(d.ctx[slot]));
        worst = std::max(worst, m);
    }
    *ms = worst;
    return HMME_OK;
}

uint64_t hmme_group_kernel_launches(hmme_group* g) {
    uint64_t n = 0;
    if (g) for (GroupDev& d : g->devs) for (int s = 0; s < kSlots; ++s) n += hmme_kernel_launches(d.ctx[s]);
    return n;
}

}  // extern "C"
