/*
 * TEncOpenCL.cpp -- see TEncOpenCL.h.  Host logic only; the arithmetic lives in libhmme_b200.so.
 * Mirrors the call order of TEncTop::xInitOpenCL (TEncTop.cpp:1129-1145):
 *   findDevice -> compileKernelSource -> createBuffers -> setEnabled(true), then per CTU
 *   calcMotionVectors + getX/getY/getRuiCost (TEncSearch.cpp:3749-3764), setLambda per slice/CTU.
 */
#include "TEncOpenCL.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "hmme_b200.h"

// ---- speculative whole-frame search (see TEncOpenCL.h) ---------------------------------------------------------------------------
namespace {
const int kSpecHyp = 3;                          // window hypotheses kept per reference picture (least recently used is replaced)

struct SpecRef {
    const Pel* hostOrigin; Int stride, marginX, marginY;
    hmme_plane plane;                            // device copy (8 bit)
};
struct SpecSlot {                                // one hypothesis: every CTU from firstCtu on, searched with the windows `lt`
    bool valid, ready; Int range, firstCtu, centreX, centreY; UInt lambda; unsigned long long stamp;
    std::vector<int32_t> lt;                     // [nctu][2]; INT32_MIN where the job had to be replaced (never hits)
};
}

struct TEncOpenCLSpec {
    hmme_ctx* ctx;                               // its own context / stream: the synchronous per-CTU calls never queue behind a whole-frame search
    const Pel* org; Int orgStride, width, height, ncx, ncy;
    hmme_plane orgPlane; bool orgAlloc;
    std::vector<SpecRef> refs;
    std::vector<SpecSlot> slots;                 // [ref][kSpecHyp]
    hmme_table* table; Int tableSlots, tableJobs;
    int32_t* mirror; size_t mirrorBytes;         // pinned host copy of the tables: [slot][4][nctu][593]
    std::vector<hmme_job> jobs;
    bool active, verify;
    unsigned long long clock;
    TEncOpenCL::SpecStats st;
    TEncOpenCLSpec() : ctx(NULL), org(NULL), orgStride(0), width(0), height(0), ncx(0), ncy(0), orgAlloc(false), table(NULL), tableSlots(0), tableJobs(0),
                       mirror(NULL), mirrorBytes(0), active(false), verify(false), clock(0) { memset(&orgPlane, 0, sizeof(orgPlane)); memset(&st, 0, sizeof(st)); }
};

static void specFatal(const char* what, hmme_ctx* c) {
    fprintf(stderr, "FATAL: speculative search: %s: %s\n", what, hmme_last_error(c));
    abort();
}

TEncOpenCL::TEncOpenCL()
    : m_ctx(NULL), deviceFound(false), compileKernel(false), deviceId(0), enabled(false), searchRange(0), m_lambdaDouble(0.0), m_lambda(0), m_spec(NULL) {
    memset(Xarray, 0, sizeof(Xarray));
    memset(Yarray, 0, sizeof(Yarray));
    memset(ruiCosts, 0, sizeof(ruiCosts));
    for (int i = 0; i < NUM_CTU_PARTS; i++) minSad[i] = 0xFFFFFFFFu;
}

TEncOpenCL::~TEncOpenCL() {
    // constructed in every encoder run, GPU or not (TEncTop.h:82): must be safe without a device
    if (m_spec) {
        if (m_spec->st.calls)
            printf("HMME_SPEC calls=%llu hits=%llu miss_block=%llu miss_window=%llu miss_other=%llu speculations=%llu jobs=%llu\n", m_spec->st.calls, m_spec->st.hits,
                   m_spec->st.missBlock, m_spec->st.missWindow, m_spec->st.missOther, m_spec->st.speculations, m_spec->st.jobsSpeculated);
        if (m_spec->ctx) {
            hmme_sync(m_spec->ctx);
            if (m_spec->table) hmme_table_destroy(m_spec->table);
            if (m_spec->mirror) hmme_host_free(m_spec->mirror);
            if (m_spec->orgAlloc) hmme_plane_free(m_spec->ctx, &m_spec->orgPlane);
            for (size_t i = 0; i < m_spec->refs.size(); i++) hmme_plane_free(m_spec->ctx, &m_spec->refs[i].plane);
            hmme_destroy(m_spec->ctx);
        }
        delete m_spec;
        m_spec = NULL;
    }
    if (m_ctx) hmme_destroy(m_ctx);
    m_ctx = NULL;
}

TEncOpenCL::SpecStats TEncOpenCL::getSpecStats() const {
    SpecStats z; memset(&z, 0, sizeof(z));
    return m_spec ? m_spec->st : z;
}

Void TEncOpenCL::beginPicture(const Pel* orgLuma, Int orgStride, Int width, Int height) {
    if (!m_ctx || !enabled) return;
    if (!m_spec) {
        m_spec = new TEncOpenCLSpec;
        if (hmme_create(&m_spec->ctx, deviceId, 64, 64, searchRange) != HMME_OK) specFatal("hmme_create", NULL);
        const char* v = getenv("HMME_SPEC_VERIFY");
        m_spec->verify = v && atoi(v) != 0;
    }
    TEncOpenCLSpec& S = *m_spec;
    hmme_sync(S.ctx);                                            // nothing of the previous picture is in flight any more
    if (S.orgAlloc && (S.width != width || S.height != height)) { hmme_plane_free(S.ctx, &S.orgPlane); S.orgAlloc = false; }
    S.org = orgLuma; S.orgStride = orgStride; S.width = width; S.height = height; S.ncx = width / 64; S.ncy = height / 64;
    if (!S.orgAlloc) {
        if (hmme_plane_alloc(S.ctx, &S.orgPlane, 1, width, height, 0, 0) != HMME_OK) specFatal("hmme_plane_alloc", S.ctx);
        S.orgAlloc = true;
    }
    for (size_t i = 0; i < S.refs.size(); i++) hmme_plane_free(S.ctx, &S.refs[i].plane);
    S.refs.clear();
    S.slots.clear();
    S.active = false;
    if (S.ncx <= 0 || S.ncy <= 0) return;
    if (hmme_plane_upload_s16_async(S.ctx, &S.orgPlane, orgLuma, orgStride) != HMME_OK) specFatal("upload of the original picture", S.ctx);
}

Void TEncOpenCL::addReferencePicture(const Pel* recLuma, Int recStride, Int marginX, Int marginY) {
    if (!m_spec || !m_spec->org || m_spec->ncx <= 0 || m_spec->ncy <= 0) return;
    TEncOpenCLSpec& S = *m_spec;
    for (size_t i = 0; i < S.refs.size(); i++)
        if (S.refs[i].hostOrigin == recLuma) return;            // the same picture in both lists
    SpecRef r;
    r.hostOrigin = recLuma; r.stride = recStride; r.marginX = marginX; r.marginY = marginY;
    if (hmme_plane_alloc(S.ctx, &r.plane, 1, S.width, S.height, marginX, marginY) != HMME_OK) specFatal("hmme_plane_alloc", S.ctx);
    // the window addressing of the reference is linear in the HOST plane (row wrap, App. B4): the device copy must have the same pitch
    if (r.plane.pitch != recStride) { hmme_plane_free(S.ctx, &r.plane); return; }
    if (hmme_plane_upload_s16_async(S.ctx, &r.plane, recLuma, recStride) != HMME_OK) specFatal("upload of a reference picture", S.ctx);
    S.refs.push_back(r);
}

// enqueue hypothesis `slot` of reference `ri`: CTUs firstCtu.. with the windows the quarter-pel centre (4cx, 4cy) gives
static void specLaunch(TEncOpenCLSpec& S, UInt lambda, Int ri, Int h, Int range, Int cx, Int cy, Int firstCtu) {
    const Int nctu = S.ncx * S.ncy, n = nctu - firstCtu;
    if (n <= 0) return;
    SpecSlot& sl = S.slots[ri * kSpecHyp + h];
    sl.valid = false;
    sl.lt.assign((size_t)nctu * 2, INT32_MIN);
    S.jobs.resize(n);
    const SpecRef& R = S.refs[ri];
    const long long rows = S.height + 2 * R.marginY, alloc = rows * (long long)R.stride;
    for (Int k = 0; k < n; k++) {
        const Int ctu = firstCtu + k, x = (ctu % S.ncx) * 64, y = (ctu / S.ncx) * 64;
        int ltx = 0, lty = 0;
        hmme_search_window(4 * cx, 4 * cy, range, x, y, S.width, S.height, &ltx, &lty, NULL, NULL);
        const long long r0 = (long long)(R.marginY + y + lty) * R.stride + R.marginX + x + ltx, r1 = r0 + (long long)(2 * range + 63) * R.stride + 2 * range + 63;
        if (r0 < 0 || r1 >= alloc) {                             // the reference would read outside its allocation here: never served from the table
            S.jobs[k].ctuX = x; S.jobs[k].ctuY = y; S.jobs[k].ltx = 0; S.jobs[k].lty = 0;
            const long long q0 = (long long)(R.marginY + y) * R.stride + R.marginX + x, q1 = q0 + (long long)(2 * range + 63) * R.stride + 2 * range + 63;
            if (q1 >= alloc) return;                             // not even a centred window fits (tiny margins): no speculation
            continue;
        }
        S.jobs[k].ctuX = x; S.jobs[k].ctuY = y; S.jobs[k].ltx = ltx; S.jobs[k].lty = lty;
        sl.lt[2 * ctu] = ltx; sl.lt[2 * ctu + 1] = lty;
    }
    hmme_set_lambda_q16(S.ctx, lambda);
    const Int ts = ri * kSpecHyp + h;
    if (hmme_search_frame_table_async(S.ctx, &S.orgPlane, &R.plane, &S.jobs[0], n, range, S.table, ts) != HMME_OK) specFatal("hmme_search_frame_table_async", S.ctx);
    const size_t plane = (size_t)S.tableJobs * NUM_CTU_PARTS;
    int32_t* m = S.mirror + (size_t)ts * 4 * plane;
    if (hmme_table_fetch_async(S.ctx, S.table, ts, 0, n, m, m + plane, reinterpret_cast<uint32_t*>(m + 2 * plane), reinterpret_cast<uint32_t*>(m + 3 * plane)) != HMME_OK)
        specFatal("hmme_table_fetch_async", S.ctx);
    sl.valid = true; sl.ready = false; sl.range = range; sl.firstCtu = firstCtu; sl.centreX = cx; sl.centreY = cy; sl.lambda = lambda; sl.stamp = ++S.clock;
    S.st.speculations++; S.st.jobsSpeculated += (unsigned long long)n;
}

Void TEncOpenCL::speculate(Int range) {
    if (!m_spec || m_spec->refs.empty() || range < 0 || range > searchRange) return;
    TEncOpenCLSpec& S = *m_spec;
    const Int nctu = S.ncx * S.ncy, want = (Int)S.refs.size() * kSpecHyp;
    if (!S.table || S.tableSlots < want || S.tableJobs < nctu) {
        hmme_sync(S.ctx);
        if (S.table) hmme_table_destroy(S.table);
        if (S.mirror) hmme_host_free(S.mirror);
        S.table = NULL; S.mirror = NULL;
        if (hmme_table_create(S.ctx, &S.table, want, nctu) != HMME_OK) specFatal("hmme_table_create", S.ctx);
        S.tableSlots = want; S.tableJobs = nctu;
        S.mirrorBytes = (size_t)want * 4 * nctu * NUM_CTU_PARTS * sizeof(int32_t);
        S.mirror = static_cast<int32_t*>(hmme_host_alloc(S.mirrorBytes));
        if (!S.mirror) specFatal("hmme_host_alloc", S.ctx);
    }
    S.slots.assign(S.refs.size() * kSpecHyp, SpecSlot());
    for (size_t i = 0; i < S.slots.size(); i++) { S.slots[i].valid = false; S.slots[i].stamp = 0; }
    for (size_t ri = 0; ri < S.refs.size(); ri++) specLaunch(S, m_lambda, (Int)ri, 0, range, 0, 0, 0);   // first hypothesis: zero predictor
    S.active = true;
}

Void TEncOpenCL::endPicture() {
    if (m_spec) { hmme_sync(m_spec->ctx); m_spec->active = false; }
}


const Char* TEncOpenCL::getLastError() const { return hmme_last_error(m_ctx); }

Bool TEncOpenCL::findDevice(Int device) {
    int n = 0;
    deviceFound = false;
    if (hmme_device_count(&n) != HMME_OK || n <= 0) {
        fprintf(stderr, "ERROR: No CUDA devices found ( %s )\n", hmme_last_error(NULL));
        return false;                                   // the caller must not enable GPU ME (the reference crashes here, App. B10)
    }
    if (device < 0 || device > n - 1) {                 // TEncOpenCL.cpp:111-115 of the reference
        device = 0;
        printf("ID device not found, use default GPU device \n");
    }
    deviceId = device;
    deviceFound = true;
    return true;
}

Bool TEncOpenCL::compileKernelSource(const Char* fileName, const Char* kernelNameCalc) {
    // Kernels are precompiled sm_100a code inside libhmme_b200.so.  The option KernelOpenCL must still be given
    // (TEncTop.cpp:1131 disables GPU ME when it is NULL) but the file is not read; only the 593-partition layout
    // (kernel "calcSAD_AMP", AMP_ENC_SPEEDUP = 0) exists.
    compileKernel = false;
    if (fileName == NULL || kernelNameCalc == NULL) {
        fprintf(stderr, "ERROR: Reading Kernel ( -1 )\n");
        return false;
    }
    if (strcmp(kernelNameCalc, "calcSAD_AMP") != 0) {
        fprintf(stderr, "ERROR: kernel '%s' is not available: only calcSAD_AMP (NUM_CTU_PARTS = 593) is implemented\n", kernelNameCalc);
        return false;
    }
    compileKernel = true;
    return true;
}

Bool TEncOpenCL::createBuffers(UInt i_maxCtuWidth, UInt i_maxCtuHeight, Int i_searchRange) {
    if (!deviceFound || !compileKernel) {
        fprintf(stderr, "ERROR: createBuffers called before findDevice/compileKernelSource succeeded\n");
        return false;
    }
    if (m_ctx) { hmme_destroy(m_ctx); m_ctx = NULL; }
    const int rc = hmme_create(&m_ctx, deviceId, (int)i_maxCtuWidth, (int)i_maxCtuHeight, i_searchRange);
    if (rc != HMME_OK) {
        fprintf(stderr, "ERROR: hmme_create ( %d ): %s\n", rc, hmme_last_error(NULL));
        m_ctx = NULL;
        return false;
    }
    searchRange = i_searchRange;
    info = hmme_device_name(m_ctx);
    hmme_set_lambda(m_ctx, m_lambdaDouble);
    printf("Using GPU device              : %s\n", info.c_str());
    return true;
}

Void TEncOpenCL::setLambda(Double lambda) {
    m_lambdaDouble = lambda;
    m_lambda = (UInt)floor(65536.0 * sqrt(lambda));
    if (m_ctx) hmme_set_lambda(m_ctx, lambda);
}

Void TEncOpenCL::calcMotionVectors(Pel* pelCtu, Pel* pelSearch, Int i_iRefStride, Int i_iCtuStride, Int i_areaSize, TComMv* pcMvSrchRngLT) {
    if (!m_ctx || !enabled) {
        fprintf(stderr, "FATAL: TEncOpenCL::calcMotionVectors called without an initialised, enabled GPU context (there is no CPU fallback)\n");
        abort();
    }
    // ---- speculative whole-frame search: is the answer already in the tables?
    Int specRef = -1, specCtu = -1;
    bool specBlockOk = false;
    if (m_spec && m_spec->active) {
        TEncOpenCLSpec& S = *m_spec;
        S.st.calls++;
        for (size_t i = 0; i < S.refs.size() && specRef < 0; i++) {
            const SpecRef& R = S.refs[i];
            if (i_iRefStride != R.stride) continue;
            const ptrdiff_t off = pelSearch - R.hostOrigin;
            if (off < 0 || off >= (ptrdiff_t)R.stride * S.height) continue;
            const Int y = (Int)(off / R.stride), x = (Int)(off - (ptrdiff_t)y * R.stride);
            if (x >= S.width || (x & 63) || (y & 63) || x + 64 > S.width || y + 64 > S.height) continue;
            specRef = (Int)i; specCtu = (y / 64) * S.ncx + x / 64;
            specBlockOk = true;                                  // the block must be the original picture's (not the bi-prediction residual block)
            for (Int r = 0; r < 64 && specBlockOk; r++)
                specBlockOk = memcmp(pelCtu + (ptrdiff_t)r * i_iCtuStride, S.org + (ptrdiff_t)(y + r) * S.orgStride + x, 64 * sizeof(Pel)) == 0;
        }
        if (specRef >= 0 && specBlockOk) {
            for (Int h = 0; h < kSpecHyp; h++) {
                SpecSlot& sl = S.slots[specRef * kSpecHyp + h];
                if (!sl.valid || sl.range != i_areaSize || sl.lambda != m_lambda || specCtu < sl.firstCtu) continue;
                if (sl.lt[2 * specCtu] != pcMvSrchRngLT->getHor() || sl.lt[2 * specCtu + 1] != pcMvSrchRngLT->getVer()) continue;
                if (!sl.ready) {                                 // first use of results that may still be in flight
                    if (hmme_sync(S.ctx) != HMME_OK) specFatal("hmme_sync", S.ctx);
                    for (size_t q = 0; q < S.slots.size(); q++) S.slots[q].ready = S.slots[q].valid;
                }
                const size_t plane = (size_t)S.tableJobs * NUM_CTU_PARTS;
                const int32_t* m = S.mirror + (size_t)(specRef * kSpecHyp + h) * 4 * plane + (size_t)(specCtu - sl.firstCtu) * NUM_CTU_PARTS;
                memcpy(Xarray, m, sizeof(Xarray)); memcpy(Yarray, m + plane, sizeof(Yarray));
                memcpy(ruiCosts, m + 2 * plane, sizeof(ruiCosts)); memcpy(minSad, m + 3 * plane, sizeof(minSad));
                sl.stamp = ++S.clock;
                S.st.hits++;
                if (S.verify) {                                  // test mode: the synchronous search must give the same 4 x 593 values
                    int32_t vx[NUM_CTU_PARTS], vy[NUM_CTU_PARTS]; uint32_t vs[NUM_CTU_PARTS], vc[NUM_CTU_PARTS];
                    if (hmme_search_ctu(m_ctx, pelCtu, i_iCtuStride, pelSearch, i_iRefStride, i_areaSize, pcMvSrchRngLT->getHor(), pcMvSrchRngLT->getVer(), vx, vy, vs, vc) != HMME_OK ||
                        memcmp(vx, Xarray, sizeof(vx)) || memcmp(vy, Yarray, sizeof(vy)) || memcmp(vs, ruiCosts, sizeof(vs)) || memcmp(vc, minSad, sizeof(vc))) {
                        fprintf(stderr, "FATAL: speculative table entry differs from the synchronous search (ctu %d, ref %d)\n", specCtu, specRef);
                        abort();
                    }
                }
                return;
            }
            S.st.missWindow++;
        } else if (specRef >= 0) S.st.missBlock++;
        else S.st.missOther++;
    }
    const int rc = hmme_search_ctu(m_ctx, pelCtu, i_iCtuStride, pelSearch, i_iRefStride, i_areaSize, pcMvSrchRngLT->getHor(),
                                   pcMvSrchRngLT->getVer(), Xarray, Yarray, ruiCosts, minSad);
    if (rc != HMME_OK) {
        fprintf(stderr, "FATAL: hmme_search_ctu ( %d ): %s\n", rc, hmme_last_error(m_ctx));
        abort();
    }
    // ---- a window the tables did not have: the CTUs still to come are re-searched with this window centre while the encoder works on this CTU
    if (specRef >= 0 && specBlockOk && i_areaSize <= searchRange) {
        TEncOpenCLSpec& S = *m_spec;
        const Int cx = pcMvSrchRngLT->getHor() + i_areaSize, cy = pcMvSrchRngLT->getVer() + i_areaSize;
        Int victim = -1;
        bool known = false;
        for (Int h = 0; h < kSpecHyp; h++) {
            const SpecSlot& sl = S.slots[specRef * kSpecHyp + h];
            if (sl.valid && sl.range == i_areaSize && sl.lambda == m_lambda && sl.centreX == cx && sl.centreY == cy && specCtu >= sl.firstCtu) known = true;   // only clipping differed here
            if (!sl.valid && victim < 0) victim = h;
        }
        if (victim < 0) {                                        // all in use: replace the least recently used hypothesis
            victim = 0;
            for (Int h = 1; h < kSpecHyp; h++)
                if (S.slots[specRef * kSpecHyp + h].stamp < S.slots[specRef * kSpecHyp + victim].stamp) victim = h;
        }
        if (!known) specLaunch(S, m_lambda, specRef, victim, i_areaSize, cx, cy, specCtu + 1);
    }
}

Distortion TEncOpenCL::refineFractional(Pel* pelKey, Int iKeyStride, Int iWidth, Int iHeight, Pel* piRefY, Int iRefStride, const TComMv& rcMvInt,
                                        const TComMv& rcMvPred, Bool bUseHadamard, TComMv& rcMvHalf, TComMv& rcMvQter) {
    if (!m_ctx || !enabled) {
        fprintf(stderr, "FATAL: TEncOpenCL::refineFractional called without an initialised, enabled GPU context (there is no CPU fallback)\n");
        abort();
    }
    int32_t qx = 0, qy = 0;
    uint32_t cost = 0;
    const int rc = hmme_refine_pu(m_ctx, pelKey, iKeyStride, piRefY, iRefStride, iWidth, iHeight, rcMvInt.getHor(), rcMvInt.getVer(),
                                  rcMvPred.getHor(), rcMvPred.getVer(), bUseHadamard ? 1 : 0, &qx, &qy, &cost, NULL);
    if (rc != HMME_OK) {
        fprintf(stderr, "FATAL: hmme_refine_pu ( %d ): %s\n", rc, hmme_last_error(m_ctx));
        abort();
    }
    // quarter-pel offset from the integer MV, in [-3, 3]: split as 2*half + qter with both parts in [-1, 1]
    const int sx = qx - 4 * rcMvInt.getHor(), sy = qy - 4 * rcMvInt.getVer();
    const int hx = sx / 2, hy = sy / 2;
    rcMvHalf = TComMv((Short)hx, (Short)hy);
    rcMvQter = TComMv((Short)(sx - 2 * hx), (Short)(sy - 2 * hy));
    return (Distortion)cost;
}

Distortion TEncOpenCL::templateDistortion(Pel* pelOrg, Int iOrgStride, Int iWidth, Int iHeight, Pel* piRefY, Int iRefStride, const TComMv& rcMvClipped,
                                          Bool bUseHadamard) {
    if (!m_ctx || !enabled) {
        fprintf(stderr, "FATAL: TEncOpenCL::templateDistortion called without an initialised, enabled GPU context (there is no CPU fallback)\n");
        abort();
    }
    uint32_t dist = 0;
    const int rc = hmme_mc_cost_pu(m_ctx, pelOrg, iOrgStride, piRefY, iRefStride, iWidth, iHeight, rcMvClipped.getHor(), rcMvClipped.getVer(),
                                   bUseHadamard ? 1 : 0, &dist);
    if (rc != HMME_OK) {
        fprintf(stderr, "FATAL: hmme_mc_cost_pu ( %d ): %s\n", rc, hmme_last_error(m_ctx));
        abort();
    }
    return (Distortion)dist;
}

Distortion TEncOpenCL::interPredictionError(Pel* pelOrg, Int iOrgStride, Int iWidth, Int iHeight, Pel* piRefY0, Int iRefStride0, const TComMv& rcMv0,
                                            Pel* piRefY1, Int iRefStride1, const TComMv& rcMv1, Bool bUseHadamard) {
    if (!m_ctx || !enabled) {
        fprintf(stderr, "FATAL: TEncOpenCL::interPredictionError called without an initialised, enabled GPU context (there is no CPU fallback)\n");
        abort();
    }
    uint32_t dist = 0;
    const int rc = piRefY1
        ? hmme_mc_cost_bi_pu(m_ctx, pelOrg, iOrgStride, piRefY0, iRefStride0, rcMv0.getHor(), rcMv0.getVer(), piRefY1, iRefStride1, rcMv1.getHor(),
                             rcMv1.getVer(), iWidth, iHeight, bUseHadamard ? 1 : 0, &dist)
        : hmme_mc_cost_pu(m_ctx, pelOrg, iOrgStride, piRefY0, iRefStride0, iWidth, iHeight, rcMv0.getHor(), rcMv0.getVer(), bUseHadamard ? 1 : 0, &dist);
    if (rc != HMME_OK) {
        fprintf(stderr, "FATAL: hmme_mc_cost%s_pu ( %d ): %s\n", piRefY1 ? "_bi" : "", rc, hmme_last_error(m_ctx));
        abort();
    }
    return (Distortion)dist;
}
