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

#include "hmme_b200.h"

TEncOpenCL::TEncOpenCL()
    : m_ctx(NULL), deviceFound(false), compileKernel(false), deviceId(0), enabled(false), searchRange(0), m_lambdaDouble(0.0), m_lambda(0) {
    memset(Xarray, 0, sizeof(Xarray));
    memset(Yarray, 0, sizeof(Yarray));
    memset(ruiCosts, 0, sizeof(ruiCosts));
    for (int i = 0; i < NUM_CTU_PARTS; i++) minSad[i] = 0xFFFFFFFFu;
}

TEncOpenCL::~TEncOpenCL() {
    // constructed in every encoder run, GPU or not (TEncTop.h:82): must be safe without a device
    if (m_ctx) hmme_destroy(m_ctx);
    m_ctx = NULL;
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
    const int rc = hmme_search_ctu(m_ctx, pelCtu, i_iCtuStride, pelSearch, i_iRefStride, i_areaSize, pcMvSrchRngLT->getHor(),
                                   pcMvSrchRngLT->getVer(), Xarray, Yarray, ruiCosts, minSad);
    if (rc != HMME_OK) {
        fprintf(stderr, "FATAL: hmme_search_ctu ( %d ): %s\n", rc, hmme_last_error(m_ctx));
        abort();
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
