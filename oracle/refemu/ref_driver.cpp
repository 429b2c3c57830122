// ref_driver.cpp -- C-ABI handle on the reference's own TEncOpenCL class (compiled unmodified from
// /root/reference/source/Lib/TLibEncoder/TEncOpenCL.cpp) running over oracle/refemu/fake_cl.cpp.
// TEST INFRASTRUCTURE ONLY: used by oracle/gen_golden.py and tests/test_oracle.py (when oracle/_ref is
// built) to pin the restated oracle against the reference itself.
#include <cstdint>
#include <cstring>

#include "TLibEncoder/TEncOpenCL.h"

namespace {
// protected members (m_lambda, minSad) are reachable from a derived class without touching the source
struct RefME : public TEncOpenCL {
    void setLambdaRaw(UInt v) { m_lambda = v; }
    UInt lambdaRaw() const { return m_lambda; }
    const Distortion* minSadPtr() const { return minSad; }
};
}  // namespace

extern "C" {

// Mirrors TEncTop::xInitOpenCL (TEncTop.cpp:1116-1162): findDevice -> compileKernelSource -> createBuffers -> setEnabled.
void* hmref_create(const char* kernelPath, int searchRange) {
    RefME* me = new RefME;
    me->setEnabled(false);                      // the reference ctor leaves `enabled` uninitialised (App. B9)
    if (!me->findDevice(0)) { delete me; return nullptr; }
    if (!me->compileKernelSource(kernelPath, "calcSAD_AMP")) { delete me; return nullptr; }
    if (!me->createBuffers(64, 64, searchRange)) { delete me; return nullptr; }
    me->setEnabled(true);
    me->setLambdaRaw(0);
    return me;
}
void hmref_destroy(void* h) { delete static_cast<RefME*>(h); }
void hmref_set_lambda(void* h, double lambda) { static_cast<RefME*>(h)->setLambda(lambda); }
void hmref_set_lambda_q16(void* h, uint32_t v) { static_cast<RefME*>(h)->setLambdaRaw(v); }
uint32_t hmref_get_lambda_q16(void* h) { return static_cast<RefME*>(h)->lambdaRaw(); }

// One TEncOpenCL::calcMotionVectors call (TEncOpenCL.cpp:240-362) + the getters TEncSearch reads (:3752-3764).
void hmref_calc(void* h, const int16_t* cur, const int16_t* refAtCtu, int refStride, int range, int ltx, int lty,
                int32_t* X, int32_t* Y, uint32_t* sad, uint32_t* cost) {
    RefME* me = static_cast<RefME*>(h);
    TComMv lt((Short)ltx, (Short)lty);
    me->calcMotionVectors(const_cast<Pel*>(cur), const_cast<Pel*>(refAtCtu), refStride, 64, range, &lt);
    std::memcpy(X, me->getX(), sizeof(Int) * NUM_CTU_PARTS);
    std::memcpy(Y, me->getY(), sizeof(Int) * NUM_CTU_PARTS);
    std::memcpy(sad, me->getRuiCost(), sizeof(Distortion) * NUM_CTU_PARTS);
    std::memcpy(cost, me->minSadPtr(), sizeof(Distortion) * NUM_CTU_PARTS);
}
int hmref_num_parts(void) { return NUM_CTU_PARTS; }

}  // extern "C"
