/*
 * TEncOpenCL.h -- drop-in replacement of HM-OpenCL's GPU motion-estimation host class.
 *
 * Same class name and public surface as /root/reference/source/Lib/TLibEncoder/TEncOpenCL.h:105-123, so that
 * TEncTop (TEncTop.h:82, TEncTop.cpp:1116-1162), TEncSearch (TEncSearch.cpp:3743-3771) and TEncSlice
 * (TEncSlice.cpp:125,150,799,893) compile and behave unchanged.  Everything behind it is new: no OpenCL, the
 * work goes through the C ABI of libhmme_b200.so (include/hmme_b200.h) to hand-written sm_100a CUDA kernels.
 *
 * Deliberately NOT reproduced from the reference: uninitialised members, use of mapped
 * pointers after unmap, dangling alloca'd device lists, clFinish(NULL) in the destructor, and the silent CPU
 * fallback -- a failed initialisation returns false with a message and leaves the object disabled; calling
 * calcMotionVectors on a disabled object aborts with a message instead of computing on the CPU.
 *
 * Build inside HM: replace TLibEncoder/TEncOpenCL.{h,cpp} by this pair, add -I<repo>/include, link
 * -L<repo>/hm-opencl_b200 -lhmme_b200 instead of -lOpenCL.
 * Build standalone (tests): -DHMME_STANDALONE supplies the handful of HM typedefs this header needs.
 */
#ifndef TENCOPENCL_H
#define TENCOPENCL_H

#ifdef HMME_STANDALONE
#include <stdint.h>
typedef void Void;
typedef bool Bool;
typedef char Char;
typedef short Short;
typedef int Int;
typedef unsigned int UInt;
typedef double Double;
typedef Short Pel;               // TypeDef.h:706
typedef UInt Distortion;         // TypeDef.h:717
#define NUM_CTU_PARTS 593        // TypeDef.h:263
class TComMv {                   // TComMv.h:54-55: a pair of Short
    Short m_iHor, m_iVer;
public:
    TComMv() : m_iHor(0), m_iVer(0) {}
    TComMv(Short h, Short v) : m_iHor(h), m_iVer(v) {}
    Short getHor() const { return m_iHor; }
    Short getVer() const { return m_iVer; }
};
#else
#include "TLibCommon/TypeDef.h"
#include "TLibCommon/TComMv.h"
#endif

#include <string>

struct hmme_ctx;
struct TEncOpenCLSpec;                          // state of the speculative whole-frame search (TEncOpenCL.cpp)

class TEncOpenCL {
protected:
    hmme_ctx*    m_ctx;                         ///< CUDA context of libhmme_b200 (NULL until createBuffers succeeds)
    Bool         deviceFound;                   ///< findDevice succeeded
    Bool         compileKernel;                 ///< compileKernelSource accepted the request
    Int          deviceId;                      ///< CUDA device ordinal (option OpenCLDevice)
    Bool         enabled;
    std::string  info;                          ///< device name (getDeviceInfo)
    Int          searchRange;                   ///< range createBuffers was sized for
    Double       m_lambdaDouble;                ///< last value given to setLambda (applied once the context exists)
    UInt         m_lambda;                      ///< (UInt)floor(65536*sqrt(lambda)), TEncOpenCL.h:121 of the reference

    Int          Xarray[NUM_CTU_PARTS];         ///< integer-pel MV x of the best candidate per partition
    Int          Yarray[NUM_CTU_PARTS];
    Distortion   minSad[NUM_CTU_PARTS];         ///< SAD + MV-bit cost of the winner
    Distortion   ruiCosts[NUM_CTU_PARTS];       ///< pure SAD of the winner
    TEncOpenCLSpec* m_spec;                     ///< NULL until beginPicture is used

public:
    TEncOpenCL();
    virtual         ~TEncOpenCL();
    Bool            compileKernelSource(const Char* fileName, const Char* kernelNameCalc);
    Bool            findDevice(Int device);
    Bool            findDevices(Int device) { return findDevice(device); }   ///< the reference README's spelling
    Bool            createBuffers(UInt i_maxCtuWidth, UInt i_maxCtuHeight, Int i_searchRange);
    Void            calcMotionVectors(Pel* pelCtu, Pel* pelSearch, Int i_iRefStride, Int i_iCtuStride, Int i_areaSize, TComMv* pcMvSrchRngLT);
    /// Addition to the reference's surface: the fractional-pel refinement of one PU on the GPU, with the
    /// arguments TEncSearch::xPatternSearchFracDIF (TEncSearch.cpp:4294-4331) has at hand: pattern key block, piRefY at the PU origin,
    /// integer MV, the predictor set in TComRdCost, HadamardME && !lossless.  Returns ruiCost; rcMvHalf / rcMvQter receive a
    /// decomposition of the winning quarter-pel offset (only their sum (half << 1) + qter is used by the caller, :3800-3803).
    Distortion      refineFractional(Pel* pelKey, Int iKeyStride, Int iWidth, Int iHeight, Pel* piRefY, Int iRefStride, const TComMv& rcMvInt,
                                     const TComMv& rcMvPred, Bool bUseHadamard, TComMv& rcMvHalf, TComMv& rcMvQter);

    /// Addition: SAD (or Hadamard SATD) between the original block and the uni-directional motion-compensated
    /// prediction at an already clipped quarter-pel MV -- the distortion TEncSearch::xGetTemplateCost (TEncSearch.cpp:3634-3674)
    /// computes with xPredInterBlk + getDistPart(DF_SAD) before it adds its rate term.
    Distortion      templateDistortion(Pel* pelOrg, Int iOrgStride, Int iWidth, Int iHeight, Pel* piRefY, Int iRefStride, const TComMv& rcMvClipped,
                                       Bool bUseHadamard);

    /// Addition (row f3): the distortion TEncSearch::xGetInterPredictionError (TEncSearch.cpp:2814-2836) computes after motionCompensation, for a
    /// uni-directional PU (piRefY1 == NULL) or a bi-directional one (xPredInterBi + addAvg); MVs already clipped, plane pointers at the PU origin.
    Distortion      interPredictionError(Pel* pelOrg, Int iOrgStride, Int iWidth, Int iHeight, Pel* piRefY0, Int iRefStride0, const TComMv& rcMv0,
                                         Pel* piRefY1, Int iRefStride1, const TComMv& rcMv1, Bool bUseHadamard);

    /// Additions: speculative whole-frame search.  The reference searches one CTU per calcMotionVectors call
    /// because the window position (pcMvSrchRngLT) depends on the AMVP predictor of the CTU being coded.  Here the slice encoder announces the
    /// picture before its CTU loop (TEncSlice::compressSlice, TEncSlice.cpp:730; the reference pictures are final and border-extended since
    /// TComSlice::setRefPicList, TComSlice.cpp:351-377): beginPicture(original luma) + addReferencePicture(reconstructed luma) per reference +
    /// speculate(range).  Every full CTU of the picture is then searched against every reference in ONE launch per reference, with the window a
    /// zero predictor gives, into device-resident [hypothesis][ctu][593] tables (hmme_table).  calcMotionVectors keeps its signature: it recognises
    /// the CTU and the reference picture from its pointers, and when block content, window position, range and lambda equal a table entry it
    /// answers from the table; otherwise it runs the synchronous search as before AND re-speculates the CTUs still to come with the window
    /// centre it just saw (motion is coherent, so the next CTUs usually hit).  Results are those of the same kernels on the same inputs: the
    /// bitstream cannot change.  HMME_SPEC_VERIFY=1 makes every hit also run the synchronous search and abort on any difference.
    Void            beginPicture(const Pel* orgLuma, Int orgStride, Int width, Int height);
    Void            addReferencePicture(const Pel* recLuma, Int recStride, Int marginX, Int marginY);
    Void            speculate(Int searchRange);
    Void            endPicture();
    struct SpecStats { unsigned long long calls, hits, missBlock, missWindow, missOther, speculations, jobsSpeculated; };
    SpecStats       getSpecStats() const;

    //======== getters and setters ================
    Int             getDeviceId         ()              { return deviceId; }
    Void            setDeviceId         ( Int i )       { deviceId = i; }
    const Char*     getDeviceInfo       ()              { return info.c_str(); }
    Distortion*     getRuiCost          ()              { return ruiCosts; }
    Int*            getX                ()              { return Xarray; }
    Int*            getY                ()              { return Yarray; }
    Void            setLambda           (Double lambda);
    Void            setEnabled          (Bool e)        { enabled = e; }
    Bool            isEnabled           () const        { return enabled; }
    const Char*     getLastError        () const;

private:
    TEncOpenCL(const TEncOpenCL&);              // owns a device context: not copyable
    TEncOpenCL& operator=(const TEncOpenCL&);
};

#endif /* TENCOPENCL_H */
