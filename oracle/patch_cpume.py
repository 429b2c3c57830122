#!/usr/bin/env python
"""Instruments a SCRATCH COPY of the reference's TEncSearch.cpp for the CPU-ME baseline:
wall-clock around the OpenCL=0 integer search calls of xMotionEstimation (TEncSearch.cpp:3774-3791: full search and TZ)
and a counter of the DistFunc invocations made inside them (:390-426, :3878), the same timer around the fractional-pel refinement
(xPatternSearchFracDIF, :4294-4331), and optional logs (search-window placements, fractional-refinement records).  Arithmetic and control flow are untouched; the bitstream stays
identical.  Used only by `make -C oracle encoders` on a temporary copy; nothing patched is ever committed."""
import sys

path = sys.argv[1]
s = open(path).read()


def once(old, new):
    global s
    assert s.count(old) == 1, (old, s.count(old))
    s = s.replace(old, new)


once('#include "TEncSearch.h"', '''#include "TEncSearch.h"
#include <time.h>
#include <stdio.h>
#include <stdlib.h>
static double g_hmmeMeSeconds = 0.0;
static unsigned long long g_hmmeDistCalls = 0, g_hmmeMeCalls = 0, g_hmmeInMe = 0;
static double hmmeNow() { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
static double g_hmmeFracSeconds = 0.0; static unsigned long long g_hmmeFracCalls = 0, g_hmmeFracPixels = 0;
static void hmmeReport() { printf("HMME_CPUME me_seconds=%.6f dist_calls=%llu me_calls=%llu frac_seconds=%.6f frac_calls=%llu frac_pixels=%llu\\n", g_hmmeMeSeconds, g_hmmeDistCalls, g_hmmeMeCalls, g_hmmeFracSeconds, g_hmmeFracCalls, g_hmmeFracPixels); }
static struct HmmeReportInit { HmmeReportInit() { atexit(hmmeReport); } } g_hmmeReportInit;''')
once('          xPatternSearch      ( pcPatternKey, piRefY, iRefStride, &cMvSrchRngLT, &cMvSrchRngRB, rcMv, ruiCost );',
     '          { const double t0_ = hmmeNow(); g_hmmeInMe = 1; xPatternSearch      ( pcPatternKey, piRefY, iRefStride, &cMvSrchRngLT, &cMvSrchRngRB, rcMv, ruiCost ); g_hmmeInMe = 0; g_hmmeMeSeconds += hmmeNow() - t0_; ++g_hmmeMeCalls; }')
once('          xPatternSearchFast  ( pcCU, pcPatternKey, piRefY, iRefStride, &cMvSrchRngLT, &cMvSrchRngRB, rcMv, ruiCost, pIntegerMv2Nx2NPred );',
     '          { const double t0_ = hmmeNow(); g_hmmeInMe = 1; xPatternSearchFast  ( pcCU, pcPatternKey, piRefY, iRefStride, &cMvSrchRngLT, &cMvSrchRngRB, rcMv, ruiCost, pIntegerMv2Nx2NPred ); g_hmmeInMe = 0; g_hmmeMeSeconds += hmmeNow() - t0_; ++g_hmmeMeCalls; }')
# every DistFunc call made while an integer search is being timed: the full-search loop (:3878) and the TZ helper
# (xTZSearchHelp, :390-426); the fractional refinement (:856) runs outside the timed calls and is not counted
n = s.count("m_cDistParam.DistFunc( &m_cDistParam )")
assert n == 5, n
s = s.replace("m_cDistParam.DistFunc( &m_cDistParam )", "(g_hmmeDistCalls += g_hmmeInMe, m_cDistParam.DistFunc( &m_cDistParam ))")
# wall-clock around the fractional-pel refinement that follows every integer search (TEncSearch.cpp:3800)
# (two call sites: xMotionEstimation :3798 and its per-PU twin used after the GPU table look-up :5738)
n = s.count('xPatternSearchFracDIF( bIsLosslessCoded, pcPatternKey, piRefY, iRefStride, &rcMv, cMvHalf, cMvQter, ruiCost ,bBi );')
assert n == 2, n
s = s.replace('xPatternSearchFracDIF( bIsLosslessCoded, pcPatternKey, piRefY, iRefStride, &rcMv, cMvHalf, cMvQter, ruiCost ,bBi );',
              '{ const double t0_ = hmmeNow(); xPatternSearchFracDIF( bIsLosslessCoded, pcPatternKey, piRefY, iRefStride, &rcMv, cMvHalf, cMvQter, ruiCost ,bBi ); g_hmmeFracSeconds += hmmeNow() - t0_; ++g_hmmeFracCalls; g_hmmeFracPixels += (unsigned long long)(iRoiWidth * iRoiHeight); }')
# optional log of every search-window placement handed to the GPU path (golden for hmme_search_window_lt)
once('            m_ppcOpenCLME->calcMotionVectors(piCtu, piRefY, iRefStride, iCtuStride, iSrchRng ,&cMvSrchRngLT);',
     '            if (getenv("HMME_LOG_LT")) { const TComMv& pm_ = bBi ? rcMv : cMvPred; printf("HMME_LT %d %d %d %d %d %d %d %d %d %d %d\\n", (int)pm_.getHor(), (int)pm_.getVer(), iSrchRng, '
     '(int)pcCU->getCUPelX(), (int)pcCU->getCUPelY(), (int)pcCU->getSlice()->getSPS()->getPicWidthInLumaSamples(), (int)pcCU->getSlice()->getSPS()->getPicHeightInLumaSamples(), '
     '(int)cMvSrchRngLT.getHor(), (int)cMvSrchRngLT.getVer(), (int)cMvSrchRngRB.getHor(), (int)cMvSrchRngRB.getVer()); }\n'
     '            m_ppcOpenCLME->calcMotionVectors(piCtu, piRefY, iRefStride, iCtuStride, iSrchRng ,&cMvSrchRngLT);')
# optional binary log of fractional-pel refinements (golden records for the frac oracle): inputs as the
# function sees them (current block, reference patch with the 8-tap apron, integer MV, predictor, lambda) and its outputs
once('#include "TEncSearch.h"', '#include "TEncSearch.h"\n#include <map>\nstatic unsigned g_hmmeFracCosts[18];')
once("""    uiDist += m_pcRdCost->getCost( cMvTest.getHor(), cMvTest.getVer() );

    if ( uiDist < uiDistBest )
    {
      uiDistBest  = uiDist;
      uiDirecBest = i;""", """    uiDist += m_pcRdCost->getCost( cMvTest.getHor(), cMvTest.getVer() );
    g_hmmeFracCosts[i + (iFrac == 2 ? 0 : 9)] = (unsigned)uiDist;

    if ( uiDist < uiDistBest )
    {
      uiDistBest  = uiDist;
      uiDirecBest = i;""")
once("""  ruiCost = xPatternRefinement( pcPatternKey, baseRefMv, 1, rcMvQter, !bIsLosslessCoded );
}""", """  ruiCost = xPatternRefinement( pcPatternKey, baseRefMv, 1, rcMvQter, !bIsLosslessCoded );
  if (const char* logName_ = getenv("HMME_LOG_FRAC"))
  {
    static FILE* f_ = fopen(logName_, "wb");
    static std::map<int, int> seen_;
    static const int cap_ = getenv("HMME_LOG_FRAC_CAP") ? atoi(getenv("HMME_LOG_FRAC_CAP")) : 8;
    static const int stride_ = getenv("HMME_LOG_FRAC_STRIDE") ? atoi(getenv("HMME_LOG_FRAC_STRIDE")) : 29;
    const int w_ = pcPatternKey->getROIYWidth(), h_ = pcPatternKey->getROIYHeight();
    const int n_ = seen_[(w_ * 100 + h_) * 2 + (biPred ? 1 : 0)]++;
    if (f_ && n_ % stride_ == 0 && n_ / stride_ < cap_)
    {
      const int hdr_[16] = { 0x46524143, w_, h_, biPred ? 1 : 0, (m_pcEncCfg->getUseHADME() && !bIsLosslessCoded) ? 1 : 0,
                             pcMvInt->getHor(), pcMvInt->getVer(), m_pcRdCost->m_mvPredictor.getHor(), m_pcRdCost->m_mvPredictor.getVer(),
                             (int)m_pcRdCost->m_uiCost, rcMvHalf.getHor(), rcMvHalf.getVer(), rcMvQter.getHor(), rcMvQter.getVer(), (int)ruiCost, 0 };
      fwrite(hdr_, sizeof(int), 16, f_);
      fwrite(g_hmmeFracCosts, sizeof(unsigned), 18, f_);
      for (int r_ = 0; r_ < h_; ++r_) fwrite(pcPatternKey->getROIY() + r_ * pcPatternKey->getPatternLStride(), sizeof(Pel), w_, f_);
      for (int r_ = -4; r_ < h_ + 4; ++r_) fwrite(piRefY + iOffset + r_ * iRefStride - 4, sizeof(Pel), w_ + 8, f_);
      fflush(f_);
    }
  }
}""")
# optional binary log of AMVP template costs (golden records for the motion-compensated-cost oracle): block, reference
# patch around the integer part of the clipped candidate MV, the MV, and the SAD xGetTemplateCost computed before its rate term
once("""  uiCost = m_pcRdCost->getDistPart( pcCU->getSlice()->getSPS()->getBitDepth(CHANNEL_TYPE_LUMA), pcTemplateCand->getAddr(COMPONENT_Y, uiPartAddr), pcTemplateCand->getStride(COMPONENT_Y), pcOrgYuv->getAddr(COMPONENT_Y, uiPartAddr), pcOrgYuv->getStride(COMPONENT_Y), iSizeX, iSizeY, COMPONENT_Y, DF_SAD );
""", """  uiCost = m_pcRdCost->getDistPart( pcCU->getSlice()->getSPS()->getBitDepth(CHANNEL_TYPE_LUMA), pcTemplateCand->getAddr(COMPONENT_Y, uiPartAddr), pcTemplateCand->getStride(COMPONENT_Y), pcOrgYuv->getAddr(COMPONENT_Y, uiPartAddr), pcOrgYuv->getStride(COMPONENT_Y), iSizeX, iSizeY, COMPONENT_Y, DF_SAD );
  if (const char* logName_ = getenv("HMME_LOG_MC"))
  {
    static FILE* f_ = fopen(logName_, "wb");
    static std::map<int, int> seen_;
    static const int cap_ = getenv("HMME_LOG_FRAC_CAP") ? atoi(getenv("HMME_LOG_FRAC_CAP")) : 8;
    static const int stride_ = getenv("HMME_LOG_FRAC_STRIDE") ? atoi(getenv("HMME_LOG_FRAC_STRIDE")) : 29;
    const int n_ = seen_[(iSizeX * 100 + iSizeY) * 16 + (cMvCand.getHor() & 3) * 4 + (cMvCand.getVer() & 3)]++;
    if (f_ && !(pcCU->getSlice()->testWeightPred() && pcCU->getSlice()->getSliceType()==P_SLICE) && n_ % stride_ == 0 && n_ / stride_ < cap_)
    {
      const int hdr_[8] = { 0x4d434f53, iSizeX, iSizeY, cMvCand.getHor(), cMvCand.getVer(), (int)uiCost, 0, 0 };
      fwrite(hdr_, sizeof(int), 8, f_);
      Pel* org_ = pcOrgYuv->getAddr(COMPONENT_Y, uiPartAddr);
      for (int r_ = 0; r_ < iSizeY; ++r_) fwrite(org_ + r_ * pcOrgYuv->getStride(COMPONENT_Y), sizeof(Pel), iSizeX, f_);
      const Int rs_ = pcPicYuvRef->getStride(COMPONENT_Y);
      Pel* ref_ = pcPicYuvRef->getAddr(COMPONENT_Y, pcCU->getCtuRsAddr(), pcCU->getZorderIdxInCtu() + uiPartAddr) + (cMvCand.getVer() >> 2) * rs_ + (cMvCand.getHor() >> 2);
      for (int r_ = -4; r_ < iSizeY + 4; ++r_) fwrite(ref_ + r_ * rs_ - 4, sizeof(Pel), iSizeX + 8, f_);
      fflush(f_);
    }
  }
""")
# optional binary log of inter-prediction errors (merge candidates and the motion-estimation result, uni- and bi-directional; golden
# records for the bi-directional oracle): what xGetInterPredictionError computed and the samples it computed it from
once("""  ruiErr = cDistParam.DistFunc( &cDistParam );
}""", """  ruiErr = cDistParam.DistFunc( &cDistParam );
  if (const char* logName_ = getenv("HMME_LOG_IPE"))
  {
    static FILE* f_ = fopen(logName_, "wb");
    static std::map<int, int> seen_;
    static const int cap_ = getenv("HMME_LOG_IPE_CAP") ? atoi(getenv("HMME_LOG_IPE_CAP")) : 4;
    static const int stride_ = getenv("HMME_LOG_IPE_STRIDE") ? atoi(getenv("HMME_LOG_IPE_STRIDE")) : 5;
    const Int ri0_ = pcCU->getCUMvField(REF_PIC_LIST_0)->getRefIdx(uiAbsPartIdx), ri1_ = pcCU->getCUMvField(REF_PIC_LIST_1)->getRefIdx(uiAbsPartIdx);
    const bool bi_ = ri0_ >= 0 && ri1_ >= 0 && !xCheckIdenticalMotion(pcCU, uiAbsPartIdx);
    const bool wp_ = (pcCU->getSlice()->getPPS()->getUseWP() && pcCU->getSlice()->getSliceType() == P_SLICE) || (pcCU->getSlice()->getPPS()->getWPBiPred() && pcCU->getSlice()->getSliceType() == B_SLICE);
    const int n_ = seen_[(iWidth * 100 + iHeight) * 2 + (bi_ ? 1 : 0)]++;
    if (f_ && !wp_ && n_ % stride_ == 0 && n_ / stride_ < cap_)
    {
      int lists_[2], nl_ = 0;
      if (bi_) { lists_[0] = 0; lists_[1] = 1; nl_ = 2; } else { lists_[0] = ri0_ >= 0 ? 0 : 1; nl_ = 1; }
      TComMv mv_[2]; Pel* rp_[2] = {0, 0}; Int rs_[2] = {0, 0};
      for (int l_ = 0; l_ < nl_; ++l_)
      {
        const RefPicList e_ = lists_[l_] ? REF_PIC_LIST_1 : REF_PIC_LIST_0;
        mv_[l_] = pcCU->getCUMvField(e_)->getMv(uiAbsPartIdx);
        pcCU->clipMv(mv_[l_]);
        TComPicYuv* py_ = pcCU->getSlice()->getRefPic(e_, pcCU->getCUMvField(e_)->getRefIdx(uiAbsPartIdx))->getPicYuvRec();
        rs_[l_] = py_->getStride(COMPONENT_Y);
        rp_[l_] = py_->getAddr(COMPONENT_Y, pcCU->getCtuRsAddr(), pcCU->getZorderIdxInCtu() + uiAbsPartIdx) + (mv_[l_].getVer() >> 2) * rs_[l_] + (mv_[l_].getHor() >> 2);
      }
      const int hdr_[12] = { 0x49504552, iWidth, iHeight, nl_, (m_pcEncCfg->getUseHADME() && (pcCU->getCUTransquantBypass(iPartIdx) == 0)) ? 1 : 0,
                             mv_[0].getHor(), mv_[0].getVer(), nl_ == 2 ? mv_[1].getHor() : 0, nl_ == 2 ? mv_[1].getVer() : 0, (int)ruiErr, 0, 0 };
      fwrite(hdr_, sizeof(int), 12, f_);
      Pel* org_ = pcYuvOrg->getAddr(COMPONENT_Y, uiAbsPartIdx);
      for (int r_ = 0; r_ < iHeight; ++r_) fwrite(org_ + r_ * pcYuvOrg->getStride(COMPONENT_Y), sizeof(Pel), iWidth, f_);
      for (int l_ = 0; l_ < nl_; ++l_)
        for (int r_ = -4; r_ < iHeight + 4; ++r_) fwrite(rp_[l_] + r_ * rs_[l_] - 4, sizeof(Pel), iWidth + 8, f_);
      fflush(f_);
    }
  }
}""")
open(path, "w").write(s)
# the log reads the predictor and lambda straight from TComRdCost: open its first private section in the scratch copy
import os
rd = os.path.join(os.path.dirname(path), "..", "TLibCommon", "TComRdCost.h")
t = open(rd).read()
assert t.count("private:\n  // for distortion") == 1
open(rd, "w").write(t.replace("private:\n  // for distortion", "public:\n  // for distortion"))
