#!/usr/bin/env python
"""Routes the reference encoder's fractional-pel refinement through the B200 library in a SCRATCH COPY of the reference sources
(bitstream-parity build `_ref/TAppEncoder_b200frac`): the body of TEncSearch::xPatternSearchFracDIF
(TEncSearch.cpp:4294-4331) becomes one call of TEncOpenCL::refineFractional when GPU ME is enabled, and TComRdCost gets a getter
for the predictor the call needs; TEncSearch::xGetTemplateCost (:3634-3674, the AMVP candidate check) takes its SAD from
TEncOpenCL::templateDistortion.  This is the integration a maintainer would write; nothing
patched is ever committed."""
import os
import sys

path = sys.argv[1]                       # .../TLibEncoder/TEncSearch.cpp of the scratch copy
s = open(path).read()
old = """  //  Reference pattern initialization (integer scale)
  TComPattern cPatternRoi;
  Int         iOffset    = pcMvInt->getHor() + pcMvInt->getVer() * iRefStride;"""
assert s.count(old) == 1
s = s.replace(old, """  if ( m_ppcOpenCLME && m_ppcOpenCLME->isEnabled() )
  {
    TComMv cPred = m_pcRdCost->getPredictor();
    ruiCost = m_ppcOpenCLME->refineFractional( pcPatternKey->getROIY(), pcPatternKey->getPatternLStride(), pcPatternKey->getROIYWidth(),
                                               pcPatternKey->getROIYHeight(), piRefY, iRefStride, *pcMvInt, cPred,
                                               m_pcEncCfg->getUseHADME() && !bIsLosslessCoded, rcMvHalf, rcMvQter );
    m_pcRdCost->setCostScale( 0 );          // the state the CPU body leaves behind
    return;
  }
""" + old)
# row f3: the AMVP candidate check (xGetTemplateCost, TEncSearch.cpp:3634-3674) takes its SAD from the library as well
old = """  // prediction pattern
  if ( pcCU->getSlice()->testWeightPred() && pcCU->getSlice()->getSliceType()==P_SLICE )
  {
    xPredInterBlk( COMPONENT_Y, pcCU, pcPicYuvRef, uiPartAddr, &cMvCand, iSizeX, iSizeY, pcTemplateCand, true,"""
assert s.count(old) == 1
s = s.replace(old, """  if ( m_ppcOpenCLME && m_ppcOpenCLME->isEnabled() && !( pcCU->getSlice()->testWeightPred() && pcCU->getSlice()->getSliceType()==P_SLICE ) )
  {
    uiCost = m_ppcOpenCLME->templateDistortion( pcOrgYuv->getAddr(COMPONENT_Y, uiPartAddr), pcOrgYuv->getStride(COMPONENT_Y), iSizeX, iSizeY,
                                                pcPicYuvRef->getAddr(COMPONENT_Y, pcCU->getCtuRsAddr(), pcCU->getZorderIdxInCtu() + uiPartAddr),
                                                pcPicYuvRef->getStride(COMPONENT_Y), cMvCand, false );
    return (UInt) m_pcRdCost->calcRdCost( m_auiMVPIdxCost[iMVPIdx][iMVPNum], uiCost, false, DF_SAD );
  }
""" + old)
# row f3, second half: xGetInterPredictionError (TEncSearch.cpp:2814-2836; merge candidates and the motion-estimation result) takes the
# luma distortion of the uni- or bi-directional prediction from the library; motionCompensation is skipped (its output is only used here)
old = """  motionCompensation( pcCU, &m_tmpYuvPred, REF_PIC_LIST_X, iPartIdx );

  UInt uiAbsPartIdx = 0;
  Int iWidth = 0;
  Int iHeight = 0;
  pcCU->getPartIndexAndSize( iPartIdx, uiAbsPartIdx, iWidth, iHeight );
"""
assert s.count(old) == 1
s = s.replace(old, """  const Bool bWp_ = ( pcCU->getSlice()->getPPS()->getUseWP() && pcCU->getSlice()->getSliceType() == P_SLICE ) ||
                    ( pcCU->getSlice()->getPPS()->getWPBiPred() && pcCU->getSlice()->getSliceType() == B_SLICE );
  if ( m_ppcOpenCLME && m_ppcOpenCLME->isEnabled() && !bWp_ )
  {
    UInt uiAddr_ = 0; Int iW_ = 0, iH_ = 0;
    pcCU->getPartIndexAndSize( iPartIdx, uiAddr_, iW_, iH_ );
    const Int iRef0_ = pcCU->getCUMvField(REF_PIC_LIST_0)->getRefIdx(uiAddr_), iRef1_ = pcCU->getCUMvField(REF_PIC_LIST_1)->getRefIdx(uiAddr_);
    const Bool bBi_ = iRef0_ >= 0 && iRef1_ >= 0 && !xCheckIdenticalMotion( pcCU, uiAddr_ );
    Pel* apRef_[2] = { NULL, NULL }; Int aiStride_[2] = { 0, 0 }; TComMv acMv_[2];
    Int n_ = 0;
    for ( Int l_ = 0; l_ < 2; l_++ )
    {
      const RefPicList e_ = l_ ? REF_PIC_LIST_1 : REF_PIC_LIST_0;
      const Int iRef_ = l_ ? iRef1_ : iRef0_;
      if ( iRef_ < 0 || ( !bBi_ && n_ == 1 ) ) continue;
      acMv_[n_] = pcCU->getCUMvField(e_)->getMv(uiAddr_);
      pcCU->clipMv( acMv_[n_] );
      TComPicYuv* pcRec_ = pcCU->getSlice()->getRefPic( e_, iRef_ )->getPicYuvRec();
      aiStride_[n_] = pcRec_->getStride(COMPONENT_Y);
      apRef_[n_] = pcRec_->getAddr( COMPONENT_Y, pcCU->getCtuRsAddr(), pcCU->getZorderIdxInCtu() + uiAddr_ );
      n_++;
    }
    ruiErr = m_ppcOpenCLME->interPredictionError( pcYuvOrg->getAddr( COMPONENT_Y, uiAddr_ ), pcYuvOrg->getStride(COMPONENT_Y), iW_, iH_,
                                                  apRef_[0], aiStride_[0], acMv_[0], bBi_ ? apRef_[1] : NULL, aiStride_[1], acMv_[1],
                                                  m_pcEncCfg->getUseHADME() && (pcCU->getCUTransquantBypass(iPartIdx) == 0) );
    return;
  }
""" + old)
open(path, "w").write(s)

rd = os.path.join(os.path.dirname(path), "..", "TLibCommon", "TComRdCost.h")
t = open(rd).read()
old = "  Void    setCostScale( Int iCostScale )    { m_iCostScale = iCostScale; }"
assert t.count(old) == 1
open(rd, "w").write(t.replace(old, old + "\n  TComMv  getPredictor() const              { return m_mvPredictor; }"))
