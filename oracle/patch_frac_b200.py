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
open(path, "w").write(s)

rd = os.path.join(os.path.dirname(path), "..", "TLibCommon", "TComRdCost.h")
t = open(rd).read()
old = "  Void    setCostScale( Int iCostScale )    { m_iCostScale = iCostScale; }"
assert t.count(old) == 1
open(rd, "w").write(t.replace(old, old + "\n  TComMv  getPredictor() const              { return m_mvPredictor; }"))
