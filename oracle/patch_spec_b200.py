#!/usr/bin/env python
"""Patches a SCRATCH COPY of the reference's TEncSlice.cpp for the `TAppEncoder_b200spec` parity build (SURVEY.md section 8 row f2,
INTEGRATION.md section 4.2): before the CTU loop of TEncSlice::compressSlice (/root/reference/source/Lib/TLibEncoder/TEncSlice.cpp:730)
the slice encoder announces the original picture and the slice's reference pictures to the drop-in TEncOpenCL, which searches every
CTU of the picture speculatively.  TEncSearch::xMotionEstimation and its calcMotionVectors call are left untouched.

    python patch_spec_b200.py <scratch>/source/Lib/TLibEncoder/TEncSlice.cpp

TEST / INTEGRATION INFRASTRUCTURE: writes only to the path given (never under /root/reference)."""
import sys

ANCHOR = "  // for every CTU in the slice segment (may terminate sooner if there is a byte limit on the slice-segment)\n"
HOOK = """  // hm-opencl_b200: announce the picture to the GPU motion estimator (speculative whole-frame search, INTEGRATION.md section 4.2)
  if ( pcSlice->getSliceType() != I_SLICE && m_pcOpenCLME->isEnabled() )
  {
    TComPicYuv* pcOrgYuv = pcPic->getPicYuvOrg();
    m_pcOpenCLME->beginPicture( pcOrgYuv->getAddr(COMPONENT_Y), pcOrgYuv->getStride(COMPONENT_Y), pcOrgYuv->getWidth(COMPONENT_Y), pcOrgYuv->getHeight(COMPONENT_Y) );
    for ( Int iList = 0; iList < ( pcSlice->isInterB() ? 2 : 1 ); iList++ )
    {
      for ( Int iRef = 0; iRef < pcSlice->getNumRefIdx( RefPicList(iList) ); iRef++ )
      {
        TComPicYuv* pcRecYuv = pcSlice->getRefPic( RefPicList(iList), iRef )->getPicYuvRec();
        m_pcOpenCLME->addReferencePicture( pcRecYuv->getAddr(COMPONENT_Y), pcRecYuv->getStride(COMPONENT_Y), pcRecYuv->getMarginX(COMPONENT_Y), pcRecYuv->getMarginY(COMPONENT_Y) );
      }
    }
    m_pcOpenCLME->speculate( m_pcCfg->getSearchRange() );
  }

"""


def main():
    path = sys.argv[1]
    assert not path.startswith("/root/reference"), "patch a scratch copy, never the reference tree"
    s = open(path).read()
    assert s.count(ANCHOR) == 1, "anchor not found exactly once in " + path
    s = s.replace(ANCHOR, HOOK + ANCHOR)
    open(path, "w").write(s)


if __name__ == "__main__":
    main()
