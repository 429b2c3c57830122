#!/usr/bin/env python
"""Instruments a SCRATCH COPY of the reference's TEncSearch.cpp for the CPU-ME baseline:
wall-clock around the OpenCL=0 integer search calls of xMotionEstimation (TEncSearch.cpp:3774-3791: full search and TZ)
and a counter of the DistFunc invocations made inside them (:390-426, :3878).  Arithmetic and control flow are untouched; the bitstream stays
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
static void hmmeReport() { printf("HMME_CPUME me_seconds=%.6f dist_calls=%llu me_calls=%llu\\n", g_hmmeMeSeconds, g_hmmeDistCalls, g_hmmeMeCalls); }
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
# optional log of every search-window placement handed to the GPU path (golden for hmme_search_window_lt)
once('            m_ppcOpenCLME->calcMotionVectors(piCtu, piRefY, iRefStride, iCtuStride, iSrchRng ,&cMvSrchRngLT);',
     '            if (getenv("HMME_LOG_LT")) { const TComMv& pm_ = bBi ? rcMv : cMvPred; printf("HMME_LT %d %d %d %d %d %d %d %d %d %d %d\\n", (int)pm_.getHor(), (int)pm_.getVer(), iSrchRng, '
     '(int)pcCU->getCUPelX(), (int)pcCU->getCUPelY(), (int)pcCU->getSlice()->getSPS()->getPicWidthInLumaSamples(), (int)pcCU->getSlice()->getSPS()->getPicHeightInLumaSamples(), '
     '(int)cMvSrchRngLT.getHor(), (int)cMvSrchRngLT.getVer(), (int)cMvSrchRngRB.getHor(), (int)cMvSrchRngRB.getVer()); }\n'
     '            m_ppcOpenCLME->calcMotionVectors(piCtu, piRefY, iRefStride, iCtuStride, iSrchRng ,&cMvSrchRngLT);')
open(path, "w").write(s)
