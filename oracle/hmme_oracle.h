/*
 * hmme_oracle.h -- CPU ORACLE for HM-OpenCL's whole-CTU integer-pel motion estimation.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product path: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library, and only as the checker / the CPU arm.  The product (hm-opencl_b200/) never links,
 * imports or calls it and has no CPU fallback.
 *
 * It is a plain-C restatement of the reference's algorithm (never a copy of its code):
 *   - 4x4 base SADs ............ /root/reference/cl/sad.cl:171-186      (calcSAD_AMP, abs_diff on short)
 *   - 593 partition SADs ....... /root/reference/cl/sad.cl:188-365      (intended = true rectangle sums)
 *   - cost / strict-< arg-min .. /root/reference/cl/sad.cl:370-408      (compareSAD)
 *   - window origin, scan order  /root/reference/source/Lib/TLibEncoder/TEncOpenCL.cpp:243-256,312-333
 *   - initial values ........... /root/reference/source/Lib/TLibEncoder/TEncOpenCL.cpp:366-392
 *   - lambda quantisation ...... /root/reference/source/Lib/TLibEncoder/TEncOpenCL.h:121
 *   - partition index layout ... /root/reference/source/Lib/TLibCommon/TComDataCU.cpp:4676-6461
 *   - NUM_CTU_PARTS / Pel ...... /root/reference/source/Lib/TLibCommon/TypeDef.h:257-266,706,717
 *
 * Parity pinning: the reference ships no golden vectors for this path, so the
 * oracle is pinned against the reference ITSELF run here: oracle/_ref/ holds the reference's own
 * cl/sad.cl + TEncOpenCL.cpp compiled for the CPU behind a lock-step OpenCL emulation
 * (oracle/Makefile, oracle/refemu/), tests/golden/ holds vectors generated from it
 * (oracle/gen_golden.py) and the 593-entry layout parsed from getIndexBlock.
 */
#ifndef HMME_ORACLE_H
#define HMME_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HMME_ORACLE_NUM_PARTS 593

typedef struct { int x, y, w, h; } hmme_oracle_rect;

/* Rectangle (luma offset inside the 64x64 CTU, size) of partition index p (0..592). */
void hmme_oracle_partition_table(hmme_oracle_rect out[HMME_ORACLE_NUM_PARTS]);

/* bits(v) of sad.cl:377-396 == TComRdCost::xGetComponentBits (TComRdCost.cpp:278-292). */
uint32_t hmme_oracle_mv_bits(int v);

/* floor(65536*sqrt(lambda)) as UInt, TEncOpenCL.h:121. */
uint32_t hmme_oracle_lambda_q16(double lambda);

/*
 * One calcMotionVectors call (TEncOpenCL.cpp:240-362), scalar and obviously-correct.
 *   cur       : 64x64 block, int16, row stride curStride (elements)
 *   refAtCtu  : pointer INTO the padded reference plane at the CTU's top-left sample
 *   refStride : plane stride (elements); addressing is linear, ref[(T+y+r)*S + L+x+c]
 *   range     : R; candidates x,y in 0..2R inclusive, y outer, x inner
 *   ltx,lty   : integer-pel search-range left/top (pcMvSrchRngLT)
 *   lambda    : m_lambda (uint32)
 * Outputs (593 each): X,Y = winning integer MV (x+ltx, y+lty), sad = pure SAD at the winner
 * (ruiCosts), cost = minSad.  A partition never updated keeps X=Y=0, sad=0, cost=0xFFFFFFFF.
 */
int hmme_oracle_search_ctu(const int16_t* cur, int curStride,
                           const int16_t* refAtCtu, int refStride,
                           int range, int ltx, int lty, uint32_t lambda,
                           int32_t* X, int32_t* Y, uint32_t* sad, uint32_t* cost);

/*
 * Frame batch used as the CPU arm of bench.py: njobs independent CTU searches over one padded
 * 8-bit-content plane pair, spread over nthreads POSIX threads (jobs are independent).
 *   curPlane / refPlane : int16 planes; origin = sample (0,0) of the picture; stride in elements
 *   jobs                : njobs x {ctuX, ctuY, ltx, lty} (pixels; lt relative to the CTU origin)
 * Outputs are [njobs][593].  Same arithmetic as hmme_oracle_search_ctu (hierarchical sums instead
 * of an integral image; checked equal in tests/test_oracle.py).
 */
int hmme_oracle_search_frame(const int16_t* curOrigin, int curStride,
                             const int16_t* refOrigin, int refStride,
                             const int32_t* jobs, int njobs, int range, uint32_t lambda,
                             int nthreads,
                             int32_t* X, int32_t* Y, uint32_t* sad, uint32_t* cost);

/*
 * Fractional-pel refinement after the integer search (hmme_frac_oracle.c has the citations):
 * TEncSearch::xPatternSearchFracDIF for a list of prediction units.  Planes are int16 with origin = picture sample (0,0)
 * (cur may hold the 16-bit bi-prediction target 2*org - pred, ref has 8-bit content); each PU gives its rectangle, the
 * integer-pel MV found by the search and the quarter-pel predictor the bit cost is relative to.  lambda is
 * TComRdCost::m_uiCost (= m_uiLambdaMotionSAD = floor(65536*sqrt(lambda))); useHad = HadamardME && !lossless.
 * Outputs per PU: mvq = final quarter-pel MV, half / qter = the two stage winners (each -1..1), cost = ruiCost as
 * xPatternSearchFracDIF returns it, dist = cost minus the MV cost of the winner.
 */
typedef struct { int32_t x, y, w, h, mvx, mvy, predx, predy; } hmme_oracle_pu;
int hmme_oracle_refine_frac(const int16_t* curOrigin, int curStride, const int16_t* refOrigin, int refStride,
                            const hmme_oracle_pu* pus, int npus, uint32_t lambda, int useHad,
                            int32_t* mvq, int32_t* half, int32_t* qter, uint32_t* cost, uint32_t* dist,
                            uint32_t* cand /* optional [npus][18]: cost of each half- then quarter-pel candidate */);

/* Distortion (SAD, or Hadamard SATD with useHad) between the current block and the motion-compensated uni-prediction at a
 * QUARTER-PEL MV: the arithmetic of xGetTemplateCost / the uni-directional merge candidates (hmme_frac_oracle.c has the citations). */
typedef struct { int32_t x, y, w, h, mvqx, mvqy; } hmme_oracle_mc_pu;
int hmme_oracle_mc_cost(const int16_t* curOrigin, int curStride, const int16_t* refOrigin, int refStride,
                        const hmme_oracle_mc_pu* pus, int npus, int useHad, uint32_t* dist);

/* Bi-directional form: two reference planes, two clipped quarter-pel MVs (xPredInterBi + addAvg). */
typedef struct { int32_t x, y, w, h, mv0x, mv0y, mv1x, mv1y; } hmme_oracle_mc_bi_pu;
int hmme_oracle_mc_cost_bi(const int16_t* curOrigin, int curStride, const int16_t* ref0Origin, int ref0Stride,
                           const int16_t* ref1Origin, int ref1Stride, const hmme_oracle_mc_bi_pu* pus, int npus, int useHad, uint32_t* dist);

#ifdef __cplusplus
}
#endif
#endif
