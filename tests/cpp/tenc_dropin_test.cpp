// tenc_dropin_test.cpp -- exercises the C++ drop-in class (hm-opencl_b200/host/TEncOpenCL) exactly the way HM does
// (TEncTop::xInitOpenCL call order, TEncSearch::xMotionEstimation per-CTU call + getters) and checks every output
// against the CPU oracle.  Test code: it may link the oracle.  Built and run by tests/test_gpu_dropin_cpp.py.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "TEncOpenCL.h"
#include "hmme_b200.h"
#include "hmme_oracle.h"

static unsigned lcg(unsigned& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

int main() {
    TEncOpenCL me;                                                       // TEncTop.h:82: a by-value member, constructed in every run
    if (!me.findDevice(0)) { printf("FAIL findDevice\n"); return 1; }
    if (me.compileKernelSource(NULL, "calcSAD_AMP")) { printf("FAIL NULL kernel file accepted\n"); return 1; }
    if (!me.compileKernelSource("cl/sad.cl", "calcSAD_AMP")) { printf("FAIL compileKernelSource\n"); return 1; }
    if (!me.createBuffers(64, 64, 64)) { printf("FAIL createBuffers: %s\n", me.getLastError()); return 1; }
    me.setEnabled(true);
    printf("device: %s\n", me.getDeviceInfo());

    const int W = 256, H = 192, M = 80, S = W + 2 * M;                   // padded plane like TComPicYuv (margin 80)
    std::vector<Pel> ref((size_t)S * (H + 2 * M)), cur((size_t)S * (H + 2 * M));
    unsigned seed = 12345;
    for (size_t i = 0; i < ref.size(); ++i) ref[i] = (Pel)(lcg(seed) & 255);
    for (int y = 0; y < H + 2 * M; ++y)
        for (int x = 0; x < S; ++x) {
            const int sy = y + 2 < H + 2 * M ? y + 2 : y, sx = x + 3 < S ? x + 3 : x;
            cur[(size_t)y * S + x] = (Pel)((ref[(size_t)sy * S + sx] + (int)(lcg(seed) % 5) - 2) & 255);
        }
    int bad = 0, calls = 0;
    const double lambdas[3] = {49.3, 4.0, 1200.0};
    const int ranges[3] = {64, 16, 4};
    for (int t = 0; t < 3; ++t) {
        me.setLambda(lambdas[t]);                                        // TEncSlice::setUpLambda
        const int R = ranges[t];
        for (int cy = 0; cy + 64 <= H; cy += 64)
            for (int cx = 0; cx + 64 <= W; cx += 64) {
                Pel blk[64 * 64];
                for (int r = 0; r < 64; ++r) memcpy(blk + 64 * r, &cur[(size_t)(M + cy + r) * S + M + cx], 64 * sizeof(Pel));
                if (t == 2)                                              // bi-pred style 16-bit block 2*org - pred
                    for (int i = 0; i < 4096; ++i) blk[i] = (Pel)(2 * blk[i] - (Pel)(lcg(seed) & 255));
                TComMv lt((Short)(-R + (cx ? 3 : 0)), (Short)(-R - (cy ? 2 : 0)));
                Pel* refAtCtu = &ref[(size_t)(M + cy) * S + M + cx];
                me.calcMotionVectors(blk, refAtCtu, S, 64, R, &lt);
                int32_t X[593], Y[593]; uint32_t sad[593], cost[593];
                hmme_oracle_search_ctu(blk, 64, refAtCtu, S, R, lt.getHor(), lt.getVer(), hmme_oracle_lambda_q16(lambdas[t]), X, Y, sad, cost);
                ++calls;
                for (int p = 0; p < 593; ++p)
                    if (me.getX()[p] != X[p] || me.getY()[p] != Y[p] || me.getRuiCost()[p] != sad[p]) {
                        if (bad++ < 5) printf("MISMATCH R=%d ctu(%d,%d) part %d: got (%d,%d,%u) want (%d,%d,%u)\n", R, cx, cy, p, me.getX()[p], me.getY()[p],
                                              me.getRuiCost()[p], X[p], Y[p], sad[p]);
                    }
            }
    }
    // the fractional-pel refinement as TEncSearch::xPatternSearchFracDIF would call it:
    // pattern key block, piRefY at the PU origin, integer MV, predictor; a few PU shapes, 8-bit and 16-bit blocks
    int fcalls = 0;
    me.setLambda(49.3);
    const int shapes[6][2] = {{64, 64}, {32, 24}, {16, 4}, {12, 16}, {8, 8}, {48, 64}};
    for (int t = 0; t < 12; ++t) {
        const int w = shapes[t % 6][0], h = shapes[t % 6][1], px = 16 + 8 * t, py = 24 + 4 * t;
        Pel blk[64 * 64];
        for (int r = 0; r < h; ++r)
            for (int c = 0; c < w; ++c) {
                const Pel v = cur[(size_t)(M + py + r) * S + M + px + c];
                blk[r * 64 + c] = t >= 6 ? (Pel)(2 * v - (Pel)(lcg(seed) & 255)) : v;
            }
        const TComMv mvInt((Short)(3 - t), (Short)(t - 5)), pred((Short)(7 * t - 30), (Short)(11 - 3 * t));
        Pel* refAtPu = &ref[(size_t)(M + py) * S + M + px];
        TComMv half, qter;
        const Distortion cost = me.refineFractional(blk, 64, w, h, refAtPu, S, mvInt, pred, (t & 1) == 0, half, qter);
        const hmme_oracle_pu pu = {0, 0, w, h, mvInt.getHor(), mvInt.getVer(), pred.getHor(), pred.getVer()};
        int32_t mvq[2], oh[2], oq[2]; uint32_t oc, od;
        hmme_oracle_refine_frac(blk, 64, refAtPu, S, &pu, 1, hmme_oracle_lambda_q16(49.3), (t & 1) == 0, mvq, oh, oq, &oc, &od, NULL);
        ++fcalls;
        const int gx = 4 * mvInt.getHor() + 2 * half.getHor() + qter.getHor(), gy = 4 * mvInt.getVer() + 2 * half.getVer() + qter.getVer();
        if (gx != mvq[0] || gy != mvq[1] || cost != oc || half.getHor() < -1 || half.getHor() > 1 || qter.getVer() < -1 || qter.getVer() > 1) {
            if (bad++ < 10) printf("MISMATCH refineFractional %dx%d: got (%d,%d,%u) want (%d,%d,%u)\n", w, h, gx, gy, cost, mvq[0], mvq[1], oc);
        }
    }
    // the AMVP candidate check's distortion as TEncSearch::xGetTemplateCost would ask for it
    int tcalls = 0;
    for (int t = 0; t < 12; ++t) {
        const int w = shapes[t % 6][0], h = shapes[t % 6][1], px = 24 + 8 * t, py = 20 + 4 * t;
        Pel blk[64 * 64];
        for (int r = 0; r < h; ++r) memcpy(blk + 64 * r, &cur[(size_t)(M + py + r) * S + M + px], w * sizeof(Pel));
        const TComMv mv((Short)(13 * t - 70), (Short)(41 - 9 * t));
        Pel* refAtPu = &ref[(size_t)(M + py) * S + M + px];
        const Distortion d = me.templateDistortion(blk, 64, w, h, refAtPu, S, mv, (t & 1) != 0);
        const hmme_oracle_mc_pu pu = {0, 0, w, h, mv.getHor(), mv.getVer()};
        uint32_t od;
        hmme_oracle_mc_cost(blk, 64, refAtPu, S, &pu, 1, (t & 1) != 0, &od);
        ++tcalls;
        if (d != od && bad++ < 10) printf("MISMATCH templateDistortion %dx%d: got %u want %u\n", w, h, d, od);
    }
    // merge candidates / ME result as TEncSearch::xGetInterPredictionError would ask: bi-directional over two reference planes (cur doubles as
    // the second one), and the uni-directional form through the same method
    int icalls = 0;
    for (int t = 0; t < 12; ++t) {
        const int w = shapes[t % 6][0], h = shapes[t % 6][1], px = 32 + 8 * t, py = 16 + 4 * t;
        Pel blk[64 * 64];
        for (int r = 0; r < h; ++r) memcpy(blk + 64 * r, &ref[(size_t)(M + py + r + 1) * S + M + px + 1], w * sizeof(Pel));
        const TComMv mv0((Short)(9 * t - 40), (Short)(23 - 5 * t)), mv1((Short)(31 - 7 * t), (Short)(3 * t - 14));
        Pel* r0 = &ref[(size_t)(M + py) * S + M + px];
        Pel* r1 = &cur[(size_t)(M + py) * S + M + px];
        const bool bi = (t % 3) != 0, had = (t & 1) == 0;
        const Distortion d = me.interPredictionError(blk, 64, w, h, r0, S, mv0, bi ? r1 : NULL, S, mv1, had);
        uint32_t od;
        if (bi) {
            const hmme_oracle_mc_bi_pu pu = {0, 0, w, h, mv0.getHor(), mv0.getVer(), mv1.getHor(), mv1.getVer()};
            hmme_oracle_mc_cost_bi(blk, 64, r0, S, r1, S, &pu, 1, had, &od);
        } else {
            const hmme_oracle_mc_pu pu = {0, 0, w, h, mv0.getHor(), mv0.getVer()};
            hmme_oracle_mc_cost(blk, 64, r0, S, &pu, 1, had, &od);
        }
        ++icalls;
        if (d != od && bad++ < 10) printf("MISMATCH interPredictionError %dx%d bi=%d: got %u want %u\n", w, h, (int)bi, d, od);
    }
    // the speculative whole-frame search as the patched TEncSlice::compressSlice drives it: announce the picture,
    // then one calcMotionVectors call per CTU in coding order.  Pass 0: windows of a zero predictor -> every call must be a table hit.
    // Pass 1: the predictor changes at CTU 5 -> one miss (synchronous search + re-speculation), hits again afterwards.  Pass 2: a
    // bi-prediction style block (not the original picture's) -> never served from the table.  Every answer is checked against the oracle.
    int scalls = 0;
    {
        const int R = 16;
        me.setLambda(49.3);
        const TEncOpenCL::SpecStats s0 = me.getSpecStats();
        for (int pass = 0; pass < 3; ++pass) {
            me.beginPicture(&cur[(size_t)M * S + M], S, W, H);
            me.addReferencePicture(&ref[(size_t)M * S + M], S, M, M);
            me.addReferencePicture(&ref[(size_t)M * S + M], S, M, M);   // the same picture in the other list: announced once
            me.speculate(R);
            int c = 0;
            for (int cy = 0; cy + 64 <= H; cy += 64)
                for (int cx = 0; cx + 64 <= W; cx += 64, ++c) {
                    const int px = (pass == 1 && c >= 5) ? 24 : 0, py = (pass == 1 && c >= 5) ? -12 : 0;
                    int ltx, lty;
                    hmme_search_window(px, py, R, cx, cy, W, H, &ltx, &lty, NULL, NULL);
                    TComMv lt((Short)ltx, (Short)lty);
                    Pel blk[64 * 64];
                    for (int r = 0; r < 64; ++r) memcpy(blk + 64 * r, &cur[(size_t)(M + cy + r) * S + M + cx], 64 * sizeof(Pel));
                    if (pass == 2) blk[100] = (Pel)(blk[100] + 300);
                    Pel* refAtCtu = &ref[(size_t)(M + cy) * S + M + cx];
                    me.calcMotionVectors(blk, refAtCtu, S, 64, R, &lt);
                    int32_t X[593], Y[593]; uint32_t sad[593], cost[593];
                    hmme_oracle_search_ctu(blk, 64, refAtCtu, S, R, ltx, lty, hmme_oracle_lambda_q16(49.3), X, Y, sad, cost);
                    ++scalls;
                    for (int p = 0; p < 593; ++p)
                        if (me.getX()[p] != X[p] || me.getY()[p] != Y[p] || me.getRuiCost()[p] != sad[p]) {
                            if (bad++ < 5) printf("MISMATCH speculative pass %d ctu %d part %d\n", pass, c, p);
                        }
                }
            me.endPicture();
            const TEncOpenCL::SpecStats st = me.getSpecStats();
            const int n = c;
            const unsigned long long calls = st.calls - s0.calls, hits = st.hits - s0.hits;
            const unsigned long long wantHits = pass == 0 ? n : pass == 1 ? 2 * n - 1 : 2 * n - 1;
            if (calls != (unsigned long long)(pass + 1) * n || hits != wantHits) {
                printf("FAIL speculative pass %d: calls %llu hits %llu (want %llu), miss_block %llu miss_window %llu\n", pass, calls, hits, wantHits,
                       st.missBlock - s0.missBlock, st.missWindow - s0.missWindow);
                ++bad;
            }
        }
    }
    printf("%d speculative calls; ", scalls);
    printf("%s: %d calcMotionVectors calls, %d refineFractional, %d templateDistortion, %d interPredictionError calls, %d mismatches\n",
           bad ? "FAIL" : "PASS", calls, fcalls, tcalls, icalls, bad);
    return bad ? 1 : 0;
}
