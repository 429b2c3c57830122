// group_test.cpp -- a C++ integrator's view of the multi-GPU entry points (include/hmme_b200.h, hmme_group_*): one process,
// every visible GPU (at most 4), one frame cut into bands by the library, results in one host table, both ways of distributing
// the reference picture.  Every CTU is checked against the CPU oracle.  Test code: it may link the oracle.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "hmme_b200.h"
#include "hmme_oracle.h"

static unsigned lcg(unsigned& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

int main() {
    int ndev = 0;
    if (hmme_device_count(&ndev) != HMME_OK || ndev < 1) { printf("FAIL no device: %s\n", hmme_last_error(NULL)); return 1; }
    if (ndev > 4) ndev = 4;
    int devs[4] = {0, 1, 2, 3};
    hmme_group* g = NULL;
    if (hmme_group_create(&g, devs, ndev, 32) != HMME_OK) { printf("FAIL hmme_group_create: %s\n", hmme_group_last_error(NULL)); return 1; }
    const int W = 704, H = 448, M = 80, R = 32, S = W + 2 * M, rows = H + 2 * M;   // 11 x 7 CTUs: bands cut mid-row
    std::vector<int16_t> ref((size_t)S * rows), cur((size_t)S * rows);
    unsigned seed = 777;
    for (size_t i = 0; i < ref.size(); ++i) ref[i] = (int16_t)(lcg(seed) & 255);
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < S; ++x) {
            const int sy = y + 1 < rows ? y + 1 : y, sx = x + 2 < S ? x + 2 : x;
            cur[(size_t)y * S + x] = (int16_t)((ref[(size_t)sy * S + sx] + (int)(lcg(seed) % 3) - 1) & 255);
        }
    std::vector<hmme_job> jobs;
    for (int cy = 0; cy + 64 <= H; cy += 64)
        for (int cx = 0; cx + 64 <= W; cx += 64) jobs.push_back(hmme_job{cx, cy, -R + ((cx / 64) % 3), -R - ((cy / 64) % 2)});
    const int n = (int)jobs.size();
    const uint32_t lambda = 460000;
    hmme_group_set_lambda_q16(g, lambda);
    std::vector<int32_t> wX((size_t)n * 593), wY((size_t)n * 593);
    std::vector<uint32_t> wS((size_t)n * 593), wC((size_t)n * 593);
    const int16_t* curO = &cur[(size_t)M * S + M];
    const int16_t* refO = &ref[(size_t)M * S + M];
    hmme_oracle_search_frame(curO, S, refO, S, (const int32_t*)jobs.data(), n, R, lambda, 8, wX.data(), wY.data(), wS.data(), wC.data());
    int bad = 0;
    for (int mode = 0; mode < 2; ++mode) {
        if (hmme_group_configure(g, W, H, M, M, mode == 0 ? HMME_REF_BAND_HALO : HMME_REF_BROADCAST) != HMME_OK) {
            printf("FAIL hmme_group_configure: %s\n", hmme_group_last_error(g)); return 1;
        }
        for (int slot = 0; slot < HMME_GROUP_SLOTS; ++slot) {
            std::vector<int32_t> X((size_t)n * 593, -7), Y((size_t)n * 593, -7);
            std::vector<uint32_t> Sd((size_t)n * 593, 7), C((size_t)n * 593, 7);
            if (hmme_group_search_frame_async(g, slot, curO, S, refO, S, 2, jobs.data(), n, R, X.data(), Y.data(), Sd.data(), C.data()) != HMME_OK ||
                hmme_group_sync(g, slot) != HMME_OK) {
                printf("FAIL frame: %s\n", hmme_group_last_error(g)); return 1;
            }
            for (size_t i = 0; i < X.size(); ++i)
                if (X[i] != wX[i] || Y[i] != wY[i] || Sd[i] != wS[i] || C[i] != wC[i]) {
                    if (bad++ < 5) printf("MISMATCH mode %d slot %d job %zu part %zu\n", mode, slot, i / 593, i % 593);
                }
        }
    }
    int first = 0, count = 0, total = 0;
    for (int i = 0; i < ndev; ++i) { hmme_group_band(g, n, i, &first, &count); if (first != total) ++bad; total += count; }
    if (total != n) ++bad;
    hmme_group_destroy(g);
    printf("%s: %d GPU(s), %d jobs x 2 modes x %d slots, %d mismatches\n", bad ? "FAIL" : "PASS", ndev, n, HMME_GROUP_SLOTS, bad);
    return bad ? 1 : 0;
}
