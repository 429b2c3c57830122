// tc_probe2.cu -- second feasibility probe (fp16 variant) for moving the 16x16-block partition sums onto tcgen05 (sm_100a).
// Here the A operand is not per-pixel differences but the 4x2-pixel SADs (two chained VABSDIFF4.U8.ACC, <= 2040): as a 16-bit
// pattern such a value IS the fp16 number s * 2^-24 (denormals and the first normal binade are linear in the bit pattern), two
// per 32-bit word, so a 16x16 block is K = 32 fp16 elements; B is 1.0/0.0 in fp16, D accumulates in fp32 on top of 0.5 (a
// constant extra MMA), whose bit pattern is then 0x3F000000 + sum -- and the constant vanishes in key = bits * 2^11 + kb mod 2^32.
// Three candidate rows per lane as in the product kernel; per warp 96 D columns + 24 A columns + 8 constant columns = 128.
// (tc_probe.cu is the u8 per-pixel variant: exact, but TMEM capacity forces two candidate rows per lane, which is shared-memory bound.)
//
// Idea under test: per (candidate, 16x16 block) the 256 per-pixel absolute differences (VABSDIFF4 without accumulate: four u8
// |a-b| per word) are stored to tensor memory as the A operand of a u8 x u8 -> s32 tcgen05.mma (kind::i8, M = 128 candidates =
// 4 warps x 32 lanes, K = 256 pixels in 4 chunks of 64, N = 32 partitions), the B operand is the constant 0/1 pixel->partition
// matrix in shared memory, and the 32 partition sums per candidate come back with tcgen05.ld for key formation and arg-min.
// This replaces the kernel's 37-add hierarchy per (candidate, block).  The probe has the real kernel's thread layout (16 compute
// warps = 16 blocks, lane = candidate column, 2 candidate rows per round sharing reference rows) plus one MMA-issuing warp, a
// synthetic window/CTU in shared memory, and checks every resulting running minimum of CTA 0 against a host computation.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/tc_probe tools/tc_probe.cu && tools/tc_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

constexpr int kN = 32;            // partitions per MMA (N)
constexpr int kPitch64 = 187;     // sliding 64-bit entries per window row (as in the product kernel)
constexpr int kWinRows = 18 + 64; // rows a round can touch: 2 candidates + 15 + block row offset 48
constexpr int kComputeWarps = 16;
constexpr int kThreads = kComputeWarps * 32;

enum { F_ALU = 1, F_ST = 2, F_MMA = 4, F_LD = 8, F_KEYS = 16, F_LBOSWAP = 32 };

struct Params {
    int rounds;
    int flags;
    uint32_t* out;        // [grid][512][33] running minima
    long long* cycles;    // [grid]
    uint32_t lbo, sbo;    // B descriptor strides (bytes)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(smem_u32(b)) : "memory"); }
__device__ volatile int* gDbg;     // mapped host memory: [0] = code of the wait that timed out, [1..] = context
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t parity) {
    uint32_t ok;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_dbg(uint64_t* b, uint32_t parity, int code, int a0, int a1) {
    const long long t0 = clock64();
    while (!mbar_try(b, parity)) {
        if (clock64() - t0 > 400000000LL) {
            if (gDbg[0] == 0) { gDbg[0] = code; gDbg[1] = a0; gDbg[2] = a1; gDbg[3] = (int)threadIdx.x; gDbg[4] = (int)blockIdx.x; }
            __threadfence_system();
            asm volatile("trap;");
        }
    }
}
#define mbar_wait(b, parity) mbar_wait_dbg(b, parity, __LINE__, r, 0)
__device__ __forceinline__ bool mbar_test(uint64_t* b, uint32_t parity) {
    uint32_t ok;
    asm volatile("{ .reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ uint32_t absdiff4(uint32_t a, uint32_t b) {     // four byte-wise |a - b|, no accumulate
    uint32_t d;
    asm("vabsdiff4.u32.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(0u));
    return d;
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
                 "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                   "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void mma_i8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc),
                 "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__host__ __device__ inline uint32_t hashw(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

// rectangle of block-level partition k (0..32) inside a 16x16 block: x, y, w, h
__host__ __device__ inline void part_rect(int k, int& x, int& y, int& w, int& h) {
    if (k < 8) { x = 8 * (k & 1); y = 4 * (k >> 1); w = 8; h = 4; }
    else if (k < 16) { x = 4 * ((k - 8) & 3); y = 8 * ((k - 8) >> 2); w = 4; h = 8; }
    else if (k < 20) { x = 8 * ((k - 16) & 1); y = 8 * ((k - 16) >> 1); w = 8; h = 8; }
    else if (k == 20) { x = 0; y = 0; w = 16; h = 4; }
    else if (k == 21) { x = 0; y = 12; w = 16; h = 4; }
    else if (k == 22) { x = 0; y = 0; w = 16; h = 12; }
    else if (k == 23) { x = 0; y = 4; w = 16; h = 12; }
    else if (k == 24) { x = 0; y = 0; w = 4; h = 16; }
    else if (k == 25) { x = 12; y = 0; w = 4; h = 16; }
    else if (k == 26) { x = 0; y = 0; w = 12; h = 16; }
    else if (k == 27) { x = 4; y = 0; w = 12; h = 16; }
    else if (k == 28) { x = 0; y = 0; w = 16; h = 8; }
    else if (k == 29) { x = 0; y = 8; w = 16; h = 8; }
    else if (k == 30) { x = 0; y = 0; w = 8; h = 16; }
    else if (k == 31) { x = 8; y = 0; w = 8; h = 16; }
    else { x = 0; y = 0; w = 16; h = 16; }
}

// window entry x of row r: bytes x..x+7 of the synthetic reference row; CTU word (row, i)
__host__ __device__ inline uint8_t ref_byte(int row, int col) { return (uint8_t)(hashw((uint32_t)(row * 4099 + col) + 77u) >> 11); }
__host__ __device__ inline uint8_t cur_byte(int row, int col) { return (uint8_t)(hashw((uint32_t)(row * 64 + col) + 900001u) >> 7); }


__device__ __forceinline__ uint32_t sad4_acc(uint32_t a, uint32_t b, uint32_t acc) {
    uint32_t d;
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(acc));
    return d;
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
                 "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ uint32_t pack16(uint32_t lo, uint32_t hi) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, 65536, %2;" : "=r"(d) : "r"(hi), "r"(lo));
    return d;
}

#define READOUT(RR) { \
            if (flags & F_MMA) mbar_wait_dbg(&doneBar[g], (uint32_t)((RR) & 1), __LINE__, RR, 1); \
            if ((flags & F_LD)) { \
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); \
                uint32_t s28[3], s29[3]; \
                for (int h = 0; h < 2; ++h) { \
                    uint32_t d0[16], d1[16], d2[16]; \
                    tmem_ld16(tcol + 16 * h, d0); \
                    tmem_ld16(tcol + 32 + 16 * h, d1); \
                    tmem_ld16(tcol + 64 + 16 * h, d2); \
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); \
                    if (flags & F_KEYS) { \
                        for (int k = 0; k < 16; ++k) { \
                            best[16 * h + k] = min(min(best[16 * h + k], d0[k] * 2048u + kbPrev[0]), d1[k] * 2048u + kbPrev[1]); \
                            best[16 * h + k] = min(best[16 * h + k], d2[k] * 2048u + kbPrev[2]); \
                        } \
                        if (h == 1) { s28[0] = d0[12]; s29[0] = d0[13]; s28[1] = d1[12]; s29[1] = d1[13]; s28[2] = d2[12]; s29[2] = d2[13]; } \
                    } else { \
                        for (int k = 0; k < 16; ++k) best[16 * h + k] ^= d0[k] + d1[k] + d2[k]; \
                        if (h == 1) { s28[0] = s29[0] = s28[1] = s29[1] = s28[2] = s29[2] = d0[0]; } \
                    } \
                } \
                best[32] = min(min(best[32], (s28[0] + s29[0]) * 2048u + kbPrev[0]), (s28[1] + s29[1]) * 2048u + kbPrev[1]); \
                best[32] = min(best[32], (s28[2] + s29[2]) * 2048u + kbPrev[2]); \
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); \
            } else { \
                for (int k = 0; k < 8; ++k) best[k] ^= a[0][k] + a[1][k] + a[2][k]; \
            } \
         }

__global__ void __launch_bounds__(kThreads, 1) tc_probe_kernel(const Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint16_t* sB = reinterpret_cast<uint16_t*>(smem);          // main: [32 K / 8][N/8][8 rows][8 fp16] = 2 KB ; const at +2048: [16 K / 8][N/8][8][8] = 1 KB
    uint2* sWin = reinterpret_cast<uint2*>(smem + 4096);       // kWinRows x kPitch64
    uint32_t* sCur = reinterpret_cast<uint32_t*>(smem + 4096 + ((kWinRows * kPitch64 * 8 + 15) & ~15));   // 64 x 16 words
    __shared__ uint64_t freeBar[4], doneBar[4];
    __shared__ uint32_t arrivals[4];
    __shared__ uint32_t tmemBase;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int i = tid; i < 32 * kN; i += kThreads) {            // K element k = 4 * rowPair + colGroup  (cell: rows 2rp..2rp+1, cols 4i..4i+3)
        const int n = i / 32, k = i % 32, px = 4 * (k & 3), py = 2 * (k >> 2);
        int x, y, w, h;
        part_rect(n, x, y, w, h);
        const uint16_t v = (px >= x && px < x + w && py >= y && py < y + h) ? 0x3C00 : 0;      // 1.0
        sB[(k / 8) * ((kN / 8) * 64) + (n / 8) * 64 + (n % 8) * 8 + (k % 8)] = v;
    }
    for (int i = tid; i < 16 * kN; i += kThreads) {
        const int n = i / 16, k = i % 16;
        sB[1024 + (k / 8) * ((kN / 8) * 64) + (n / 8) * 64 + (n % 8) * 8 + (k % 8)] = (k == 0) ? 0x3800 : 0;   // 0.5
    }
    for (int i = tid; i < kWinRows * kPitch64; i += kThreads) {
        const int r = i / kPitch64, x = i % kPitch64;
        uint32_t lo = 0, hi = 0;
        for (int b = 0; b < 4; ++b) { lo |= (uint32_t)ref_byte(r, x + b) << (8 * b); hi |= (uint32_t)ref_byte(r, x + 4 + b) << (8 * b); }
        sWin[i] = make_uint2(lo, hi);
    }
    for (int i = tid; i < 1024; i += kThreads) {
        uint32_t v = 0;
        for (int b = 0; b < 4; ++b) v |= (uint32_t)cur_byte(i >> 4, 4 * (i & 15) + b) << (8 * b);
        sCur[i] = v;
    }
    if (tid == 0) {
        for (int g = 0; g < 4; ++g) {
            arrivals[g] = 0;
            mbar_init(&freeBar[g], 1);
            mbar_init(&doneBar[g], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmemBase)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tmemBase;
    const int flags = p.flags;
    if (warp < kComputeWarps) {                                // constant A columns: fp16 element 0 = 1.0, the other 15 = 0
        const uint32_t tcol = tbase + (uint32_t)((warp >> 2) * 128) + ((uint32_t)(32 * (warp & 3)) << 16);
        const uint32_t one[8] = {0x00003C00u, 0, 0, 0, 0, 0, 0, 0};
        tmem_st8(tcol + 120, one);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const long long t0 = clock64();
    const uint32_t idesc = (1u << 4) | ((uint32_t)(kN >> 3) << 17) | ((128u >> 4) << 24);   // F32 accumulate, F16 x F16, K-major A and B, N, M = 128
    const uint64_t descHi = ((uint64_t)((p.sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)((p.lbo >> 4) & 0x3FFF) << 16);
    const uint32_t bAddr = smem_u32(sB);

    if (warp < kComputeWarps) {
        const int b = warp, g = warp >> 2, q = warp & 3;
        const int bx = (b & 3) * 16, by = (b >> 2) * 16;
        const uint32_t tcol = tbase + (uint32_t)(g * 128) + ((uint32_t)(32 * q) << 16);
        // columns: D[j] at 32*j (j = 0..2); A[j] at 96 + 8*j; constant at 120
        uint32_t best[33];
#pragma unroll
        for (int k = 0; k < 33; ++k) best[k] = 0xFFFFFFFFu;
        const uint32_t* cp = sCur + by * 16 + (bx >> 2);
        uint32_t kbPrev[3] = {0, 0, 0};
        long long tAtom = 0, tIssue = 0; int nIssue = 0;
        uint32_t a[3][8] = {};
        for (int r = 0; r < p.rounds; ++r) {
            const int ux = (lane + 32 * r) % 129, y0 = 3 * ((r * 5) % 5);
            const uint2* wp = sWin + (y0 + by) * kPitch64 + ux + bx;
            uint32_t kb[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) kb[j] = ((uint32_t)((r * 37 + lane * 11 + 5 * j) & 1023) << 11) | ((uint32_t)((3 * r + j) * 32 + lane) & 0x7FFu);
            uint32_t acc[3][4];
            uint4 cw[3];
#pragma unroll
            for (int rho = 0; rho < 18; ++rho) {
                uint32_t r0 = 0, r1 = 0, r2 = 0, r3 = 0;
                if (flags & F_ALU) {
                    const uint2 ra = wp[rho * kPitch64], rb = wp[rho * kPitch64 + 8];
                    r0 = ra.x; r1 = ra.y; r2 = rb.x; r3 = rb.y;
                    if (rho < 16) cw[rho % 3] = *reinterpret_cast<const uint4*>(cp + rho * 16);
                }
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const int rr = rho - j;
                    if (rr < 0 || rr > 15) continue;
                    const uint4 c = cw[rr % 3];
                    if (rr & 1) {
                        acc[j][0] = sad4_acc(c.x, r0, acc[j][0]); acc[j][1] = sad4_acc(c.y, r1, acc[j][1]);
                        acc[j][2] = sad4_acc(c.z, r2, acc[j][2]); acc[j][3] = sad4_acc(c.w, r3, acc[j][3]);
                        const int w = 2 * ((rr >> 1) & 3);
                        a[j][w] = pack16(acc[j][0], acc[j][1]);
                        a[j][w + 1] = pack16(acc[j][2], acc[j][3]);
                    } else {
                        acc[j][0] = sad4_acc(c.x, r0, 0); acc[j][1] = sad4_acc(c.y, r1, 0);
                        acc[j][2] = sad4_acc(c.z, r2, 0); acc[j][3] = sad4_acc(c.w, r3, 0);
                    }
                    if (rr == 7 && j == 0 && r > 0) READOUT(r - 1)       // deferred by half a round: the MMAs of round r-1 finished long ago
                    if ((flags & F_ST) && (rr & 7) == 7) {     // 8 rows of candidate j complete: one K step
                        const int ch = rr >> 3;
                        if (ch == 1 && j == 0) {
                            if (flags & F_MMA) mbar_wait(&freeBar[g], (uint32_t)(r & 1));
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        }
                        tmem_st8(tcol + 96 + 8 * j, a[j]);
                        if (j == 2) {
                            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                            __syncwarp();
                            if (lane == 0 && (flags & F_MMA)) {
                                // the last of the group's four warps to deliver this chunk issues its MMAs
                                uint32_t old;
                                const long long ta = clock64();
                                asm volatile("atom.shared.acq_rel.cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(&arrivals[g])) : "memory");
                                const long long tb2 = clock64();
                                tAtom += tb2 - ta;
                                if ((old & 3u) == 3u) {
                                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                                    const uint32_t tg = tbase + (uint32_t)(g * 128);
#pragma unroll
                                    for (int jj = 0; jj < 3; ++jj) {
                                        if (ch == 0) mma_f16_ts(tg + 32 * jj, tg + 120, descHi | (uint64_t)(((bAddr + 2048) >> 4) & 0x3FFF), idesc, 0u);   // D = 0.5
                                        mma_f16_ts(tg + 32 * jj, tg + 96 + 8 * jj, descHi | (uint64_t)(((bAddr + (uint32_t)(2 * ch) * p.lbo) >> 4) & 0x3FFF), idesc, 1u);
                                    }
                                    if (ch == 0) mma_commit(&freeBar[g]);
                                    else mma_commit(&doneBar[g]);
                                    tIssue += clock64() - tb2; ++nIssue;
                                }
                            }
                            __syncwarp();
                        }
                    }
                }
            }
            kbPrev[0] = kb[0]; kbPrev[1] = kb[1]; kbPrev[2] = kb[2];
        }
        READOUT(p.rounds - 1)
        if (lane == 0 && blockIdx.x == 0) { p.cycles[200 + warp * 3] = tAtom; p.cycles[201 + warp * 3] = tIssue; p.cycles[202 + warp * 3] = nIssue; }
        uint32_t* o = p.out + ((size_t)blockIdx.x * 512 + tid) * 33;
        for (int k = 0; k < 33; ++k) o[k] = best[k];
    }
    const long long t1 = clock64();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) p.cycles[blockIdx.x] = t1 - t0;
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
}

static void host_reference(int rounds, std::vector<uint32_t>& best) {
    best.assign(512 * 33, 0xFFFFFFFFu);
    std::vector<int> rx(33), ry(33), rw(33), rh(33);
    for (int k = 0; k < 33; ++k) part_rect(k, rx[k], ry[k], rw[k], rh[k]);
    for (int tid = 0; tid < 512; ++tid) {
        const int lane = tid & 31, b = tid >> 5, bx = (b & 3) * 16, by = (b >> 2) * 16;
        for (int r = 0; r < rounds; ++r) {
            const int ux = (lane + 32 * r) % 129, y0 = 3 * ((r * 5) % 5);
            for (int j = 0; j < 3; ++j) {
                int ad[16][16];
                for (int y = 0; y < 16; ++y)
                    for (int x = 0; x < 16; ++x)
                        ad[y][x] = abs((int)cur_byte(by + y, bx + x) - (int)ref_byte(y0 + j + by + y, ux + bx + x));
                const uint32_t kb = ((uint32_t)((r * 37 + lane * 11 + 5 * j) & 1023) << 11) | ((uint32_t)((3 * r + j) * 32 + lane) & 0x7FFu);
                for (int k = 0; k < 33; ++k) {
                    uint32_t s = 0;
                    for (int y = ry[k]; y < ry[k] + rh[k]; ++y)
                        for (int x = rx[k]; x < rx[k] + rw[k]; ++x) s += ad[y][x];
                    uint32_t& bb = best[tid * 33 + k];
                    bb = std::min(bb, s * 2048u + kb);
                }
            }
        }
    }
}

int main(int argc, char** argv) {
    int dev = 0;
    setvbuf(stdout, NULL, _IONBF, 0);
    CK(cudaSetDevice(dev));
    int* hDbg;
    CK(cudaHostAlloc(&hDbg, 64, cudaHostAllocMapped));
    memset(hDbg, 0, 64);
    {
        int* dDbgPtr;
        CK(cudaHostGetDevicePointer(&dDbgPtr, hDbg, 0));
        CK(cudaMemcpyToSymbol(gDbg, &dDbgPtr, sizeof(dDbgPtr)));
    }
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    printf("device %s sm_%d%d, %d SMs\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
    const size_t smemBytes = 4096 + (size_t)kWinRows * kPitch64 * 8 + 16 + 4096 + 1024;
    CK(cudaFuncSetAttribute(tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes));
    const int grid = prop.multiProcessorCount;
    uint32_t* dOut;
    long long* dCyc;
    CK(cudaMalloc(&dOut, (size_t)grid * 512 * 33 * 4));
    CK(cudaMalloc(&dCyc, (grid + 200) * sizeof(long long)));
    std::vector<uint32_t> got(512 * 33), want;
    std::vector<long long> cyc(grid);

    const int vr = 7;
    host_reference(vr, want);
    {
        Params p{vr, F_ALU | F_ST | F_MMA | F_LD | F_KEYS, dOut, dCyc, (kN / 8) * 128, 128};
        CK(cudaMemset(dOut, 0, (size_t)grid * 512 * 33 * 4));
        tc_probe_kernel<<<1, kThreads, smemBytes>>>(p);
        CK(cudaGetLastError());
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("verify: kernel failed: %s; dbg code(line)=%d ctx=%d %d %d %d\n", cudaGetErrorString(e), hDbg[0], hDbg[1], hDbg[2], hDbg[3], hDbg[4]); return 3; }
        CK(cudaMemcpy(got.data(), dOut, got.size() * 4, cudaMemcpyDeviceToHost));
        size_t bad = 0, firstBad = 0;
        for (size_t i = 0; i < got.size(); ++i)
            if (got[i] != want[i]) { if (!bad) firstBad = i; ++bad; }
        printf("verify fp16 variant: %zu of %zu running minima differ", bad, got.size());
        if (bad) printf(" (first: thread %zu key %zu got %08x want %08x)", firstBad / 33, firstBad % 33, got[firstBad], want[firstBad]);
        printf("\n");
    }

    const int rounds = argc > 1 ? atoi(argv[1]) : 200;
    const int sets[] = {F_ALU, F_ALU | F_ST, F_ALU | F_ST | F_MMA, F_ALU | F_ST | F_MMA | F_LD, F_ALU | F_ST | F_MMA | F_LD | F_KEYS,
                        F_ST | F_MMA | F_LD | F_KEYS, F_LD | F_KEYS, F_LD, F_ST};
    const char* names[] = {"alu(lds+sad+pack)", "alu+st", "alu+st+mma", "alu+st+mma+ld", "alu+st+mma+ld+keys (full)", "st+mma+ld+keys (no lds)", "ld+keys only", "ld only", "st only"};
    for (size_t i = 0; i < sizeof(sets) / sizeof(sets[0]); ++i) {
        Params p{rounds, sets[i], dOut, dCyc, (kN / 8) * 128, 128};
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        tc_probe_kernel<<<grid, kThreads, smemBytes>>>(p);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        tc_probe_kernel<<<grid, kThreads, smemBytes>>>(p);
        CK(cudaEventRecord(e1));
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: kernel failed: %s; dbg code(line)=%d ctx=%d %d %d %d\n", names[i], cudaGetErrorString(e), hDbg[0], hDbg[1], hDbg[2], hDbg[3], hDbg[4]); return 4; }
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        CK(cudaMemcpy(cyc.data(), dCyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
        double avg = 0;
        for (int k = 0; k < grid; ++k) avg += (double)cyc[k];
        avg /= grid;
        // a round = 16 blocks x 32 lanes x 3 candidates = 96 CTU-candidates per SM
        if (sets[i] & F_MMA) {
            long long t[48];
            CK(cudaMemcpy(t, dCyc + 200, sizeof(t), cudaMemcpyDeviceToHost));
            long long sa = 0, si = 0, ni = 0;
            for (int w = 0; w < 16; ++w) { sa += t[3 * w]; si += t[3 * w + 1]; ni += t[3 * w + 2]; }
            printf("    CTA 0: atom %.1f cycles per arrival, MMA issue block %.1f cycles (x%lld)\n", (double)sa / (16.0 * 2 * rounds), ni ? (double)si / ni : 0.0, ni);
        }
        printf("%-34s %9.1f cycles/round  = %6.2f cycles per CTU-candidate per SM   (%.3f ms for %d rounds; product kernel: ~46 cycles per CTU-candidate)\n",
               names[i], avg / rounds, avg / rounds / 96.0, ms, rounds);
    }
    return 0;
}
