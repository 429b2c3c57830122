// tc_probe.cu -- feasibility probe for moving the 16x16-block partition sums of the search kernel onto tcgen05 (sm_100a).
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
constexpr int kWinRows = 17 + 64; // rows a round can touch: 2 candidates + 15 + block row offset 48
constexpr int kComputeWarps = 16;
constexpr int kThreads = (kComputeWarps + 1) * 32;

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

__global__ void __launch_bounds__(kThreads, 1) tc_probe_kernel(const Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sB = smem;                                        // [256/16 k-blocks][N/8 n-groups][8 rows][16 bytes] = 8 KB (no swizzle, K-major)
    uint2* sWin = reinterpret_cast<uint2*>(smem + 8192);       // kWinRows x kPitch64
    uint32_t* sCur = reinterpret_cast<uint32_t*>(smem + 8192 + ((kWinRows * kPitch64 * 8 + 15) & ~15));   // 64 x 16 words
    __shared__ uint64_t fullBar[4][2], freeBar[4][2], doneBar[4];
    __shared__ uint32_t tmemBase;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int i = tid; i < 256 * kN; i += kThreads) {
        const int n = i / 256, k = i % 256, px = k & 15, py = k >> 4;
        int x, y, w, h;
        part_rect(n, x, y, w, h);
        const uint8_t v = (px >= x && px < x + w && py >= y && py < y + h) ? 1 : 0;
        sB[(k / 16) * ((kN / 8) * 128) + (n / 8) * 128 + (n % 8) * 16 + (k % 16)] = v;
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
            mbar_init(&fullBar[g][0], 4); mbar_init(&fullBar[g][1], 4);
            mbar_init(&freeBar[g][0], 1); mbar_init(&freeBar[g][1], 1);
            mbar_init(&doneBar[g], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmemBase)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes of sB -> visible to the tensor core's async proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tmemBase;
    const int flags = p.flags;
    const long long t0 = clock64();

    if (warp < kComputeWarps) {
        const int b = warp, g = warp >> 2, q = warp & 3;
        const int bx = (b & 3) * 16, by = (b >> 2) * 16;
        const uint32_t tcol = tbase + (uint32_t)(g * 128) + ((uint32_t)(32 * q) << 16);   // group's 128 columns, this warp's lane quarter
        // columns: D[j] at 32*j (j = 0,1); A[buf][j] at 64 + 32*buf + 16*j
        uint32_t best[33];
#pragma unroll
        for (int k = 0; k < 33; ++k) best[k] = 0xFFFFFFFFu;
        const uint32_t* cp = sCur + by * 16 + (bx >> 2);
        for (int r = 0; r < p.rounds; ++r) {
            // unit of this lane in round r: candidate column ux (0..128), candidate rows y0, y0+1
            const int ux = (lane + 32 * r) % 129, y0 = 2 * ((r * 5) % 8);
            const uint2* wp = sWin + (y0 + by) * kPitch64 + ux + bx;
            const uint32_t kb0 = ((uint32_t)((r * 37 + lane * 11) & 1023) << 11) | (uint32_t)((2 * r) * 32 + lane) & 0x7FFu;
            const uint32_t kb1 = ((uint32_t)((r * 37 + lane * 11 + 5) & 1023) << 11) | (uint32_t)((2 * r + 1) * 32 + lane) & 0x7FFu;
            uint32_t a0[16], a1[16];
            uint4 cw[2];
#pragma unroll
            for (int rho = 0; rho < 17; ++rho) {
                uint32_t r0 = 0, r1 = 0, r2 = 0, r3 = 0;
                if (flags & F_ALU) {
                    const uint2 ra = wp[rho * kPitch64], rb = wp[rho * kPitch64 + 8];
                    r0 = ra.x; r1 = ra.y; r2 = rb.x; r3 = rb.y;
                    if (rho < 16) cw[rho & 1] = *reinterpret_cast<const uint4*>(cp + rho * 16);
                }
                if (rho < 16) {                       // candidate 0 meets block row rho
                    const uint4 c = cw[rho & 1];
                    const int o = 4 * (rho & 3);
                    a0[o] = absdiff4(c.x, r0); a0[o + 1] = absdiff4(c.y, r1); a0[o + 2] = absdiff4(c.z, r2); a0[o + 3] = absdiff4(c.w, r3);
                }
                if (rho >= 1) {                       // candidate 1 meets block row rho - 1
                    const uint4 c = cw[(rho - 1) & 1];
                    const int o = 4 * ((rho - 1) & 3);
                    a1[o] = absdiff4(c.x, r0); a1[o + 1] = absdiff4(c.y, r1); a1[o + 2] = absdiff4(c.z, r2); a1[o + 3] = absdiff4(c.w, r3);
                }
                if ((flags & F_ST) && (rho & 3) == 3) {            // strip T of candidate 0 complete
                    const int T = rho >> 2, buf = T & 1;
                    if (T >= 2) {
                        if (flags & F_MMA) mbar_wait(&freeBar[g][buf], (uint32_t)(r & 1));
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    }
                    tmem_st16(tcol + 64 + 32 * buf, a0);
                }
                if ((flags & F_ST) && rho >= 4 && (rho & 3) == 0) {   // strip T of candidate 1 complete: hand the chunk to the MMA warp
                    const int T = rho / 4 - 1, buf = T & 1;
                    tmem_st16(tcol + 64 + 32 * buf + 16, a1);
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&fullBar[g][buf]);
                }
            }
            uint32_t sum16[2] = {0, 0};
            if ((flags & F_MMA) && !(flags & F_LD)) mbar_wait(&doneBar[g], (uint32_t)(r & 1));
            if ((flags & F_LD)) {
                if (flags & F_MMA) mbar_wait(&doneBar[g], (uint32_t)(r & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t d0[16], d1[16];
                    tmem_ld16(tcol + 16 * h, d0);
                    tmem_ld16(tcol + 32 + 16 * h, d1);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (flags & F_KEYS) {
#pragma unroll
                        for (int k = 0; k < 16; ++k) best[16 * h + k] = min(min(best[16 * h + k], d0[k] * 2048u + kb0), d1[k] * 2048u + kb1);
                        if (h == 1) { sum16[0] = d0[12] + d0[13]; sum16[1] = d1[12] + d1[13]; }     // 16x16 = 16x8 top + 16x8 bottom (keys 28, 29)
                    } else {
#pragma unroll
                        for (int k = 0; k < 16; ++k) best[16 * h + k] ^= d0[k] + d1[k];
                    }
                }
                best[32] = min(min(best[32], sum16[0] * 2048u + kb0), sum16[1] * 2048u + kb1);
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            } else {
#pragma unroll
                for (int k = 0; k < 16; ++k) best[k] ^= a0[k] + a1[k];
            }
        }
        uint32_t* o = p.out + ((size_t)blockIdx.x * 512 + tid) * 33;
        for (int k = 0; k < 33; ++k) o[k] = best[k];
    } else if (lane == 0 && (flags & F_MMA)) {
        // ---- MMA issuer: polls the four groups' "chunk stored" barriers and issues 2 candidates x 2 K-steps per chunk
        const uint32_t idesc = (2u << 4) | ((uint32_t)(kN >> 3) << 17) | ((128u >> 4) << 24);   // S32 accumulate, u8 x u8, K-major A and B, N, M = 128
        const uint32_t lbo = p.lbo, sbo = p.sbo;
        const uint64_t descHi = ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);   // version 1, no swizzle
        const uint32_t bAddr = smem_u32(sB);
        int step[4] = {0, 0, 0, 0};
        const int total = 4 * p.rounds;
        int remaining = 4 * total;
        while (remaining > 0) {
            bool any = false;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int s = step[g];
                if (s >= total) continue;
                const int T = s & 3, buf = T & 1;
                if (!mbar_test(&fullBar[g][buf], (uint32_t)((s >> 1) & 1))) continue;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tg = tbase + (uint32_t)(g * 128);
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        const uint32_t kblk = (uint32_t)(4 * T + 2 * ks);                    // 16-byte K block index
                        const uint64_t desc = descHi | (uint64_t)(((bAddr + kblk * lbo) >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16);
                        mma_i8_ts(tg + 32 * j, tg + 64 + 32 * buf + 16 * j + 8 * ks, desc, idesc, (T > 0 || ks > 0) ? 1u : 0u);
                    }
                if (T < 2) mma_commit(&freeBar[g][buf]);
                else if (T == 3) mma_commit(&doneBar[g]);
                step[g] = s + 1;
                --remaining;
                any = true;
            }
            if (!any) {
                __nanosleep(40);
                if (clock64() - t0 > 400000000LL) {
                    if (gDbg[0] == 0) { gDbg[0] = 9999; gDbg[1] = step[0]; gDbg[2] = step[1]; gDbg[3] = step[2]; gDbg[4] = step[3]; }
                    __threadfence_system();
                    asm volatile("trap;");
                }
            }
        }
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
            const int ux = (lane + 32 * r) % 129, y0 = 2 * ((r * 5) % 8);
            for (int j = 0; j < 2; ++j) {
                int ad[16][16];
                for (int y = 0; y < 16; ++y)
                    for (int x = 0; x < 16; ++x)
                        ad[y][x] = abs((int)cur_byte(by + y, bx + x) - (int)ref_byte(y0 + j + by + y, ux + bx + x));
                const uint32_t kb = j == 0 ? (((uint32_t)((r * 37 + lane * 11) & 1023) << 11) | ((uint32_t)((2 * r) * 32 + lane) & 0x7FFu))
                                           : (((uint32_t)((r * 37 + lane * 11 + 5) & 1023) << 11) | ((uint32_t)((2 * r + 1) * 32 + lane) & 0x7FFu));
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
    const size_t smemBytes = 8192 + (size_t)kWinRows * kPitch64 * 8 + 4096 + 1024;
    CK(cudaFuncSetAttribute(tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes));
    const int grid = prop.multiProcessorCount;
    uint32_t* dOut;
    long long* dCyc;
    CK(cudaMalloc(&dOut, (size_t)grid * 512 * 33 * 4));
    CK(cudaMalloc(&dCyc, grid * sizeof(long long)));
    std::vector<uint32_t> got(512 * 33), want;
    std::vector<long long> cyc(grid);

    // ---- correctness: full pipeline, 6 rounds, both LBO/SBO assignments
    const int vr = 6;
    host_reference(vr, want);
    int goodSwap = -1;
    for (int swap = 0; swap < 2; ++swap) {
        Params p{vr, F_ALU | F_ST | F_MMA | F_LD | F_KEYS, dOut, dCyc, 0, 0};
        // K-major, no swizzle: core matrix = 8 rows x 16 bytes (128 B); n-groups 128 B apart, 16-byte K blocks (kN/8)*128 B apart
        p.lbo = swap ? 128 : (kN / 8) * 128;
        p.sbo = swap ? (kN / 8) * 128 : 128;
        CK(cudaMemset(dOut, 0, (size_t)grid * 512 * 33 * 4));
        tc_probe_kernel<<<1, kThreads, smemBytes>>>(p);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("verify swap=%d: kernel failed: %s; dbg code(line)=%d ctx=%d %d %d %d\n", swap, cudaGetErrorString(e), hDbg[0], hDbg[1], hDbg[2], hDbg[3], hDbg[4]); return 3; }
        CK(cudaMemcpy(got.data(), dOut, got.size() * 4, cudaMemcpyDeviceToHost));
        size_t bad = 0, firstBad = 0;
        for (size_t i = 0; i < got.size(); ++i)
            if (got[i] != want[i]) { if (!bad) firstBad = i; ++bad; }
        printf("verify lbo=%u sbo=%u: %zu of %zu running minima differ", p.lbo, p.sbo, bad, got.size());
        if (bad) printf(" (first: thread %zu key %zu got %08x want %08x)", firstBad / 33, firstBad % 33, got[firstBad], want[firstBad]);
        printf("\n");
        if (!bad) goodSwap = swap;
    }
    if (goodSwap < 0) { printf("RESULT: tcgen05 i8 path NOT exact with either descriptor\n"); }

    // ---- throughput: cycles per round with stages switched on one by one, one CTA per SM
    const int rounds = argc > 1 ? atoi(argv[1]) : 200;
    const int sets[][2] = {{F_ALU, 0}, {F_ALU | F_ST, 0}, {F_ALU | F_ST | F_MMA, 0}, {F_ALU | F_ST | F_MMA | F_LD, 0}, {F_ALU | F_ST | F_MMA | F_LD | F_KEYS, 0},
                           {F_ST | F_MMA | F_LD | F_KEYS, 0}, {F_LD | F_KEYS, 0}, {F_LD, 0}, {F_ST, 0}};
    const char* names[] = {"alu(lds+absdiff)", "alu+st", "alu+st+mma", "alu+st+mma+ld", "alu+st+mma+ld+keys (full)", "st+mma+ld+keys (no lds)", "ld+keys only", "ld only", "st only"};
    for (size_t i = 0; i < sizeof(sets) / sizeof(sets[0]); ++i) {
        Params p{rounds, sets[i][0], dOut, dCyc, 0, 0};
        const int swap = goodSwap < 0 ? 0 : goodSwap;
        p.lbo = swap ? 128 : (kN / 8) * 128;
        p.sbo = swap ? (kN / 8) * 128 : 128;
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
        // a round = 16 blocks x 32 lanes x 2 candidates = 64 CTU-candidates per SM
        printf("%-34s %9.1f cycles/round  = %6.2f cycles per CTU-candidate per SM   (%.3f ms for %d rounds; product kernel: ~46 cycles per CTU-candidate)\n",
               names[i], avg / rounds, avg / rounds / 64.0, ms, rounds);
    }
    return 0;
}
