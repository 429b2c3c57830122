// me_generic_kernel.cuh -- exact kernel for everything the packed 8-bit kernel does not take:
// 16-bit "current" blocks (bi-prediction refinement: cur = 2*org - pred in [-255, 510],
// /root/reference/source/Lib/TLibEncoder/TEncSearch.cpp:3702-3712, range 4), planes that are not
// 8-bit, and the support kernels (narrowing upload, result initialisation, finalisation).
//
// One CTA handles one job x one contiguous chunk of candidates in scan order.  Per candidate, 256 threads
// compute the 256 4x4 SADs (exact |a-b| on 16-bit samples like OpenCL abs_diff, sad.cl:171-186), a 17x17
// integral image is built in shared memory, and each thread keeps the running 64-bit arg-min key of up to
// three partitions (rectangle sums from the integral image).  Results merge through the same
// atomicMin(best[job][593]) as the fast kernel, so both kernels share the finalize step.
#pragma once
#include "me_common.cuh"

namespace hmme {

constexpr int kGenThreads = 256;
constexpr int kGenBatch = 4;     // candidates per shared-memory round

struct GenericParams {
    const void* cur;           // picture sample (0,0)
    const void* ref;
    long long curPitch, refPitch;   // elements
    const int4* jobs;
    unsigned long long* best;
    uint32_t lambda;
    int W;                     // 2R+1
    int chunk;                 // candidates per CTA
    int nChunks;
};

template <typename T> __device__ __forceinline__ int ld_px(const void* base, long long idx) {
    return (int)reinterpret_cast<const T*>(base)[idx];
}

template <typename TC, typename TR>
__global__ void __launch_bounds__(kGenThreads) me_generic_kernel(const GenericParams p) {
    __shared__ int sCur[64 * 64];
    __shared__ uint32_t sA[kGenBatch][16][16];
    __shared__ uint32_t sB[kGenBatch][16][16];
    __shared__ uint32_t sII[kGenBatch][17][17];
    __shared__ uint32_t sMvc[kGenBatch];

    const int tid = threadIdx.x;
    const int job = blockIdx.x / p.nChunks, chunkId = blockIdx.x - job * p.nChunks;
    const int4 jb = p.jobs[job];
    const int nCand = p.W * p.W;
    const int c0 = chunkId * p.chunk, c1 = min(nCand, c0 + p.chunk);

    for (int idx = tid; idx < 4096; idx += kGenThreads)
        sCur[idx] = ld_px<TC>(p.cur, (long long)(jb.y + (idx >> 6)) * p.curPitch + jb.x + (idx & 63));
    for (int idx = tid; idx < kGenBatch * 17 * 17; idx += kGenThreads) (&sII[0][0][0])[idx] = 0;

    // up to three partitions per thread: corners in 4x4 units
    int px0[3], py0[3], px1[3], py1[3];
    unsigned long long best[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int part = tid + k * kGenThreads;
        const PartRect r = part_rect(part < HMME_NPARTS ? part : 0);
        px0[k] = r.x >> 2; py0[k] = r.y >> 2; px1[k] = (r.x + r.w) >> 2; py1[k] = (r.y + r.h) >> 2;
        best[k] = kNoWinner;
    }
    __syncthreads();

    const int bi = tid & 15, bj = tid >> 4;
    const long long refBase = (long long)(jb.y + jb.w) * p.refPitch + (jb.x + jb.z);   // window origin, linear
    for (int cb = c0; cb < c1; cb += kGenBatch) {
        // 4x4 SADs
#pragma unroll
        for (int n = 0; n < kGenBatch; ++n) {
            const int c = cb + n;
            uint32_t s = 0;
            if (c < c1) {
                const int y = c / p.W, x = c - y * p.W;
                const long long o = refBase + (long long)(y + 4 * bj) * p.refPitch + x + 4 * bi;
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int d = sCur[(4 * bj + r) * 64 + 4 * bi + q] - ld_px<TR>(p.ref, o + r * p.refPitch + q);
                        s += (uint32_t)(d < 0 ? -d : d);
                    }
                if (tid == 0) sMvc[n] = mv_cost(p.lambda, x + jb.z, y + jb.w);
            }
            sA[n][bj][bi] = s;
        }
        __syncthreads();
        // row prefix
#pragma unroll
        for (int n = 0; n < kGenBatch; ++n) {
            uint32_t s = 0;
            for (int i = 0; i <= bi; ++i) s += sA[n][bj][i];
            sB[n][bj][bi] = s;
        }
        __syncthreads();
        // column prefix -> integral image (row/column 0 stay zero)
#pragma unroll
        for (int n = 0; n < kGenBatch; ++n) {
            uint32_t s = 0;
            for (int j = 0; j <= bj; ++j) s += sB[n][j][bi];
            sII[n][bj + 1][bi + 1] = s;
        }
        __syncthreads();
#pragma unroll
        for (int n = 0; n < kGenBatch; ++n) {
            const int c = cb + n;
            if (c < c1) {
                const uint32_t mvc = sMvc[n];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const uint32_t s = sII[n][py1[k]][px1[k]] - sII[n][py0[k]][px1[k]] - sII[n][py1[k]][px0[k]] + sII[n][py0[k]][px0[k]];
                    const unsigned long long key = ((unsigned long long)(uint32_t)(s + mvc) << 32) | (uint32_t)c;
                    best[k] = key < best[k] ? key : best[k];
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int part = tid + k * kGenThreads;
        if (part < HMME_NPARTS && best[k] != kNoWinner) atomicMin(p.best + (size_t)job * HMME_NPARTS + part, best[k]);
    }
}

// best[] <- "no winner" (TEncOpenCL.cpp:366-392: minSad = UINT_MAX, X = Y = 0); run once per allocation, the searches
// restore this state themselves when they read their results out
__global__ void me_init_kernel(unsigned long long* best, size_t n) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) best[i] = kNoWinner;
}

// key -> X, Y, sad (ruiCosts), cost (minSad) for result slot i of a job with search-range origin (ltx, lty)
__device__ __forceinline__ void decode_key(unsigned long long key, int ltx, int lty, int W, uint32_t lambda, int i,
                                           int32_t* X, int32_t* Y, uint32_t* sad, uint32_t* cost) {
    if (key == kNoWinner) { X[i] = 0; Y[i] = 0; sad[i] = 0; cost[i] = 0xFFFFFFFFu; return; }
    const uint32_t c = (uint32_t)(key >> 32), g = (uint32_t)key;
    const int y = (int)(g / (uint32_t)W), x = (int)(g - (uint32_t)y * (uint32_t)W);
    const int mvx = x + ltx, mvy = y + lty;
    X[i] = mvx; Y[i] = mvy; cost[i] = c; sad[i] = c - mv_cost(lambda, mvx, mvy);
}

// ---- bi-prediction blocks on the packed 8-bit kernel.  The refinement's block is cur = 2*org - pred, a signed 16-bit value
// (TEncSearch.cpp:3702-3712), searched against an 8-bit picture.  For r in [0, 255] and any integer c:
//     |c - r| = |clamp(c, 0, 255) - r| + |c - clamp(c, 0, 255)|
// (c < 0: r - c = (r - 0) + (0 - c);  c > 255: c - r = (c - 255) + (255 - r)).  The second term does not depend on the candidate, so
// the SAD of a partition is the 8-bit SAD of the clamped block plus a per-partition constant; the arg-min and its tie-break are
// unchanged, and sad / cost get the constant added at finalisation.  This kernel prepares both: per job the clamped 64x64 block
// (dense 4 KiB record) and the 593 constants (4x4 sums -> integral image -> rectangle sums).  One CTA per job, 256 threads.
__global__ void __launch_bounds__(256) me_bipred_prep_kernel(const int16_t* __restrict__ cur, long long curPitch, const int4* __restrict__ jobs,
                                                             uint8_t* __restrict__ blocks, uint32_t* __restrict__ offsets) {
    __shared__ uint32_t sA[16][16], sB[16][16], sII[17][17];
    const int tid = threadIdx.x, job = blockIdx.x, bi = tid & 15, bj = tid >> 4;
    const int4 jb = jobs[job];
    const int16_t* c0 = cur + (long long)(jb.y + 4 * bj) * curPitch + jb.x + 4 * bi;
    uint8_t* out = blocks + (size_t)job * 4096 + (4 * bj) * 64 + 4 * bi;
    uint32_t ex = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        uint32_t w = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int v = c0[r * curPitch + q];
            const int cl = min(max(v, 0), 255);
            ex += (uint32_t)abs(v - cl);
            w |= (uint32_t)cl << (8 * q);
        }
        *reinterpret_cast<uint32_t*>(out + r * 64) = w;
    }
    sA[bj][bi] = ex;
    if (tid < 17) { sII[0][tid] = 0; sII[tid][0] = 0; }
    __syncthreads();
    uint32_t s = 0;
    for (int i = 0; i <= bi; ++i) s += sA[bj][i];
    sB[bj][bi] = s;
    __syncthreads();
    s = 0;
    for (int j = 0; j <= bj; ++j) s += sB[j][bi];
    sII[bj + 1][bi + 1] = s;
    __syncthreads();
    for (int part = tid; part < HMME_NPARTS; part += 256) {
        const PartRect r = part_rect(part);
        const int x0 = r.x >> 2, y0 = r.y >> 2, x1 = (r.x + r.w) >> 2, y1 = (r.y + r.h) >> 2;
        offsets[(size_t)job * HMME_NPARTS + part] = sII[y1][x1] - sII[y0][x1] - sII[y1][x0] + sII[y0][x0];
    }
}

// Finalisation pass of the generic path; outputs are four planes of [njobs][593].  Every key is handed back as "no winner"
// (the state me_init_kernel set up once), so the arg-min scratch is always ready for the next search.
// `offsets` (NULL, or [njobs][593]): the candidate-independent part of a bi-prediction block's SADs (me_bipred_prep_kernel).
__global__ void me_finalize_kernel(unsigned long long* best, const int4* jobs, int njobs, int W, uint32_t lambda,
                                   int32_t* X, int32_t* Y, uint32_t* sad, uint32_t* cost, const uint32_t* __restrict__ offsets) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= njobs * HMME_NPARTS) return;
    const unsigned long long key = best[i];
    best[i] = kNoWinner;
    const int4 jb = jobs[i / HMME_NPARTS];
    decode_key(key, jb.z, jb.w, W, lambda, i, X, Y, sad, cost);
    if (offsets && key != kNoWinner) { const uint32_t o = offsets[i]; sad[i] += o; cost[i] += o; }
}

// int16 -> uint8 narrowing of a whole padded plane with a content check (the reference path is only
// defined for 8-bit internal depth); *flag != 0 afterwards means out-of-range samples.
__global__ void me_narrow_kernel(const int16_t* __restrict__ src, uint8_t* __restrict__ dst, size_t n, int* flag) {
    size_t i = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * 8;
    if (i >= n) return;
    bool bad = false;
    if (i + 8 <= n && ((reinterpret_cast<uintptr_t>(src + i) & 15) == 0) && ((reinterpret_cast<uintptr_t>(dst + i) & 7) == 0)) {
        const uint4 v = *reinterpret_cast<const uint4*>(src + i);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            bad |= (w[k] & 0xFF00FF00u) != 0;
            const uint32_t two = (w[k] & 0xFFu) | ((w[k] >> 8) & 0xFF00u);
            if (k < 2) lo |= two << (16 * k); else hi |= two << (16 * (k - 2));
        }
        *reinterpret_cast<uint2*>(dst + i) = make_uint2(lo, hi);
    } else {
        for (size_t k = i; k < n && k < i + 8; ++k) { const int s = src[k]; bad |= (s < 0 || s > 255); dst[k] = (uint8_t)s; }
    }
    if (bad) atomicOr(flag, 1);
}

// Same narrowing for a rectangle: src = dense staging rows of `spitch` int16 (spitch a multiple of 8, so every row starts 16-byte
// aligned), dst = the rectangle's top-left sample inside an 8-bit plane of pitch `dpitch`.  One thread per 8 samples of a row.
__global__ void me_narrow_rect_kernel(const int16_t* __restrict__ src, int spitch, uint8_t* __restrict__ dst, long long dpitch, int cols, int rows,
                                      int* flag) {
    const int perRow = (cols + 7) >> 3;
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= (long long)perRow * rows) return;
    const int r = (int)(t / perRow), q = (int)(t - (long long)r * perRow) * 8;
    const int16_t* s = src + (size_t)r * spitch + q;
    uint8_t* d = dst + r * dpitch + q;
    bool bad = false;
    if (q + 8 <= cols) {
        const uint4 v = *reinterpret_cast<const uint4*>(s);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            bad |= (w[k] & 0xFF00FF00u) != 0;
            const uint32_t two = (w[k] & 0xFFu) | ((w[k] >> 8) & 0xFF00u);
            if (k < 2) lo |= two << (16 * k); else hi |= two << (16 * (k - 2));
        }
        if ((reinterpret_cast<uintptr_t>(d) & 7) == 0) *reinterpret_cast<uint2*>(d) = make_uint2(lo, hi);
        else {
#pragma unroll
            for (int k = 0; k < 4; ++k) { d[k] = (uint8_t)(lo >> (8 * k)); d[4 + k] = (uint8_t)(hi >> (8 * k)); }
        }
    } else {
        for (int k = 0; q + k < cols; ++k) { const int x = s[k]; bad |= (x < 0 || x > 255); d[k] = (uint8_t)x; }
    }
    if (bad) atomicOr(flag, 1);
}

// Integer-ALU issue-rate probe: a dependent-free stream of VABSDIFF4.U8.ACC (the kernel's dominant ALU
// instruction); lanes/clk/SM from clock64, SM MHz from clock64 vs the event time.
__global__ void __launch_bounds__(1024) me_alu_probe_kernel(uint32_t* out, unsigned long long* cyc, int iters, uint32_t seed) {
    uint32_t a[8], b = seed * 0x9E3779B9u + threadIdx.x, c = seed ^ 0x5bd1e995u;
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = threadIdx.x * 31 + k + seed;
    const unsigned long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int s = 0; s < 32; ++s)
#pragma unroll
            for (int k = 0; k < 8; ++k) asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(a[k]) : "r"(b), "r"(c));
    }
    const unsigned long long t1 = clock64();
    uint32_t r = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) r ^= a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

}  // namespace hmme
