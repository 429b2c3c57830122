// tmem_bw.cu -- tcgen05.ld / tcgen05.st bandwidth per SM by instruction shape (sm_100a).  16 warps, every warp moves 64 columns x 32 lanes
// per iteration (8 KB) with x16 / x32 / x64 shapes, optionally with several loads in flight before the wait.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmem_bw tools/tmem_bw.cu && tools/tmem_bw
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

template <int N> struct Regs { uint32_t v[N]; };

__device__ __forceinline__ void ld16(uint32_t t, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                   "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(t) : "memory");
}
__device__ __forceinline__ void ld32(uint32_t t, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                   "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                   "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(t) : "memory");
}
__device__ __forceinline__ void ld16x256(uint32_t t, uint32_t* v) {     // 16x256b.x4: 16 lanes x 4 x 256 bit -> 16 registers per thread
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                   "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(t) : "memory");
}
__device__ __forceinline__ void st16(uint32_t t, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(t),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
                 "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}

__global__ void __launch_bounds__(512, 1) k(int mode, int iters, int nwarps, long long* cyc, uint32_t* sink) {
    __shared__ uint32_t tb;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tb)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t t = tb + (uint32_t)((warp >> 2) * 128) + ((uint32_t)(32 * (warp & 3)) << 16);
    uint32_t acc = threadIdx.x;
    uint32_t v[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] = threadIdx.x * 64 + i;
    st16(t, v); st16(t + 16, v + 16); st16(t + 32, v + 32); st16(t + 48, v + 48);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    __syncthreads();
    const long long t0 = clock64();
    if (warp < nwarps) {
        for (int it = 0; it < iters; ++it) {
            if (mode == 0) {            // 4 x (x16, wait)
#pragma unroll
                for (int h = 0; h < 4; ++h) { ld16(t + 16 * h, v); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); acc ^= v[0] + v[15]; }
            } else if (mode == 1) {     // 4 x x16 in flight, one wait
                ld16(t, v); ld16(t + 16, v + 16); ld16(t + 32, v + 32); ld16(t + 48, v + 48);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc ^= v[0] + v[31] + v[63];
            } else if (mode == 2) {     // 2 x x32 in flight
                ld32(t, v); ld32(t + 32, v + 32);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc ^= v[0] + v[31] + v[63];
            } else if (mode == 3) {     // 16x256b.x4 (x4 of them; half the lanes per instruction)
                ld16x256(t, v); ld16x256(t + 32, v + 16); ld16x256(t + ((16u) << 16), v + 32); ld16x256(t + 32 + ((16u) << 16), v + 48);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc ^= v[0] + v[31] + v[63];
            } else if (mode == 4) {     // stores: 4 x x16, one wait
                v[0] = acc;
                st16(t, v); st16(t + 16, v + 16); st16(t + 32, v + 32); st16(t + 48, v + 48);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                acc += 1;
            }
        }
    }
    const long long t1 = clock64();
    sink[blockIdx.x * 512 + threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

int main() {
    setvbuf(stdout, NULL, _IONBF, 0);
    long long* dC; uint32_t* dS;
    CK(cudaMalloc(&dC, 148 * 8)); CK(cudaMalloc(&dS, 148 * 512 * 4));
    const char* names[] = {"ld 32x32b.x16, wait each", "ld 4 x 32x32b.x16, one wait", "ld 2 x 32x32b.x32, one wait", "ld 4 x 16x256b.x4, one wait", "st 4 x 32x32b.x16, one wait"};
    const int iters = 2000;
    for (int nw = 16; nw >= 1; nw /= 4)
        for (int mode = 0; mode < 5; ++mode) {
            k<<<148, 512>>>(mode, iters, nw, dC, dS);
            CK(cudaGetLastError());
            CK(cudaDeviceSynchronize());
            long long c[148];
            CK(cudaMemcpy(c, dC, sizeof(c), cudaMemcpyDeviceToHost));
            double avg = 0; for (int i = 0; i < 148; ++i) avg += (double)c[i]; avg /= 148;
            printf("%2d warps  %-30s %8.1f cycles/iter  %7.1f B/clk/SM\n", nw, names[mode], avg / iters, nw * 8192.0 * iters / avg);
        }
    return 0;
}
