// tma_probe.cu -- can tensor-map TMA build the search kernel's sliding window directly?  Loads the same (rows x 208 byte) box of a u8 plane
// eight times with the column coordinate shifted by 0..7 bytes (and a per-copy skew), checks the bytes, and times issue -> completion.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_probe tools/tma_probe.cu -lcuda && tools/tma_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

struct alignas(64) Maps { CUtensorMap ref; };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(512, 1) k(const __grid_constant__ Maps m, int c0, int r0, int rows, int inner, int ncopies, int colStep, int bufStride, int dstOff,
                                            uint8_t* out, long long* cyc) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    const int tid = threadIdx.x;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long t0 = clock64();
    if (tid == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"((uint32_t)(ncopies * rows * inner)) : "memory");
        for (int s = 0; s < ncopies; ++s) {
            const int cx = c0 + s * colStep, cy = r0;
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                             smem_u32(smem + dstOff + s * bufStride)),
                         "l"(&m.ref), "r"(cx), "r"(cy), "r"(smem_u32(&bar))
                         : "memory");
        }
    }
    const long long t1 = clock64();
    asm volatile("{ .reg .pred p;\nW_%=: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@!p bra W_%=;\n}" ::"r"(smem_u32(&bar)) : "memory");
    const long long t2 = clock64();
    if (blockIdx.x == 0)
        for (int i = tid; i < ncopies * rows * inner; i += 512) {
            const int s = i / (rows * inner), o = i % (rows * inner);
            out[i] = smem[dstOff + s * bufStride + o];
        }
    if (tid == 0) { cyc[2 * blockIdx.x] = t1 - t0; cyc[2 * blockIdx.x + 1] = t2 - t0; }
}

int main(int argc, char** argv) {
    const int only = argc > 1 ? atoi(argv[1]) : -1;
    setvbuf(stdout, NULL, _IONBF, 0);
    const int pitch = 2080, prow = 1240, rows = 84, inner = 208;
    std::vector<uint8_t> h((size_t)pitch * prow);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)((i * 2654435761u) >> 13);
    uint8_t* d; CK(cudaMalloc(&d, h.size()));
    CK(cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice));
    Maps m;
    cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)prow};
    cuuint64_t strides[1] = {(cuuint64_t)pitch};
    cuuint32_t box[2] = {(cuuint32_t)inner, (cuuint32_t)rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = cuTensorMapEncodeTiled(&m.ref, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); return 3; }
    uint8_t* dOut; long long* dC;
    const int maxCopies = 8;
    CK(cudaMalloc(&dOut, (size_t)maxCopies * rows * inner)); CK(cudaMalloc(&dC, 148 * 16));
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    struct Case { const char* name; int c0, r0, colStep, bufStride, dstOff, grid; } cases[] = {
        {"aligned column 336, col step 16", 336, 100, 16, ((rows * inner + 127) / 128) * 128, 0, 1},
        {"column 336, col step 4", 336, 100, 4, ((rows * inner + 127) / 128) * 128, 0, 1},
        {"column 336, col step 1", 336, 100, 1, ((rows * inner + 127) / 128) * 128, 0, 1},
        {"aligned dst, col step 1", 333, 100, 1, ((rows * inner + 127) / 128) * 128, 0, 1},
        {"dst 16-byte skew per copy (stride = 128k + 16)", 333, 100, 1, ((rows * inner + 127) / 128) * 128 + 16, 0, 1},
        {"negative column, rows past the end (zero fill)", -5, prow - 40, 1, ((rows * inner + 127) / 128) * 128, 0, 1},
        {"all SMs, aligned", 333, 100, 1, ((rows * inner + 127) / 128) * 128, 0, 148},
    };
    std::vector<uint8_t> got((size_t)maxCopies * rows * inner);
    int ci = -1;
    for (auto& cs : cases) {
        if (++ci != only && only >= 0) continue;
        CK(cudaMemset(dOut, 0xEE, got.size()));
        k<<<cs.grid, 512, 200 * 1024>>>(m, cs.c0, cs.r0, rows, inner, maxCopies, cs.colStep, cs.bufStride, cs.dstOff, dOut, dC);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%-50s kernel failed: %s\n", cs.name, cudaGetErrorString(e)); return 4; }
        CK(cudaMemcpy(got.data(), dOut, got.size(), cudaMemcpyDeviceToHost));
        size_t bad = 0;
        for (int s = 0; s < maxCopies; ++s)
            for (int y = 0; y < rows; ++y)
                for (int x = 0; x < inner; ++x) {
                    const int gx = cs.c0 + s * cs.colStep + x, gy = cs.r0 + y;
                    const uint8_t want = (gx >= 0 && gx < pitch && gy >= 0 && gy < prow) ? h[(size_t)gy * pitch + gx] : 0;
                    bad += got[((size_t)s * rows + y) * inner + x] != want;
                }
        long long c[296];
        CK(cudaMemcpy(c, dC, sizeof(long long) * 2 * cs.grid, cudaMemcpyDeviceToHost));
        double a = 0, b = 0; for (int i = 0; i < cs.grid; ++i) { a += c[2 * i]; b += c[2 * i + 1]; }
        printf("%-50s %zu wrong bytes; issue %.0f cycles, complete %.0f cycles after the first instruction (%d copies of %d x %d bytes, %d CTAs)\n", cs.name, bad, a / cs.grid,
               b / cs.grid, maxCopies, rows, inner, cs.grid);
    }
    return 0;
}
