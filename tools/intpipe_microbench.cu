// Integer-pipe issue-rate micro-benchmark for sm_100a (B200).
//
// the roofline that bounds whole-CTU block matching is the SM integer ALU, and
// MEASURED_PEAKS.json only carries HBM and bf16 numbers.  This program measures, per SM and per
// clock, how many 32-bit lanes of each instruction class the kernel family relies on can retire:
//   VABSDIFF4.U8.ACC (packed 4xu8 SAD + accumulate), IADD3, IMAD, VIMNMX, LEA-style shift-add,
//   mixtures (ALU + FMA pipes together), CREDUX.MIN, SHFL, LDS.32 / LDS.128.
// Output: one JSON object on stdout; lanes/clk/SM derived from clock64() deltas inside the kernel
// (so it is independent of the DVFS state) and, separately, wall-clock ops/s from CUDA events.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o intpipe_microbench intpipe_microbench.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <string>
#include <algorithm>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

constexpr int ILP = 8;        // independent chains per thread
constexpr int INNER = 64;     // unrolled ops per chain per outer iteration

enum Op { OP_VSAD4 = 0, OP_IADD3, OP_IMAD, OP_VIMNMX, OP_LEA, OP_MIX_IADD_IMAD, OP_MIX_VSAD_IMAD,
          OP_MIX_VSAD_IADD, OP_CREDUX, OP_SHFL, OP_LDS32, OP_LDS128, OP_VIADDMNMX, OP_SETP_SEL,
          OP_MIX_SAD_KEY_MIN, OP_VIADDMNMX_U16X2, OP_COUNT };

static const char* kNames[OP_COUNT] = {
  "vabsdiff4_acc", "iadd3", "imad", "vimnmx_u32", "lea_shift_add", "mix_iadd3_imad", "mix_vabsdiff4_imad",
  "mix_vabsdiff4_iadd3", "credux_min", "shfl_xor", "lds32", "lds128", "viaddmnmx_u32", "isetp_sel",
  "mix_sad_imadkey_min", "viaddmnmx_u16x2" };

// lane-ops counted per inner step per chain (a "mix" step issues more than one instruction)
__host__ __device__ constexpr int opsPerStep(int op) {
  return (op == OP_MIX_IADD_IMAD || op == OP_MIX_VSAD_IMAD || op == OP_MIX_VSAD_IADD) ? 2 :
         (op == OP_SETP_SEL) ? 2 : (op == OP_MIX_SAD_KEY_MIN) ? 4 : 1;
}

template <int OP>
__global__ void __launch_bounds__(256) bench(uint32_t* out, unsigned long long* cyc, int iters, uint32_t seed) {
  __shared__ uint32_t sm[2048];
  uint32_t a[ILP], b = seed * 0x9E3779B9u + threadIdx.x, c = seed ^ 0x5bd1e995u;
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = i * seed;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < ILP; ++k) a[k] = threadIdx.x * 31 + k + seed;
  unsigned long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int s = 0; s < INNER; ++s) {
#pragma unroll
      for (int k = 0; k < ILP; ++k) {
        if (OP == OP_VSAD4) {
          asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(a[k]) : "r"(b), "r"(c));
        } else if (OP == OP_IADD3) {
          asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(a[k]) : "r"(b), "r"(c));
        } else if (OP == OP_IMAD) {
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(b), "r"(c));
        } else if (OP == OP_VIMNMX) {
          asm volatile("min.u32 %0, %0, %1;" : "+r"(a[k]) : "r"(b + s));
        } else if (OP == OP_LEA) {
          asm volatile("{ .reg .u32 t; shl.b32 t, %0, 5; add.u32 %0, t, %1; }" : "+r"(a[k]) : "r"(b));
        } else if (OP == OP_MIX_IADD_IMAD) {
          if (k & 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(b), "r"(c));
          else asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(a[k]) : "r"(b), "r"(c));
        } else if (OP == OP_MIX_VSAD_IMAD) {
          if (k & 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(b), "r"(c));
          else asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(a[k]) : "r"(b), "r"(c));
        } else if (OP == OP_MIX_VSAD_IADD) {
          if (k & 1) asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(a[k]) : "r"(b), "r"(c));
          else asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(a[k]) : "r"(b), "r"(c));
        } else if (OP == OP_CREDUX) {
          a[k] = __reduce_min_sync(0xffffffffu, a[k] + s);
        } else if (OP == OP_SHFL) {
          a[k] = __shfl_xor_sync(0xffffffffu, a[k], 1 + (s & 15));
        } else if (OP == OP_LDS32) {
          a[k] = sm[(a[k] + threadIdx.x) & 2047];
        } else if (OP == OP_LDS128) {
          uint4 v = *reinterpret_cast<const uint4*>(&sm[((a[k] + threadIdx.x) & 511) * 4]);
          a[k] = v.x ^ v.y ^ v.z ^ v.w;
        } else if (OP == OP_VIADDMNMX) {
          a[k] = __viaddmin_u32(a[k], b, c + s);
        } else if (OP == OP_VIADDMNMX_U16X2) {
          a[k] = __viaddmin_u16x2(a[k], b, c + s);
        } else if (OP == OP_SETP_SEL) {
          uint32_t v = b + s * 7 + k;
          asm volatile("{ .reg .pred p; setp.lt.u32 p, %1, %0; selp.u32 %0, %1, %0, p; }" : "+r"(a[k]) : "r"(v));
        } else if (OP == OP_MIX_SAD_KEY_MIN) {
          // the kernel's steady-state mixture: 2 packed SADs (ALU?), 1 IMAD key (FMA pipe), 1 min (ALU)
          uint32_t t = a[k];
          asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(t) : "r"(b), "r"(c));
          asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(t) : "r"(c), "r"(b));
          uint32_t key;
          asm volatile("mad.lo.u32 %0, %1, 32768, %2;" : "=r"(key) : "r"(t), "r"(c));
          asm volatile("min.u32 %0, %1, %2;" : "=r"(a[k]) : "r"(key), "r"(t));
        }
      }
    }
  }
  unsigned long long t1 = clock64();
  uint32_t r = 0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) r ^= a[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
static void run(int sms, int ctasPerSm, int iters, uint32_t* dOut, unsigned long long* dCyc, std::string& json) {
  const int threads = 256, grid = sms * ctasPerSm;
  bench<OP><<<grid, threads>>>(dOut, dCyc, 4, 1u);   // warm-up
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  bench<OP><<<grid, threads>>>(dOut, dCyc, iters, 3u);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<unsigned long long> cyc(grid);
  CK(cudaMemcpy(cyc.data(), dCyc, grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  std::sort(cyc.begin(), cyc.end());
  const double laneOpsPerCta = double(threads) * ILP * INNER * double(iters) * opsPerStep(OP);
  // all ctasPerSm CTAs of an SM are co-resident (256 thr, few regs) and run concurrently: lanes/clk/SM
  // = ctasPerSm * laneOpsPerCta / median CTA cycles.
  const double med = double(cyc[grid / 2]);
  const double lanesPerClkSm = ctasPerSm * laneOpsPerCta / med;
  const double opsPerSec = laneOpsPerCta * grid / (ms * 1e-3);
  const double impliedMhz = med / (ms * 1e-3) / 1e6;
  char buf[512];
  snprintf(buf, sizeof buf, "%s\"%s\": {\"lanes_per_clk_sm\": %.2f, \"lane_ops_per_s\": %.4e, \"ms\": %.3f, \"median_cta_cycles\": %.0f, \"implied_sm_mhz\": %.0f}",
           json.empty() ? "" : ", ", kNames[OP], lanesPerClkSm, opsPerSec, ms, med, impliedMhz);
  json += buf;
}

int main(int argc, char** argv) {
  int iters = argc > 1 ? atoi(argv[1]) : 2000;
  int ctasPerSm = argc > 2 ? atoi(argv[2]) : 4;
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  uint32_t* dOut; unsigned long long* dCyc;
  CK(cudaMalloc(&dOut, size_t(sms) * ctasPerSm * 256 * 4));
  CK(cudaMalloc(&dCyc, size_t(sms) * ctasPerSm * 8));
  std::string js;
  run<OP_VSAD4>(sms, ctasPerSm, iters, dOut, dCyc, js);
  run<OP_IADD3>(sms, ctasPerSm, iters, dOut, dCyc, js);
  run<OP_IMAD>(sms, ctasPerSm, iters, dOut, dCyc, js);
  run<OP_VIMNMX>(sms, ctasPerSm, iters, dOut, dCyc, js);
  run<OP_LEA>(sms, ctasPerSm, iters, dOut, dCyc, js);
  run<OP_MIX_IADD_IMAD>(sms, ctasPerSm, iters, dOut, dCyc, js);
  run<OP_MIX_VSAD_IMAD>(sms, ctasPerSm, iters, dOut, dCyc, js);
  run<OP_MIX_VSAD_IADD>(sms, ctasPerSm, iters, dOut, dCyc, js);
  run<OP_CREDUX>(sms, ctasPerSm, iters / 4, dOut, dCyc, js);
  run<OP_SHFL>(sms, ctasPerSm, iters / 4, dOut, dCyc, js);
  run<OP_LDS32>(sms, ctasPerSm, iters / 4, dOut, dCyc, js);
  run<OP_LDS128>(sms, ctasPerSm, iters / 4, dOut, dCyc, js);
  run<OP_VIADDMNMX>(sms, ctasPerSm, iters, dOut, dCyc, js);
  run<OP_VIADDMNMX_U16X2>(sms, ctasPerSm, iters, dOut, dCyc, js);
  run<OP_SETP_SEL>(sms, ctasPerSm, iters, dOut, dCyc, js);
  run<OP_MIX_SAD_KEY_MIN>(sms, ctasPerSm, iters, dOut, dCyc, js);
  printf("{\"device\": \"%s\", \"sms\": %d, \"ctas_per_sm\": %d, \"threads\": 256, \"ilp\": %d, \"clock_khz_max\": %d, %s}\n",
         p.name, sms, ctasPerSm, ILP, p.clockRate, js.c_str());
  return 0;
}
