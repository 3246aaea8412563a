// Issue rate of scalar vs packed fp32 on sm_100a: FFMA / FMUL / FADD against FFMA2 / FMUL2 / FADD2
// (fma.rn.f32x2 ...), register and immediate forms.  Prints lane-operations per clock per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_rate f32x2_rate.cu && ./f32x2_rate
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define CHAINS 8
#define ITERS 4096
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float a, float b) {
  float f[CHAINS * 2];
  u64 p[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS * 2; i++) f[i] = threadIdx.x * 1e-3f + i;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(f[2 * i]), "f"(f[2 * i + 1]));
  u64 a2, b2;
  asm("mov.b64 %0, {%1, %1};" : "=l"(a2) : "f"(a));
  asm("mov.b64 %0, {%1, %1};" : "=l"(b2) : "f"(b));
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
      if (MODE == 0) {  // scalar FFMA, three registers (two per pair element)
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[2 * i]) : "f"(a), "f"(b));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[2 * i + 1]) : "f"(a), "f"(b));
      } else if (MODE == 1) {  // scalar FFMA, immediate addend
        asm volatile("fma.rn.f32 %0, %0, %1, 0f3F000000;" : "+f"(f[2 * i]) : "f"(a));
        asm volatile("fma.rn.f32 %0, %0, %1, 0f3F000000;" : "+f"(f[2 * i + 1]) : "f"(a));
      } else if (MODE == 2) {  // FFMA2, three register pairs
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(a2), "l"(b2));
      } else if (MODE == 3) {  // FMUL2
        asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(a2));
      } else if (MODE == 4) {  // FADD2
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(a2));
      } else if (MODE == 5) {  // scalar FMUL
        asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[2 * i]) : "f"(a));
        asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[2 * i + 1]) : "f"(a));
      } else if (MODE == 6) {  // scalar FADD
        asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[2 * i]) : "f"(a));
        asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[2 * i + 1]) : "f"(a));
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i]));
    s += lo + hi + f[2 * i] + f[2 * i + 1];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, float* out, int sms, int khz) {
  const int blocks = sms * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  k<MODE><<<blocks, 256>>>(out, 0.999f, 1e-3f);
  cudaEventRecord(e0);
  k<MODE><<<blocks, 256>>>(out, 0.999f, 1e-3f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double lane_ops = (double)blocks * 256 * ITERS * CHAINS * 2;  // fp32 lane operations
  const double clocks = ms * 1e-3 * khz * 1e3;
  printf("%-28s %8.3f ms  %7.1f fp32 lane-ops / clk / SM (nominal clock %d MHz)\n", name, ms, lane_ops / clocks / sms, khz / 1000);
}
int main() {
  cudaDeviceProp pr;
  cudaGetDeviceProperties(&pr, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  float* out;
  cudaMalloc(&out, (size_t)pr.multiProcessorCount * 8 * 256 * 4);
  run<0>("FFMA  reg,reg,reg", out, pr.multiProcessorCount, khz);
  run<1>("FFMA  reg,reg,imm", out, pr.multiProcessorCount, khz);
  run<2>("FFMA2 reg,reg,reg", out, pr.multiProcessorCount, khz);
  run<5>("FMUL  reg,reg", out, pr.multiProcessorCount, khz);
  run<3>("FMUL2 reg,reg", out, pr.multiProcessorCount, khz);
  run<6>("FADD  reg,reg", out, pr.multiProcessorCount, khz);
  run<4>("FADD2 reg,reg", out, pr.multiProcessorCount, khz);
  return 0;
}
