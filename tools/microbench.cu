// Issue-rate micro-benchmarks for the pipes the sweep kernels load (SURVEY.md
// section 8d asks for measured FP32 / INT / XU peaks instead of nominal ones).
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/microbench tools/microbench.cu
// Prints one JSON line per instruction mix: lane-ops/s and ops/clk/SM at the
// SM clock sampled from the device (clock64 deltas vs. wall time).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

#define ITERS 4096
#define ILP 8

template <int KIND>
__global__ void __launch_bounds__(256) bench(float* out, unsigned long long* cycles, float a, float b, int iters) {
  float x[ILP];
  unsigned long long xx[ILP];
  unsigned int u[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) {
    x[i] = a + threadIdx.x * 1e-3f + i;
    xx[i] = ((unsigned long long)__float_as_uint(x[i]) << 32) | __float_as_uint(x[i] + 1.f);
    u[i] = threadIdx.x * 2654435761u + i;
  }
  unsigned long long bb = ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(b);
  unsigned int p = threadIdx.x & 1;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) {
        if (KIND == 0) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(b));
        if (KIND == 1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(b), "f"(a));
        if (KIND == 2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(xx[i]) : "l"(bb));
        if (KIND == 3) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(xx[i]) : "l"(bb));
        if (KIND == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(p), "r"(u[(i + 1) % ILP]));
        if (KIND == 5) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(p | 3u), "r"(u[(i + 1) % ILP]));
        if (KIND == 6) asm volatile("{.reg .pred q; setp.ne.u32 q, %2, 0; selp.f32 %0, %0, %1, q;}" : "+f"(x[i]) : "f"(b), "r"(u[i] & 1));
        if (KIND == 7) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
        if (KIND == 8) { asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(b));
                         asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(p), "r"(u[(i + 1) % ILP])); }
        if (KIND == 9) { asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(xx[i]) : "l"(bb));
                         asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(p), "r"(u[(i + 1) % ILP])); }
        if (KIND == 10) { unsigned int hi; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(hi) : "r"(u[i]), "r"(0xD2511F53u)); u[i] ^= hi; }
        if (KIND == 11) asm volatile("add.f64 %0, %0, %1;" : "+d"(*(double*)&xx[i]) : "d"(1.0));
      }
    }
  }
  long long t1 = clock64();
  float acc = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc += x[i] + __uint_as_float((unsigned)xx[i]) + __uint_as_float(u[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = (unsigned long long)(t1 - t0);
}

template <int KIND>
void run(const char* name, double ops_per_instr, int instr_per_slot) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int blocks = sms * 8, threads = 256;
  float* out; unsigned long long* cyc;
  cudaMalloc(&out, blocks * threads * sizeof(float)); cudaMalloc(&cyc, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  bench<KIND><<<blocks, threads>>>(out, cyc, 1.0001f, 0.9999f, 64);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    bench<KIND><<<blocks, threads>>>(out, cyc, 1.0001f, 0.9999f, ITERS);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  unsigned long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double instr = (double)blocks * threads * ITERS * 8.0 * ILP * instr_per_slot;
  const double lane_instr_per_s = instr / (best * 1e-3);
  printf("{\"mix\": \"%s\", \"ms\": %.3f, \"lane_instr_per_s\": %.4e, \"lane_ops_per_s\": %.4e, "
         "\"warp_instr_per_clk_per_sm_at_1965MHz\": %.3f, \"block0_cycles\": %llu}\n",
         name, best, lane_instr_per_s, lane_instr_per_s * ops_per_instr / instr_per_slot * 1.0,
         lane_instr_per_s / 32.0 / sms / 1.965e9, c);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("FMUL", 1, 1);
  run<1>("FFMA", 1, 1);
  run<2>("FMUL2", 2, 1);
  run<3>("FFMA2", 2, 1);
  run<4>("LOP3", 1, 1);
  run<5>("IMAD", 1, 1);
  run<6>("SETP+FSEL", 1, 2);
  run<7>("MUFU.EX2", 1, 1);
  run<8>("FMUL+LOP3", 1, 2);
  run<9>("FMUL2+LOP3", 1.5, 2);
  run<10>("IMAD.HI+LOP3", 1, 2);
  run<11>("DADD", 1, 1);
  return 0;
}
