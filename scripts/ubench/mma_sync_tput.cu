// Throughput / latency of legacy mma.sync on sm_100a: m16n8k8 tf32 and m16n8k16 bf16, NACC independent accumulators per warp.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
template <int KIND, int NACC>
__global__ void k(float* out, int iters) {
  float c[NACC][4];
  for (int j = 0; j < NACC; ++j) c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.f;
  uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 9u}, b[2] = {threadIdx.x * 5u, 11u};
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < NACC; ++j) {
      if (KIND == 0)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
      else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
  }
  float s = 0.f;
  for (int j = 0; j < NACC; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int KIND, int NACC>
void run(const char* name, int warps) {
  float* out; cudaMalloc(&out, 148 * 1024 * 4);
  const int iters = 4096;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<KIND, NACC><<<148, warps * 32>>>(out, 16);
  cudaEventRecord(e0);
  k<KIND, NACC><<<148, warps * 32>>>(out, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double mmas = (double)iters * NACC * warps;            // per SM
  const double cyc = ms * 1e-3 * 1.965e9;
  const double macs = (KIND == 0 ? 1024.0 : 2048.0);
  printf("%s warps/SM=%2d nacc=%d: %.1f cycles per mma per warp-chain step, %.2f mma/clk/SM, %.0f MAC/clk/SM, %.1f TFLOP/s chip\n", name, warps, NACC,
         cyc / iters / NACC * 1.0, mmas / cyc, mmas * macs / cyc, mmas * macs * 2 * 148 / (ms * 1e-3) / 1e12);
  cudaFree(out);
}
int main() {
  run<0, 1>("tf32 m16n8k8 ", 1); run<0, 1>("tf32 m16n8k8 ", 4); run<0, 2>("tf32 m16n8k8 ", 8); run<0, 4>("tf32 m16n8k8 ", 8); run<0, 8>("tf32 m16n8k8 ", 16);
  run<1, 1>("bf16 m16n8k16", 1); run<1, 1>("bf16 m16n8k16", 4); run<1, 2>("bf16 m16n8k16", 8); run<1, 4>("bf16 m16n8k16", 8); run<1, 8>("bf16 m16n8k16", 16);
  return 0;
}
