// umma_tput.cu -- cycles per tcgen05.mma (M=128, K=16, bf16, SS mode) as a function of N, smem layout (SW64 / SW128 K-major),
// A row offset alignment and operand reuse.  Operands are whatever is in shared memory (zeros).  Bring-up probe.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_tput umma_tput.cu && ./umma_tput
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include "../../transmf_ad_b200/csrc/umma.cuh"
using namespace tmf::umma;

// mode: 0 = same A and B every MMA; 1 = A walks through a 64 KB region (row shifts), B walks through 27 taps
__global__ void __launch_bounds__(128, 1) tput(int N, int sw128, int a_row_off, int mode, int iters, int ksteps, uint32_t idesc, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(gen)[i] = 0;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(smem_u32(&tmem_slot), 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t pitch = sw128 ? 128u : 64u;
    const uint32_t layout = sw128 ? LAYOUT_SW128 : LAYOUT_SW64;
    const uint32_t smA = base, smB = base + 96 * 1024;
    const uint64_t hi = make_smem_desc(0, 16, 8 * pitch, layout, 0) & 0xFFFFFFFF00000000ull;
    const uint32_t lo_c = (uint32_t)(make_smem_desc(0, 16, 8 * pitch, layout, 0) & 0xFFFF0000ull);
    const uint32_t a0 = lo_c | (((smA + a_row_off * pitch) & 0x3FFFFu) >> 4);
    const uint32_t b0 = lo_c | ((smB & 0x3FFFFu) >> 4);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      uint32_t a = a0, b = b0;
      if (mode == 1) {
        a += ((uint32_t)((i % 9) * 47 * pitch)) >> 4;             // row shifts inside a slab
        { int nb = (int)((96u * 1024u) / ((uint32_t)N * pitch)); nb = nb > 27 ? 27 : (nb < 1 ? 1 : nb); b += ((uint32_t)((i % nb) * N) * pitch) >> 4; }
      }
      for (int k = 0; k < ksteps; ++k) mma_bf16_ss(tm, hi | (uint64_t)(a + 2u * k), hi | (uint64_t)(b + 2u * k), idesc, 1u);
    }
    mma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    out[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 8);
  { cudaError_t e = cudaFuncSetAttribute(tput, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024); if (e != cudaSuccess) printf("attr: %s\n", cudaGetErrorString(e)); }
  const int iters = 2000;
  printf("cycles per MMA (M=128, K=16), 148 CTAs concurrently\n");
  for (int sw128 = 0; sw128 <= 1; ++sw128)
    for (int mode = 0; mode <= 1; ++mode)
      for (int off : {0, 3}) {
        printf("%s mode %d A row offset %d:", sw128 ? "SW128" : "SW64 ", mode, off);
        for (int N : {32, 64, 96, 128, 192, 256}) {
          const int ks = sw128 ? 4 : 2;
          tput<<<148, 128, 201 * 1024 + 2048>>>(N, sw128, off, mode, iters, ks, make_idesc_bf16(128, N, 0, 0), d);
          cudaError_t e = cudaGetLastError();
          if (e == cudaSuccess) e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf(" N=%d: %s", N, cudaGetErrorString(e)); break; }
          long long h[148];
          cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
          double avg = 0;
          for (int i = 0; i < 148; ++i) avg += (double)h[i];
          printf("  N=%3d: %6.1f", N, avg / 148 / iters / ks);
        }
        printf("\n");
      }
  return 0;
}
