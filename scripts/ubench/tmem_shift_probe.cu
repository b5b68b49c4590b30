// tmem_shift_probe.cu -- what does tcgen05.shift.down do?  Fill TMEM with value = lane*1000 + column, shift, read back.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tmem_shift_probe tmem_shift_probe.cu && ./tmem_shift_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(float* out, int col_of_shift, int lane_of_shift, int nshift) {
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_slot;
  const uint32_t taddr = base + ((uint32_t)(warp * 32) << 16);
  uint32_t v[32];
  for (int j = 0; j < 32; ++j) v[j] = __float_as_uint((float)(threadIdx.x * 1000 + j));
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
        "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
        "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x == 0) {
    const uint32_t sa = base + ((uint32_t)lane_of_shift << 16) + (uint32_t)col_of_shift;
    for (int i = 0; i < nshift; ++i) asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(sa) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  {
    uint32_t ok = 0;
    while (!ok) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.b32 %0, 1, 0, p;\n\t}"
          : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int j = 0; j < 32; ++j) out[threadIdx.x * 32 + j] = __uint_as_float(r[j]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(base) : "memory");
}

static void run(int col, int lane, int nshift) {
  float* d; cudaMalloc(&d, 128 * 32 * 4);
  probe<<<1, 128>>>(d, col, lane, nshift);
  cudaError_t e = cudaDeviceSynchronize();
  printf("== shift at col %d lane %d x%d: %s\n", col, lane, nshift, cudaGetErrorString(e));
  if (e != cudaSuccess) return;
  static float h[128 * 32];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  // for each column: which source lane does lane L now hold (value/1000), report as delta pattern
  for (int c = 0; c < 32; ++c) {
    int changed = 0, d1 = 0, dm1 = 0, other = 0;
    for (int L = 0; L < 128; ++L) {
      const int src = (int)(h[L * 32 + c] / 1000.f + 0.001f), cc = (int)(h[L * 32 + c]) % 1000;
      if (src != L || cc != c) { ++changed; if (src == L + 1) ++d1; else if (src == L - 1) ++dm1; else ++other; }
    }
    if (changed) printf("  col %2d: changed %3d lanes (holds lane+1: %d, lane-1: %d, other: %d)  lane0=%g lane1=%g lane31=%g lane32=%g lane33=%g lane126=%g lane127=%g\n",
                        c, changed, d1, dm1, other, h[c], h[32 + c], h[31 * 32 + c], h[32 * 32 + c], h[33 * 32 + c], h[126 * 32 + c], h[127 * 32 + c]);
  }
  cudaFree(d);
}

int main() {
  run(0, 0, 1);
  run(8, 0, 1);
  run(4, 0, 1);
  run(0, 32, 1);
  run(0, 0, 2);
  return 0;
}
