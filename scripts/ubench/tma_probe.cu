// tma_probe.cu -- which tensor-map / box configurations does the TMA unit accept?  (bring-up probe, not product code)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../transmf_ad_b200/csrc/umma.cuh"
using namespace tmf::umma;

struct alignas(64) Args { CUtensorMap tm; int rank; int c[5]; uint32_t bytes; };

__global__ void probe(const __grid_constant__ Args a, uint8_t* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar = base + 65536;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
    mbar_expect_tx(bar, a.bytes);
    if (a.rank == 5) tma_load_5d(base, &a.tm, bar, a.c[0], a.c[1], a.c[2], a.c[3], a.c[4]);
    else if (a.rank == 4) tma_load_4d(base, &a.tm, bar, a.c[0], a.c[1], a.c[2], a.c[3]);
    else if (a.rank == 3) tma_load_3d(base, &a.tm, bar, a.c[0], a.c[1], a.c[2]);
    else tma_load_2d(base, &a.tm, bar, a.c[0], a.c[1]);
    mbar_wait(bar, 0);
    for (uint32_t i = 0; i < a.bytes; ++i) out[i] = gen[i];
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int run(const char* name, int rank, std::vector<uint64_t> dims, std::vector<uint32_t> box, std::vector<int> coord,
        CUtensorMapSwizzle swz, EncodeTiledFn enc, void* dptr, uint8_t* out) {
  Args a{};
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5] = {1, 1, 1, 1, 1};
  uint64_t stride = 2;
  uint32_t bytes = 2;
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i]; bx[i] = box[i]; bytes *= box[i];
    stride *= dims[i];
    if (i < rank - 1) gs[i] = stride;
    a.c[i] = coord[i];
  }
  CUresult r = enc(&a.tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, dptr, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("%-40s encode failed %d\n", name, (int)r); return 1; }
  a.rank = rank; a.bytes = bytes;
  probe<<<1, 32, 1024 + 65536 + 64>>>(a, out);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%-40s bytes %6u -> %s\n", name, bytes, cudaGetErrorString(e));
  return e != cudaSuccess;
}

int main(int argc, char** argv) {
  int which = argc > 1 ? atoi(argv[1]) : 0;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)sym;
  void* d; uint8_t* out;
  cudaMalloc(&d, 64 << 20); cudaMemset(d, 0x11, 64 << 20);
  cudaMalloc(&out, 1 << 20);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 + 65536 + 64);
  // each case in its own process (a faulting TMA poisons the context): the shell loop passes `which`
  switch (which) {
    case 0: return run("4d x dims{24,19,17,2} box{32,4,4,1} sw64", 4, {24, 19, 17, 2}, {32, 4, 4, 1}, {0, -1, -1, 0}, CU_TENSOR_MAP_SWIZZLE_64B, enc, d, out);
    case 1: return run("4d x dims{96,109,91,8} box{32,4,4,1} sw64", 4, {96, 109, 91, 8}, {32, 4, 4, 1}, {1, -1, -1, 0}, CU_TENSOR_MAP_SWIZZLE_64B, enc, d, out);
    case 2: return run("4d x dims{96,..} box{32,4,4,1} c0=0", 4, {96, 109, 91, 8}, {32, 4, 4, 1}, {0, 3, 3, 0}, CU_TENSOR_MAP_SWIZZLE_64B, enc, d, out);
    case 3: return run("5d y dims{32,18,19,17,2} box{32,32,2,2,1} sw64", 5, {32, 18, 19, 17, 2}, {32, 32, 2, 2, 1}, {0, 0, 2, 2, 0}, CU_TENSOR_MAP_SWIZZLE_64B, enc, d, out);
    case 4: return run("5d g dims{32,9,9,8,2} box{32,9,1,1,1} none", 5, {32, 9, 9, 8, 2}, {32, 9, 1, 1, 1}, {0, 0, 1, 1, 0}, CU_TENSOR_MAP_SWIZZLE_NONE, enc, d, out);
    case 5: return run("4d x box{32,4,4,1} c0=32 (aligned)", 4, {96, 109, 91, 8}, {32, 4, 4, 1}, {32, 3, 3, 0}, CU_TENSOR_MAP_SWIZZLE_64B, enc, d, out);
    case 6: return run("4d x box{32,4,4,1} c0=8", 4, {96, 109, 91, 8}, {32, 4, 4, 1}, {8, 3, 3, 0}, CU_TENSOR_MAP_SWIZZLE_64B, enc, d, out);
    case 7: return run("4d x box{32,4,4,1} no swizzle c0=1", 4, {96, 109, 91, 8}, {32, 4, 4, 1}, {1, 3, 3, 0}, CU_TENSOR_MAP_SWIZZLE_NONE, enc, d, out);
    case 8: return run("5d g dims{32,45,54,45,8} box{32,46,1,1,1}", 5, {32, 45, 54, 45, 8}, {32, 46, 1, 1, 1}, {0, 0, 1, 1, 0}, CU_TENSOR_MAP_SWIZZLE_NONE, enc, d, out);
    case 9: return run("5d y dims{32,91,..} box{32,96,2,2,1} sw64", 5, {32, 91, 109, 91, 2}, {32, 96, 2, 2, 1}, {0, 0, 2, 2, 0}, CU_TENSOR_MAP_SWIZZLE_64B, enc, d, out);
  }
  return 0;
}
