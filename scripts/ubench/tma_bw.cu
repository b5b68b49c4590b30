// tma_bw.cu -- TMA tensor-load throughput per SM for the box shapes the conv kernels use (bring-up probe, not product).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_bw tma_bw.cu -lcuda && ./tma_bw
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include "../../transmf_ad_b200/csrc/umma.cuh"
using namespace tmf::umma;

struct alignas(64) Args {
  CUtensorMap tm;
  int rank, nslots, iters;
  uint32_t bytes, slot_bytes;
  int c1, c2;          // start coordinates in dims 1, 2
  int n3, n4;          // extent to cycle through in dims 3, 4 (planes, samples); for rank 2: n3 = number of row blocks
  int rows_per_box;
  int nl;              // issuing lanes per warp
};

__global__ void __launch_bounds__(256, 1) bw_kernel(const __grid_constant__ Args a, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int nw = (blockDim.x >> 5) * a.nl, wid = (threadIdx.x >> 5) * a.nl + (threadIdx.x & 31);   // issuer index
  const uint32_t bars_all = base + (uint32_t)(a.nslots * nw) * a.slot_bytes;
  if (threadIdx.x == 0) {
    for (int i = 0; i < a.nslots * nw; ++i) mbar_init(bars_all + 8 * i, 1);
    fence_barrier_init();
  }
  __syncthreads();
  const uint32_t bars = bars_all + 8 * a.nslots * wid;
  if ((threadIdx.x & 31) < a.nl) {
    const long long t0 = clock64();
    int issued = 0, done = 0;
    uint32_t ph = 0;
    int slot_w = 0;
    while (done < a.iters) {
      while (issued < a.iters && issued - done < a.nslots) {
        const int s = issued % a.nslots;
        const int k = (issued * nw + wid) * gridDim.x + blockIdx.x;  // distinct box per (CTA, warp, iteration)
        mbar_expect_tx(bars + 8 * s, a.bytes);
        const uint32_t dst = base + (uint32_t)(wid * a.nslots + s) * a.slot_bytes;
        if (a.rank == 5) tma_load_5d(dst, &a.tm, bars + 8 * s, 0, a.c1, a.c2, k % a.n3, (k / a.n3) % a.n4);
        else if (a.rank == 3) tma_load_3d(dst, &a.tm, bars + 8 * s, 0, 0, k % a.n3);
        else tma_load_2d(dst, &a.tm, bars + 8 * s, 0, (k % a.n3) * a.rows_per_box);
        ++issued;
      }
      mbar_wait(bars + 8 * slot_w, ph);
      if (++slot_w == a.nslots) { slot_w = 0; ph ^= 1u; }
      ++done;
    }
    if (wid == 0) cycles[blockIdx.x] = clock64() - t0;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn enc;
static void* dptr;
static long long* dcyc;

static void run(const char* name, int rank, std::vector<uint64_t> dims, std::vector<uint32_t> box, int c1, int c2, int n3,
                int n4, CUtensorMapSwizzle swz, int nslots, int iters, CUtensorMapL2promotion prom = CU_TENSOR_MAP_L2_PROMOTION_L2_128B, int nw = 1, int nl = 1) {
  Args a{};
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5] = {1, 1, 1, 1, 1};
  uint64_t stride = 2;
  uint32_t bytes = 2;
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i]; bx[i] = box[i];
    stride *= dims[i];
    if (i < rank - 1) gs[i] = stride;
    bytes *= box[i];
  }
  CUresult r = enc(&a.tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, dptr, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, prom,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("%-44s encode failed %d\n", name, (int)r); return; }
  a.nl = nl; a.rank = rank; a.nslots = nslots; a.iters = iters; a.bytes = bytes; a.slot_bytes = (bytes + 1023u) & ~1023u;
  a.c1 = c1; a.c2 = c2; a.n3 = n3; a.n4 = n4; a.rows_per_box = rank == 2 ? box[1] : 0;
  const size_t smem = 1024 + (size_t)nslots * nw * nl * a.slot_bytes + 8 * nslots * nw * nl + 64;
  if (smem > 227 * 1024) { printf("%-44s smem too large\n", name); return; }
  cudaFuncSetAttribute(bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  for (int rep = 0; rep < 2; ++rep) {
    bw_kernel<<<148, 32 * nw, smem>>>(a, dcyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-44s failed: %s\n", name, cudaGetErrorString(e)); return; }
  }
  std::vector<long long> h(148);
  cudaMemcpy(h.data(), dcyc, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (auto v : h) avg += (double)v;
  avg /= 148;
  uint32_t rows = bytes / (box[0] * 2);
  printf("%-44s box %6u B (%4u rows x %3u B) warps %d lanes %d slots %d: %7.0f cyc/box  %6.2f B/clk/SM\n", name, bytes, rows,
         box[0] * 2, nw, nl, nslots, avg / iters / (nw * nl), (double)bytes * iters * nw * nl / avg);
}

int main() {
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
  enc = (EncodeTiledFn)sym;
  cudaMalloc(&dptr, (size_t)256 << 20);
  cudaMemset(dptr, 0, (size_t)256 << 20);
  cudaMalloc(&dcyc, 148 * sizeof(long long));
  const int IT = 64;
  const uint64_t W = 45, H = 54, D = 45, B = 8;
  const auto P = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
  for (int nl : {1, 2, 4})
    run("5D c32 {32,46,6} SW64", 5, {32, W, H, D, B}, {32, 46, 6, 1, 1}, -1, 20, 45, 8, CU_TENSOR_MAP_SWIZZLE_64B, 2, IT, P, 1, nl);
  for (int nl : {1, 2, 4, 8})
    run("2D c64 {64,64} SW128 8KB", 2, {64, W * H * D * B}, {64, 64}, 0, 0, 12000, 1, CU_TENSOR_MAP_SWIZZLE_128B, 2, IT, P, 1, nl);
  return 0;
}
