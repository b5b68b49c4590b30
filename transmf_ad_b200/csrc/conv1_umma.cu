// conv1_umma.cu -- conv1.0 (Cin = 1, 3x3x3, pad 1; reference models/networks.py:22) on the tcgen05 tensor cores.
//
// The layer is only 26 FLOP/byte, but on CUDA cores it is FMA-issue bound (864 FMAs per voxel).  Here the 27-tap
// neighbourhood of every voxel is written by producer warps as one K = 32 row of an im2col tile in shared memory
// (K-major, 64-byte rows, 64B swizzle -- the layout a TMA box would have produced), and a [128 voxels x Cout x 32]
// GEMM runs on tcgen05 with the accumulator in TMEM.  fp32 fidelity is kept with a bf16 hi/lo split of BOTH operands:
//     x*w ~= xh*wh + xl*wh + xh*wl        (dropped term xl*wl ~ 2^-16 relative)
// i.e. three MMAs per K step; the image is never rounded to bf16 (that costs ~10x logit error).
// Epilogue as in conv_umma.cu: +bias, round to bf16, 16-byte stores, BatchNorm sum / sum-of-squares of the stored values.
#include <cuda.h>

#include "common.cuh"
#include "umma.cuh"

namespace tmf {
using namespace umma;

constexpr int C1U_PRODUCERS = 128;                 // warps 0-3: one im2col row (voxel) per thread
constexpr int C1U_THREADS = 288;                   // + warp 4: MMA issuer / TMEM owner, warps 5-8: epilogue
constexpr int C1U_STAGES = 4;
constexpr uint32_t C1U_TILE_BYTES = 128 * 64;      // one [128 x 32] bf16 operand tile

struct C1UParams {
  const float* x[TMF_MAX_GROUPS];
  const float* w[TMF_MAX_GROUPS];
  const float* bias[TMF_MAX_GROUPS];
  __nv_bfloat16* y[TMF_MAX_GROUPS];
  double* stats[TMF_MAX_GROUPS];
  int ng, B, D, H, W, cout;
  long long M;
  int ntiles;
  uint32_t idesc, tmem_cols;
};

// byte offset of 16-byte chunk j of row r inside a K-major, 64-byte-row, 64B-swizzled tile (tile base 1024-aligned)
__device__ __forceinline__ uint32_t sw64_off(int r, int j) { return (uint32_t)r * 64u + (uint32_t)((j ^ ((r >> 1) & 3)) << 4); }

__device__ __forceinline__ void split_bf16(float v, float& hi, float& lo) {
  hi = round_bf16(v);
  lo = v - hi;
}

__global__ void __launch_bounds__(C1U_THREADS, 1) conv1_umma_fwd_kernel(const __grid_constant__ C1UParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (smem_base - smem_u32(smem_raw));
  // layout: [stages][A_hi | A_lo] (16 KB each) | W_hi | W_lo (cout x 64 B each, 4 KB reserved each) | barriers | stats
  const uint32_t smA = smem_base;
  const uint32_t smW = smA + C1U_STAGES * 2 * C1U_TILE_BYTES;
  const uint32_t bars = smW + 2 * 4096;
  const uint32_t full = bars, empty = bars + 8 * C1U_STAGES, acc_full = empty + 8 * C1U_STAGES, acc_empty = acc_full + 16;
  const uint32_t tmem_slot = acc_empty + 16;
  float* stats_ptr = reinterpret_cast<float*>(gen + (tmem_slot + 16 - smem_base));        // float [2][64]
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.x % p.ng;
  const int cta = blockIdx.x / p.ng, ncta = gridDim.x / p.ng;

  if (threadIdx.x == 0) {
    for (int i = 0; i < C1U_STAGES; ++i) { mbar_init(full + 8 * i, C1U_PRODUCERS); mbar_init(empty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full + 8 * i, 1); mbar_init(acc_empty + 8 * i, 128); }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 128; i += C1U_THREADS) stats_ptr[i] = 0.f;
  // weights: fp32 (Cout,27) -> bf16 hi / lo tiles, K-major rows of 32 taps (taps 27..31 zero)
  for (int i = threadIdx.x; i < p.cout * 4; i += C1U_THREADS) {
    const int co = i >> 2, j = i & 3;
    float hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int t = j * 8 + e;
      const float v = (t < 27) ? p.w[g][co * 27 + t] : 0.f;
      split_bf16(v, hi[e], lo[e]);
    }
    *reinterpret_cast<uint4*>(gen + (smW - smem_base) + sw64_off(co, j)) = pack8(hi);
    *reinterpret_cast<uint4*>(gen + (smW + 4096 - smem_base) + sw64_off(co, j)) = pack8(lo);
  }
  fence_proxy_async();
  if (warp == 4) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp < 4) {
    // ======================================= im2col producers =======================================
    const int r = threadIdx.x;                       // row of the tile
    const float* xg = p.x[g];
    int s = 0;
    uint32_t ph = 0;
    for (int tile = cta; tile < p.ntiles; tile += ncta) {
      const long long m = (long long)tile * 128 + r;
      float in[32];
#pragma unroll
      for (int t = 27; t < 32; ++t) in[t] = 0.f;
      if (m < p.M) {
        const int wq = (int)(m % p.W);
        const int hq = (int)((m / p.W) % p.H);
        const int dq = (int)((m / ((long long)p.W * p.H)) % p.D);
        const float* xp = xg + m;
#pragma unroll
        for (int kd = 0; kd < 3; ++kd)
#pragma unroll
          for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              const int dd = dq + kd - 1, hh = hq + kh - 1, ww = wq + kw - 1;
              const bool ok = dd >= 0 && dd < p.D && hh >= 0 && hh < p.H && ww >= 0 && ww < p.W;
              in[(kd * 3 + kh) * 3 + kw] = ok ? __ldg(xp + ((long long)(kd - 1) * p.H + (kh - 1)) * p.W + (kw - 1)) : 0.f;
            }
      } else {
#pragma unroll
        for (int t = 0; t < 27; ++t) in[t] = 0.f;
      }
      mbar_wait(empty + 8 * s, ph ^ 1u);
      uint8_t* a_hi = gen + (smA - smem_base) + (size_t)s * 2 * C1U_TILE_BYTES;
      uint8_t* a_lo = a_hi + C1U_TILE_BYTES;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float hi[8], lo[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) split_bf16(in[j * 8 + e], hi[e], lo[e]);
        const uint32_t off = sw64_off(r, j);
        *reinterpret_cast<uint4*>(a_hi + off) = pack8(hi);
        *reinterpret_cast<uint4*>(a_lo + off) = pack8(lo);
      }
      fence_proxy_async();                           // generic-proxy smem writes -> visible to the tensor core
      mbar_arrive(full + 8 * s);
      if (++s == C1U_STAGES) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 4) {
    // ======================================= MMA issuer =============================================
    int s = 0;
    uint32_t ph = 0;
    const uint64_t desc_hi = make_smem_desc(0, 16, 512, LAYOUT_SW64, 0) & 0xFFFFFFFF00000000ull;
    const uint32_t lo_const = (uint32_t)(make_smem_desc(0, 16, 512, LAYOUT_SW64, 0) & 0xFFFF0000ull);
    const uint32_t a0 = lo_const | ((smA & 0x3FFFFu) >> 4);
    const uint32_t wh = lo_const | ((smW & 0x3FFFFu) >> 4);
    const uint32_t wl = lo_const | (((smW + 4096) & 0x3FFFFu) >> 4);
    int it = 0;
    for (int tile = cta; tile < p.ntiles; tile += ncta, ++it) {
      const int as = it & 1;
      const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
      mbar_wait(acc_empty + 8 * as, acc_ph ^ 1u);
      mbar_wait(full + 8 * s, ph);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(as * p.cout);
      const uint32_t ah = a0 + (uint32_t)s * (2 * C1U_TILE_BYTES >> 4);
      const uint32_t al = ah + (C1U_TILE_BYTES >> 4);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          mma_bf16_ss(d_tmem, desc_hi | (uint64_t)(ah + 2u * k), desc_hi | (uint64_t)(wh + 2u * k), p.idesc, k ? 1u : 0u);
          mma_bf16_ss(d_tmem, desc_hi | (uint64_t)(al + 2u * k), desc_hi | (uint64_t)(wh + 2u * k), p.idesc, 1u);
          mma_bf16_ss(d_tmem, desc_hi | (uint64_t)(ah + 2u * k), desc_hi | (uint64_t)(wl + 2u * k), p.idesc, 1u);
        }
        mma_commit(empty + 8 * s);
        mma_commit(acc_full + 8 * as);
      }
      __syncwarp();
      if (++s == C1U_STAGES) { s = 0; ph ^= 1u; }
    }
  } else {
    // ======================================= epilogue ===============================================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const float* bias = p.bias[g];
    __nv_bfloat16* yg = p.y[g];
    const bool want_stats = p.stats[g] != nullptr;
    float ssum[2] = {0.f, 0.f}, ssq[2] = {0.f, 0.f};
    int it = 0;
    for (int tile = cta; tile < p.ntiles; tile += ncta, ++it) {
      const long long m = (long long)tile * 128 + row;
      const bool valid = m < p.M;
      const int as = it & 1;
      const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
      mbar_wait(acc_full + 8 * as, acc_ph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * p.cout);
      for (int c0 = 0; c0 < p.cout; c0 += 32) {
        uint32_t raw[32];
        tmem_ld32(taddr + (uint32_t)c0, raw);
        tmem_ld_wait();
        float v[32], sq[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float a = __uint_as_float(raw[j]);
          if (bias != nullptr) a += __ldg(bias + c0 + j);
          v[j] = valid ? round_bf16(a) : 0.f;
          sq[j] = v[j] * v[j];
        }
        if (valid) {
          __nv_bfloat16* yrow = yg + m * p.cout + c0;
#pragma unroll
          for (int qd = 0; qd < 4; ++qd) *reinterpret_cast<uint4*>(yrow + qd * 8) = pack8(&v[qd * 8]);
        }
        if (want_stats) {
#pragma unroll
          for (int off = 16; off >= 1; off >>= 1) {
            const bool upper = (lane & off) != 0;
#pragma unroll
            for (int i = 0; i < off; ++i) {
              const float s_send = upper ? v[i] : v[i + off];
              const float s_keep = upper ? v[i + off] : v[i];
              v[i] = s_keep + __shfl_xor_sync(0xffffffffu, s_send, off);
              const float q_send = upper ? sq[i] : sq[i + off];
              const float q_keep = upper ? sq[i + off] : sq[i];
              sq[i] = q_keep + __shfl_xor_sync(0xffffffffu, q_send, off);
            }
          }
          ssum[c0 >> 5] += v[0];
          ssq[c0 >> 5] += sq[0];
        }
      }
      tc_fence_before();
      mbar_arrive(acc_empty + 8 * as);
    }
    if (want_stats) {
      for (int i = 0; i * 32 < p.cout; ++i) {
        atomicAdd(&stats_ptr[i * 32 + lane], ssum[i]);
        atomicAdd(&stats_ptr[64 + i * 32 + lane], ssq[i]);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int e = threadIdx.x - 160;             // 0..127 over the four epilogue warps
      if (e < p.cout) {
        atomicAdd(&p.stats[g][e], (double)stats_ptr[e]);
        atomicAdd(&p.stats[g][p.cout + e], (double)stats_ptr[64 + e]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

}  // namespace tmf

using namespace tmf;

bool tmf_conv1_fwd_umma_supported(int cout) {
  if (getenv("TMF_DISABLE_UMMA") != nullptr || getenv("TMF_DISABLE_UMMA_CONV1") != nullptr) return false;
  return cout == 32 || cout == 64;
}

int tmf_conv1_fwd_umma(int ng, const float* const* x, const float* const* w, const float* const* bias, void* const* y,
                       double* const* stats, int B, int D, int H, int W, int cout, void* stream) {
  TMF_CHECK_NG(ng);
  TMF_REQUIRE(cout == 32 || cout == 64, "conv1_fwd_umma: Cout must be 32 or 64 (got %d)", cout);
  C1UParams p{};
  p.ng = ng; p.B = B; p.D = D; p.H = H; p.W = W; p.cout = cout;
  p.M = (long long)B * D * H * W;
  p.ntiles = (int)((p.M + 127) / 128);
  p.idesc = make_idesc_bf16(128, cout, 0, 0);
  p.tmem_cols = (cout == 32) ? 64 : 128;
  cudaStream_t st = (cudaStream_t)stream;
  for (int g = 0; g < ng; ++g) {
    TMF_REQUIRE(x[g] && w[g] && y[g], "conv1_fwd_umma: NULL device pointer");
    p.x[g] = x[g]; p.w[g] = w[g];
    p.bias[g] = bias ? bias[g] : nullptr;
    p.y[g] = (__nv_bfloat16*)y[g];
    p.stats[g] = stats ? stats[g] : nullptr;
    if (stats && stats[g]) TMF_CUDA(cudaMemsetAsync(stats[g], 0, sizeof(double) * 2 * cout, st));
  }
  const uint32_t smem = 1024 + C1U_STAGES * 2 * C1U_TILE_BYTES + 2 * 4096 + 8 * (2 * C1U_STAGES) + 64 + 512 + 64;
  static bool attr_done = false;
  if (!attr_done) {
    TMF_CUDA(cudaFuncSetAttribute(conv1_umma_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int per_group = sms / ng;
  if (per_group > p.ntiles) per_group = p.ntiles;
  if (per_group < 1) per_group = 1;
  conv1_umma_fwd_kernel<<<dim3(per_group * ng), C1U_THREADS, smem, st>>>(p);
  TMF_LAUNCH_CHECK();
  return 0;
}
