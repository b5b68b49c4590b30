// conv_direct.cu -- CUDA-core convolution kernels of the sNet stack:
//   * weight packing (fp32 master -> bf16 operand layouts),
//   * conv1.0 (Cin = 1) forward and weight-gradient (fp32 input; memory-bound, stays on CUDA cores),
//   * a shared-memory-tiled implicit-GEMM conv3d forward/dgrad and weight-gradient used as the bring-up and
//     cross-check implementation (TMF_CONV_DIRECT) for the tcgen05 kernels in conv_umma.cu.
// Semantics follow nn.Conv3d(k=3, padding=1) / nn.Conv3d(k=1) of reference models/networks.py:22-49.
#include "common.cuh"

namespace tmf {

// ------------------------------------------------------------------------------------------------------------
// weight packing
// ------------------------------------------------------------------------------------------------------------
__global__ void pack_conv_weights_kernel(GroupPtr<const float> w, GroupPtr<__nv_bfloat16> wf,
                                         GroupPtr<__nv_bfloat16> wd, int cout, int cin, int taps) {
  const int g = blockIdx.z;
  const int64_t total = (int64_t)cout * cin * taps;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)(idx % taps);
    const int ci = (int)((idx / taps) % cin);
    const int co = (int)(idx / ((int64_t)taps * cin));
    const __nv_bfloat16 v = __float2bfloat16_rn(w.p[g][idx]);
    wf.p[g][((int64_t)t * cout + co) * cin + ci] = v;
    if (wd.p[g] != nullptr) wd.p[g][((int64_t)(taps - 1 - t) * cin + ci) * cout + co] = v;
  }
}

// ------------------------------------------------------------------------------------------------------------
// conv1.0 forward: one thread per output voxel, all Cout (<= 64) channels, fp32 operands.
// ------------------------------------------------------------------------------------------------------------
// After the call lane l holds sum over the warp of v[l] (v has 32 entries per lane).
__device__ __forceinline__ float warp_transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = upper ? v[i] : v[i + off];
      const float keep = upper ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

constexpr int C1_THREADS = 256;

// Persistent blocks (grid-stride over 256-voxel chunks) so that the per-channel statistics cost ONE flush of
// 2*Cout double atomics per block instead of one per 256 voxels.
__global__ void __launch_bounds__(C1_THREADS)
conv1_fwd_kernel(GroupPtr<const float> x, GroupPtr<const float> w, GroupPtr<const float> bias,
                 GroupPtr<__nv_bfloat16> y, GroupPtr<double> stats, int B, int D, int H, int W, int cout) {
  const int g = blockIdx.z;
  __shared__ __align__(16) float ws[27][64];
  __shared__ float bs[64];
  __shared__ float red[2][64];
  for (int i = threadIdx.x; i < 27 * 64; i += C1_THREADS) {
    const int t = i / 64, co = i % 64;
    ws[t][co] = (co < cout) ? w.p[g][co * 27 + t] : 0.f;
  }
  if (threadIdx.x < 64) {
    bs[threadIdx.x] = (threadIdx.x < cout && bias.p[g] != nullptr) ? bias.p[g][threadIdx.x] : 0.f;
    red[0][threadIdx.x] = 0.f;
    red[1][threadIdx.x] = 0.f;
  }
  __syncthreads();

  const int64_t M = (int64_t)B * D * H * W;
  const int lane = threadIdx.x & 31;
  const bool want_stats = stats.p[g] != nullptr;
  float ssum[2] = {0.f, 0.f}, ssq[2] = {0.f, 0.f};           // lane-owned channel totals (channel = 32*i + lane)
  for (int64_t m0 = (int64_t)blockIdx.x * C1_THREADS; m0 < M; m0 += (int64_t)gridDim.x * C1_THREADS) {
    const int64_t m = m0 + threadIdx.x;
    const bool active = m < M;
    float in[27];
    if (active) {
      const int wq = (int)(m % W);
      const int hq = (int)((m / W) % H);
      const int dq = (int)((m / ((int64_t)W * H)) % D);
      const float* xp = x.p[g] + m;
#pragma unroll
      for (int kd = 0; kd < 3; ++kd)
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const int dd = dq + kd - 1, hh = hq + kh - 1, ww = wq + kw - 1;
            const bool ok = dd >= 0 && dd < D && hh >= 0 && hh < H && ww >= 0 && ww < W;
            in[(kd * 3 + kh) * 3 + kw] = ok ? __ldg(xp + ((int64_t)(kd - 1) * H + (kh - 1)) * W + (kw - 1)) : 0.f;
          }
    } else {
#pragma unroll
      for (int t = 0; t < 27; ++t) in[t] = 0.f;
    }
    for (int c0 = 0; c0 < cout; c0 += 32) {
      float acc[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) acc[c] = bs[c0 + c];
#pragma unroll
      for (int t = 0; t < 27; ++t) {
        const float xv = in[t];
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const float4 wv = *reinterpret_cast<const float4*>(&ws[t][c0 + c4 * 4]);
          acc[c4 * 4 + 0] = fmaf(xv, wv.x, acc[c4 * 4 + 0]);
          acc[c4 * 4 + 1] = fmaf(xv, wv.y, acc[c4 * 4 + 1]);
          acc[c4 * 4 + 2] = fmaf(xv, wv.z, acc[c4 * 4 + 2]);
          acc[c4 * 4 + 3] = fmaf(xv, wv.w, acc[c4 * 4 + 3]);
        }
      }
      float sq[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        acc[c] = active ? round_bf16(acc[c]) : 0.f;
        sq[c] = acc[c] * acc[c];
      }
      if (active) {
        __nv_bfloat16* yp = y.p[g] + m * cout + c0;
        const int nvalid = min(32, cout - c0);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (q * 8 < nvalid) *reinterpret_cast<uint4*>(yp + q * 8) = pack8(&acc[q * 8]);
        }
      }
      if (want_stats) {
        ssum[c0 >> 5] += warp_transpose_reduce32(acc, lane);
        ssq[c0 >> 5] += warp_transpose_reduce32(sq, lane);
      }
    }
  }
  if (want_stats) {
    for (int i = 0; i * 32 < cout; ++i) {
      atomicAdd(&red[0][i * 32 + lane], ssum[i]);
      atomicAdd(&red[1][i * 32 + lane], ssq[i]);
    }
    __syncthreads();
    if (threadIdx.x < cout) {
      double* row = stats.p[g] + (size_t)(blockIdx.x % TMF_STAT_ROWS) * 2 * cout;   // bring-up path: atomics, spread over rows
      atomicAdd(&row[threadIdx.x], (double)red[0][threadIdx.x]);
      atomicAdd(&row[cout + threadIdx.x], (double)red[1][threadIdx.x]);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// conv1.0 weight gradient: dW[co][tap] = sum_v dy[v,co] * x[v + shift(tap)]      (fp32 accumulate, fp32 x)
// A block walks strips of C1W_ROWS output rows of one (sample, plane): the zero-padded x neighbourhood (3 planes x
// (rows+2) x (W+2) floats) and the dy strip are staged in shared memory, so the inner loop has no index arithmetic
// and no bounds checks.  lane = a PAIR of adjacent voxels along w, warp = 4 output channels: 108 register
// accumulators per thread (4 co x 27 taps), 20 shared loads per 216 FMAs.  One transposing warp reduction and 108
// atomics per warp at the very end.
// ------------------------------------------------------------------------------------------------------------
constexpr int C1W_THREADS = 256;
constexpr int C1W_ROWS = 4;

__global__ void __launch_bounds__(C1W_THREADS)
conv1_wgrad_kernel(GroupPtr<const __nv_bfloat16> dy, GroupPtr<const float> x, GroupPtr<float> dw, int B, int D,
                   int H, int W, int cout, int xpitch, int dypitch) {
  extern __shared__ __align__(16) uint8_t c1w_smem[];
  float* xs = reinterpret_cast<float*>(c1w_smem);                                  // [3][ROWS+2][xpitch]
  uint8_t* dys = c1w_smem + sizeof(float) * 3 * (C1W_ROWS + 2) * xpitch;           // [ROWS][W] x dypitch bytes
  const int g = blockIdx.z;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ngroups = cout >> 2;
  const int strips_h = (H + C1W_ROWS - 1) / C1W_ROWS;
  const int64_t nstrips = (int64_t)B * D * strips_h;
  const float* xg = x.p[g];
  const __nv_bfloat16* dyg = dy.p[g];
  const int cq = cout >> 3;                                                          // 16-byte chunks per voxel
  for (int cg0 = 0; cg0 < ngroups; cg0 += C1W_THREADS / 32) {
    const int cg = cg0 + warp;
    float acc[4][27];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int t = 0; t < 27; ++t) acc[c][t] = 0.f;
    for (int64_t sidx = blockIdx.x; sidx < nstrips; sidx += gridDim.x) {
      const int hs = (int)(sidx % strips_h);
      const int d = (int)((sidx / strips_h) % D);
      const int n = (int)(sidx / ((int64_t)strips_h * D));
      const int h0 = hs * C1W_ROWS;
      const int rows = min(C1W_ROWS, H - h0);
      __syncthreads();
      // ---- stage x (zero padded) ----
      for (int i = threadIdx.x; i < 3 * (C1W_ROWS + 2) * xpitch; i += C1W_THREADS) {
        const int col = i % xpitch;
        const int rr = (i / xpitch) % (C1W_ROWS + 2);
        const int pl = i / (xpitch * (C1W_ROWS + 2));
        const int dd = d + pl - 1, hh = h0 + rr - 1, ww = col - 1;
        float v = 0.f;
        if (dd >= 0 && dd < D && hh >= 0 && hh < H && ww >= 0 && ww < W)
          v = __ldg(xg + (((int64_t)n * D + dd) * H + hh) * W + ww);
        xs[i] = v;
      }
      // ---- stage dy strip (rows x W voxels x cout bf16), 16-byte chunks ----
      for (int i = threadIdx.x; i < rows * W * cq; i += C1W_THREADS) {
        const int q = i % cq;
        const int vox = i / cq;                                   // row-major (row, w)
        const int64_t src = ((((int64_t)n * D + d) * H + h0) * W + vox) * cout + q * 8;
        const uint4 v4 = *reinterpret_cast<const uint4*>(dyg + src);
        uint2* dst = reinterpret_cast<uint2*>(dys + (size_t)vox * dypitch + q * 16);   // pitch is only 8-byte aligned
        dst[0] = make_uint2(v4.x, v4.y);
        dst[1] = make_uint2(v4.z, v4.w);
      }
      __syncthreads();
      if (cg < ngroups) {
        for (int r = 0; r < rows; ++r) {
          for (int w0 = 2 * lane; w0 < W; w0 += 64) {
            const bool two = (w0 + 1) < W;
            const uint2 r0 = *reinterpret_cast<const uint2*>(dys + (size_t)(r * W + w0) * dypitch + cg * 8);
            uint2 r1 = make_uint2(0u, 0u);
            if (two) r1 = *reinterpret_cast<const uint2*>(dys + (size_t)(r * W + w0 + 1) * dypitch + cg * 8);
            const float a0[4] = {bf16_lo(r0.x), bf16_hi(r0.x), bf16_lo(r0.y), bf16_hi(r0.y)};
            const float a1[4] = {bf16_lo(r1.x), bf16_hi(r1.x), bf16_lo(r1.y), bf16_hi(r1.y)};
#pragma unroll
            for (int kd = 0; kd < 3; ++kd)
#pragma unroll
              for (int kh = 0; kh < 3; ++kh) {
                const float* xr = xs + ((kd * (C1W_ROWS + 2)) + r + kh) * xpitch + w0;   // x[w0-1 .. w0+2]
                const float2 lo = *reinterpret_cast<const float2*>(xr);
                const float2 hi = *reinterpret_cast<const float2*>(xr + 2);
                const float xv[4] = {lo.x, lo.y, hi.x, hi.y};
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                  const int t = (kd * 3 + kh) * 3 + kw;
#pragma unroll
                  for (int c = 0; c < 4; ++c) {
                    acc[c][t] = fmaf(a0[c], xv[kw], acc[c][t]);
                    acc[c][t] = fmaf(a1[c], xv[kw + 1], acc[c][t]);
                  }
                }
              }
          }
        }
      }
    }
    if (cg < ngroups) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float v[32];
#pragma unroll
        for (int t = 0; t < 32; ++t) v[t] = (t < 27) ? acc[c][t] : 0.f;
        const float tot = warp_transpose_reduce32(v, lane);
        if (lane < 27) atomicAdd(&dw.p[g][(cg * 4 + c) * 27 + lane], tot);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// generic conv3d forward (3x3x3 pad 1 or 1x1x1), CUDA-core implicit GEMM, bf16 operands, fp32 accumulate
// ------------------------------------------------------------------------------------------------------------
constexpr int DT_M = 64, DT_N = 64, DT_K = 32, DT_PAD = 4;

struct ConvDirectArgs {
  GroupPtr<const __nv_bfloat16> a, wf;
  GroupPtr<const float> bias;
  GroupPtr<__nv_bfloat16> y;
  GroupPtr<double> stats;
  int B, D, H, W, cin, cout, ks;
  int64_t M;
};

__global__ void __launch_bounds__(256) conv3d_direct_kernel(ConvDirectArgs p) {
  const int g = blockIdx.z;
  __shared__ __align__(16) float As[DT_K][DT_M + DT_PAD];
  __shared__ __align__(16) float Bs[DT_K][DT_N + DT_PAD];
  __shared__ float red[2][DT_N];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * DT_M;
  const int co0 = blockIdx.y * DT_N;
  // loader coordinates: this thread always loads voxel lm / weight row lm, channel chunk lq
  const int lm = tid & 63, lq = tid >> 6;
  const int64_t m_ld = m0 + lm;
  const bool m_ok = m_ld < p.M;
  int wq = 0, hq = 0, dq = 0;
  if (m_ok) {
    wq = (int)(m_ld % p.W);
    hq = (int)((m_ld / p.W) % p.H);
    dq = (int)((m_ld / ((int64_t)p.W * p.H)) % p.D);
  }
  if (tid < DT_N) { red[0][tid] = 0.f; red[1][tid] = 0.f; }
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int taps = p.ks * p.ks * p.ks;
  const int half = p.ks / 2;
  const __nv_bfloat16* ag = p.a.p[g];
  const __nv_bfloat16* wg = p.wf.p[g];
  for (int t = 0; t < taps; ++t) {
    const int kd = t / (p.ks * p.ks) - half, kh = (t / p.ks) % p.ks - half, kw = t % p.ks - half;
    const int dd = dq + kd, hh = hq + kh, ww = wq + kw;
    const bool ok = m_ok && dd >= 0 && dd < p.D && hh >= 0 && hh < p.H && ww >= 0 && ww < p.W;
    const int64_t src = m_ld + ((int64_t)kd * p.H + kh) * p.W + kw;
    for (int c0 = 0; c0 < p.cin; c0 += DT_K) {
      float fa[8], fb[8];
      const int ch = c0 + lq * 8;
      if (ok && ch < p.cin) {
        unpack8(*reinterpret_cast<const uint4*>(ag + src * p.cin + ch), fa);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) fa[j] = 0.f;
      }
      if (co0 + lm < p.cout && ch < p.cin) {
        unpack8(*reinterpret_cast<const uint4*>(wg + ((int64_t)t * p.cout + co0 + lm) * p.cin + ch), fb);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) fb[j] = 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        As[lq * 8 + j][lm] = fa[j];
        Bs[lq * 8 + j][lm] = fb[j];
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < DT_K; ++k) {
        const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        const float a4[4] = {av.x, av.y, av.z, av.w};
        const float b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
      }
    }
  }
  // epilogue
  const int co = co0 + tx * 4;
  float bv[4] = {0.f, 0.f, 0.f, 0.f};
  if (p.bias.p[g] != nullptr) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (co + j < p.cout) bv[j] = p.bias.p[g][co + j];
  }
  float s[4] = {0.f, 0.f, 0.f, 0.f}, ss[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m < p.M && co < p.cout) {
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[j] = round_bf16(acc[i][j] + bv[j]);
        s[j] += v[j];
        ss[j] += v[j] * v[j];
      }
      uint2 o;
      o.x = pack_bf16(v[0], v[1]);
      o.y = pack_bf16(v[2], v[3]);
      *reinterpret_cast<uint2*>(p.y.p[g] + m * p.cout + co) = o;
    }
  }
  if (p.stats.p[g] != nullptr) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      atomicAdd(&red[0][tx * 4 + j], s[j]);
      atomicAdd(&red[1][tx * 4 + j], ss[j]);
    }
    __syncthreads();
    if (tid < DT_N && co0 + tid < p.cout) {
      double* row = p.stats.p[g] + (size_t)(blockIdx.x % TMF_STAT_ROWS) * 2 * p.cout;
      atomicAdd(&row[co0 + tid], (double)red[0][tid]);
      atomicAdd(&row[p.cout + co0 + tid], (double)red[1][tid]);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// generic weight gradient on CUDA cores: dW[co][ci][tap] = sum_v dy[v,co] * a[v + shift(tap), ci]
// ------------------------------------------------------------------------------------------------------------
constexpr int WG_V = 32;

struct WgradDirectArgs {
  GroupPtr<const __nv_bfloat16> dy, a;
  GroupPtr<float> dw;
  int B, D, H, W, cin, cout, ks;
  int64_t M;
  int ntile_ci, ntile_co;
  int64_t vox_per_block;  // multiple of WG_V
};

__global__ void __launch_bounds__(256) conv3d_wgrad_direct_kernel(WgradDirectArgs p) {
  const int g = blockIdx.z;
  __shared__ __align__(16) float Ys[WG_V][64 + DT_PAD];
  __shared__ __align__(16) float Xs[WG_V][64 + DT_PAD];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int taps = p.ks * p.ks * p.ks, half = p.ks / 2;
  int by = blockIdx.y;
  const int tci = by % p.ntile_ci; by /= p.ntile_ci;
  const int tco = by % p.ntile_co; by /= p.ntile_co;
  const int t = by;
  const int kd = t / (p.ks * p.ks) - half, kh = (t / p.ks) % p.ks - half, kw = t % p.ks - half;
  const int co0 = tco * 64, ci0 = tci * 64;
  const int lv = tid >> 3, lq = tid & 7;  // loader: voxel lv (0..31), channel chunk lq (8 channels)
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int64_t mb = (int64_t)blockIdx.x * p.vox_per_block;
  const int64_t me = min(p.M, mb + p.vox_per_block);
  for (int64_t m0 = mb; m0 < me; m0 += WG_V) {
    const int64_t m = m0 + lv;
    float fy[8], fx[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { fy[j] = 0.f; fx[j] = 0.f; }
    if (m < me) {
      if (co0 + lq * 8 < p.cout)
        unpack8(*reinterpret_cast<const uint4*>(p.dy.p[g] + m * p.cout + co0 + lq * 8), fy);
      const int wq = (int)(m % p.W), hq = (int)((m / p.W) % p.H), dq = (int)((m / ((int64_t)p.W * p.H)) % p.D);
      const int dd = dq + kd, hh = hq + kh, ww = wq + kw;
      if (dd >= 0 && dd < p.D && hh >= 0 && hh < p.H && ww >= 0 && ww < p.W && ci0 + lq * 8 < p.cin) {
        const int64_t src = m + ((int64_t)kd * p.H + kh) * p.W + kw;
        unpack8(*reinterpret_cast<const uint4*>(p.a.p[g] + src * p.cin + ci0 + lq * 8), fx);
      }
    }
    __syncthreads();
    *reinterpret_cast<float4*>(&Ys[lv][lq * 8]) = make_float4(fy[0], fy[1], fy[2], fy[3]);
    *reinterpret_cast<float4*>(&Ys[lv][lq * 8 + 4]) = make_float4(fy[4], fy[5], fy[6], fy[7]);
    *reinterpret_cast<float4*>(&Xs[lv][lq * 8]) = make_float4(fx[0], fx[1], fx[2], fx[3]);
    *reinterpret_cast<float4*>(&Xs[lv][lq * 8 + 4]) = make_float4(fx[4], fx[5], fx[6], fx[7]);
    __syncthreads();
#pragma unroll
    for (int v = 0; v < WG_V; ++v) {
      const float4 av = *reinterpret_cast<const float4*>(&Ys[v][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Xs[v][tx * 4]);
      const float a4[4] = {av.x, av.y, av.z, av.w};
      const float b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= p.cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = ci0 + tx * 4 + j;
      if (ci < p.cin) atomicAdd(&p.dw.p[g][((int64_t)co * p.cin + ci) * taps + t], acc[i][j]);
    }
  }
}

}  // namespace tmf

using namespace tmf;

// implemented in conv1_umma.cu
bool tmf_conv1_fwd_umma_supported(int cout);
int tmf_conv1_fwd_umma(int ng, const float* const* x, const float* const* w, const float* const* bias, void* const* y,
                       double* const* stats, int B, int D, int H, int W, int cout, void* stream);
bool tmf_conv1_wgrad_umma_supported(int W, int cout);
size_t tmf_conv1_wgrad_umma_workspace(int ng, int cout);
int tmf_conv1_wgrad_umma(int ng, const void* const* dy, const float* const* x, float* const* dw, int B, int D, int H,
                         int W, int cout, void* ws, size_t ws_bytes, void* stream);
// implemented in conv_umma.cu
int tmf_conv3d_fwd_umma(int ng, const void* const* a, const void* const* wf, const float* const* bias,
                        void* const* y, double* const* stats, int B, int D, int H, int W, int cin, int cout,
                        int ksize, void* stream);
bool tmf_conv3d_fwd_umma_supported(int D, int H, int W, int cin, int cout, int ksize);
// conv_umma_col.cu: input-stationary rolling-plane, kw-stacked variant (Cin in {32, 64}, Cout a multiple of 32)
int tmf_conv3d_fwd_col(int ng, const void* const* a, const void* const* wf, const float* const* bias,
                       void* const* y, double* const* stats, int B, int D, int H, int W, int cin, int cout, int ksize,
                       void* stream);
bool tmf_conv3d_fwd_col_supported(int ng, int D, int H, int W, int cin, int cout, int ksize);
// wgrad_umma_col.cu: rolling-plane weight gradient, 9 taps per MMA (Cin, Cout multiples of 32)
bool tmf_conv3d_wgrad_col_supported(int ng, int D, int H, int W, int cin, int cout, int ksize);
size_t tmf_conv3d_wgrad_col_workspace(int ng, int D, int H, int W, int cin, int cout, int ksize);
int tmf_conv3d_wgrad_col(int ng, const void* const* dy, const void* const* a, float* const* dw, int B, int D, int H, int W,
                         int cin, int cout, int ksize, void* ws, size_t ws_bytes, void* stream);
int tmf_conv3d_wgrad_umma(int ng, const void* const* dy, const void* const* a, float* const* dw, int B, int D, int H,
                          int W, int cin, int cout, int ksize, void* ws, size_t ws_bytes, void* stream);
size_t tmf_conv3d_wgrad_umma_workspace(int ng, int B, int D, int H, int W, int cin, int cout, int ksize);
bool tmf_conv3d_wgrad_umma_supported(int D, int H, int W, int cin, int cout, int ksize);

extern "C" {

int tmf_conv3d_supported(int op, int impl, int D, int H, int W, int cin, int cout, int ksize) {
  if (ksize != 1 && ksize != 3) return 0;
  if (impl == TMF_CONV_DIRECT) return (cin % 8 == 0 && cout % (op == 0 ? 4 : 8) == 0) ? 1 : 0;
  if (impl == TMF_CONV_UMMA)
    return (op == 0 ? tmf_conv3d_fwd_umma_supported(D, H, W, cin, cout, ksize)
                    : tmf_conv3d_wgrad_umma_supported(D, H, W, cin, cout, ksize))
               ? 1
               : 0;
  return 0;
}

int tmf_pack_conv_weights(int ng, const float* const* w, void* const* wf, void* const* wd, int cout, int cin,
                          int ksize, void* stream) {
  TMF_CHECK_NG(ng);
  TMF_REQUIRE(ksize == 1 || ksize == 3, "pack_conv_weights: ksize must be 1 or 3 (got %d)", ksize);
  GroupPtr<const float> gw;
  GroupPtr<__nv_bfloat16> gwf, gwd;
  if (!load_group(gw, w, ng, true, "w")) return 1;
  if (!load_group(gwf, (__nv_bfloat16* const*)wf, ng, true, "wf")) return 1;
  if (!load_group(gwd, (__nv_bfloat16* const*)wd, ng, false, "wd")) return 1;
  const int taps = ksize * ksize * ksize;
  const int64_t total = (int64_t)cout * cin * taps;
  dim3 grid(min(ceil_div(total, 256), 1184), 1, ng);
  pack_conv_weights_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(gw, gwf, gwd, cout, cin, taps);
  TMF_LAUNCH_CHECK();
  return 0;
}

int tmf_conv1_fwd(int ng, const float* const* x, const float* const* w, const float* const* bias, void* const* y,
                  double* const* stats, int B, int D, int H, int W, int cout, int impl, void* stream) {
  TMF_CHECK_NG(ng);
  if (impl == TMF_CONV_AUTO) impl = tmf_conv1_fwd_umma_supported(cout) ? TMF_CONV_UMMA : TMF_CONV_DIRECT;
  if (impl == TMF_CONV_UMMA) {
    TMF_REQUIRE(tmf_conv1_fwd_umma_supported(cout), "conv1_fwd: tcgen05 path needs Cout 32 or 64 (got %d)", cout);
    return tmf_conv1_fwd_umma(ng, x, w, bias, y, stats, B, D, H, W, cout, stream);
  }
  TMF_REQUIRE(cout >= 8 && cout <= 64 && cout % 8 == 0, "conv1_fwd: Cout must be a multiple of 8 in [8,64] (got %d)",
              cout);
  GroupPtr<const float> gx, gw, gb;
  GroupPtr<__nv_bfloat16> gy;
  GroupPtr<double> gs;
  if (!load_group(gx, x, ng, true, "x") || !load_group(gw, w, ng, true, "w") ||
      !load_group(gb, bias, ng, false, "bias") || !load_group(gy, (__nv_bfloat16* const*)y, ng, true, "y") ||
      !load_group(gs, stats, ng, false, "stats"))
    return 1;
  cudaStream_t st = (cudaStream_t)stream;
  if (stats != nullptr)
    for (int g = 0; g < ng; ++g) TMF_CUDA(cudaMemsetAsync(stats[g], 0, sizeof(double) * 2 * cout * TMF_STAT_ROWS, st));
  const int64_t M = (int64_t)B * D * H * W;
  dim3 grid(min(ceil_div(M, C1_THREADS), 148 * 6), 1, ng);
  conv1_fwd_kernel<<<grid, C1_THREADS, 0, st>>>(gx, gw, gb, gy, gs, B, D, H, W, cout);
  TMF_LAUNCH_CHECK();
  return 0;
}

int64_t tmf_conv1_wgrad_workspace_bytes(int ng, int impl, int W, int cout) {
  if (impl == TMF_CONV_AUTO) impl = tmf_conv1_wgrad_umma_supported(W, cout) ? TMF_CONV_UMMA : TMF_CONV_DIRECT;
  return impl == TMF_CONV_UMMA ? (int64_t)tmf_conv1_wgrad_umma_workspace(ng, cout) : 0;
}

int tmf_conv1_wgrad(int ng, const void* const* dy, const float* const* x, float* const* dw, int B, int D, int H,
                    int W, int cout, int impl, void* ws, size_t ws_bytes, void* stream) {
  TMF_CHECK_NG(ng);
  if (impl == TMF_CONV_AUTO) impl = tmf_conv1_wgrad_umma_supported(W, cout) ? TMF_CONV_UMMA : TMF_CONV_DIRECT;
  if (impl == TMF_CONV_UMMA) {
    TMF_REQUIRE(tmf_conv1_wgrad_umma_supported(W, cout), "conv1_wgrad: tcgen05 path needs Cout = 32 (got %d)", cout);
    return tmf_conv1_wgrad_umma(ng, dy, x, dw, B, D, H, W, cout, ws, ws_bytes, stream);
  }
  TMF_REQUIRE(cout >= 8 && cout <= 64 && cout % 8 == 0, "conv1_wgrad: Cout must be a multiple of 8 in [8,64] (got %d)",
              cout);
  GroupPtr<const __nv_bfloat16> gdy;
  GroupPtr<const float> gx;
  GroupPtr<float> gdw;
  if (!load_group(gdy, (const __nv_bfloat16* const*)dy, ng, true, "dy") || !load_group(gx, x, ng, true, "x") ||
      !load_group(gdw, dw, ng, true, "dw"))
    return 1;
  cudaStream_t st = (cudaStream_t)stream;
  for (int g = 0; g < ng; ++g) TMF_CUDA(cudaMemsetAsync(dw[g], 0, sizeof(float) * 27 * cout, st));
  const int xpitch = (W + 2 + 3) & ~1;                 // even, >= W + 4 (pairs read x[w0-1 .. w0+2])
  const int dypitch = cout * 2 + 8;                    // bytes per voxel row in smem (+8: spreads banks)
  const size_t smem = sizeof(float) * 3 * (C1W_ROWS + 2) * xpitch + (size_t)C1W_ROWS * W * dypitch;
  TMF_REQUIRE(smem <= 227 * 1024, "conv1_wgrad: W=%d too wide for the shared-memory strip", W);
  if (smem > 48 * 1024)
    TMF_CUDA(cudaFuncSetAttribute(conv1_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t nstrips = (int64_t)B * D * ceil_div(H, C1W_ROWS);
  dim3 grid((unsigned)min(nstrips, (int64_t)148 * 2), 1, ng);
  conv1_wgrad_kernel<<<grid, C1W_THREADS, smem, st>>>(gdy, gx, gdw, B, D, H, W, cout, xpitch, dypitch);
  TMF_LAUNCH_CHECK();
  return 0;
}

int tmf_conv3d_fwd(int ng, const void* const* a, const void* const* wf, const float* const* bias, void* const* y,
                   double* const* stats, int B, int D, int H, int W, int cin, int cout, int ksize, int impl,
                   void* stream) {
  TMF_CHECK_NG(ng);
  TMF_REQUIRE(ksize == 1 || ksize == 3, "conv3d_fwd: ksize must be 1 or 3 (got %d)", ksize);
  TMF_REQUIRE(cin % 8 == 0 && cout % 4 == 0, "conv3d_fwd: need Cin %% 8 == 0 and Cout %% 4 == 0 (got %d, %d)", cin,
              cout);
  if (impl == TMF_CONV_AUTO)
    impl = tmf_conv3d_fwd_umma_supported(D, H, W, cin, cout, ksize) ? TMF_CONV_UMMA : TMF_CONV_DIRECT;
  if (impl == TMF_CONV_UMMA) {
    TMF_REQUIRE(tmf_conv3d_fwd_umma_supported(D, H, W, cin, cout, ksize),
                "conv3d_fwd: tcgen05 path does not support D,H,W=%d,%d,%d Cin=%d Cout=%d k=%d", D, H, W, cin, cout,
                ksize);
    if (tmf_conv3d_fwd_col_supported(ng, D, H, W, cin, cout, ksize))
      return tmf_conv3d_fwd_col(ng, a, wf, bias, y, stats, B, D, H, W, cin, cout, ksize, stream);
    return tmf_conv3d_fwd_umma(ng, a, wf, bias, y, stats, B, D, H, W, cin, cout, ksize, stream);
  }
  ConvDirectArgs p;
  if (!load_group(p.a, (const __nv_bfloat16* const*)a, ng, true, "a") ||
      !load_group(p.wf, (const __nv_bfloat16* const*)wf, ng, true, "wf") ||
      !load_group(p.bias, bias, ng, false, "bias") || !load_group(p.y, (__nv_bfloat16* const*)y, ng, true, "y") ||
      !load_group(p.stats, stats, ng, false, "stats"))
    return 1;
  p.B = B; p.D = D; p.H = H; p.W = W; p.cin = cin; p.cout = cout; p.ks = ksize;
  p.M = (int64_t)B * D * H * W;
  cudaStream_t st = (cudaStream_t)stream;
  if (stats != nullptr)
    for (int g = 0; g < ng; ++g) TMF_CUDA(cudaMemsetAsync(stats[g], 0, sizeof(double) * 2 * cout * TMF_STAT_ROWS, st));
  dim3 grid(ceil_div(p.M, DT_M), ceil_div(cout, DT_N), ng);
  conv3d_direct_kernel<<<grid, 256, 0, st>>>(p);
  TMF_LAUNCH_CHECK();
  return 0;
}

int64_t tmf_conv3d_wgrad_workspace_bytes(int ng, int impl, int B, int D, int H, int W, int cin, int cout, int ksize) {
  if (impl == TMF_CONV_DIRECT) return 0;
  if (!tmf_conv3d_wgrad_umma_supported(D, H, W, cin, cout, ksize)) return 0;
  const size_t a = tmf_conv3d_wgrad_umma_workspace(ng, B, D, H, W, cin, cout, ksize);
  const size_t b = tmf_conv3d_wgrad_col_workspace(ng, D, H, W, cin, cout, ksize);
  return (int64_t)(a > b ? a : b);
}

int tmf_conv3d_wgrad(int ng, const void* const* dy, const void* const* a, float* const* dw, int B, int D, int H,
                     int W, int cin, int cout, int ksize, int impl, void* ws, size_t ws_bytes, void* stream) {
  TMF_CHECK_NG(ng);
  TMF_REQUIRE(ksize == 1 || ksize == 3, "conv3d_wgrad: ksize must be 1 or 3 (got %d)", ksize);
  TMF_REQUIRE(cin % 8 == 0 && cout % 8 == 0, "conv3d_wgrad: need Cin %% 8 == 0 and Cout %% 8 == 0 (got %d, %d)", cin,
              cout);
  if (impl == TMF_CONV_AUTO)
    impl = tmf_conv3d_wgrad_umma_supported(D, H, W, cin, cout, ksize) ? TMF_CONV_UMMA : TMF_CONV_DIRECT;
  if (impl == TMF_CONV_UMMA) {
    TMF_REQUIRE(tmf_conv3d_wgrad_umma_supported(D, H, W, cin, cout, ksize),
                "conv3d_wgrad: tcgen05 path does not support D,H,W=%d,%d,%d Cin=%d Cout=%d k=%d", D, H, W, cin, cout,
                ksize);
    if (tmf_conv3d_wgrad_col_supported(ng, D, H, W, cin, cout, ksize))
      return tmf_conv3d_wgrad_col(ng, dy, a, dw, B, D, H, W, cin, cout, ksize, ws, ws_bytes, stream);
    return tmf_conv3d_wgrad_umma(ng, dy, a, dw, B, D, H, W, cin, cout, ksize, ws, ws_bytes, stream);
  }
  WgradDirectArgs p;
  if (!load_group(p.dy, (const __nv_bfloat16* const*)dy, ng, true, "dy") ||
      !load_group(p.a, (const __nv_bfloat16* const*)a, ng, true, "a") || !load_group(p.dw, dw, ng, true, "dw"))
    return 1;
  p.B = B; p.D = D; p.H = H; p.W = W; p.cin = cin; p.cout = cout; p.ks = ksize;
  p.M = (int64_t)B * D * H * W;
  p.ntile_ci = ceil_div(cin, 64);
  p.ntile_co = ceil_div(cout, 64);
  const int taps = ksize * ksize * ksize;
  const int ny = taps * p.ntile_ci * p.ntile_co;
  int nx = max(1, (148 * 8) / ny);
  nx = min(nx, ceil_div(p.M, WG_V));
  p.vox_per_block = (int64_t)ceil_div(ceil_div(p.M, nx), WG_V) * WG_V;
  nx = ceil_div(p.M, p.vox_per_block);
  cudaStream_t st = (cudaStream_t)stream;
  for (int g = 0; g < ng; ++g) TMF_CUDA(cudaMemsetAsync(dw[g], 0, sizeof(float) * (size_t)cout * cin * taps, st));
  dim3 grid(nx, ny, ng);
  conv3d_wgrad_direct_kernel<<<grid, 256, 0, st>>>(p);
  TMF_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
