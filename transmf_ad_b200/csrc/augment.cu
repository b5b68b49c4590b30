// augment.cu -- the input side of the hot path on the device: what the reference does per sample on the
// host with MONAI, synchronously on the training thread (datasets/ADNI.py:59-84, DataLoader num_workers = 0):
//   ScaleIntensityd            per-volume min-max scaling to [0,1]                                   (:64)
//   RandFlipd(spatial_axis=0)  flip along the first spatial axis                                     (:66)
//   RandRotated(range_x=0.05)  rotation about the first spatial axis by theta, bilinear, border pad  (:67)
//   RandZoomd(0.95..1)         isotropic zoom about the centre, size kept, edge padding              (:68)
// Here: one reduction kernel (min / max per volume) and ONE gather kernel that applies scaling, flip, rotation and zoom
// as a single affine resampling (trilinear, border clamp) -- the three geometric transforms compose into one map, so the
// volume is interpolated once instead of three times.  The random draws stay on the host (transmf_ad_b200/data.py):
// one parameter record per subject, shared by its MRI and PET volume like MONAI's dictionary transforms.
#include "common.cuh"

namespace tmf {

// one block per volume: min / max over n voxels -> out[2*v], out[2*v+1]   (min / max are order-independent: deterministic)
__global__ void __launch_bounds__(1024) volume_minmax_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t n) {
  const float* xv = x + (size_t)blockIdx.x * n;
  float lo = INFINITY, hi = -INFINITY;
  const int64_t n4 = ((uintptr_t)xv & 15) == 0 ? (n >> 2) : 0;
  for (int64_t i = threadIdx.x; i < n4; i += 1024) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(xv) + i);
    lo = fminf(lo, fminf(fminf(v.x, v.y), fminf(v.z, v.w)));
    hi = fmaxf(hi, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
  }
  for (int64_t i = n4 * 4 + threadIdx.x; i < n; i += 1024) {
    const float v = __ldg(xv + i);
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
  __shared__ float slo[32], shi[32];
  lo = -warp_max(-lo);
  hi = warp_max(hi);
  if ((threadIdx.x & 31) == 0) { slo[threadIdx.x >> 5] = lo; shi[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x < 32) {
    lo = -warp_max(-slo[threadIdx.x]);
    hi = warp_max(shi[threadIdx.x]);
    if (threadIdx.x == 0) { out[2 * blockIdx.x] = lo; out[2 * blockIdx.x + 1] = hi; }
  }
}

// params[v] = {flip (0/1), cos(theta), sin(theta), 1/zoom}.  Output voxel (d,h,w), centred p = (d,h,w) - c, c = (size-1)/2:
//   source (centred)  q = (1/zoom) * Rx(theta)^-1 p,  Rx rotating the (h,w) plane;  then the flip mirrors d.
// Sample = trilinear interpolation of the min-max scaled input at q + c with coordinates clamped to the volume (border).
__global__ void __launch_bounds__(256) augment_volume_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                             const float* __restrict__ minmax, const float* __restrict__ params,
                                                             int D, int H, int W, int vols_per_subject) {
  const int v = blockIdx.y;
  const int64_t n = (int64_t)D * H * W;
  const float* s = src + (size_t)v * n;
  const float lo = minmax[2 * v], hi = minmax[2 * v + 1];
  const float sc = (hi > lo) ? 1.f / (hi - lo) : 0.f;          // MONAI rescale_array: a constant volume maps to minv = 0
  const float* pr = params + (size_t)(v / vols_per_subject) * 4;
  const bool flip = pr[0] != 0.f;
  const float cs = pr[1], sn = pr[2], iz = pr[3];
  const bool identity = !flip && cs == 1.f && sn == 0.f && iz == 1.f;
  const float cd = 0.5f * (D - 1), ch = 0.5f * (H - 1), cw = 0.5f * (W - 1);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float val;
    if (identity) {
      val = __ldg(s + i);
    } else {
      const int w = (int)(i % W), h = (int)((i / W) % H), d = (int)(i / ((int64_t)W * H));
      const float pd = d - cd, ph = h - ch, pw = w - cw;
      float qd = iz * pd + cd;
      const float qh = iz * (cs * ph + sn * pw) + ch;
      const float qw = iz * (-sn * ph + cs * pw) + cw;
      if (flip) qd = (D - 1) - qd;
      const float fd = fminf(fmaxf(qd, 0.f), (float)(D - 1)), fh = fminf(fmaxf(qh, 0.f), (float)(H - 1)),
                  fw = fminf(fmaxf(qw, 0.f), (float)(W - 1));
      const int d0 = (int)floorf(fd), h0 = (int)floorf(fh), w0 = (int)floorf(fw);
      const int d1 = min(d0 + 1, D - 1), h1 = min(h0 + 1, H - 1), w1 = min(w0 + 1, W - 1);
      const float td = fd - d0, th = fh - h0, tw = fw - w0;
      auto at = [&](int a, int b, int c) { return __ldg(s + ((int64_t)a * H + b) * W + c); };
      const float c00 = at(d0, h0, w0) * (1.f - tw) + at(d0, h0, w1) * tw;
      const float c01 = at(d0, h1, w0) * (1.f - tw) + at(d0, h1, w1) * tw;
      const float c10 = at(d1, h0, w0) * (1.f - tw) + at(d1, h0, w1) * tw;
      const float c11 = at(d1, h1, w0) * (1.f - tw) + at(d1, h1, w1) * tw;
      val = (c00 * (1.f - th) + c01 * th) * (1.f - td) + (c10 * (1.f - th) + c11 * th) * td;
    }
    dst[(size_t)v * n + i] = (val - lo) * sc;
  }
}

}  // namespace tmf

using namespace tmf;

extern "C" {

int tmf_volume_minmax(const float* x, float* minmax, int nvol, int64_t voxels, void* stream) {
  TMF_REQUIRE(x && minmax && nvol > 0 && voxels > 0, "volume_minmax: bad arguments");
  volume_minmax_kernel<<<nvol, 1024, 0, (cudaStream_t)stream>>>(x, minmax, voxels);
  TMF_LAUNCH_CHECK();
  return 0;
}

int tmf_augment_volumes(const float* src, float* dst, const float* minmax, const float* params, int nvol,
                        int vols_per_subject, int D, int H, int W, void* stream) {
  TMF_REQUIRE(src && dst && minmax && params && src != dst, "augment_volumes: bad pointers (in-place is not supported)");
  TMF_REQUIRE(nvol > 0 && vols_per_subject > 0 && nvol % vols_per_subject == 0 && D > 0 && H > 0 && W > 0,
              "augment_volumes: bad extents");
  const int64_t n = (int64_t)D * H * W;
  int bx = ceil_div(n, 256 * 4);
  if (bx > 148 * 8) bx = 148 * 8;
  augment_volume_kernel<<<dim3(bx, nvol), 256, 0, (cudaStream_t)stream>>>(src, dst, minmax, params, D, H, W, vols_per_subject);
  TMF_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
