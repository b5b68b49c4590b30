// conv_umma.cu -- tcgen05 / TMEM / TMA implicit-GEMM convolution kernels (placeholder until bring-up).
#include "common.cuh"

bool tmf_conv3d_fwd_umma_supported(int, int, int, int, int, int) { return false; }
int tmf_conv3d_fwd_umma(int, const void* const*, const void* const*, const float* const*, void* const*,
                        double* const*, int, int, int, int, int, int, int, void*) {
  tmf::set_error("conv3d_fwd: tcgen05 path not built");
  return 1;
}
bool tmf_conv3d_wgrad_umma_supported(int, int, int, int, int, int) { return false; }
int tmf_conv3d_wgrad_umma(int, const void* const*, const void* const*, float* const*, int, int, int, int, int, int,
                          int, void*) {
  tmf::set_error("conv3d_wgrad: tcgen05 path not built");
  return 1;
}
