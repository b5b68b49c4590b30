// eval_ops.cu -- inference-side kernels (reference kfold_train_adversarial.py:144-187):
//   * BatchNorm folding: in eval mode BatchNorm3d is the affine map z = scale*(conv + b) + shift with scale = gamma *
//     rsqrt(running_var + eps), shift = beta - running_mean*scale, so it is folded into the conv operands once --
//     w' = w*scale (packed bf16 [tap][Cout][Cin], or fp32 (Cout,27) for conv1.0), b' = b*scale + shift -- and the conv
//     epilogue then writes the post-BN value directly: no statistics, no finalize launch, one conv + one
//     LeakyReLU/pool pass per layer;
//   * the metric glue of val_step: arg-max labels (ConfusionMatrix / Accuracy input, :172) and the positive-class
//     softmax probability (ROC_AUC input, :186) from the logits, plus the four confusion counts.
#include "common.cuh"

namespace tmf {

__global__ void fold_bn_pack_kernel(GroupPtr<const float> w, GroupPtr<const float> cbias, GroupPtr<const float> gamma,
                                    GroupPtr<const float> beta, GroupPtr<const float> rmean, GroupPtr<const float> rvar,
                                    GroupPtr<__nv_bfloat16> wf, GroupPtr<float> w32, GroupPtr<float> bias_out, int cout,
                                    int cin, int taps, float eps) {
  const int g = blockIdx.z;
  const int64_t total = (int64_t)cout * cin * taps;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)(idx % taps);
    const int ci = (int)((idx / taps) % cin);
    const int co = (int)(idx / ((int64_t)taps * cin));
    const float scale = gamma.p[g][co] * rsqrtf(rvar.p[g][co] + eps);
    const float v = w.p[g][idx] * scale;
    if (wf.p[g] != nullptr) wf.p[g][((int64_t)t * cout + co) * cin + ci] = __float2bfloat16_rn(v);
    if (w32.p[g] != nullptr) w32.p[g][idx] = v;
    if (t == 0 && ci == 0) {
      const float b = cbias.p[g] != nullptr ? cbias.p[g][co] : 0.f;
      bias_out.p[g][co] = (b - rmean.p[g][co]) * scale + beta.p[g][co];
    }
  }
}

// logits [B][C] -> pred[B] (first maximum, like torch.argmax), prob_last[B] = softmax(logits)[:, C-1]; with labels (may be
// NULL) also counts[4] = {TN, FP, FN, TP} for C == 2 (ConfusionMatrix c[label][pred]), accumulated with integer atomics.
__global__ void eval_head_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels,
                                 int64_t* __restrict__ pred, float* __restrict__ prob_last, unsigned long long* counts, int B, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const float* row = logits + (size_t)i * C;
  float mx = row[0];
  int am = 0;
  for (int c = 1; c < C; ++c)
    if (row[c] > mx) { mx = row[c]; am = c; }
  float sum = 0.f;
  for (int c = 0; c < C; ++c) sum += expf(row[c] - mx);
  pred[i] = am;
  prob_last[i] = expf(row[C - 1] - mx) / sum;
  if (counts != nullptr && labels != nullptr && C == 2) {
    const int y = (int)labels[i];
    if (y == 0 || y == 1) atomicAdd(counts + (y * 2 + am), 1ull);
  }
}

}  // namespace tmf

using namespace tmf;

extern "C" {

int tmf_fold_bn_pack(int ng, const float* const* w, const float* const* conv_bias, const float* const* gamma,
                     const float* const* beta, const float* const* running_mean, const float* const* running_var,
                     void* const* wf, float* const* w32, float* const* bias_out, int cout, int cin, int ksize, float eps,
                     void* stream) {
  TMF_CHECK_NG(ng);
  TMF_REQUIRE(ksize == 1 || ksize == 3, "fold_bn_pack: kernel size must be 1 or 3");
  GroupPtr<const float> gw, gcb, gg, gb, gm, gv;
  GroupPtr<__nv_bfloat16> gwf;
  GroupPtr<float> gw32, gbo;
  if (!load_group(gw, w, ng, true, "w") || !load_group(gcb, conv_bias, ng, false, "conv_bias") ||
      !load_group(gg, gamma, ng, true, "gamma") || !load_group(gb, beta, ng, true, "beta") ||
      !load_group(gm, running_mean, ng, true, "running_mean") || !load_group(gv, running_var, ng, true, "running_var") ||
      !load_group(gwf, (__nv_bfloat16* const*)wf, ng, false, "wf") || !load_group(gw32, w32, ng, false, "w32") ||
      !load_group(gbo, bias_out, ng, true, "bias_out"))
    return 1;
  TMF_REQUIRE(wf != nullptr || w32 != nullptr, "fold_bn_pack: no output");
  const int taps = ksize * ksize * ksize;
  const int64_t total = (int64_t)cout * cin * taps;
  dim3 grid((unsigned)(ceil_div(total, 256) < 592 ? ceil_div(total, 256) : 592), 1, ng);
  fold_bn_pack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(gw, gcb, gg, gb, gm, gv, gwf, gw32, gbo, cout, cin, taps, eps);
  TMF_LAUNCH_CHECK();
  return 0;
}

int tmf_eval_head(const float* logits, const int64_t* labels, int64_t* pred, float* prob_last, void* counts4, int B, int C,
                  void* stream) {
  TMF_REQUIRE(logits && pred && prob_last && B > 0 && C >= 1, "eval_head: bad arguments");
  eval_head_kernel<<<ceil_div(B, 128), 128, 0, (cudaStream_t)stream>>>(logits, labels, pred, prob_last,
                                                                        (unsigned long long*)counts4, B, C);
  TMF_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
