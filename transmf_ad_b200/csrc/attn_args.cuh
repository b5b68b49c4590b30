// attn_args.cuh -- argument block shared by the attention kernels (fusion_ops.cu: row-per-warp kernels; attention.cu:
// register-tiled kernels).
#pragma once

namespace tmf {

struct AttnArgs {
  const float* q; const float* kv; const float* out; const float* lse_in; const float* dout;
  float* o; float* lse; float* dq; float* dkv;
  int B, Nq, Nk, heads, dh;
  float scale;
};

// register-tiled kernels (attention.cu); return 0 on success, -1 if the shape is outside their envelope
int attn_tiled_fwd(const AttnArgs& p, cudaStream_t st);
int attn_tiled_bwd(const AttnArgs& p, cudaStream_t st);
// tensor-core kernels (attention_mma.cu): Nq, Nk <= 160, dim_head 16 / 32 / 64; same return convention
int attn_mma_fwd(const AttnArgs& p, cudaStream_t st);
int attn_mma_bwd(const AttnArgs& p, cudaStream_t st);

}  // namespace tmf
