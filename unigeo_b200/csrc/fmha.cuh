// Fused multi-head self-attention over the tokens of one frame, head_dim 64 (spatial attention of the
// SVD UNet).  Replaces xformers memory_efficient_attention (reference enables it at
// /root/reference/model/depthcrafter.py:33) for q/k/v packed as [F*N][3C].
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace ug {

struct FmhaArgs {
  int N, C, heads, F;
  float scale_log2;     // softmax scale * log2(e)
  void* out;            // [F*N][C] 16-bit
  int fmt;              // 0 fp16, 1 bf16
};

// tm: 2-D SWIZZLE_128B map over qkv viewed as [F*N rows][3C cols], box (64 cols, 128 rows)
int launch_fmha_d64(const CUtensorMap& tm, const FmhaArgs& args, cudaStream_t stream);

}  // namespace ug
