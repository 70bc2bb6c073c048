// LayerNorm over one token row held by ONE warp in registers: lane l owns the 8-element vectors l, l + 32, ... of the
// row (kLnMaxVec of them at most, i.e. D <= 1024).  Semantics of torch.nn.LayerNorm as the 3M-ASR blocks use it
// (trainer_3m_fix/layer/fmoe_transformer.py:54-65: eps = 1e-12, biased variance, fp32 gamma / beta); the TensorRT
// plugin that replaces it computes the same thing in fp32 (TRTAPI++/plugin/layer_norm_plugin/layer_norm_kernel.cu).
// Two passes over the registers (mean, then the centred sum of squares): no E[x^2] - mu^2 cancellation.
#pragma once
#include "common.cuh"

namespace b200moe {

constexpr int kLnMaxVec = 4;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// v[k][i]: element (k * 32 + lane) * 8 + i of the row; vectors with (k * 32 + lane) * 8 >= D are ignored.
__device__ __forceinline__ void ln_row_registers(float (&v)[kLnMaxVec][8], int D, int lane,
                                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                                 float eps) {
  const int nvec = D >> 3;
  float s = 0.0f;
#pragma unroll
  for (int k = 0; k < kLnMaxVec; ++k)
    if (k * 32 + lane < nvec) {
#pragma unroll
      for (int i = 0; i < 8; ++i) s += v[k][i];
    }
  const float mean = warp_sum(s) / static_cast<float>(D);
  float q = 0.0f;
#pragma unroll
  for (int k = 0; k < kLnMaxVec; ++k)
    if (k * 32 + lane < nvec) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float d = v[k][i] - mean;
        q = fmaf(d, d, q);
      }
    }
  const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(D) + eps);
#pragma unroll
  for (int k = 0; k < kLnMaxVec; ++k)
    if (k * 32 + lane < nvec) {
      const int f0 = (k * 32 + lane) * 8;
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + f0));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + f0) + 1);
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + f0));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + f0) + 1);
      const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) v[k][i] = fmaf((v[k][i] - mean) * rstd, g[i], b[i]);
    }
}

}  // namespace b200moe
