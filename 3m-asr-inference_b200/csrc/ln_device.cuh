// LayerNorm over token rows held by ONE warp in registers: lane l owns the 8-element vectors l, l + 32, ... of a row
// (kVec of them at most).  Semantics of torch.nn.LayerNorm as the 3M-ASR blocks use it
// (trainer_3m_fix/layer/fmoe_transformer.py:54-65: eps = 1e-12, biased variance, fp32 gamma / beta); the TensorRT
// plugin that replaces it computes the same thing in fp32 (TRTAPI++/plugin/layer_norm_plugin/layer_norm_kernel.cu).
// Two passes over the registers (mean, then the centred sum of squares): no E[x^2] - mu^2 cancellation.
// R rows are processed together so that their shuffle reductions overlap; the arithmetic of one row does not depend on
// R or kVec (partial sums run over the lane's valid vectors in increasing order, then a xor-shuffle tree), so every
// kernel that normalises through this file produces the same bits for the same row.
#pragma once
#include "common.cuh"

namespace b200moe {

constexpr int kLnMaxVec = 4;  // D <= 1024

// gamma / beta of the vectors a lane owns: the same for every row, loaded once
template <int kVec>
struct LnAffine {
  float g[kVec][8];
  float b[kVec][8];
  __device__ __forceinline__ void load(const float* __restrict__ gamma, const float* __restrict__ beta, int D, int lane) {
    const int nvec = D >> 3;
#pragma unroll
    for (int k = 0; k < kVec; ++k) {
      if (k * 32 + lane < nvec) {
        const int f0 = (k * 32 + lane) * 8;
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + f0));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + f0) + 1);
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + f0));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + f0) + 1);
        g[k][0] = g0.x; g[k][1] = g0.y; g[k][2] = g0.z; g[k][3] = g0.w;
        g[k][4] = g1.x; g[k][5] = g1.y; g[k][6] = g1.z; g[k][7] = g1.w;
        b[k][0] = b0.x; b[k][1] = b0.y; b[k][2] = b0.z; b[k][3] = b0.w;
        b[k][4] = b1.x; b[k][5] = b1.y; b[k][6] = b1.z; b[k][7] = b1.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) g[k][i] = b[k][i] = 0.0f;
      }
    }
  }
};

// v[r][k][i]: element (k * 32 + lane) * 8 + i of row r; vectors with (k * 32 + lane) * 8 >= D are ignored.
// Statistics only: mean[r] and rstd[r] = rsqrt(var + eps), the same on every lane.
template <int kVec, int R>
__device__ __forceinline__ void ln_rows_stats(const float (&v)[R][kVec][8], int D, int lane, float eps, float (&mean)[R],
                                              float (&rstd)[R]) {
  const int nvec = D >> 3;
  const float inv_d = 1.0f / static_cast<float>(D);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    mean[r] = 0.0f;
#pragma unroll
    for (int k = 0; k < kVec; ++k)
      if (k * 32 + lane < nvec) {
#pragma unroll
        for (int i = 0; i < 8; ++i) mean[r] += v[r][k][i];
      }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int r = 0; r < R; ++r) mean[r] += __shfl_xor_sync(0xffffffffu, mean[r], o);
  float q[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    mean[r] *= inv_d;
    q[r] = 0.0f;
#pragma unroll
    for (int k = 0; k < kVec; ++k)
      if (k * 32 + lane < nvec) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float d = v[r][k][i] - mean[r];
          q[r] = fmaf(d, d, q[r]);
        }
      }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int r = 0; r < R; ++r) q[r] += __shfl_xor_sync(0xffffffffu, q[r], o);
#pragma unroll
  for (int r = 0; r < R; ++r) rstd[r] = rsqrtf(q[r] * inv_d + eps);
}

// one element: THE expression every kernel uses, so that all of them agree to the bit
__device__ __forceinline__ float ln_apply(float v, float mean, float rstd, float g, float b) {
  return fmaf((v - mean) * rstd, g, b);
}

template <int kVec, int R>
__device__ __forceinline__ void ln_rows_registers(float (&v)[R][kVec][8], int D, int lane, const LnAffine<kVec>& a,
                                                  float eps) {
  const int nvec = D >> 3;
  float mean[R], rstd[R];
  ln_rows_stats<kVec, R>(v, D, lane, eps, mean, rstd);
#pragma unroll
  for (int r = 0; r < R; ++r) {
#pragma unroll
    for (int k = 0; k < kVec; ++k)
      if (k * 32 + lane < nvec) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[r][k][i] = ln_apply(v[r][k][i], mean[r], rstd[r], a.g[k][i], a.b[k][i]);
      }
  }
}

// One row, kVec vectors per lane (the row-pass and combine kernels: kVec = 2 for D <= 512, else kLnMaxVec).  gamma / beta
// are fetched when they are applied, not held across the reductions: these kernels want many resident warps, not ILP.
template <int kVec>
__device__ __forceinline__ void ln_row_registers(float (&v)[kVec][8], int D, int lane, const float* __restrict__ gamma,
                                                 const float* __restrict__ beta, float eps) {
  const int nvec = D >> 3;
  float mean[1], rstd[1];
  ln_rows_stats<kVec, 1>(reinterpret_cast<const float (&)[1][kVec][8]>(v), D, lane, eps, mean, rstd);
#pragma unroll
  for (int k = 0; k < kVec; ++k)
    if (k * 32 + lane < nvec) {
      const int f0 = (k * 32 + lane) * 8;
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + f0));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + f0) + 1);
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + f0));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + f0) + 1);
      const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) v[k][i] = ln_apply(v[k][i], mean[0], rstd[0], g[i], b[i]);
    }
}

}  // namespace b200moe
