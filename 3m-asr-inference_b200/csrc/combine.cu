// Weighted gather / combine with the residual add, and the bf16 weight packer.
//
// Reference behaviour fused here into one pass over the rows:
//   gather       out[s] = buf[mapping[s]]                 TRTAPI++/plugin/fmoe_expert_plugin/fmoe_expert_kernel.cu:191-227,
//                                                         trainer_3m_fix/fmoe/functions.py:194 (local_gather)
//   x gate value expert_outputs * gate_value              trainer_3m_fix/layer/positionwise_feed_forward.py:257-258
//   top-k mix    bmm(gate_score[N,1,k], y[N,k,d])         trainer_3m_fix/fmoe/layers.py:204-206
//   x ff_scale, + residual                                trainer_3m_fix/layer/fmoe_transformer.py:155-158
// HBM-bound: one warp per token row, 128-bit loads and stores, fp32 accumulation.
#include "common.cuh"
#include "ln_device.cuh"
#include "ptx.cuh"

namespace b200moe {

namespace {

constexpr int kCombineThreads = 256;

template <typename T>
struct Vec8;  // 8 elements of T as 128-bit words

template <>
struct Vec8<bf16> {
  uint4 v;
  __device__ __forceinline__ void load(const bf16* p) { v = __ldg(reinterpret_cast<const uint4*>(p)); }
  __device__ __forceinline__ void store(bf16* p) const { *reinterpret_cast<uint4*>(p) = v; }
  __device__ __forceinline__ void to_f(float (&o)[8]) const {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      o[2 * i] = __uint_as_float(w[i] << 16);
      o[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ __forceinline__ void from_f(const float (&o)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 p = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&p);
    }
    v = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

template <>
struct Vec8<__half> {
  uint4 v;
  __device__ __forceinline__ void load(const __half* p) { v = __ldg(reinterpret_cast<const uint4*>(p)); }
  __device__ __forceinline__ void store(__half* p) const { *reinterpret_cast<uint4*>(p) = v; }
  __device__ __forceinline__ void to_f(float (&o)[8]) const {
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(h[i]);
      o[2 * i] = f.x;
      o[2 * i + 1] = f.y;
    }
  }
  __device__ __forceinline__ void from_f(const float (&o)[8]) {
    __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(o[2 * i], o[2 * i + 1]);
  }
};

template <>
struct Vec8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = __ldg(reinterpret_cast<const float4*>(p));
    b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  }
  __device__ __forceinline__ void store(float* p) const {
    reinterpret_cast<float4*>(p)[0] = a;
    reinterpret_cast<float4*>(p)[1] = b;
  }
  __device__ __forceinline__ void to_f(float (&o)[8]) const {
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
    o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
  }
  __device__ __forceinline__ void from_f(const float (&o)[8]) {
    a = make_float4(o[0], o[1], o[2], o[3]);
    b = make_float4(o[4], o[5], o[6], o[7]);
  }
};

template <typename T, bool kLn, int kVec>
__global__ void __launch_bounds__(kCombineThreads)
combine_kernel(const T* __restrict__ ybuf, const int* __restrict__ mapping, const float* __restrict__ score,
               const T* __restrict__ residual, float ff_scale, int S, int D, int top_k, T* __restrict__ out,
               const float* __restrict__ ln_gamma, const float* __restrict__ ln_beta, float ln_eps) {
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int warps_per_block = kCombineThreads / 32;
  if constexpr (kLn) {
    // norm_final fused behind the residual add (fmoe_transformer.py:164-166): the warp holds the whole row
    const int nvec = D >> 3;
    LnAffine<kVec> aff;  // constants of the layer, fetched once per warp
    aff.load(ln_gamma, ln_beta, D, lane);
    for (int s = blockIdx.x * warps_per_block + warp; s < S; s += gridDim.x * warps_per_block) {
      float o[1][kVec][8];
#pragma unroll
      for (int k = 0; k < kVec; ++k) {
        const int v = k * 32 + lane;
        if (v >= nvec) continue;
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
        for (int j = 0; j < top_k; ++j) {
          const int row = mapping[s * top_k + j];
          if (row < 0) continue;
          const float w = score ? score[s * top_k + j] : 1.0f;
          Vec8<T> y;
          y.load(ybuf + static_cast<size_t>(row) * D + v * 8);
          float f[8];
          y.to_f(f);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] = fmaf(w, f[i], acc[i]);
        }
        if (residual) {
          Vec8<T> r;
          r.load(residual + static_cast<size_t>(s) * D + v * 8);
          r.to_f(o[0][k]);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[0][k][i] = fmaf(ff_scale, acc[i], o[0][k][i]);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) o[0][k][i] = ff_scale * acc[i];
        }
      }
      ln_rows_registers<kVec, 1>(o, D, lane, aff, ln_eps);
#pragma unroll
      for (int k = 0; k < kVec; ++k) {
        const int v = k * 32 + lane;
        if (v >= nvec) continue;
        Vec8<T> ov;
        ov.from_f(o[0][k]);
        ov.store(out + static_cast<size_t>(s) * D + v * 8);
      }
    }
    return;
  }
  for (int s = blockIdx.x * warps_per_block + warp; s < S; s += gridDim.x * warps_per_block) {
    for (int v = lane; v < D / 8; v += 32) {
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
      for (int j = 0; j < top_k; ++j) {
        const int row = mapping[s * top_k + j];
        if (row < 0) continue;
        const float w = score ? score[s * top_k + j] : 1.0f;
        Vec8<T> y;
        y.load(ybuf + static_cast<size_t>(row) * D + v * 8);
        float f[8];
        y.to_f(f);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(w, f[i], acc[i]);
      }
      float o[8];
      if (residual) {
        Vec8<T> r;
        r.load(residual + static_cast<size_t>(s) * D + v * 8);
        r.to_f(o);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = fmaf(ff_scale, acc[i], o[i]);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = ff_scale * acc[i];
      }
      Vec8<T> ov;
      ov.from_f(o);
      ov.store(out + static_cast<size_t>(s) * D + v * 8);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) pack_bf16_kernel(const T* __restrict__ src, bf16* __restrict__ dst, size_t n) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  const size_t n8 = n / 8;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += stride) {
    Vec8<T> in;
    in.load(src + i * 8);
    float f[8];
    in.to_f(f);
    Vec8<bf16> o;
    o.from_f(f);
    o.store(dst + i * 8);
  }
  // tail (n not a multiple of 8)
  for (size_t i = n8 * 8 + static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = __float2bfloat16_rn(to_float(src[i]));
}

// out[index[i]] = in[i]: the row scatter of fmoe_cuda.local_gather (trainer_3m_fix/fmoe/functions.py:194 with pos as the
// index).  One warp per row, 128-bit accesses; row_bytes is a multiple of 16.
__global__ void __launch_bounds__(kCombineThreads)
scatter_rows_kernel(const uint4* __restrict__ in, const int* __restrict__ index, int n, int n_out, int row_vec,
                    uint4* __restrict__ out) {
  const int wpb = blockDim.x / 32, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = blockIdx.x * wpb + warp; i < n; i += gridDim.x * wpb) {
    const int d = index[i];
    if (d < 0 || d >= n_out) continue;
    for (int v = lane; v < row_vec; v += 32) out[static_cast<size_t>(d) * row_vec + v] = __ldg(in + static_cast<size_t>(i) * row_vec + v);
  }
}

}  // namespace

cudaError_t launch_scatter_rows(const void* in, const int* index, int n, int n_out, int row_bytes, void* out,
                                cudaStream_t stream) {
  if (row_bytes % 16 != 0) return cudaErrorInvalidValue;
  if (n <= 0) return cudaSuccess;
  int blocks = (n + 7) / 8;
  if (blocks > 8 * 148) blocks = 8 * 148;
  scatter_rows_kernel<<<blocks, kCombineThreads, 0, stream>>>(static_cast<const uint4*>(in), index, n, n_out,
                                                              row_bytes / 16, static_cast<uint4*>(out));
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_combine(const void* ybuf, const int* mapping, const float* score, const void* residual,
                           float ff_scale, int S, int D, int top_k, int dtype, void* out, cudaStream_t stream,
                           const float* ln_gamma, const float* ln_beta, float ln_eps) {
  if (S == 0) return cudaSuccess;
  if (D % 8 != 0 || top_k < 1) return cudaErrorInvalidValue;
  if (ln_gamma != nullptr && (!layernorm_supported(D) || ln_beta == nullptr)) return cudaErrorInvalidValue;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int grid = (S + kCombineThreads / 32 - 1) / (kCombineThreads / 32);
  if (grid > sms * 8) grid = sms * 8;
#define B200MOE_COMBINE_LN(T, LN, VEC)                                                                              \
  combine_kernel<T, LN, VEC><<<grid, kCombineThreads, 0, stream>>>(static_cast<const T*>(ybuf), mapping, score,           \
                                                              static_cast<const T*>(residual), ff_scale, S, D,      \
                                                              top_k, static_cast<T*>(out), ln_gamma, ln_beta, ln_eps)
#define B200MOE_COMBINE(T)                 \
  do {                                     \
    if (ln_gamma == nullptr)               \
      B200MOE_COMBINE_LN(T, false, 1);     \
    else if (D <= 512)                     \
      B200MOE_COMBINE_LN(T, true, 2);      \
    else                                   \
      B200MOE_COMBINE_LN(T, true, 4);      \
  } while (0)
  switch (dtype) {
    case B200MOE_F32:
      B200MOE_COMBINE(float);
      break;
    case B200MOE_F16:
      B200MOE_COMBINE(__half);
      break;
    case B200MOE_BF16:
      B200MOE_COMBINE(bf16);
      break;
    default:
      return cudaErrorInvalidValue;
  }
#undef B200MOE_COMBINE
#undef B200MOE_COMBINE_LN
  count_launch();
  return cudaGetLastError();
}

namespace {
__global__ void __launch_bounds__(256) pack_tf32_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t n) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = ptx::round_tf32(src[i]);
}
}  // namespace

cudaError_t launch_pack_tf32(const float* src, float* dst, size_t n, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  size_t want = (n + 255) / 256;
  const int grid = static_cast<int>(want > 148 * 16 ? 148 * 16 : want);
  pack_tf32_kernel<<<grid, 256, 0, stream>>>(src, dst, n);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_pack_bf16(const void* src, int src_dtype, bf16* dst, size_t n, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  size_t want = (n / 8 + 255) / 256;
  int grid = static_cast<int>(want < 1 ? 1 : (want > static_cast<size_t>(sms) * 16 ? static_cast<size_t>(sms) * 16 : want));
  switch (src_dtype) {
    case B200MOE_F32:
      pack_bf16_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(src), dst, n);
      break;
    case B200MOE_F16:
      pack_bf16_kernel<__half><<<grid, 256, 0, stream>>>(static_cast<const __half*>(src), dst, n);
      break;
    case B200MOE_BF16:
      pack_bf16_kernel<bf16><<<grid, 256, 0, stream>>>(static_cast<const bf16*>(src), dst, n);
      break;
    default:
      return cudaErrorInvalidValue;
  }
  count_launch();
  return cudaGetLastError();
}

}  // namespace b200moe
