// Row LayerNorm either side of the MoE path: `norm_ff` in front of the gate, `norm_final`
// behind the residual add (trainer_3m_fix/layer/fmoe_transformer.py:144-166).  One warp per token row, the row lives
// in registers between the single read and the single write, so the kernel may run in place.  HBM-bound:
// 2 * D * sizeof(T) bytes per token.
#include "common.cuh"
#include "ln_device.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

namespace b200moe {

namespace {

constexpr int kLnThreads = 256;

template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&o)[8]);
template <typename T>
__device__ __forceinline__ void store8(T* p, const float (&o)[8]);

// ordinary (coherent) loads: the kernel may normalise a buffer in place
template <>
__device__ __forceinline__ void load8<bf16>(const bf16* p, float (&o)[8]) {
  const uint4 v = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    o[2 * i] = __uint_as_float(w[i] << 16);
    o[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
template <>
__device__ __forceinline__ void load8<__half>(const __half* p, float (&o)[8]) {
  const uint4 v = *reinterpret_cast<const uint4*>(p);
  const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(h[i]);
    o[2 * i] = f.x;
    o[2 * i + 1] = f.y;
  }
}
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&o)[8]) {
  const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
  o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}
template <>
__device__ __forceinline__ void store8<bf16>(bf16* p, const float (&o)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 pk = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&pk);
  }
  *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}
template <>
__device__ __forceinline__ void store8<__half>(__half* p, const float (&o)[8]) {
  uint4 v;
  __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(o[2 * i], o[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = v;
}
template <>
__device__ __forceinline__ void store8<float>(float* p, const float (&o)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(o[0], o[1], o[2], o[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(o[4], o[5], o[6], o[7]);
}

template <typename T, int kVec>
__global__ void __launch_bounds__(kLnThreads)
layernorm_rows_kernel(const T* in, const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int S,
                      int D, T* out) {
  // `in` may be the output of the kernel in front (programmatic dependent launch): nothing is read before the wait.
  // The kernel behind may start its own prologue at once (its wait returns when this grid has completed).
  ptx::pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // gamma / beta are constants of the layer: fetched before the wait for the producer of `in`, so that their latency
  // (the second dependent L2 round trip of a one-row-per-warp kernel) is off the chain
  LnAffine<kVec> aff;
  aff.load(gamma, beta, D, lane);
  ptx::pdl_wait();
  const int wpb = kLnThreads / 32;
  const int nvec = D >> 3;
  for (int s = blockIdx.x * wpb + warp; s < S; s += gridDim.x * wpb) {
    float v[1][kVec][8];
    const T* src = in + static_cast<size_t>(s) * D;
#pragma unroll
    for (int k = 0; k < kVec; ++k)
      if (k * 32 + lane < nvec) load8<T>(src + (k * 32 + lane) * 8, v[0][k]);
    ln_rows_registers<kVec, 1>(v, D, lane, aff, eps);
    T* dst = out + static_cast<size_t>(s) * D;
#pragma unroll
    for (int k = 0; k < kVec; ++k)
      if (k * 32 + lane < nvec) store8<T>(dst + (k * 32 + lane) * 8, v[0][k]);
  }
}

}  // namespace

bool layernorm_supported(int D) { return D > 0 && D % 8 == 0 && D <= kLnMaxVec * 32 * 8; }

cudaError_t launch_layernorm(const void* in, const float* gamma, const float* beta, float eps, int S, int D, int dtype,
                             void* out, cudaStream_t stream) {
  if (S == 0) return cudaSuccess;
  if (!layernorm_supported(D) || !in || !out || !gamma || !beta) return cudaErrorInvalidValue;
  int grid = (S + kLnThreads / 32 - 1) / (kLnThreads / 32);
  if (grid > num_sms() * 8) grid = num_sms() * 8;
  cudaError_t e;
#define B200MOE_LN_LAUNCH(T)                                                                                          \
  e = D <= 512 ? launch_kernel(layernorm_rows_kernel<T, 2>, dim3(grid), dim3(kLnThreads), 0, stream, kPdlLn,           \
                               static_cast<const T*>(in), gamma, beta, eps, S, D, static_cast<T*>(out))                \
               : launch_kernel(layernorm_rows_kernel<T, kLnMaxVec>, dim3(grid), dim3(kLnThreads), 0, stream, kPdlLn,   \
                               static_cast<const T*>(in), gamma, beta, eps, S, D, static_cast<T*>(out))
  switch (dtype) {
    case B200MOE_F32:
      B200MOE_LN_LAUNCH(float);
      break;
    case B200MOE_F16:
      B200MOE_LN_LAUNCH(__half);
      break;
    case B200MOE_BF16:
      B200MOE_LN_LAUNCH(bf16);
      break;
    default:
      return cudaErrorInvalidValue;
  }
#undef B200MOE_LN_LAUNCH
  count_launch();
  return e;
}

}  // namespace b200moe
