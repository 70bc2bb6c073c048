// The other encoder plugins of the reference as sm_100a kernels: what sits between the fast_moe
// blocks in the TensorRT graph the reference builds.  All four are bound by HBM bytes (no reuse), so: one pass, 16-byte
// accesses where the layout allows, fp32 arithmetic, the activation dtype at the boundary.
//
//   att_masked_softmax   AttMaskedSoftmaxPluginDynamic, TRTAPI++/plugin/att_masked_softmax_plugin/att_masked_softmax_kernel.cu
//                        :198-277: rows of [B, N, S, ld] scores, out = softmax(scale * x) over the first mask[b] keys, 0 behind
//   glu                  GluPluginDynamic, glu_plugin/glu_kernel.cu:24-37: [M, 2C, N] -> [M, C, N], x1 * sigmoid(x2)
//   masked_fill          MaskedFillPluginDynamic, masked_fill_plugin/masked_fill_kernel.cu:25-39: [B, dim, T], t >= mask[b] -> fill
//   rel_pos_encoding     RelPositionalEncodingPluginDynamic, rel_positional_encoding_plugin/rel_positional_encoding_kernel.cu
//                        :61-71: out = x * scale, pos_emb = pe[:T]
#include <cfloat>

#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

namespace b200moe {

namespace {

constexpr int kThreads = 256;

// One warp per row.  ld <= 32 * kMaxPerLane keys are held in registers between the two reductions (one read, one write).
constexpr int kMaxPerLane = 32;

template <typename T>
__global__ void __launch_bounds__(kThreads)
masked_softmax_kernel(const T* __restrict__ in, const int* __restrict__ mask, float scale, int rows_per_batch, int ld,
                      long long n_rows, T* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (long long r = static_cast<long long>(blockIdx.x) * wpb + warp; r < n_rows; r += static_cast<long long>(gridDim.x) * wpb) {
    const int b = static_cast<int>(r / rows_per_batch);
    const int valid = mask ? min(ld, max(mask[b], 0)) : ld;
    const T* src = in + r * ld;
    T* dst = out + r * ld;
    float v[kMaxPerLane];
    float mx = -FLT_MAX;
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i) {
      const int c = i * 32 + lane;
      v[i] = c < valid ? to_float(src[c]) : -FLT_MAX;
      mx = fmaxf(mx, v[i]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i) {
      const int c = i * 32 + lane;
      v[i] = c < valid ? __expf(scale * (v[i] - mx)) : 0.0f;   // (the reference scales the difference: :70,:107)
      sum += v[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float rz = valid > 0 ? 1.0f / sum : 0.0f;            // a row without valid keys gives zeros (reference: NaN)
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i) {
      const int c = i * 32 + lane;
      if (c < ld) dst[c] = from_float<T>(v[i] * rz);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
glu_kernel(const T* __restrict__ x, long long M, int C, int N, T* __restrict__ y) {
  const long long total = M * C * N;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long m = i / (static_cast<long long>(C) * N);
    const long long rest = i - m * C * N;               // j * N + k
    const float a = to_float(x[m * 2 * C * N + rest]);
    const float g = to_float(x[m * 2 * C * N + static_cast<long long>(C) * N + rest]);
    y[i] = from_float<T>(a / (1.0f + __expf(-g)));
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
masked_fill_kernel(const T* __restrict__ in, const int* __restrict__ mask, float fill, int dim, int T_len, long long total,
                   T* __restrict__ out) {
  const T f = from_float<T>(fill);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int t = static_cast<int>(i % T_len);
    const int b = static_cast<int>(i / (static_cast<long long>(dim) * T_len));
    out[i] = t >= mask[b] ? f : in[i];
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
rel_pos_kernel(const T* __restrict__ in, const T* __restrict__ pe, float scale, long long x_size, long long pos_size,
               T* __restrict__ out, T* __restrict__ pos_emb) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < x_size;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    out[i] = from_float<T>(to_float(in[i]) * scale);
    if (i < pos_size) pos_emb[i] = pe[i];
  }
}

int grid_for_elems(long long n) {
  long long g = (n + kThreads - 1) / kThreads;
  const long long cap = static_cast<long long>(num_sms()) * 8;
  return static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

bool masked_softmax_supported(int ld) { return ld > 0 && ld <= 32 * kMaxPerLane; }

#define B200MOE_BY_DTYPE(CALL)                      \
  switch (dtype) {                                  \
    case B200MOE_F32: { using T = float; CALL; } break;   \
    case B200MOE_F16: { using T = __half; CALL; } break;  \
    case B200MOE_BF16: { using T = bf16; CALL; } break;   \
    default: return cudaErrorInvalidValue;          \
  }

cudaError_t launch_att_masked_softmax(const void* in, const int* mask, float scale, int B, int N, int S, int ld, int dtype,
                                      void* out, cudaStream_t stream) {
  const long long rows = static_cast<long long>(B) * N * S;
  if (rows == 0) return cudaSuccess;
  if (!masked_softmax_supported(ld)) return cudaErrorInvalidValue;
  long long g = (rows + kThreads / 32 - 1) / (kThreads / 32);
  const long long cap = static_cast<long long>(num_sms()) * 8;
  const int grid = static_cast<int>(g > cap ? cap : g);
  B200MOE_BY_DTYPE((masked_softmax_kernel<T><<<grid, kThreads, 0, stream>>>(static_cast<const T*>(in), mask, scale, N * S, ld,
                                                                            rows, static_cast<T*>(out))));
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_glu(const void* x, long long M, int C, int N, int dtype, void* y, cudaStream_t stream) {
  const long long total = M * C * N;
  if (total == 0) return cudaSuccess;
  B200MOE_BY_DTYPE((glu_kernel<T><<<grid_for_elems(total), kThreads, 0, stream>>>(static_cast<const T*>(x), M, C, N,
                                                                                 static_cast<T*>(y))));
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_masked_fill(const void* in, const int* mask, float fill, int B, int dim, int T_len, int dtype, void* out,
                               cudaStream_t stream) {
  const long long total = static_cast<long long>(B) * dim * T_len;
  if (total == 0) return cudaSuccess;
  B200MOE_BY_DTYPE((masked_fill_kernel<T><<<grid_for_elems(total), kThreads, 0, stream>>>(
      static_cast<const T*>(in), mask, fill, dim, T_len, total, static_cast<T*>(out))));
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_rel_pos_encoding(const void* in, const void* pe, float scale, int B, int T_len, int D, int dtype,
                                    void* out, void* pos_emb, cudaStream_t stream) {
  const long long pos = static_cast<long long>(T_len) * D, total = pos * B;
  if (total == 0) return cudaSuccess;
  B200MOE_BY_DTYPE((rel_pos_kernel<T><<<grid_for_elems(total), kThreads, 0, stream>>>(
      static_cast<const T*>(in), static_cast<const T*>(pe), scale, total, pos, static_cast<T*>(out),
      static_cast<T*>(pos_emb))));
  count_launch();
  return cudaGetLastError();
}

#undef B200MOE_BY_DTYPE

}  // namespace b200moe
