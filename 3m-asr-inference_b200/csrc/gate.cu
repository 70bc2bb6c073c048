// Gate: fused router linear + softmax + top-k.
//
// Reference behaviour:
//   3M-ASR router (gate_mode 0): logits = cat(embed, x) . router_weights (+ router_bias); probs = softmax(logits);
//     (value, idx) = probs.max(-1)      trainer_3m_fix/model/dfsmn_base_fmoe_localComm_catEmbed.py:166-181,210-211,
//     built for TensorRT as concat + MatMul + SoftmaxTopKPluginDynamic (layer/positionwise_feed_forward.py:169-207,
//     225; TRTAPI++/plugin/softmax_topk_plugin/softmax_topk_kernel.cu:26-89: value = 1 / sum(exp(l - max))).
//   FastMoE NaiveGate (gate_mode 1): logits = Linear(x); top-k of the logits; softmax over the k selected logits
//     (trainer_3m_fix/fmoe/gates.py:51-66).
// Ties resolve to the lowest expert index (torch.max / torch.topk on CPU); the reference kernel's smem tree picks a
// different winner on exact ties (softmax_topk_kernel.cu:55-64), which only matters for bit-identical logits.
//
// Fast path (E <= 32): the router matrix lives in shared memory as fp32 (128 B per k-row, XOR-swizzled 16 B
// chunks so that the 8 lanes of a quarter-warp hit 8 different bank groups).  A warp owns kTok tokens at a time;
// lane = (k-slot, expert half): it streams 8 consecutive k of its tokens with 128-bit loads, reads 16 router
// weights per k from smem and keeps kTok x 16 fp32 accumulators.  The 16 k-slots are then combined with a
// reduce-scatter butterfly of warp shuffles (15 shuffles per token), after which lane l holds the finished logit
// of one expert and softmax / arg-max / top-k are 5-step shuffle reductions.  fp32 FMA throughout, so routing is
// reproducible against an fp32/fp64 oracle whenever the top-1/top-2 margin exceeds fp32 summation noise.
#include <math_constants.h>

#include "common.cuh"

namespace b200moe {

namespace {

constexpr int kGateThreads = 256;
constexpr int kGateWarps = kGateThreads / 32;

// ---- shared final stage: values distributed as expert = lane + 32 * j -------------------------------------------
template <int NJ>
__device__ __forceinline__ void warp_select_store(float (&v)[NJ], int E, int top_k, int gate_mode, int lane,
                                                  int* __restrict__ idx_out, float* __restrict__ score_out) {
#pragma unroll
  for (int j = 0; j < NJ; ++j)
    if (lane + 32 * j >= E) v[j] = -CUDART_INF_F;

  float first_val = 0.0f;
  float denom = 0.0f;
  float sel_val[8];
  int sel_idx[8];
  const int kk = top_k < 8 ? top_k : 8;
  for (int s = 0; s < kk; ++s) {
    // local best (lowest index wins ties because j ascends and the compare is strict)
    float bv = v[0];
    int bi = lane;
#pragma unroll
    for (int j = 1; j < NJ; ++j) {
      if (v[j] > bv) {
        bv = v[j];
        bi = lane + 32 * j;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) {
        bv = ov;
        bi = oi;
      }
    }
    sel_val[s] = bv;
    sel_idx[s] = bi;
    if (s == 0) {
      first_val = bv;
      if (gate_mode == B200MOE_GATE_3M) {
        // softmax denominator over ALL experts, relative to the maximum
        float part = 0.0f;
#pragma unroll
        for (int j = 0; j < NJ; ++j)
          if (lane + 32 * j < E) part += expf(v[j] - first_val);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        denom = part;
      }
    }
    // remove the winner for the next round
#pragma unroll
    for (int j = 0; j < NJ; ++j)
      if (lane + 32 * j == bi) v[j] = -CUDART_INF_F;
  }
  if (gate_mode != B200MOE_GATE_3M) {
    denom = 0.0f;
    for (int s = 0; s < kk; ++s) denom += expf(sel_val[s] - first_val);
  }
  if (lane == 0) {
    for (int s = 0; s < kk; ++s) {
      idx_out[s] = sel_idx[s];
      score_out[s] = expf(sel_val[s] - first_val) / denom;
    }
  }
}

// 8 consecutive input elements as raw 128-bit words (1 word for 16-bit types, 2 for fp32) + conversion to fp32.
template <typename InT>
struct Raw8 {
  static constexpr int kVec = sizeof(InT) == 4 ? 2 : 1;
  uint4 v[kVec];
  __device__ __forceinline__ void load(const InT* __restrict__ p) {
#pragma unroll
    for (int i = 0; i < kVec; ++i) v[i] = __ldg(reinterpret_cast<const uint4*>(p) + i);
  }
  __device__ __forceinline__ void to_f(float (&o)[8]) const;
};
template <>
__device__ __forceinline__ void Raw8<float>::to_f(float (&o)[8]) const {
  o[0] = __uint_as_float(v[0].x); o[1] = __uint_as_float(v[0].y);
  o[2] = __uint_as_float(v[0].z); o[3] = __uint_as_float(v[0].w);
  o[4] = __uint_as_float(v[1].x); o[5] = __uint_as_float(v[1].y);
  o[6] = __uint_as_float(v[1].z); o[7] = __uint_as_float(v[1].w);
}
template <>
__device__ __forceinline__ void Raw8<bf16>::to_f(float (&o)[8]) const {
  const uint32_t w[4] = {v[0].x, v[0].y, v[0].z, v[0].w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    o[2 * i] = __uint_as_float(w[i] << 16);
    o[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
template <>
__device__ __forceinline__ void Raw8<__half>::to_f(float (&o)[8]) const {
  const __half2* h = reinterpret_cast<const __half2*>(&v[0]);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(h[i]);
    o[2 * i] = f.x;
    o[2 * i + 1] = f.y;
  }
}

__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ---- fast path ---------------------------------------------------------------------------------------------------
template <typename InT, int kTok>
__global__ void __launch_bounds__(kGateThreads, 1)
gate_smem_kernel(const InT* __restrict__ x, const InT* __restrict__ embed, const float* __restrict__ Wr,
                 const float* __restrict__ br, const int* __restrict__ x_len, int S, int T, int D, int Demb, int E,
                 int top_k, int gate_mode, int* __restrict__ idx, float* __restrict__ score) {
  extern __shared__ __align__(16) float s_w[];  // [Rpad][32] fp32, 16-byte chunks XOR-swizzled by (k >> 3) & 7
  const int R = D + Demb;
  const int Rpad = (R + 127) / 128 * 128;
  if (E == 32) {
    // rows are exactly 128 B: asynchronous 16-byte copies straight into the swizzled position (one latency in total)
    for (int i = threadIdx.x; i < R * 8; i += kGateThreads) {
      const int k = i >> 3;
      const int c = i & 7;
      cp_async_16(s_w + k * 32 + ((c ^ ((k >> 3) & 7)) << 2), Wr + static_cast<size_t>(k) * 32 + c * 4);
    }
    for (int i = R * 32 + threadIdx.x; i < Rpad * 32; i += kGateThreads) s_w[i] = 0.0f;
    cp_async_wait_all();
  } else {
    for (int i = threadIdx.x; i < Rpad * 32; i += kGateThreads) {
      const int k = i >> 5;
      const int e = i & 31;
      const float val = (k < R && e < E) ? Wr[static_cast<size_t>(k) * E + e] : 0.0f;
      const int pc = (e >> 2) ^ ((k >> 3) & 7);
      s_w[k * 32 + pc * 4 + (e & 3)] = val;
    }
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int hsel = lane & 1;  // expert half: experts [16*hsel, 16*hsel + 16)
  const int ks = lane >> 1;   // k-slot: 8 consecutive k per 128-wide iteration
  const int swz = ks & 7;
  const int n_tasks = (S + kTok - 1) / kTok;
  const int nkb = Rpad / 128;
  // register ring of prefetched activations: kPf iterations of 8 k per token are in flight at any time
  constexpr int kVec = Raw8<InT>::kVec;
  constexpr int kPf = (kTok * kVec >= 4) ? 2 : (kTok * kVec == 2 ? 4 : 8);

  for (int task = blockIdx.x * kGateWarps + warp; task < n_tasks; task += gridDim.x * kGateWarps) {
    const int t0 = task * kTok;
    float acc[kTok][16];
#pragma unroll
    for (int t = 0; t < kTok; ++t)
#pragma unroll
      for (int e = 0; e < 16; ++e) acc[t][e] = 0.0f;

    Raw8<InT> ring[kPf][kTok];
    auto fetch = [&](Raw8<InT>(&dst)[kTok], int kb) {
      const int k0 = kb * 128 + ks * 8;
      if (k0 < R) {
#pragma unroll
        for (int t = 0; t < kTok; ++t) {
          const int tok = min(t0 + t, S - 1);
          const InT* src = (k0 < Demb) ? embed + static_cast<size_t>(tok) * Demb + k0
                                       : x + static_cast<size_t>(tok) * D + (k0 - Demb);
          dst[t].load(src);
        }
      }
    };
#pragma unroll
    for (int s = 0; s < kPf; ++s)
      if (s < nkb) fetch(ring[s], s);

    for (int kb0 = 0; kb0 < nkb; kb0 += kPf) {
#pragma unroll
      for (int s = 0; s < kPf; ++s) {
        const int kb = kb0 + s;
        if (kb < nkb) {
          const int k0 = kb * 128 + ks * 8;
          float xv[kTok][8];
          if (k0 < R) {
#pragma unroll
            for (int t = 0; t < kTok; ++t) ring[s][t].to_f(xv[t]);
          }
          if (kb + kPf < nkb) fetch(ring[s], kb + kPf);
          if (k0 < R) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4* row = reinterpret_cast<const float4*>(s_w + (k0 + j) * 32);
              float w[16];
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const float4 q = row[(hsel * 4 + c) ^ swz];
                w[4 * c] = q.x; w[4 * c + 1] = q.y; w[4 * c + 2] = q.z; w[4 * c + 3] = q.w;
              }
#pragma unroll
              for (int t = 0; t < kTok; ++t)
#pragma unroll
                for (int e = 0; e < 16; ++e) acc[t][e] = fmaf(xv[t][j], w[e], acc[t][e]);
            }
          }
        }
      }
    }

#pragma unroll
    for (int t = 0; t < kTok; ++t) {
      // reduce-scatter over the 16 k-slots (lane bits 4..1): 16 -> 8 -> 4 -> 2 -> 1 values per lane
      float v8[8], v4[4], v2[2], v1;
      {
        const bool up = (lane & 16) != 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float send = up ? acc[t][i] : acc[t][i + 8];
          const float keep = up ? acc[t][i + 8] : acc[t][i];
          v8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
      }
      {
        const bool up = (lane & 8) != 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float send = up ? v8[i] : v8[i + 4];
          const float keep = up ? v8[i + 4] : v8[i];
          v4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
      }
      {
        const bool up = (lane & 4) != 0;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const float send = up ? v4[i] : v4[i + 2];
          const float keep = up ? v4[i + 2] : v4[i];
          v2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
      }
      {
        const bool up = (lane & 2) != 0;
        const float send = up ? v2[0] : v2[1];
        const float keep = up ? v2[1] : v2[0];
        v1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
      }
      // this lane now holds expert (lane & 1) * 16 + (lane >> 1); move expert e to lane e
      const int src_lane = ((lane & 15) << 1) | (lane >> 4);
      float logit[1];
      logit[0] = __shfl_sync(0xffffffffu, v1, src_lane);
      if (br != nullptr && lane < E) logit[0] += br[lane];

      const int tok = t0 + t;
      if (tok < S) {  // warp-uniform
        bool valid = true;
        if (x_len != nullptr) valid = (tok % T) < x_len[tok / T];
        if (valid) {
          warp_select_store<1>(logit, E, top_k, gate_mode, lane, idx + static_cast<size_t>(tok) * top_k,
                               score + static_cast<size_t>(tok) * top_k);
        } else if (lane < top_k) {
          idx[static_cast<size_t>(tok) * top_k + lane] = -1;
          score[static_cast<size_t>(tok) * top_k + lane] = 0.0f;
        }
      }
    }
  }
}

// ---- generic path: any E <= 256, any R; one warp per token, router matrix read from L2 ------------------------------
template <typename InT>
__global__ void __launch_bounds__(kGateThreads)
gate_generic_kernel(const InT* __restrict__ x, const InT* __restrict__ embed, const float* __restrict__ Wr,
                    const float* __restrict__ br, const int* __restrict__ x_len, int S, int T, int D, int Demb, int E,
                    int top_k, int gate_mode, int* __restrict__ idx, float* __restrict__ score) {
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int tok = blockIdx.x * kGateWarps + warp; tok < S; tok += gridDim.x * kGateWarps) {
    bool valid = true;
    if (x_len != nullptr) valid = (tok % T) < x_len[tok / T];
    if (!valid) {
      if (lane < top_k) {
        idx[static_cast<size_t>(tok) * top_k + lane] = -1;
        score[static_cast<size_t>(tok) * top_k + lane] = 0.0f;
      }
      continue;
    }
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
    const int R = D + Demb;
    for (int k = 0; k < R; ++k) {
      const float xv = (k < Demb) ? to_float(embed[static_cast<size_t>(tok) * Demb + k])
                                  : to_float(x[static_cast<size_t>(tok) * D + (k - Demb)]);
      const float* wrow = Wr + static_cast<size_t>(k) * E;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int e = lane + 32 * j;
        if (e < E) acc[j] = fmaf(xv, wrow[e], acc[j]);
      }
    }
    if (br != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int e = lane + 32 * j;
        if (e < E) acc[j] += br[e];
      }
    }
    warp_select_store<8>(acc, E, top_k, gate_mode, lane, idx + static_cast<size_t>(tok) * top_k,
                         score + static_cast<size_t>(tok) * top_k);
  }
}

// ---- SoftmaxTopKPluginDynamic: logits given -------------------------------------------------------------------------
template <typename InT>
__global__ void __launch_bounds__(kGateThreads)
softmax_top1_kernel(const InT* __restrict__ logits, const int* __restrict__ mask, int S, int T, int E,
                    InT* __restrict__ value, int* __restrict__ idx) {
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int tok = blockIdx.x * kGateWarps + warp; tok < S; tok += gridDim.x * kGateWarps) {
    const bool valid = mask == nullptr || (tok % T) < mask[tok / T];
    if (!valid) {
      if (lane == 0) {
        idx[tok] = -1;
        value[tok] = from_float<InT>(0.0f);
      }
      continue;
    }
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int e = lane + 32 * j;
      v[j] = e < E ? to_float(logits[static_cast<size_t>(tok) * E + e]) : -CUDART_INF_F;
    }
    int i_out = 0;
    float s_out = 0.0f;
    // one winner: reuse the shared routine through registers (lane 0 holds the result)
    int* ip = &i_out;
    float* sp = &s_out;
    warp_select_store<8>(v, E, 1, B200MOE_GATE_3M, lane, ip, sp);
    if (lane == 0) {
      idx[tok] = i_out;
      value[tok] = from_float<InT>(s_out);
    }
  }
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <typename InT, int kTok>
cudaError_t launch_gate_smem(const InT* x, const InT* embed, const float* Wr, const float* br, const int* x_len, int S,
                             int T, int D, int Demb, int E, int top_k, int gate_mode, int* idx, float* score,
                             cudaStream_t stream) {
  const int R = D + Demb;
  const int Rpad = (R + 127) / 128 * 128;
  const size_t smem = static_cast<size_t>(Rpad) * 32 * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gate_smem_kernel<InT, kTok>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int n_tasks = (S + kTok - 1) / kTok;
  int grid = (n_tasks + kGateWarps - 1) / kGateWarps;
  if (grid > sm_count()) grid = sm_count();
  gate_smem_kernel<InT, kTok><<<grid, kGateThreads, smem, stream>>>(x, embed, Wr, br, x_len, S, T, D, Demb, E, top_k,
                                                                     gate_mode, idx, score);
  count_launch();
  return cudaGetLastError();
}

template <typename InT>
cudaError_t launch_gate_typed(const void* xv, const void* ev, const float* Wr, const float* br, const int* x_len,
                              int B, int T, int D, int Demb, int E, int top_k, int gate_mode, int* idx, float* score,
                              cudaStream_t stream) {
  const InT* x = static_cast<const InT*>(xv);
  const InT* embed = static_cast<const InT*>(ev);
  const int S = B * T;
  const int R = D + Demb;
  const bool fast = E <= 32 && D % 8 == 0 && Demb % 8 == 0 && static_cast<size_t>((R + 127) / 128 * 128) * 128 <= 200 * 1024;
  if (fast) {
    // tokens per warp pass: as many as possible (weight reuse from smem) while still giving every warp work
    const int warps = sm_count() * kGateWarps;
    if (S >= 4 * warps)
      return launch_gate_smem<InT, 4>(x, embed, Wr, br, x_len, S, T, D, Demb, E, top_k, gate_mode, idx, score, stream);
    if (S >= 2 * warps)
      return launch_gate_smem<InT, 2>(x, embed, Wr, br, x_len, S, T, D, Demb, E, top_k, gate_mode, idx, score, stream);
    return launch_gate_smem<InT, 1>(x, embed, Wr, br, x_len, S, T, D, Demb, E, top_k, gate_mode, idx, score, stream);
  }
  int grid = (S + kGateWarps - 1) / kGateWarps;
  if (grid > 8 * sm_count()) grid = 8 * sm_count();
  gate_generic_kernel<InT><<<grid, kGateThreads, 0, stream>>>(x, embed, Wr, br, x_len, S, T, D, Demb, E, top_k,
                                                              gate_mode, idx, score);
  count_launch();
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_gate(const void* x, const void* embed, const float* Wr, const float* br, const int* x_len, int B,
                        int T, int D, int Demb, int E, int top_k, int gate_mode, int dtype, int* idx, float* score,
                        cudaStream_t stream) {
  if (B * T == 0) return cudaSuccess;
  if (E < 1 || E > kMaxExperts || top_k < 1 || top_k > 8 || top_k > E) return cudaErrorInvalidValue;
  if (embed == nullptr) Demb = 0;
  if (gate_mode == B200MOE_GATE_3M && top_k != 1) return cudaErrorInvalidValue;
  switch (dtype) {
    case B200MOE_F32:
      return launch_gate_typed<float>(x, embed, Wr, br, x_len, B, T, D, Demb, E, top_k, gate_mode, idx, score, stream);
    case B200MOE_F16:
      return launch_gate_typed<__half>(x, embed, Wr, br, x_len, B, T, D, Demb, E, top_k, gate_mode, idx, score, stream);
    case B200MOE_BF16:
      return launch_gate_typed<bf16>(x, embed, Wr, br, x_len, B, T, D, Demb, E, top_k, gate_mode, idx, score, stream);
    default:
      return cudaErrorInvalidValue;
  }
}

cudaError_t launch_softmax_topk(const void* logits, const int* mask, int B, int T, int E, int dtype, void* value,
                                int* idx, cudaStream_t stream) {
  const int S = B * T;
  if (S == 0) return cudaSuccess;
  if (E < 1 || E > kMaxExperts) return cudaErrorInvalidValue;
  int grid = (S + kGateWarps - 1) / kGateWarps;
  if (grid > 8 * sm_count()) grid = 8 * sm_count();
  switch (dtype) {
    case B200MOE_F32:
      softmax_top1_kernel<float><<<grid, kGateThreads, 0, stream>>>(static_cast<const float*>(logits), mask, S, T, E,
                                                                    static_cast<float*>(value), idx);
      break;
    case B200MOE_F16:
      softmax_top1_kernel<__half><<<grid, kGateThreads, 0, stream>>>(static_cast<const __half*>(logits), mask, S, T,
                                                                     E, static_cast<__half*>(value), idx);
      break;
    case B200MOE_BF16:
      softmax_top1_kernel<bf16><<<grid, kGateThreads, 0, stream>>>(static_cast<const bf16*>(logits), mask, S, T, E,
                                                                   static_cast<bf16*>(value), idx);
      break;
    default:
      return cudaErrorInvalidValue;
  }
  count_launch();
  return cudaGetLastError();
}

}  // namespace b200moe
