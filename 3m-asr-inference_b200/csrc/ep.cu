// Expert parallelism over peer-mapped memory: experts are partitioned contiguously over the GPUs of one NVLink / NVSwitch
// domain (expert e on rank e / E_local, the reference's layout: trainer_3m_fix/model/..._hier.py:259-273), tokens stay
// data-parallel, and rows travel by plain stores into the destination GPU's memory.
//
// Reference behaviour replaced (trainer_3m_fix/fmoe/functions.py):
//   :37-44   fmoe_cuda.expert_exchange   all-to-all of per-expert counts      -> count vectors stored into every rank by one
//                                                                               CTA of the dispatch / route kernel
//   :48-50   .cpu() of the counts        host synchronisation                 -> none: counts are only read on the device
//   :74-80   fmoe_cuda.global_scatter    all-to-all-v of token rows           -> the dispatch / route kernel pushes rows
//                                                                               to their final place in the owner's buffer
//   :185-191 fmoe_cuda.global_gather     all-to-all-v of expert outputs       -> the FFN kernel's second-GEMM epilogue
//                                                                               stores each row on its source rank
// The protocol is described next to EpLayout (common.cuh).  Every wait is a bounded spin on a flag in LOCAL memory that
// a peer raises with st.release.sys after its data.
#include <cstring>

#include "common.cuh"
#include "ep_device.cuh"
#include "ln_device.cuh"
#include "ptx.cuh"

namespace b200moe {

namespace {

// Staged drivers only (in the one-call path the expert kernel's producer does this wait itself).
__global__ void __launch_bounds__(32) ep_wait_rows_kernel(const EpPeers ep) {
  ep_wait_counters(ep, ep.lay.arrive, kEpCtrlArrive, kEpErrDispatchTimeout);
}

// Folded path (the owners write finished output rows into this rank's `out`): whatever consumes `out` next -- the next
// layer's gate, a copy to the host -- is ordered behind this one-CTA kernel, which returns once every owner's done
// counter has reached what it announced for the current layer call.  A missing peer poisons `out` (NaN) and leaves the
// status word.
__global__ void __launch_bounds__(256)
ep_wait_done_kernel(const EpPeers ep, bf16* __restrict__ out, size_t n8) {
  ptx::pdl_launch_dependents();  // the next layer's route kernel may set itself up (it touches constants only until its wait)
  ptx::pdl_wait();               // this rank's own expert kernel (and through it the dispatch that set the targets) has completed
  if (threadIdx.x < 32) ep_wait_counters(ep, ep.lay.done, kEpCtrlDone, kEpErrReturnTimeout);
  __syncthreads();
  if (*reinterpret_cast<volatile int*>(&ep_ctrl(ep)[3]) != 0 && out != nullptr) {
    const uint4 nan8 = make_uint4(0x7fc07fc0u, 0x7fc07fc0u, 0x7fc07fc0u, 0x7fc07fc0u);
    for (size_t i = threadIdx.x; i < n8; i += blockDim.x) reinterpret_cast<uint4*>(out)[i] = nan8;
  }
}

__device__ __forceinline__ void bf16x8_to_f(const uint4& v, float (&o)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    o[2 * i] = __uint_as_float(w[i] << 16);
    o[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

// One warp per token row.  ret_y was written by other GPUs during this kernel's lifetime at the latest, so it is read
// with ordinary (L2-coherent) loads after the acquire, never through the read-only path.
// kLn: norm_final fused in (own instantiation: holding a whole row per warp triples the registers)
template <bool kLn, int kVec>
__global__ void __launch_bounds__(256)
ep_combine_kernel(const EpPeers ep, const int* __restrict__ mapping, const float* __restrict__ score,
                  const bf16* __restrict__ residual, float ff_scale, int S, int D, int top_k, bf16* __restrict__ out,
                  const float* __restrict__ ln_gamma, const float* __restrict__ ln_beta, float ln_eps) {
  // Programmatic dependent launch, both ways.  (1) The next layer's route kernel may start now: it only touches
  // constants (embed, router) until its own griddepcontrol.wait, which returns when this grid has completed.  (2) This
  // kernel is itself launched early (the FFN kernel releases its dependents at its start) and does its own
  // griddepcontrol.wait before it reads anything an earlier kernel wrote: the counter targets, mapping and score come
  // from this layer's route / dispatch kernel, whose stores are only guaranteed visible through the chain of completed
  // grids (a stale target would let the wait pass on the previous layer's counts).  This rank's own expert kernel is one
  // of the owners waited for anyway, so the wait costs nothing the counters would not have cost.
  ptx::pdl_launch_dependents();
  // (kLn) gamma / beta are constants of the layer: fetched before the waits
  LnAffine<kVec> aff;
  if constexpr (kLn) aff.load(ln_gamma, ln_beta, D, threadIdx.x & 31);
  ptx::pdl_wait();
  int* ctrl = ep_ctrl(ep);
  if (threadIdx.x < 32) ep_wait_counters(ep, ep.lay.done, kEpCtrlDone, kEpErrReturnTimeout);
  __syncthreads();
  // A peer that never delivered (rows out or rows back): the layer's output is POISONED with NaNs rather than built
  // from stale rows, so that a stalled rank cannot pass for a result; ctrl[3] (b200moe_ep_status) says which wait failed.
  if (*reinterpret_cast<volatile int*>(&ctrl[3]) != 0) {
    const size_t n8 = static_cast<size_t>(S) * D / 8;
    const uint4 nan8 = make_uint4(0x7fc07fc0u, 0x7fc07fc0u, 0x7fc07fc0u, 0x7fc07fc0u);
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8;
         i += static_cast<size_t>(gridDim.x) * blockDim.x)
      reinterpret_cast<uint4*>(out)[i] = nan8;
    return;
  }
  const bf16* ret_y = reinterpret_cast<const bf16*>(ep.base[ep.rank] + ep.lay.ret_y);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wpb = blockDim.x / 32;
  if constexpr (kLn) {
    // norm_final fused behind the residual add (fmoe_transformer.py:164-166): the warp holds the whole row
    const int nvec = D >> 3;
    for (int s = blockIdx.x * wpb + warp; s < S; s += gridDim.x * wpb) {
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
          const uint4 y = *reinterpret_cast<const uint4*>(ret_y + static_cast<size_t>(row) * D + v * 8);
          float f[8];
          bf16x8_to_f(y, f);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] = fmaf(w, f[i], acc[i]);
        }
        if (residual) {
          const uint4 r = __ldg(reinterpret_cast<const uint4*>(residual + static_cast<size_t>(s) * D + v * 8));
          bf16x8_to_f(r, o[0][k]);
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
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          __nv_bfloat162 pk = __floats2bfloat162_rn(o[0][k][2 * i], o[0][k][2 * i + 1]);
          w[i] = *reinterpret_cast<uint32_t*>(&pk);
        }
        *reinterpret_cast<uint4*>(out + static_cast<size_t>(s) * D + v * 8) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    return;
  }
  for (int s = blockIdx.x * wpb + warp; s < S; s += gridDim.x * wpb) {
    for (int v = lane; v < D / 8; v += 32) {
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
      for (int j = 0; j < top_k; ++j) {
        const int row = mapping[s * top_k + j];
        if (row < 0) continue;
        const float w = score ? score[s * top_k + j] : 1.0f;
        const uint4 y = *reinterpret_cast<const uint4*>(ret_y + static_cast<size_t>(row) * D + v * 8);
        float f[8];
        bf16x8_to_f(y, f);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(w, f[i], acc[i]);
      }
      float o[8];
      if (residual) {
        const uint4 r = __ldg(reinterpret_cast<const uint4*>(residual + static_cast<size_t>(s) * D + v * 8));
        bf16x8_to_f(r, o);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = fmaf(ff_scale, acc[i], o[i]);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = ff_scale * acc[i];
      }
      uint32_t w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        __nv_bfloat162 pk = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
        w[i] = *reinterpret_cast<uint32_t*>(&pk);
      }
      *reinterpret_cast<uint4*>(out + static_cast<size_t>(s) * D + v * 8) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

}  // namespace

cudaError_t launch_ep_wait_rows(const EpPeers& ep, cudaStream_t stream) {
  ep_wait_rows_kernel<<<1, 32, 0, stream>>>(ep);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_ep_wait_done(const EpPeers& ep, void* out, int S, int D, cudaStream_t stream) {
  cudaError_t e = launch_kernel(ep_wait_done_kernel, dim3(1), dim3(256), 0, stream, kPdlFfn, ep, static_cast<bf16*>(out),
                                static_cast<size_t>(S) * D / 8);
  count_launch();
  return e;
}

cudaError_t launch_ep_combine(const EpPeers& ep, const int* mapping, const float* score, const void* residual,
                              float ff_scale, int S, int D, int top_k, void* out, cudaStream_t stream,
                              const float* ln_gamma, const float* ln_beta, float ln_eps) {
  if (D % 8 != 0) return cudaErrorInvalidValue;
  if (ln_gamma != nullptr && (!layernorm_supported(D) || ln_beta == nullptr)) return cudaErrorInvalidValue;
  int blocks = (S + 7) / 8;
  if (blocks < 1) blocks = 1;  // the wait on the return flags must happen even for a rank without tokens
  if (blocks > 4 * 148) blocks = 4 * 148;
  cudaError_t e = launch_kernel(ln_gamma == nullptr ? ep_combine_kernel<false, 1>
                                                     : (D <= 512 ? ep_combine_kernel<true, 2> : ep_combine_kernel<true, kLnMaxVec>),
                                dim3(blocks),
                                dim3(256), 0, stream, kPdlFfn, ep, mapping, score,
                                static_cast<const bf16*>(residual), ff_scale, S, D, top_k, static_cast<bf16*>(out),
                                ln_gamma, ln_beta, ln_eps);
  count_launch();
  return e;
}

}  // namespace b200moe
