// Device-side pieces of the expert-parallel path shared by dispatch.cu and ep.cu.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace b200moe {

constexpr int kEpErrDispatchTimeout = 1;
constexpr int kEpErrReturnTimeout = 2;

// Spin until *flag >= want (acquire, system scope) or the deadline passes. Returns false on timeout.
__device__ __forceinline__ bool ep_wait_flag_sys(const int* flag, int want, unsigned long long deadline_ns) {
  while (ptx::ld_acquire_sys(flag) < want) {
    __nanosleep(64);
    if (ptx::globaltimer_ns() > deadline_ns) return false;
  }
  return true;
}

// Called by every thread of ONE CTA.  Waits until the rows of layer call `seq` from every rank have landed in this
// rank's receive buffer, then builds the FFN group table over it (expert-major, source rank inside an expert, so that
// consecutive tiles reuse an expert's weights while they are hot in L2) and clears the h flags.
//   s_cnt: shared, world * (E_local + 1) ints;  s_g0: shared, E_local * world + 1 ints.
__device__ __forceinline__ void ep_wait_and_build_groups(const EpPeers& ep, int seq, int bn, GroupRec* groups,
                                                         int* n_groups, int* h_ready, int gmax, int* s_cnt, int* s_g0) {
  int* ctrl = reinterpret_cast<int*>(ep.base[ep.rank] + ep.lay.ctrl);
  const int* flags = reinterpret_cast<const int*>(ep.base[ep.rank] + ep.lay.disp_flag);
  const int W = ep.world, El = ep.E_local, stride = El + 1;
  if (static_cast<int>(threadIdx.x) < W) {
    const unsigned long long deadline = ptx::globaltimer_ns() + 1000000ull * static_cast<unsigned>(ep.timeout_ms);
    if (!ep_wait_flag_sys(flags + threadIdx.x, seq, deadline)) atomicExch(&ctrl[3], kEpErrDispatchTimeout);
  }
  __syncthreads();
  const bool failed = *reinterpret_cast<volatile int*>(&ctrl[3]) != 0;
  const int* rc = reinterpret_cast<const int*>(ep.base[ep.rank] + ep.lay.recv_cnt);
  // a peer that never showed up: run the rest of the layer over nothing rather than over garbage
  for (int i = threadIdx.x; i < W * stride; i += blockDim.x) s_cnt[i] = failed ? 0 : rc[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int e = 0; e < El; ++e)
      for (int s = 0; s < W; ++s) {
        s_g0[e * W + s] = acc;
        acc += (s_cnt[s * stride + e] + bn - 1) / bn;
      }
    s_g0[El * W] = acc;
    n_groups[0] = acc < gmax ? acc : gmax;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < El * W; i += blockDim.x) {
    const int e = i / W, s = i - e * W;
    const int c = s_cnt[s * stride + e];
    int off = 0;  // rows of source s that precede expert e in its segment
    for (int k = 0; k < e; ++k) off += s_cnt[s * stride + k];
    const int nt = (c + bn - 1) / bn;
    const int g0 = s_g0[i];
    for (int j = 0; j < nt && g0 + j < gmax; ++j) {
      GroupRec r;
      r.expert = e;
      r.row0 = s * ep.cap + off + j * bn;
      r.nrows = min(bn, c - j * bn);
      r.src = s;
      r.orow0 = s_cnt[s * stride + El] + off + j * bn;  // row in rank s's own expert-ordered entries
      r.pad[0] = r.pad[1] = r.pad[2] = 0;
      groups[g0 + j] = r;
    }
  }
  for (int g = threadIdx.x; g < gmax; g += blockDim.x) h_ready[g] = 0;
}

}  // namespace b200moe
