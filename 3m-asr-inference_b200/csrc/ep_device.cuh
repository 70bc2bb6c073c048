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
// rank's receive buffer, then builds the FFN group table over it and clears the h flags.  Groups of rows that came from
// OTHER ranks are listed first, this rank's own rows last: the FFN kernel works through the table in order, so the
// results that have to cross NVLink leave early and the system-scope fence at the end of the kernel mostly finds local
// stores outstanding.  Every expert's weights are thus streamed twice, the second time from L2.
// (Measured and dropped: telling the remote ranks "your rows are back" as soon as the remote groups are done, from inside
// the running kernel.  It needs a fence.acq_rel.sys on every SM in mid-kernel, and that stalls the SM's other warps for
// the NVLink round trip: +6 us on a 30 us kernel at 2 GPUs, against the ~3 us the early flag could save.)
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
  // table order: (expert, source) for the REMOTE sources first, expert-major, then this rank's own rows per expert
  const int n_rem = El * (W - 1);
  auto entry = [&](int i, int& e, int& s) {
    if (i < n_rem) {
      e = i / (W - 1);
      s = i - e * (W - 1);
      s += s >= ep.rank;
    } else {
      e = i - n_rem;
      s = ep.rank;
    }
  };
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int i = 0; i < El * W; ++i) {
      int e, s;
      entry(i, e, s);
      s_g0[i] = acc;
      acc += (s_cnt[s * stride + e] + bn - 1) / bn;
    }
    s_g0[El * W] = acc;
    n_groups[0] = acc < gmax ? acc : gmax;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < El * W; i += blockDim.x) {
    int e, s;
    entry(i, e, s);
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
