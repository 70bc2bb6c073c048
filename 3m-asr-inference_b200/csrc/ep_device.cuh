// Device-side pieces of the expert-parallel path shared by route.cu, dispatch.cu and ep.cu (protocol: common.cuh,
// EpLayout).
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace b200moe {

constexpr int kEpErrDispatchTimeout = 1;  // a peer's rows did not arrive
constexpr int kEpErrReturnTimeout = 2;    // an owner's results did not arrive
constexpr int kEpErrCountTimeout = 3;     // a peer's counts did not arrive
constexpr int kEpErrModeMismatch = 4;     // the ranks disagree on the call's mode (fold / residual)

// Spin until *flag >= want (acquire, system scope) or the deadline passes. Returns false on timeout.
__device__ __forceinline__ bool ep_wait_flag_sys(const int* flag, int want, unsigned long long deadline_ns) {
  while (ptx::ld_acquire_sys(flag) < want) {
    __nanosleep(64);
    if (ptx::globaltimer_ns() > deadline_ns) return false;
  }
  return true;
}

__device__ __forceinline__ int* ep_ctrl(const EpPeers& ep) { return reinterpret_cast<int*>(ep.base[ep.rank] + ep.lay.ctrl); }

// Group table over expert-contiguous rows: expert e contributes ceil(count[e] / bn) groups, in expert order.  Runs in one
// CTA (every thread calls it); offsets_sm[E + 1] in shared or global memory, scratch[E + 1] in shared memory.
__device__ __forceinline__ void build_groups_block(const int* offsets_sm, int E, int bn, GroupRec* groups, int* n_groups,
                                                   int* h_ready, int gmax, int* scratch) {
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int e = 0; e < E; ++e) {
      scratch[e] = acc;
      const int c = offsets_sm[e + 1] - offsets_sm[e];
      acc += (c + bn - 1) / bn;
    }
    scratch[E] = acc;
    n_groups[0] = acc < gmax ? acc : gmax;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    const int c = offsets_sm[e + 1] - offsets_sm[e];
    const int nt = (c + bn - 1) / bn;
    const int g0 = scratch[e];
    for (int j = 0; j < nt && g0 + j < gmax; ++j) {
      GroupRec r;
      r.expert = e;
      r.row0 = offsets_sm[e] + j * bn;
      r.nrows = min(bn, c - j * bn);
      r.src = 0;
      r.orow0 = r.row0;
      r.pad[0] = r.pad[1] = r.pad[2] = 0;
      groups[g0 + j] = r;
    }
  }
  for (int g = threadIdx.x; g < gmax; g += blockDim.x) h_ready[g] = 0;
}

// Step 1 of the protocol.  Called by every thread of ONE CTA once s_total[E] (this rank's rows per global expert) is
// complete and visible to the CTA: thread t < world stores the counts into rank t and raises the count flag there.
__device__ __forceinline__ void ep_send_counts(const EpPeers& ep, int seq, const int* s_total, int E, int mode) {
  const int t = threadIdx.x;
  if (t < ep.world) {
    int* dst = reinterpret_cast<int*>(ep.base[t] + ep.lay.cnt_all) + ((seq & 1) * kMaxEpWorld + ep.rank) * (E + 1);
    for (int e = 0; e < E; ++e) dst[e] = s_total[e];
    dst[E] = mode;
    // release at system scope: the counts above (same thread) are performed at rank t before the flag is
    ptx::st_release_sys(reinterpret_cast<int*>(ep.base[t] + ep.lay.cnt_flag) + ep.rank, seq);
  }
}

// Step 2.  Called by every thread of a CTA.  Waits for every rank's counts of call `seq`, copies the matrix into s_cnt
// (world x (E + 1) ints) and derives s_base[e]: the row of the owner's receive buffer where THIS rank's first row for
// global expert e belongs (owner rows: expert-major, source-major inside an expert).  Returns false when a peer did not
// deliver or the ranks disagree on the mode: the caller then pushes nothing (the output is poisoned further down).
__device__ __forceinline__ bool ep_wait_counts(const EpPeers& ep, int seq, int E, int mode, int* s_cnt, int* s_base) {
  int* ctrl = ep_ctrl(ep);
  const int W = ep.world, El = ep.E_local, stride = E + 1;
  if (static_cast<int>(threadIdx.x) < W) {
    const int* flags = reinterpret_cast<const int*>(ep.base[ep.rank] + ep.lay.cnt_flag);
    const unsigned long long deadline = ptx::globaltimer_ns() + 1000000ull * static_cast<unsigned>(ep.timeout_ms);
    if (!ep_wait_flag_sys(flags + threadIdx.x, seq, deadline)) atomicExch(&ctrl[3], kEpErrCountTimeout);
  }
  __syncthreads();
  const bool failed = *reinterpret_cast<volatile int*>(&ctrl[3]) != 0;
  // (written by other GPUs during this kernel's lifetime: L2-coherent loads, never the L1 / read-only path)
  const int* cnt = reinterpret_cast<const int*>(ep.base[ep.rank] + ep.lay.cnt_all) + (seq & 1) * kMaxEpWorld * stride;
  for (int i = threadIdx.x; i < W * stride; i += blockDim.x) s_cnt[i] = failed ? 0 : __ldcg(cnt + i);
  __syncthreads();
  bool ok = !failed;
  if (ok)
    for (int s = 0; s < W; ++s) ok = ok && ((s_cnt[s * stride + E] ^ mode) & kEpModeFold) == 0;
  if (!failed && !ok && threadIdx.x == 0) atomicExch(&ctrl[3], kEpErrModeMismatch);
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    const int owner = e / El;
    int row = 0;
    for (int e2 = owner * El; e2 < e; ++e2)
      for (int s = 0; s < W; ++s) row += s_cnt[s * stride + e2];
    for (int s = 0; s < ep.rank; ++s) row += s_cnt[s * stride + e];
    s_base[e] = row;
  }
  __syncthreads();
  return ok;
}

// Owner side of step 2, one CTA (every thread calls it, after ep_wait_counts): the expert kernel's group table over the
// merged rows of this rank's E_local experts.  s_off / s_scr: E_local + 1 ints of shared memory each.
__device__ __forceinline__ void ep_build_groups_merged(const EpPeers& ep, int E, const int* s_cnt, bool ok, int bn,
                                                       GroupRec* groups, int* n_groups, int* h_ready, int gmax,
                                                       int* s_off, int* s_scr) {
  const int W = ep.world, El = ep.E_local, stride = E + 1;
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int e = 0; e < El; ++e) {
      s_off[e] = acc;
      if (ok)
        for (int s = 0; s < W; ++s) acc += s_cnt[s * stride + ep.rank * El + e];
    }
    const int cap_rows = W * ep.cap;
    s_off[El] = acc < cap_rows ? acc : cap_rows;
  }
  __syncthreads();
  build_groups_block(s_off, El, bn, groups, n_groups, h_ready, gmax, s_scr);
}

// Step 3, receiving side: wait until every rank's rows of call `seq` have landed.  Threads [0, world) of one CTA.
__device__ __forceinline__ void ep_wait_rows(const EpPeers& ep, int seq) {
  if (static_cast<int>(threadIdx.x) < ep.world) {
    int* ctrl = ep_ctrl(ep);
    const int* flags = reinterpret_cast<const int*>(ep.base[ep.rank] + ep.lay.disp_flag);
    const unsigned long long deadline = ptx::globaltimer_ns() + 1000000ull * static_cast<unsigned>(ep.timeout_ms);
    if (!ep_wait_flag_sys(flags + threadIdx.x, seq, deadline)) atomicExch(&ctrl[3], kEpErrDispatchTimeout);
  }
}

}  // namespace b200moe
