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
// complete and visible to the CTA.  push_ctas / ffn_ctas: how many CTAs of this rank will signal "delivered" in this call
// (dispatch side) and "results delivered" (expert kernel); their running totals go out with the counts.  Every word
// travels self-validating, (seq << 32) | value, stored straight into every rank: no flag and no release fence behind the
// data (a system-scope release costs an NVLink round trip of its own), the receiver simply re-reads a word until it
// carries this call's sequence number.  s_scr: two ints of shared memory.
__device__ __forceinline__ void ep_send_counts(const EpPeers& ep, int seq, const int* s_total, int E, int mode,
                                               int push_ctas, int ffn_ctas, int* s_scr) {
  if (threadIdx.x == 0) {
    int* ctrl = ep_ctrl(ep);
    s_scr[0] = ctrl[5] = ctrl[5] + push_ctas;
    s_scr[1] = ctrl[6] = ctrl[6] + ffn_ctas;
  }
  __syncthreads();
  const int stride = E + kEpCntExtra;
  const unsigned long long tag = static_cast<unsigned long long>(static_cast<unsigned>(seq)) << 32;
  for (int i = threadIdx.x; i < ep.world * stride; i += blockDim.x) {
    const int t = i / stride, e = i - t * stride;
    unsigned long long* dst =
        reinterpret_cast<unsigned long long*>(ep.base[t] + ep.lay.cnt_all) + static_cast<size_t>(ep.rank) * stride + e;
    const int v = e < E ? s_total[e] : (e == E ? mode : s_scr[e - E - 1]);
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(tag | static_cast<unsigned>(v)) : "memory");
  }
}

// Where this rank's rows go (step 2, sender side; every thread of a CTA calls it): every owner keeps one segment of `cap`
// rows per SOURCE rank, and a source lays its rows for an owner out in its own expert order -- so the row of the owner's
// receive buffer where THIS rank's first row for global expert e belongs follows from this rank's own offsets alone.
// (Round 2 first merged the sources' rows of an expert into contiguous rows; that needs every rank's counts before the
// first row can leave, i.e. a system-scope round trip in front of the pushes: ~6 us per layer on 2 GPUs.)
__device__ __forceinline__ void ep_local_bases(const EpPeers& ep, int E, const int* s_off, int* s_base) {
  for (int e = threadIdx.x; e < E; e += blockDim.x)
    s_base[e] = ep.rank * ep.cap + s_off[e] - s_off[(e / ep.E_local) * ep.E_local];
  __syncthreads();
}

// Owner side of step 2.  Called by every thread of ONE CTA (the one that builds the expert kernel's group table).  Waits for
// every rank's counts of call `seq` and copies the matrix into s_cnt (world x (E + kEpCntExtra) ints).  Returns false when
// a peer did not deliver or the ranks disagree on the mode (status word raised; the output is poisoned further down).
__device__ __forceinline__ bool ep_wait_counts(const EpPeers& ep, int seq, int E, int mode, int* s_cnt) {
  int* ctrl = ep_ctrl(ep);
  const int W = ep.world, stride = E + kEpCntExtra;
  {
    // (written by other GPUs during this kernel's lifetime: system-scope loads that bypass L1)
    const unsigned long long* cnt = reinterpret_cast<const unsigned long long*>(ep.base[ep.rank] + ep.lay.cnt_all);
    const unsigned long long deadline = ptx::globaltimer_ns() + 1000000ull * static_cast<unsigned>(ep.timeout_ms);
    bool gave_up = false;
    for (int i = threadIdx.x; i < W * stride; i += blockDim.x) {
      unsigned long long w;
      while (true) {
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(cnt + i) : "memory");
        if (static_cast<int>(w >> 32) == seq) break;
        if (gave_up || ptx::globaltimer_ns() > deadline) {
          gave_up = true;
          w = 0;
          break;
        }
        __nanosleep(40);
      }
      s_cnt[i] = static_cast<int>(w & 0xffffffffu);
    }
    if (gave_up) atomicExch(&ctrl[3], kEpErrCountTimeout);
  }
  __syncthreads();
  const bool failed = *reinterpret_cast<volatile int*>(&ctrl[3]) != 0;
  bool ok = !failed;
  if (ok)
    for (int s = 0; s < W; ++s) ok = ok && ((s_cnt[s * stride + E] ^ mode) & kEpModeFold) == 0;
  if (!failed && !ok && threadIdx.x == 0) atomicExch(&ctrl[3], kEpErrModeMismatch);
  __syncthreads();
  return ok;
}

// Owner side of step 2, one CTA (every thread calls it, after ep_wait_counts): what this rank's expert kernel and its
// consumers will wait for in this call (local bookkeeping), and the expert kernel's group table over the received rows:
// one run of token tiles per (local expert, source rank); GroupRec::src = the source rank.  Rows of source s for local
// expert e start at  s * cap + (rows of s for this rank's experts before e).
// s_scr: E + 1 ints of shared memory.
__device__ __forceinline__ void ep_build_groups_segmented(const EpPeers& ep, int E, const int* s_cnt, bool ok, int bn,
                                                          GroupRec* groups, int* n_groups, int* h_ready, int gmax,
                                                          int* s_scr) {
  const int W = ep.world, El = ep.E_local, stride = E + kEpCntExtra;
  if (static_cast<int>(threadIdx.x) < W && ok) {
    int* ctrl = ep_ctrl(ep);
    ctrl[kEpCtrlArrive + threadIdx.x] = s_cnt[threadIdx.x * stride + E + 1];
    ctrl[kEpCtrlDone + threadIdx.x] = s_cnt[threadIdx.x * stride + E + 2];
  }
  auto count = [&](int s, int el) { return ok ? min(s_cnt[s * stride + ep.rank * El + el], ep.cap) : 0; };
  // Order of the runs: this rank's OWN rows first (they are there when the kernel starts: the expert kernel works on them
  // while the peers' rows are still crossing NVLink), then the other sources' rows, expert-major.
  auto run = [&](int i, int& el, int& s) {
    if (i < El) {
      el = i;
      s = ep.rank;
    } else {
      const int j = i - El, k = j % (W - 1);
      el = j / (W - 1);
      s = k < ep.rank ? k : k + 1;
    }
  };
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int i = 0; i < El * W; ++i) {
      int el, s2;
      run(i, el, s2);
      s_scr[i] = acc;
      acc += (count(s2, el) + bn - 1) / bn;
    }
    n_groups[0] = acc < gmax ? acc : gmax;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < El * W; i += blockDim.x) {
    int el, s;
    run(i, el, s);
    const int c = count(s, el);
    int base = s * ep.cap;
    for (int e2 = 0; e2 < el; ++e2) base += count(s, e2);
    const int nt = (c + bn - 1) / bn;
    for (int j = 0; j < nt && s_scr[i] + j < gmax; ++j) {
      GroupRec r;
      r.expert = el;
      r.row0 = base + j * bn;
      r.nrows = min(bn, c - j * bn);
      r.src = s;
      r.orow0 = r.row0;
      r.pad[0] = r.pad[1] = r.pad[2] = 0;
      groups[s_scr[i] + j] = r;
    }
  }
  for (int g = threadIdx.x; g < gmax; g += blockDim.x) h_ready[g] = 0;
}

// Step 3, sending side.  Every CTA of the dispatch / route kernel, after a __syncthreads behind its last row store: the
// fence returns once this CTA's pushes have been performed at the peers (one NVLink round trip), then every rank's
// arrive[my rank] goes up by one.  Relaxed increments: the fence in front of them is what orders the rows.
__device__ __forceinline__ void ep_signal(const EpPeers& ep, size_t counters_off) {
  if (threadIdx.x == 0) {
    ptx::fence_acq_rel_sys();
    for (int t = 0; t < ep.world; ++t) {
      int* c = reinterpret_cast<int*>(ep.base[t] + counters_off) + ep.rank;
      asm volatile("red.relaxed.sys.global.add.s32 [%0], 1;" ::"l"(c) : "memory");
    }
  }
}

// Waits until every rank's counter has reached the value announced for this call (ctrl[slot + s], written by this rank's
// own dispatch kernel).  Lanes [0, world) of one warp call it; the loads are relaxed and overlap, one acquire fence
// behind them orders everything that follows.  Returns false (and raises the status word) on a timeout.
__device__ __forceinline__ bool ep_wait_counters(const EpPeers& ep, size_t counters_off, int slot, int err) {
  int* ctrl = ep_ctrl(ep);
  bool ok = true;
  const int s = threadIdx.x & 31;
  if (s < ep.world) {
    const int want = *reinterpret_cast<volatile int*>(&ctrl[slot + s]);
    const int* c = reinterpret_cast<const int*>(ep.base[ep.rank] + counters_off) + s;
    const unsigned long long deadline = ptx::globaltimer_ns() + 1000000ull * static_cast<unsigned>(ep.timeout_ms);
    while (true) {
      int v;
      asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(c) : "memory");
      if (v - want >= 0) break;
      if (ptx::globaltimer_ns() > deadline) {
        atomicExch(&ctrl[3], err);
        ok = false;
        break;
      }
      __nanosleep(40);
    }
    ptx::fence_acq_rel_sys();
  }
  return ok;
}

// The same wait by ONE thread (the expert kernel's TMA producer): all counters are read together each round.
__device__ __forceinline__ bool ep_wait_counters_1t(const EpPeers& ep, size_t counters_off, int slot, int err) {
  int* ctrl = ep_ctrl(ep);
  const int* c = reinterpret_cast<const int*>(ep.base[ep.rank] + counters_off);
  int want[kMaxEpWorld];
#pragma unroll
  for (int s = 0; s < kMaxEpWorld; ++s) want[s] = s < ep.world ? *reinterpret_cast<volatile int*>(&ctrl[slot + s]) : 0;
  const unsigned long long deadline = ptx::globaltimer_ns() + 1000000ull * static_cast<unsigned>(ep.timeout_ms);
  bool ok = true;
  while (true) {
    int v[kMaxEpWorld];
#pragma unroll
    for (int s = 0; s < kMaxEpWorld; ++s) {
      v[s] = want[s];
      if (s < ep.world) asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(v[s]) : "l"(c + s) : "memory");
    }
    bool all = true;
#pragma unroll
    for (int s = 0; s < kMaxEpWorld; ++s) all = all && (v[s] - want[s] >= 0);
    if (all) break;
    if (ptx::globaltimer_ns() > deadline) {
      atomicExch(&ctrl[3], err);
      ok = false;
      break;
    }
    __nanosleep(40);
  }
  ptx::fence_acq_rel_sys();
  return ok;
}

}  // namespace b200moe
