// Token dispatch: per-expert count, exclusive prefix sum, STABLE scatter of token rows into expert-contiguous order.
//
// Reference behaviour: TRTAPI++/plugin/fmoe_expert_plugin/fmoe_expert_kernel.cu:25-128 (one CTA, shared-memory
// atomicAdd ranks -- order within an expert is nondeterministic -- thread-0 serial scan, then a scalar copy
// kernel) and trainer_3m_fix/fmoe/functions.py:29-35,72 (torch.sort + unique + fmoe_cuda.local_scatter).
// Here the within-expert order is defined (stable: by entry index), so mapping is reproducible bit for bit:
//   mapping[i] = offsets[idx[i]] + #{ j < i : idx[j] == idx[i] },   pos = mapping^-1.
//
// Two launches: count (per-chunk histograms) and scatter (every CTA derives its chunk's base offsets from the
// histograms, ranks its entries with warp match/ballot, and copies rows with 128-bit accesses, casting to bf16).
// CTA 0 of the scatter kernel also emits counts / offsets / the FFN group table and clears the FFN flags.
#include "common.cuh"
#include "ep_device.cuh"
#include "ptx.cuh"

namespace b200moe {

namespace {

struct Chunking {
  int chunk_tok;  // tokens per CTA (multiple of 32)
  int chunk;      // entries per CTA = chunk_tok * top_k
  int nchunks;    // grid size (<= kMaxChunks)
};

Chunking make_chunking(int S, int top_k) {
  // ~2 CTAs per SM; a CTA's chunk is a multiple of 32 tokens so that small batches still spread over many SMs
  // (the row copy is latency bound: 3 200 tokens -> 100 CTAs x 32 rows, 4 rows per warp in flight at once) and so
  // that chunk boundaries coincide with the 32-token histogram rows the tensor-core gate emits.
  Chunking c;
  int chunk = (S + 295) / 296;
  chunk = (chunk + 31) / 32 * 32;
  if (chunk < 32) chunk = 32;
  c.chunk_tok = chunk;
  c.chunk = chunk * top_k;
  c.nchunks = (S + chunk - 1) / chunk;
  if (c.nchunks < 1) c.nchunks = 1;
  return c;
}

__global__ void __launch_bounds__(kDispatchThreads)
dispatch_count_kernel(const int* __restrict__ idx, int Sk, int E, int chunk, int* __restrict__ chunk_hist) {
  __shared__ int hist[kMaxExperts];
  for (int e = threadIdx.x; e < E; e += blockDim.x) hist[e] = 0;
  __syncthreads();
  const int begin = blockIdx.x * chunk;
  const int end = min(begin + chunk, Sk);
  for (int i = begin + threadIdx.x; i < end; i += blockDim.x) {
    const int e = idx[i];
    if (e >= 0 && e < E) atomicAdd(&hist[e], 1);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < E; e += blockDim.x) chunk_hist[blockIdx.x * E + e] = hist[e];
}

// 8 consecutive elements -> 8 bf16 packed in a uint4 (one 128-bit store).
template <typename InT>
__device__ __forceinline__ uint4 load8_as_bf16(const InT* __restrict__ p);

template <>
__device__ __forceinline__ uint4 load8_as_bf16<bf16>(const bf16* __restrict__ p) {
  return __ldg(reinterpret_cast<const uint4*>(p));
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

template <>
__device__ __forceinline__ uint4 load8_as_bf16<float>(const float* __restrict__ p) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  uint4 o;
  o.x = pack_bf16x2(a.x, a.y);
  o.y = pack_bf16x2(a.z, a.w);
  o.z = pack_bf16x2(b.x, b.y);
  o.w = pack_bf16x2(b.z, b.w);
  return o;
}

template <>
__device__ __forceinline__ uint4 load8_as_bf16<__half>(const __half* __restrict__ p) {
  uint4 in = __ldg(reinterpret_cast<const uint4*>(p));
  const __half2* h = reinterpret_cast<const __half2*>(&in);
  uint4 o;
  uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 f = __half22float2(h[k]);
    ow[k] = pack_bf16x2(f.x, f.y);
  }
  return o;
}

// kEp: expert parallelism (protocol: common.cuh, EpLayout).  Expert e lives on rank e / E_local.  CTA 0 sends this rank's
// per-expert counts to every rank, every CTA waits for all ranks' counts and pushes its rows over peer-mapped memory
// (NVLink) to where they belong in the owner's receive buffer -- expert-major, source-major inside an expert -- together
// with 8 bytes of routing data per row; the last CTA to finish raises the arrival flag with a system-scope release (the
// reference does this with two NCCL all-to-alls and a host round trip: trainer_3m_fix/fmoe/functions.py:37-50,74-80).
// kXf32: the expert-order buffer keeps fp32 rows (TF32 compute; fp32 activations only) instead of bf16.
template <typename InT, bool kEp, bool kXf32>
__global__ void __launch_bounds__(kDispatchThreads)
dispatch_scatter_kernel(const InT* __restrict__ x, const int* __restrict__ idx, const float* __restrict__ score,
                        int Sk, int D, int E, int top_k, int chunk, int nchunks,
                        const int* __restrict__ chunk_hist, int hist_rows, int rows_per_chunk, int bn, int gmax, int* __restrict__ counts,
                        int* __restrict__ offsets, int* __restrict__ mapping, int* __restrict__ pos,
                        float* __restrict__ row_score, bf16* __restrict__ xbuf, GroupRec* groups, int* n_groups,
                        int* h_ready, int* counts_out, int* offsets_out, int* mapping_out, InT* __restrict__ drop_out,
                        const InT* __restrict__ drop_residual, int early_trigger, const EpPeers ep,
                        int ep_fold_wait, int ep_mode, int ep_send, int ep_ffn_ctas) {
  constexpr int kWarps = kDispatchThreads / 32;
  constexpr int kRowsPerBatch = 8;            // rows a warp keeps in flight during the copy (x 16 B per lane: 4 KiB per warp and pass)
  constexpr int kMaxParts = 8;
  // All shared memory is dynamic and sized by E (about 4.6 KB at E = 32), so that a CTA of the expert-FFN kernel
  // (launched early under programmatic dependent launch, ~211 KB) fits on the same SM beside a dispatch CTA.
  extern __shared__ int s_dyn[];
  int* s_total = s_dyn;                       // [E]     tokens per expert over all chunks
  int* s_cursor = s_total + E;                // [E]     next expert-order row for this CTA's entries
  int* s_off = s_cursor + E;                  // [E + 1] global exclusive offsets
  int* s_scratch = s_off + E + 1;             // [E + 1]
  int* s_dst = s_scratch + E + 1;             // [kDispatchThreads] destination row of each entry of the segment
  int* s_wcnt = s_dst + kDispatchThreads;     // [kWarps][E] per-warp counts of the current segment
  int* s_part = s_wcnt + kWarps * E;          // [nparts][2][E] partial column sums (before / total)
  int* s_exp = s_part + kMaxParts * 2 * E;    // [kDispatchThreads] expert of each entry of the segment (kEp)
  int* s_cnt = s_exp + kDispatchThreads;      // [world][E + kEpCntExtra] every rank's count message (kEp)
  int* s_base = s_cnt + kMaxEpWorld * (E + kEpCntExtra);  // [E] owner row of this rank's first row per expert (kEp)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (early_trigger) ptx::pdl_launch_dependents();  // the expert-FFN kernel may set itself up; it blocks in pdl_wait()
  ptx::pdl_wait();               // idx / score / hist32 come from the gate; xbuf, pos, groups may still be read by
                                 // the previous layer's FFN kernel

  // 1. column sums of the chunk histograms: totals and the part that precedes this chunk.  All 256 threads take
  //    part: thread (part, e) sums the chunks c = part, part + nparts, ... (independent loads, one latency), then
  //    the partial sums are combined through shared memory.
  {
    const int nparts = max(1, min(kMaxParts, static_cast<int>(blockDim.x) / E));
    const int e = threadIdx.x % E;
    const int part = threadIdx.x / E;
    if (part < nparts) {
      int before = 0, total = 0;
      // histogram rows are per dispatch chunk (count kernel) or per 32 tokens (tensor-core gate)
      const int my_first_row = static_cast<int>(blockIdx.x) * rows_per_chunk;
      for (int c = part; c < hist_rows; c += nparts) {
        const int v = chunk_hist[c * E + e];
        if (c < my_first_row) before += v;
        total += v;
      }
      s_part[(part * 2) * E + e] = before;
      s_part[(part * 2 + 1) * E + e] = total;
    }
    for (int i = threadIdx.x; i < kWarps * E; i += blockDim.x) s_wcnt[i] = 0;
    __syncthreads();
    for (int ee = threadIdx.x; ee < E; ee += blockDim.x) {
      int before = 0, total = 0;
      for (int pt = 0; pt < nparts; ++pt) {
        before += s_part[(pt * 2) * E + ee];
        total += s_part[(pt * 2 + 1) * E + ee];
      }
      s_cursor[ee] = before;
      s_total[ee] = total;
    }
  }
  __syncthreads();
  // 2. exclusive scan over experts (E <= 256: a serial scan by one thread is a few hundred cycles)
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int e = 0; e < E; ++e) {
      s_off[e] = acc;
      acc += s_total[e];
    }
    s_off[E] = acc;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < E; e += blockDim.x) s_cursor[e] += s_off[e];
  __syncthreads();
  int ep_seq = 0;
  bool ep_ok = true;
  if (kEp) {
    ep_seq = ep_ctrl(ep)[0] + 1;   // (only this kernel's last CTA advances the word, after every CTA has read it)
    if (blockIdx.x == 0 && ep_send) ep_send_counts(ep, ep_seq, s_total, E, ep_mode, gridDim.x, ep_ffn_ctas, s_dst);
    ep_local_bases(ep, E, s_off, s_base);  // (no wait: rows go into this rank's own segment of the owners' buffers)
  }

  const int begin = blockIdx.x * chunk;
  const int end = min(begin + chunk, Sk);
  for (int seg = begin; seg < end; seg += kDispatchThreads) {
    const int n = min(kDispatchThreads, end - seg);
    // 3. stable rank of the segment's entries: (earlier segments) + (earlier warps) + (earlier lanes)
    const int i = seg + threadIdx.x;
    int e = -1;
    if (static_cast<int>(threadIdx.x) < n) {
      e = idx[i];
      if (e < 0 || e >= E) e = -1;
    }
    const unsigned peers = __match_any_sync(0xffffffffu, e);
    if (e >= 0 && lane == __ffs(peers) - 1) s_wcnt[warp * E + e] = __popc(peers);
    __syncthreads();
    int dst = -1;
    if (e >= 0) {
      int base = s_cursor[e];
      for (int w = 0; w < warp; ++w) base += s_wcnt[w * E + e];
      dst = base + __popc(peers & ((1u << lane) - 1u));
    }
    s_dst[threadIdx.x] = dst;
    if (kEp) s_exp[threadIdx.x] = e;
    if (static_cast<int>(threadIdx.x) < n) {
      mapping[i] = dst;
      if (mapping_out) mapping_out[i] = dst;
      if (dst >= 0) {
        pos[dst] = i;
        const float sc = score ? score[i] : 1.0f;
        row_score[dst] = sc;
        if (kEp && ep_ok) {
          const int owner = e / ep.E_local;
          int2* meta = reinterpret_cast<int2*>(ep.base[owner] + ep.lay.meta);
          meta[s_base[e] + dst - s_off[e]] =
              make_int2(ep_meta_word(ep.rank, (ep_mode & kEpModeFold) ? i / top_k : dst, (ep_mode & kEpModeResidual) != 0),
                        __float_as_int(sc));
        }
      }
    }
    __syncthreads();
    for (int ee = threadIdx.x; ee < E; ee += blockDim.x) {
      int add = 0;
      for (int w = 0; w < kWarps; ++w) {
        add += s_wcnt[w * E + ee];
        s_wcnt[w * E + ee] = 0;
      }
      s_cursor[ee] += add;
    }
    // 4. row copy: warp w moves rows w, w+8, ... of the segment, kRowsPerBatch rows (x 16 B per lane) in flight
    //    (xbuf == nullptr: routing tables only -- moe_prepare_forward, trainer_3m_fix/fmoe/functions.py:13-52)
    if (xbuf != nullptr || kEp)
    for (int j0 = warp; j0 < n; j0 += kWarps * kRowsPerBatch) {
      int d[kRowsPerBatch];
      const InT* src[kRowsPerBatch];
#pragma unroll
      for (int r = 0; r < kRowsPerBatch; ++r) {
        const int j = j0 + r * kWarps;
        d[r] = j < n ? s_dst[j] : -1;
        src[r] = x + static_cast<size_t>((seg + (j < n ? j : 0)) / top_k) * D;
      }
      bf16* drow[kRowsPerBatch];
      bool drop[kRowsPerBatch];
#pragma unroll
      for (int r = 0; r < kRowsPerBatch; ++r) {
        drow[r] = xbuf + static_cast<size_t>(d[r] < 0 ? 0 : d[r]) * D;
        drop[r] = d[r] < 0;
        if (kEp && d[r] >= 0) {
          const int ex = s_exp[j0 + r * kWarps];
          drow[r] = reinterpret_cast<bf16*>(ep.base[ex / ep.E_local] + ep.lay.recv_x) +
                    static_cast<size_t>(s_base[ex] + d[r] - s_off[ex]) * D;
          if (!ep_ok) d[r] = -1;  // a peer's counts are missing: push nothing
        }
      }
      if constexpr (kXf32) {
        // fp32 rows: xbuf is really float [Sk, D]
        for (int v = lane; v < D / 4; v += 32) {
          float4 regs[kRowsPerBatch];
#pragma unroll
          for (int r = 0; r < kRowsPerBatch; ++r)
            if (d[r] >= 0) {
              regs[r] = __ldg(reinterpret_cast<const float4*>(src[r]) + v);
              // the tensor cores truncate fp32 words to TF32: round to nearest here instead
              regs[r].x = ptx::round_tf32(regs[r].x);
              regs[r].y = ptx::round_tf32(regs[r].y);
              regs[r].z = ptx::round_tf32(regs[r].z);
              regs[r].w = ptx::round_tf32(regs[r].w);
            }
#pragma unroll
          for (int r = 0; r < kRowsPerBatch; ++r)
            if (d[r] >= 0)
              reinterpret_cast<float4*>(reinterpret_cast<float*>(xbuf) + static_cast<size_t>(d[r]) * D)[v] = regs[r];
        }
      } else {
      for (int v = lane; v < D / 8; v += 32) {
        uint4 regs[kRowsPerBatch];
#pragma unroll
        for (int r = 0; r < kRowsPerBatch; ++r)
          if (d[r] >= 0) regs[r] = load8_as_bf16<InT>(src[r] + v * 8);
#pragma unroll
        for (int r = 0; r < kRowsPerBatch; ++r)
          if (d[r] >= 0) reinterpret_cast<uint4*>(drow[r])[v] = regs[r];
      }
      }
      if (drop_out != nullptr) {
        // dropped tokens (padding / invalid expert): output row = residual row (or zero)
#pragma unroll
        for (int r = 0; r < kRowsPerBatch; ++r) {
          const int j = j0 + r * kWarps;
          if (j < n && drop[r]) {
            // 16-byte accesses, four loads in flight before their stores (the two buffers may alias as far as the
            // compiler knows: element-wise this was one dependent L2 round trip per element)
            const size_t row = static_cast<size_t>(seg + j) / top_k * D;
            uint4* orow = reinterpret_cast<uint4*>(drop_out + row);
            const uint4* rrow = reinterpret_cast<const uint4*>(drop_residual + row);
            const int nv = D * static_cast<int>(sizeof(InT)) / 16;  // D % 8 == 0
            for (int c0 = 0; c0 < nv; c0 += 128) {
              uint4 v[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int c = c0 + u * 32 + lane;
                v[u] = (drop_residual != nullptr && c < nv) ? __ldg(rrow + c) : make_uint4(0u, 0u, 0u, 0u);
              }
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int c = c0 + u * 32 + lane;
                if (c < nv) orow[c] = v[u];
              }
            }
          }
        }
      }
    }
    __syncthreads();  // s_dst / s_wcnt / s_cursor are reused by the next segment
  }

  // 5. CTA 0 publishes counts / offsets / FFN group table
  if (blockIdx.x == 0) {
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
      counts[e] = s_total[e];
      if (counts_out) counts_out[e] = s_total[e];
    }
    for (int e = threadIdx.x; e <= E; e += blockDim.x) {
      offsets[e] = s_off[e];
      if (offsets_out) offsets_out[e] = s_off[e];
    }
    if (!kEp) build_groups_block(s_off, E, bn, groups, n_groups, h_ready, gmax, s_scratch);
    // this rank as an OWNER: every rank's counts are needed only now, for the group table over the received rows
    // (s_part is free by now)
    else {
      const bool ok = ep_wait_counts(ep, ep_seq, E, ep_mode, s_cnt);
      ep_build_groups_segmented(ep, E, s_cnt, ok, bn, groups, n_groups, h_ready, gmax, s_part);
    }
  }

  if (kEp) {
    // step 3: this CTA's rows (and routing data) have landed -- fence, then one increment on every rank.  The expert
    // kernels wait for the counters themselves, so nobody has to be the last CTA and nobody waits here.
    __syncthreads();
    ep_signal(ep, ep.lay.arrive);
    if (threadIdx.x == 0) {
      // the sequence number advances once every CTA of this launch has read it (a later wave's CTA may start long
      // after the first ones have finished)
      int* ctrl = ep_ctrl(ep);
      if (atomicAdd(&ctrl[1], 1) == static_cast<int>(gridDim.x) - 1) {
        ctrl[1] = 0;
        ctrl[0] = ep_seq;
      }
    }
  }
}

// Expert parallelism driven stage by stage (tests that emulate several ranks on one GPU: a kernel must never wait for a
// kernel that is queued behind it): the counts alone, from the chunk histograms, without the scatter.
__global__ void __launch_bounds__(kDispatchThreads)
dispatch_ep_counts_kernel(const int* __restrict__ chunk_hist, int hist_rows, int E, const EpPeers ep, int ep_mode,
                          int push_ctas, int ffn_ctas) {
  extern __shared__ int s_dyn[];
  int* s_total = s_dyn;
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    int t = 0;
    for (int c = 0; c < hist_rows; ++c) t += chunk_hist[c * E + e];
    s_total[e] = t;
  }
  __syncthreads();
  ep_send_counts(ep, ep_ctrl(ep)[0] + 1, s_total, E, ep_mode, push_ctas, ffn_ctas, s_total + E);
}

__global__ void __launch_bounds__(256)
build_groups_kernel(const int* __restrict__ offsets, int E, int bn, GroupRec* groups, int* n_groups, int* h_ready,
                    int gmax) {
  __shared__ int s_off[kMaxExperts + 1];
  __shared__ int s_scratch[kMaxExperts + 1];
  for (int e = threadIdx.x; e <= E; e += blockDim.x) s_off[e] = offsets[e];
  __syncthreads();
  build_groups_block(s_off, E, bn, groups, n_groups, h_ready, gmax, s_scratch);
}

}  // namespace

// Token-tile width of the FFN kernel: the smallest of 32/64/128/256 that holds ~1.25x the mean tokens per expert, so a
// balanced router gives one tile per expert and weights are streamed once -- but 256 (CTA pairs) only when there are
// rows enough for several waves of such tiles.  Below that the kernel is a latency chain (first GEMM -> h -> second GEMM
// per group), a 256-token tile's K loops are twice as long and 74 pairs take half as many tiles at a time: at 3 200 rows
// over 16 experts, 128-token tiles finish in about half the time of 256-token pair tiles.
int choose_bn(int Sk, int E) {
  const long long need = (static_cast<long long>(Sk) * 5 + 4LL * E - 1) / (4LL * E);
  if (need <= 32) return 32;
  if (need <= 64) return 64;
  if (need <= 128 || Sk < 12288) return 128;
  return 256;
}

cudaError_t launch_dispatch(const void* x, const int* idx, const float* score, int S, int D, int E, int top_k,
                            int dtype, int bn, const RouteWs& ws, int* counts_out, int* offsets_out,
                            int* mapping_out, bf16* xbuf, void* drop_out, const void* drop_residual,
                            const int* hist32, cudaStream_t stream, const EpPeers* ep, bool ep_fold_wait,
                            bool xbuf_f32, int ep_mode, int ep_phase, int ep_ffn_ctas) {
  if (xbuf_f32 && (dtype != B200MOE_F32 || ep != nullptr || D % 4 != 0)) return cudaErrorInvalidValue;
  const int Sk = S * top_k;
  if (top_k != 1) drop_out = nullptr;  // (expert parallelism: only the folded path passes one)
  if (ep != nullptr && (ep->world * ep->E_local != E || Sk > ep->cap || ep->D != D)) return cudaErrorInvalidValue;
  if (E > kMaxExperts || E < 1 || D % 8 != 0) return cudaErrorInvalidValue;
  const int gmax = ep ? max_groups(ep->world * ep->cap, ep->world * ep->E_local, bn) : max_groups(Sk, E, bn);
  if (Sk == 0 && ep == nullptr) {
    // nothing to route: still publish zero counts / offsets and an empty group table
    cudaError_t e = cudaMemsetAsync(ws.counts, 0, sizeof(int) * E, stream);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(ws.offsets, 0, sizeof(int) * (E + 1), stream);
    if (e != cudaSuccess) return e;
    if (counts_out) cudaMemsetAsync(counts_out, 0, sizeof(int) * E, stream);
    if (offsets_out) cudaMemsetAsync(offsets_out, 0, sizeof(int) * (E + 1), stream);
    return cudaMemsetAsync(ws.n_groups, 0, sizeof(int), stream);
  }
  const Chunking ck = make_chunking(S, top_k);
  const int rows32 = (S + 31) / 32;
  const int* hist = ws.chunk_hist;
  int hist_rows = ck.nchunks;
  int rows_per_chunk = 1;
  if (hist32 != nullptr && rows32 <= kMaxHistRows) {
    hist = hist32;  // counts came with the gate: no count kernel
    hist_rows = rows32;
    rows_per_chunk = ck.chunk_tok / 32;
  } else if (ep_phase != 2) {   // (phase 2 = scatter only: the histograms are there from phase 1)
    dispatch_count_kernel<<<ck.nchunks, kDispatchThreads, 0, stream>>>(idx, Sk, E, ck.chunk, ws.chunk_hist);
    count_launch();
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return err;
  }
  const int nparts_max = 8;
  const size_t dyn = sizeof(int) * (2 * E + 2 * (E + 1) + kDispatchThreads + (kDispatchThreads / 32) * E +
                                    nparts_max * 2 * E + kDispatchThreads + kMaxEpWorld * (E + kEpCntExtra) + E);
  cudaError_t lerr = cudaSuccess;
  EpPeers epv{};
  if (ep) epv = *ep;
  if (ep != nullptr && ep_phase == 1) {
    // counts only (ranks emulated on one GPU are driven phase by phase: nobody may wait for a kernel queued behind it)
    dispatch_ep_counts_kernel<<<1, kDispatchThreads, sizeof(int) * (E + 2), stream>>>(hist, hist_rows, E, epv, ep_mode,
                                                                                     ck.nchunks, ep_ffn_ctas);
    count_launch();
    return cudaGetLastError();
  }
  const int ep_send = ep_phase == 2 ? 0 : 1;
#define B200MOE_SCATTER_K(T, EP)                                                                                  \
  lerr = launch_kernel(dispatch_scatter_kernel<T, EP, false>, dim3(ck.nchunks), dim3(kDispatchThreads), dyn, stream, \
      kPdlDispatch,                                                                                               \
      static_cast<const T*>(x), idx, score, Sk, D, E, top_k, ck.chunk, ck.nchunks, hist, hist_rows,               \
      rows_per_chunk, bn, gmax,                                                                                   \
      ws.counts, ws.offsets, ws.mapping, ws.pos, ws.row_score, xbuf, ws.groups, ws.n_groups, ws.h_ready,          \
      counts_out, offsets_out, mapping_out, static_cast<T*>(drop_out), static_cast<const T*>(drop_residual),      \
      (pdl_trigger() & kPdlDispatch) ? 1 : 0, epv, ep_fold_wait ? 1 : 0, ep_mode, ep_send, ep_ffn_ctas)
#define B200MOE_SCATTER(T)          \
  if (ep)                           \
    B200MOE_SCATTER_K(T, true);     \
  else                              \
    B200MOE_SCATTER_K(T, false)
  switch (dtype) {
    case B200MOE_F32:
      if (xbuf_f32) {
        lerr = launch_kernel(dispatch_scatter_kernel<float, false, true>, dim3(ck.nchunks), dim3(kDispatchThreads), dyn,
                             stream, kPdlDispatch, static_cast<const float*>(x), idx, score, Sk, D, E, top_k, ck.chunk,
                             ck.nchunks, hist, hist_rows, rows_per_chunk, bn, gmax, ws.counts, ws.offsets, ws.mapping,
                             ws.pos, ws.row_score, xbuf, ws.groups, ws.n_groups, ws.h_ready, counts_out, offsets_out,
                             mapping_out, static_cast<float*>(drop_out), static_cast<const float*>(drop_residual),
                             (pdl_trigger() & kPdlDispatch) ? 1 : 0, epv, 0, 0, 0, 0);
        break;
      }
      B200MOE_SCATTER(float);
      break;
    case B200MOE_F16:
      B200MOE_SCATTER(__half);
      break;
    case B200MOE_BF16:
      B200MOE_SCATTER(bf16);
      break;
    default:
      return cudaErrorInvalidValue;
  }
#undef B200MOE_SCATTER
#undef B200MOE_SCATTER_K
  count_launch();
  return lerr;
}

cudaError_t launch_build_groups(const int* offsets, int E, int bn, GroupRec* groups, int* n_groups, int* h_ready,
                                int gmax, cudaStream_t stream) {
  if (E > kMaxExperts) return cudaErrorInvalidValue;
  build_groups_kernel<<<1, 256, 0, stream>>>(offsets, E, bn, groups, n_groups, h_ready, gmax);
  count_launch();
  return cudaGetLastError();
}

}  // namespace b200moe
