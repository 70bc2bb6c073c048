// Fused gate + dispatch for small batches ("route" kernel): router GEMM + softmax / top-1 and the stable scatter of the
// token rows in ONE launch, separated by a grid-wide barrier instead of a kernel boundary.
//
// Reference behaviour covered (same as gate_tc.cu + dispatch.cu): router MatMul + SoftmaxTopKPluginDynamic
// (trainer_3m_fix/layer/positionwise_feed_forward.py:169-207,225; TRTAPI++/plugin/softmax_topk_plugin/
// softmax_topk_kernel.cu:26-120), NaiveGate with k = 1 (trainer_3m_fix/fmoe/gates.py:51-66), ScatterMapping +
// ScatterMappingCopy (TRTAPI++/plugin/fmoe_expert_plugin/fmoe_expert_kernel.cu:25-128), moe_prepare_forward + MOEScatter
// (trainer_3m_fix/fmoe/functions.py:13-52,62-86).
//
// Why: at the 3M-ASR batch sizes (50 .. a few thousand tokens per layer) gate and dispatch each move a few MB but cost
// ~9 us apiece in launch / prologue / dependent-load latency, and the gate's 128-token tiles occupy only S/128 SMs, each
// of which then has to pull 384 KB through one SM's memory port.  Here a tile is 32 tokens, so 3 200 tokens spread over
// 100 SMs, and the token rows a CTA scatters are the ones it has just read.
//
//   phase 1 (per 32-token tile): "swap-AB" router GEMM on tcgen05 -- A = packed router, hi rows 0..31 / lo rows 32..63
//     of a 128-row UMMA tile whose upper 64 rows are never loaded (their accumulator lanes are garbage and never read),
//     B = 32 tokens x K, D[expert lane, token column] in TMEM.  Two warps transpose hi / lo through shared memory, then
//     one thread per token does softmax / arg-max in registers exactly like gate_tc_kernel and the warp emits the tile's
//     expert histogram.
//   grid barrier (all CTAs are co-resident: grid <= number of SMs, one CTA per SM)
//   phase 2 (per 32-token chunk): prefix over the histograms -> offsets, stable ranks by match_any, mapping / pos /
//     row_score, 128-bit row copies into expert order (or, under expert parallelism, into the owner GPU's receive buffer).
//     CTA 0 publishes counts / offsets / the FFN group table.
//   route_kernel<.., kLn = true>: the Conformer block's norm_ff (trainer_3m_fix/layer/fmoe_transformer.py:145-148) folded
//     into the router algebraically, see ln_stats_in_ring below.
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <random>

#include <math_constants.h>

#include "common.cuh"
#include "ep_device.cuh"
#include "ln_device.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

namespace b200moe {

namespace {

constexpr int kTok = 32;                    // tokens per tile == UMMA N == dispatch chunk
constexpr int kRK = 64;                     // k-block
constexpr int kHalfKb = 8;                  // k-blocks per pipeline slot: one TMA instruction per operand per slot
constexpr int kRABlk = 64 * kRK * 2;        // 8 KiB: 64 packed router rows (hi | lo) of one k-block
constexpr int kRBBlk = kTok * kRK * 2;      // 4 KiB: 32 tokens of one k-block
constexpr int kRSlotA = kHalfKb * kRABlk;   // 64 KiB
constexpr int kRSlotB = kHalfKb * kRBBlk;   // 32 KiB  (also the slack the 128-row UMMA reads past the last A block)
constexpr int kRSlot = kRSlotA + kRSlotB;   // 96 KiB
constexpr int kRSlots = 2;
constexpr int kRThreads = 256;
constexpr uint32_t kRTmemCols = 64;         // 2 accumulator buffers x 32 token columns
constexpr int kLnWarps = 4;                 // warps 2, 3, 6, 7: row statistics and in-place normalisation when norm_ff is folded in
constexpr unsigned long long kRouteTimeoutMs = 2000;  // grid-barrier wait before the launch gives up (status word)
constexpr int kStatusRouteTimeout = 1;

// Device status word of the library's non-EP kernels (b200moe_status): 0 = ok.  A module-scope variable rather than a
// word in the caller's scratch, which nobody initialises.
__device__ int g_route_status = 0;

// norm_ff fused into the router ALGEBRAICALLY (kLn).  With mu, r the mean and reciprocal standard deviation of a token row,
//   LN(x) . Wr_x = r * ( x . W' - mu * c1 ) + c0,    W' = diag(gamma) Wr_x,  c1 = gamma^T Wr_x,  c0 = beta^T Wr_x,
// so the router MMAs run on the RAW rows against a pre-scaled router (b200moe_pack_router_ln) the moment the tile lands,
// into an accumulator of their own, while four otherwise idle warps compute mu and r from the same tile; the arg-max
// epilogue combines the embed accumulator, the x accumulator, mu, r, c1 and c0.  Normalising the rows in place in front of
// the MMAs instead put ~2.8 us of arithmetic (4 warps, 16 K elements) on the critical path of every tile.
// The rows the experts consume are normalised (and rounded to bf16) in place by the same warps once the MMAs have read
// the tile, off the critical path (CTAs with a second tile re-read and normalise the rows from global memory instead).
//
// The x operand in the ring: [k-block][32 rows][64 bf16], 128-byte swizzle (the 16-byte chunk c of row r sits at chunk
// c ^ (r & 7)).  Lane l owns the 8-element vectors l and l + 32 of a row (D <= 512 here) exactly like
// layernorm_rows_kernel, so the statistics have the same bits.
constexpr int kRowsPerLnWarp = kTok / kLnWarps;  // 8
constexpr int kLnBatch = 4;                       // rows a warp reduces together
constexpr int kRouteLnVec = 2;

__device__ __forceinline__ void ln_stats_in_ring(const uint8_t* sx, int r0, int D, int lane, float eps, float* s_mu,
                                                 float* s_rs) {
  const int nvec = D >> 3;
#pragma unroll 1
  for (int rb = r0; rb < r0 + kRowsPerLnWarp; rb += kLnBatch) {
    float v[kLnBatch][kRouteLnVec][8];
#pragma unroll
    for (int i = 0; i < kLnBatch; ++i) {
      const int r = rb + i;
#pragma unroll
      for (int k = 0; k < kRouteLnVec; ++k) {
        const int vec = k * 32 + lane;
        if (vec < nvec) {
          const uint4 w4 =
              *reinterpret_cast<const uint4*>(sx + (vec >> 3) * kRBBlk + r * 128 + (((vec & 7) ^ (r & 7)) << 4));
          const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            v[i][k][2 * j] = __uint_as_float(w[j] << 16);
            v[i][k][2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
          }
        }
      }
    }
    float mean[kLnBatch], rstd[kLnBatch];
    ln_rows_stats<kRouteLnVec, kLnBatch>(v, D, lane, eps, mean, rstd);
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < kLnBatch; ++i) {
        s_mu[rb + i] = mean[i];
        s_rs[rb + i] = rstd[i];
      }
    }
  }
}

// Normalise kRowsPerLnWarp rows in place in the ring (statistics already known); same expression as everywhere else.
__device__ __forceinline__ void ln_apply_in_ring(uint8_t* sx, int r0, int D, int lane, const float* s_mu,
                                                 const float* s_rs, const float* s_gamma, const float* s_beta) {
  const int nvec = D >> 3;
#pragma unroll 2
  for (int i = 0; i < kRowsPerLnWarp; ++i) {
    const int r = r0 + i;
    const float mu = s_mu[r], rs = s_rs[r];
#pragma unroll
    for (int k = 0; k < kRouteLnVec; ++k) {
      const int vec = k * 32 + lane;
      if (vec < nvec) {
        uint4* cell = reinterpret_cast<uint4*>(sx + (vec >> 3) * kRBBlk + r * 128 + (((vec & 7) ^ (r & 7)) << 4));
        const uint4 w4 = *cell;
        const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
        const float4 g0 = *reinterpret_cast<const float4*>(s_gamma + vec * 8);
        const float4 g1 = *reinterpret_cast<const float4*>(s_gamma + vec * 8 + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(s_beta + vec * 8);
        const float4 b1 = *reinterpret_cast<const float4*>(s_beta + vec * 8 + 4);
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        uint32_t o[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float v0 = ln_apply(__uint_as_float(w[u] << 16), mu, rs, g[2 * u], b[2 * u]);
          const float v1 = ln_apply(__uint_as_float(w[u] & 0xffff0000u), mu, rs, g[2 * u + 1], b[2 * u + 1]);
          __nv_bfloat162 pk = __floats2bfloat162_rn(v0, v1);
          o[u] = *reinterpret_cast<uint32_t*>(&pk);
        }
        *cell = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
  }
}

struct RouteParams {
  // gate
  const float* br;
  const int* x_len;
  int* idx;
  float* score;
  unsigned long long* hist64;  // [n_tiles, E] self-validating words: (launch tag << 8) | count, see the grid barrier
  int S, T, D, Demb, E, gate_mode;
  // dispatch
  const bf16* x;
  bf16* xbuf;
  int* counts;
  int* offsets;
  int* mapping;
  int* pos;
  float* row_score;
  GroupRec* groups;
  int* n_groups;
  int* h_ready;
  int bn, gmax;
  int* counts_out;
  int* offsets_out;
  int* mapping_out;
  bf16* drop_out;
  const bf16* drop_residual;
  int keep_expert_output;
  // grid barrier: two 64-bit state words {nonce, count}; see grid_arrive
  unsigned long long* bar;
  unsigned nonce;
  unsigned long long tag;  // 56 random bits (<< 8) drawn per launch: what makes a histogram word this launch's
  int ep_fold_wait;
  int ep_mode;   // kEpModeFold / kEpModeResidual bits of this call (checked against the peers')
  int ep_ffn_ctas;  // grid of the expert kernel that follows (announced with the counts)
  // norm_ff fused into the router (block call, route_kernel<.., kLn = true>); null = off
  const float* ln_gamma;
  const float* ln_beta;
  const float* ln_c;   // c1[32], c0[32] behind the pre-scaled packed router (b200moe_pack_router_ln)
  float ln_eps;
  int warm_mma;
  int table_cta;  // the CTA that writes counts / offsets / the expert kernel's group table: 0, or an extra CTA without a
                  // token tile (grid = tiles + 1) when an SM is free for it -- the table then is off CTA 0's tail
  uint4* trace;  // debug timeline: 16 records per CTA {event, clock64 lo, hi, -}; slots 14 / 15 hold %globaltimer
  unsigned long long* tl;  // cross-kernel timeline slot of this launch (common.cuh: set_timeline) or null
};

__device__ __forceinline__ void rtrace(const RouteParams& p, int ev) {
  if (p.trace != nullptr) {
    const unsigned long long c = static_cast<unsigned long long>(clock64());
    p.trace[blockIdx.x * 16 + ev] = make_uint4(static_cast<uint32_t>(ev), static_cast<uint32_t>(c),
                                               static_cast<uint32_t>(c >> 32), 1u);
  }
}
__device__ __forceinline__ void rtrace_sync(const RouteParams& p, int slot) {
  if (p.trace != nullptr) {
    const unsigned long long g = ptx::globaltimer_ns();
    p.trace[blockIdx.x * 16 + slot] = make_uint4(static_cast<uint32_t>(slot), static_cast<uint32_t>(g),
                                                 static_cast<uint32_t>(g >> 32), 1u);
  }
}

// ---- grid barrier that needs no zero-initialised memory ----------------------------------------------------------------
// state = (nonce << 32) | arrivals.  The first arrival of a launch finds a foreign nonce in the upper half (garbage, or the
// complement the previous launch left behind) and installs {nonce, 1} with a compare-and-swap; everybody else adds 1.
// The last CTA to finish the kernel overwrites both words with the complemented nonce, so a CUDA-graph replay (same
// nonce every time) starts clean as well.
__device__ __forceinline__ unsigned grid_arrive(unsigned long long* st, unsigned nonce) {
  unsigned long long old = *reinterpret_cast<volatile unsigned long long*>(st);
  while (true) {
    if (static_cast<unsigned>(old >> 32) == nonce) {
      const unsigned long long prev = atomicAdd(st, 1ull);
      return static_cast<unsigned>(prev & 0xffffffffu) + 1u;
    }
    const unsigned long long nw = (static_cast<unsigned long long>(nonce) << 32) | 1ull;
    const unsigned long long prev = atomicCAS(st, old, nw);
    if (prev == old) return 1u;
    old = prev;
  }
}

// kLn: the block's norm_ff is fused in (its own instantiation: the LayerNorm code doubles the kernel's registers and
// instruction footprint, which cost the plain kernel 1.2-1.9 us per layer when it was a run-time switch)
template <bool kEp, bool kLn>
__global__ void __launch_bounds__(kRThreads, 1)
route_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_e,
             const __grid_constant__ CUtensorMap tm_wx, const __grid_constant__ CUtensorMap tm_we, const RouteParams p,
             const EpPeers ep) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = ptx::smem_u32(smem_raw);
  if ((smem_base & 1023u) != 0) __trap();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  auto mark = [&](int i) {  // cross-kernel timeline (debug; off = one predicated branch)
    if (p.tl != nullptr) p.tl[blockIdx.x * kTimelineMarks + i] = ptx::globaltimer_ns();
  };
  if (threadIdx.x == 0) mark(0);
  const uint32_t bar_base = smem_base + kRSlots * kRSlot;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kRSlots + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kRSlots + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kRSlots + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kRSlots + 4);
  auto ln_bar = [&](int s) { return tmem_slot + 16u + 8u * s; };  // statistics warps -> softmax warp, per accumulator stage
  uint8_t* misc = smem_raw + (tmem_slot - smem_base) + 32;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(misc - 32);
  float* s_br = reinterpret_cast<float*>(misc);               // [32]
  float* s_hi = s_br + 32;                                     // [32][33]
  float* s_lo = s_hi + 32 * 33;                                // [32][33]
  int* s_hist = reinterpret_cast<int*>(s_lo + 32 * 33);       // [32]
  int* s_total = s_hist + 32;                                  // [32]  tokens per expert over all chunks
  int* s_before = s_total + 32;                                // [32]  ... over the chunks before the current one
  int* s_off = s_before + 32;                                  // [33]  exclusive offsets
  int* s_scratch = s_off + 33;                                 // [33]
  int* s_dst = s_scratch + 33;                                 // [32]  destination row of each token of the chunk
  int* s_exp = s_dst + 32;                                     // [32]
  int* s_part = s_exp + 32;                                    // [8][2][32]
  int* s_part2 = s_part + 512;                                 // [8][32]  prefix of this CTA's second tile
  int* s_before2 = s_part2 + 256;                              // [32]
  int* s_lastp = s_before2 + 32;                               // [1]
  // kLn only (the host sizes the allocation accordingly)
  float* s_hx = reinterpret_cast<float*>(s_lastp + 6);         // [32][33]  x-part accumulator, hi rows (16 B aligned)
  float* s_lx = s_hx + 32 * 33;                                // [32][33]  ... lo rows
  float* s_mu = s_lx + 32 * 33;                                // [2][32]   per accumulator stage: row means
  float* s_rs = s_mu + 64;                                     // [2][32]   ... reciprocal standard deviations
  float* s_c1 = s_rs + 64;                                     // [32]
  float* s_c0 = s_c1 + 32;                                     // [32]
  float* s_gamma = s_c0 + 32;                                  // [D]
  float* s_beta = s_gamma + kHalfKb * kRK;                     // [D]

  if (threadIdx.x == 0) {
    rtrace_sync(p, 14);
    rtrace(p, 0);
  }
  const int E = p.E;
  const int n_tiles = (p.S + kTok - 1) / kTok;
  // the K dimension comes in (up to) two parts: the embed columns and the x columns (the reference's concat,
  // positionwise_feed_forward.py:225); each part is one pipeline slot filled by two TMA instructions
  const int kb_e = p.Demb / kRK;
  const int kb_x = p.D / kRK;
  const int n_parts = kb_e > 0 ? 2 : 1;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tm_x);
    ptx::prefetch_tensormap(&tm_wx);
    if (kb_e > 0) {
      ptx::prefetch_tensormap(&tm_e);
      ptx::prefetch_tensormap(&tm_we);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kRSlots; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(tfull_bar(s), 1);
      ptx::mbar_init(tempty_bar(s), 2);
      if (kLn) ptx::mbar_init(ln_bar(s), kLnWarps);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) ptx::tmem_alloc<kLn ? 2 * kRTmemCols : kRTmemCols>(tmem_slot);
  if (warp == 3) s_br[lane] = (p.br != nullptr && lane < E) ? p.br[lane] : 0.0f;
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  if (threadIdx.x == 0) rtrace(p, 1);
  if (warp == 1 && lane == 0 && p.warm_mma && static_cast<int>(blockIdx.x) < n_tiles) {
    // the first tcgen05.mma of a kernel is ~2 us slower to issue than any later one (see ffn.cu): one throw-away
    // instruction over whatever the ring holds, into the first accumulator stage (overwritten by the first real MMA),
    // pays that while the tile is still on its way
    ptx::umma_f16_ss(tmem_base, ptx::make_kmajor_sw128_desc(smem_base), ptx::make_kmajor_sw128_desc(smem_base + kRSlotA),
                     ptx::make_idesc(1u, 128, kTok), 0u);
  }

  // ================================================ phase 1: gate ================================================
  if (warp == 0) {
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      bool waited = false;
      int it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        for (int part = 0; part < n_parts; ++part) {
          const bool is_e = n_parts == 2 && part == 0;
          const int nkb = is_e ? kb_e : kb_x;
          if constexpr (kLn) {
            // The MMAs release a ring slot when THEY have read it, but the statistics warps read the x tile as well
            // and take longer: a second tile's rows must not land on top of the first tile's before its statistics
            // are done (ln_bar of that tile; with two K parts the x tiles of consecutive tiles share a slot, and a
            // CTA has at most two tiles).
            if (!is_e && it > 0) ptx::mbar_wait(ln_bar((it - 1) & 1), ((it - 1) >> 1) & 1);
          }
          // Programmatic dependent launch: the embed half of the router GEMM (embed and the router are constants) runs
          // while the previous layer's FFN kernel is still draining; x, its output, is only touched after this wait.
          ptx::mbar_wait(empty_bar(slot), phase ^ 1u);
          ptx::mbar_arrive_expect_tx(full_bar(slot), static_cast<uint32_t>(nkb) * (kRABlk + kRBBlk));
          const uint32_t sa = smem_base + slot * kRSlot;
          // router k-blocks of this part: rows 0..63 of the packed router, k-blocks [kb0, kb0 + nkb)
          ptx::tma_load_3d(sa, is_e ? &tm_we : &tm_wx, full_bar(slot), 0, 0, is_e ? 0 : kb_e, ptx::kEvictLast);
          // (The x half of the router is a constant as well: it is on its way before the wait, only the 32 KiB of token
          // rows are requested behind it.)
          if (!is_e && !waited) {
            ptx::pdl_wait();
            mark(1);
            waited = true;
          }
          // the tile's 32 tokens, all k-blocks of the part
          if (is_e)
            ptx::tma_load_3d(sa + kRSlotA, &tm_e, full_bar(slot), 0, t * kTok, 0, ptx::kEvictFirst);
          else
            ptx::tma_load_3d(sa + kRSlotA, &tm_x, full_bar(slot), 0, t * kTok, 0, ptx::kEvictLast);
          if (++slot == kRSlots) {
            slot = 0;
            phase ^= 1u;
          }
        }
      }
      rtrace(p, 2);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = ptx::make_idesc(1u /*bf16*/, 128, kTok);
      int slot = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        ptx::mbar_wait(tempty_bar(as), aphase ^ 1u);
        ptx::tc_fence_after();
        for (int part = 0; part < n_parts; ++part) {
          const int nkb = (n_parts == 2 && part == 0) ? kb_e : kb_x;
          // kLn: the x part accumulates on its own (columns 32..63 of the stage's 64), it is rescaled per token later
          const bool own_acc = kLn && part == n_parts - 1;
          const uint32_t tmem_d = tmem_base + as * (kLn ? 2 * kTok : kTok) + (own_acc ? kTok : 0);
          ptx::mbar_wait(full_bar(slot), phase);
          ptx::tc_fence_after();
          const uint32_t sa = smem_base + slot * kRSlot;
          for (int j = 0; j < nkb; ++j) {
            // A: 64 loaded rows + 64 rows of whatever follows (accumulator lanes 64..127 are never read)
            const uint64_t a_desc = ptx::make_kmajor_sw128_desc(sa + j * kRABlk);
            const uint64_t b_desc = ptx::make_kmajor_sw128_desc(sa + kRSlotA + j * kRBBlk);
#pragma unroll
            for (int k = 0; k < kRK / 16; ++k)
              ptx::umma_f16_ss(tmem_d, a_desc + 2u * k, b_desc + 2u * k, idesc,
                               ((own_acc ? 0 : part) | j | k) != 0 ? 1u : 0u);
          }
          ptx::umma_commit(empty_bar(slot));
          if (part == n_parts - 1) ptx::umma_commit(tfull_bar(as));
          if (++slot == kRSlots) {
            slot = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (kLn && (warp == 2 || warp == 3 || warp >= 6)) {
    // norm_ff's statistics (fmoe_transformer.py:145-148) of the tile's 32 token rows, 8 rows per warp, while the MMAs run
    const int lw = (warp & 1) | ((warp >> 2) << 1);  // warps 2, 3, 6, 7 -> 0 .. 3
    // gamma / beta into shared memory while the tile is on its way.  Every one of the four warps writes the whole
    // vectors (identical values), so that none of them depends on another one's stores.
    for (int i = lane * 4; i < p.D; i += 128) {
      *reinterpret_cast<float4*>(s_gamma + i) = __ldg(reinterpret_cast<const float4*>(p.ln_gamma + i));
      *reinterpret_cast<float4*>(s_beta + i) = __ldg(reinterpret_cast<const float4*>(p.ln_beta + i));
    }
    __syncwarp();
    int slot = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      for (int part = 0; part < n_parts; ++part) {
        if (part == n_parts - 1) {
          ptx::mbar_wait(full_bar(slot), phase);  // the TMA's bytes are there (a CTA has at most two tiles: the stage
                                                  // it & 1 of s_mu / s_rs is not in use any more)
          ln_stats_in_ring(smem_raw + slot * kRSlot + kRSlotA, lw * kRowsPerLnWarp, p.D, lane, p.ln_eps,
                           s_mu + (it & 1) * 32, s_rs + (it & 1) * 32);
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(ln_bar(it & 1));  // release.cta: the softmax warp waits on it
          if (n_tiles <= static_cast<int>(gridDim.x)) {
            // One tile per CTA: the ring keeps the tile until phase 2 copies the rows out.  Once the MMAs have read
            // it (accumulator complete), normalise the rows in place -- while the arg-max epilogue and the wait for the
            // other CTAs' histograms run -- so that phase 2 is the plain copy it is without a LayerNorm.
            ptx::mbar_wait(tfull_bar(it & 1), (it >> 1) & 1);
            ln_apply_in_ring(smem_raw + slot * kRSlot + kRSlotA, lw * kRowsPerLnWarp, p.D, lane, s_mu + (it & 1) * 32,
                             s_rs + (it & 1) * 32, s_gamma, s_beta);
          }
        }
        if (++slot == kRSlots) {
          slot = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 4 || warp == 5) {
    // warp 4: TMEM lanes 0..31 = hi logits of experts 0..31; warp 5: lanes 32..63 = lo parts
    const int q = warp & 3;
    float* dstm = q == 0 ? s_hi : s_lo;
    int it = 0;
    if constexpr (kLn) {
      if (q == 0) {  // constants of the layer, read back by this warp only
        s_c1[lane] = p.ln_c[lane];
        s_c0[lane] = p.ln_c[32 + lane];
        __syncwarp();
      }
    }
    ptx::pdl_wait();  // idx / score / histogram words may still be read by the previous layer's kernels
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      ptx::mbar_wait(tfull_bar(as), aphase);
      ptx::tc_fence_after();
      if (q == 0 && lane == 0 && it == 0) rtrace(p, 3);
      uint32_t r[32];
      const uint32_t tacc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * (kLn ? 2 * kTok : kTok);
      if (!kLn || n_parts == 2) {  // (kLn without an embed part: the first accumulator was never written)
        ptx::tmem_ld_32x32b_x32(tacc, r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) dstm[lane * 33 + j] = __uint_as_float(r[j]);
      }
      if constexpr (kLn) {
        float* dstx = q == 0 ? s_hx : s_lx;
        ptx::tmem_ld_32x32b_x32(tacc + kTok, r);  // the x part's own accumulator
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) dstx[lane * 33 + j] = __uint_as_float(r[j]);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(tempty_bar(as));
      if (q == 0) s_hist[lane] = 0;
      ptx::named_bar_sync(1, 64);
      if (q == 0) {
        // one thread per token: softmax / arg-max in registers (same arithmetic and tie rule as gate_tc_kernel)
        const int tok = t * kTok + lane;
        bool valid = tok < p.S;
        if (valid && p.x_len != nullptr) valid = (tok % p.T) < p.x_len[tok / p.T];
        float l[32];
        if constexpr (kLn) {
          // logits = embed part + r * (x . W' - mu * c1) + c0  (see ln_stats_in_ring)
          ptx::mbar_wait(ln_bar(as), aphase);  // acquire.cta: the statistics warps' s_mu / s_rs of this tile
          const float mu = s_mu[as * 32 + lane], rs = s_rs[as * 32 + lane];
          const bool has_e = n_parts == 2;
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const float le = has_e ? s_hi[e * 33 + lane] + s_lo[e * 33 + lane] : 0.0f;
            const float lx = s_hx[e * 33 + lane] + s_lx[e * 33 + lane];
            l[e] = le + fmaf(rs, fmaf(-mu, s_c1[e], lx), s_c0[e]) + s_br[e];
            if (e >= E) l[e] = -CUDART_INF_F;
          }
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            l[e] = s_hi[e * 33 + lane] + s_lo[e * 33 + lane] + s_br[e];
            if (e >= E) l[e] = -CUDART_INF_F;
          }
        }
        float bv = l[0];
        int bi = 0;
#pragma unroll
        for (int e = 1; e < 32; ++e) {
          if (l[e] > bv) {  // strict: the lowest index wins exact ties
            bv = l[e];
            bi = e;
          }
        }
        float denom = 1.0f;
        if (p.gate_mode == B200MOE_GATE_3M) {
          denom = 0.0f;
#pragma unroll
          for (int e = 0; e < 32; ++e) denom += expf(l[e] - bv);  // exp(-inf) = 0 for the padded experts
        }
        const int sel = valid ? bi : -1;
        if (tok < p.S) {
          p.idx[tok] = sel;
          p.score[tok] = valid ? expf(bv - bv) / denom : 0.0f;
        }
        const unsigned peers = __match_any_sync(0xffffffffu, sel);
        if (sel >= 0 && lane == __ffs(peers) - 1) s_hist[sel] = __popc(peers);
        __syncwarp();
        // Every histogram word validates itself: (56-bit launch tag << 8) | count, one 64-bit store.  Other CTAs need
        // nothing else from this tile (idx / score are re-read by this CTA only), so no fence and no release store
        // stand between the arg-max and the other CTAs seeing the row.
        if (lane < E)
          p.hist64[static_cast<size_t>(t) * E + lane] = p.tag | static_cast<unsigned long long>(s_hist[lane]);
        if (lane == 0 && it == 0) rtrace(p, 4);
      }
      ptx::named_bar_sync(1, 64);  // s_hi / s_lo / s_hist free for the next tile
    }
  }

  // ================================================ grid barrier + prefix ================================================
  // There is no counter to contend on and no separate wait: every histogram word carries this launch's tag, every CTA
  // needs every row for its prefix sums anyway, so it simply re-reads a word until the tag is there.  The workspace is
  // caller-owned scratch that nobody clears: the tag is 56 RANDOM bits drawn per launch, so that whatever the words
  // hold beforehand (another layout's mapping / counts, another layer's data, an earlier launch's words) does not pass
  // for a published tile, and a count above the 32 tokens of a tile is refused as well.  The CTAs of the grid are meant
  // to be co-resident (grid <= SM count, one CTA per SM); when a foreign kernel holds SMs they become resident as it
  // drains, and a wait that outlasts the deadline raises the device status word (b200moe_status) instead of trapping:
  // the launch then completes over whatever counts were seen -- wrong rows, but every index stays in range.
  __syncthreads();
  ptx::pdl_wait();  // (every thread: phase 2 reads and writes the workspace and the row buffers)
  if (threadIdx.x == 0) mark(2);
  if (threadIdx.x == 0) rtrace(p, 5);
  {
    const int e = threadIdx.x & 31;
    const int part = threadIdx.x >> 5;  // 8 parts
    const int c_first = blockIdx.x;
    const unsigned long long tag = p.tag;
    const unsigned long long deadline = ptx::globaltimer_ns() + 1000000ull * kRouteTimeoutMs;
    bool gave_up = false;
    // `before2`: the same for this CTA's second tile (a CTA has at most two: route_supported), so that the second chunk
    // needs no further pass over the histograms (148 dependent-latency loads per expert when it had one)
    const int c_second = c_first + static_cast<int>(gridDim.x);
    int total = 0, before = 0, before2 = 0;
    if (e < E)
      for (int r0 = part; r0 < n_tiles; r0 += 8 * 8) {
        // eight rows per batch: the loads are independent and all in flight together (one L2 round trip per batch
        // when the rows are already there), only a word whose tag is still missing is polled on its own
        unsigned long long w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = r0 + 8 * i;
          w[i] = r < n_tiles ? *(const volatile unsigned long long*)(p.hist64 + static_cast<size_t>(r) * E + e) : tag;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = r0 + 8 * i;
          while ((w[i] & ~0xffull) != tag || (w[i] & 0xffull) > static_cast<unsigned long long>(kTok)) {
            if (gave_up || ptx::globaltimer_ns() > deadline) {
              gave_up = true;
              w[i] = tag;  // count 0
              break;
            }
            __nanosleep(20);
            w[i] = *(const volatile unsigned long long*)(p.hist64 + static_cast<size_t>(r) * E + e);
          }
          const int v = static_cast<int>(w[i] & 0xffull);
          total += v;
          if (r < c_first) before += v;
          if (r < c_second) before2 += v;
        }
      }
    if (gave_up) atomicExch(&g_route_status, kStatusRouteTimeout);
    s_part[part * 32 + e] = total;
    s_part[256 + part * 32 + e] = before;
    s_part2[part * 32 + e] = before2;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    int tot = 0, bef = 0, bef2 = 0;
#pragma unroll
    for (int pt = 0; pt < 8; ++pt) {
      tot += s_part[pt * 32 + threadIdx.x];
      bef += s_part[256 + pt * 32 + threadIdx.x];
      bef2 += s_part2[pt * 32 + threadIdx.x];
    }
    tot = static_cast<int>(threadIdx.x) < E ? tot : 0;
    s_total[threadIdx.x] = tot;
    s_before[threadIdx.x] = bef;
    s_before2[threadIdx.x] = bef2;
    // exclusive scan over the experts within the warp
    int incl = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, incl, d);
      if (static_cast<int>(threadIdx.x) >= d) incl += n;
    }
    s_off[threadIdx.x] = incl - tot;
    if (threadIdx.x == 31) s_off[32] = incl;
    if (static_cast<int>(threadIdx.x) == E - 1) s_off[E] = incl;
  }
  __syncthreads();
  if (threadIdx.x == 0) rtrace(p, 6);
  // Expert parallelism, steps 1 and 2 of the protocol (common.cuh): CTA 0 tells every rank how many rows this rank has for
  // each expert; where this rank's rows belong in the owners' receive buffers follows from its own offsets.
  // s_part is free again: [0, W * (E + 1)) count matrix, then s_base[E], then scratch for the group table.
  int* s_cnt = s_part;
  int* s_base = s_part + kMaxEpWorld * (32 + kEpCntExtra);
  int ep_seq = 0;
  bool ep_ok = true;
  if (kEp) {
    ep_seq = ep_ctrl(ep)[0] + 1;   // (only this kernel's last CTA advances the word, after every CTA has read it)
    if (blockIdx.x == 0) ep_send_counts(ep, ep_seq, s_total, E, p.ep_mode, gridDim.x, p.ep_ffn_ctas, s_dst);
    ep_local_bases(ep, E, s_off, s_base);  // (no wait: rows go into this rank's own segment of the owners' buffers)
  }
  if (threadIdx.x == 0) rtrace(p, 7);

  for (int c = blockIdx.x; c < n_tiles; c += gridDim.x) {
    if (c != static_cast<int>(blockIdx.x)) {
      // second chunk of this CTA (more tiles than SMs): its prefix was accumulated in the same pass
      if (threadIdx.x < 32) s_before[threadIdx.x] = s_before2[threadIdx.x];
      __syncthreads();
    }
    if (warp == 0) {
      const int tok = c * kTok + lane;
      int e = -1;
      if (tok < p.S) {
        e = p.idx[tok];
        if (e < 0 || e >= E) e = -1;
      }
      const unsigned peers = __match_any_sync(0xffffffffu, e);
      int dst = -1;
      if (e >= 0) dst = s_off[e] + s_before[e] + __popc(peers & ((1u << lane) - 1u));
      s_dst[lane] = dst;
      s_exp[lane] = e;
      if (tok < p.S) {
        p.mapping[tok] = dst;
        if (p.mapping_out) p.mapping_out[tok] = dst;
        if (dst >= 0) {
          p.pos[dst] = tok;
          const float sc = p.keep_expert_output ? 1.0f : p.score[tok];
          p.row_score[dst] = sc;
          if (kEp && ep_ok) {
            // routing data travels with the row: where the owner sends the result (the token's output row when the
            // combine is folded into the owner's epilogue, else this rank's expert-order row) and the gate score
            const int owner = e / ep.E_local;
            const int slot = s_base[e] + dst - s_off[e];
            int2* meta = reinterpret_cast<int2*>(ep.base[owner] + ep.lay.meta);
            meta[slot] = make_int2(ep_meta_word(ep.rank, (p.ep_mode & kEpModeFold) ? tok : dst,
                                                (p.ep_mode & kEpModeResidual) != 0), __float_as_int(sc));
          }
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0 && c == static_cast<int>(blockIdx.x)) rtrace(p, 8);
    if (n_tiles <= static_cast<int>(gridDim.x)) {
      // One tile per CTA: the 32 token rows are still in the ring (the B operand of the x part, 128B-swizzled K-major
      // blocks of 32 rows), so they are copied out of shared memory instead of being fetched from L2 a second time.
      // Lane l of warp w: row 4w + l/8, 16-byte chunk l%8 of every k-block (4 rows x 128 B per instruction, no bank
      // conflicts: the XOR swizzle permutes chunks inside a row's 128 bytes).
      const int r = warp * 4 + (lane >> 3);
      const int ch = lane & 7;
      const int tok = c * kTok + r;
      const int d = tok < p.S ? s_dst[r] : -1;
      if (d >= 0 && (!kEp || ep_ok)) {
        bf16* drow = p.xbuf + static_cast<size_t>(d) * p.D;
        if (kEp) {
          const int ex = s_exp[r];
          drow = reinterpret_cast<bf16*>(ep.base[ex / ep.E_local] + ep.lay.recv_x) +
                 static_cast<size_t>(s_base[ex] + d - s_off[ex]) * p.D;
        }
        const uint8_t* sx = smem_raw + (n_parts - 1) * kRSlot + kRSlotA + r * 128 + ((ch ^ (r & 7)) << 4);
        // (kLn: the statistics warps have normalised the rows in place by now)
#pragma unroll 4
        for (int j = 0; j < kb_x; ++j)
          reinterpret_cast<uint4*>(drow)[j * 8 + ch] = *reinterpret_cast<const uint4*>(sx + j * kRBBlk);
      } else if (d < 0 && tok < p.S && p.drop_out != nullptr) {
        // dropped token (padding): output row = residual row (or zero); the fused FFN epilogue never touches it
        // (16-byte accesses, every load of the row issued before its stores: the compiler cannot prove that the two
        // buffers do not alias, and one dependent L2 round trip per element made a padded batch 4x slower)
        const size_t row = static_cast<size_t>(tok) * p.D;
        uint4* orow = reinterpret_cast<uint4*>(p.drop_out + row);
        const uint4* rrow = reinterpret_cast<const uint4*>(p.drop_residual + row);
        uint4 v[kHalfKb];
#pragma unroll
        for (int j = 0; j < kHalfKb; ++j)
          v[j] = (p.drop_residual != nullptr && j < kb_x) ? __ldg(rrow + j * 8 + ch) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int j = 0; j < kHalfKb; ++j)
          if (j < kb_x) orow[j * 8 + ch] = v[j];
      }
    } else
    // row copies: warp w moves rows w, w + 8, w + 16, w + 24 of the chunk, all four in flight (2 x 16 B per lane each)
    {
      const bf16* src[4];
      bf16* drow[4];
      int d[4];
      bool drop[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int j = warp + r * 8;
        const int tok = c * kTok + j;
        d[r] = tok < p.S ? s_dst[j] : -1;
        src[r] = p.x + static_cast<size_t>(tok < p.S ? tok : 0) * p.D;
        drow[r] = p.xbuf + static_cast<size_t>(d[r] < 0 ? 0 : d[r]) * p.D;
        drop[r] = tok < p.S && d[r] < 0;
        if (kEp && d[r] >= 0) {
          const int ex = s_exp[j];
          drow[r] = reinterpret_cast<bf16*>(ep.base[ex / ep.E_local] + ep.lay.recv_x) +
                    static_cast<size_t>(s_base[ex] + d[r] - s_off[ex]) * p.D;
          if (!ep_ok) d[r] = -1;  // a peer's counts are missing: push nothing
        }
      }
      if constexpr (kLn) {
        // more tiles than SMs: the ring has moved on, the rows come from global memory and are normalised again on
        // the way (same lane <-> vector assignment as in the ring, hence the same bits the router saw)
        const int nvec = p.D >> 3;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          if (d[r] < 0) continue;  // (warp-uniform)
          float v[kLnMaxVec][8];
#pragma unroll
          for (int k = 0; k < kLnMaxVec; ++k)
            if (k * 32 + lane < nvec) {
              const uint4 w4 = __ldg(reinterpret_cast<const uint4*>(src[r]) + k * 32 + lane);
              const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                v[k][2 * i] = __uint_as_float(w[i] << 16);
                v[k][2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
              }
            }
          ln_row_registers<kLnMaxVec>(v, p.D, lane, p.ln_gamma, p.ln_beta, p.ln_eps);
#pragma unroll
          for (int k = 0; k < kLnMaxVec; ++k)
            if (k * 32 + lane < nvec) {
              uint32_t w[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                __nv_bfloat162 pk = __floats2bfloat162_rn(v[k][2 * i], v[k][2 * i + 1]);
                w[i] = *reinterpret_cast<uint32_t*>(&pk);
              }
              reinterpret_cast<uint4*>(drow[r])[k * 32 + lane] = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
      } else
      for (int v = lane; v < p.D / 8; v += 32) {
        uint4 regs[4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
          if (d[r] >= 0) regs[r] = __ldg(reinterpret_cast<const uint4*>(src[r]) + v);
#pragma unroll
        for (int r = 0; r < 4; ++r)
          if (d[r] >= 0) reinterpret_cast<uint4*>(drow[r])[v] = regs[r];
      }
      if (p.drop_out != nullptr) {
        // dropped tokens (padding): output row = residual row (or zero); the fused FFN epilogue never touches them
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int tok = c * kTok + warp + r * 8;
          if (drop[r]) {
            const size_t row = static_cast<size_t>(tok) * p.D;
            uint4* orow = reinterpret_cast<uint4*>(p.drop_out + row);
            const uint4* rrow = reinterpret_cast<const uint4*>(p.drop_residual + row);
            uint4 v[kRouteLnVec];   // D <= 512: at most two 16-byte vectors per lane
#pragma unroll
            for (int k = 0; k < kRouteLnVec; ++k)
              v[k] = (p.drop_residual != nullptr && k * 32 + lane < p.D / 8) ? __ldg(rrow + k * 32 + lane)
                                                                            : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (int k = 0; k < kRouteLnVec; ++k)
              if (k * 32 + lane < p.D / 8) orow[k * 32 + lane] = v[k];
          }
        }
      }
    }
    __syncthreads();
  }

  if (static_cast<int>(blockIdx.x) == (kEp ? 0 : p.table_cta)) {
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
      p.counts[e] = s_total[e];
      if (p.counts_out) p.counts_out[e] = s_total[e];
    }
    for (int e = threadIdx.x; e <= E; e += blockDim.x) {
      p.offsets[e] = s_off[e];
      if (p.offsets_out) p.offsets_out[e] = s_off[e];
    }
    if (!kEp) {
      // FFN group table: expert e contributes ceil(count / bn) groups, in expert order
      if (threadIdx.x == 0) {
        int acc = 0;
        for (int e = 0; e < E; ++e) {
          s_scratch[e] = acc;
          acc += (s_total[e] + p.bn - 1) / p.bn;
        }
        p.n_groups[0] = acc;
      }
      __syncthreads();
      for (int e = threadIdx.x; e < E; e += blockDim.x) {
        const int cnt = s_total[e];
        const int nt = (cnt + p.bn - 1) / p.bn;
        for (int j = 0; j < nt; ++j) {
          GroupRec r;
          r.expert = e;
          r.row0 = s_off[e] + j * p.bn;
          r.nrows = min(p.bn, cnt - j * p.bn);
          r.src = 0;
          r.orow0 = r.row0;
          r.pad[0] = r.pad[1] = r.pad[2] = 0;
          p.groups[s_scratch[e] + j] = r;
        }
      }
      for (int g = threadIdx.x; g < p.gmax; g += blockDim.x) p.h_ready[g] = 0;
    } else {
      // this rank as an OWNER: every rank's counts are needed only now (they left the peers before their rows did), for
      // the expert kernel's group table over the received rows
      const bool ok = ep_wait_counts(ep, ep_seq, E, p.ep_mode, s_cnt);
      ep_build_groups_segmented(ep, E, s_cnt, ok, p.bn, p.groups, p.n_groups, p.h_ready, p.gmax, s_base + 40);
    }
  }

  // ================================================ completion ================================================
  // The histogram words must not survive into a replay of the same CUDA graph (same tag): the expert-FFN kernel that
  // always follows clears them (FfnLaunch::clear_ptr).
  __syncthreads();
  if (threadIdx.x == 0) rtrace(p, 9);
  if (kEp) {
    // step 3: this CTA's rows (and routing data) have landed -- fence, then one increment on every rank.  The expert
    // kernels wait for the counters themselves, so nobody has to be the last CTA and nobody waits here.
    ep_signal(ep, ep.lay.arrive);
    if (threadIdx.x == 0) {
      // (the sequence number advances once every CTA has read it: they all have by the grid barrier above)
      if (grid_arrive(p.bar, p.nonce) == gridDim.x) {
        ep_ctrl(ep)[0] = ep_seq;
        p.bar[0] = static_cast<unsigned long long>(~p.nonce) << 32;  // a foreign nonce: the next launch starts from scratch
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<kLn ? 2 * kRTmemCols : kRTmemCols>(tmem_base);
  }
  if (threadIdx.x == 0) {
    rtrace(p, 10);
    rtrace_sync(p, 15);
    mark(5);
  }
}

std::atomic<unsigned> g_route_nonce{1};
void* g_route_trace = nullptr;

// 56 random bits per launch (splitmix64 over a counter seeded from the clock and an address): see the grid barrier.
unsigned long long next_route_tag() {
  static std::atomic<unsigned long long> ctr{[] {
    unsigned long long s = static_cast<unsigned long long>(std::chrono::steady_clock::now().time_since_epoch().count());
    s ^= static_cast<unsigned long long>(reinterpret_cast<uintptr_t>(&g_route_trace)) << 17;
    s ^= static_cast<unsigned long long>(std::random_device{}()) << 32;
    return s;
  }()};
  unsigned long long z = ctr.fetch_add(0x9e3779b97f4a7c15ull, std::memory_order_relaxed) + 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  z ^= z >> 31;
  z &= ~0xffull;
  return z != 0 ? z : 0x5bd1e99500ull;  // never the all-zero tag (cleared words)
}

}  // namespace

cudaError_t read_route_status(int* host_status, bool clear) {
  cudaError_t e = cudaMemcpyFromSymbol(host_status, g_route_status, sizeof(int));
  if (e == cudaSuccess && clear && *host_status != 0) {
    const int zero = 0;
    e = cudaMemcpyToSymbol(g_route_status, &zero, sizeof(int));
  }
  return e;
}

void set_route_trace(void* dev_buf) { g_route_trace = dev_buf; }

bool route_supported(int S, int D, int Demb, int E, int top_k, int dtype) {
  // up to two 32-token tiles per SM; beyond that the 128-token gate re-reads the router far less often
  return route_mode() != 0 && dtype == B200MOE_BF16 && top_k == 1 && E <= 32 && S > 0 && S <= 2 * kTok * num_sms() &&
         D % kRK == 0 && Demb % kRK == 0 && D > 0 && D <= kHalfKb * kRK && Demb <= kHalfKb * kRK;
}

cudaError_t launch_route(const void* x, const void* embed, const void* wr_packed, const float* br, const int* x_len,
                         int B, int T, int D, int Demb, int E, int gate_mode, int keep_expert_output, int* idx,
                         float* score, int bn, const RouteWs& ws, int* counts_out, int* offsets_out, int* mapping_out,
                         bf16* xbuf, void* drop_out, const void* drop_residual, cudaStream_t stream, const EpPeers* ep,
                         bool ep_fold_wait, const float* ln_gamma, const float* ln_beta, float ln_eps, const float* ln_c,
                         int ep_mode, int ep_ffn_ctas) {
  const int S = B * T;
  if (embed == nullptr) Demb = 0;
  if (!route_supported(S, D, Demb, E, 1, B200MOE_BF16)) return cudaErrorInvalidValue;
  if (ep != nullptr && (ep->world * ep->E_local != E || S > ep->cap || ep->D != D)) return cudaErrorInvalidValue;
  // [K/64][rows][64] views: one TMA instruction per operand and K part (see make_tmap_bf16_kblocks)
  CUtensorMap tx, te, twx, twe;
  if (!make_tmap_bf16_kblocks(&tx, x, S, D, kTok, D / kRK)) return cudaErrorInvalidValue;
  if (!make_tmap_bf16_kblocks(&twx, wr_packed, 64, static_cast<uint64_t>(D + Demb), 64, D / kRK))
    return cudaErrorInvalidValue;
  if (Demb > 0) {
    if (!make_tmap_bf16_kblocks(&te, embed, S, Demb, kTok, Demb / kRK)) return cudaErrorInvalidValue;
    if (!make_tmap_bf16_kblocks(&twe, wr_packed, 64, static_cast<uint64_t>(D + Demb), 64, Demb / kRK))
      return cudaErrorInvalidValue;
  } else {
    te = tx;
    twe = twx;
  }
  RouteParams p;
  p.br = br;
  p.x_len = x_len;
  p.idx = idx;
  p.score = score;
  p.hist64 = reinterpret_cast<unsigned long long*>(ws.hist32);
  p.S = S;
  p.T = T;
  p.D = D;
  p.Demb = Demb;
  p.E = E;
  p.gate_mode = gate_mode;
  p.x = static_cast<const bf16*>(x);
  p.xbuf = xbuf;
  p.counts = ws.counts;
  p.offsets = ws.offsets;
  p.mapping = ws.mapping;
  p.pos = ws.pos;
  p.row_score = ws.row_score;
  p.groups = ws.groups;
  p.n_groups = ws.n_groups;
  p.h_ready = ws.h_ready;
  p.bn = bn;
  p.gmax = ep ? max_groups(ep->world * ep->cap, ep->world * ep->E_local, bn) : max_groups(S, E, bn);
  p.counts_out = counts_out;
  p.offsets_out = offsets_out;
  p.mapping_out = mapping_out;
  p.drop_out = static_cast<bf16*>(drop_out);  // (expert parallelism: only the folded path passes one)
  p.drop_residual = static_cast<const bf16*>(drop_residual);
  p.keep_expert_output = keep_expert_output;
  p.bar = reinterpret_cast<unsigned long long*>(ws.n_groups + 2);  // completion counter {nonce, CTAs done}
  // 24-bit launch tag, never 0: a plain count left in the same workspace by the split gate kernel has tag 0
  unsigned nonce = g_route_nonce.fetch_add(1, std::memory_order_relaxed) & 0x00ffffffu;
  if (nonce == 0) nonce = g_route_nonce.fetch_add(1, std::memory_order_relaxed) & 0x00ffffffu;
  p.nonce = nonce;
  p.tag = next_route_tag();
  p.ep_fold_wait = ep_fold_wait ? 1 : 0;
  p.ep_mode = ep_mode;
  p.ep_ffn_ctas = ep_ffn_ctas;
  // kLn: `wr_packed` is the pre-scaled router of b200moe_pack_router_ln and ln_c its c1 / c0 tail
  p.ln_gamma = (ln_gamma != nullptr && ln_beta != nullptr && ln_c != nullptr) ? ln_gamma : nullptr;
  p.ln_beta = ln_beta;
  p.ln_c = ln_c;
  p.ln_eps = ln_eps;
  static const int warm = [] {
    const char* v = std::getenv("B200MOE_WARM");
    return (v && *v) ? std::atoi(v) : 1;
  }();
  p.warm_mma = warm;
  p.trace = static_cast<uint4*>(g_route_trace);
  p.tl = next_timeline_slot(1);
  EpPeers epv{};
  if (ep) epv = *ep;
  const size_t smem_plain =
      kRSlots * kRSlot + 8 * (2 * kRSlots + 4) + 32 + 4 * (32 + 2 * 32 * 33 + 7 * 33 + 512 + 256 + 32 + 64);
  // kLn: x-part accumulator copies, statistics, c1 / c0, gamma / beta
  const size_t smem_ln = smem_plain + 4 * (2 * 32 * 33 + 4 * 32 + 2 * 32 + 2 * kHalfKb * kRK);
  const size_t smem = p.ln_gamma != nullptr ? smem_ln : smem_plain;
  static bool attr_set = false;
  if (!attr_set) {
    const int bytes = static_cast<int>(smem_plain), bytes_ln = static_cast<int>(smem_ln);
    cudaError_t e = cudaFuncSetAttribute(route_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(route_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(route_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes_ln);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(route_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes_ln);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int n_tiles = (S + kTok - 1) / kTok;
  int grid = n_tiles < num_sms() ? n_tiles : num_sms();
  // Fewer tiles than SMs: one more CTA, without a tile, takes the barrier like the others and writes the tables while they
  // copy their rows (CTA 0 used to end 0.8-1 us after the rest, and the expert kernel waits for the last CTA): measured
  // on one box -0.24 us per layer at 3 200 tokens (100 tiles), -1.07 us at 50 tokens (2 tiles).
  static const int tcta = [] {
    const char* v = std::getenv("B200MOE_ROUTE_TCTA");
    return (v && *v) ? std::atoi(v) : 1;   // (0: CTA 0 writes the tables, as it does under expert parallelism)
  }();
  p.table_cta = 0;
  if (tcta != 0 && ep == nullptr && n_tiles < num_sms()) {
    p.table_cta = n_tiles;
    grid = n_tiles + 1;
  }
  const bool ln = p.ln_gamma != nullptr;
  auto kernel = ep ? (ln ? route_kernel<true, true> : route_kernel<true, false>)
                   : (ln ? route_kernel<false, true> : route_kernel<false, false>);
  cudaError_t e = launch_kernel(kernel, dim3(grid), dim3(kRThreads), smem, stream, kPdlGate, tx, te, twx, twe, p, epv);
  count_launch();
  return e;
}

}  // namespace b200moe
