// Expert FFN: ONE persistent kernel for  y = act(x . W1[e]^T + b1[e]) . W2[e]^T + b2[e]  over all experts.
//
// Replaces the reference's per-expert loop of cublasSgemm + BiasSilu + cublasSgemm + Bias on 8 streams with two
// host synchronisations (TRTAPI++/plugin/fmoe_expert_plugin/fmoe_expert_plugin.cpp:75-130) and the absent
// fmoe_cuda.forward grouped GEMM (trainer_3m_fix/fmoe/functions.py:135-152).
//
// Design (B200 / sm_100a):
//  * "swap-AB" orientation: the WEIGHT rows are the UMMA M dimension (always a full 128-row tile), the expert's
//    TOKENS are the UMMA N dimension (BN = 32..256, chosen per launch from the average tokens per expert).  The
//    3M-ASR regime is a few to ~100 tokens per expert, where the layer is bound by streaming 64 MiB of bf16
//    weights from HBM: no tensor-core work is wasted on padding rows and every SM streams weight tiles.
//  * tile = (group, 128 weight rows).  A group is one expert x one BN-token tile.  Phase-1 tiles compute
//    h^T[128 hidden, tokens] = W1 tile . x^T (K = D), add b1, apply the activation and store h as bf16 [row, H];
//    phase-2 tiles compute y^T[128 features, tokens] = W2 tile . h^T (K = H), add b2 and either store y in
//    expert order or (top-1) write  out[token] = residual + ff_scale * score * y  straight to token order
//    (the reference's gather, x gate_value, x ff_scale and + residual: fmoe_expert_kernel.cu:191-227,
//    positionwise_feed_forward.py:257-258, fmoe_transformer.py:155-158).
//  * both phases live in one static round-robin tile list; phase-2 tiles of a group are scheduled ~3 waves after
//    its phase-1 tiles and wait on a per-group counter in global memory, so h only ever travels through L2.
//  * warp roles: warp 0 = TMA producer (cp.async.bulk.tensor, 128B swizzle, multi-stage mbarrier ring),
//    warp 1 = single-thread tcgen05.mma issuer (fp32 accumulators in TMEM, double buffered: 2 x 256 columns),
//    warp 2 = TMEM allocator, warps 4-7 = epilogue (tcgen05.ld -> bias/activation -> global).
#include <mutex>

#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

namespace b200moe {

namespace {

constexpr int kBlockM = 128;                           // weight rows per tile == UMMA M
constexpr int kBlockK = 64;                            // bf16 elements per k-block == one 128 B swizzle row
constexpr int kUmmaK = 16;                             // K per tcgen05.mma for 16-bit inputs
constexpr int kMaxStages = 10;
constexpr int kAStageBytes = kBlockM * kBlockK * 2;    // 16 KiB
constexpr int kStageBudget = 200 * 1024;
constexpr int kThreads = 256;
constexpr int kEpiThreads = 128;
constexpr uint32_t kTmemCols = 512;
constexpr int kAccStride = 256;                        // TMEM columns per accumulator buffer
constexpr int kStagingBytes = 32 * kBlockM * 4;        // epilogue transpose buffer: 32 columns x 128 features fp32

struct FfnParams {
  const GroupRec* groups;
  const int* n_groups;
  int* h_ready;
  const float* b1;
  const float* b2;
  bf16* hbuf;
  void* out;
  const void* residual;
  const int* pos;
  const float* row_score;
  float ff_scale;
  int top_k;
  int E, D, H, bn, act, fused;
  int stages;
  int lag;  // groups between a group's phase-1 and phase-2 tiles in the schedule
  // debug timeline (ffn_kernel<.., true> only): per CTA `trace_cap` records of {tile, event, globaltimer lo, hi}
  uint4* trace;
  int trace_cap;
};

enum TraceEvent : int {
  kEvKernelStart = 0, kEvProdTileStart = 1, kEvProdDepOk = 2, kEvProdIssued = 3, kEvMmaAccFree = 4,
  kEvMmaFirstData = 5, kEvMmaIssued = 6, kEvEpiAccReady = 7, kEvEpiAccReleased = 8, kEvEpiStored = 9,
  kEvEpiPublished = 10, kEvKernelEnd = 11
};

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// One slot counter per role (0 producer, 1 mma, 2 epilogue) so the three recording threads never collide.
template <bool kTrace>
struct Tracer {
  uint4* base;
  int cap, n;
  __device__ __forceinline__ Tracer(const FfnParams& p, int role) : base(nullptr), cap(0), n(0) {
    if constexpr (kTrace) {
      if (p.trace != nullptr) {
        cap = p.trace_cap / 3;
        base = p.trace + (static_cast<size_t>(blockIdx.x) * 3 + role) * cap;
      }
    }
  }
  __device__ __forceinline__ void rec(int tile, int ev) {
    if constexpr (kTrace) {
      if (base != nullptr && n < cap) {
        const unsigned long long t = global_ns();
        base[n++] = make_uint4(static_cast<uint32_t>(tile), static_cast<uint32_t>(ev), static_cast<uint32_t>(t),
                               static_cast<uint32_t>(t >> 32));
      }
    }
  }
};

struct Tile {
  int phase;  // 1 or 2
  int g;      // group
  int mb;     // 128-row block of the weight matrix
};

// Static schedule: L "head" slots hold only phase-1 tiles, the middle slots hold phase-1 tiles of group s and
// phase-2 tiles of group s-L, the last L slots only phase-2 tiles.  Every dependency of a tile has a smaller index.
__device__ __forceinline__ Tile decode_tile(int t, int ng, int lag, int m1, int m2) {
  Tile r;
  const int head = lag * m1;
  if (t < head) {
    r.phase = 1;
    r.g = t / m1;
    r.mb = t - r.g * m1;
    return r;
  }
  const int per = m1 + m2;
  const int mid_slots = ng - lag;
  int u = t - head;
  if (u < mid_slots * per) {
    const int s = u / per;
    const int w = u - s * per;
    if (w < m1) {
      r.phase = 1;
      r.g = lag + s;
      r.mb = w;
    } else {
      r.phase = 2;
      r.g = s;
      r.mb = w - m1;
    }
    return r;
  }
  u -= mid_slots * per;
  r.phase = 2;
  r.g = mid_slots + u / m2;
  r.mb = u % m2;
  return r;
}

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == B200MOE_ACT_SILU) {
    // x * sigmoid(x)  (trainer_3m_fix/utils/common.py:24-28; TRTAPI++/plugin/common/common.cuh sigmoid)
    return __fdividef(v, 1.0f + __expf(-v));
  } else if (act == B200MOE_ACT_RELU) {
    return fmaxf(v, 0.0f);
  } else {
    return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
  }
}

// Four consecutive boundary elements as one vector access (8 B for 16-bit types, 16 B for fp32), predicated so that
// the code stays branch-free (a C++ `if` around a store makes the compiler sink the value's whole computation into a
// per-column basic block and the columns' dependency chains stop interleaving).
template <typename T>
struct Vec4Io;
template <>
struct Vec4Io<float> {
  struct raw_t { uint32_t a, b, c, d; };
  static __device__ __forceinline__ raw_t load(const float* p, bool pred) {
    raw_t r;
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b32 %0, 0;\n\tmov.b32 %1, 0;\n\tmov.b32 %2, 0;\n\t"
        "mov.b32 %3, 0;\n\t@p ld.global.v4.b32 {%0, %1, %2, %3}, [%4];\n\t}"
        : "=r"(r.a), "=r"(r.b), "=r"(r.c), "=r"(r.d)
        : "l"(p), "r"(static_cast<int>(pred)));
    return r;
  }
  static __device__ __forceinline__ void to_f(const raw_t& r, float (&o)[4]) {
    o[0] = __uint_as_float(r.a); o[1] = __uint_as_float(r.b); o[2] = __uint_as_float(r.c); o[3] = __uint_as_float(r.d);
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4], bool pred) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t@p st.global.v4.b32 [%0], {%1, %2, %3, %4};\n\t}"
        :
        : "l"(p), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
          "r"(__float_as_uint(v[3])), "r"(static_cast<int>(pred))
        : "memory");
  }
};
struct Raw2 { uint32_t a, b; };
__device__ __forceinline__ Raw2 ld_pred_v2(const void* p, bool pred) {
  Raw2 r;
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %3, 0;\n\tmov.b32 %0, 0;\n\tmov.b32 %1, 0;\n\t"
      "@p ld.global.v2.b32 {%0, %1}, [%2];\n\t}"
      : "=r"(r.a), "=r"(r.b)
      : "l"(p), "r"(static_cast<int>(pred)));
  return r;
}
__device__ __forceinline__ void st_pred_v2(void* p, uint32_t a, uint32_t b, bool pred) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %3, 0;\n\t@p st.global.v2.b32 [%0], {%1, %2};\n\t}"
               :
               : "l"(p), "r"(a), "r"(b), "r"(static_cast<int>(pred))
               : "memory");
}
__device__ __forceinline__ void st_pred_v4(void* p, const uint4& v, bool pred) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t@p st.global.v4.b32 [%0], {%1, %2, %3, %4};\n\t}"
               :
               : "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(static_cast<int>(pred))
               : "memory");
}
template <>
struct Vec4Io<bf16> {
  using raw_t = Raw2;
  static __device__ __forceinline__ raw_t load(const bf16* p, bool pred) { return ld_pred_v2(p, pred); }
  static __device__ __forceinline__ void to_f(const raw_t& r, float (&o)[4]) {
    o[0] = __uint_as_float(r.a << 16); o[1] = __uint_as_float(r.a & 0xffff0000u);
    o[2] = __uint_as_float(r.b << 16); o[3] = __uint_as_float(r.b & 0xffff0000u);
  }
  static __device__ __forceinline__ void store(bf16* p, const float (&v)[4], bool pred) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]);
    __nv_bfloat162 hi = __floats2bfloat162_rn(v[2], v[3]);
    st_pred_v2(p, *reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi), pred);
  }
};
template <>
struct Vec4Io<__half> {
  using raw_t = Raw2;
  static __device__ __forceinline__ raw_t load(const __half* p, bool pred) { return ld_pred_v2(p, pred); }
  static __device__ __forceinline__ void to_f(const raw_t& r, float (&o)[4]) {
    const float2 x = __half22float2(*reinterpret_cast<const __half2*>(&r.a));
    const float2 y = __half22float2(*reinterpret_cast<const __half2*>(&r.b));
    o[0] = x.x; o[1] = x.y; o[2] = y.x; o[3] = y.y;
  }
  static __device__ __forceinline__ void store(__half* p, const float (&v)[4], bool pred) {
    __half2 lo = __floats2half2_rn(v[0], v[1]);
    __half2 hi = __floats2half2_rn(v[2], v[3]);
    st_pred_v2(p, *reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi), pred);
  }
};

template <typename OutT, bool kTrace>
__global__ void __launch_bounds__(kThreads, 1)
ffn_kernel(const __grid_constant__ CUtensorMap tm_w1, const __grid_constant__ CUtensorMap tm_w2,
           const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_h, const FfnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B needs 1024 B alignment
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int stages = p.stages;
  const uint32_t b_stage_bytes = static_cast<uint32_t>(p.bn) * kBlockK * 2;
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_base + stages * kAStageBytes;
  const uint32_t bar_base = smem_b + stages * b_stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kMaxStages + 4);
  // epilogue staging: one chunk of 32 token columns x 128 features (fp32, or bf16 for h) + per-column routing tables
  const uint32_t stage_off = ((tmem_slot + 16u + 15u) & ~15u) - ptx::smem_u32(smem_raw);
  float* stg_f = reinterpret_cast<float*>(smem_raw + stage_off);
  int* s_tok = reinterpret_cast<int*>(smem_raw + stage_off + kStagingBytes);
  float* s_sc = reinterpret_cast<float*>(smem_raw + stage_off + kStagingBytes + 256 * 4);
  // generic pointer to the tmem slot for reading it back
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tm_w1);
    ptx::prefetch_tensormap(&tm_w2);
    ptx::prefetch_tensormap(&tm_x);
    ptx::prefetch_tensormap(&tm_h);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < stages; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(tfull_bar(s), 1);
      ptx::mbar_init(tempty_bar(s), kEpiThreads / 32);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc<kTmemCols>(tmem_slot);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int ng = *p.n_groups;
  const int m1 = p.H / kBlockM;
  const int m2 = p.D / kBlockM;
  const int lag = min(p.lag, ng);
  const int n_tiles = ng * (m1 + m2);
  const int kb1 = p.D / kBlockK;  // k-blocks of the first GEMM
  const int kb2 = p.H / kBlockK;

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = kAStageBytes + b_stage_bytes;
      Tracer<kTrace> tr(p, 0);
      tr.rec(-1, kEvKernelStart);
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const Tile tl = decode_tile(t, ng, lag, m1, m2);
        const GroupRec gr = p.groups[tl.g];
        tr.rec(t, kEvProdTileStart);
        const CUtensorMap* tm_a = tl.phase == 1 ? &tm_w1 : &tm_w2;
        const CUtensorMap* tm_b = tl.phase == 1 ? &tm_x : &tm_h;
        const int a_row = gr.expert * (tl.phase == 1 ? p.H : p.D) + tl.mb * kBlockM;
        const int nkb = tl.phase == 1 ? kb1 : kb2;
        if (tl.phase == 2) {
          // wait until every phase-1 tile of this group has published its slice of h
          while (ptx::ld_acquire_gpu(p.h_ready + tl.g) < m1) __nanosleep(40);
          ptx::fence_proxy_async_all();  // generic-proxy writes of h -> async-proxy (TMA) reads
        }
        tr.rec(t, kEvProdDepOk);
        for (int kb = 0; kb < nkb; ++kb) {
          ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
          ptx::mbar_arrive_expect_tx(full_bar(stage), tx_bytes);
          ptx::tma_load_2d(smem_a + stage * kAStageBytes, tm_a, full_bar(stage), kb * kBlockK, a_row,
                           ptx::kEvictNormal);
          ptx::tma_load_2d(smem_b + stage * b_stage_bytes, tm_b, full_bar(stage), kb * kBlockK, gr.row0,
                           ptx::kEvictLast);
          if (++stage == stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        tr.rec(t, kEvProdIssued);
      }
      tr.rec(-1, kEvKernelEnd);
    }
  } else if (warp == 1) {
    // ============================ MMA issuer (one thread) ============================
    if (lane == 0) {
      const uint32_t idesc = ptx::make_idesc(1u /*bf16*/, kBlockM, static_cast<uint32_t>(p.bn));
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      Tracer<kTrace> tr(p, 1);
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const Tile tl = decode_tile(t, ng, lag, m1, m2);
        const int nkb = tl.phase == 1 ? kb1 : kb2;
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        ptx::mbar_wait(tempty_bar(as), aphase ^ 1u);  // epilogue has drained this accumulator buffer
        ptx::tc_fence_after();
        tr.rec(t, kEvMmaAccFree);
        const uint32_t tmem_d = tmem_base + as * kAccStride;
        for (int kb = 0; kb < nkb; ++kb) {
          ptx::mbar_wait(full_bar(stage), phase);
          ptx::tc_fence_after();
          if (kb == 0) tr.rec(t, kEvMmaFirstData);
          const uint64_t a_desc = ptx::make_kmajor_sw128_desc(smem_a + stage * kAStageBytes);
          const uint64_t b_desc = ptx::make_kmajor_sw128_desc(smem_b + stage * b_stage_bytes);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            // advance both descriptors by k * 16 elements * 2 B = 32 B -> +2 in 16-byte units
            ptx::umma_f16_ss(tmem_d, a_desc + 2u * k, b_desc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          ptx::umma_commit(empty_bar(stage));  // smem slot free once these MMAs have read it
          if (kb == nkb - 1) ptx::umma_commit(tfull_bar(as));
          if (++stage == stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        tr.rec(t, kEvMmaIssued);
      }
    }
  } else if (warp >= 4) {
    // ============================ epilogue (4 warps, one TMEM lane quarter each) ============================
    // tcgen05.ld gives thread (q, lane) ONE feature and 32 token columns; global memory wants whole token rows.
    // Each chunk of 32 columns is therefore transposed through shared memory: the feature-major values are written
    // column by column (conflict-free), then warp q turns columns 8q..8q+7 into row segments of 128 features with
    // 8/16-byte vector accesses (256 / 512 contiguous bytes per warp instruction).  All global loads of a chunk
    // (residual rows) are issued before anything depends on them; routing records come from a table filled before
    // the accumulator wait.  Global memory latency under full HBM load is 1-2 us, so no dependent load chains.
    const int q = warp & 3;  // tcgen05.ld: warp w may touch lanes 32*(w%4) .. +31
    const int et = threadIdx.x - (kThreads - kEpiThreads);
    const int feat_l = q * 32 + lane;  // feature within the tile owned in the column phase
    int it = 0;
    Tracer<kTrace> tr(p, 2);
    const bool tracer_thread = et == 0;
    bf16* stg_h = reinterpret_cast<bf16*>(stg_f);
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const Tile tl = decode_tile(t, ng, lag, m1, m2);
      const GroupRec gr = p.groups[tl.g];
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * kAccStride;
      const int feat0 = tl.mb * kBlockM;  // first feature (weight row) of the tile
      const int nrows = gr.nrows;
      if (tl.phase == 1) {
        const float bias = p.b1 ? p.b1[static_cast<size_t>(gr.expert) * p.H + feat0 + feat_l] : 0.0f;
        ptx::mbar_wait(tfull_bar(as), aphase);
        ptx::tc_fence_after();
        if (tracer_thread) tr.rec(t, kEvEpiAccReady);
        for (int c0 = 0; c0 < nrows; c0 += 32) {
          uint32_t r[32];
          ptx::tmem_ld_32x32b_x32(taddr + c0, r);
          ptx::tmem_ld_wait();
          if (c0 + 32 >= nrows) {  // last chunk is in registers: hand the accumulator back to the MMA warp
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(tempty_bar(as));
            if (tracer_thread) tr.rec(t, kEvEpiAccReleased);
          }
          if (p.act == B200MOE_ACT_SILU) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              stg_h[j * kBlockM + feat_l] =
                  __float2bfloat16_rn(apply_act(__uint_as_float(r[j]) + bias, B200MOE_ACT_SILU));
          } else if (p.act == B200MOE_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              stg_h[j * kBlockM + feat_l] =
                  __float2bfloat16_rn(apply_act(__uint_as_float(r[j]) + bias, B200MOE_ACT_RELU));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              stg_h[j * kBlockM + feat_l] =
                  __float2bfloat16_rn(apply_act(__uint_as_float(r[j]) + bias, B200MOE_ACT_GELU));
          }
          ptx::named_bar_sync(1, kEpiThreads);
          // row phase: warp q owns columns 8q .. 8q+7; a half-warp writes one 256-byte row segment of h
          {
            const int half = lane >> 4;
            const int l16 = lane & 15;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int j = q * 8 + 2 * i + half;
              const uint4 val = *reinterpret_cast<const uint4*>(stg_h + j * kBlockM + l16 * 8);
              bf16* dst = p.hbuf + static_cast<size_t>(gr.row0 + c0 + j) * p.H + feat0 + l16 * 8;
              st_pred_v4(dst, val, c0 + j < nrows);
            }
          }
          ptx::named_bar_sync(1, kEpiThreads);  // staging buffer free again; all h stores of the chunk issued
        }
        // publish: every epilogue thread's stores precede the barrier above; one thread releases them at gpu scope
        if (tracer_thread) tr.rec(t, kEvEpiStored);
        if (et == 0) {
          ptx::fence_proxy_async_all();
          ptx::red_release_gpu_add(p.h_ready + tl.g, 1);
          tr.rec(t, kEvEpiPublished);
        }
      } else {
        const float bias = p.b2 ? p.b2[static_cast<size_t>(gr.expert) * p.D + feat0 + feat_l] : 0.0f;
        OutT* out = static_cast<OutT*>(p.out);
        const OutT* res = static_cast<const OutT*>(p.residual);
        const bool with_res = p.fused && res != nullptr;
        // routing table of the tile's columns: output row and scale (filled while the MMAs are still running)
        for (int i = et; i < nrows; i += kEpiThreads) {
          const int row = gr.row0 + i;
          int tok = row;
          float sc = 1.0f;
          if (p.fused) {
            tok = p.pos[row] / p.top_k;
            sc = p.ff_scale * (p.row_score ? p.row_score[row] : 1.0f);
          }
          s_tok[i] = tok;
          s_sc[i] = sc;
        }
        ptx::mbar_wait(tfull_bar(as), aphase);
        ptx::tc_fence_after();
        if (tracer_thread) tr.rec(t, kEvEpiAccReady);
        using Io = Vec4Io<OutT>;
        for (int c0 = 0; c0 < nrows; c0 += 32) {
          uint32_t r[32];
          ptx::tmem_ld_32x32b_x32(taddr + c0, r);
          ptx::tmem_ld_wait();
          if (c0 + 32 >= nrows) {
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(tempty_bar(as));
            if (tracer_thread) tr.rec(t, kEvEpiAccReleased);
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) stg_f[j * kBlockM + feat_l] = __uint_as_float(r[j]) + bias;
          ptx::named_bar_sync(1, kEpiThreads);  // staging (and, for the first chunk, the routing table) complete
          {
            // row phase: warp q owns columns 8q .. 8q+7, lane l the 4 features 4l .. 4l+3 of each
            typename Io::raw_t rres[8];
            size_t off[8];
            bool valid[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int col = c0 + q * 8 + i;
              valid[i] = col < nrows;
              const int tok = valid[i] ? s_tok[col] : 0;
              off[i] = static_cast<size_t>(tok) * p.D + feat0 + lane * 4;
              rres[i] = Io::load(res + off[i], valid[i] && with_res);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int j = q * 8 + i;
              const float4 a = *reinterpret_cast<const float4*>(stg_f + j * kBlockM + lane * 4);
              const float sc = valid[i] ? s_sc[c0 + j] : 0.0f;
              float rv[4], o[4];
              Io::to_f(rres[i], rv);
              o[0] = fmaf(sc, a.x, rv[0]);
              o[1] = fmaf(sc, a.y, rv[1]);
              o[2] = fmaf(sc, a.z, rv[2]);
              o[3] = fmaf(sc, a.w, rv[3]);
              Io::store(out + off[i], o, valid[i]);
            }
          }
          ptx::named_bar_sync(1, kEpiThreads);  // staging buffer (and routing table) free again
        }
        if (tracer_thread) tr.rec(t, kEvEpiStored);
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Host side: tensor maps + launch
// ------------------------------------------------------------------------------------------------------------
void* g_trace_buf = nullptr;
int g_trace_cap = 0;

int stages_for_bn(int bn) {
  int s = kStageBudget / (kAStageBytes + bn * kBlockK * 2);
  return s > kMaxStages ? kMaxStages : s;
}

size_t smem_bytes_for(int bn, int stages) {
  return 1024 + static_cast<size_t>(stages) * (kAStageBytes + bn * kBlockK * 2) + 8 * (2 * kMaxStages + 4) + 32 +
         kStagingBytes + 2 * 256 * 4;
}

template <typename OutT>
cudaError_t launch_typed(const FfnLaunch& a, const CUtensorMap& tw1, const CUtensorMap& tw2, const CUtensorMap& tx,
                         const CUtensorMap& th, const FfnParams& p, cudaStream_t stream) {
  const size_t smem = smem_bytes_for(a.bn, p.stages);
  static bool attr_set = false;  // per OutT instantiation
  if (!attr_set) {
    cudaError_t e =
        cudaFuncSetAttribute(ffn_kernel<OutT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(ffn_kernel<OutT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int m1 = a.H / kBlockM, m2 = a.D / kBlockM;
  long long tiles_ub = static_cast<long long>(a.gmax) * (m1 + m2);
  int grid = num_sms();
  if (tiles_ub < grid) grid = static_cast<int>(tiles_ub);
  if (grid < 1) grid = 1;
  if (p.trace != nullptr)
    ffn_kernel<OutT, true><<<grid, kThreads, smem, stream>>>(tw1, tw2, tx, th, p);
  else
    ffn_kernel<OutT, false><<<grid, kThreads, smem, stream>>>(tw1, tw2, tx, th, p);
  count_launch();
  return cudaGetLastError();
}

}  // namespace

void set_ffn_trace(void* dev_buf, int records_per_cta) {
  g_trace_buf = dev_buf;
  g_trace_cap = records_per_cta;
}

cudaError_t launch_ffn(const FfnLaunch& a, cudaStream_t stream) {
  if (a.n_rows <= 0) return cudaSuccess;
  if (a.D % kBlockM != 0 || a.H % kBlockM != 0) return cudaErrorInvalidValue;
  if (a.bn % 16 != 0 || a.bn < 16 || a.bn > 256) return cudaErrorInvalidValue;
  if (a.fused && a.top_k != 1) return cudaErrorInvalidValue;
  CUtensorMap tw1, tw2, tx, th;
  if (!make_tmap_bf16(&tw1, a.W1, static_cast<uint64_t>(a.E) * a.H, a.D, kBlockM, kBlockK) ||
      !make_tmap_bf16(&tw2, a.W2, static_cast<uint64_t>(a.E) * a.D, a.H, kBlockM, kBlockK) ||
      !make_tmap_bf16(&tx, a.xbuf, a.n_rows, a.D, a.bn, kBlockK) ||
      !make_tmap_bf16(&th, a.hbuf, a.n_rows, a.H, a.bn, kBlockK)) {
    return cudaErrorInvalidValue;
  }
  FfnParams p;
  p.groups = a.groups;
  p.n_groups = a.n_groups;
  p.h_ready = a.h_ready;
  p.b1 = a.b1;
  p.b2 = a.b2;
  p.hbuf = a.hbuf;
  p.out = a.out;
  p.residual = a.residual;
  p.pos = a.pos;
  p.row_score = a.row_score;
  p.ff_scale = a.ff_scale;
  p.top_k = a.top_k < 1 ? 1 : a.top_k;
  p.E = a.E;
  p.D = a.D;
  p.H = a.H;
  p.bn = a.bn;
  p.act = a.act;
  p.fused = a.fused;
  p.stages = stages_for_bn(a.bn);
  p.trace = static_cast<uint4*>(g_trace_buf);
  p.trace_cap = g_trace_cap;
  // phase-2 tiles trail their group's phase-1 tiles by ~3 waves of the grid
  const int m1 = a.H / kBlockM;
  p.lag = (3 * num_sms() + m1 - 1) / m1;
  switch (a.out_dtype) {
    case B200MOE_F32:
      return launch_typed<float>(a, tw1, tw2, tx, th, p, stream);
    case B200MOE_F16:
      return launch_typed<__half>(a, tw1, tw2, tx, th, p, stream);
    case B200MOE_BF16:
      return launch_typed<bf16>(a, tw1, tw2, tx, th, p, stream);
    default:
      return cudaErrorInvalidValue;
  }
}

}  // namespace b200moe
