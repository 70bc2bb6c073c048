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
//  * warp roles (384 threads): warp 0 = TMA producer (cp.async.bulk.tensor, 128B swizzle, mbarrier ring of stages
//    holding two k-blocks each), warp 1 = single-thread tcgen05.mma issuer (fp32 accumulators in TMEM, double
//    buffered: 2 x 256 columns), warp 2 = TMEM allocator and second TMA producer (token rows; warp 0 then requests
//    the weights only), warp 3 = publisher of the h flags, warps 4-11 = epilogue
//    (two sets of four warps; tcgen05.ld -> bias/activation -> smem transpose -> row-wise vector stores).
//  * 256-token tiles run on CTA pairs (cta_group::2, kCtas = 2): see ffn_kernel.
//  * expert parallelism: the second GEMM's epilogue stores its rows straight into the source GPU's return buffer and
//    the last CTA raises the "rows are back" flags (trainer_3m_fix/fmoe/functions.py:185-191 without the collective).
#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "ep_device.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

namespace b200moe {

namespace {

constexpr int kBlockM = 128;                           // weight rows per tile == UMMA M
constexpr int kBlockK = 64;                            // bf16 elements per k-block == one 128 B swizzle row
constexpr int kUmmaK = 16;                             // K per tcgen05.mma for 16-bit inputs
constexpr int kMaxStages = 10;
constexpr int kAStageBytes = kBlockM * kBlockK * 2;    // 16 KiB
constexpr int kStageBudget = 192 * 1024;
constexpr int kThreads = 384;                          // 4 control warps + 8 epilogue warps
constexpr int kEpiThreads = 256;                       // two sets of four warps (one TMEM lane quarter each)
constexpr int kSetThreads = 128;
constexpr uint32_t kTmemCols = 512;
constexpr int kAccStride = 256;                        // TMEM columns per accumulator buffer
constexpr int kStagingBytes = 32 * kBlockM * 4;        // per epilogue set: 32 columns x 128 features fp32 (or 2 x bf16)
constexpr int kActNone = 3;                            // internal: no activation (grouped linear, b200moe_expert_linear)

struct FfnParams {
  const GroupRec* groups;
  const int* n_groups;
  int* h_ready;
  const float* b1;
  const float* b2;
  void* hbuf;  // bf16 [rows, H], or fp32 for TF32 compute
  void* out;
  const void* residual;
  const int* pos;
  const float* row_score;
  float ff_scale;
  int top_k;
  int E, D, H, bn, act, fused;
  int stages;
  int gmax;  // capacity of the group table
  int lag;  // groups between a group's phase-1 and phase-2 tiles in the schedule
  int kps;  // k-blocks (of 64) per pipeline stage: one TMA instruction per operand brings all of them
  int p1_only;  // first GEMM only (a single grouped linear): the tile list has no second-GEMM tiles
  int pdl_trigger;  // release the dependent kernel at the start (1) or at exit (0)
  int warm_mma;     // issue one throw-away MMA before the first tile (B200MOE_WARM, default 1)
  int two_prod;     // 1: weights and token rows are requested by two threads in two warps (see the producers)
  // expert parallelism: results go to the source rank's return buffer over peer-mapped memory (NVLink)
  int ep;                           // 0 = off
  int ep_world, ep_rank;
  int ep_fold;                      // 1: the epilogue writes the finished layer output (residual + ff_scale * score * y)
                                    //    into the source rank's `out`; 0: the bare y row into the source's ret_y
  const int2* ep_meta;              // per received row: {source rank << 27 | row index at the source, gate score bits}
  uint8_t* ep_out[kMaxEpWorld];     // where rank r's rows go: its `out` buffer (fold) or its ret_y
  EpPeers ep_peers;                 // (arrival / done counters live in the symmetric buffers)
  int* clear_ptr;   // zeroed at kernel start, spread over the CTAs (tagged histogram words of the route kernel)
  int clear_ints;
  int dbg;          // timing experiments only (B200MOE_DBG): 1 = no phase-2 stores, 2 = no phase-1 stores, 4 = no residual,
                    // 8 / 16 = no B-operand loads in GEMM 1 / 2, 32 = no wait for h
  uint64_t w_policy;  // L2 eviction policy of the weight tiles: evict-first when every tile is read once
  // debug timeline (ffn_kernel<.., true> only): per CTA `trace_cap` records of {tile, event, globaltimer lo, hi}
  uint4* trace;
  int trace_cap;
  unsigned long long* tl;  // cross-kernel timeline slot of this launch (common.cuh: set_timeline) or null
};

enum TraceEvent : int {
  kEvKernelStart = 0, kEvProdTileStart = 1, kEvProdDepOk = 2, kEvProdIssued = 3, kEvMmaAccFree = 4,
  kEvMmaFirstData = 5, kEvMmaIssued = 6, kEvEpiAccReady = 7, kEvEpiAccReleased = 8, kEvEpiStored = 9,
  kEvEpiPublished = 10, kEvKernelEnd = 11, kEvEpiChunkLd = 12, kEvEpiChunkStaged = 13, kEvEpiChunkDone = 14,
  kEvClock = 15,        // time fields = %globaltimer (ns) ...
  kEvClockCycles = 16,  // ... and clock64 read right after it
  kEvProdSlotFree = 17, // first two tiles of a CTA only: the producer's wait for ring slot of stage `kb` is over (tile field = kb)
  kEvMmaStageData = 18, // ... the MMA thread's wait for the data of stage `kb` is over
  kEvMmaStageIssued = 19,
  kEvMmaInstr = 20,     // first stage of a CTA's first tile only: after every single tcgen05.mma (tile field = j * 4 + k)
  kEvWarmIssued = 21,   // trace only: a batch of 8 throw-away MMAs issued at kernel start (tile field = batch) ...
  kEvWarmDone = 22      // ... and completed (commit -> mbarrier -> wait)
};

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// One slot counter per role (0 producer, 1 mma, 2 epilogue, 3 publisher) so the recording threads never collide.
template <bool kTrace>
struct Tracer {
  uint4* base;
  int cap, n;
  __device__ __forceinline__ Tracer(const FfnParams& p, int role, int tail_slots = 0) : base(nullptr), cap(0), n(0) {
    if constexpr (kTrace) {
      if (p.trace != nullptr) {
        cap = p.trace_cap / 4;
        base = p.trace + (static_cast<size_t>(blockIdx.x) * 4 + role) * cap;
        if (tail_slots > 0) n = cap - tail_slots;  // a second recorder of the same role writes the last slots
      }
    }
  }
  // Events carry the SM's cycle counter (a register read); reading %globaltimer costs a few hundred ns and would
  // distort what is being measured.  sync() records one (globaltimer, clock64) pair so that the host can put every
  // CTA's cycle counts on the common ns time base.
  __device__ __forceinline__ void rec(int tile, int ev) {
    if constexpr (kTrace) {
      if (base != nullptr && n < cap - 2) {  // the last two slots stay free for the closing sync() pair
        const unsigned long long t = static_cast<unsigned long long>(clock64());
        base[n++] = make_uint4(static_cast<uint32_t>(tile), static_cast<uint32_t>(ev), static_cast<uint32_t>(t),
                               static_cast<uint32_t>(t >> 32));
      }
    }
  }
  __device__ __forceinline__ void sync() {
    if constexpr (kTrace) {
      if (base != nullptr && n + 1 < cap) {
        const unsigned long long g = global_ns();
        const unsigned long long c = static_cast<unsigned long long>(clock64());
        base[n++] = make_uint4(0u, static_cast<uint32_t>(kEvClock), static_cast<uint32_t>(g),
                               static_cast<uint32_t>(g >> 32));
        base[n++] = make_uint4(0u, static_cast<uint32_t>(kEvClockCycles), static_cast<uint32_t>(c),
                               static_cast<uint32_t>(c >> 32));
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
__device__ __forceinline__ uint4 ld_pred_v4(const void* p, bool pred) {
  uint4 r;
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b32 %0, 0;\n\tmov.b32 %1, 0;\n\tmov.b32 %2, 0;\n\t"
      "mov.b32 %3, 0;\n\t@p ld.global.v4.b32 {%0, %1, %2, %3}, [%4];\n\t}"
      : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
      : "l"(p), "r"(static_cast<int>(pred)));
  return r;
}
// two 16-bit elements: a + b with one rounding (the exact sum of two bf16 / fp16 values, rounded to nearest)
template <typename T>
__device__ __forceinline__ uint32_t add_packed(uint32_t a, uint32_t b);
template <>
__device__ __forceinline__ uint32_t add_packed<bf16>(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
template <>
__device__ __forceinline__ uint32_t add_packed<__half>(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("add.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
template <>
__device__ __forceinline__ uint32_t add_packed<float>(uint32_t a, uint32_t) { return a; }  // never used (fp32 path)
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

// kCtas == 2: CTA pairs (cta_group::2).  The two CTAs of a cluster sit on neighbouring SMs and issue ONE tcgen05.mma
// over 256 weight rows x BN tokens: each CTA stages its own 128 weight rows and HALF of the token tile, the tensor
// cores read the other half out of the peer's shared memory.  Per CTA and k-block that is 16 KiB + BN/2 x 128 B through
// the SM's memory port instead of 16 KiB + BN x 128 B -- at BN = 256 a third less L2 -> SM traffic and shared-memory
// bandwidth, which is what bounds the single-CTA kernel in the compute-bound regime (128 x 256 tiles: 85 FLOP per
// byte moved into the SM, while the MMA rate asks for ~190).  Everything else (tile schedule over pairs, the h
// dependency flags, the epilogue of each CTA over its own 128 accumulator rows) is the single-CTA design.
// kTf32: operands are fp32 in memory (x rows, the reference's fp32 FMoELinear weights, h) and the tensor cores read
// them as TF32 (tcgen05.mma.kind::tf32: 10-bit mantissa, fp32 accumulation) -- the <= 1e-3 flavour of the path.  A
// 128-byte swizzle row then holds 32 elements instead of 64, i.e. twice the k-blocks and twice the bytes per tile.
// kMode: 0 = product, 1 = per-event tracing (tools/ffn_trace.py), 2 = the six cross-kernel timeline marks only.
template <typename OutT, int kMode, int kCtas, bool kTf32>
__global__ void __launch_bounds__(kThreads, 1)
ffn_kernel(const __grid_constant__ CUtensorMap tm_w1, const __grid_constant__ CUtensorMap tm_w2,
           const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_h, const FfnParams p) {
  constexpr bool kTrace = kMode == 1;
  extern __shared__ __align__(1024) uint8_t smem_raw[];   // SWIZZLE_128B tiles need 1024 B alignment
  const uint32_t smem_base = ptx::smem_u32(smem_raw);
  if ((smem_base & 1023u) != 0) __trap();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const uint32_t cta_rank = kCtas == 2 ? ptx::cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  const int stages = p.stages;
  const int b_rows = p.bn / kCtas;  // token rows this CTA stages per k-block
  // A pipeline stage holds kps k-blocks of both operands.  Issuing a TMA costs the producer thread ~130 ns whatever
  // the box size, and at 128 x 256 x 64 per k-block the tensor cores need a new k-block every ~300 ns: with one k-block
  // per instruction the producer, not the memory system, paces the MMAs (measured: 70-76 % of the MMA rate).
  const int kps = p.kps;
  const uint32_t a_blk_bytes = kBlockM * kBlockK * 2;                            // 16 KiB per k-block
  const uint32_t b_blk_bytes = static_cast<uint32_t>(b_rows) * kBlockK * 2;
  const uint32_t a_stage_bytes = a_blk_bytes * kps;
  const uint32_t b_stage_bytes = b_blk_bytes * kps;
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_base + stages * a_stage_bytes;
  const uint32_t bar_base = smem_b + stages * b_stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + 2 + s); };
  auto pfull_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + 4 + s); };   // epilogue -> publisher
  auto pempty_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + 6 + s); };  // publisher -> epilogue
  const uint32_t tmem_slot = bar_base + 8u * (2 * kMaxStages + 8);
  // epilogue staging: one chunk of 32 token columns x 128 features (fp32, or bf16 for h) + per-column routing tables
  const uint32_t stage_off = ((tmem_slot + 16u + 15u) & ~15u) - ptx::smem_u32(smem_raw);
  int* s_tok = reinterpret_cast<int*>(smem_raw + stage_off + 2 * kStagingBytes);
  float* s_sc = reinterpret_cast<float*>(smem_raw + stage_off + 2 * kStagingBytes + 256 * 4);
  // (expert parallelism packs more into s_tok: residual? << 31 | destination rank << 27 | row index, like EpLayout::meta)
  // generic pointer to the tmem slot for reading it back
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));

  if (p.pdl_trigger) ptx::pdl_launch_dependents();  // the next kernel's prologue may overlap this whole kernel
  // cross-kernel timeline (debug): a variant of its own -- as run-time branches in the product variant the six marks cost
  // 0.2-0.3 us per layer (A/B on one box), and the per-event tracing variant is ~3 us slower than the product kernel
  auto mark = [&](int i) {
    if constexpr (kMode != 0) {
      if (p.tl != nullptr) p.tl[blockIdx.x * kTimelineMarks + i] = global_ns();
    }
  };
  if (threadIdx.x == 0) mark(0);
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tm_w1);
    ptx::prefetch_tensormap(&tm_w2);
    ptx::prefetch_tensormap(&tm_x);
    ptx::prefetch_tensormap(&tm_h);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < stages; ++s) {
      // pair: the leader's barrier takes one arrival from each CTA's producer (two producers: one from each of them)
      ptx::mbar_init(full_bar(s), kCtas * (p.two_prod ? 2 : 1));
      ptx::mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(tfull_bar(s), 1);
      ptx::mbar_init(tempty_bar(s), kCtas * kEpiThreads / 32);  // pair: both CTAs' epilogue warps release the leader's
      ptx::mbar_init(pfull_bar(s), kEpiThreads / 32);
      ptx::mbar_init(pempty_bar(s), 1);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    if constexpr (kCtas == 2) ptx::tmem_alloc_pair<kTmemCols>(tmem_slot);
    else ptx::tmem_alloc<kTmemCols>(tmem_slot);
  }
  if (warp == 3 && p.clear_ptr != nullptr) {
    ptx::pdl_wait();  // (the route kernel that wrote these words has completed: ordinary stream order or PDL)
    for (int i = blockIdx.x * 32 + lane; i < p.clear_ints; i += gridDim.x * 32) p.clear_ptr[i] = 0;
  }
  ptx::tc_fence_before();
  if constexpr (kCtas == 2) ptx::cluster_sync();  // the peer's barriers must exist before anything signals them
  else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 1 && lane == 0 && (kCtas == 1 || leader) && p.warm_mma) {
    // The first tcgen05.mma an SM executes in a kernel takes ~2 us to issue (measured: 2.2 us for the first stage's
    // eight instructions against 0.3 us for every later stage).  One throw-away instruction over whatever the ring
    // holds, into the first accumulator (the first real MMA of a tile overwrites it), pays that here -- on-chip state
    // only, so under programmatic dependent launch it overlaps the previous kernel's tail.
    const uint32_t idesc = ptx::make_idesc(kTf32 ? 2u /*tf32*/ : 1u /*bf16*/, kBlockM * kCtas,
                                           static_cast<uint32_t>(p.bn));
    const uint64_t a_desc = ptx::make_kmajor_sw128_desc(smem_a);
    const uint64_t b_desc = ptx::make_kmajor_sw128_desc(smem_b);
    for (int i = 0; i < p.warm_mma; ++i) {
      if constexpr (kTf32) ptx::umma_tf32_ss(tmem_base, a_desc, b_desc, idesc, 0u);
      else if constexpr (kCtas == 2) ptx::umma_f16_ss_pair(tmem_base, a_desc, b_desc, idesc, 0u);
      else ptx::umma_f16_ss(tmem_base, a_desc, b_desc, idesc, 0u);
    }
    if constexpr (kTrace && kCtas == 1 && !kTf32) {
      // experiment: how long do 8 MMAs take on an otherwise idle SM, first batch vs later batches?
      if (p.dbg & 64) {
        Tracer<kTrace> tr(p, 1, 16);
        const uint32_t wbar = tmem_slot + 8;
        ptx::mbar_init(wbar, 1);
        ptx::fence_mbar_init();
        for (int b = 0; b < 4; ++b) {
          tr.rec(b, kEvWarmIssued);
          for (int i = 0; i < 8; ++i) ptx::umma_f16_ss(tmem_base, a_desc, b_desc, idesc, 0u);
          ptx::umma_commit(wbar);
          tr.rec(b, kEvWarmIssued);
          ptx::mbar_wait(wbar, b & 1);
          tr.rec(b, kEvWarmDone);
        }
      }
    }
  }

  // Everything above touched only kernel parameters and on-chip state, so under programmatic dependent launch it
  // overlaps the tail of the dispatch kernel.  The routing tables, xbuf and every output come after this wait.
  ptx::pdl_wait();
  if (threadIdx.x == 0) mark(1);
  // a tile is kCtas x 128 weight rows of one group; this CTA's share is rows (mb * kCtas + rank) * 128 ...
  const int m1 = p.H / (kBlockM * kCtas);
  const int m2 = p.D / (kBlockM * kCtas);
  const int tile0 = kCtas == 2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int tile_step = kCtas == 2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  // The first tile of a CTA is a first-GEMM tile of group tile0 / m1 whenever it lies in the head of the schedule
  // (decode_tile), whatever the number of groups turns out to be: its table entry is requested together with the group
  // count instead of one L2 round trip after it.
  const int g_first = min(tile0 / m1, p.gmax - 1);
  const GroupRec gr_first = p.groups[g_first];
  const int ng = *p.n_groups;
  const int m1_flags = p.H / kBlockM;  // every CTA publishes its own 128-row slice of h: flags per group
  const int lag = p.p1_only ? ng : min(p.lag, ng);
  const int n_tiles = p.p1_only ? ng * m1 : ng * (m1 + m2);  // (p1_only: lag == ng puts every tile in the head)
  constexpr int kKel = kTf32 ? kBlockK / 2 : kBlockK;  // elements per 128-byte k-block
  const int kb1 = p.D / (kKel * kps);  // pipeline stages per tile of the first GEMM
  const int kb2 = p.H / (kKel * kps);

  // arm a stage: the leader expects the bytes of both CTAs, the other CTA only announces that its loads are issued
  auto arm = [&](int st, uint32_t bytes) {
    if (kCtas == 1 || leader) ptx::mbar_arrive_expect_tx(full_bar(st), bytes);
    else ptx::mbar_arrive_cluster(full_bar(st) & ptx::kPeerBitMask);
  };
  // rows [row, ..) x k-blocks [kb * kps, (kb + 1) * kps) of a [K/64][rows][64] view
  auto load = [&](uint32_t dst, const CUtensorMap* tm, int st, int kb, int row, uint64_t pol) {
    if constexpr (kCtas == 2)
      ptx::tma_load_3d_pair(dst, tm, full_bar(st) & ptx::kPeerBitMask, 0, row, kb * kps, pol);
    else
      ptx::tma_load_3d(dst, tm, full_bar(st), 0, row, kb * kps, pol);
  };

  if (p.two_prod && (warp == 0 || warp == 2)) {
    // ============================ two TMA producers ============================
    // Warp 0 requests the weight tiles, warp 2 the token rows (x for the first GEMM, h for the second); a stage's full
    // barrier takes one arrival and the bytes of each.  The weight stream depends on nothing but the routing table, so it
    // never stops: only the token producer waits for h (or, under expert parallelism, for the peers' rows), and the ring
    // fills with the tile's weights meanwhile.  Two issuing threads also get a stage's requests out of the SM sooner: one
    // thread manages ~2.5 bulk copies per us whatever their size, two threads ~4.7 (tools/tma_issue_bench.cu,
    // profiles/r02_tma_issue_bench.txt).  Measured against the single producer on one box: -0.2 us per layer at 3 200
    // tokens, -0.5 us at 50, -1.5 us at 65 536.
    if (lane == 0) {
      const bool weights = warp == 0;
      int stage = 0;
      uint32_t phase = 0;
      bool rows_ready = !p.ep;
      for (int t = tile0; t < n_tiles; t += tile_step) {
        const Tile tl = decode_tile(t, ng, lag, m1, m2);
        const GroupRec gr = (t == tile0 && tl.g == g_first) ? gr_first : p.groups[tl.g];
        const int nkb = tl.phase == 1 ? kb1 : kb2;
        const CUtensorMap* tm;
        int row;
        if (weights) {
          tm = tl.phase == 1 ? &tm_w1 : &tm_w2;
          row = gr.expert * (tl.phase == 1 ? p.H : p.D) + (tl.mb * kCtas + static_cast<int>(cta_rank)) * kBlockM;
        } else {
          tm = tl.phase == 1 ? &tm_x : &tm_h;
          row = gr.row0 + static_cast<int>(cta_rank) * b_rows;
          if (tl.phase == 2) {
            if (!(p.dbg & 32))
              while (ptx::ld_acquire_gpu(p.h_ready + tl.g) < m1_flags) __nanosleep(32);
            ptx::fence_proxy_async_all();  // h was written through the generic proxy, TMA reads it through the async one
          } else if (!rows_ready && gr.src != p.ep_rank) {
            ep_wait_counters_1t(p.ep_peers, p.ep_peers.lay.arrive, kEpCtrlArrive, kEpErrDispatchTimeout);
            rows_ready = true;
            ptx::fence_proxy_async_all();
          }
        }
        const uint32_t dst0 = weights ? smem_a : smem_b;
        const uint32_t sbytes = weights ? a_stage_bytes : b_stage_bytes;
        const uint64_t pol = weights ? p.w_policy : ptx::kEvictLast;
        for (int kb = 0; kb < nkb; ++kb) {
          ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
          arm(stage, sbytes * kCtas);
          load(dst0 + stage * sbytes, tm, stage, kb, row, pol);
          if (++stage == stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
      if (weights) mark(3);
    }
  } else if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = (a_stage_bytes + b_stage_bytes) * kCtas;  // both CTAs' tiles land on the leader's barrier
      Tracer<kTrace> tr(p, 0);
      tr.sync();
      tr.rec(-1, kEvKernelStart);
      // expert parallelism: the token rows may still be landing (the peers' dispatch kernels push them); the first
      // weight tiles are requested at once, the rows only when every rank's arrival counter says they are there
      bool rows_ready = !p.ep;
      for (int t = tile0; t < n_tiles; t += tile_step) {
        const Tile tl = decode_tile(t, ng, lag, m1, m2);
        const GroupRec gr = (t == tile0 && tl.g == g_first) ? gr_first : p.groups[tl.g];
        tr.rec(t, kEvProdTileStart);
        const CUtensorMap* tm_a = tl.phase == 1 ? &tm_w1 : &tm_w2;
        const CUtensorMap* tm_b = tl.phase == 1 ? &tm_x : &tm_h;
        const int a_row = gr.expert * (tl.phase == 1 ? p.H : p.D) + (tl.mb * kCtas + static_cast<int>(cta_rank)) * kBlockM;
        const int b_row = gr.row0 + static_cast<int>(cta_rank) * b_rows;
        const int nkb = tl.phase == 1 ? kb1 : kb2;
        int pre = 0;
        // timing experiments (B200MOE_DBG): 8 / 16 = no token-operand loads in the first / second GEMM (garbage results:
        // how much of the time is the B operand's L2 -> SM traffic?), 32 = no wait for h (what does the hand-off cost?)
        const bool skip_b = (tl.phase == 1 && (p.dbg & 8)) || (tl.phase == 2 && (p.dbg & 16));
        const uint32_t st_bytes = skip_b ? a_stage_bytes * kCtas : tx_bytes;
        // (expert parallelism: rows this rank sent to itself were written by the kernel in front -- no counter to wait for)
        const bool dep_pending =
            (tl.phase == 2 && !(p.dbg & 32) && ptx::ld_acquire_gpu(p.h_ready + tl.g) < m1_flags) ||
            (tl.phase == 1 && !rows_ready && gr.src != p.ep_rank);
        if (tl.phase == 2 && !dep_pending) ptx::fence_proxy_async_all();  // h was written through the generic proxy
        if (dep_pending) {
          // The W2 tiles do not depend on h: fill the ring with them first, then wait until every phase-1 tile of this
          // group has published its slice of h and add the h tiles to the same stages (one full barrier per stage
          // expects both).  (Only when h is in fact late: in steady state this order would hold back the first h tile
          // until the whole ring has been re-filled with weights, a ~1.5 us bubble per tile.)
          pre = nkb < stages ? nkb : stages;
          const int stage0 = stage;
          for (int kb = 0; kb < pre; ++kb) {
            ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
            arm(stage, st_bytes);
            load(smem_a + stage * a_stage_bytes, tm_a, stage, kb, a_row, p.w_policy);
            if (++stage == stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
          if (tl.phase == 2) {
            while (ptx::ld_acquire_gpu(p.h_ready + tl.g) < m1_flags) __nanosleep(32);
          } else {
            ep_wait_counters_1t(p.ep_peers, p.ep_peers.lay.arrive, kEpCtrlArrive, kEpErrDispatchTimeout);
            rows_ready = true;
          }
          ptx::fence_proxy_async_all();  // generic-proxy writes of h / of the pushed rows -> async-proxy (TMA) reads
          int s2 = stage0;
          for (int kb = 0; kb < pre; ++kb) {
            if (!skip_b) load(smem_b + s2 * b_stage_bytes, tm_b, s2, kb, b_row, ptx::kEvictLast);
            if (++s2 == stages) s2 = 0;
          }
        }
        tr.rec(t, kEvProdDepOk);
        for (int kb = pre; kb < nkb; ++kb) {
          ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
          if (t < tile0 + 2 * tile_step) tr.rec(kb, kEvProdSlotFree);
          arm(stage, st_bytes);
          load(smem_a + stage * a_stage_bytes, tm_a, stage, kb, a_row, p.w_policy);
          if (!skip_b) load(smem_b + stage * b_stage_bytes, tm_b, stage, kb, b_row, ptx::kEvictLast);
          if (++stage == stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        tr.rec(t, kEvProdIssued);
      }
      mark(3);
      tr.rec(-1, kEvKernelEnd);
      tr.sync();
    }
  } else if (warp == 1) {
    // ============================ MMA issuer (one thread) ============================
    if (lane == 0 && (kCtas == 1 || leader)) {  // pair: the leader issues for both CTAs
      const uint32_t idesc = ptx::make_idesc(kTf32 ? 2u /*tf32*/ : 1u /*bf16*/, kBlockM * kCtas,
                                             static_cast<uint32_t>(p.bn));
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      Tracer<kTrace> tr(p, 1);
      for (int t = tile0; t < n_tiles; t += tile_step, ++it) {
        const Tile tl = decode_tile(t, ng, lag, m1, m2);
        const int nkb = tl.phase == 1 ? kb1 : kb2;
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        ptx::mbar_wait(tempty_bar(as), aphase ^ 1u);  // epilogue has drained this accumulator buffer
        ptx::tc_fence_after();
        tr.rec(t, kEvMmaAccFree);
        const uint32_t tmem_d = tmem_base + as * kAccStride;
        if ((p.dbg & 128) && t == tile0) {
          // experiment: let every stage of the ring land before the first MMA (does the first stage's slowness come from
          // the TMA writes that are still streaming into shared memory?)
          for (int s2 = 1; s2 < stages && s2 < nkb; ++s2) ptx::mbar_wait(full_bar(s2), 0);
        }
        for (int kb = 0; kb < nkb; ++kb) {
          ptx::mbar_wait(full_bar(stage), phase);
          ptx::tc_fence_after();
          if (kb == 0) tr.rec(t, kEvMmaFirstData);
          if (kb == 0 && it == 0) mark(2);
          if (t < tile0 + 2 * tile_step) tr.rec(kb, kEvMmaStageData);
          for (int j = 0; j < kps; ++j) {
            const uint64_t a_desc = ptx::make_kmajor_sw128_desc(smem_a + stage * a_stage_bytes + j * a_blk_bytes);
            const uint64_t b_desc = ptx::make_kmajor_sw128_desc(smem_b + stage * b_stage_bytes + j * b_blk_bytes);
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k) {
              // advance both descriptors by k * 16 elements * 2 B = 32 B -> +2 in 16-byte units
              if constexpr (kTf32)
                ptx::umma_tf32_ss(tmem_d, a_desc + 2u * k, b_desc + 2u * k, idesc, (kb | j | k) != 0 ? 1u : 0u);
              else if constexpr (kCtas == 2)
                ptx::umma_f16_ss_pair(tmem_d, a_desc + 2u * k, b_desc + 2u * k, idesc, (kb | j | k) != 0 ? 1u : 0u);
              else
                ptx::umma_f16_ss(tmem_d, a_desc + 2u * k, b_desc + 2u * k, idesc, (kb | j | k) != 0 ? 1u : 0u);
              if (kTrace && t == tile0 && kb == 0) tr.rec(j * 4 + k, kEvMmaInstr);
            }
          }
          if (t < tile0 + 2 * tile_step) tr.rec(kb, kEvMmaStageIssued);
          if constexpr (kCtas == 2) {
            // the slot is free in BOTH CTAs once these MMAs have read it; both epilogues get the accumulator signal
            ptx::umma_commit_pair(empty_bar(stage), 0x3);
            if (kb == nkb - 1) ptx::umma_commit_pair(tfull_bar(as), 0x3);
          } else {
            ptx::umma_commit(empty_bar(stage));  // smem slot free once these MMAs have read it
            if (kb == nkb - 1) ptx::umma_commit(tfull_bar(as));
          }
          if (++stage == stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        tr.rec(t, kEvMmaIssued);
      }
      mark(4);
    }
  } else if (warp == 3) {
    // ============================ publisher (one thread) ============================
    // Releasing a phase-1 tile's slice of h at gpu scope waits until the tile's stores have drained to L2, which under
    // full memory load takes 1-2 us.  It runs here so that the epilogue warps can start on their next tile at once.
    if (lane == 0) {
      Tracer<kTrace> tr(p, 3);
      int k = 0;  // phase-1 tiles of this CTA so far
      for (int t = tile0; t < n_tiles; t += tile_step) {
        const Tile tl = decode_tile(t, ng, lag, m1, m2);
        if (tl.phase != 1) continue;
        const int slot = k & 1;
        ptx::mbar_wait(pfull_bar(slot), (k >> 1) & 1);  // acquire.cta: the four epilogue warps' h stores
        ptx::fence_proxy_async_all();                   // generic-proxy writes -> the consumers' TMA reads
        ptx::red_release_gpu_add(p.h_ready + tl.g, 1);
        ptx::mbar_arrive(pempty_bar(slot));
        tr.rec(t, kEvEpiPublished);
        ++k;
      }
    }
  } else if (warp >= 4) {
    // ============================ epilogue (8 warps = 2 sets x 4 TMEM lane quarters) ============================
    // tcgen05.ld gives thread (q, lane) ONE feature and 32 token columns, while global memory holds token rows, so each
    // chunk of 32 columns is transposed through shared memory: the feature-major values are written column by column
    // (conflict-free, immediate offsets), then warp q of the set turns columns 8q..8q+7 into row segments of 128 features
    // with 16/8-byte vector accesses.  Set s owns the chunks c with c % 2 == s and has its own staging buffer and named
    // barrier, so the two sets never wait for each other inside a tile.
    //
    // What bounds this code is the instruction issue rate of the four schedulers (a warp-wide FP32 op occupies a
    // 16-lane pipe for two cycles; measured: ~2.5 cycles per SASS instruction with one warp per scheduler), not memory:
    // switching every store and residual load off did not change its duration.  Hence: as few instructions per
    // element as possible (SiLU = fma, tanh, fma), two warps per scheduler, and no per-element address arithmetic.
    const int q = warp & 3;                 // tcgen05.ld: warp w may touch lanes 32*(w%4) .. +31
    const int set = (warp - 4) >> 2;        // 0 or 1
    const int et = threadIdx.x - 128;       // 0 .. 255
    const int feat_l = q * 32 + lane;       // feature within the tile owned in the column phase
    const uint32_t set_bar = 1u + set;
    uint8_t* stg_raw = smem_raw + stage_off + set * kStagingBytes;
    float* stg_f = reinterpret_cast<float*>(stg_raw);
    int it = 0;
    int k1 = 0;  // phase-1 tiles so far (publisher hand-off ring)
    Tracer<kTrace> tr(p, 2);
    const bool tracer_thread = et == 0;
    for (int t = tile0; t < n_tiles; t += tile_step, ++it) {
      const Tile tl = decode_tile(t, ng, lag, m1, m2);
      const GroupRec gr = p.groups[tl.g];
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * kAccStride;
      const int feat0 = (tl.mb * kCtas + static_cast<int>(cta_rank)) * kBlockM;  // first weight row of this CTA's part
      const int nrows = gr.nrows;
      if (tl.phase == 1) {
        const float bias = p.b1 ? p.b1[static_cast<size_t>(gr.expert) * p.H + feat0 + feat_l] : 0.0f;
        const float hbias = 0.5f * bias;
        const int act = p.act;
        const int nst = (p.dbg & 2) ? 0 : nrows;
        ptx::mbar_wait(tfull_bar(as), aphase);
        ptx::tc_fence_after();
        if (tracer_thread) tr.rec(t, kEvEpiAccReady);
        int ci = 0;
        if constexpr (kTf32) {
          // fp32 h: fp32 staging (one 16 KiB buffer per set, two barriers per chunk), exact activation
          float* hb = static_cast<float*>(p.hbuf);
#pragma unroll 1
          for (int c0 = set * 32; c0 < nrows; c0 += 64) {
            uint32_t r[32];
            ptx::tmem_ld_32x32b_x32(taddr + c0, r);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float v = __uint_as_float(r[j]) + bias;
              if (act == B200MOE_ACT_SILU) v = __fdividef(v, 1.0f + __expf(-v));
              else if (act == B200MOE_ACT_RELU) v = fmaxf(v, 0.0f);
              else v = 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
              stg_f[j * kBlockM + feat_l] = ptx::round_tf32(v);  // the second GEMM would truncate it otherwise
            }
            ptx::named_bar_sync(set_bar, kSetThreads);
            float* hrow = hb + static_cast<size_t>(gr.row0 + c0 + q * 8) * p.H + feat0 + lane * 4;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int j = q * 8 + i;
              const float4 val = *reinterpret_cast<const float4*>(stg_f + j * kBlockM + lane * 4);
              st_pred_v4(hrow + static_cast<size_t>(i) * p.H,
                         make_uint4(__float_as_uint(val.x), __float_as_uint(val.y), __float_as_uint(val.z),
                                    __float_as_uint(val.w)),
                         c0 + j < nst);
            }
            ptx::named_bar_sync(set_bar, kSetThreads);
          }
        } else
#pragma unroll 1
        for (int c0 = set * 32; c0 < nrows; c0 += 64, ++ci) {
          bf16* sb = reinterpret_cast<bf16*>(stg_raw) + (ci & 1) * (32 * kBlockM);  // two 8 KB buffers per set
          uint32_t r[32];
          ptx::tmem_ld_32x32b_x32(taddr + c0, r);
          ptx::tmem_ld_wait();
          if (act == B200MOE_ACT_SILU) {
            // silu(v) = v * sigmoid(v) with sigmoid(v) = (tanh(v / 2) + 1) / 2, so with h = v / 2: h * tanh(h) + h.
            // ONE special-function op (tanh.approx, ~2^-11 relative) instead of ex2 + rcp: the 16 MUFU results per
            // clock of an SM are what this loop would otherwise wait for.  The result is rounded to bf16 (2^-9).
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float hv = fmaf(__uint_as_float(r[j]), 0.5f, hbias);
              float th;
              asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(hv));
              sb[j * kBlockM + feat_l] = __float2bfloat16_rn(fmaf(hv, th, hv));
            }
          } else if (act == B200MOE_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              sb[j * kBlockM + feat_l] = __float2bfloat16_rn(fmaxf(__uint_as_float(r[j]) + bias, 0.0f));
          } else if (act == kActNone) {
#pragma unroll
            for (int j = 0; j < 32; ++j) sb[j * kBlockM + feat_l] = __float2bfloat16_rn(__uint_as_float(r[j]) + bias);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float v = __uint_as_float(r[j]) + bias;
              sb[j * kBlockM + feat_l] = __float2bfloat16_rn(0.5f * v * (1.0f + erff(v * 0.70710678118654752f)));
            }
          }
          ptx::named_bar_sync(set_bar, kSetThreads);  // this chunk is staged; the other buffer's readers are done
          if (tracer_thread) tr.rec(t, kEvEpiChunkStaged);
          // row phase: warp q owns columns 8q .. 8q+7; a half-warp writes one 256-byte row segment of h
          {
            const int half = lane >> 4;
            const int l16 = lane & 15;
            bf16* hrow =
                static_cast<bf16*>(p.hbuf) + static_cast<size_t>(gr.row0 + c0 + q * 8 + half) * p.H + feat0 + l16 * 8;
            const size_t step = 2 * static_cast<size_t>(p.H);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int j = q * 8 + 2 * i + half;
              const uint4 val = *reinterpret_cast<const uint4*>(sb + j * kBlockM + l16 * 8);
              st_pred_v4(hrow + i * step, val, c0 + j < nst);
            }
          }
        }
        // every column this warp owns has left TMEM: hand the accumulator back to the MMA warp
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (kCtas == 2) ptx::mbar_arrive_cluster(tempty_bar(as) & ptx::kPeerBitMask);
          else ptx::mbar_arrive(tempty_bar(as));
        }
        if (tracer_thread) tr.rec(t, kEvEpiStored);
        // hand the tile to the publisher: this warp's stores are ordered before lane 0's arrive (release.cta)
        if (lane == 0) {
          const int slot = k1 & 1;
          ptx::mbar_wait(pempty_bar(slot), ((k1 >> 1) & 1) ^ 1u);
          ptx::mbar_arrive(pfull_bar(slot));
        }
        ++k1;
        ptx::named_bar_sync(set_bar, kSetThreads);  // staging buffers free before the next tile reuses them
      } else {
        const float bias = p.b2 ? p.b2[static_cast<size_t>(gr.expert) * p.D + feat0 + feat_l] : 0.0f;
        OutT* out = static_cast<OutT*>(p.out);
        const OutT* res = static_cast<const OutT*>(p.residual);
        const bool with_res = (p.fused || p.ep_fold) && res != nullptr && !(p.dbg & 4);
        const bool st2 = !(p.dbg & 1);
        // routing table of the tile's columns: output row and scale (filled while the MMAs are still running)
        ptx::named_bar_sync(3, kEpiThreads);  // every warp is done with the previous table
        for (int i = et; i < nrows; i += kEpiThreads) {
          const int row = gr.row0 + i;
          int tok = gr.orow0 + i;  // un-fused: same row of the output buffer
          float sc = 1.0f;
          if (p.fused) {
            tok = p.pos[row] / p.top_k;
            sc = p.ff_scale * (p.row_score ? p.row_score[row] : 1.0f);
          }
          if (p.ep) {
            // expert parallelism: the row's routing data came with it (written by the source GPU: L2-coherent load)
            const int2 m = __ldcg(p.ep_meta + row);
            tok = m.x;
            if (((m.x >> kEpMetaRankShift) & kEpMetaRankMask) >= p.ep_world)  // (never for a row that was really sent)
              tok = ep_meta_word(p.ep_rank, m.x & kEpMetaIndexMask, false);
            if (p.ep_fold) sc = p.ff_scale * __int_as_float(m.y);
          }
          s_tok[i] = tok;
          s_sc[i] = sc;
        }
        ptx::named_bar_sync(3, kEpiThreads);  // routing table visible to all epilogue warps
        if constexpr (sizeof(OutT) == 2) {
          // ---- 16-bit outputs.  The column phase stages  sc * (y + b2)  already rounded to the output type (8 KiB per
          // chunk, two buffers per set, ONE barrier per chunk); the row phase adds the residual with packed adds
          // (add.rn.{bf16x2,f16x2}: one rounding of the exact sum) and writes 256-byte row segments, a half-warp per row.
          // Rounding the MoE term before the residual add perturbs `out` by 2^-9 of a term that is itself ~1e-2 of the
          // residual; without a residual (sc * y alone) it is the same single rounding as before.
          const int half = lane >> 4;
          const int l16 = lane & 15;
          const size_t fo16 = static_cast<size_t>(feat0 + l16 * 8);
          // Residual rows of this set's first two chunks are requested now, while the MMAs of the tile still run (a
          // global load under full HBM load takes 1-2 us); the rows of chunk k+2 as soon as chunk k is done.
          // ALL of this set's chunks (up to four at 256-token tiles) are requested here.  Loads issued inside the chunk
          // loop instead stalled the next chunk's column phase until they returned (6.2 vs 4.4 us per 256-token tile).
          uint4 rres[4][4];
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            if (cc < 2 || nrows > 128) {  // (uniform) tiles of up to 128 tokens have two chunks per set
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int col = (2 * cc + set) * 32 + q * 8 + 2 * i + half;
                const bool v = col < nrows;
                // residual row: the token's (fused), or -- expert parallelism -- the received row itself when the source had one
                const int w = v ? s_tok[col] : 0;
                const bool hr = v && (!p.ep || w < 0);
                const int rr = p.ep ? gr.row0 + col : w;
                rres[cc][i] = ld_pred_v4(res + static_cast<size_t>(hr ? rr : 0) * p.D + fo16, hr && with_res);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i) rres[cc][i] = make_uint4(0u, 0u, 0u, 0u);
            }
          }
          ptx::mbar_wait(tfull_bar(as), aphase);
          ptx::tc_fence_after();
          if (tracer_thread) tr.rec(t, kEvEpiAccReady);
          int ci = 0;
          {
#pragma unroll
            for (int cs = 0; cs < 4; ++cs) {  // spelled out: the residual registers need static indices
              const int c0 = set * 32 + cs * 64;
              if (c0 < nrows) {               // uniform over the set
                OutT* sb = reinterpret_cast<OutT*>(stg_raw) + (ci & 1) * (32 * kBlockM);
                ++ci;
                {
                  uint32_t r[32];
                  ptx::tmem_ld_32x32b_x32(taddr + c0, r);
                  float scv[32];  // per-column scale, fetched while the TMEM load is in flight
#pragma unroll
                  for (int j4 = 0; j4 < 8; ++j4) {
                    const float4 v4 = *reinterpret_cast<const float4*>(s_sc + c0 + 4 * j4);  // broadcast reads
                    scv[4 * j4] = v4.x;
                    scv[4 * j4 + 1] = v4.y;
                    scv[4 * j4 + 2] = v4.z;
                    scv[4 * j4 + 3] = v4.w;
                  }
                  ptx::tmem_ld_wait();
#pragma unroll
                  for (int j = 0; j < 32; ++j)
                    sb[j * kBlockM + feat_l] = from_float<OutT>((__uint_as_float(r[j]) + bias) * scv[j]);
                }
                if (tracer_thread) tr.rec(t, kEvEpiChunkLd);
                ptx::named_bar_sync(set_bar, kSetThreads);  // this chunk is staged; the other buffer's readers are done
                if (tracer_thread) tr.rec(t, kEvEpiChunkStaged);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const int j = q * 8 + 2 * i + half;
                  const int col = c0 + j;
                  const bool v = col < nrows;
                  const int w = v ? s_tok[col] : 0;
                  const int tok = p.ep ? (w & kEpMetaIndexMask) : w;
                  OutT* ob = p.ep ? reinterpret_cast<OutT*>(p.ep_out[(w >> kEpMetaRankShift) & kEpMetaRankMask]) : out;
                  const uint4 a = *reinterpret_cast<const uint4*>(sb + j * kBlockM + l16 * 8);
                  uint4 o;
                  o.x = add_packed<OutT>(rres[cs][i].x, a.x);
                  o.y = add_packed<OutT>(rres[cs][i].y, a.y);
                  o.z = add_packed<OutT>(rres[cs][i].z, a.z);
                  o.w = add_packed<OutT>(rres[cs][i].w, a.w);
                  st_pred_v4(ob + static_cast<size_t>(tok) * p.D + fo16, o, v && st2);
                }
                if (tracer_thread) tr.rec(t, kEvEpiChunkDone);
              }
            }
          }
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (kCtas == 2) ptx::mbar_arrive_cluster(tempty_bar(as) & ptx::kPeerBitMask);
            else ptx::mbar_arrive(tempty_bar(as));
          }
          if (tracer_thread) tr.rec(t, kEvEpiStored);
          ptx::named_bar_sync(set_bar, kSetThreads);  // staging buffers free before the next tile reuses them
        } else {
          // ---- fp32 outputs: fp32 staging (16 KiB per chunk, one buffer per set, two barriers per chunk)
          using Io = Vec4Io<OutT>;
          const size_t fo = static_cast<size_t>(feat0 + lane * 4);
          // Residual rows of this set's first two chunks are requested now, while the MMAs of the tile still run (a global
          // load under full HBM load takes 1-2 us); the rows of chunk k+2 are requested as soon as chunk k is done.
          typename Io::raw_t rres[2][8];
  #pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
  #pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int col = (2 * cc + set) * 32 + q * 8 + i;
              const bool v = col < nrows;
              const int tok = v ? s_tok[col] : 0;
              rres[cc][i] = Io::load(res + static_cast<size_t>(tok) * p.D + fo, v && with_res);
            }
          }
          ptx::mbar_wait(tfull_bar(as), aphase);
          ptx::tc_fence_after();
          if (tracer_thread) tr.rec(t, kEvEpiAccReady);
  #pragma unroll 1
          for (int cb = set * 32; cb < nrows; cb += 128) {
  #pragma unroll
            for (int cs = 0; cs < 2; ++cs) {  // spelled out twice: the residual registers need static indices
              const int c0 = cb + cs * 64;
              if (c0 < nrows) {               // uniform over the set
                {
                  uint32_t r[32];
                  ptx::tmem_ld_32x32b_x32(taddr + c0, r);
                  ptx::tmem_ld_wait();
  #pragma unroll
                  for (int j = 0; j < 32; ++j) stg_f[j * kBlockM + feat_l] = __uint_as_float(r[j]) + bias;
                }
                if (tracer_thread) tr.rec(t, kEvEpiChunkLd);
                ptx::named_bar_sync(set_bar, kSetThreads);  // staging complete
                if (tracer_thread) tr.rec(t, kEvEpiChunkStaged);
                // row phase: warp q owns columns 8q .. 8q+7, lane l the 4 features 4l .. 4l+3 of each
                float4 a[8];
                float sc[8];
                size_t ofs[8];
                bool vv[8];
  #pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const int j = q * 8 + i;
                  const int col = c0 + j;
                  vv[i] = col < nrows;
                  const int tok = vv[i] ? s_tok[col] : 0;
                  sc[i] = vv[i] ? s_sc[col] : 0.0f;
                  a[i] = *reinterpret_cast<const float4*>(stg_f + j * kBlockM + lane * 4);
                  ofs[i] = static_cast<size_t>(tok) * p.D + fo;
                }
                ptx::named_bar_sync(set_bar, kSetThreads);  // staging buffer free again (values are in registers)
                float o[8][4];
  #pragma unroll
                for (int i = 0; i < 8; ++i) {
                  float rv[4];
                  Io::to_f(rres[cs][i], rv);
                  o[i][0] = fmaf(sc[i], a[i].x, rv[0]);
                  o[i][1] = fmaf(sc[i], a[i].y, rv[1]);
                  o[i][2] = fmaf(sc[i], a[i].z, rv[2]);
                  o[i][3] = fmaf(sc[i], a[i].w, rv[3]);
                }
  #pragma unroll
                for (int i = 0; i < 8; ++i) Io::store(out + ofs[i], o[i], vv[i] && st2);
                // residual rows of the chunk two steps ahead
  #pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const int col = c0 + 128 + q * 8 + i;
                  const bool v = col < nrows;
                  const int tok = v ? s_tok[col] : 0;
                  rres[cs][i] = Io::load(res + static_cast<size_t>(tok) * p.D + fo, v && with_res);
                }
                if (tracer_thread) tr.rec(t, kEvEpiChunkDone);
              }
            }
          }
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (kCtas == 2) ptx::mbar_arrive_cluster(tempty_bar(as) & ptx::kPeerBitMask);
            else ptx::mbar_arrive(tempty_bar(as));
          }
          if (tracer_thread) tr.rec(t, kEvEpiStored);
        }
      }
    }
  }

  ptx::tc_fence_before();
  if constexpr (kCtas == 2) ptx::cluster_sync();  // neither CTA may leave while the other can still signal its barriers
  else __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    if constexpr (kCtas == 2) ptx::tmem_dealloc_pair<kTmemCols>(tmem_base);
    else ptx::tmem_dealloc<kTmemCols>(tmem_base);
  }
  if (threadIdx.x == 0) mark(5);
  if (p.ep) {
    // Every store of this CTA into the peers' buffers precedes the barrier above: fence (one NVLink round trip), then one
    // increment of done[my rank] on every rank (functions.py:185-191's all-to-all, without the host and without a last
    // CTA: how many CTAs there are went out with the counts).
    if (threadIdx.x == 0) {
      Tracer<kTrace> tr(p, 3, 4);
      tr.rec(-2, kEvEpiChunkLd);
    }
    ep_signal(p.ep_peers, p.ep_peers.lay.done);
    if (threadIdx.x == 0) {
      Tracer<kTrace> tr(p, 3, 3);
      tr.rec(-2, kEvEpiChunkStaged);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Host side: tensor maps + launch
// ------------------------------------------------------------------------------------------------------------
void* g_trace_buf = nullptr;
int g_trace_cap = 0;

// b_rows: token rows one CTA stages per k-block (BN, or BN / 2 for CTA pairs); kps: k-blocks per stage
int stages_for(int b_rows, int kps) {
  int s = kStageBudget / (kps * (kAStageBytes + b_rows * kBlockK * 2));
  return s > kMaxStages ? kMaxStages : s;
}

size_t smem_bytes_for(int b_rows, int kps, int stages) {
  return static_cast<size_t>(stages) * kps * (kAStageBytes + b_rows * kBlockK * 2) + 8 * (2 * kMaxStages + 8) + 32 +
         2 * kStagingBytes + 2 * 256 * 4;
}

int grid_for(int gmax, int D, int H, int ctas) {
  const int m1 = H / (kBlockM * ctas), m2 = D / (kBlockM * ctas);
  long long tiles_ub = static_cast<long long>(gmax) * (m1 + m2) * ctas;  // in CTAs
  int grid = num_sms();
  if (tiles_ub < grid) grid = static_cast<int>(tiles_ub);
  if (ctas == 2) grid &= ~1;  // whole pairs
  if (grid < ctas) grid = ctas;
  return grid;
}

// CTA pairs for the compute-bound regime: full 256-token tiles and an even number of 128-row blocks per GEMM
bool use_pairs(int bn, int D, int H, bool tf32) {
  static const int pair_env = [] {
    const char* v = std::getenv("B200MOE_PAIR");
    return (v && *v) ? std::atoi(v) : 1;
  }();
  // (B200MOE_PAIR=2, experiments: pairs for 128-token tiles as well)
  return !tf32 && pair_env != 0 && (bn == 256 || (pair_env == 2 && bn == 128)) && (H / kBlockM) % 2 == 0 &&
         (D / kBlockM) % 2 == 0;
}

template <typename OutT, int kCtas, bool kTf32 = false>
cudaError_t launch_typed(const FfnLaunch& a, const CUtensorMap& tw1, const CUtensorMap& tw2, const CUtensorMap& tx,
                         const CUtensorMap& th, const FfnParams& p, cudaStream_t stream) {
  const size_t smem = smem_bytes_for(a.bn / kCtas, p.kps, p.stages);
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(ffn_kernel<OutT, 0, kCtas, kTf32>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(ffn_kernel<OutT, 1, kCtas, kTf32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(ffn_kernel<OutT, 2, kCtas, kTf32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int grid = grid_for(a.gmax, a.D, a.H, kCtas);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (kPdlFfn & pdl_mask()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (kCtas == 2) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  cudaError_t e;
  if (p.trace != nullptr)
    e = cudaLaunchKernelEx(&cfg, ffn_kernel<OutT, 1, kCtas, kTf32>, tw1, tw2, tx, th, p);
  else if (p.tl != nullptr)
    e = cudaLaunchKernelEx(&cfg, ffn_kernel<OutT, 2, kCtas, kTf32>, tw1, tw2, tx, th, p);
  else
    e = cudaLaunchKernelEx(&cfg, ffn_kernel<OutT, 0, kCtas, kTf32>, tw1, tw2, tx, th, p);
  count_launch();
  return e;
}

}  // namespace

namespace {
unsigned long long* g_tl_buf = nullptr;
int g_tl_cap = 0, g_tl_used = 0;
int g_tl_kind[4096];
}  // namespace
void set_timeline(void* dev_buf, int max_launches) {
  g_tl_buf = static_cast<unsigned long long*>(dev_buf);
  g_tl_cap = dev_buf ? (max_launches < 4096 ? max_launches : 4096) : 0;
  g_tl_used = 0;
}
unsigned long long* next_timeline_slot(int kind) {
  if (g_tl_buf == nullptr || g_tl_used >= g_tl_cap) return nullptr;
  g_tl_kind[g_tl_used] = kind;
  return g_tl_buf + static_cast<size_t>(g_tl_used++) * 148 * kTimelineMarks;
}
int timeline_kind(int slot) { return slot >= 0 && slot < g_tl_used ? g_tl_kind[slot] : 0; }

void set_ffn_trace(void* dev_buf, int records_per_cta) {
  g_trace_buf = dev_buf;
  g_trace_cap = records_per_cta;
}

int ffn_grid_ctas(int bn, int gmax, int D, int H, bool tf32) {
  return grid_for(gmax > 0 ? gmax : 1, D, H, use_pairs(bn, D, H, tf32) ? 2 : 1);
}

cudaError_t launch_ffn(const FfnLaunch& a, cudaStream_t stream) {
  if (a.n_rows <= 0) return cudaSuccess;
  if (a.D % kBlockM != 0 || a.H % kBlockM != 0) return cudaErrorInvalidValue;
  if (a.p1_only && (a.tf32 || a.ep != nullptr || a.fused)) return cudaErrorInvalidValue;
  if (a.bn % 16 != 0 || a.bn < 16 || a.bn > 256) return cudaErrorInvalidValue;
  if (a.fused && a.top_k != 1) return cudaErrorInvalidValue;
  const bool tf32 = a.tf32 != 0;
  if (tf32 && (a.out_dtype != B200MOE_F32 || a.ep != nullptr)) return cudaErrorInvalidValue;
  const bool pair = use_pairs(a.bn, a.D, a.H, tf32);
  const int ctas = pair ? 2 : 1;
  // two k-blocks per pipeline stage (one TMA instruction per operand and stage) wherever that still leaves >= 3 stages
  static const int kps_env = [] {
    const char* v = std::getenv("B200MOE_KPS");
    return (v && *v) ? std::atoi(v) : 2;
  }();
  int kps = kps_env == 1 ? 1 : 2;
  const int kel = tf32 ? kBlockK / 2 : kBlockK;  // elements per 128-byte k-block
  if (stages_for(a.bn / ctas, kps) < 3 || (a.D / kel) % kps != 0 || (a.H / kel) % kps != 0) kps = 1;
  CUtensorMap tw1, tw2, tx, th;
  if (tf32) {
    if (!make_tmap_f32_kblocks(&tw1, a.W1, static_cast<uint64_t>(a.E) * a.H, a.D, kBlockM, kps) ||
        !make_tmap_f32_kblocks(&tw2, a.W2, static_cast<uint64_t>(a.E) * a.D, a.H, kBlockM, kps) ||
        !make_tmap_f32_kblocks(&tx, a.xbuf, a.n_rows, a.D, a.bn, kps) ||
        !make_tmap_f32_kblocks(&th, a.hbuf, a.n_rows, a.H, a.bn, kps)) {
      return cudaErrorInvalidValue;
    }
  } else if (a.p1_only) {
    if (!make_tmap_bf16_kblocks(&tw1, a.W1, static_cast<uint64_t>(a.E) * a.H, a.D, kBlockM, kps) ||
        !make_tmap_bf16_kblocks(&tx, a.xbuf, a.n_rows, a.D, a.bn / ctas, kps))
      return cudaErrorInvalidValue;
    tw2 = tw1;  // (never used: there are no second-GEMM tiles)
    th = tx;
  } else if (!make_tmap_bf16_kblocks(&tw1, a.W1, static_cast<uint64_t>(a.E) * a.H, a.D, kBlockM, kps) ||
      !make_tmap_bf16_kblocks(&tw2, a.W2, static_cast<uint64_t>(a.E) * a.D, a.H, kBlockM, kps) ||
      !make_tmap_bf16_kblocks(&tx, a.xbuf, a.n_rows, a.D, a.bn / ctas, kps) ||
      !make_tmap_bf16_kblocks(&th, a.hbuf, a.n_rows, a.H, a.bn / ctas, kps)) {
    return cudaErrorInvalidValue;
  }
  FfnParams p;
  p.ep = 0;
  p.ep_world = p.ep_rank = 0;
  p.ep_peers = EpPeers{};
  for (int r = 0; r < kMaxEpWorld; ++r) p.ep_out[r] = nullptr;
  p.ep_fold = 0;
  p.ep_meta = nullptr;
  if (a.ep != nullptr) {
    if (a.fused || a.out_dtype != B200MOE_BF16) return cudaErrorInvalidValue;
    p.ep = 1;
    p.ep_world = a.ep->world;
    p.ep_rank = a.ep->rank;
    p.ep_peers = *a.ep;
    p.ep_fold = a.ep_fold ? 1 : 0;
    p.ep_meta = reinterpret_cast<const int2*>(a.ep->base[a.ep->rank] + a.ep->lay.meta);
    // fold: the same offset inside every rank's symmetric buffer as this rank's own `out` (symmetric allocation)
    for (int r = 0; r < a.ep->world; ++r) p.ep_out[r] = a.ep->base[r] + (a.ep_fold ? a.ep_out_off : a.ep->lay.ret_y);
  }
  p.clear_ptr = a.clear_ptr;
  p.clear_ints = a.clear_ptr ? a.clear_ints : 0;
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
  p.kps = kps;
  p.p1_only = a.p1_only ? 1 : 0;
  p.stages = stages_for(a.bn / ctas, kps);
  {
    static const int dbg = [] {
      const char* v = std::getenv("B200MOE_DBG");
      return (v && *v) ? std::atoi(v) : 0;
    }();
    p.dbg = dbg;
    p.pdl_trigger = (pdl_trigger() & kPdlFfn) ? 1 : 0;
    static const int warm = [] {
      const char* v = std::getenv("B200MOE_WARM");
      return (v && *v) ? std::atoi(v) : 1;
    }();
    p.warm_mma = warm;
    static const int two = [] {
      const char* v = std::getenv("B200MOE_2PROD");
      return (v && *v) ? std::atoi(v) : 1;   // (0 = the single-producer code below, kept for A/B runs and the traces)
    }();
    p.two_prod = (two && g_trace_buf == nullptr) ? 1 : 0;  // (the per-event trace instruments the single producer)
  }
  // about one token tile per expert: every weight tile is read exactly once, so it can leave L2 right after
  p.w_policy = (static_cast<long long>(a.n_rows) <= static_cast<long long>(a.E) * a.bn) ? ptx::kEvictFirst
                                                                                      : ptx::kEvictNormal;
  p.trace = static_cast<uint4*>(g_trace_buf);
  p.trace_cap = g_trace_cap;
  p.tl = next_timeline_slot(2);
  // phase-2 tiles trail their group's phase-1 tiles by ~3 waves of the grid
  const int m1 = a.H / (kBlockM * ctas);
  p.lag = (3 * (num_sms() / ctas) + m1 - 1) / m1;
  p.gmax = a.gmax > 0 ? a.gmax : 1;
  if (tf32) return launch_typed<float, 1, true>(a, tw1, tw2, tx, th, p, stream);
  switch (a.out_dtype) {
    case B200MOE_F32:
      return pair ? launch_typed<float, 2>(a, tw1, tw2, tx, th, p, stream)
                  : launch_typed<float, 1>(a, tw1, tw2, tx, th, p, stream);
    case B200MOE_F16:
      return pair ? launch_typed<__half, 2>(a, tw1, tw2, tx, th, p, stream)
                  : launch_typed<__half, 1>(a, tw1, tw2, tx, th, p, stream);
    case B200MOE_BF16:
      return pair ? launch_typed<bf16, 2>(a, tw1, tw2, tx, th, p, stream)
                  : launch_typed<bf16, 1>(a, tw1, tw2, tx, th, p, stream);
    default:
      return cudaErrorInvalidValue;
  }
}

}  // namespace b200moe
