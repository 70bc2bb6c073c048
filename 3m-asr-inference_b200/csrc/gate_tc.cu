// Tensor-core gate (bf16 activations, E <= 32): router GEMM on tcgen05 + softmax / top-k + per-32-token histograms.
//
// Same reference behaviour as gate.cu (router MatMul + SoftmaxTopKPluginDynamic, NaiveGate); this kernel exists because
// the SIMT version is FMA-bound at ~1/3 of the HBM roofline (32 FMAs per loaded activation element), while the router
// GEMM [tokens, R] x [R, E] is a perfectly ordinary tensor-core problem:
//   * A = activations, 128 tokens x 64 k per stage, TMA-loaded straight from `embed` (k < Demb) and `x` (k >= Demb):
//     the reference's concat (positionwise_feed_forward.py:225) costs nothing;
//   * B = the router matrix packed once as bf16 K-major [2*32, R]: rows 0..31 hold hi = bf16(Wr^T), rows 32..63 hold
//     lo = bf16(Wr^T - hi).  bf16 x bf16 products are exact in the fp32 accumulator, so logits = acc_hi + acc_lo carry
//     the fp32 router weights to ~2^-17 relative -- the arg-max agrees with an fp32/fp64 oracle whenever the top-1/2
//     margin exceeds that, exactly like the SIMT kernel;
//   * D = [128 tokens (TMEM lanes), 64 columns] fp32 in TMEM, double buffered; one epilogue thread per token reads its
//     64 columns and does softmax / arg-max / top-k entirely in registers (no shuffles, no shared memory);
//   * each epilogue warp (32 consecutive tokens) also emits the expert histogram of its tokens, which is what the
//     dispatch kernel needs as per-chunk counts -- the separate count kernel disappears.
#include <math_constants.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

namespace b200moe {

namespace {

constexpr int kTokTile = 128;   // tokens per tile == UMMA M
constexpr int kBlkK = 64;
constexpr int kNCols = 64;      // UMMA N: 32 hi + 32 lo expert columns
// A pipeline stage holds kKpsG k-blocks of both operands, each brought by ONE TMA instruction through a
// [K/64][rows][64] view: one producer thread issues ~3.7 bulk copies per us, so with one 16 KiB + one 8 KiB instruction
// per k-block the issue rate, not HBM, paced the kernel (4.2 TB/s of activations at 65 536 tokens).
constexpr int kKpsG = 2;
constexpr int kStagesG = 4;
constexpr int kABlk = kTokTile * kBlkK * 2;    // 16 KiB: 128 tokens of one k-block
constexpr int kBBlk = kNCols * kBlkK * 2;      // 8 KiB: 64 packed router rows of one k-block
constexpr int kAStage = kKpsG * kABlk;         // 32 KiB
constexpr int kBStage = kKpsG * kBBlk;         // 16 KiB
constexpr int kThreadsG = 256;
constexpr uint32_t kTmemColsG = 128;

struct GateTcParams {
  const float* br;
  const int* x_len;
  int* idx;
  float* score;
  int* hist32;  // [ceil(S/32), E] per-32-token expert counts, or null
  int S, T, D, Demb, E, top_k, gate_mode;
  // optional DRAM -> L2 prefetch of the layer's expert weights, spread over all CTAs (the gate itself moves ~2 KiB per
  // token, so HBM is otherwise idle while the gate and the dispatch run)
  const uint8_t* pf_ptr[2];
  unsigned long long pf_bytes[2];
  int pdl_trigger;
  int pf_mode;  // 0 off, 1 before the dependency wait (weights are constants), 2 after it
  int kps;      // k-blocks per pipeline stage (see kKpsG)
};

constexpr unsigned kPfChunk = 16384;

__device__ __forceinline__ void prefetch_weights(const GateTcParams& p, int lane) {
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const unsigned long long n = p.pf_bytes[r];
    if (p.pf_ptr[r] == nullptr || n == 0) continue;
    const unsigned long long nchunks = (n + kPfChunk - 1) / kPfChunk;
    for (unsigned long long c = static_cast<unsigned long long>(blockIdx.x) * 32 + lane; c < nchunks;
         c += static_cast<unsigned long long>(gridDim.x) * 32) {
      const unsigned long long off = c * kPfChunk;
      const unsigned long long left = n - off;
      const unsigned bytes = left < kPfChunk ? static_cast<unsigned>(left & ~15ull) : kPfChunk;
      if (bytes) ptx::prefetch_l2_bulk(p.pf_ptr[r] + off, bytes);
    }
  }
}

__global__ void __launch_bounds__(kThreadsG, 1)
gate_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_e,
               const __grid_constant__ CUtensorMap tm_w, const GateTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_base + kStagesG * kAStage;
  const uint32_t bar_base = smem_b + kStagesG * kBStage;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStagesG + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kStagesG + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kStagesG + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStagesG + 4);
  const uint32_t misc_off = tmem_slot + 16u - ptx::smem_u32(smem_raw);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + misc_off - 16u);
  float* s_br = reinterpret_cast<float*>(smem_raw + misc_off);            // [32]
  int* s_hist = reinterpret_cast<int*>(smem_raw + misc_off + 32 * 4);     // [4][32]

  if (p.pdl_trigger) ptx::pdl_launch_dependents();  // dispatch may start its prologue; it blocks in its pdl_wait()
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tm_x);
    ptx::prefetch_tensormap(&tm_e);
    ptx::prefetch_tensormap(&tm_w);
  }
  if (warp == 3 && p.pf_mode == 1) prefetch_weights(p, lane);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStagesG; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(tfull_bar(s), 1);
      ptx::mbar_init(tempty_bar(s), 4);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) ptx::tmem_alloc<kTmemColsG>(tmem_slot);
  if (warp == 3) s_br[lane] = (p.br != nullptr && lane < p.E) ? p.br[lane] : 0.0f;
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int n_tiles = (p.S + kTokTile - 1) / kTokTile;
  const int kps = p.kps;                        // k-blocks per stage: kKpsG, or 1 when Demb / 64 or D / 64 is odd
  const int kb_e = p.Demb / (kBlkK * kps);      // stages of the embed part
  const int nkb = (p.Demb + p.D) / (kBlkK * kps);

  if (warp == 3 && p.pf_mode == 2) {
    ptx::pdl_wait();
    prefetch_weights(p, lane);
  }
  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      ptx::pdl_wait();  // x is the previous layer's output; idx / score / hist32 may still be read by its kernels
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        for (int kb = 0; kb < nkb; ++kb) {
          ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
          ptx::mbar_arrive_expect_tx(full_bar(stage), static_cast<uint32_t>(kps) * (kABlk + kBBlk));
          if (kb < kb_e)
            ptx::tma_load_3d(smem_a + stage * kAStage, &tm_e, full_bar(stage), 0, t * kTokTile, kb * kps,
                             ptx::kEvictFirst);
          else
            ptx::tma_load_3d(smem_a + stage * kAStage, &tm_x, full_bar(stage), 0, t * kTokTile, (kb - kb_e) * kps,
                             ptx::kEvictNormal);  // x is read again by the dispatch kernel
          ptx::tma_load_3d(smem_b + stage * kBStage, &tm_w, full_bar(stage), 0, 0, kb * kps, ptx::kEvictLast);
          if (++stage == kStagesG) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = ptx::make_idesc(1u /*bf16*/, kTokTile, kNCols);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        ptx::mbar_wait(tempty_bar(as), aphase ^ 1u);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * kNCols;
        for (int kb = 0; kb < nkb; ++kb) {
          ptx::mbar_wait(full_bar(stage), phase);
          ptx::tc_fence_after();
          for (int j = 0; j < kps; ++j) {
            const uint64_t a_desc = ptx::make_kmajor_sw128_desc(smem_a + stage * kAStage + j * kABlk);
            const uint64_t b_desc = ptx::make_kmajor_sw128_desc(smem_b + stage * kBStage + j * kBBlk);
#pragma unroll
            for (int k = 0; k < kBlkK / 16; ++k)
              ptx::umma_f16_ss(tmem_d, a_desc + 2u * k, b_desc + 2u * k, idesc, (kb | j | k) != 0 ? 1u : 0u);
          }
          ptx::umma_commit(empty_bar(stage));
          if (kb == nkb - 1) ptx::umma_commit(tfull_bar(as));
          if (++stage == kStagesG) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    int it = 0;
    ptx::pdl_wait();
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int tok = t * kTokTile + q * 32 + lane;
      bool valid = tok < p.S;
      if (valid && p.x_len != nullptr) valid = (tok % p.T) < p.x_len[tok / p.T];
      s_hist[q * 32 + lane] = 0;
      ptx::mbar_wait(tfull_bar(as), aphase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * kNCols;
      uint32_t hi[32], lo[32];
      ptx::tmem_ld_32x32b_x32(taddr, hi);
      ptx::tmem_ld_32x32b_x32(taddr + 32, lo);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(tempty_bar(as));

      float l[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        l[e] = __uint_as_float(hi[e]) + __uint_as_float(lo[e]) + s_br[e];
        if (e >= p.E) l[e] = -CUDART_INF_F;
      }
      const int kk = p.top_k;
      float first = 0.0f;
      float sel_v[8];
      int sel_i[8];
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        if (s < kk) {
          float bv = l[0];
          int bi = 0;
#pragma unroll
          for (int e = 1; e < 32; ++e) {
            if (l[e] > bv) {  // strict: the lowest index wins exact ties
              bv = l[e];
              bi = e;
            }
          }
          sel_v[s] = bv;
          sel_i[s] = bi;
          if (s == 0) first = bv;
          if (s + 1 < kk) {
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (e == bi) l[e] = -CUDART_INF_F;
          }
        }
      }
      float denom = 0.0f;
      if (p.gate_mode == B200MOE_GATE_3M) {
        // softmax over ALL experts relative to the maximum (kk == 1, so l[] is still intact)
#pragma unroll
        for (int e = 0; e < 32; ++e) denom += expf(l[e] - first);  // exp(-inf) = 0 for the padded experts
      } else {
#pragma unroll
        for (int s = 0; s < 8; ++s)
          if (s < kk) denom += expf(sel_v[s] - first);
      }
      if (tok < p.S) {
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          if (s < kk) {
            p.idx[static_cast<size_t>(tok) * kk + s] = valid ? sel_i[s] : -1;
            p.score[static_cast<size_t>(tok) * kk + s] = valid ? expf(sel_v[s] - first) / denom : 0.0f;
          }
        }
      }
      if (p.hist32 != nullptr) {
        // expert histogram of this warp's 32 tokens (all top_k entries), for the dispatch kernel
        __syncwarp();
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          if (s < kk) {
            const int e = valid ? sel_i[s] : -1;
            const unsigned peers = __match_any_sync(0xffffffffu, e);
            if (e >= 0 && lane == __ffs(peers) - 1) s_hist[q * 32 + e] += __popc(peers);
            __syncwarp();
          }
        }
        const int chunk = t * (kTokTile / 32) + q;
        if (chunk * 32 < p.S && lane < p.E) p.hist32[static_cast<size_t>(chunk) * p.E + lane] = s_hist[q * 32 + lane];
        __syncwarp();
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<kTmemColsG>(tmem_base);
  }
}

__global__ void __launch_bounds__(256)
pack_router_kernel(const float* __restrict__ Wr, int R, int E, bf16* __restrict__ packed) {
  // packed [64, R]: row e = bf16(Wr[:, e]) (hi), row 32 + e = bf16(Wr[:, e] - hi) (lo); rows >= E are zero
  const int n = 64 * R;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int row = i / R;
    const int k = i - row * R;
    const int e = row & 31;
    float v = 0.0f;
    if (e < E) {
      const float w = Wr[static_cast<size_t>(k) * E + e];
      const float hi = __bfloat162float(__float2bfloat16_rn(w));
      v = row < 32 ? hi : (w - hi);
    }
    packed[i] = __float2bfloat16_rn(v);
  }
}

// norm_ff folded into the router (route.cu, kLn): rows k >= R - D of Wr are scaled by gamma[k - (R - D)] before the hi / lo
// split; block 0 then appends c1[e] = sum_k (hi + lo)[k, e] over those rows (what the tensor cores will actually multiply)
// and c0[e] = sum_k beta[k] * Wr_x[k, e].
__global__ void __launch_bounds__(256)
pack_router_ln_kernel(const float* __restrict__ Wr, int R, int E, int D, const float* __restrict__ gamma,
                      const float* __restrict__ beta, bf16* __restrict__ packed, float* __restrict__ c) {
  const int n = 64 * R;
  const int kx = R - D;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int row = i / R;
    const int k = i - row * R;
    const int e = row & 31;
    float v = 0.0f;
    if (e < E) {
      float w = Wr[static_cast<size_t>(k) * E + e];
      if (k >= kx) w *= gamma[k - kx];
      const float hi = __bfloat162float(__float2bfloat16_rn(w));
      v = row < 32 ? hi : (w - hi);
    }
    packed[i] = __float2bfloat16_rn(v);
  }
  if (blockIdx.x == 0 && threadIdx.x < 64) {
    const int e = threadIdx.x & 31;
    double acc = 0.0;
    if (e < E) {
      for (int k = kx; k < R; ++k) {
        const float w0 = Wr[static_cast<size_t>(k) * E + e];
        if (threadIdx.x < 32) {
          const float w = w0 * gamma[k - kx];
          const float hi = __bfloat162float(__float2bfloat16_rn(w));
          const float lo = __bfloat162float(__float2bfloat16_rn(w - hi));
          acc += static_cast<double>(hi) + static_cast<double>(lo);
        } else {
          acc += static_cast<double>(beta[k - kx]) * static_cast<double>(w0);
        }
      }
    }
    c[threadIdx.x] = static_cast<float>(acc);  // [0, 32): c1, [32, 64): c0
  }
}

}  // namespace

size_t router_ln_pack_bytes(int R) { return static_cast<size_t>(64) * R * sizeof(bf16) + 64 * sizeof(float); }

cudaError_t launch_pack_router_ln(const float* Wr, int R, int E, int D, const float* gamma, const float* beta,
                                  void* packed, cudaStream_t stream) {
  if (E > 32 || R < 1 || D < 1 || D > R) return cudaErrorInvalidValue;
  float* c = reinterpret_cast<float*>(static_cast<uint8_t*>(packed) + static_cast<size_t>(64) * R * sizeof(bf16));
  pack_router_ln_kernel<<<64, 256, 0, stream>>>(Wr, R, E, D, gamma, beta, static_cast<bf16*>(packed), c);
  count_launch();
  return cudaGetLastError();
}

bool gate_tc_supported(int D, int Demb, int E, int top_k, int dtype) {
  return dtype == B200MOE_BF16 && E <= 32 && top_k <= 8 && D % kBlkK == 0 && Demb % kBlkK == 0 && D > 0;
}

size_t router_pack_bytes(int R) { return static_cast<size_t>(64) * R * sizeof(bf16); }

cudaError_t launch_pack_router(const float* Wr, int R, int E, void* packed, cudaStream_t stream) {
  if (E > 32 || R < 1) return cudaErrorInvalidValue;
  pack_router_kernel<<<64, 256, 0, stream>>>(Wr, R, E, static_cast<bf16*>(packed));
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_gate_tc(const void* x, const void* embed, const void* wr_packed, const float* br, const int* x_len,
                           int B, int T, int D, int Demb, int E, int top_k, int gate_mode, int* idx, float* score,
                           int* hist32, const void* pf0, size_t pf0_bytes, const void* pf1, size_t pf1_bytes,
                           cudaStream_t stream) {
  const int S = B * T;
  if (S == 0) return cudaSuccess;
  if (embed == nullptr) Demb = 0;
  if (!gate_tc_supported(D, Demb, E, top_k, B200MOE_BF16)) return cudaErrorInvalidValue;
  CUtensorMap tx, te, tw;
  const int kps = ((D / kBlkK) % kKpsG == 0 && (Demb / kBlkK) % kKpsG == 0) ? kKpsG : 1;
  if (!make_tmap_bf16_kblocks(&tx, x, S, D, kTokTile, kps)) return cudaErrorInvalidValue;
  if (Demb > 0) {
    if (!make_tmap_bf16_kblocks(&te, embed, S, Demb, kTokTile, kps)) return cudaErrorInvalidValue;
  } else {
    te = tx;
  }
  if (!make_tmap_bf16_kblocks(&tw, wr_packed, 64, static_cast<uint64_t>(D + Demb), kNCols, kps))
    return cudaErrorInvalidValue;
  GateTcParams p;
  p.kps = kps;
  p.br = br;
  p.x_len = x_len;
  p.idx = idx;
  p.score = score;
  p.hist32 = hist32;
  p.S = S;
  p.T = T;
  p.D = D;
  p.Demb = Demb;
  p.E = E;
  p.top_k = top_k;
  p.gate_mode = gate_mode;
  const size_t smem = 1024 + kStagesG * (kAStage + kBStage) + 8 * (2 * kStagesG + 4) + 16 + 32 * 4 + 4 * 32 * 4 + 64;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gate_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int n_tiles = (S + kTokTile - 1) / kTokTile;
  int grid = n_tiles < num_sms() ? n_tiles : num_sms();
  p.pf_mode = 0;
  p.pdl_trigger = (pdl_trigger() & kPdlGate) ? 1 : 0;
  p.pf_ptr[0] = p.pf_ptr[1] = nullptr;
  p.pf_bytes[0] = p.pf_bytes[1] = 0;
  if (prefetch_mode() != 0 && (pf0 != nullptr || pf1 != nullptr)) {
    p.pf_mode = prefetch_mode() == 2 ? 2 : 1;
    p.pf_ptr[0] = static_cast<const uint8_t*>(pf0);
    p.pf_bytes[0] = pf0 ? pf0_bytes : 0;
    p.pf_ptr[1] = static_cast<const uint8_t*>(pf1);
    p.pf_bytes[1] = pf1 ? pf1_bytes : 0;
    grid = num_sms();  // CTAs without a token tile only prefetch
  }
  // PDL only pays when there is work to do before the dependency (the weight prefetch); otherwise ordinary ordering
  cudaError_t e = launch_kernel(gate_tc_kernel, dim3(grid), dim3(kThreadsG), smem, stream,
                                p.pf_mode == 1 ? kPdlGate : 0, tx, te, tw, p);
  count_launch();
  return e;
}

}  // namespace b200moe
