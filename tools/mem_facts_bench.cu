// Micro-benchmarks (GPU box only) for the facts the weight-bound expert kernel is designed around:
//   1. l2hit    TMA bulk streaming out of an L2-resident window, every CTA its own addresses / groups of 8 CTAs the same
//               addresses (how the token operand of a group is read): what does L2 -> SM deliver next to HBM -> SM?
//   2. mixed    half the stages from HBM (unique addresses), half from the L2 window: do the two add up?
//   3. latency  a small dependent global load issued while X KB of bulk loads are in flight on the same SM (flag polls,
//               group records, instruction fetches behind a deep weight prefetch)
//   4. dsmem    cp.async.bulk shared::cta -> shared::cluster between the CTAs of a cluster, and plain 16-byte remote stores
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I3m-asr-inference_b200/csrc tools/mem_facts_bench.cu -o gpurun_out/mem_facts_bench
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ptx.cuh"

using namespace b200moe;

__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :
               : "r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// mode 0: every CTA streams its own part of [base, base + window) (window > L2: HBM; window small: L2 hits after pass 1)
// mode 1: CTAs 8g .. 8g+7 stream the SAME addresses (shared token operand)
// mode 2: even stages from the big HBM buffer, odd stages from the L2 window
__global__ void __launch_bounds__(128, 1)
stream_kernel(const uint8_t* hbm, size_t hbm_bytes, const uint8_t* win, size_t win_bytes, int mode, int stage_bytes,
              int stages, long long units_per_cta, unsigned long long* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = ptx::smem_u32(smem);
  const uint32_t bar_base = sbase + stages * stage_bytes;
  auto full = [&](int s) { return bar_base + 8u * s; };
  auto empty = [&](int s) { return bar_base + 8u * (16 + s); };
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      ptx::mbar_init(full(s), 1);
      ptx::mbar_init(empty(s), 1);
    }
    ptx::fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int st = 0;
    uint32_t ph = 0;
    const int who = mode == 1 ? blockIdx.x / 8 : blockIdx.x;
    const int n_who = mode == 1 ? (gridDim.x + 7) / 8 : gridDim.x;
    for (long long u = 0; u < units_per_cta; ++u) {
      ptx::mbar_wait(empty(st), ph ^ 1u);
      ptx::mbar_arrive_expect_tx(full(st), stage_bytes);
      const size_t lin = (static_cast<size_t>(u) * n_who + who) * stage_bytes;
      const uint8_t* src;
      if (mode == 2 && (u & 1) == 0) src = hbm + lin % hbm_bytes;
      else if (mode == 2 || win_bytes < hbm_bytes) src = win + lin % win_bytes;
      else src = hbm + lin % hbm_bytes;
      bulk_load(sbase + st * stage_bytes, src, stage_bytes, full(st));
      if (++st == stages) {
        st = 0;
        ph ^= 1u;
      }
    }
  } else if (threadIdx.x == 32) {
    int st = 0;
    uint32_t ph = 0;
    unsigned long long acc = 0;
    for (long long u = 0; u < units_per_cta; ++u) {
      ptx::mbar_wait(full(st), ph);
      acc += *reinterpret_cast<volatile unsigned*>(smem + st * stage_bytes);
      ptx::mbar_arrive(empty(st));
      if (++st == stages) {
        st = 0;
        ph ^= 1u;
      }
    }
    if (acc == 0x1234567ull) sink[0] = acc;
  }
}

// X KB of bulk loads in flight (unique HBM addresses), then one dependent 4-byte load: cycles until it returns.
// probe_hit: the probed word was touched by a previous pass (L2 hit) or not.
__global__ void __launch_bounds__(128, 1)
latency_kernel(const uint8_t* hbm, size_t hbm_bytes, const unsigned* probe, int inflight_kb, int salt,
               unsigned* out_cycles, unsigned long long* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = ptx::smem_u32(smem);
  const uint32_t bar = sbase + 200 * 1024;
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar, 1);
    ptx::fence_mbar_init();
    const uint32_t bytes = static_cast<uint32_t>(inflight_kb) * 1024u;
    if (bytes) ptx::mbar_arrive_expect_tx(bar, bytes);
    const size_t base = ((static_cast<size_t>(blockIdx.x) + static_cast<size_t>(salt) * gridDim.x) * 256 * 1024) % hbm_bytes;
    for (uint32_t o = 0; o < bytes; o += 16384) bulk_load(sbase + o, hbm + base + o, 16384, bar);
    const long long t0 = clock64();
    unsigned v;
    asm volatile("ld.global.cv.u32 %0, [%1];" : "=r"(v) : "l"(probe + (blockIdx.x + salt * 977) * 64 % (1 << 20)));
    unsigned w;
    asm volatile("mov.u32 %0, %1;" : "=r"(w) : "r"(v));  // (volatile: ordered before the clock read, and it needs v)
    const long long t1 = clock64() + (w == 0x12345 ? 1 : 0);
    out_cycles[blockIdx.x] = static_cast<unsigned>(t1 - t0);
    if (bytes) ptx::mbar_wait(bar, 0);
    out_cycles[gridDim.x + blockIdx.x] = static_cast<unsigned>(clock64() - t0);
    if (v == 0x7654321) sink[0] = v;
  }
}

// DSMEM: every CTA of a cluster pushes `bytes` from its own shared memory into the next CTA's, `reps` times.
// mode 0: cp.async.bulk shared::cta -> shared::cluster (16 KiB chunks, completion on the destination's mbarrier)
// mode 1: st.shared::cluster.v4 by 128 threads
__global__ void __launch_bounds__(128, 1)
dsmem_kernel(int mode, int bytes, int reps, unsigned* out_cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = ptx::smem_u32(smem);
  const uint32_t src = sbase, dst = sbase + 96 * 1024, bar = sbase + 200 * 1024;
  uint32_t csize, crank;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csize));
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const uint32_t peer = (crank + 1) % csize;
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar, 1);
    ptx::fence_mbar_init();
  }
  ptx::cluster_sync();
  uint32_t rdst, rbar;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rdst) : "r"(dst), "r"(peer));
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(bar), "r"(peer));
  const long long t0 = clock64();
  if (mode == 0) {
    if (threadIdx.x == 0) {
      for (int r = 0; r < reps; ++r) {
        ptx::mbar_arrive_expect_tx(bar, bytes);  // what THIS CTA will receive from its predecessor
        for (int o = 0; o < bytes; o += 16384)
          asm volatile(
              "cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(rdst + o),
              "r"(src + o), "r"(16384), "r"(rbar)
              : "memory");
        ptx::mbar_wait(bar, r & 1);
      }
    }
  } else {
    for (int r = 0; r < reps; ++r) {
      for (int o = threadIdx.x * 16; o < bytes; o += 128 * 16)
        asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(rdst + o), "r"(r) : "memory");
      ptx::cluster_sync();
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) out_cycles[blockIdx.x] = static_cast<unsigned>(clock64() - t0);
  ptx::cluster_sync();
}

static float time_ms(cudaEvent_t e0, cudaEvent_t e1) {
  float ms = 0;
  cudaEventSynchronize(e1);
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  const size_t hbm_bytes = size_t(1) << 30;
  const size_t win_bytes = size_t(24) << 20;  // well inside the 126 MB L2
  uint8_t *hbm = nullptr, *win = nullptr;
  unsigned long long* sink = nullptr;
  unsigned* cyc = nullptr;
  cudaMalloc(&hbm, hbm_bytes);
  cudaMalloc(&win, win_bytes);
  cudaMalloc(&sink, 8);
  cudaMalloc(&cyc, 4096 * 4);
  cudaMemset(hbm, 1, hbm_bytes);
  cudaMemset(win, 2, win_bytes);
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(latency_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(dsmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(dsmem_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  int nsm = 148, khz = 0;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  printf("SMs %d, nominal SM clock %.0f MHz\n", nsm, khz / 1e3);

  printf("\n== streaming (cp.async.bulk, 1 GiB moved per run) ==\n%-26s %6s %6s %8s %10s\n", "source", "stage", "stages",
         "KB/SM", "GB/s");
  struct Case { const char* name; int mode; bool window; };
  const Case cases[] = {{"HBM, own addresses", 0, false}, {"L2 window, own addresses", 0, true},
                        {"L2 window, 8 CTAs share", 1, true}, {"HBM, 8 CTAs share", 1, false},
                        {"half HBM / half L2 window", 2, true}};
  for (const Case& c : cases)
    for (int stage_kb : {16, 32})
      for (int stages : {2, 4, 6, 12}) {
        const int stage_bytes = stage_kb * 1024;
        const size_t smem = size_t(stages) * stage_bytes + 8 * 32 + 64;
        if (smem > 220 * 1024 || (stage_kb == 32 && stages == 12)) continue;
        const long long units = static_cast<long long>(hbm_bytes / stage_bytes / nsm);
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
          cudaEventRecord(e0);
          stream_kernel<<<nsm, 128, smem>>>(hbm, hbm_bytes, win, c.window ? win_bytes : hbm_bytes, c.mode, stage_bytes,
                                            stages, units, sink);
          cudaEventRecord(e1);
          best = std::min(best, time_ms(e0, e1));
        }
        if (cudaGetLastError() != cudaSuccess) { printf("error\n"); return 1; }
        const double moved = double(units) * nsm * stage_bytes;
        printf("%-26s %5dK %6d %8d %10.1f\n", c.name, stage_kb, stages, stages * stage_kb, moved / (best * 1e-3) / 1e9);
      }

  printf("\n== streaming from HBM with fewer CTAs than SMs (per-SM ingest when the others are idle) ==\n%6s %6s %6s %10s %12s\n",
         "CTAs", "stage", "stages", "GB/s", "GB/s per SM");
  for (int grid : {8, 16, 32, 64, 92, 128, 148})
    for (int cfg = 0; cfg < 3; ++cfg) {
      const int stage_kb = cfg == 0 ? 16 : 32, stages = cfg == 0 ? 12 : (cfg == 1 ? 4 : 6);
      const int stage_bytes = stage_kb * 1024;
      const size_t smem = size_t(stages) * stage_bytes + 8 * 32 + 64;
      const long long units = static_cast<long long>((hbm_bytes / 4) / stage_bytes / grid);
      float best = 1e30f;
      for (int rep = 0; rep < 3; ++rep) {
        cudaMemsetAsync(win, rep, win_bytes);
        cudaEventRecord(e0);
        stream_kernel<<<grid, 128, smem>>>(hbm + (size_t(rep) << 28), hbm_bytes / 4, win, hbm_bytes / 4, 0, stage_bytes, stages,
                                           units, sink);
        cudaEventRecord(e1);
        best = std::min(best, time_ms(e0, e1));
      }
      const double moved = double(units) * grid * stage_bytes;
      printf("%6d %5dK %6d %10.1f %12.1f\n", grid, stage_kb, stages, moved / (best * 1e-3) / 1e9,
             moved / (best * 1e-3) / 1e9 / grid);
    }

  printf("\n== one dependent 4-byte load behind X KB of bulk loads on the same SM (all SMs at once) ==\n");
  printf("%8s %10s %14s %14s %16s\n", "X KB", "probe", "median cyc", "max cyc", "bulk done (cyc)");
  std::vector<unsigned> h(2 * nsm);
  unsigned* probe = reinterpret_cast<unsigned*>(win);
  int salt = 0;
  for (int hit = 0; hit < 2; ++hit)
    for (int kb : {0, 16, 32, 64, 128, 192}) {
      std::vector<unsigned> med, done;
      for (int rep = 0; rep < 5; ++rep) {
        ++salt;
        if (hit) {  // bring the probed words into L2 first (no bulk traffic)
          latency_kernel<<<nsm, 128, 210 * 1024>>>(hbm, hbm_bytes, probe, 0, salt, cyc, sink);
        } else {    // evict: stream 256 MB through L2
          cudaMemsetAsync(hbm, rep, size_t(256) << 20);
        }
        latency_kernel<<<nsm, 128, 210 * 1024>>>(hbm, hbm_bytes, probe, kb, salt, cyc, sink);
        cudaMemcpy(h.data(), cyc, 2 * nsm * 4, cudaMemcpyDeviceToHost);
        std::vector<unsigned> a(h.begin(), h.begin() + nsm), b(h.begin() + nsm, h.end());
        std::sort(a.begin(), a.end());
        std::sort(b.begin(), b.end());
        med.push_back(a[nsm / 2]);
        med.push_back(a[nsm - 1]);
        done.push_back(b[nsm / 2]);
      }
      std::sort(done.begin(), done.end());
      std::vector<unsigned> m, x;
      for (size_t i = 0; i < med.size(); i += 2) { m.push_back(med[i]); x.push_back(med[i + 1]); }
      std::sort(m.begin(), m.end());
      std::sort(x.begin(), x.end());
      printf("%8d %10s %14u %14u %16u\n", kb, hit ? "L2 hit" : "L2 miss", m[m.size() / 2], x[x.size() / 2],
             done[done.size() / 2]);
    }
  if (cudaGetLastError() != cudaSuccess) { printf("latency error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }

  printf("\n== DSMEM: each CTA pushes to the next CTA of its cluster (64 KiB x 8 reps per CTA) ==\n");
  printf("%-10s %8s %14s %14s\n", "mode", "cluster", "cyc / 64 KiB", "B / cyc / SM");
  for (int mode = 0; mode < 2; ++mode)
    for (int cs : {2, 4}) {
      if (mode == 1) continue;
      const int bytes = 64 * 1024, reps = 8;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((nsm / cs) * cs >= 16 * cs ? 16 * cs : cs);
      cfg.blockDim = dim3(128);
      cfg.dynamicSmemBytes = 210 * 1024;
      cudaLaunchAttribute at;
      at.id = cudaLaunchAttributeClusterDimension;
      at.val.clusterDim.x = cs;
      at.val.clusterDim.y = at.val.clusterDim.z = 1;
      cfg.attrs = &at;
      cfg.numAttrs = 1;
      cudaError_t e = cudaLaunchKernelEx(&cfg, dsmem_kernel, mode, bytes, reps, cyc);
      if (e != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
        printf("dsmem launch failed: %s\n", cudaGetErrorString(cudaGetLastError()));
        continue;
      }
      const int n = cfg.gridDim.x;
      cudaMemcpy(h.data(), cyc, n * 4, cudaMemcpyDeviceToHost);
      std::vector<unsigned> a(h.begin(), h.begin() + n);
      std::sort(a.begin(), a.end());
      const double per = double(a[n / 2]) / reps;
      printf("%-10s %8d %14.0f %14.1f\n", mode == 0 ? "bulk" : "st.v4", cs, per, bytes / per);
    }
  return 0;
}
