// Micro-benchmark (GPU box only): what paces bulk copies into ONE SM -- the bytes, the number of instructions, or the
// thread that issues them?  Every CTA streams out of an L2-resident window (so HBM is not the bound) through a ring of
// `stages` slots of `stage_bytes`; the slots are dealt round-robin to P producer threads in P different warps, one
// consumer thread frees them in order.  DESIGN.md section 4 ("what bounds the expert kernel") quotes the result.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I3m-asr-inference_b200/csrc tools/tma_issue_bench.cu -o tools/bin/tma_issue_bench
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "ptx.cuh"

using namespace b200moe;

__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :
               : "r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// split: every slot is filled by `split` instructions of stage_bytes / split (same barrier): instructions per byte
__global__ void __launch_bounds__(192, 1)
issue_kernel(const uint8_t* win, size_t win_bytes, int stage_bytes, int stages, int producers, int split,
             long long units_per_cta, unsigned long long* sink) {
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
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0 && warp < producers) {
    // producer `warp` owns slots warp, warp + P, ... of every round
    const int part = stage_bytes / split;
    for (long long u = warp; u < units_per_cta; u += producers) {
      const int st = static_cast<int>(u % stages);
      const uint32_t ph = static_cast<uint32_t>((u / stages) & 1);
      ptx::mbar_wait(empty(st), ph ^ 1u);
      ptx::mbar_arrive_expect_tx(full(st), stage_bytes);
      const size_t lin = (static_cast<size_t>(u) * gridDim.x + blockIdx.x) * stage_bytes;
      const uint8_t* src = win + lin % win_bytes;
      for (int i = 0; i < split; ++i) bulk_load(sbase + st * stage_bytes + i * part, src + i * part, part, full(st));
    }
  } else if (threadIdx.x == 160) {
    unsigned long long acc = 0;
    for (long long u = 0; u < units_per_cta; ++u) {
      const int st = static_cast<int>(u % stages);
      const uint32_t ph = static_cast<uint32_t>((u / stages) & 1);
      ptx::mbar_wait(full(st), ph);
      acc += *reinterpret_cast<volatile unsigned*>(smem + st * stage_bytes);
      ptx::mbar_arrive(empty(st));
    }
    if (acc == 0x1234567ull) sink[0] = acc;
  }
}

int main() {
  const size_t win_bytes = size_t(24) << 20;
  uint8_t* win = nullptr;
  unsigned long long* sink = nullptr;
  cudaMalloc(&win, win_bytes);
  cudaMalloc(&sink, 8);
  cudaMemset(win, 2, win_bytes);
  cudaFuncSetAttribute(issue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  int nsm = 148;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  printf("bulk copies out of a 24 MiB L2-resident window; ring of `stages` slots, P producer threads (one per warp)\n");
  printf("%6s %7s %6s %6s %3s %10s %12s %14s\n", "CTAs", "slot", "split", "slots", "P", "GB/s", "GB/s per SM", "instr/us/SM");
  for (int grid : {8, nsm})
    for (int stage_kb : {16, 32, 64})
      for (int split : {1, 2, 4})
        for (int producers : {1, 2, 4}) {
          const int stages = 192 / stage_kb;
          if (split > 1 && (stage_kb != 32 || producers != 1)) continue;
          if (producers > stages) continue;
          const int stage_bytes = stage_kb * 1024;
          const size_t smem = size_t(stages) * stage_bytes + 8 * 32 + 64;
          const long long units = static_cast<long long>((size_t(4) << 30) / stage_bytes / nsm) * (grid == nsm ? 1 : 1);
          float best = 1e30f;
          for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            issue_kernel<<<grid, 192, smem>>>(win, win_bytes, stage_bytes, stages, producers, split, units, sink);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            best = std::min(best, ms);
          }
          const cudaError_t err = cudaGetLastError();
          if (err != cudaSuccess) {
            printf("error: %s\n", cudaGetErrorString(err));
            return 1;
          }
          const double moved = double(units) * grid * stage_bytes;
          const double gbs = moved / (best * 1e-3) / 1e9;
          printf("%6d %6dK %6d %6d %3d %10.1f %12.1f %14.2f\n", grid, stage_kb, split, stages, producers, gbs, gbs / grid,
                 double(units) * split / (best * 1e3));
        }
  return 0;
}
