// Micro-benchmark (GPU box only): how many cycles does one tcgen05.mma.cta_group::1.kind::f16 (M = 128, K = 16, both
// operands in shared memory, 128B swizzle) take as a function of N, alone and while bulk copies stream into the other
// half of shared memory at the expert kernel's rate?  (the 128-token tiles of the weight-bound
// regime read 8 KiB of operands per instruction, the whole shared-memory bandwidth of the SM.)
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I3m-asr-inference_b200/csrc tools/mma_rate_bench.cu -o tools/bin/mma_rate_bench
#include <algorithm>
#include <cstdio>
#include <vector>

#include "ptx.cuh"

using namespace b200moe;

__global__ void __launch_bounds__(128, 1)
mma_kernel(const uint8_t* src, int n, int batches, int stream_kb_per_batch, int random_data, int pipelined, int rotate, unsigned* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = ptx::smem_u32(smem);
  const uint32_t a = sbase, b = sbase + 32768, dump = sbase + 65536;   // A: 2 k-blocks of 16 KiB, B: 2 x <= 32 KiB... (64-wide)
  const uint32_t bar = sbase + 200 * 1024, lbar = bar + 8, slot = bar + 16, bar2 = bar + 24;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar, 1);
    ptx::mbar_init(lbar, 1);
    ptx::mbar_init(bar2, 1);
    ptx::fence_mbar_init();
  }
  if (random_data) {  // operands with random mantissas and exponents near 1.0 (bf16), like real activations / weights
    unsigned x = 0x9E3779B9u * (threadIdx.x + 1 + blockIdx.x * 131);
    for (int i = threadIdx.x; i < 196608 / 4; i += 128) {
      x = x * 1664525u + 1013904223u;
      const unsigned lo = 0x3F00u | ((x >> 8) & 0x80FFu), hi = 0x3F00u | ((x >> 20) & 0x80FFu);
      reinterpret_cast<unsigned*>(smem)[i] = lo | (hi << 16);
    }
  } else {
    for (int i = threadIdx.x; i < 196608 / 4; i += 128) reinterpret_cast<unsigned*>(smem)[i] = 0;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 1) ptx::tmem_alloc<512>(slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + 200 * 1024 + 16);
  if (warp == 0 && lane == 0) {
    const uint32_t idesc = ptx::make_idesc(1u, 128, static_cast<uint32_t>(n));
    const uint64_t da = ptx::make_kmajor_sw128_desc(a), db = ptx::make_kmajor_sw128_desc(b);
    long long t_first = 0, total = 0;
    for (int it = 0; it < batches; ++it) {
      if (stream_kb_per_batch > 0) {  // bulk copies into the other half of shared memory while the MMAs run
        const uint32_t bytes = static_cast<uint32_t>(stream_kb_per_batch) * 1024u;
        ptx::mbar_arrive_expect_tx(lbar, bytes);
        for (uint32_t o = 0; o < bytes; o += 32768)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           dump + o),
                       "l"(src + (static_cast<size_t>(blockIdx.x) * batches + it) * bytes % (size_t(1) << 29) + o),
                       "r"(32768u), "r"(lbar)
                       : "memory");
      }
      const long long t0 = clock64();
      // rotate: batch `it` reads its operands from stage (it % 3) of a 3 x 64 KiB ring, like the expert kernel does
      const uint32_t rot = rotate ? static_cast<uint32_t>((it % 3) * 65536 >> 4) : 0u;
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          ptx::umma_f16_ss(tmem, da + rot + (j * 16384 >> 4) + 2u * k, db + rot + (j * 16384 >> 4) + 2u * k, idesc, 1u);
      if (!pipelined) {
        ptx::umma_commit(bar);
        ptx::mbar_wait(bar, it & 1);
      } else {  // two barriers in turn, waiting one batch behind: the queue never drains
        ptx::umma_commit((it & 1) ? bar2 : bar);
        if (it > 0) ptx::mbar_wait(((it - 1) & 1) ? bar2 : bar, ((it - 1) >> 1) & 1);
      }
      const long long t1 = clock64();
      if (it == 0) t_first = t1 - t0;
      else total += t1 - t0;
      if (stream_kb_per_batch > 0) ptx::mbar_wait(lbar, it & 1);
    }
    if (pipelined) ptx::mbar_wait(((batches - 1) & 1) ? bar2 : bar, ((batches - 1) >> 1) & 1);
    out[blockIdx.x * 2] = static_cast<unsigned>(t_first);
    out[blockIdx.x * 2 + 1] = static_cast<unsigned>(total / (batches - 1));
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem);
  }
}

int main() {
  uint8_t* src = nullptr;
  unsigned* out = nullptr;
  cudaMalloc(&src, size_t(1) << 30);
  cudaMemset(src, 0, size_t(1) << 30);
  cudaMalloc(&out, 148 * 8);
  cudaFuncSetAttribute(mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
  std::vector<unsigned> h(296);
  printf("8 MMAs (M=128, K=16 each) + commit + wait, all 148 SMs: cycles for the batch (median over SMs)\n");
  printf("%6s %14s %12s %12s %14s\n", "N", "stream KB/batch", "first batch", "later", "cyc per MMA");
  for (int n : {64, 128, 256})
   for (int rot = 0; rot < 2; ++rot)
    for (int pipe = 0; pipe < 2; ++pipe)
    for (int kb : {0}) {
      const int rnd = 1;
      mma_kernel<<<148, 128, 210 * 1024>>>(src, n, 66, kb, rnd, pipe, rot, out);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
      cudaMemcpy(h.data(), out, 296 * 4, cudaMemcpyDeviceToHost);
      std::vector<unsigned> f, l;
      for (int i = 0; i < 148; ++i) { f.push_back(h[2 * i]); l.push_back(h[2 * i + 1]); }
      std::sort(f.begin(), f.end());
      std::sort(l.begin(), l.end());
      printf("%6d %14d %12u %12u %14.1f  %s %s\n", n, kb, f[74], l[74], l[74] / 8.0, rot ? "rotating operands" : "same operands",
             pipe ? "pipelined" : "wait per batch");
    }
  return 0;
}
