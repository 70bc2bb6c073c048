// Micro-benchmark (GPU box only): how fast can 148 persistent CTAs stream weights from HBM into shared memory with TMA,
// as a function of the access pattern and of the bytes in flight per SM?  Sets the ceiling for the expert-FFN kernel in
// the weight-bound regime.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I3m-asr-inference_b200/csrc tools/stream_bench.cu -o gpurun_out/stream_bench
//   gpurun_out/stream_bench
//
// Patterns:
//   tile3d  the FFN kernel's A-operand load: [K/64][rows][64] view, box = 64 x 128 rows x kps k-blocks, 128B swizzle,
//           tiles walked like the kernel walks them (one 128-row block after the other, all its k-blocks in order)
//   bulk    cp.async.bulk of contiguous stage-sized chunks (no tensor map): the friendliest pattern there is
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ptx.cuh"
#include "tma_host.cuh"

using namespace b200moe;

__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :
               : "r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// mode 0: tile3d, mode 1: bulk.  cols = K of the weight matrix (512 or 1024), kps k-blocks of 64 per stage.
__global__ void __launch_bounds__(128, 1)
stream_kernel(const __grid_constant__ CUtensorMap tm, const uint8_t* base, size_t total_bytes, int mode, int cols, int kps,
              int stages, int rows_total, unsigned long long* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = ptx::smem_u32(smem);
  const uint32_t stage_bytes = 128u * 128u * kps;  // 128 rows x 128 B x kps
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
  const int kb_per_row_block = cols / 64 / kps;                 // stages per 128-row block
  const long long n_blocks = rows_total / 128;                   // 128-row blocks in the matrix
  const long long n_units = n_blocks * kb_per_row_block;         // stage-sized units
  if (threadIdx.x == 0) {
    int st = 0;
    uint32_t ph = 0;
    // each CTA takes whole 128-row blocks (contiguous 128 x cols x 2 bytes), round-robin
    for (long long rb = blockIdx.x; rb < n_blocks; rb += gridDim.x) {
      for (int j = 0; j < kb_per_row_block; ++j) {
        ptx::mbar_wait(empty(st), ph ^ 1u);
        ptx::mbar_arrive_expect_tx(full(st), stage_bytes);
        if (mode == 0) {
          ptx::tma_load_3d(sbase + st * stage_bytes, &tm, full(st), 0, static_cast<int>(rb * 128), j * kps,
                           ptx::kEvictFirst);
        } else {
          const size_t off = (static_cast<size_t>(rb) * kb_per_row_block + j) * stage_bytes;
          bulk_load(sbase + st * stage_bytes, base + off, stage_bytes, full(st));
        }
        if (++st == stages) {
          st = 0;
          ph ^= 1u;
        }
      }
    }
  } else if (threadIdx.x == 32) {
    int st = 0;
    uint32_t ph = 0;
    unsigned long long acc = 0;
    for (long long rb = blockIdx.x; rb < n_blocks; rb += gridDim.x) {
      for (int j = 0; j < kb_per_row_block; ++j) {
        ptx::mbar_wait(full(st), ph);
        acc += *reinterpret_cast<volatile unsigned*>(smem + st * stage_bytes);
        ptx::mbar_arrive(empty(st));
        if (++st == stages) {
          st = 0;
          ph ^= 1u;
        }
      }
    }
    if (acc == 0x1234567ull) sink[0] = acc;
  }
  (void)n_units;
  (void)total_bytes;
}

int main() {
  const size_t total = size_t(1) << 30;  // 1 GiB: far beyond the 126 MB L2
  uint8_t* buf = nullptr;
  unsigned long long* sink = nullptr;
  cudaMalloc(&buf, total);
  cudaMalloc(&sink, 8);
  cudaMemset(buf, 1, total);
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  int nsm = 148;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  printf("SMs %d\n", nsm);
  printf("%-8s %5s %4s %6s %9s %10s %9s\n", "mode", "cols", "kps", "stages", "KB/SM", "GB/s", "us/64MiB");
  for (int mode = 0; mode < 2; ++mode)
    for (int cols : {512, 1024})
      for (int kps : {1, 2, 4})
        for (int stages : {2, 3, 4, 6, 8, 12}) {
          const uint32_t stage_bytes = 128u * 128u * kps;
          const size_t smem = size_t(stages) * stage_bytes + 8 * 32 + 64;
          if (smem > 220 * 1024) continue;
          if (mode == 1 && cols == 1024) continue;
          const int rows_total = static_cast<int>(total / (size_t(cols) * 2));
          CUtensorMap tm;
          if (!make_tmap_bf16_kblocks(&tm, buf, rows_total, cols, 128, kps)) {
            printf("tensor map failed\n");
            return 1;
          }
          float best = 1e30f;
          for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            stream_kernel<<<nsm, 128, smem>>>(tm, buf, total, mode, cols, kps, stages, rows_total, sink);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
          }
          cudaError_t err = cudaGetLastError();
          if (err != cudaSuccess) {
            printf("error: %s\n", cudaGetErrorString(err));
            return 1;
          }
          const double gbs = double(total) / (best * 1e-3) / 1e9;
          printf("%-8s %5d %4d %6d %9.0f %10.1f %9.2f\n", mode == 0 ? "tile3d" : "bulk", cols, kps, stages,
                 stages * stage_bytes / 1024.0, gbs, 64.0 * 1048576 / (gbs * 1e9) * 1e6);
        }
  // reference point: plain device-to-device copy and a read-only reduction by ordinary loads
  {
    uint8_t* dst = nullptr;
    cudaMalloc(&dst, total);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      cudaMemcpyAsync(dst, buf, total, cudaMemcpyDeviceToDevice);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    printf("cudaMemcpy D2D: %.1f GB/s read + the same written\n", double(total) / (best * 1e-3) / 1e9);
    cudaFree(dst);
  }
  return 0;
}
