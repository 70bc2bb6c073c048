// Host-side TMA descriptor helpers shared by the tcgen05 kernels (no libcuda link: the driver entry point is fetched
// through the runtime).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <mutex>

#include "common.cuh"

namespace b200moe {

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
  });
  return fn;
}

// 2-D bf16 row-major tensor [rows, cols]; box = [box_rows, box_cols (64 => one 128 B swizzle span)].
inline bool make_tmap_bf16(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                           uint32_t box_cols) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * sizeof(bf16)};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// The same row-major bf16 matrix [rows, cols] viewed as [cols / 64][rows][64]: a box of `box_kblocks` k-blocks x
// `box_rows` rows lands in shared memory as box_kblocks consecutive K-major tiles of box_rows x 128 bytes, each
// 128B-swizzled -- i.e. several pipeline stages' worth of operand with ONE TMA instruction (issuing a TMA costs
// ~130 ns of the producer thread regardless of the box size).  cols must be a multiple of 64.
inline bool make_tmap_bf16_kblocks(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                                   uint32_t box_kblocks) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc || cols % 64 != 0 || box_kblocks == 0 || box_kblocks > 256 || box_rows > 256) return false;
  cuuint64_t dims[3] = {64, rows, cols / 64};
  cuuint64_t strides[2] = {cols * sizeof(bf16), 64 * sizeof(bf16)};
  cuuint32_t box[3] = {64, box_rows, box_kblocks};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// fp32 flavour (TF32 compute): a 128-byte swizzle row holds 32 elements, so the view is [cols / 32][rows][32].
inline bool make_tmap_f32_kblocks(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                                  uint32_t box_kblocks) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc || cols % 32 != 0 || box_kblocks == 0 || box_kblocks > 256 || box_rows > 256) return false;
  cuuint64_t dims[3] = {32, rows, cols / 32};
  cuuint64_t strides[2] = {cols * sizeof(float), 32 * sizeof(float)};
  cuuint32_t box[3] = {32, box_rows, box_kblocks};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace b200moe
