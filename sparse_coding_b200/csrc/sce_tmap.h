// sce_tmap.h — host-side construction of the TMA tensor maps the GEMM kernels consume.
// The driver entry point is resolved at run time through the runtime API so that the shared
// library has no link-time dependency on libcuda (it must load, and export its symbols, on a
// machine without a driver — the CPU-side tests check exactly that).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

namespace sce {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || p == nullptr) return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// A bf16 tensor seen as [models][rows][cols] (cols contiguous), tiled in boxes of
// [1][box_rows][64] with the 128-byte swizzle. `models == 1` + coordinate 0 expresses an operand
// shared by the whole ensemble. Out-of-bounds box elements read as zero.
inline bool make_tmap_bf16_box(CUtensorMap* map, const void* base, uint64_t models, uint64_t rows,
                               uint64_t cols, uint64_t row_pitch_elems, uint64_t model_pitch_elems,
                               uint32_t box_cols, uint32_t box_rows, CUtensorMapSwizzle swizzle) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[3] = {cols, rows, models};
  cuuint64_t strides[2] = {row_pitch_elems * 2, model_pitch_elems * 2};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims,
                  strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

inline bool make_tmap_bf16(CUtensorMap* map, const void* base, uint64_t models, uint64_t rows,
                           uint64_t cols, uint64_t row_pitch_elems, uint64_t model_pitch_elems,
                           uint32_t box_rows) {
  return make_tmap_bf16_box(map, base, models, rows, cols, row_pitch_elems, model_pitch_elems, 64, box_rows,
                            CU_TENSOR_MAP_SWIZZLE_128B);
}

// Epilogue store map: boxes of [1][32 rows][32 cols] bf16 (64-byte rows, 64-byte swizzle).
inline bool make_tmap_bf16_store32(CUtensorMap* map, const void* base, uint64_t models, uint64_t rows,
                                   uint64_t cols, uint64_t model_pitch_elems) {
  return make_tmap_bf16_box(map, base, models, rows, cols, cols, model_pitch_elems, 32, 32,
                            CU_TENSOR_MAP_SWIZZLE_64B);
}

// An 8-bit tensor (the e5m2 planes of the f16f8 arithmetic) seen as [models][rows][cols], boxes of
// [1][box_rows][box_cols] bytes. K-major operand tiles: box_cols = BK with the 64-byte (BK = 64) or 32-byte
// (BK = 32) swizzle; MN-major tiles: box_cols = 128 with the 128-byte swizzle; epilogue stores: 32 x 32, 32-byte swizzle.
inline bool make_tmap_u8_box(CUtensorMap* map, const void* base, uint64_t models, uint64_t rows, uint64_t cols,
                             uint64_t row_pitch_elems, uint64_t model_pitch_elems, uint32_t box_cols,
                             uint32_t box_rows, CUtensorMapSwizzle swizzle) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[3] = {cols, rows, models};
  cuuint64_t strides[2] = {row_pitch_elems, model_pitch_elems};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// An fp32 tensor [models][rows][cols] written in boxes of [1][32 rows][32 cols] (128-byte rows, 128-byte swizzle):
// the epilogue store map of the top-k scores.
inline bool make_tmap_f32_store32(CUtensorMap* map, const void* base, uint64_t models, uint64_t rows, uint64_t cols,
                                  uint64_t model_pitch_elems) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[3] = {cols, rows, models};
  cuuint64_t strides[2] = {cols * 4, model_pitch_elems * 4};
  cuuint32_t box[3] = {32, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

}  // namespace sce
