// sce_epilogues.cuh — the fused epilogues of the four GEMMs of one ensemble training step.
// Each functor is constructed per (thread, tile) by gemm_split_kernel, receives the fp32
// accumulator of its row in 32-column chunks straight from TMEM, and writes what the next GEMM
// needs — as (hi, lo) bf16 pairs — so the fp32 code tensor [M,B,n] never exists in HBM.
//
// Reference arithmetic being fused (HoagyC/sparse_coding @ 69c5ae0):
//   encode  c = clamp(x W^T + b, min=0) [masked_fill]        autoencoders/sae_ensemble.py:141-143, 356
//   decode  x^ = c W ; l_rec = mean((x^ - x)^2)               :145, :148
//   l1      alpha * mean_b sum_n |c|                          :149
//   dcode   backward of the above
#pragma once
#include "sce_gemm.cuh"

namespace sce {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Store 32 consecutive bf16 (64 B) from packed registers.
__device__ __forceinline__ void store_bf16x32(__nv_bfloat16* dst, const uint32_t (&w)[16], int ncols_valid) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (j * 8 < ncols_valid) {  // host guarantees n % 8 == 0
      uint4 v = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
      *reinterpret_cast<uint4*>(dst + j * 8) = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Warp-private staging tile: 32 rows x 64 bf16 (128 B per row, 4 KB), 16-byte chunks XOR-swizzled
// by (row & 7) so that both the row-per-thread access (epilogue math) and the 8-lanes-per-row access
// (coalesced 128-byte global lines) hit the 4-wavefront minimum of a 512-byte warp access.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4* stage_slot(uint8_t* buf, int row, int chunk) {
  return reinterpret_cast<uint4*>(buf + row * 128 + ((chunk ^ (row & 7)) << 4));
}
// thread `lane` owns row `lane`: write its 64 bf16 (32 packed words)
__device__ __forceinline__ void stage_put_row(uint8_t* buf, int lane, const uint32_t (&w)[32]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) *stage_slot(buf, lane, q) = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
}
__device__ __forceinline__ void stage_get_row(uint8_t* buf, int lane, uint32_t (&w)[32]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const uint4 v = *stage_slot(buf, lane, q);
    w[4 * q] = v.x;
    w[4 * q + 1] = v.y;
    w[4 * q + 2] = v.z;
    w[4 * q + 3] = v.w;
  }
}
// staging tile -> global rows [row0, row0+32) x cols [col, col+64): 8 lanes write one 128-byte line
__device__ __forceinline__ void stage_flush(uint8_t* buf, int lane, __nv_bfloat16* gbase /*row0, col*/, int ld,
                                            int rows_valid, int cols_valid) {
  const int q = lane & 7;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = it * 4 + (lane >> 3);
    if (r < rows_valid && q * 8 < cols_valid)
      *reinterpret_cast<uint4*>(gbase + (long long)r * ld + q * 8) = *stage_slot(buf, r, q);
  }
}
// global rows -> staging tile (same mapping)
__device__ __forceinline__ void stage_fill(uint8_t* buf, int lane, const __nv_bfloat16* gbase, int ld,
                                           int rows_valid, int cols_valid) {
  const int q = lane & 7;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = it * 4 + (lane >> 3);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < rows_valid && q * 8 < cols_valid) v = *reinterpret_cast<const uint4*>(gbase + (long long)r * ld + q * 8);
    *stage_slot(buf, r, q) = v;
  }
}

// ------------------------------------------------------------------------------------------------
// encode:  c = relu(acc + bias) -> (c_hi, c_lo);  per-tile partial sums of |c| and count(c > 0)
// A score of exactly 0 is recorded as c_hi = -0.0 so that the backward pass can reproduce
// clamp(min=0)'s gradient of 1 at z == 0 without keeping z.
// ------------------------------------------------------------------------------------------------
struct EpiEncode {
  static constexpr int kCols = 64;
  static constexpr int kWarpStageBytes = 8192;  // hi tile + lo tile
  struct Params {
    const float* bias;             // [M, n] or nullptr
    const unsigned char* mask;     // [M, n] (1 = coefficient unused) or nullptr
    __nv_bfloat16* c_hi;           // [M, B, n]
    __nv_bfloat16* c_lo;
    float* part;                   // [M][tiles_m*4][tiles_n][2]  (sum c, nnz)
    long long c_model_stride;      // batch_max*n
    int ldc;                       // n
    int tiles_m, tiles_n;
    int flag_zero;                 // 1: mark z == 0 with -0.0 (clamp semantics), 0: relu semantics
  };
  const Params& P;
  const TileCoord& T;
  int m_total, n_total;
  uint8_t* stage;
  float l1 = 0.f, nnz = 0.f;
  __device__ EpiEncode(const Params& p, const TileCoord& t, int m, int n, uint8_t* st)
      : P(p), T(t), m_total(m), n_total(n), stage(st) {}

  __device__ __forceinline__ void chunk(int c, const uint32_t (&r)[64]) {
    const int col = T.col0 + c;
    if (col >= n_total) return;  // warp-uniform
    const bool row_ok = T.row < m_total;
    uint32_t whi[32], wlo[32];
    const float* bias = P.bias ? P.bias + (long long)T.model * n_total + col : nullptr;
    const unsigned char* mask = P.mask ? P.mask + (long long)T.model * n_total + col : nullptr;
#pragma unroll
    for (int j = 0; j < 64; j += 2) {
      __nv_bfloat16 h[2], l[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const bool col_ok = col + j + u < n_total;
        float z = __uint_as_float(r[j + u]) + ((bias && col_ok) ? __ldg(bias + j + u) : 0.f);
        const bool masked = !col_ok || (mask && __ldg(mask + j + u));
        float cv = (z > 0.f && !masked) ? z : 0.f;
        split_bf16(cv, h[u], l[u]);
        if (P.flag_zero && z == 0.f && !masked) h[u] = __ushort_as_bfloat16(0x8000);
        if (row_ok) {
          l1 += cv;
          nnz += cv > 0.f ? 1.f : 0.f;
        }
      }
      whi[j >> 1] = pack_bf16(h[0], h[1]);
      wlo[j >> 1] = pack_bf16(l[0], l[1]);
    }
    const int row0 = T.m_blk * kBM + T.warp_q * 32;
    const long long off = (long long)T.model * P.c_model_stride + (long long)row0 * P.ldc + col;
    __syncwarp();  // previous chunk's flush has finished reading the staging tiles
    stage_put_row(stage, T.lane, whi);
    stage_put_row(stage + 4096, T.lane, wlo);
    __syncwarp();
    stage_flush(stage, T.lane, P.c_hi + off, P.ldc, m_total - row0, n_total - col);
    stage_flush(stage + 4096, T.lane, P.c_lo + off, P.ldc, m_total - row0, n_total - col);
  }
  __device__ __forceinline__ void finish() {
    const float a = warp_sum(l1), b = warp_sum(nnz);
    if (T.lane == 0) {
      float* o = P.part + ((((long long)T.model * P.tiles_m + T.m_blk) * 4 + T.warp_q) * P.tiles_n + T.n_blk) * 2;
      o[0] = a;
      o[1] = b;
    }
  }
};

// ------------------------------------------------------------------------------------------------
// decode:  r = acc - x;  partial sum r^2;  g = r * 2/(B d) -> (g_hi, g_lo);  optional x^ store
// ------------------------------------------------------------------------------------------------
struct EpiDecode {
  static constexpr int kCols = 32;
  static constexpr int kWarpStageBytes = 0;
  struct Params {
    const float* x;                // [B, d] (x_model_stride = 0) or [M, B, d]
    long long x_model_stride;
    __nv_bfloat16* g_hi;           // [M, B, d]
    __nv_bfloat16* g_lo;
    float* x_hat;                  // optional [M, B, d] fp32 (evaluation / parity tests)
    float* part;                   // [M][tiles_m*4][tiles_n]  (sum r^2)
    long long g_model_stride;      // batch_max*d (workspace pitch)
    long long xhat_model_stride;   // B*d (caller's tensor)
    int ld;                        // d
    int tiles_m, tiles_n;
    float gscale;                  // 2 / (B * d)
  };
  const Params& P;
  const TileCoord& T;
  int m_total, n_total;
  float sq = 0.f;
  __device__ EpiDecode(const Params& p, const TileCoord& t, int m, int n, uint8_t*) : P(p), T(t), m_total(m), n_total(n) {}

  __device__ __forceinline__ void chunk(int c, const uint32_t (&r)[32]) {
    const int col = T.col0 + c;
    if (col >= n_total || T.row >= m_total) return;
    const float* x = P.x + (long long)T.model * P.x_model_stride + (long long)T.row * P.ld + col;
    const long long off = (long long)T.model * P.g_model_stride + (long long)T.row * P.ld + col;
    uint32_t whi[16], wlo[16];
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
      const bool ok = col + j < n_total;  // d % 4 == 0
      if (ok) xv = *reinterpret_cast<const float4*>(x + j);
      const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
      __nv_bfloat16 h[4], l[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float rr = ok ? __uint_as_float(r[j + u]) - xs[u] : 0.f;
        sq += rr * rr;
        split_bf16(rr * P.gscale, h[u], l[u]);
      }
      whi[j >> 1] = pack_bf16(h[0], h[1]);
      whi[(j >> 1) + 1] = pack_bf16(h[2], h[3]);
      wlo[j >> 1] = pack_bf16(l[0], l[1]);
      wlo[(j >> 1) + 1] = pack_bf16(l[2], l[3]);
      if (P.x_hat && ok)
        *reinterpret_cast<float4*>(P.x_hat + (long long)T.model * P.xhat_model_stride + (long long)T.row * P.ld + col + j) =
            make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                        __uint_as_float(r[j + 3]));
    }
    store_bf16x32(P.g_hi + off, whi, n_total - col);
    store_bf16x32(P.g_lo + off, wlo, n_total - col);
  }
  __device__ __forceinline__ void finish() {
    const float a = warp_sum(sq);
    if (T.lane == 0)
      P.part[(((long long)T.model * P.tiles_m + T.m_blk) * 4 + T.warp_q) * P.tiles_n + T.n_blk] = a;
  }
};

// ------------------------------------------------------------------------------------------------
// dcode:  dz = (acc + (alpha/B) [c > 0]) * [z >= 0]  -> (dz_hi, dz_lo);
//         per-warp column sums of dz (32 rows) -> bias-gradient partials
// ------------------------------------------------------------------------------------------------
struct EpiDcode {
  static constexpr int kCols = 64;
  static constexpr int kWarpStageBytes = 8192;
  struct Params {
    const __nv_bfloat16* c_hi;     // [M, B, n]
    const float* l1_over_b;        // [M]: alpha_m / B
    __nv_bfloat16* dz_hi;          // [M, B, n]
    __nv_bfloat16* dz_lo;
    float* db_part;                // [M][tiles_m*4][n] or nullptr (no bias)
    long long c_model_stride;      // batch_max*n
    int ldc;                       // n
    int tiles_m;
  };
  const Params& P;
  const TileCoord& T;
  int m_total, n_total;
  uint8_t* stage;
  float aB;
  __device__ EpiDcode(const Params& p, const TileCoord& t, int m, int n, uint8_t* st)
      : P(p), T(t), m_total(m), n_total(n), stage(st) {
    aB = __ldg(P.l1_over_b + T.model);
  }

  // 32 lanes x 32 columns -> lane j holds the sum of column j (31 shuffles)
  __device__ __forceinline__ float transpose_reduce(float (&v)[32]) {
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
      const bool upper = (T.lane & half) != 0;
#pragma unroll
      for (int i = 0; i < half; ++i) {
        const float send = upper ? v[i] : v[i + half];
        const float keep = upper ? v[i + half] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
      }
    }
    return v[0];
  }

  __device__ __forceinline__ void chunk(int c, const uint32_t (&r)[64]) {
    const int col = T.col0 + c;
    if (col >= n_total) return;  // warp-uniform
    const int row0 = T.m_blk * kBM + T.warp_q * 32;
    const long long off = (long long)T.model * P.c_model_stride + (long long)row0 * P.ldc + col;
    // the code tile (for the activity pattern), fetched with full 128-byte lines
    uint32_t cw[32];
    __syncwarp();
    stage_fill(stage, T.lane, P.c_hi + off, P.ldc, m_total - row0, n_total - col);
    __syncwarp();
    stage_get_row(stage, T.lane, cw);
    float dz0[32], dz1[32];
    uint32_t whi[32], wlo[32];
#pragma unroll
    for (int j = 0; j < 64; j += 2) {
      __nv_bfloat16 h[2], l[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const uint32_t bits = (cw[j >> 1] >> (16 * u)) & 0xFFFFu;
        const bool pos = bits != 0u && !(bits & 0x8000u);   // c > 0
        const bool gate = pos || bits == 0x8000u;           // z >= 0 (z == 0 flagged as -0.0)
        const float v = gate ? __uint_as_float(r[j + u]) + (pos ? aB : 0.f) : 0.f;
        if (j + u < 32) dz0[j + u] = v; else dz1[j + u - 32] = v;
        split_bf16(v, h[u], l[u]);
      }
      whi[j >> 1] = pack_bf16(h[0], h[1]);
      wlo[j >> 1] = pack_bf16(l[0], l[1]);
    }
    __syncwarp();  // everyone has read its code row
    stage_put_row(stage, T.lane, whi);
    stage_put_row(stage + 4096, T.lane, wlo);
    __syncwarp();
    stage_flush(stage, T.lane, P.dz_hi + off, P.ldc, m_total - row0, n_total - col);
    stage_flush(stage + 4096, T.lane, P.dz_lo + off, P.ldc, m_total - row0, n_total - col);
    if (P.db_part) {
      float* o = P.db_part + (((long long)T.model * P.tiles_m + T.m_blk) * 4 + T.warp_q) * n_total + col;
      const float s0 = transpose_reduce(dz0);
      if (col + T.lane < n_total) o[T.lane] = s0;
      const float s1 = transpose_reduce(dz1);
      if (col + 32 + T.lane < n_total) o[32 + T.lane] = s1;
    }
  }
  __device__ __forceinline__ void finish() {}
};

}  // namespace sce
