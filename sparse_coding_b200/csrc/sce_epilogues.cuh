// sce_epilogues.cuh — the fused epilogues of the four GEMMs of one ensemble training step.
// Each functor is constructed per (thread, tile) by gemm_split_kernel, receives the fp32
// accumulator of its row in 32-column chunks straight from TMEM, and writes what the next GEMM
// needs — as operand planes (fp16 + two E5M2 planes, or a (hi, lo) bf16 pair; 4 bytes per element either way) — so
// the fp32 code tensor [M,B,n] never exists in HBM.
//
// Reference arithmetic being fused (HoagyC/sparse_coding @ 69c5ae0):
//   encode  c = clamp(x W^T + b, min=0) [masked_fill]        autoencoders/sae_ensemble.py:141-143, 356
//   decode  x^ = c W ; l_rec = mean((x^ - x)^2)               :145, :148
//   l1      alpha * mean_b sum_n |c|                          :149
//   dcode   backward of the above
#pragma once
#include "sce_gemm.cuh"

#ifndef SCE_EPI_PAIR
// f16f8 encode / dcode epilogues: 1 = two adjacent chunks per bulk store (stage_pair_and_store), 0 = one chunk per store.
// Measured slower (ncu, same box: encode 0.905 vs 0.841 ms, dcode 0.981 vs 0.954 ms; profiles/r02b_epilogue_writeout_experiments.txt):
// kept as a build option only.
#define SCE_EPI_PAIR 0
#endif

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
// Epilogue output staging: each epilogue warp owns two 2 KB shared-memory tiles (hi, lo) of 32 rows x
// 32 bf16 (64-byte rows, TMA 64-byte swizzle: 16-byte chunk index ^= (row >> 1) & 3). A thread writes its
// own row with four conflict-free 16-byte stores; one lane then hands the tile to the TMA engine, which
// writes full lines to HBM and clips rows/columns outside the tensor. Replaces 32-line scattered STG.128
// (the r01b profile had the encode/dcode epilogues bound by L1 line requests, ~16 k per tile).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_row32(uint8_t* tile, int lane, const uint32_t (&w)[16]) {
  const int sw = (lane >> 1) & 3;
#pragma unroll
  for (int q = 0; q < 4; ++q)
    *reinterpret_cast<uint4*>(tile + lane * 64 + ((q ^ sw) << 4)) =
        make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
}
// 8-bit plane tile: 32 rows x 32 bytes, TMA 32-byte swizzle (16-byte chunk index ^= (row >> 2) & 1)
__device__ __forceinline__ void stage_row32_u8(uint8_t* tile, int lane, const uint32_t* w /*[8]*/) {
  const int sw = (lane >> 2) & 1;
#pragma unroll
  for (int q = 0; q < 2; ++q)
    *reinterpret_cast<uint4*>(tile + lane * 32 + ((q ^ sw) << 4)) = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
}
// whole warp: wait until the previous tiles have been read out, write the new ones, launch their stores.
// bf16x3: whi / wx are the hi / lo planes. f16f8: whi is the fp16 plane, wx[0..7] the value-e5m2 plane and
// wx[8..15] the residual-e5m2 plane (maps m_lo / m_x8).
// `planes`: which of the f16f8 8-bit planes a consumer will read (bit 0: value plane, bit 1: residual plane); planes
// nobody reads are neither staged nor stored (warp-uniform). bf16x3 always writes both of its planes.
template <int ARITH>
__device__ __forceinline__ void stage_and_store(uint8_t* stage, int lane, const uint32_t (&whi)[16],
                                                const uint32_t (&wx)[16], const CUtensorMap* m_hi,
                                                const CUtensorMap* m_lo, const CUtensorMap* m_x8, int col, int row0,
                                                int model, int planes = 3) {
  if (lane == 0) tma_store_wait_read();
  __syncwarp();
  stage_row32(stage, lane, whi);
  if constexpr (ARITH == kArithF16F8) {
    if (planes & 1) stage_row32_u8(stage + 2048, lane, &wx[0]);
    if (planes & 2) stage_row32_u8(stage + 3072, lane, &wx[8]);
  } else {
    stage_row32(stage + 2048, lane, wx);
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    tma_store_3d(m_hi, stage, col, row0, model);
    if constexpr (ARITH == kArithF16F8) {
      if (planes & 1) tma_store_3d(m_lo, stage + 2048, col, row0, model);
      if (planes & 2) tma_store_3d(m_x8, stage + 3072, col, row0, model);
    } else {
      tma_store_3d(m_lo, stage + 2048, col, row0, model);
    }
    tma_store_commit();
  }
}

// ------------------------------------------------------------------------------------------------
// f16f8, two adjacent 32-column chunks per bulk store (Epi::kPairChunks): an epilogue warp stages the chunks 2q and 2q+1 of
// its rows side by side — fp16 tile of 32 rows x 128 B (128-byte swizzle), two 8-bit tiles of 32 rows x 64 B (64-byte
// swizzle), 8 KB per warp — and hands each tile to the TMA engine ONCE per pair: three bulk stores and one wait for the
// staging tile per 64 columns instead of per 32 (the write-out of these epilogues is bound by the latency of the store
// queue x requests in flight, profiles/r02b_epilogue_writeout_experiments.txt), and the fp16 plane goes out in full
// 128-byte lines. `half` = which chunk of the pair this is; `last` = no further chunk of this pair follows (second half,
// or the first half at the ragged right edge: the engine clips the columns beyond the tensor).
// ------------------------------------------------------------------------------------------------
constexpr int kPairStageBytes = 8192;
__device__ __forceinline__ void stage_pair_and_store(uint8_t* stage, int lane, int half, bool last, const uint32_t (&whi)[16],
                                                     const uint32_t (&wx)[16], const CUtensorMap* m_hi,
                                                     const CUtensorMap* m_lo, const CUtensorMap* m_x8, int pair_col, int row0,
                                                     int model, int planes) {
  if (half == 0) {
    if (lane == 0) tma_store_wait_read();   // the previous pair's stores have read the staging tiles
    __syncwarp();
  }
  {
    const int sw = lane & 7;   // 128-byte rows: 16-byte piece index ^= row & 7
#pragma unroll
    for (int q = 0; q < 4; ++q)
      *reinterpret_cast<uint4*>(stage + lane * 128 + (((4 * half + q) ^ sw) << 4)) =
          make_uint4(whi[4 * q], whi[4 * q + 1], whi[4 * q + 2], whi[4 * q + 3]);
  }
  {
    const int sw = (lane >> 1) & 3;   // 64-byte rows: piece index ^= (row >> 1) & 3
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int off = lane * 64 + (((2 * half + q) ^ sw) << 4);
      if (planes & 1) *reinterpret_cast<uint4*>(stage + 4096 + off) = make_uint4(wx[4 * q], wx[4 * q + 1], wx[4 * q + 2], wx[4 * q + 3]);
      if (planes & 2)
        *reinterpret_cast<uint4*>(stage + 6144 + off) = make_uint4(wx[8 + 4 * q], wx[9 + 4 * q], wx[10 + 4 * q], wx[11 + 4 * q]);
    }
  }
  if (last) {
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_3d(m_hi, stage, pair_col, row0, model);
      if (planes & 1) tma_store_3d(m_lo, stage + 4096, pair_col, row0, model);
      if (planes & 2) tma_store_3d(m_x8, stage + 6144, pair_col, row0, model);
      tma_store_commit();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// (hi, lo) split of two fp32 values into packed bf16x2 words: one packed conversion per pair for hi
// and one for lo (cvt.rn.bf16x2.f32), the residual formed on the fp32 pipe.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi2, uint32_t& lo2) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  hi2 = *reinterpret_cast<const uint32_t*>(&h);
  const float ha = __uint_as_float(hi2 << 16), hb = __uint_as_float(hi2 & 0xFFFF0000u);
  const __nv_bfloat162 l = __floats2bfloat162_rn(a - ha, b - hb);
  lo2 = *reinterpret_cast<const uint32_t*>(&l);
}

// Pair `i` (0..15, compile-time after unrolling; pairs are produced in increasing order) of a 32-column chunk into
// the packed plane words the stores take.
template <int ARITH>
__device__ __forceinline__ void split_pair(float a, float b, int i, uint32_t (&whi)[16], uint32_t (&wx)[16]) {
  if constexpr (ARITH == kArithF16F8) {
    uint32_t h8, l8;
    split2_f16f8(a, b, whi[i], h8, l8);
    if (i & 1) {
      wx[i >> 1] |= h8 << 16;
      wx[8 + (i >> 1)] |= l8 << 16;
    } else {
      wx[i >> 1] = h8;
      wx[8 + (i >> 1)] = l8;
    }
  } else {
    split2(a, b, whi[i], wx[i]);
  }
}

// ------------------------------------------------------------------------------------------------
// Activity masks: what the backward pass needs to know about the forward pass per coefficient is two bits —
// [c > 0] (the sparsity term and the ReLU gate) and [z == 0] (clamp(min=0) passes the gradient at exactly 0, SURVEY
// Q4). encode (and the top-k selection) write them as two planes of 32-column words laid out CHUNK-major,
// word(model, chunk, row) at ((model * n_chunks + chunk) * batch_max + row): the 32 lanes of an epilogue warp own 32
// consecutive rows, so one coalesced 128-byte request per chunk replaces the 32 scattered 64-byte reads of the code's
// 16-bit plane the dcode epilogue used to make (1 GB per step at config 2; same-box timing without those reads: -12 %
// on dcode). Bit (31 - j) of a word is column 32 * chunk + j.
// ------------------------------------------------------------------------------------------------
struct ActMask {
  uint32_t* pos;    // [M][n_chunks][batch_max]
  uint32_t* zero;   // same shape: z == 0 exactly; nullptr with relu semantics (top-k), where it would stay empty
  int n_chunks, batch_max;
  __device__ __forceinline__ long long at(int model, int chunk, int row) const {
    return ((long long)model * n_chunks + chunk) * batch_max + row;
  }
};

// ------------------------------------------------------------------------------------------------
// encode:  c = relu(acc + bias) -> (c_hi, c_lo);  per-tile partial sums of |c| and count(c > 0)
// [c > 0] and [z == 0] (clamp(min=0)'s gradient of 1 at exactly 0) go to the activity masks, so the
// backward pass needs neither z nor the code.
// ------------------------------------------------------------------------------------------------
template <int ARITH>
struct EpiEncodeT {
  static constexpr int kCols = 32;
  static constexpr bool kPairChunks = ARITH == kArithF16F8 && SCE_EPI_PAIR != 0;
  static constexpr int kWarpStageBytes = kPairChunks ? kPairStageBytes : 4096;
  struct Params {
    CUtensorMap out_hi, out_lo, out_x8;  // store maps of the code planes: [M][B][n], box 32 x 32
    const float* bias;             // [M, n] or nullptr
    const unsigned char* mask;     // [M, n] (1 = coefficient unused) or nullptr
    float* part;                   // [M][tiles_m*8][tiles_n][2]  (sum c, nnz)
    int tiles_m, tiles_n;
    int flag_zero;                 // 1: record z == 0 (clamp semantics), 0: relu semantics
    ActMask act;                   // activity masks for the backward pass
  };
  const Params& P;
  const TileCoord& T;
  int m_total, n_total;
  uint8_t* stage;
  float l1 = 0.f;
  int nnz = 0;
  __device__ EpiEncodeT(const Params& p, const TileCoord& t, int m, int n, uint8_t* st)
      : P(p), T(t), m_total(m), n_total(n), stage(st) {}

  __device__ __forceinline__ void chunk(int c, const uint32_t (&r)[32]) {
    const int col = T.col0 + c;
    if (col >= n_total) return;  // warp-uniform
    const bool row_ok = T.row < m_total;
    uint32_t whi[16], wlo[16];
    const float* bias = P.bias ? P.bias + (long long)T.model * n_total + col : nullptr;
    float ls = 0.f;
    uint32_t pos = 0, zero = 0;   // activity-mask words of this row and chunk: bit (31 - j) is column j
    if (col + 32 <= n_total && !P.mask && bias) {
      // fast path: whole chunk in range, no coefficient mask; bias fetched as 8 uniform float4.
      // Lean on purpose (this GEMM is bound by the SM's data paths, not by the tensor pipe): relu as max, the sign
      // bits shifted into one word, and ONE tracker — the smallest |z| — for the rare exact zeros, resolved after the loop
      float zmin = 1.f;
      uint32_t neg = 0;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + j));
        const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int u = 0; u < 4; u += 2) {
          const float z0 = __uint_as_float(r[j + u]) + bb[u], z1 = __uint_as_float(r[j + u + 1]) + bb[u + 1];
          zmin = fminf(zmin, fminf(fabsf(z0), fabsf(z1)));
          neg = __funnelshift_l(__float_as_uint(z1), __funnelshift_l(__float_as_uint(z0), neg, 1), 1);
          const float c0 = fmaxf(z0, 0.f), c1 = fmaxf(z1, 0.f);
          split_pair<ARITH>(c0, c1, (j + u) >> 1, whi, wlo);
          ls += c0 + c1;
        }
      }
      pos = ~neg;  // no zero among the 32 scores: positive <=> sign bit clear
      if (zmin == 0.f) {  // some score is exactly +-0: not positive; recorded for the clamp semantics
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float z = __uint_as_float(r[j]) + __ldg(bias + j);
          if (z == 0.f) {
            pos &= ~(0x80000000u >> j);
            if (P.flag_zero) zero |= 0x80000000u >> j;
          }
        }
      }
    } else {
      const unsigned char* mask = P.mask ? P.mask + (long long)T.model * n_total + col : nullptr;
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        float cv[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const bool col_ok = col + j + u < n_total;
          const float z = __uint_as_float(r[j + u]) + ((bias && col_ok) ? __ldg(bias + j + u) : 0.f);
          const bool masked = !col_ok || (mask && __ldg(mask + j + u));
          cv[u] = (z > 0.f && !masked) ? z : 0.f;
          if (cv[u] > 0.f) pos |= 0x80000000u >> (j + u);
          if (P.flag_zero && z == 0.f && !masked) zero |= 0x80000000u >> (j + u);
          ls += cv[u];
        }
        split_pair<ARITH>(cv[0], cv[1], j >> 1, whi, wlo);
      }
    }
    if (row_ok) {
      l1 += ls;
      nnz += __popc(pos);
      const long long w = P.act.at(T.model, col >> 5, T.row);   // 32 lanes = 32 consecutive rows: coalesced
      P.act.pos[w] = pos;
      if (P.act.zero) P.act.zero[w] = zero;
    }
    if constexpr (kPairChunks) {
      const int half = (c >> 5) & 1;
      stage_pair_and_store(stage, T.lane, half, half == 1 || col + 32 >= n_total, whi, wlo, &P.out_hi, &P.out_lo, &P.out_x8,
                           col - 32 * half, T.m_blk * kBM + T.warp_q * 32, T.model, 3);
    } else {
      stage_and_store<ARITH>(stage, T.lane, whi, wlo, &P.out_hi, &P.out_lo, &P.out_x8, col,
                             T.m_blk * kBM + T.warp_q * 32, T.model);
    }
  }
  __device__ __forceinline__ void finish() {
    if (T.lane == 0) tma_store_wait_read();  // the staging tiles must outlive their bulk stores
    const float a = warp_sum(l1), b = warp_sum(float(nnz));
    if (T.lane == 0 && T.m_blk * kBM < m_total) {  // (a CTA pair's second half may lie wholly past the batch)
      float* o = P.part +
                 ((((long long)T.model * P.tiles_m + T.m_blk) * 8 + T.grp * 4 + T.warp_q) * P.tiles_n + T.n_blk) * 2;
      o[0] = a;
      o[1] = b;
    }
  }
};

// ------------------------------------------------------------------------------------------------
// decode:  r = acc - x;  partial sum r^2;  g = r * gscale -> planes of g;  optional x^ store
// bf16x3: gscale = 2/(B d) (g is the loss gradient). f16f8: gscale = 1 — the fp16 plane could not hold 2r/(Bd)
// (~1e-7), so the backward pass runs on the residual itself and its consumers carry the factor (dcode adds
// alpha d/2 instead of alpha/B; the weight- and bias-gradient outputs are multiplied by 2/(B d)).
// ------------------------------------------------------------------------------------------------
template <int ARITH>
struct EpiDecodeT {
  static constexpr int kCols = 32;
  static constexpr int kWarpStageBytes = 0;
  struct Params {
    const float* x;                // [B, d] (x_model_stride = 0) or [M, B, d]
    long long x_model_stride;
    uint16_t* g_hi;                // [M, B, d] 16-bit plane
    uint8_t* g_lo;                 // bf16x3: lo plane (2 B / element); f16f8: value-e5m2 plane
    uint8_t* g_x8;                 // f16f8: residual-e5m2 plane
    float* x_hat;                  // optional [M, B, d] fp32 (evaluation / parity tests)
    float* part;                   // [M][tiles_m*8][tiles_n]  (sum r^2)
    long long g_model_stride;      // batch_max*d (workspace pitch)
    long long xhat_model_stride;   // B*d (caller's tensor)
    int ld;                        // d
    int tiles_m, tiles_n;
    float gscale;                  // 2 / (B * d), or 1 (f16f8)
  };
  const Params& P;
  const TileCoord& T;
  int m_total, n_total;
  float sq = 0.f;
  float4 xn[8];  // the input row's next 32 columns, fetched one chunk ahead (hides the L2 latency of x behind
                 // the tail of the main loop / the previous chunk; with SPLIT_ACC the epilogue is on the critical path)
  __device__ __forceinline__ void fetch_x(int c) {
    const int col = T.col0 + c;
    const bool row_ok = T.row < m_total;
    const float* x = P.x + (long long)T.model * P.x_model_stride + (long long)T.row * P.ld + col;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      xn[j] = (row_ok && col + 4 * j < n_total) ? __ldg(reinterpret_cast<const float4*>(x + 4 * j))
                                               : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __device__ EpiDecodeT(const Params& p, const TileCoord& t, int m, int n, uint8_t*) : P(p), T(t), m_total(m), n_total(n) {
    fetch_x(T.grp * 32);
  }

  __device__ __forceinline__ void chunk(int c, const uint32_t (&r)[32]) {
    const int col = T.col0 + c;
    float4 xc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) xc[j] = xn[j];
    fetch_x(c + 64);  // this warp's next chunk (harmless past the tile: predicated on n_total, unused)
    if (col >= n_total || T.row >= m_total) return;
    const long long off = (long long)T.model * P.g_model_stride + (long long)T.row * P.ld + col;
    uint32_t whi[16], wlo[16];
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 xv = xc[j >> 2];
      const bool ok = col + j < n_total;  // d % 4 == 0
      const float r0 = ok ? __uint_as_float(r[j]) - xv.x : 0.f, r1 = ok ? __uint_as_float(r[j + 1]) - xv.y : 0.f;
      const float r2 = ok ? __uint_as_float(r[j + 2]) - xv.z : 0.f, r3 = ok ? __uint_as_float(r[j + 3]) - xv.w : 0.f;
      sq += r0 * r0 + r1 * r1 + r2 * r2 + r3 * r3;
      split_pair<ARITH>(r0 * P.gscale, r1 * P.gscale, j >> 1, whi, wlo);
      split_pair<ARITH>(r2 * P.gscale, r3 * P.gscale, (j >> 1) + 1, whi, wlo);
      if (P.x_hat && ok)
        *reinterpret_cast<float4*>(P.x_hat + (long long)T.model * P.xhat_model_stride + (long long)T.row * P.ld + col + j) =
            make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                        __uint_as_float(r[j + 3]));
    }
    store_bf16x32(reinterpret_cast<__nv_bfloat16*>(P.g_hi) + off, whi, n_total - col);
    if constexpr (ARITH == kArithF16F8) {
#pragma unroll
      for (int q = 0; q < 2; ++q)
        if (q * 16 < n_total - col) {  // d % 16 == 0 in this arithmetic
          *reinterpret_cast<uint4*>(P.g_lo + off + q * 16) = make_uint4(wlo[4 * q], wlo[4 * q + 1], wlo[4 * q + 2], wlo[4 * q + 3]);
          *reinterpret_cast<uint4*>(P.g_x8 + off + q * 16) = make_uint4(wlo[8 + 4 * q], wlo[9 + 4 * q], wlo[10 + 4 * q], wlo[11 + 4 * q]);
        }
    } else {
      store_bf16x32(reinterpret_cast<__nv_bfloat16*>(P.g_lo) + off, wlo, n_total - col);
    }
  }
  __device__ __forceinline__ void finish() {
    const float a = warp_sum(sq);
    if (T.lane == 0 && T.m_blk * kBM < m_total)
      P.part[(((long long)T.model * P.tiles_m + T.m_blk) * 8 + T.grp * 4 + T.warp_q) * P.tiles_n + T.n_blk] = a;
  }
};

// ------------------------------------------------------------------------------------------------
// dcode:  dz = (acc + (alpha/B) [c > 0]) * [z >= 0]  -> (dz_hi, dz_lo);
//         per-warp column sums of dz (32 rows) -> bias-gradient partials
// ------------------------------------------------------------------------------------------------
template <int ARITH>
struct EpiDcodeT {
  static constexpr int kCols = 32;
  static constexpr bool kPairChunks = ARITH == kArithF16F8 && SCE_EPI_PAIR != 0;
  static constexpr int kWarpStageBytes = kPairChunks ? kPairStageBytes : 4096;
  // column offset of this warp's next chunk after the one at offset c (see the epilogue loop of gemm_split_kernel)
  static __device__ __forceinline__ int next_chunk(int c) { return kPairChunks ? (((c >> 5) & 1) ? c + 96 : c + 32) : c + 64; }
  struct Params {
    CUtensorMap out_hi, out_lo, out_x8;  // store maps of the dz planes: [M][B][n], box 32 x 32
    ActMask act;                   // [c > 0] / [z == 0] written by encode (or the top-k selection)
    const float* l1_over_b;        // [M]: alpha_m / B (f16f8: alpha_m d / 2, see EpiDecodeT)
    float* db_part;                // [M][tiles_m*4][n] or nullptr (no bias)
    int tiles_m;
    // f16f8: the 8-bit planes of dz the weight-gradient GEMM will read. dz meets x there (dz^T x): its value plane
    // multiplies x's RESIDUAL plane, which is all zeros for fp16-exact activations (*x_res_flag == 0: that cross term
    // is skipped, GemmParams::b_res_flag) — and with single-pass backward GEMMs neither 8-bit plane is read at all.
    const uint32_t* x_res_flag;    // device flag written by the batch split, or nullptr (unknown: write the plane)
    int planes;                    // planes the consumer reads at most (3, or 0 with single-pass backward)
  };
  const Params& P;
  const TileCoord& T;
  int m_total, n_total;
  uint8_t* stage;
  float aB;
  int planes;              // 8-bit planes of dz to write (see Params)
  uint32_t pos_n, zero_n;  // the mask words of this warp's next chunk, fetched one chunk ahead
  __device__ __forceinline__ void fetch_mask(int c) {
    const int col = T.col0 + c;
    pos_n = zero_n = 0u;
    if (T.row < m_total && col < n_total) {
      const long long w = P.act.at(T.model, col >> 5, T.row);
      pos_n = __ldg(P.act.pos + w);
      if (P.act.zero) zero_n = __ldg(P.act.zero + w);
    }
  }
  __device__ EpiDcodeT(const Params& p, const TileCoord& t, int m, int n, uint8_t* st)
      : P(p), T(t), m_total(m), n_total(n), stage(st) {
    aB = __ldg(P.l1_over_b + T.model);
    planes = P.planes;
    if (P.x_res_flag && __ldg(P.x_res_flag) == 0u) planes &= ~1;
    fetch_mask(kPairChunks ? T.grp * 64 : T.grp * 32);
  }

  __device__ __forceinline__ void chunk(int c, const uint32_t (&r)[32]) {
    const int col = T.col0 + c;
    const uint32_t pos = pos_n, zero = zero_n;
    fetch_mask(next_chunk(c));  // this warp's next chunk
    if (col >= n_total) return;  // warp-uniform
    float dz[32];
    uint32_t whi[16], wlo[16];
    if (zero == 0u) {   // the common case: gradient passes exactly where the coefficient is active
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float v0 = (pos & (0x80000000u >> j)) ? __uint_as_float(r[j]) + aB : 0.f;
        const float v1 = (pos & (0x40000000u >> j)) ? __uint_as_float(r[j + 1]) + aB : 0.f;
        dz[j] = v0;
        dz[j + 1] = v1;
        split_pair<ARITH>(v0, v1, j >> 1, whi, wlo);
      }
    } else {            // some z == 0: clamp passes the reconstruction gradient there, without the sparsity term
      const uint32_t gate = pos | zero;
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float v0 = (gate & (0x80000000u >> j)) ? __uint_as_float(r[j]) + ((pos & (0x80000000u >> j)) ? aB : 0.f) : 0.f;
        const float v1 = (gate & (0x40000000u >> j)) ? __uint_as_float(r[j + 1]) + ((pos & (0x40000000u >> j)) ? aB : 0.f) : 0.f;
        dz[j] = v0;
        dz[j + 1] = v1;
        split_pair<ARITH>(v0, v1, j >> 1, whi, wlo);
      }
    }
    if constexpr (kPairChunks) {
      const int half = (c >> 5) & 1;
      stage_pair_and_store(stage, T.lane, half, half == 1 || col + 32 >= n_total, whi, wlo, &P.out_hi, &P.out_lo, &P.out_x8,
                           col - 32 * half, T.m_blk * kBM + T.warp_q * 32, T.model, planes);
    } else {
      stage_and_store<ARITH>(stage, T.lane, whi, wlo, &P.out_hi, &P.out_lo, &P.out_x8, col,
                             T.m_blk * kBM + T.warp_q * 32, T.model, planes);
    }
    if (P.db_part && T.m_blk * kBM < m_total) {  // warp-uniform
      // transpose-reduce: 32 lanes x 32 columns -> lane j holds the sum of column j (31 shuffles)
#pragma unroll
      for (int half = 16; half >= 1; half >>= 1) {
        const bool upper = (T.lane & half) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
          const float send = upper ? dz[i] : dz[i + half];
          const float keep = upper ? dz[i + half] : dz[i];
          dz[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
      }
      if (col + T.lane < n_total)
        P.db_part[(((long long)T.model * P.tiles_m + T.m_blk) * 4 + T.warp_q) * n_total + col + T.lane] = dz[0];
    }
  }
  __device__ __forceinline__ void finish() {
    if (T.lane == 0) tma_store_wait_read();
  }
};

// ------------------------------------------------------------------------------------------------
// scores of the top-k variant: acc -> fp32 [M][B][n] through the same staging + bulk-store path as the code planes
// (each epilogue warp owns a 4 KB tile of 32 rows x 128 B, 128-byte swizzle: a thread writes its row with eight
// conflict-free 16-byte stores, one lane hands the tile to the TMA engine, which writes full lines and clips the
// ragged edges). The plain per-thread stores of EpiStoreF32 touch 32 different lines per instruction — fine for
// the small weight-gradient output, but the scores GEMM (K = d only, 4 B per element out) was bound by them.
// ------------------------------------------------------------------------------------------------
struct EpiScoresTma {
  static constexpr int kCols = 32;
  static constexpr int kWarpStageBytes = 4096;
  struct Params {
    CUtensorMap out;   // [M][B][n] fp32, box 32 x 32
    // optional: largest order-preserving key (f2key) of every 32-column chunk of every row, [M][batch_max][n_chunks] —
    // the selection then reads these (1/32 of the scores) and only the chunks that can hold one of the k largest
    uint32_t* cmax = nullptr;
    int n_chunks = 0;
    long long cmax_model_stride = 0;   // batch_max * n_chunks
  };
  const Params& P;
  const TileCoord& T;
  int m_total, n_total;
  uint8_t* stage;
  __device__ EpiScoresTma(const Params& p, const TileCoord& t, int m, int n, uint8_t* st)
      : P(p), T(t), m_total(m), n_total(n), stage(st) {}
  __device__ __forceinline__ void chunk(int c, const uint32_t (&r)[32]) {
    const int col = T.col0 + c;
    if (col >= n_total) return;   // warp-uniform
    if (T.lane == 0) tma_store_wait_read();
    __syncwarp();
    const int sw = T.lane & 7;
#pragma unroll
    for (int q = 0; q < 8; ++q)
      *reinterpret_cast<uint4*>(stage + T.lane * 128 + ((q ^ sw) << 4)) = make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
    fence_proxy_async_smem();
    __syncwarp();
    if (T.lane == 0) {
      tma_store_3d(&P.out, stage, col, T.m_blk * kBM + T.warp_q * 32, T.model);
      tma_store_commit();
    }
    if (P.cmax) {   // (kernel-uniform)
      const int valid = n_total - col;   // columns of this chunk inside the matrix (a multiple of 8)
      uint32_t mx = 0;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const uint32_t u = r[j];
        const uint32_t key = u ^ ((uint32_t)((int32_t)u >> 31) | 0x80000000u);   // == f2key
        if (valid >= 32 || j < valid) mx = max(mx, key);
      }
      if (T.row < m_total)
        P.cmax[(long long)T.model * P.cmax_model_stride + (long long)T.row * P.n_chunks + (col >> 5)] = mx;
    }
  }
  __device__ __forceinline__ void finish() {
    if (T.lane == 0) tma_store_wait_read();
  }
};

// ------------------------------------------------------------------------------------------------
// centring GEMM (FunctionalTiedSAE.center, sae_ensemble.py:126-128): acc = rot (x - trans); out = acc * scale[col], fp32
// [M][B][d] — the per-model batch every later kernel of the step reads (split into operand planes by split_rows_kernel,
// subtracted from x^ by the decode epilogue).
// ------------------------------------------------------------------------------------------------
struct EpiCenter {
  static constexpr int kCols = 32;
  static constexpr int kWarpStageBytes = 0;
  struct Params {
    float* out;                // [M][B][ld]
    long long model_stride;    // elements between models of `out`
    int ld;
    const float* col_scale;    // [M][ld]
  };
  const Params& P;
  const TileCoord& T;
  int m_total, n_total;
  __device__ EpiCenter(const Params& p, const TileCoord& t, int m, int n, uint8_t*) : P(p), T(t), m_total(m), n_total(n) {}
  __device__ __forceinline__ void chunk(int c, const uint32_t (&r)[32]) {
    if (T.row >= m_total) return;
    const int col = T.col0 + c;
    float* o = P.out + (long long)T.model * P.model_stride + (long long)T.row * P.ld + col;
    const float* sc = P.col_scale + (long long)T.model * P.ld + col;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      if (col + j < n_total) {   // n_total % 4 == 0
        const float4 s4 = __ldg(reinterpret_cast<const float4*>(sc + j));
        *reinterpret_cast<float4*>(o + j) = make_float4(__uint_as_float(r[j]) * s4.x, __uint_as_float(r[j + 1]) * s4.y,
                                                        __uint_as_float(r[j + 2]) * s4.z, __uint_as_float(r[j + 3]) * s4.w);
      }
    }
  }
  __device__ __forceinline__ void finish() {}
};

using EpiEncode = EpiEncodeT<kArithBf16x3>;
using EpiDecode = EpiDecodeT<kArithBf16x3>;
using EpiDcode = EpiDcodeT<kArithBf16x3>;

}  // namespace sce
