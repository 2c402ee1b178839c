// sce_gemm.cuh — persistent, warp-specialised, batched split-operand GEMM on tcgen05 / TMEM / TMA.
//
//   D[model][i][j] = sum_set sum_k A_set[model][i,k] * B_set[model][j,k]            (fp32 in TMEM)
//
// This is how the engine reaches the reference's true-FP32 results on the low-precision tensor
// pipes. Every fp32 operand is carried as operand planes and the product formed from partial products:
//   ARITH = bf16x3: x ~= hi + lo (two bf16 planes); hi*hi + hi*lo + lo*hi, three kind::f16 passes (~2^-16);
//   ARITH = f16f8 : x ~= h + l, h = fp16(x); h*h as one kind::f16 pass, the two cross terms as kind::f8f6f4 passes
//                   on E5M2 planes at twice the rate, rescaled inside the accumulator (see sce_ptx.cuh) — 2 pass
//                   equivalents, the default.
// `passes == 1` keeps only the 16-bit plane product. Operands may be K-major (reduction index contiguous in HBM) or
// MN-major (row/column index contiguous), so no transposed copies of activations/codes are ever written.
//
// One CTA per SM, 384 threads: warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM
// allocator, warps 4..11 = epilogue (TMEM -> registers -> fused epilogue -> HBM). Accumulators are
// double-buffered in TMEM so the epilogue of tile t overlaps the main loop of tile t+1.
// Eight epilogue warps (two per TMEM lane quarter, alternating 32-column chunks) because a single
// warp per scheduler is latency-bound: the r01a profile showed ~37 issued instructions per element
// at IPC ~0.25 holding the tensor pipe at 31 % in the encode GEMM.
#pragma once
#include <type_traits>
#include "sce_ptx.cuh"

namespace sce {

constexpr int kBM = 128;        // rows of the output tile == TMEM lanes
constexpr int kGemmThreads = 384;
constexpr int kEpiWarps = 8;
constexpr int kMaxSets = 2;

// What the epilogue functor sees for each tile.
struct TileCoord {
  int model;    // ensemble index
  int m_blk;    // tile row index
  int n_blk;    // tile column index
  int row;      // global output row owned by this thread (may be >= m_total: predicate!)
  int col0;     // first global output column of the tile
  int warp_q;   // epilogue warp quarter 0..3 (rows 32*warp_q .. +31 of the tile)
  int grp;      // epilogue warp group 0..1 (handles the 32-column chunks with chunk % 2 == grp)
  int lane;
};

template <class EpiParams>
struct GemmParams {
  // bf16x3: (hi, lo) bf16 planes. f16f8: hi = fp16 plane, lo = e5m2(x) plane, x8 = e5m2((x - fp16(x)) * 2^kLoShift) plane.
  CUtensorMap a_hi[kMaxSets], a_lo[kMaxSets], b_hi[kMaxSets], b_lo[kMaxSets];
  CUtensorMap a_x8[kMaxSets], b_x8[kMaxSets];
  // f16f8: optional device flags, one per operand: *flag == 0 says "the residual plane (x8) of this operand is all
  // zeros for this launch" (e.g. activations that are exactly fp16, the reference's chunk format). The cross term that
  // multiplies that plane is then skipped together with the loads of its two planes: 25 % fewer operand bytes and
  // 8-bit instructions for that operand pair, bit-identical results. nullptr = no such knowledge.
  const uint32_t* a_res_flag[kMaxSets];
  const uint32_t* b_res_flag[kMaxSets];
  int a_batched[kMaxSets], b_batched[kMaxSets];  // 0: operand shared by all models
  int nsets;      // number of (A,B) operand pairs accumulated into the same tile
  int k_total;    // reduction length of each pair
  int passes;     // 3: hi*hi + hi*lo + lo*hi (f16f8: fp16 hh + two fp8 cross terms), 1: hi*hi
  int n_models, m_total, n_total;
  int tiles_m, tiles_n;
  // NSUB == 2 only: the last `tail_rows` row blocks (model-major order) are processed as single-width tiles, two per
  // row block, so that the final wave of the persistent grid is not half empty (requires tiles_n == 1). 0 = none.
  int tail_rows;
  // NSUB == 2 only: issue the two MMAs of a K slice as collector::a::fill / ::lastuse (A read from shared memory once)
  int a_collector;
  EpiParams epi;
};

template <int BN, int BK, bool A_MN, bool B_MN, int STAGES, int EPI_WARP_BYTES = 0, bool CTA2 = false, int ARITH = 0,
          int NSUB = 1>
struct GemmSmem {
  static constexpr int kATile = kBM * BK * 2;   // bytes, one of hi/lo
  static constexpr int kBSub = CTA2 ? BN / 2 : BN;   // B rows of ONE sub-tile held by this CTA (a pair splits B)
  static constexpr int kBRows = NSUB * kBSub;        // NSUB sub-tiles of BN output columns share one A tile
  static constexpr int kBTile = kBRows * BK * 2;
  // bf16x3: hi and lo planes of A and B. f16f8: either the fp16 planes or the four 8-bit planes (same bytes).
  static constexpr int kStage = ARITH == 1 ? kATile + kBTile : 2 * kATile + 2 * kBTile;
  static constexpr int kBarOff = STAGES * kStage;
  static constexpr int kEpiOff = kBarOff + 1024;  // barriers + tmem ptr live in the 1 KB before (keeps 1 KB alignment)
  static constexpr int kBytes = kEpiOff + kEpiWarps * EPI_WARP_BYTES + 1024 /*align slack*/;
  static_assert(kBytes <= 232448, "exceeds the 227 KB of shared memory one CTA may use");
};

// ------------------------------------------------------------------------------------------------
// The kernel
// ------------------------------------------------------------------------------------------------
// SPLIT_ACC: keep the dominant hi*hi products and the small cross terms (hi*lo, lo*hi) in two separate
// TMEM accumulators (summed in the epilogue). The tensor core's fp32 accumulation truncates, which biases a
// result by about -2e-8 per accumulated MMA (measured: 1.3e-4 at K = 32768 with all three passes in one
// chain); the cross terms are 2^-8 of the total, so moving them out shortens the chain that matters 3x.
// Costs the second accumulator stage (tile epilogue no longer overlaps the next main loop), so it is used
// where the reduction is long and the epilogue short: the decode GEMM (K = n).
//
// CTA2: the tile is 256 x BN and is computed by a CTA pair (cluster of 2, tcgen05 cta_group::2). Each CTA
// keeps 128 accumulator rows in its own TMEM, loads its own 128 rows of A and HALF of the B tile, so a stage
// is 1/3 smaller (three K=64 stages fit instead of two) and each SM reads a third less shared memory per
// MMA. `tiles_m` then counts 256-row tiles; TileCoord::m_blk stays in 128-row units (2*tile_m + cta rank).
//
// ARITH = kArithF16F8 (see sce_ptx.cuh, "fp16 + fp8 arithmetic"): a tile makes TWO sweeps over K. Sweep 1 streams the
// 8-bit planes (a stage holds A.h8, A.l8, B.h8, B.l8 — the same bytes as A.f16 + B.f16) and issues the cross terms
// as kind::f8f6f4; sweep 2 streams the fp16 planes and issues hh as kind::f16, its first instruction rescaling the
// accumulator by 2^-kLoShift. One accumulator, so the double-buffered TMEM stages stay, and the chain of dominant
// products is as short as with SPLIT_ACC.
constexpr int kArithBf16x3 = 0, kArithF16F8 = 1;

// epilogues may declare `static constexpr bool kPairChunks = true` (see the epilogue loop of gemm_split_kernel)
template <class Epi, class = void>
struct epi_pairs_chunks : std::false_type {};
template <class Epi>
struct epi_pairs_chunks<Epi, std::void_t<decltype(Epi::kPairChunks)>> : std::bool_constant<Epi::kPairChunks> {};

//
// NSUB = 2 (f16f8 only): the output tile is 256 x (2 * BN): both BN-column halves are accumulated from ONE A tile per
// K block (two MMAs per K slice, two accumulators filling all 512 TMEM columns, so no accumulator double-buffering —
// for GEMMs whose epilogue is negligible). Shared memory then takes 96 instead of 128 KB per two tiles' worth of MMAs
// (the f16f8 main loop otherwise runs exactly at the SM's shared-memory bandwidth), and the A operand is read from
// L2 / HBM once instead of once per column half.
template <class Epi, int BN, int BK, bool A_MN, bool B_MN, int STAGES, bool SPLIT_ACC = false, bool CTA2 = false,
          int ARITH = kArithBf16x3, int NSUB = 1>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_split_kernel(const __grid_constant__ GemmParams<typename Epi::Params> p) {
  constexpr bool F8 = ARITH == kArithF16F8;
  static_assert(NSUB == 1 || (NSUB == 2 && F8 && BN == 256), "two sub-tiles: f16f8, 256 columns each");
  static_assert(!F8 || !SPLIT_ACC, "f16f8 rescales in the accumulator; no split accumulators");
  static_assert(!F8 || BK % 32 == 0, "an fp8 instruction covers K = 32");
  static_assert(BN % 64 == 0 && BN <= 256, "BN must be a multiple of 64, at most 256");
  static_assert(BK % 16 == 0 && BK <= 64, "BK in {16,32,48,64}");
  static_assert(A_MN || BK == 64 || BK == 32, "K-major A: one swizzled row per tile row, 128 B (BK=64) or 64 B (BK=32)");
  static_assert(B_MN || BK == 64 || BK == 32, "K-major B: one swizzled row per tile row, 128 B (BK=64) or 64 B (BK=32)");
  using SM = GemmSmem<BN, BK, A_MN, B_MN, STAGES, Epi::kWarpStageBytes, CTA2, ARITH, NSUB>;
  static_assert(!CTA2 || BN % 128 == 0, "a CTA pair splits B in halves of whole 64-column boxes");
  constexpr int EC = Epi::kCols;  // accumulator columns handed to the epilogue per call (32 or 64)
  static_assert(EC == 32 || EC == 64, "epilogue chunk is 32 or 64 columns");
  static_assert(BN % EC == 0, "tile width must be a multiple of the epilogue chunk");
  constexpr int kAccStages = (SPLIT_ACC || NSUB == 2) ? 1 : 2;
  constexpr uint32_t kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128
                                 : (2 * BN <= 256) ? 256 : 512;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SM::kBarOff);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cta_rank = CTA2 ? int(cluster_ctarank()) : 0;
  const int tile0 = CTA2 ? int(blockIdx.x >> 1) : int(blockIdx.x);       // first tile of this CTA (pair)
  const int tile_step = CTA2 ? int(gridDim.x >> 1) : int(gridDim.x);
  const int big_tiles = (p.n_models * p.tiles_m - (NSUB == 2 ? p.tail_rows : 0)) * p.tiles_n;
  const int num_tiles = big_tiles + (NSUB == 2 ? 2 * p.tail_rows : 0);
  // tile index -> (model, tile row, first BN-column block, number of BN-column sub-tiles)
  auto decode_tile = [&](int tile, int& model, int& tile_m, int& n0, int& nsub) {
    if (NSUB == 2 && tile >= big_tiles) {
      const int t = tile - big_tiles;
      const int rb = big_tiles + (t >> 1);   // tiles_n == 1 here: row block index == big tile index
      model = rb / p.tiles_m;
      tile_m = rb - model * p.tiles_m;
      n0 = t & 1;
      nsub = 1;
    } else {
      model = tile / (p.tiles_m * p.tiles_n);
      const int rem = tile - model * (p.tiles_m * p.tiles_n);
      tile_m = rem / p.tiles_n;
      n0 = (rem % p.tiles_n) * NSUB;
      nsub = NSUB;
    }
  };
  const int kblocks = (p.k_total + BK - 1) / BK;
  const bool three = p.passes >= 3;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nsets; ++s) {
      tma_prefetch_desc(&p.a_hi[s]);
      tma_prefetch_desc(&p.b_hi[s]);
      if (three) {
        tma_prefetch_desc(&p.a_lo[s]);
        tma_prefetch_desc(&p.b_lo[s]);
        if constexpr (F8) {
          tma_prefetch_desc(&p.a_x8[s]);
          tma_prefetch_desc(&p.b_x8[s]);
        }
      }
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], CTA2 ? 2 * kEpiWarps : kEpiWarps);  // pair: both CTAs' epilogues report to CTA 0
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    if constexpr (CTA2) {
      tmem_alloc_2cta(tmem_ptr, kTmemCols);
      tmem_relinquish_2cta();
    } else {
      tmem_alloc(tmem_ptr, kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CTA2) cluster_sync_all();  // the peer's barriers must be initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // f16f8: which cross terms each operand pair needs (see GemmParams::a_res_flag); the same for every CTA of the launch
  [[maybe_unused]] bool term_lh[kMaxSets], term_hl[kMaxSets];
  if constexpr (F8) {
#pragma unroll
    for (int s = 0; s < kMaxSets; ++s) {
      term_lh[s] = s < p.nsets && (p.a_res_flag[s] == nullptr || __ldg(p.a_res_flag[s]) != 0u);
      term_hl[s] = s < p.nsets && (p.b_res_flag[s] == nullptr || __ldg(p.b_res_flag[s]) != 0u);
    }
  }

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      const uint32_t stage_bytes_full = (three && !F8) ? uint32_t(SM::kStage) : uint32_t(SM::kATile + SM::kBTile);
      // one copy: same-CTA barrier, or (pair) the cta_group::2 form that completes on CTA 0's barrier
      auto load = [&](void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
        if constexpr (CTA2) tma_load_3d_2cta(dst, m, bar, c0, c1, c2);
        else tma_load_3d(dst, m, bar, c0, c1, c2);
      };
      constexpr int kBHalf = SM::kBSub;  // B rows (N index) of one sub-tile this CTA loads
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        int model, tile_m, n0, nsub;
        decode_tile(tile, model, tile_m, n0, nsub);
        const int m_blk = tile_m * (CTA2 ? 2 : 1) + cta_rank;
        const int b_row0 = n0 * BN + cta_rank * kBHalf;  // (+ sub * BN for the second sub-tile)
        [[maybe_unused]] const int n_blk = n0;
        // bytes of a 16-bit stage of THIS tile (a tail tile of an NSUB = 2 launch loads one sub-tile of B only)
        const uint32_t stage_bytes = F8 ? uint32_t(SM::kATile + nsub * (SM::kBSub * BK * 2)) : stage_bytes_full;
        if constexpr (F8) {
          // sweep 1: the 8-bit planes (skipped for passes == 1); sweep 2: the fp16 planes
          for (int sweep = three ? 0 : 1; sweep < 2; ++sweep)
            for (int set = 0; set < p.nsets; ++set) {
              const int am = p.a_batched[set] ? model : 0;
              const int bm = p.b_batched[set] ? model : 0;
              // cross terms of this operand pair: t_lh = A.l8 x B.h8 (needs A's residual), t_hl = A.h8 x B.l8
              const bool t_lh = term_lh[set], t_hl = term_hl[set];
              if (sweep == 0 && !t_lh && !t_hl) continue;
              const uint32_t bytes = sweep == 1 ? stage_bytes : (uint32_t(t_lh) + uint32_t(t_hl)) * (stage_bytes / 2);
              for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* st = smem + stage * SM::kStage;
                if (!CTA2 || cta_rank == 0) mbar_expect_tx(&full_bar[stage], CTA2 ? 2 * bytes : bytes);
                const int k0 = kb * BK;
                if (sweep == 1) {
                  uint8_t* sa = st;
                  uint8_t* sb = st + SM::kATile;
                  if constexpr (!A_MN) load(sa, &p.a_hi[set], &full_bar[stage], k0, m_blk * kBM, am);
                  else {
#pragma unroll
                    for (int j = 0; j < kBM / 64; ++j)
                      load(sa + j * (BK * 128), &p.a_hi[set], &full_bar[stage], m_blk * kBM + j * 64, k0, am);
                  }
                  for (int sub = 0; sub < nsub; ++sub) {
                    uint8_t* sbs = sb + sub * (kBHalf * BK * 2);
                    const int r0 = b_row0 + sub * BN;
                    if constexpr (!B_MN) load(sbs, &p.b_hi[set], &full_bar[stage], k0, r0, bm);
                    else {
#pragma unroll
                      for (int j = 0; j < kBHalf / 64; ++j)
                        load(sbs + j * (BK * 128), &p.b_hi[set], &full_bar[stage], r0 + j * 64, k0, bm);
                    }
                  }
                } else {
                  uint8_t* sa_h = st;
                  uint8_t* sa_l = st + SM::kATile / 2;
                  uint8_t* sb_h = st + SM::kATile;
                  uint8_t* sb_l = sb_h + SM::kBTile / 2;
                  if constexpr (!A_MN) {
                    if (t_hl) load(sa_h, &p.a_lo[set], &full_bar[stage], k0, m_blk * kBM, am);
                    if (t_lh) load(sa_l, &p.a_x8[set], &full_bar[stage], k0, m_blk * kBM, am);
                  } else {
                    static_assert(!F8 || !A_MN || kBM == 128, "one 128-element box per MN-major 8-bit A tile");
                    if (t_hl) load(sa_h, &p.a_lo[set], &full_bar[stage], m_blk * kBM, k0, am);
                    if (t_lh) load(sa_l, &p.a_x8[set], &full_bar[stage], m_blk * kBM, k0, am);
                  }
                  for (int sub = 0; sub < nsub; ++sub) {
                    uint8_t* sbh = sb_h + sub * (kBHalf * BK);
                    uint8_t* sbl = sb_l + sub * (kBHalf * BK);
                    const int r0 = b_row0 + sub * BN;
                    if constexpr (!B_MN) {
                      if (t_lh) load(sbh, &p.b_lo[set], &full_bar[stage], k0, r0, bm);
                      if (t_hl) load(sbl, &p.b_x8[set], &full_bar[stage], k0, r0, bm);
                    } else {
                      static_assert(!F8 || !B_MN || kBHalf % 128 == 0, "MN-major 8-bit B tiles come in 128-element boxes");
#pragma unroll
                      for (int j = 0; j < kBHalf / 128; ++j) {
                        if (t_lh) load(sbh + j * (BK * 128), &p.b_lo[set], &full_bar[stage], r0 + j * 128, k0, bm);
                        if (t_hl) load(sbl + j * (BK * 128), &p.b_x8[set], &full_bar[stage], r0 + j * 128, k0, bm);
                      }
                    }
                  }
                }
                if (++stage == STAGES) {
                  stage = 0;
                  phase ^= 1;
                }
              }
            }
        } else {
        for (int set = 0; set < p.nsets; ++set) {
          const int am = p.a_batched[set] ? model : 0;
          const int bm = p.b_batched[set] ? model : 0;
          for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa_hi = smem + stage * SM::kStage;
            uint8_t* sa_lo = sa_hi + SM::kATile;
            uint8_t* sb_hi = sa_lo + SM::kATile;
            uint8_t* sb_lo = sb_hi + SM::kBTile;
            // pair: CTA 0 announces the bytes of BOTH CTAs; the peer's copies complete on CTA 0's barrier
            if (!CTA2 || cta_rank == 0) mbar_expect_tx(&full_bar[stage], CTA2 ? 2 * stage_bytes : stage_bytes);
            const int k0 = kb * BK;
            if constexpr (!A_MN) {
              load(sa_hi, &p.a_hi[set], &full_bar[stage], k0, m_blk * kBM, am);
              if (three) load(sa_lo, &p.a_lo[set], &full_bar[stage], k0, m_blk * kBM, am);
            } else {
#pragma unroll
              for (int j = 0; j < kBM / 64; ++j) {
                load(sa_hi + j * (BK * 128), &p.a_hi[set], &full_bar[stage], m_blk * kBM + j * 64, k0, am);
                if (three) load(sa_lo + j * (BK * 128), &p.a_lo[set], &full_bar[stage], m_blk * kBM + j * 64, k0, am);
              }
            }
            if constexpr (!B_MN) {
              load(sb_hi, &p.b_hi[set], &full_bar[stage], k0, b_row0, bm);
              if (three) load(sb_lo, &p.b_lo[set], &full_bar[stage], k0, b_row0, bm);
            } else {
#pragma unroll
              for (int j = 0; j < kBHalf / 64; ++j) {
                load(sb_hi + j * (BK * 128), &p.b_hi[set], &full_bar[stage], b_row0 + j * 64, k0, bm);
                if (three) load(sb_lo + j * (BK * 128), &p.b_lo[set], &full_bar[stage], b_row0 + j * 64, k0, bm);
              }
            }
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    if (lane == 0 && cta_rank == 0) {  // in a pair only CTA 0 issues; its MMAs drive both SMs
      constexpr uint32_t idesc = make_idesc_bf16(CTA2 ? 2 * kBM : kBM, BN, A_MN, B_MN);
      auto mma = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t acc_flag) {
        if constexpr (CTA2) umma_bf16_2cta(d, a, b, idesc, acc_flag);
        else umma_bf16(d, a, b, idesc, acc_flag);
      };
      auto mma16 = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc_flag) {
        if constexpr (CTA2) umma_bf16_2cta(d, a, b, id, acc_flag);
        else umma_bf16(d, a, b, id, acc_flag);
      };
      auto commit = [&](uint64_t* bar) {
        if constexpr (CTA2) umma_commit_2cta(bar);
        else umma_commit(bar);
      };
      constexpr uint32_t a_lbo = A_MN ? BK * 128 : 16;
      constexpr uint32_t b_lbo = B_MN ? BK * 128 : 16;
      constexpr uint32_t a_kstep = A_MN ? 2048 : 32;  // bytes per K=16 slice
      constexpr uint32_t b_kstep = B_MN ? 2048 : 32;
      // MN-major tiles always use 128-byte rows; K-major tiles have BK*2-byte rows (128B or 64B swizzle)
      constexpr uint32_t a_sbo = (A_MN || BK == 64) ? 1024 : 512, a_lt = (A_MN || BK == 64) ? 2 : 4;
      constexpr uint32_t b_sbo = (B_MN || BK == 64) ? 1024 : 512, b_lt = (B_MN || BK == 64) ? 2 : 4;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        int nsub = NSUB;
        if constexpr (NSUB == 2) nsub = tile >= big_tiles ? 1 : 2;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(acc * BN);   // (NSUB == 2: one stage, sub-tile s at column s * BN)
        const uint32_t d_cross = SPLIT_ACC ? tmem_base + uint32_t(BN) : d_tmem;
        uint32_t accumulate = 0;
        if constexpr (F8) {
          constexpr uint32_t i16 = make_idesc_fmt(CTA2 ? 2 * kBM : kBM, BN, A_MN, B_MN, 0, 0);  // f16 x f16
          constexpr uint32_t i8 = make_idesc_fmt(CTA2 ? 2 * kBM : kBM, BN, A_MN, B_MN, 1, 1);   // e5m2 x e5m2
          // 8-bit planes: K-major rows are BK bytes (64-byte or 32-byte swizzle); MN-major rows hold 128 elements
          constexpr uint32_t a8_sbo = A_MN ? 1024 : (BK == 64 ? 512 : 256), a8_lt = A_MN ? 2 : (BK == 64 ? 4 : 6);
          constexpr uint32_t b8_sbo = B_MN ? 1024 : (BK == 64 ? 512 : 256), b8_lt = B_MN ? 2 : (BK == 64 ? 4 : 6);
          constexpr uint32_t a8_kstep = A_MN ? 4096 : 32, b8_kstep = B_MN ? 4096 : 32;  // bytes per K = 32 slice
          const int iters = p.nsets * kblocks;
          if (three) {
            for (int set = 0; set < p.nsets; ++set) {
            const bool t_lh = term_lh[set], t_hl = term_hl[set];
            if (!t_lh && !t_hl) continue;
            for (int kb = 0; kb < kblocks; ++kb) {
              mbar_wait(&full_bar[stage], phase);
              tc_fence_after();
              const uint32_t sa_h = smem_u32(smem + stage * SM::kStage);
              const uint32_t sa_l = sa_h + SM::kATile / 2;
              const uint32_t sb_h = sa_h + SM::kATile;
              const uint32_t sb_l = sb_h + SM::kBTile / 2;
#pragma unroll
              for (int k = 0; k < BK / 32; ++k) {
                const uint64_t ah = make_sdesc(sa_h + k * a8_kstep, a_lbo, a8_sbo, a8_lt);
                const uint64_t al = make_sdesc(sa_l + k * a8_kstep, a_lbo, a8_sbo, a8_lt);
                uint32_t acc_after = accumulate;
                if constexpr (NSUB == 2 && CTA2) {
                  if (nsub == 2 && p.a_collector) {
                    const uint32_t sub8 = uint32_t(SM::kBSub * BK);
                    const uint64_t bh0 = make_sdesc(sb_h + k * b8_kstep, b_lbo, b8_sbo, b8_lt);
                    const uint64_t bh1 = make_sdesc(sb_h + sub8 + k * b8_kstep, b_lbo, b8_sbo, b8_lt);
                    const uint64_t bl0 = make_sdesc(sb_l + k * b8_kstep, b_lbo, b8_sbo, b8_lt);
                    const uint64_t bl1 = make_sdesc(sb_l + sub8 + k * b8_kstep, b_lbo, b8_sbo, b8_lt);
                    if (t_lh) {
                      umma_f8_2cta_coll<1>(d_tmem, al, bh0, i8, accumulate);
                      umma_f8_2cta_coll<2>(d_tmem + uint32_t(BN), al, bh1, i8, accumulate);
                    }
                    if (t_hl) {
                      const uint32_t a2 = t_lh ? 1u : accumulate;
                      umma_f8_2cta_coll<1>(d_tmem, ah, bl0, i8, a2);
                      umma_f8_2cta_coll<2>(d_tmem + uint32_t(BN), ah, bl1, i8, a2);
                    }
                    accumulate = 1;
                    continue;
                  }
                }
                for (int sub = 0; sub < nsub; ++sub) {
                  const uint32_t sub8 = uint32_t(sub * (SM::kBSub * BK));   // bytes of one sub-tile's 8-bit plane
                  const uint64_t bh = make_sdesc(sb_h + sub8 + k * b8_kstep, b_lbo, b8_sbo, b8_lt);
                  const uint64_t bl = make_sdesc(sb_l + sub8 + k * b8_kstep, b_lbo, b8_sbo, b8_lt);
                  uint32_t acc_s = accumulate;
                  if (t_lh) {
                    umma_f8<CTA2>(d_tmem + uint32_t(sub * BN), al, bh, i8, acc_s);
                    acc_s = 1;
                  }
                  if (t_hl) {
                    umma_f8<CTA2>(d_tmem + uint32_t(sub * BN), ah, bl, i8, acc_s);
                    acc_s = 1;
                  }
                  acc_after = acc_s;
                }
                accumulate = acc_after;
              }
              commit(&empty_bar[stage]);
              if (++stage == STAGES) {
                stage = 0;
                phase ^= 1;
              }
            }
            }
          }
          bool rescale = accumulate != 0;  // some cross term was accumulated (at 2^kLoShift): rescale with the first hh
          for (int it = 0; it < iters; ++it) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * SM::kStage);
            const uint32_t sb = sa + SM::kATile;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t ah = make_sdesc(sa + k * a_kstep, a_lbo, a_sbo, a_lt);
              if constexpr (NSUB == 2 && CTA2) {
                if (nsub == 2 && p.a_collector) {
                  const uint64_t bh0 = make_sdesc(sb + k * b_kstep, b_lbo, b_sbo, b_lt);
                  const uint64_t bh1 = make_sdesc(sb + uint32_t(SM::kBSub * BK * 2) + k * b_kstep, b_lbo, b_sbo, b_lt);
                  if (k == 0 && rescale) {
                    umma_f16_rescale_2cta_coll<1>(d_tmem, ah, bh0, i16);
                    umma_f16_rescale_2cta_coll<2>(d_tmem + uint32_t(BN), ah, bh1, i16);
                  } else {
                    umma_f16_2cta_coll<1>(d_tmem, ah, bh0, i16, accumulate);
                    umma_f16_2cta_coll<2>(d_tmem + uint32_t(BN), ah, bh1, i16, accumulate);
                  }
                  if (k == 0) rescale = false;
                  accumulate = 1;
                  continue;
                }
              }
              for (int sub = 0; sub < nsub; ++sub) {
                const uint64_t bh = make_sdesc(sb + uint32_t(sub * (SM::kBSub * BK * 2)) + k * b_kstep, b_lbo, b_sbo, b_lt);
                if (k == 0 && rescale) umma_f16_rescale<CTA2>(d_tmem + uint32_t(sub * BN), ah, bh, i16);  // D = ah*bh + D * 2^-kLoShift
                else mma16(d_tmem + uint32_t(sub * BN), ah, bh, i16, accumulate);
              }
              if (k == 0) rescale = false;
              accumulate = 1;
            }
            commit(&empty_bar[stage]);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
        if constexpr (!F8)
        for (int it = 0; it < p.nsets * kblocks; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa_hi = smem_u32(smem + stage * SM::kStage);
          const uint32_t sa_lo = sa_hi + SM::kATile;
          const uint32_t sb_hi = sa_lo + SM::kATile;
          const uint32_t sb_lo = sb_hi + SM::kBTile;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t ah = make_sdesc(sa_hi + k * a_kstep, a_lbo, a_sbo, a_lt);
            const uint64_t bh = make_sdesc(sb_hi + k * b_kstep, b_lbo, b_sbo, b_lt);
            if (three) {
              const uint64_t al = make_sdesc(sa_lo + k * a_kstep, a_lbo, a_sbo, a_lt);
              const uint64_t bl = make_sdesc(sb_lo + k * b_kstep, b_lbo, b_sbo, b_lt);
              // small cross terms first, then the dominant hi*hi term
              mma(d_cross, al, bh, accumulate);
              mma(d_cross, ah, bl, 1);
              mma(d_tmem, ah, bh, SPLIT_ACC ? accumulate : 1u);
            } else {
              mma(d_tmem, ah, bh, accumulate);
            }
            accumulate = 1;
          }
          commit(&empty_bar[stage]);  // smem slot is free (in both CTAs of a pair) once these MMAs have read it
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        commit(&tfull_bar[acc]);  // accumulator complete -> epilogue (of both CTAs)
        if (++acc == kAccStages) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ======================= epilogue =======================
    const int wq = warp & 3;
    const int grp = (warp - 4) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = tile0; tile < num_tiles; tile += tile_step) {
      TileCoord tc;
      int tile_m, n0, nsub;
      decode_tile(tile, tc.model, tile_m, n0, nsub);
      tc.m_blk = tile_m * (CTA2 ? 2 : 1) + cta_rank;
      tc.row = tc.m_blk * kBM + wq * 32 + lane;
      tc.warp_q = wq;
      tc.grp = grp;
      tc.lane = lane;

      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int sub = 0; sub < nsub; ++sub) {
      tc.n_blk = n0 + sub;   // in units of BN columns
      tc.col0 = tc.n_blk * BN;
      Epi epi(p.epi, tc, p.m_total, p.n_total, smem + SM::kEpiOff + (grp * 4 + wq) * Epi::kWarpStageBytes);
      const uint32_t taddr = tmem_base + uint32_t((acc * NSUB + sub) * BN) + (uint32_t(wq * 32) << 16);
      constexpr int kChunks = BN / EC;
      static_assert(kChunks % 2 == 0, "the two epilogue warp groups alternate chunks");
      // chunk order of an epilogue warp group: alternating chunks (grp, grp + 2, ...) or, for epilogues that stage two
      // adjacent chunks per bulk store (Epi::kPairChunks), alternating PAIRS: (2 grp, 2 grp + 1), (2 grp + 4, 2 grp + 5), ...
      constexpr bool kPairs = epi_pairs_chunks<Epi>::value;
      static_assert(!kPairs || (EC == 32 && kChunks % 4 == 0), "paired chunks: 32-column chunks, whole pairs per group");
#pragma unroll 1
      for (int it = 0; it < kChunks / 2; ++it) {
        const int c = kPairs ? ((it >> 1) * 4 + 2 * grp + (it & 1)) : (grp + 2 * it);
        uint32_t r[EC];
        tmem_ld32(taddr + uint32_t(c * EC), *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
        if constexpr (EC == 64) tmem_ld32(taddr + uint32_t(c * EC + 32), *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
        if constexpr (SPLIT_ACC) {
          static_assert(!SPLIT_ACC || EC == 32, "split accumulators are read in 32-column chunks");
          if (three) {
            uint32_t x[32];
            tmem_ld32(taddr + uint32_t(BN + c * EC), x);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(x[i]));
          }
        }
        tmem_ld_wait();
        if (it == kChunks / 2 - 1 && sub == nsub - 1) {
          // all TMEM reads of this accumulator are done: hand it back to the MMA warp early
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (CTA2) mbar_arrive_cta0(&tempty_bar[acc]);
            else mbar_arrive(&tempty_bar[acc]);
          }
        }
        epi.chunk(c * EC, r);
      }
      epi.finish();
      }
      if (++acc == kAccStages) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CTA2) cluster_sync_all();  // neither CTA may free TMEM / exit while the pair's MMAs or signals are in flight
  if (warp == 2) {
    tc_fence_after();
    if constexpr (CTA2) tmem_dealloc_2cta(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// Epilogue: plain fp32 store  out[model][row][col] = acc   (dW tiles; self-test)
// ------------------------------------------------------------------------------------------------
struct EpiStoreF32 {
  static constexpr int kCols = 32;
  static constexpr int kWarpStageBytes = 0;
  struct Params {
    float* out;
    long long model_stride;  // elements
    int ld;                  // elements
    float scale;             // out = acc * scale (0 is read as 1: zero-initialised params keep working)
  };
  const Params& P;
  const TileCoord& T;
  int m_total, n_total;
  __device__ EpiStoreF32(const Params& p, const TileCoord& t, int m, int n, uint8_t*)
      : P(p), T(t), m_total(m), n_total(n) {}
  __device__ __forceinline__ void chunk(int c, const uint32_t (&r)[32]) {
    if (T.row >= m_total) return;
    float* o = P.out + (long long)T.model * P.model_stride + (long long)T.row * P.ld + T.col0 + c;
    const float sc = P.scale == 0.f ? 1.f : P.scale;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      if (T.col0 + c + j < n_total) {  // n_total % 4 == 0 is required by the host
        float4 v = make_float4(__uint_as_float(r[j]) * sc, __uint_as_float(r[j + 1]) * sc,
                               __uint_as_float(r[j + 2]) * sc, __uint_as_float(r[j + 3]) * sc);
        *reinterpret_cast<float4*>(o + j) = v;
      }
    }
  }
  __device__ __forceinline__ void finish() {}
};

}  // namespace sce
