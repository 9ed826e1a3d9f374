// tcgen05 / TMEM / TMA contraction kernel for sm_100a: plain GEMM, 3x3x3 implicit-GEMM Conv3d and
// kernel==stride ConvTranspose3d share one persistent, warp-specialised kernel (see include/ctunet_b200.h).
//
// One CTA per SM loops over output tiles (128 voxels x BLOCK_N channels, N fastest so that CTAs running
// concurrently share the activation rows in L2):
//   warp 0 (1 lane) : TMA producer — per K block one 5-D box of activations (128 voxels x 64 channels, shifted
//                     by the filter tap; out-of-range voxels are zero-filled by the TMA unit = conv padding) and
//                     one 2-D box of packed weights (BLOCK_N x 64) into a STAGES-deep smem ring that runs ahead
//                     across tile boundaries.
//   warp 1 (1 lane) : MMA issuer — 4 x tcgen05.mma (M128 x BLOCK_N x K16) per K block into one of TWO TMEM
//                     accumulators; tcgen05.commit frees the smem stage / publishes the accumulator.
//   warps 2..       : epilogue, 4 or 8 warps (EPI_WARPS; with 8, warps w and w+4 share a TMEM lane quarter and split
//                     the columns) — tcgen05.ld (one voxel row per thread), bias / GELU / residual, then either a
//                     swizzled smem staging tile written back with one TMA store per 64-channel slab (bf16 rows),
//                     or direct stores (fp32 rows, channel-first heads, transposed-conv scatter).  InstanceNorm
//                     partial statistics of a staged bf16 tile are taken from the STAGED tile by a column-parallel
//                     pass (each thread walks a pair of columns down a band of rows: 7 instructions per 2 elements,
//                     no shuffles), accumulated per CTA in registers and flushed with fp64 atomics when the batch
//                     item changes; the other output modes keep the warp transpose-reduce.  The epilogue of tile i
//                     overlaps the main loop of tile i+1 through the second accumulator.
#include "common.cuh"
#include "../../include/ctunet_b200.h"
#include "host_util.h"
#include <stdlib.h>
#include <type_traits>

namespace ctu {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int SLAB_BYTES = BLOCK_M * 128;  // 128 rows x 64 bf16

struct GemmParams {
  int b1, b2, b3;
  int T1, T2, T3;
  int d1, d2, d3;
  int n_tiles, total_tiles;
  int num_kb, cblocks, a_c;
  int k1, k2;
  int pad;
  // epilogue
  void* out;
  const float* bias;
  const void* residual;
  double* stats;
  int n_real, out_mode, ldc, act, res_mode, ldr;
  int convt_cout, u1, u2, u3;
  int stats_ld, out_col0;
  int tma_store;
  int fast;  // bf16 rows through the TMA store (any activation, no fp32 residual): the epilogue takes the lean chunk
  int stage_res;  // fast path, flat rows, bf16 residual-like operand: staged through the output staging tile (cp.async)
};

struct TileCoord {
  int x1, x2, x3, t4, n0;
};

__device__ __forceinline__ TileCoord decode_tile(const GemmParams& p, int tile, int bn) {
  TileCoord c;
  const int n_tile = tile % p.n_tiles;
  int m_tile = tile / p.n_tiles;
  const int t1 = m_tile % p.T1;
  m_tile /= p.T1;
  const int t2 = m_tile % p.T2;
  m_tile /= p.T2;
  const int t3 = m_tile % p.T3;
  c.t4 = m_tile / p.T3;
  c.x1 = t1 * p.b1;
  c.x2 = t2 * p.b2;
  c.x3 = t3 * p.b3;
  c.n0 = n_tile * bn;
  return c;
}

// The tiles of a persistent CTA are `step` apart (the grid size, a multiple of n_tiles): instead of eight integer divisions
// per tile and thread (decode_tile), the coordinates are advanced by the decomposed step with one carry per dimension.
struct TileWalker {
  int t1, t2, t3, t4, n0;
  int s1, s2, s3, s4;
  __device__ __forceinline__ void init(const GemmParams& p, int tile0, int step, int bn) {
    n0 = (tile0 % p.n_tiles) * bn;
    int m = tile0 / p.n_tiles;
    t1 = m % p.T1; m /= p.T1;
    t2 = m % p.T2; m /= p.T2;
    t3 = m % p.T3;
    t4 = m / p.T3;
    int sm = step / p.n_tiles;   // (step % n_tiles == 0, or the CTA has a single tile)
    s1 = sm % p.T1; sm /= p.T1;
    s2 = sm % p.T2; sm /= p.T2;
    s3 = sm % p.T3;
    s4 = sm / p.T3;
  }
  __device__ __forceinline__ void next(const GemmParams& p) {
    t1 += s1;
    if (t1 >= p.T1) { t1 -= p.T1; ++t2; }
    t2 += s2;
    if (t2 >= p.T2) { t2 -= p.T2; ++t3; }
    t3 += s3;
    if (t3 >= p.T3) { t3 -= p.T3; ++t4; }
    t4 += s4;
  }
  __device__ __forceinline__ TileCoord coord(const GemmParams& p) const {
    TileCoord c;
    c.x1 = t1 * p.b1; c.x2 = t2 * p.b2; c.x3 = t3 * p.b3; c.t4 = t4; c.n0 = n0;
    return c;
  }
};

template <int BN, int STAGES, int OUT_BUFS, int CTAS_PER_SM, int EPI_WARPS>
__global__ void __launch_bounds__(64 + 32 * EPI_WARPS, CTAS_PER_SM) umma_gemm_kernel(const __grid_constant__ CUtensorMap tmA,
                                                           const __grid_constant__ CUtensorMap tmB,
                                                           const __grid_constant__ CUtensorMap tmC,
                                                           const GemmParams p) {
  constexpr int B_STAGE_BYTES = BN * BLOCK_K * 2;
  constexpr int ACC_COLS = 2 * BN;
  constexpr int TMEM_COLS = ACC_COLS < 32 ? 32 : ACC_COLS;  // 32, 64, 128, 256 or 512
  constexpr int CH = BN < 32 ? BN : 32;                       // columns per tcgen05.ld
  constexpr int SLABS = BN / 64;                              // 64-channel staging slabs (0: no TMA store path)
  constexpr uint32_t IDESC = umma_idesc_bf16(BLOCK_M, BN);
  constexpr int EPI_THREADS = 32 * EPI_WARPS;
  constexpr int COL_SPLIT = EPI_WARPS / 4;                    // warps sharing one TMEM lane quarter
  constexpr int COLS_PER_WARP = BN / COL_SPLIT;
  constexpr int SCRATCH = (4 * BN > 2 * EPI_THREADS) ? 4 * BN : 2 * EPI_THREADS;   // float2 entries
  static_assert(EPI_WARPS == 4 || EPI_WARPS == 8, "epilogue warps");
  static_assert(COLS_PER_WARP % CH == 0, "column split");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem_a + STAGES * A_STAGE_BYTES;
  uint8_t* smem_c = smem_b + STAGES * B_STAGE_BYTES;  // [OUT_BUFS][SLABS][SLAB_BYTES], 1024-aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_c + OUT_BUFS * SLABS * SLAB_BYTES);
  uint64_t* bar_full = bars;
  uint64_t* bar_empty = bars + STAGES;
  uint64_t* bar_tfull = bars + 2 * STAGES;
  uint64_t* bar_tempty = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  float2* stat_scratch = reinterpret_cast<float2*>((reinterpret_cast<uintptr_t>(tmem_slot + 2) + 15) & ~uintptr_t(15));  // [SCRATCH]
  float* bias_s = reinterpret_cast<float*>(stat_scratch + SCRATCH);  // [BN]

  pdl_trigger();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.tma_store) tma_prefetch_desc(&tmC);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bar_tfull[s]), 1);
      mbar_init(smem_u32(&bar_tempty[s]), EPI_WARPS);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // everything above (barriers, TMEM, descriptor prefetch) overlapped the previous kernel's tail

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t it = 0;
      TileWalker tw;
      tw.init(p, blockIdx.x, gridDim.x, BN);
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, tw.next(p)) {
        const TileCoord tc = tw.coord(p);
        for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1);
          const uint32_t full = smem_u32(&bar_full[s]);
          mbar_expect_tx(full, A_STAGE_BYTES + B_STAGE_BYTES);
          const int tap = kb / p.cblocks;
          const int cb = kb - tap * p.cblocks;
          const int f1 = tap % p.k1;
          const int f2 = (tap / p.k1) % p.k2;
          const int f3 = tap / (p.k1 * p.k2);
          tma_load_5d(smem_u32(smem_a + s * A_STAGE_BYTES), &tmA, full, cb * BLOCK_K, tc.x1 + f1 - p.pad,
                      tc.x2 + f2 - p.pad, tc.x3 + f3 - p.pad, tc.t4);
          tma_load_2d(smem_u32(smem_b + s * B_STAGE_BYTES), &tmB, full, tap * p.a_c + cb * BLOCK_K, tc.n0);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      uint32_t it = 0;
      int lt = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++lt) {
        const int slot = lt & 1;
        const uint32_t uph = (lt >> 1) & 1;
        mbar_wait(smem_u32(&bar_tempty[slot]), uph ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)(slot * BN);
        for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(smem_u32(&bar_full[s]), ph);
          tc_fence_after();
          const uint64_t da = umma_desc_k_sw128(smem_u32(smem_a + s * A_STAGE_BYTES));
          const uint64_t db = umma_desc_k_sw128(smem_u32(smem_b + s * B_STAGE_BYTES));
          // four K16 steps: +32 bytes inside the 128-byte swizzled row = +2 in the (addr >> 4) field (one asm block, one
          // predicate: the issuing thread is on the critical path of the K-heavy shapes)
          umma_bf16_k4(acc, da, db, IDESC, kb != 0 ? 1u : 0u);
          umma_commit(smem_u32(&bar_empty[s]));
        }
        umma_commit(smem_u32(&bar_tfull[slot]));
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2.., EPI_THREADS threads)
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int r = q * 32 + lane;
    const int e = threadIdx.x - 64;
    const int col_lo = ((warp - 2) >> 2) * COLS_PER_WARP;  // this warp's share of the tile's columns
    const int i1 = r % p.b1;
    const int i2 = (r / p.b1) % p.b2;
    const int i3 = r / (p.b1 * p.b2);
    const bool use_tma = (SLABS > 0) && (p.tma_store != 0);
    const bool sync_tiles = use_tma || (p.stats != nullptr);
    const long long o1 = (long long)p.d1 * p.u1, o2 = (long long)p.d2 * p.u2, o3 = (long long)p.d3 * p.u3;

    constexpr int STAT_PER_THREAD = (BN + EPI_THREADS - 1) / EPI_THREADS;
    float acc_s[STAT_PER_THREAD], acc_q[STAT_PER_THREAD];
#pragma unroll
    for (int k = 0; k < STAT_PER_THREAD; ++k) acc_s[k] = acc_q[k] = 0.f;
    int stat_batch = -1;
    const int n0_cta = (blockIdx.x % p.n_tiles) * BN;
    if (p.bias != nullptr) {  // the grid is a multiple of n_tiles: this CTA only ever sees the N tile n0_cta
      for (int c = e; c < BN; c += EPI_THREADS) bias_s[c] = (n0_cta + c < p.n_real) ? __ldg(p.bias + n0_cta + c) : 0.f;
      named_bar_sync(1, EPI_THREADS);
    }
    auto flush_stats = [&](int b) {
      if (b < 0) return;
#pragma unroll
      for (int k = 0; k < STAT_PER_THREAD; ++k) {
        const int c = e + EPI_THREADS * k;
        if (c < BN && n0_cta + c < p.n_real) {
          double* dst = p.stats + ((long long)b * p.stats_ld + n0_cta + c) * 2;
          atomicAdd(dst, (double)acc_s[k]);
          atomicAdd(dst + 1, (double)acc_q[k]);
        }
        acc_s[k] = acc_q[k] = 0.f;
      }
    };

    int lt = 0;
    TileWalker tw;
    tw.init(p, blockIdx.x, gridDim.x, BN);
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++lt, tw.next(p)) {
      const int slot = lt & 1;
      const uint32_t uph = (lt >> 1) & 1;
      const TileCoord tc = tw.coord(p);
      const int n0 = tc.n0;
      const int v1 = tc.x1 + i1, v2 = tc.x2 + i2, v3 = tc.x3 + i3;
      const bool valid = (v1 < p.d1) && (v2 < p.d2) && (v3 < p.d3);
      int a1 = 0, a2 = 0, a3 = 0, colbase = n0;
      if (p.convt_cout > 0) {
        const int sub = n0 / p.convt_cout;
        colbase = n0 - sub * p.convt_cout;
        a1 = sub % p.u1;
        a2 = (sub / p.u1) % p.u2;
        a3 = sub / (p.u1 * p.u2);
      }
      const long long out_row = ((tc.t4 * o3 + (long long)v3 * p.u3 + a3) * o2 + ((long long)v2 * p.u2 + a2)) * o1 +
                                ((long long)v1 * p.u1 + a1);
      uint8_t* cbuf = smem_c + (size_t)(OUT_BUFS > 0 ? (lt % (OUT_BUFS > 0 ? OUT_BUFS : 1)) : 0) * SLABS * SLAB_BYTES;

      // Residual-like operand (a gradient that has already arrived, or the pre-activation of a GELU) of a flat-row GEMM:
      // staged into the output staging tile with COALESCED 16-byte async copies while the accumulator is still being
      // computed — a thread reading its own row straight from global memory costs 32 L1 wavefronts per load instruction
      // (2,048 per 128 x 128 tile), which bounds the +res / +gelu_bwd shapes at 2.0-2.9 TB/s.
      bool stage_res = false;
      if constexpr (CH == 32 && SLABS > 0) {
        stage_res = p.stage_res != 0 && use_tma;
        if (stage_res) {
          if (e == 0) {
            if (OUT_BUFS > 1) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          }
          named_bar_sync(1, EPI_THREADS);
          constexpr int CHUNKS_PER_ROW = BN / 8;                       // 16-byte chunks of one tile row
          const __nv_bfloat16* rbase = reinterpret_cast<const __nv_bfloat16*>(p.residual) +
                                       ((long long)tc.t4 * p.d1 + tc.x1) * p.ldr + p.out_col0 + n0;
#pragma unroll
          for (int id = e; id < BLOCK_M * CHUNKS_PER_ROW; id += EPI_THREADS) {
            const int rr = id / CHUNKS_PER_ROW, c16 = id % CHUNKS_PER_ROW;
            if (tc.x1 + rr < p.d1 && n0 + c16 * 8 < p.n_real) {
              const uint32_t dst = smem_u32(cbuf) + (uint32_t)((c16 >> 3) * SLAB_BYTES + rr * 128 + (((c16 & 7) ^ (rr & 7)) << 4));
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(rbase + (long long)rr * p.ldr + c16 * 8)
                           : "memory");
            }
          }
          asm volatile("cp.async.commit_group;" ::: "memory");
        }
      }

      mbar_wait(smem_u32(&bar_tfull[slot]), uph);
      tc_fence_after();
      if (stage_res) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        named_bar_sync(1, EPI_THREADS);
      } else if (sync_tiles) {
        // the staging buffer about to be overwritten must have been read by its TMA store (OUT_BUFS tiles ago),
        // and every thread must be done with the previous tile's statistics scratch
        if (use_tma && e == 0) {
          if (OUT_BUFS > 1) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        named_bar_sync(1, EPI_THREADS);
      }
      // statistics of a staged bf16 tile come from the staged tile itself (column pass below)
      const bool col_stats = use_tma && (p.stats != nullptr);

      // One 32-column chunk of the tile.  kFull (every column of the chunk is a real output column — always, except in
      // the last N tile of the 14-logit heads) folds the per-element range checks away: the HBM-bound GEMMs of this
      // path (K = 64..512) are otherwise bound by the instruction count of this loop, not by memory.
      auto chunk = [&](int c0, auto full_tag) {
        constexpr bool kFull = decltype(full_tag)::value;
        uint32_t raw[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(slot * BN + c0);
        if constexpr (CH == 32) tmem_ld32(taddr, raw);
        else tmem_ld16(taddr, raw);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < CH; ++j) v[j] = __uint_as_float(raw[j]);
#pragma unroll
        for (int j = CH; j < 32; ++j) v[j] = 0.f;

        const int gcol = n0 + c0;  // column in the GEMM's N space (bias / stats / n_real)
        if (p.bias != nullptr) {   // staged once per CTA in shared memory, zero beyond n_real
          const float4* bp = reinterpret_cast<const float4*>(bias_s + c0);
#pragma unroll
          for (int j = 0; j < CH; j += 4) {
            const float4 bv = bp[j >> 2];
            v[j] += bv.x; v[j + 1] += bv.y; v[j + 2] += bv.z; v[j + 3] += bv.w;
          }
        }
        if (p.act == CTU_ACT_GELU) {
#pragma unroll
          for (int j = 0; j < CH; j += 2) gelu_pair<false>(v[j], v[j + 1], v[j], v[j + 1]);
        }
        const int ocol = p.out_col0 + colbase + c0;  // column inside an output row
        if (p.res_mode == CTU_RES_F32 && valid) {
          const float* rp = reinterpret_cast<const float*>(p.residual) + out_row * p.ldr + ocol;
#pragma unroll
          for (int j = 0; j < CH; j += 4) {
            if (kFull || gcol + j < p.n_real) {
              const float4 rv = *reinterpret_cast<const float4*>(rp + j);
              v[j] += rv.x; v[j + 1] += rv.y; v[j + 2] += rv.z; v[j + 3] += rv.w;
            }
          }
        } else if (p.res_mode == CTU_RES_BF16 && valid) {
          const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(p.residual) + out_row * p.ldr + ocol;
#pragma unroll
          for (int j = 0; j < CH; j += 8) {
            if (kFull || gcol + j < p.n_real) {
              const uint4 rv = *reinterpret_cast<const uint4*>(rp + j);
              float2 f;
              f = unpack_bf16x2(rv.x); v[j] += f.x; v[j + 1] += f.y;
              f = unpack_bf16x2(rv.y); v[j + 2] += f.x; v[j + 3] += f.y;
              f = unpack_bf16x2(rv.z); v[j + 4] += f.x; v[j + 5] += f.y;
              f = unpack_bf16x2(rv.w); v[j + 6] += f.x; v[j + 7] += f.y;
            }
          }
        } else if (p.res_mode == CTU_RES_GELU_BWD && valid) {
          const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(p.residual) + out_row * p.ldr + ocol;
#pragma unroll
          for (int j = 0; j < CH; j += 8) {
            if (kFull || gcol + j < p.n_real) {
              const uint4 rv = *reinterpret_cast<const uint4*>(rp + j);
              float2 f;
              const uint32_t u[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
              for (int h = 0; h < 4; ++h) {
                f = unpack_bf16x2(u[h]);
                float ga, gb;
                gelu_pair<true>(f.x, f.y, ga, gb);
                v[j + 2 * h] *= ga;
                v[j + 2 * h + 1] *= gb;
              }
            }
          }
        }

        if (p.out_mode == CTU_OUT_BF16_ROWS) {
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < CH / 2; ++j) pk[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
          if (col_stats && !valid) {  // rows outside the volume (clipped by the TMA store) must not count
#pragma unroll
            for (int j = 0; j < CH / 2; ++j) pk[j] = 0u;
          }
          if (use_tma) {
            // SWIZZLE_128B staging: row r at r*128, 16-byte chunk c at position (c ^ (r & 7)); the TMA store clips
            // rows / columns outside the output tensor
            if constexpr (CH == 32 && SLABS > 0) {
              const uint32_t base = smem_u32(cbuf) + (uint32_t)((c0 >> 6) * SLAB_BYTES + r * 128);
              const int cb = (c0 & 63) >> 3;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint32_t addr = base + (uint32_t)(((cb + i) ^ (r & 7)) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * i]), "r"(pk[4 * i + 1]),
                             "r"(pk[4 * i + 2]), "r"(pk[4 * i + 3])
                             : "memory");
              }
            }
          } else if (valid) {
            __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + out_row * p.ldc + ocol;
#pragma unroll
            for (int j = 0; j < CH; j += 8) {
              if (kFull || gcol + j < p.n_real)
                *reinterpret_cast<uint4*>(op + j) = make_uint4(pk[j / 2], pk[j / 2 + 1], pk[j / 2 + 2], pk[j / 2 + 3]);
            }
          }
          if (p.stats != nullptr && !col_stats) {
            // statistics of the values as stored (bf16-rounded), masked rows contribute zero
#pragma unroll
            for (int j = 0; j < CH / 2; ++j) {
              const float2 f = unpack_bf16x2(pk[j]);
              v[2 * j] = valid ? f.x : 0.f;
              v[2 * j + 1] = valid ? f.y : 0.f;
            }
          }
        } else if (p.out_mode == CTU_OUT_F32_ROWS) {
          if (valid) {
            float* op = reinterpret_cast<float*>(p.out) + out_row * p.ldc + ocol;
#pragma unroll
            for (int j = 0; j < CH; j += 4) {
              if (kFull || gcol + j < p.n_real) *reinterpret_cast<float4*>(op + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
          }
          if (p.stats != nullptr) {
#pragma unroll
            for (int j = 0; j < CH; ++j) v[j] = valid ? v[j] : 0.f;
          }
        } else {  // CTU_OUT_F32_CF: lanes hold consecutive voxels, so each column is a coalesced 128-byte store
          if (valid) {
            const long long S = (long long)p.d1 * p.d2 * p.d3;
            const long long s_idx = ((long long)v3 * p.d2 + v2) * p.d1 + v1;
            float* op = reinterpret_cast<float*>(p.out) + ((long long)tc.t4 * p.n_real) * S + s_idx;
#pragma unroll
            for (int j = 0; j < CH; ++j)
              if (kFull || gcol + j < p.n_real) op[(long long)(gcol + j) * S] = v[j];
          }
        }

        if (p.stats != nullptr && !col_stats) {
          float sq[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) sq[j] = v[j] * v[j];
          const float s_sum = warp_transpose_reduce(v, lane);
          const float s_sq = warp_transpose_reduce(sq, lane);
          if (lane < CH) stat_scratch[q * BN + c0 + lane] = make_float2(s_sum, s_sq);
        }
      };
      if constexpr (CH == 32 && SLABS > 0) {
        if (p.fast) {
          // bf16 rows through the staging tile, no activation, no residual (the 1x1x1 convolutions, qkv / output
          // projections, FFN up-projection of the training forward): nothing but TMEM load, optional bias, pack, store
          const bool zero_row = col_stats && !valid;   // rows outside the volume must not count in the statistics
#pragma unroll 1
          for (int c0 = col_lo; c0 < col_lo + COLS_PER_WARP; c0 += 32) {
            uint32_t raw[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(slot * BN + c0), raw);
            tmem_ld_wait();
            if (p.bias != nullptr) {
              const float4* bp = reinterpret_cast<const float4*>(bias_s + c0);
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 bv = bp[j >> 2];
                raw[j] = __float_as_uint(__uint_as_float(raw[j]) + bv.x);
                raw[j + 1] = __float_as_uint(__uint_as_float(raw[j + 1]) + bv.y);
                raw[j + 2] = __float_as_uint(__uint_as_float(raw[j + 2]) + bv.z);
                raw[j + 3] = __float_as_uint(__uint_as_float(raw[j + 3]) + bv.w);
              }
            }
            if (p.act == CTU_ACT_GELU) {
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                float ya, yb;
                gelu_pair<false>(__uint_as_float(raw[j]), __uint_as_float(raw[j + 1]), ya, yb);
                raw[j] = __float_as_uint(ya);
                raw[j + 1] = __float_as_uint(yb);
              }
            }
            if (p.res_mode != CTU_RES_NONE && valid) {
              // bf16 rows shaped like the output: a gradient that has already arrived (added), or the pre-activation of
              // the GELU whose derivative scales this input gradient
              uint4 rv[4];
              if (stage_res) {   // this thread's row of the staged tile (same swizzled positions the result goes back to)
                const uint32_t rb = smem_u32(cbuf) + (uint32_t)((c0 >> 6) * SLAB_BYTES + r * 128);
                const int cbk = (c0 & 63) >> 3;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rv[i].x), "=r"(rv[i].y), "=r"(rv[i].z), "=r"(rv[i].w)
                               : "r"(rb + (uint32_t)(((cbk + i) ^ (r & 7)) << 4)) : "memory");
              } else {
                const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.residual) +
                                                                 out_row * p.ldr + p.out_col0 + colbase + c0);
#pragma unroll
                for (int i = 0; i < 4; ++i) rv[i] = rp[i];
              }
              const bool gelu_bwd = p.res_mode == CTU_RES_GELU_BWD;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint32_t u[4] = {rv[i].x, rv[i].y, rv[i].z, rv[i].w};
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                  const float2 f = unpack_bf16x2(u[h]);
                  const int j = 8 * i + 2 * h;
                  if (gelu_bwd) {
                    float ga, gb;
                    gelu_pair<true>(f.x, f.y, ga, gb);
                    raw[j] = __float_as_uint(__uint_as_float(raw[j]) * ga);
                    raw[j + 1] = __float_as_uint(__uint_as_float(raw[j + 1]) * gb);
                  } else {
                    raw[j] = __float_as_uint(__uint_as_float(raw[j]) + f.x);
                    raw[j + 1] = __float_as_uint(__uint_as_float(raw[j + 1]) + f.y);
                  }
                }
              }
            }
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 16; ++j)
              pk[j] = zero_row ? 0u : pack_bf16x2(__uint_as_float(raw[2 * j]), __uint_as_float(raw[2 * j + 1]));
            const uint32_t base = smem_u32(cbuf) + (uint32_t)((c0 >> 6) * SLAB_BYTES + r * 128);
            const int cb = (c0 & 63) >> 3;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint32_t addr = base + (uint32_t)(((cb + i) ^ (r & 7)) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * i]), "r"(pk[4 * i + 1]),
                           "r"(pk[4 * i + 2]), "r"(pk[4 * i + 3])
                           : "memory");
            }
          }
        }
      }
      if (!(CH == 32 && SLABS > 0) || !p.fast) {
#pragma unroll 1
        for (int c0 = col_lo; c0 < col_lo + COLS_PER_WARP; c0 += CH) {
          if (n0 + c0 + CH <= p.n_real) chunk(c0, std::true_type{});
          else chunk(c0, std::false_type{});
        }
      }

      // accumulator fully read: hand it back to the MMA warp before the (slower) global write-back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[slot]));

      if (sync_tiles) {
        if (use_tma) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        named_bar_sync(2, EPI_THREADS);
        if (use_tma && e == 0) {
          if constexpr (SLABS > 0) {
#pragma unroll
            for (int sl = 0; sl < SLABS; ++sl) {
              if (n0 + sl * 64 < p.n_real) {
                // (u = 1, a = 0, colbase = n0 for everything but the up-sampling GEMMs)
                asm volatile(
                    "cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(&tmC),
                    "r"(smem_u32(cbuf + sl * SLAB_BYTES)), "r"(colbase + sl * 64), "r"(tc.x1 * p.u1 + a1), "r"(tc.x2 * p.u2 + a2),
                    "r"(tc.x3 * p.u3 + a3), "r"(tc.t4)
                    : "memory");
              }
            }
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (p.stats != nullptr) {
          // per-CTA running column sums (the host sizes the grid as a multiple of n_tiles, so a CTA always works on
          // the same N tile); flushed with fp64 atomics only when the batch item changes and at the end
          if (tc.t4 != stat_batch) {
            flush_stats(stat_batch);
            stat_batch = tc.t4;
          }
          if constexpr (SLABS > 0) {
            if (col_stats) {
              // Column pass over the staged tile (128 rows x BN bf16, SWIZZLE_128B slabs): thread -> one pair of
              // adjacent columns (one 32-bit word per row) and a band of RPT rows.  A warp reads 128 contiguous bytes of
              // one row per step (the swizzle permutes 16-byte chunks inside the row: conflict-free); the eight swizzle
              // variants of the address are precomputed, so a row costs LDS + 2 unpack + 2 FADD + 2 FFMA.
              constexpr int PAIRS = BN / 2;
              constexpr int G = EPI_THREADS / PAIRS;   // row bands
              constexpr int RPT = BLOCK_M / G;         // rows per thread (multiple of 8)
              static_assert(EPI_THREADS % PAIRS == 0 && BLOCK_M % G == 0 && RPT % 8 == 0, "column pass tiling");
              const int pr = e % PAIRS, g = e / PAIRS;
              const int col = 2 * pr;
              const uint32_t band = smem_u32(cbuf) + (uint32_t)((col >> 6) * SLAB_BYTES + g * RPT * 128 + (col & 7) * 2);
              const int chunk16 = (col & 63) >> 3;
              uint32_t addr8[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) addr8[j] = band + (uint32_t)((chunk16 ^ j) << 4);
              float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
              for (int i = 0; i < RPT; ++i) {
                uint32_t w;
                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w) : "r"(addr8[i & 7] + (uint32_t)(i * 128)) : "memory");
                const float lo = __uint_as_float(w << 16), hi = __uint_as_float(w & 0xffff0000u);
                s0 += lo; s1 += hi;
                q0 = fmaf(lo, lo, q0); q1 = fmaf(hi, hi, q1);
              }
              *reinterpret_cast<float4*>(&stat_scratch[g * BN + col]) = make_float4(s0, q0, s1, q1);
              named_bar_sync(3, EPI_THREADS);
#pragma unroll
              for (int k = 0; k < STAT_PER_THREAD; ++k) {
                const int c = e + EPI_THREADS * k;
                if (c < BN) {
#pragma unroll
                  for (int gg = 0; gg < G; ++gg) {
                    const float2 sv = stat_scratch[gg * BN + c];
                    acc_s[k] += sv.x;
                    acc_q[k] += sv.y;
                  }
                }
              }
            }
          }
          if (!col_stats) {
#pragma unroll
            for (int k = 0; k < STAT_PER_THREAD; ++k) {
              const int c = e + EPI_THREADS * k;
              if (c < BN) {
                const float2 s0 = stat_scratch[c], s1 = stat_scratch[BN + c], s2 = stat_scratch[2 * BN + c],
                             s3 = stat_scratch[3 * BN + c];
                acc_s[k] += (s0.x + s1.x) + (s2.x + s3.x);
                acc_q[k] += (s0.y + s1.y) + (s2.y + s3.y);
              }
            }
          }
        }
      }
    }
    if (p.stats != nullptr) flush_stats(stat_batch);
    if (use_tma && e == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int BN, int STAGES, int OUT_BUFS, int EPI_WARPS>
constexpr int gemm_smem_bytes() {
  constexpr int scratch = (4 * BN > 64 * EPI_WARPS) ? 4 * BN : 64 * EPI_WARPS;
  return 1024 + STAGES * (A_STAGE_BYTES + BN * BLOCK_K * 2) + OUT_BUFS * (BN / 64) * SLAB_BYTES + (2 * STAGES + 4) * 8 +
         32 + scratch * 8 + BN * 4 + 16;
}

static int sm_count() {
  static int n = [] {
    int dev = 0, v = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v;
  }();
  return n;
}

template <int BN, int STAGES, int OUT_BUFS, int CTAS_PER_SM, int EPI_WARPS = 4>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const GemmParams& p,
                       cudaStream_t stream) {
  constexpr int smem = gemm_smem_bytes<BN, STAGES, OUT_BUFS, EPI_WARPS>();
  static_assert(CTAS_PER_SM * (smem + 1024) <= 228 * 1024, "shared memory budget");
  static_assert(CTAS_PER_SM * 2 * BN <= 512, "TMEM budget");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(umma_gemm_kernel<BN, STAGES, OUT_BUFS, CTAS_PER_SM, EPI_WARPS>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  // grid: persistent CTAs, a multiple of n_tiles so that every CTA keeps one N tile (per-CTA statistics, L2 reuse)
  int cap = persistent_sms(sm_count()) * CTAS_PER_SM;
  if (cap > p.n_tiles) cap -= cap % p.n_tiles;
  const int grid = p.total_tiles < cap ? p.total_tiles : cap;
  if (grid < p.total_tiles && grid % p.n_tiles != 0) return CTU_E_UNSUPPORTED;   // persistent CTAs must keep their N tile
  const cudaError_t le = launch_pdl(umma_gemm_kernel<BN, STAGES, OUT_BUFS, CTAS_PER_SM, EPI_WARPS>, dim3(grid), dim3(64 + 32 * EPI_WARPS), smem, stream, tmA, tmB, tmC, p);
  count_launch();
  return le != cudaSuccess ? (int)le : (int)cudaGetLastError();
}

}  // namespace ctu

extern "C" int ctu_umma_gemm(const ctu_gemm_desc* d, void* stream_) {
  using namespace ctu;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!d || !d->a || !d->w || !d->out) return CTU_E_BADARG;
  if (d->b1 * d->b2 * d->b3 != BLOCK_M) return CTU_E_BADARG;
  if (d->b1 > 256 || d->b2 > 256 || d->b3 > 256) return CTU_E_BADARG;
  const int taps = d->k1 * d->k2 * d->k3;
  if (!((d->k1 == 1 && d->k2 == 1 && d->k3 == 1) || (d->k1 == 3 && d->k2 == 3 && d->k3 == 3))) return CTU_E_UNSUPPORTED;
  if (taps > 1 && (d->a_c % BLOCK_K) != 0) return CTU_E_UNSUPPORTED;
  if (d->k_total != taps * d->a_c) return CTU_E_BADARG;
  if ((d->lda % 8) != 0 || (d->k_total % 8) != 0 || (d->a_c % 8) != 0) return CTU_E_BADARG;
  if (d->n_pad % d->block_n != 0 || d->n_real > d->n_pad || d->n_real <= 0) return CTU_E_BADARG;
  if ((reinterpret_cast<uintptr_t>(d->a) & 15) || (reinterpret_cast<uintptr_t>(d->w) & 15) ||
      (reinterpret_cast<uintptr_t>(d->out) & 15))
    return CTU_E_BADARG;
  if (d->out_mode != CTU_OUT_F32_CF) {
    if ((d->ldc % 8) != 0 || (d->out_col0 % 8) != 0 || (d->n_real % 8) != 0) return CTU_E_BADARG;
  }
  if (d->res_mode != CTU_RES_NONE && (!d->residual || (d->ldr % 8) != 0)) return CTU_E_BADARG;
  if (d->convt_cout > 0) {
    if (d->convt_cout % d->block_n != 0 || d->n_real != d->convt_cout * d->u1 * d->u2 * d->u3) return CTU_E_BADARG;
    if (d->stats != nullptr) return CTU_E_UNSUPPORTED;
  }
  if (d->stats != nullptr && d->out_mode == CTU_OUT_F32_CF) return CTU_E_UNSUPPORTED;
  if (!tma_encoder()) return CTU_E_DRIVER;
  if (d->k1 == 3) {  // large 64/128-channel layers: halo-reuse kernel
    const int rc = conv3_halo_dispatch(d, stream);
    if (rc != CTU_E_UNSUPPORTED) return rc;
  }

  const CUtensorMapL2promotion l2p = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  CUtensorMap tmA, tmB, tmC;
  {
    // A: 5-D channels-last tensor map (C, d1, d2, d3, d4)
    cuuint64_t dims[5] = {(cuuint64_t)d->a_c, (cuuint64_t)d->d1, (cuuint64_t)d->d2, (cuuint64_t)d->d3, (cuuint64_t)d->d4};
    cuuint64_t strides[4];
    strides[0] = (cuuint64_t)d->lda * 2;
    strides[1] = strides[0] * d->d1;
    strides[2] = strides[1] * d->d2;
    strides[3] = strides[2] * d->d3;
    cuuint32_t box[5] = {(cuuint32_t)BLOCK_K, (cuuint32_t)d->b1, (cuuint32_t)d->b2, (cuuint32_t)d->b3, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = tma_encoder()(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(d->a), dims, strides, box,
                               es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, l2p,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return CTU_E_DRIVER;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)d->k_total, (cuuint64_t)d->n_pad};
    cuuint64_t strides[1] = {(cuuint64_t)d->k_total * 2};
    cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)d->block_n};
    cuuint32_t es[2] = {1, 1};
    CUresult r = tma_encoder()(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(d->w), dims, strides, box,
                               es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, l2p,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return CTU_E_DRIVER;
  }
  // C: bf16 row outputs of plain GEMMs / convs go through a swizzled smem tile and TMA stores
  // (transposed convolution / pixel shuffle: the tile's voxels land on a stride-u lattice of the up-sampled output — the
  // same TMA store with ELEMENT STRIDES u along the spatial axes and the sub-voxel offset in the start coordinates)
  static const int convt_tma = [] { const char* e = getenv("CTU_CONVT_TMA_STORE"); return e ? atoi(e) : 1; }();
  const bool convt = d->convt_cout > 0;
  const bool tma_store = d->out_mode == CTU_OUT_BF16_ROWS && d->block_n >= 64 &&
                         (!convt || (convt_tma && d->convt_cout % 64 == 0 && d->u1 <= 8 && d->u2 <= 8 && d->u3 <= 8 &&
                                     d->b1 * d->u1 <= 256 && d->b2 * d->u2 <= 256 && d->b3 * d->u3 <= 256));
  if (tma_store) {
    const int u1 = convt ? d->u1 : 1, u2 = convt ? d->u2 : 1, u3 = convt ? d->u3 : 1;
    cuuint64_t dims[5] = {(cuuint64_t)(convt ? d->convt_cout : d->n_real), (cuuint64_t)d->d1 * u1, (cuuint64_t)d->d2 * u2,
                          (cuuint64_t)d->d3 * u3, (cuuint64_t)d->d4};
    cuuint64_t strides[4];
    strides[0] = (cuuint64_t)d->ldc * 2;
    strides[1] = strides[0] * dims[1];
    strides[2] = strides[1] * dims[2];
    strides[3] = strides[2] * dims[3];
    cuuint32_t box[5] = {64, (cuuint32_t)(d->b1 * u1), (cuuint32_t)(d->b2 * u2), (cuuint32_t)(d->b3 * u3), 1};
    cuuint32_t es[5] = {1, (cuuint32_t)u1, (cuuint32_t)u2, (cuuint32_t)u3, 1};
    void* base = reinterpret_cast<__nv_bfloat16*>(d->out) + d->out_col0;
    CUresult r = tma_encoder()(&tmC, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, dims, strides, box, es,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, l2p,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return CTU_E_DRIVER;
  } else {
    tmC = tmA;  // unused
  }

  GemmParams p;
  p.b1 = d->b1; p.b2 = d->b2; p.b3 = d->b3;
  p.T1 = (d->d1 + d->b1 - 1) / d->b1;
  p.T2 = (d->d2 + d->b2 - 1) / d->b2;
  p.T3 = (d->d3 + d->b3 - 1) / d->b3;
  p.d1 = d->d1; p.d2 = d->d2; p.d3 = d->d3;
  p.n_tiles = d->n_pad / d->block_n;
  p.cblocks = (d->a_c + BLOCK_K - 1) / BLOCK_K;
  p.num_kb = taps * p.cblocks;
  p.a_c = d->a_c;
  p.k1 = d->k1; p.k2 = d->k2;
  p.pad = (d->k1 == 3) ? 1 : 0;
  p.out = d->out; p.bias = d->bias; p.residual = d->residual; p.stats = d->stats;
  p.n_real = d->n_real; p.out_mode = d->out_mode; p.ldc = d->ldc; p.act = d->act;
  p.res_mode = d->res_mode; p.ldr = d->ldr;
  p.convt_cout = d->convt_cout;
  p.u1 = d->convt_cout > 0 ? d->u1 : 1;
  p.u2 = d->convt_cout > 0 ? d->u2 : 1;
  p.u3 = d->convt_cout > 0 ? d->u3 : 1;
  p.stats_ld = d->stats_ld; p.out_col0 = d->out_col0;
  p.tma_store = tma_store ? 1 : 0;
  p.fast = (tma_store && d->res_mode != CTU_RES_F32) ? 1 : 0;
  static const int stage_env = [] { const char* e = getenv("CTU_GEMM_STAGE_RES"); return e ? atoi(e) : 1; }();
  p.stage_res = (stage_env && p.fast && (d->res_mode == CTU_RES_BF16 || d->res_mode == CTU_RES_GELU_BWD) && d->b1 == BLOCK_M &&
                 d->d2 == 1 && d->d3 == 1 && d->n_real % 8 == 0 && d->k1 == 1 &&
                 (reinterpret_cast<uintptr_t>(d->residual) & 15) == 0) ? 1 : 0;
  const long long tiles_ll = (long long)p.T1 * p.T2 * p.T3 * d->d4 * p.n_tiles;
  if (tiles_ll <= 0 || tiles_ll > 0x7fffffffLL) return CTU_E_BADARG;
  p.total_tiles = (int)tiles_ll;

  // Launch shape: K-heavy tiles (convolutions) run two CTAs per SM — two MMA issuers and two TMA streams hide the
  // shared-memory-bound operand feed of small-N tcgen05.mma; thin-K GEMMs run one CTA per SM with a deep ring.
  // Two (or three) persistent CTAs per SM: several MMA issuers / TMA streams per SM hide the latency of the
  // shared-memory-fed small-N tcgen05.mma and overlap one CTA's epilogue with another's main loop (measured: the
  // 64->64 conv at 96^3 runs 1.11 ms with 2 CTAs x 3 stages against 2.05 ms with 1 CTA x 6 stages).
  static const int variant = [] { const char* e = getenv("CTU_GEMM_VARIANT"); return e ? atoi(e) : 0; }();
  switch (d->block_n) {
    case 16: return launch_gemm<16, 4, 0, 2>(tmA, tmB, tmC, p, stream);
    case 32: return launch_gemm<32, 4, 0, 2>(tmA, tmB, tmC, p, stream);
    case 64:
      // 3x3x3 convolutions with 64 output channels: three CTAs per SM (measured 1.04 vs 1.10 ms at 96^3 x 4)
      if (variant == 3 || (variant == 0 && p.num_kb >= 27)) return launch_gemm<64, 2, 1, 3>(tmA, tmB, tmC, p, stream);
      return launch_gemm<64, 3, 1, 2, 8>(tmA, tmB, tmC, p, stream);
    case 128:
      // few tiles with a long K loop (ViT GEMMs: 42-168 tiles of 12-48 K blocks): one CTA per SM with a deep ring hides
      // the L2 latency that two 2-stage CTAs cannot when most SMs hold a single tile
      if (variant == 5 || (variant == 0 && p.num_kb >= 8 && p.total_tiles <= 2 * sm_count()))
        return launch_gemm<128, 5, 1, 1, 8>(tmA, tmB, tmC, p, stream);
      return launch_gemm<128, 2, 1, 2, 8>(tmA, tmB, tmC, p, stream);
    case 256: return launch_gemm<256, 3, 1, 1, 8>(tmA, tmB, tmC, p, stream);
    default: return CTU_E_UNSUPPORTED;
  }
}
