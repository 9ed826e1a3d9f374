// Stride-1 "same" 3x3x3 Conv3d as a tcgen05 implicit GEMM with HALO REUSE of the activation operand (sm_100a).
//
// umma_gemm_kernel fetches one 128-voxel x 64-channel box per filter tap: 27 boxes (432 KB) per output tile, which
// makes the 64-channel layers bound by the L2 -> shared-memory port, not by the tensor pipe.  Here an output tile is
// 8 (z) x 16 (y) voxels of one x-plane and the producer loads, per x-tap and 64-channel block, ONE halo box of
// 10 x 18 voxels (23 KB).  All nine (y, z) taps of that plane are then shifted VIEWS of the same shared-memory tile:
// rows are 128-byte voxel lines, an 8-voxel z-run of output row y is the 8-row group starting at row
// (y + ty) * 10 + tz, so the K-major SWIZZLE_128B descriptor of tap (ty, tz) is the tile base + (ty*10 + tz) * 128
// bytes with a stride-byte-offset of 1280 between groups (the swizzle is a function of the shared-memory address, so
// TMA's writes and the shifted tensor-core reads agree).  Activation traffic drops 6.4x (3 x 22.5 KB instead of
// 27 x 16 KB per tile); the weights (one 64 x BN box per tap) stream through their own ring.
// Epilogue: bf16 rows through a swizzled staging tile + TMA store, InstanceNorm statistics (fp64 atomics), as in
// umma_gemm.cu.  Used for the 64/128-channel layers whose (z, y) extents are multiples of (8, 16).
//
// ZT = 2 (BN = 64, z extent a multiple of 16, opt-in): one CTA computes TWO z-adjacent 128-voxel tiles from one 18 x 18
// halo box and ONE stream of weight boxes — every weight box feeds two accumulators (L2 -> shared-memory traffic per
// 128 voxels 285 KB -> 170 KB).  Measured neutral: see conv3_halo_dispatch.
#include "common.cuh"
#include "../../include/ctunet_b200.h"
#include "host_util.h"
#include <stdlib.h>

namespace ctu {

constexpr int HALO_Y = 18;
constexpr int HALO_SLAB_BYTES = 128 * 128;
__host__ __device__ constexpr int halo_z(int zt) { return 8 * zt + 2; }
__host__ __device__ constexpr int halo_a_bytes(int zt) { return halo_z(zt) * HALO_Y * 128; }                 // written by TMA
__host__ __device__ constexpr int halo_a_stage(int zt) { return (halo_a_bytes(zt) + 1023) / 1024 * 1024; }  // ring pitch

struct HaloParams {
  int T1, T2, d1, d2, d3, d4;
  int n_tiles, total_tiles, cblocks, a_c;
  double* stats;
  int n_real, stats_ld;
  int base_offset_mode;
  const __nv_bfloat16* residual;  // bf16 rows shaped like the output (may BE the output: in-place accumulation) or NULL
  int ldr, res_col0;
  int ksteps;  // K16 steps per 64-channel block that can be non-zero (4, or fewer for zero-padded channel rows)
};

struct HaloTile {
  int z0, y0, x, b, n0;
};

__device__ __forceinline__ HaloTile halo_decode(const HaloParams& p, int tile, int bn, int zext) {
  HaloTile t;
  const int n_tile = tile % p.n_tiles;
  int m = tile / p.n_tiles;
  t.z0 = (m % p.T1) * zext; m /= p.T1;
  t.y0 = (m % p.T2) * 16; m /= p.T2;
  t.x = m % p.d3;
  t.b = m / p.d3;
  t.n0 = n_tile * bn;
  return t;
}

__device__ __forceinline__ uint64_t halo_desc_a(uint32_t saddr, int base_offset_mode, int hz) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((hz * 128) >> 4) << 32;  // 8-row groups are one halo line (10 or 18 rows) apart
  d |= (uint64_t)1 << 46;
  if (base_offset_mode) d |= (uint64_t)((saddr >> 7) & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int BN, int SA, int SB, int CTAS_PER_SM, int ZT>
__global__ void __launch_bounds__(192, CTAS_PER_SM) conv3_halo_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                      const __grid_constant__ CUtensorMap tmB,
                                                                      const __grid_constant__ CUtensorMap tmC,
                                                                      const HaloParams p) {
  constexpr int B_STAGE_BYTES = BN * 128;
  constexpr int SLABS = BN / 64;
  constexpr int TMEM_COLS = 2 * ZT * BN;   // two accumulator slots, ZT tiles each
  constexpr uint32_t IDESC = umma_idesc_bf16(128, BN);
  constexpr int HALO_Z = halo_z(ZT);
  constexpr int HALO_A_BYTES = halo_a_bytes(ZT);
  constexpr int HALO_A_STAGE = halo_a_stage(ZT);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem_a + SA * HALO_A_STAGE;
  uint8_t* smem_c = smem_b + SB * B_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_c + ZT * SLABS * HALO_SLAB_BYTES);
  uint64_t* full_a = bars;
  uint64_t* empty_a = bars + SA;
  uint64_t* full_b = bars + 2 * SA;
  uint64_t* empty_b = bars + 2 * SA + SB;
  uint64_t* bar_tfull = bars + 2 * SA + 2 * SB;
  uint64_t* bar_tempty = bar_tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tempty + 2);
  float2* stat_scratch = reinterpret_cast<float2*>(tmem_slot + 2);  // [4][BN]

  pdl_trigger();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    for (int s = 0; s < SA; ++s) { mbar_init(smem_u32(&full_a[s]), 1); mbar_init(smem_u32(&empty_a[s]), 1); }
    for (int s = 0; s < SB; ++s) { mbar_init(smem_u32(&full_b[s]), 1); mbar_init(smem_u32(&empty_b[s]), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&bar_tfull[s]), 1); mbar_init(smem_u32(&bar_tempty[s]), 4); }
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
      uint32_t ia = 0, ib = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const HaloTile t = halo_decode(p, tile, BN, 8 * ZT);
        for (int t3 = 0; t3 < 3; ++t3) {
          for (int cb = 0; cb < p.cblocks; ++cb) {
            {
              const int s = ia % SA;
              mbar_wait(smem_u32(&empty_a[s]), ((ia / SA) & 1) ^ 1);
              const uint32_t full = smem_u32(&full_a[s]);
              mbar_expect_tx(full, HALO_A_BYTES);
              tma_load_5d(smem_u32(smem_a + s * HALO_A_STAGE), &tmA, full, cb * 64, t.z0 - 1, t.y0 - 1, t.x + t3 - 1, t.b);
              ++ia;
            }
            for (int t21 = 0; t21 < 9; ++t21) {
              const int tap = t3 * 9 + t21;
              const int s = ib % SB;
              mbar_wait(smem_u32(&empty_b[s]), ((ib / SB) & 1) ^ 1);
              const uint32_t full = smem_u32(&full_b[s]);
              mbar_expect_tx(full, B_STAGE_BYTES);
              tma_load_2d(smem_u32(smem_b + s * B_STAGE_BYTES), &tmB, full, tap * p.a_c + cb * 64, t.n0);
              ++ib;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      uint32_t ia = 0, ib = 0;
      int lt = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++lt) {
        const int slot = lt & 1;
        mbar_wait(smem_u32(&bar_tempty[slot]), ((lt >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)(slot * ZT * BN);
        uint32_t first = 1;
        for (int t3 = 0; t3 < 3; ++t3) {
          for (int cb = 0; cb < p.cblocks; ++cb) {
            const int sa = ia % SA;
            mbar_wait(smem_u32(&full_a[sa]), (ia / SA) & 1);
            tc_fence_after();
            const uint32_t a_base = smem_u32(smem_a + sa * HALO_A_STAGE);
            for (int t2 = 0; t2 < 3; ++t2) {
#pragma unroll
              for (int t1 = 0; t1 < 3; ++t1) {
                const int sb = ib % SB;
                mbar_wait(smem_u32(&full_b[sb]), (ib / SB) & 1);
                tc_fence_after();
                const uint64_t db = umma_desc_k_sw128(smem_u32(smem_b + sb * B_STAGE_BYTES));
#pragma unroll
                for (int h = 0; h < ZT; ++h) {   // the z-adjacent tiles: same weights, views 8 halo rows apart
                  const uint64_t da = halo_desc_a(a_base + (uint32_t)((t2 * HALO_Z + t1 + 8 * h) * 128), p.base_offset_mode, HALO_Z);
                  if (p.ksteps == 4) {
                    umma_bf16_k4(acc + (uint32_t)(h * BN), da, db, IDESC, first ? 0u : 1u);
                  } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                      if (k < p.ksteps) umma_bf16(acc + (uint32_t)(h * BN), da + 2 * k, db + 2 * k, IDESC, (first && k == 0) ? 0u : 1u);
                    }
                  }
                }
                first = 0;
                umma_commit(smem_u32(&empty_b[sb]));
                ++ib;
              }
            }
            umma_commit(smem_u32(&empty_a[sa]));
            ++ia;
          }
        }
        umma_commit(smem_u32(&bar_tfull[slot]));
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int e = threadIdx.x - 64;
    const int i1 = r & 7, i2 = r >> 3;
    constexpr int STAT_PER_THREAD = (BN + 127) / 128;
    float acc_s[STAT_PER_THREAD], acc_q[STAT_PER_THREAD];
#pragma unroll
    for (int k = 0; k < STAT_PER_THREAD; ++k) acc_s[k] = acc_q[k] = 0.f;
    int stat_batch = -1;
    const int n0_cta = (blockIdx.x % p.n_tiles) * BN;
    auto flush_stats = [&](int b) {
      if (b < 0) return;
#pragma unroll
      for (int k = 0; k < STAT_PER_THREAD; ++k) {
        const int c = e + 128 * k;
        if (c < BN && n0_cta + c < p.n_real) {
          double* dst = p.stats + ((long long)b * p.stats_ld + n0_cta + c) * 2;
          atomicAdd(dst, (double)acc_s[k]);
          atomicAdd(dst + 1, (double)acc_q[k]);
        }
        acc_s[k] = acc_q[k] = 0.f;
      }
    };

    int lt = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++lt) {
      const int slot = lt & 1;
      const HaloTile t = halo_decode(p, tile, BN, 8 * ZT);
      mbar_wait(smem_u32(&bar_tfull[slot]), (lt >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int h = 0; h < ZT; ++h) {   // the z-adjacent 128-voxel tiles of this CTA tile, one staging buffer each
        const int zt0 = t.z0 + 8 * h;
        const bool valid = (zt0 + i1 < p.d1) && (t.y0 + i2 < p.d2);
        uint8_t* cbuf = smem_c + (size_t)h * SLABS * HALO_SLAB_BYTES;
        // this staging buffer must have been read by its previous TMA store; the statistics scratch must be free
        if (e == 0) {
          if (ZT > 1) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        named_bar_sync(1, 128);
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t raw[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((slot * ZT + h) * BN + c0), raw);
          tmem_ld_wait();
          if (p.residual != nullptr && valid) {
            // out = acc + residual (the input-gradient convolution adding to a gradient that has already arrived)
            const long long row = (((long long)t.b * p.d3 + t.x) * p.d2 + (t.y0 + i2)) * p.d1 + (zt0 + i1);
            const __nv_bfloat16* rp = p.residual + row * p.ldr + p.res_col0 + t.n0 + c0;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              const uint4 rv = *reinterpret_cast<const uint4*>(rp + j);
              float2 f;
              f = unpack_bf16x2(rv.x); raw[j] = __float_as_uint(__uint_as_float(raw[j]) + f.x); raw[j + 1] = __float_as_uint(__uint_as_float(raw[j + 1]) + f.y);
              f = unpack_bf16x2(rv.y); raw[j + 2] = __float_as_uint(__uint_as_float(raw[j + 2]) + f.x); raw[j + 3] = __float_as_uint(__uint_as_float(raw[j + 3]) + f.y);
              f = unpack_bf16x2(rv.z); raw[j + 4] = __float_as_uint(__uint_as_float(raw[j + 4]) + f.x); raw[j + 5] = __float_as_uint(__uint_as_float(raw[j + 5]) + f.y);
              f = unpack_bf16x2(rv.w); raw[j + 6] = __float_as_uint(__uint_as_float(raw[j + 6]) + f.x); raw[j + 7] = __float_as_uint(__uint_as_float(raw[j + 7]) + f.y);
            }
          }
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(__uint_as_float(raw[2 * j]), __uint_as_float(raw[2 * j + 1]));
          const uint32_t base = smem_u32(cbuf) + (uint32_t)((c0 >> 6) * HALO_SLAB_BYTES + r * 128);
          const int cbk = (c0 & 63) >> 3;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t addr = base + (uint32_t)(((cbk + i) ^ (r & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * i]), "r"(pk[4 * i + 1]),
                         "r"(pk[4 * i + 2]), "r"(pk[4 * i + 3])
                         : "memory");
          }
          if (p.stats != nullptr) {
            float v[32], sq[32];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float2 f = unpack_bf16x2(pk[j]);
              v[2 * j] = valid ? f.x : 0.f;
              v[2 * j + 1] = valid ? f.y : 0.f;
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) sq[j] = v[j] * v[j];
            const float s_sum = warp_transpose_reduce(v, lane);
            const float s_sq = warp_transpose_reduce(sq, lane);
            stat_scratch[q * BN + c0 + lane] = make_float2(s_sum, s_sq);
          }
        }
        if (h == ZT - 1) {   // both halves of the accumulator slot have been read
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[slot]));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        named_bar_sync(2, 128);
        if (e == 0) {
#pragma unroll
          for (int sl = 0; sl < SLABS; ++sl) {
            if (t.n0 + sl * 64 < p.n_real) {
              asm volatile(
                  "cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(&tmC),
                  "r"(smem_u32(cbuf + sl * HALO_SLAB_BYTES)), "r"(t.n0 + sl * 64), "r"(zt0), "r"(t.y0), "r"(t.x), "r"(t.b)
                  : "memory");
            }
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (p.stats != nullptr) {
          if (t.b != stat_batch) {
            flush_stats(stat_batch);
            stat_batch = t.b;
          }
#pragma unroll
          for (int k = 0; k < STAT_PER_THREAD; ++k) {
            const int c = e + 128 * k;
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
    if (p.stats != nullptr) flush_stats(stat_batch);
    if (e == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------------------------
// TWO OUTPUT x-PLANES PER TILE for the 64-output-channel layers (conv3_halo_x2_kernel).
//
// An M128 x N64 x K16 tcgen05.mma re-reads a 4 KB A tile and a 2 KB B tile from shared memory per 32 tensor-pipe cycles:
// the N = 64 layers are bound by that operand feed (~2/3 of the N = 128 rate, see conv3_halo_dispatch).  Here a tile is the
// same 8 (z) x 16 (y) voxels of TWO adjacent planes x, x+1 and the accumulator is 128 columns wide, [plane x | plane x+1].
// The halo box of input plane p is the A operand of BOTH planes' taps that read it (dx = p - x for plane x, dx - 1 for
// plane x+1), so for p = x and p = x+1 one N = 128 instruction does the work of two N = 64 ones with ONE read of the A
// tile (8 KB per 64 cycles: balanced); p = x-1 and p = x+2 feed one plane each (N = 64).  Per (y, z) tap and K16 step:
// 2 x 64 + 2 x 48 = 224 cycles for two planes instead of 6 x 48 = 288, and four halo boxes instead of six.
// The weights come from a re-laid copy (CTU_PACK_X3_FROM_PACKED): per (y,z)-tap and N tile the three x-taps are 192
// consecutive rows [W(dx=+1) | W(0) | W(-1)], so [W(0); W(-1)] (p = x) and [W(+1); W(0)] (p = x+1) are contiguous 128-row
// operands.  Plane order p = x, x-1, x+1, x+2: the first instruction of a tile (N = 128, no accumulate) initialises both
// halves of the accumulator.
// ZT = 2: the tile also spans two z-adjacent 8 x 16 voxel blocks (one 18 x 18 halo box per input plane): every weight stage
// then feeds four accumulators — the two-plane kernel with ZT = 1 moves 3.7 GB from L2 to shared memory per launch of
// 64 -> 64 @96^3 x 2 (9.6 TB/s, the L2 throughput cap; ncu), 83 % of it weights.
// SLOTS = 1: a single accumulator slot and ONE staging buffer per CTA (72 KB of shared memory, 128 TMEM columns): three
// CTAs per SM — the epilogue of a CTA then blocks its own main loop, the two other CTAs keep the tensor pipe fed.
// NI = ZT = 2: one issuing warp per z block (same operand stages, different accumulators) — one thread cannot issue N = 64
// instructions fast enough to keep the tensor pipe busy.
template <int SB, int CTAS_PER_SM, int ZT, int SA, int SLOTS, int NI>
__global__ void __launch_bounds__(160 + 32 * NI, CTAS_PER_SM) conv3_halo_x2_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                         const __grid_constant__ CUtensorMap tmB,
                                                                         const __grid_constant__ CUtensorMap tmC,
                                                                         const HaloParams p) {
  constexpr int BN = 64;
  constexpr int B_STAGE_BYTES = 128 * 128;        // up to 128 weight rows x 64 K
  constexpr int SLOT_COLS = ZT * 2 * BN;          // one accumulator slot: ZT x [plane x | plane x+1]
  constexpr int TMEM_COLS = SLOTS * SLOT_COLS;
  constexpr int CBUFS = SLOTS == 1 ? 1 : 2 * ZT;  // staging buffers
  constexpr int NSUB = 2 * ZT;                    // (z block, plane) sub-tiles of 128 voxels x 64 channels
  constexpr uint32_t IDESC64 = umma_idesc_bf16(128, 64);
  constexpr uint32_t IDESC128 = umma_idesc_bf16(128, 128);
  constexpr int HALO_Z = halo_z(ZT);
  constexpr int HALO_A_BYTES = halo_a_bytes(ZT);
  constexpr int HALO_A_STAGE = halo_a_stage(ZT);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                                  // SA halo stages
  uint8_t* smem_b = smem_a + SA * HALO_A_STAGE;
  uint8_t* smem_c = smem_b + SB * B_STAGE_BYTES;           // [NSUB][128 rows x 64 bf16]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_c + CBUFS * HALO_SLAB_BYTES);
  uint64_t* full_a = bars;
  uint64_t* empty_a = bars + SA;
  uint64_t* full_b = bars + 2 * SA;
  uint64_t* empty_b = bars + 2 * SA + SB;
  uint64_t* bar_tfull = bars + 2 * SA + 2 * SB;
  uint64_t* bar_tempty = bar_tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tempty + 2);
  float2* stat_scratch = reinterpret_cast<float2*>(tmem_slot + 2);  // [4][BN]

  pdl_trigger();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    for (int st = 0; st < SA; ++st) { mbar_init(smem_u32(&full_a[st]), 1); mbar_init(smem_u32(&empty_a[st]), NI); }
    for (int s = 0; s < SB; ++s) { mbar_init(smem_u32(&full_b[s]), 1); mbar_init(smem_u32(&empty_b[s]), NI); }
    for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&bar_tfull[s]), NI); mbar_init(smem_u32(&bar_tempty[s]), 4); }
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
  pdl_wait();

  // tile -> (n tile, z tile, y tile, plane pair, batch); p.d3 is the number of plane PAIRS here
  auto decode = [&](int tile) {
    HaloTile t;
    const int n_tile = tile % p.n_tiles;
    int m = tile / p.n_tiles;
    t.z0 = (m % p.T1) * (8 * ZT); m /= p.T1;
    t.y0 = (m % p.T2) * 16; m /= p.T2;
    t.x = (m % p.d3) * 2;
    t.b = m / p.d3;
    t.n0 = n_tile * BN;
    return t;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t ia = 0, ib = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const HaloTile t = decode(tile);
        const int nt = t.n0 / BN;
#pragma unroll 1
        for (int pi = 0; pi < 4; ++pi) {
          const int pl = pi == 0 ? 1 : (pi == 1 ? 0 : pi);        // input plane x - 1 + pl, order 1, 0, 2, 3
          const int rows = (pl == 1 || pl == 2) ? 128 : 64;
          const int dxi0 = pl == 0 ? 2 : (pl == 1 ? 1 : 0);       // first x-tap block of the operand: 0: +1, 1: 0, 2: -1
          for (int cb = 0; cb < p.cblocks; ++cb) {
            {
              const int sa = ia % SA;
              mbar_wait(smem_u32(&empty_a[sa]), ((ia / SA) & 1) ^ 1);
              const uint32_t full = smem_u32(&full_a[sa]);
              mbar_expect_tx(full, HALO_A_BYTES);
              tma_load_5d(smem_u32(smem_a + sa * HALO_A_STAGE), &tmA, full, cb * 64, t.z0 - 1, t.y0 - 1, t.x + pl - 1, t.b);
              ++ia;
            }
            for (int t21 = 0; t21 < 9; ++t21) {
              const int s = ib % SB;
              mbar_wait(smem_u32(&empty_b[s]), ((ib / SB) & 1) ^ 1);
              const uint32_t full = smem_u32(&full_b[s]);
              mbar_expect_tx(full, rows * 128);
              const int row0 = ((t21 * p.n_tiles + nt) * 3 + dxi0) * 64;
              tma_load_2d(smem_u32(smem_b + s * B_STAGE_BYTES), &tmB, full, cb * 64, row0);
              if (rows == 128) tma_load_2d(smem_u32(smem_b + s * B_STAGE_BYTES + 64 * 128), &tmB, full, cb * 64, row0 + 64);
              ++ib;
            }
          }
        }
      }
    }
  } else if (warp <= NI) {
    // ------------------------------------------------------------------ MMA issuer(s)
    if (lane == 0) {
      const int iw = warp - 1;   // NI = 2: this warp issues for z block iw
      uint32_t ia = 0, sbph = 0;
      int sbi = 0;
      int lt = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++lt) {
        const int slot = SLOTS == 2 ? (lt & 1) : 0;
        mbar_wait(smem_u32(&bar_tempty[slot]), ((SLOTS == 2 ? (lt >> 1) : lt) & 1) ^ 1);
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)(slot * SLOT_COLS);
        uint32_t first = 1;
        // the issuing thread is on the critical path (two CTAs per SM share the tensor pipe): descriptors are a constant
        // plus a stage / tap offset, the stage index and phase are carried instead of divided out, and the four K16
        // instructions of a tap are one asm block with a single predicate
        const uint64_t db0 = umma_desc_k_sw128(smem_u32(smem_b));
#pragma unroll 1
        for (int pi = 0; pi < 4; ++pi) {
          const int pl = pi == 0 ? 1 : (pi == 1 ? 0 : pi);
          const uint32_t idesc = (pl == 1 || pl == 2) ? IDESC128 : IDESC64;
          const uint32_t dst = acc + (pl == 3 ? (uint32_t)BN : 0u);
          for (int cb = 0; cb < p.cblocks; ++cb) {
            const int sa = ia % SA;
            mbar_wait(smem_u32(&full_a[sa]), (ia / SA) & 1);
            tc_fence_after();
            const uint64_t da0 = halo_desc_a(smem_u32(smem_a + sa * HALO_A_STAGE), 0, HALO_Z);
#pragma unroll
            for (int t21 = 0; t21 < 9; ++t21) {
              mbar_wait(smem_u32(&full_b[sbi]), sbph);
              tc_fence_after();
              const uint64_t db = db0 + (uint64_t)(sbi * (B_STAGE_BYTES >> 4));
              const uint64_t da = da0 + (uint64_t)((((t21 / 3) * HALO_Z + (t21 % 3)) * 128) >> 4);
#pragma unroll
              for (int h = 0; h < ZT; ++h) {   // the z-adjacent blocks: same weights, views 8 halo rows further
                if (NI == 1 || h == iw)
                  umma_bf16_k4(dst + (uint32_t)(h * 2 * BN), da + (uint64_t)((8 * h * 128) >> 4), db, idesc, first ? 0u : 1u);
              }
              first = 0;
              umma_commit(smem_u32(&empty_b[sbi]));
              if (++sbi == SB) { sbi = 0; sbph ^= 1u; }
            }
            umma_commit(smem_u32(&empty_a[sa]));
            ++ia;
          }
        }
        umma_commit(smem_u32(&bar_tfull[slot]));
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (four warps after the issuers)
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int e = threadIdx.x - 32 * (1 + NI);
    const int i1 = r & 7, i2 = r >> 3;
    float acc_s = 0.f, acc_q = 0.f;   // BN = 64 <= 128 epilogue threads: one statistics column per thread
    int stat_batch = -1;
    const int n0_cta = (blockIdx.x % p.n_tiles) * BN;
    auto flush_stats = [&](int b) {
      if (b < 0) return;
      if (e < BN && n0_cta + e < p.n_real) {
        double* dstp = p.stats + ((long long)b * p.stats_ld + n0_cta + e) * 2;
        atomicAdd(dstp, (double)acc_s);
        atomicAdd(dstp + 1, (double)acc_q);
      }
      acc_s = acc_q = 0.f;
    };

    int lt = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++lt) {
      const int slot = SLOTS == 2 ? (lt & 1) : 0;
      const HaloTile t = decode(tile);
      mbar_wait(smem_u32(&bar_tfull[slot]), (SLOTS == 2 ? (lt >> 1) : lt) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int sub = 0; sub < NSUB; ++sub) {   // z block zb, plane x + h: columns [sub * 64, sub * 64 + 64) of the slot
        const int zb = sub >> 1, h = sub & 1;
        const int zt0 = t.z0 + 8 * zb;
        const bool valid = (zt0 + i1 < p.d1) && (t.y0 + i2 < p.d2);
        uint8_t* cbuf = smem_c + (size_t)(sub % CBUFS) * HALO_SLAB_BYTES;
        if (e == 0) {   // this buffer's previous store (CBUFS groups ago) has read it
          if (CBUFS == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          else if (CBUFS == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          else asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
        }
        named_bar_sync(1, 128);
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t raw[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(slot * SLOT_COLS + sub * BN + c0), raw);
          tmem_ld_wait();
          if (p.residual != nullptr && valid) {
            const long long row = (((long long)t.b * (2 * p.d3) + (t.x + h)) * p.d2 + (t.y0 + i2)) * p.d1 + (zt0 + i1);
            const __nv_bfloat16* rp = p.residual + row * p.ldr + p.res_col0 + t.n0 + c0;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              const uint4 rv = *reinterpret_cast<const uint4*>(rp + j);
              float2 f;
              f = unpack_bf16x2(rv.x); raw[j] = __float_as_uint(__uint_as_float(raw[j]) + f.x); raw[j + 1] = __float_as_uint(__uint_as_float(raw[j + 1]) + f.y);
              f = unpack_bf16x2(rv.y); raw[j + 2] = __float_as_uint(__uint_as_float(raw[j + 2]) + f.x); raw[j + 3] = __float_as_uint(__uint_as_float(raw[j + 3]) + f.y);
              f = unpack_bf16x2(rv.z); raw[j + 4] = __float_as_uint(__uint_as_float(raw[j + 4]) + f.x); raw[j + 5] = __float_as_uint(__uint_as_float(raw[j + 5]) + f.y);
              f = unpack_bf16x2(rv.w); raw[j + 6] = __float_as_uint(__uint_as_float(raw[j + 6]) + f.x); raw[j + 7] = __float_as_uint(__uint_as_float(raw[j + 7]) + f.y);
            }
          }
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(__uint_as_float(raw[2 * j]), __uint_as_float(raw[2 * j + 1]));
          const uint32_t base = smem_u32(cbuf) + (uint32_t)(r * 128);
          const int cbk = (c0 & 63) >> 3;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t addr = base + (uint32_t)(((cbk + i) ^ (r & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * i]), "r"(pk[4 * i + 1]),
                         "r"(pk[4 * i + 2]), "r"(pk[4 * i + 3])
                         : "memory");
          }
          if (p.stats != nullptr) {
            float v[32], sq[32];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float2 f = unpack_bf16x2(pk[j]);
              v[2 * j] = valid ? f.x : 0.f;
              v[2 * j + 1] = valid ? f.y : 0.f;
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) sq[j] = v[j] * v[j];
            const float s_sum = warp_transpose_reduce(v, lane);
            const float s_sq = warp_transpose_reduce(sq, lane);
            stat_scratch[q * BN + c0 + lane] = make_float2(s_sum, s_sq);
          }
        }
        if (sub == NSUB - 1) {   // every sub-tile of the accumulator slot has been read
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[slot]));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        named_bar_sync(2, 128);
        if (e == 0) {
          if (t.n0 < p.n_real) {
            asm volatile(
                "cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(&tmC),
                "r"(smem_u32(cbuf)), "r"(t.n0), "r"(zt0), "r"(t.y0), "r"(t.x + h), "r"(t.b)
                : "memory");
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (p.stats != nullptr) {
          if (t.b != stat_batch) {
            flush_stats(stat_batch);
            stat_batch = t.b;
          }
          if (e < BN) {
            const float2 s0 = stat_scratch[e], s1 = stat_scratch[BN + e], s2 = stat_scratch[2 * BN + e], s3 = stat_scratch[3 * BN + e];
            acc_s += (s0.x + s1.x) + (s2.x + s3.x);
            acc_q += (s0.y + s1.y) + (s2.y + s3.y);
          }
        }
      }
    }
    if (p.stats != nullptr) flush_stats(stat_batch);
    if (e == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

static int halo_sm_count() {
  static int n = [] {
    int dev = 0, v = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v;
  }();
  return n;
}

template <int BN, int SA, int SB, int CTAS_PER_SM, int ZT = 1>
static int launch_halo(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const HaloParams& p,
                       cudaStream_t stream) {
  constexpr int smem = 1024 + SA * halo_a_stage(ZT) + SB * BN * 128 + ZT * (BN / 64) * HALO_SLAB_BYTES +
                       (2 * SA + 2 * SB + 4) * 8 + 16 + 4 * BN * 8;
  static_assert(CTAS_PER_SM * (smem + 1024) <= 228 * 1024, "shared memory budget");
  static_assert(CTAS_PER_SM * 2 * ZT * BN <= 512, "TMEM budget");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv3_halo_kernel<BN, SA, SB, CTAS_PER_SM, ZT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  int cap = persistent_sms(halo_sm_count()) * CTAS_PER_SM;
  if (cap > p.n_tiles) cap -= cap % p.n_tiles;
  const int grid = p.total_tiles < cap ? p.total_tiles : cap;
  const cudaError_t le = launch_pdl(conv3_halo_kernel<BN, SA, SB, CTAS_PER_SM, ZT>, dim3(grid), dim3(192), smem, stream, tmA, tmB, tmC, p);
  count_launch();
  return le != cudaSuccess ? (int)le : (int)cudaGetLastError();
}

template <int SB, int CTAS_PER_SM, int ZT = 1, int SA = 1, int SLOTS = 2, int NI = 1>
static int launch_halo_x2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const HaloParams& p,
                          cudaStream_t stream) {
  constexpr int smem = 1024 + SA * halo_a_stage(ZT) + SB * 128 * 128 + (SLOTS == 1 ? 1 : 2 * ZT) * HALO_SLAB_BYTES + (2 * SA + 2 * SB + 4) * 8 + 16 +
                       4 * 64 * 8;
  static_assert(CTAS_PER_SM * (smem + 1024) <= 228 * 1024, "shared memory budget");
  static_assert(CTAS_PER_SM * SLOTS * 2 * ZT * 64 <= 512, "TMEM budget");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv3_halo_x2_kernel<SB, CTAS_PER_SM, ZT, SA, SLOTS, NI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  int cap = persistent_sms(halo_sm_count()) * CTAS_PER_SM;
  if (cap > p.n_tiles) cap -= cap % p.n_tiles;
  const int grid = p.total_tiles < cap ? p.total_tiles : cap;
  const cudaError_t le = launch_pdl(conv3_halo_x2_kernel<SB, CTAS_PER_SM, ZT, SA, SLOTS, NI>, dim3(grid), dim3(160 + 32 * NI), smem, stream, tmA, tmB, tmC, p);
  count_launch();
  return le != cudaSuccess ? (int)le : (int)cudaGetLastError();
}

// Returns CTU_E_UNSUPPORTED when the problem does not fit this kernel (the caller then uses umma_gemm_kernel).
int conv3_halo_dispatch(const ctu_gemm_desc* d, cudaStream_t stream) {
  // CTU_CONV_HALO=0 switches this kernel off (A/B comparisons).  Measured on B200, batch 2 (profiles/r01_halo_sweep.txt):
  // 64->64 @96^3 744 -> 1008 TFLOP/s, 128->64 @96^3 782 -> 1074, 128->128 @48x48x96 1035 -> 1298.
  // (Setting the descriptor's base-offset field for the shifted views gives WRONG results: the hardware applies the
  // swizzle to the absolute shared-memory address, so the field stays 0.)
  static const int mode = [] { const char* e = getenv("CTU_CONV_HALO"); return e ? atoi(e) : 1; }();
  if (mode == 0) return CTU_E_UNSUPPORTED;
  if (d->k1 != 3 || d->k2 != 3 || d->k3 != 3) return CTU_E_UNSUPPORTED;
  if (d->out_mode != CTU_OUT_BF16_ROWS || d->bias || d->act != CTU_ACT_NONE || d->convt_cout > 0) return CTU_E_UNSUPPORTED;
  if (d->residual && (d->res_mode != CTU_RES_BF16 || (d->ldr % 8) != 0)) return CTU_E_UNSUPPORTED;
  if (d->block_n != 64 && d->block_n != 128) return CTU_E_UNSUPPORTED;
  if (d->d1 % 8 != 0 || d->d2 % 16 != 0 || d->a_c % 64 != 0 || d->n_real % 64 != 0) return CTU_E_UNSUPPORTED;
  if (!tma_encoder()) return CTU_E_DRIVER;
  const CUtensorMapL2promotion l2p = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  static const int variant = [] { const char* e = getenv("CTU_CONV_HALO_VARIANT"); return e ? atoi(e) : 0; }();
  // CTU_CONV_HALO_VARIANT=3: two z-adjacent tiles per CTA sharing one weight stream (64-output-channel layers).  Measured
  // on B200 (profiles/r02_halo_zt2.txt): 64->64 @96^3 0.402 vs 0.398 ms, 128->64 0.760 vs 0.740, 64->64 @48x48x96 0.113 vs
  // 0.118 — halving the weight traffic changes nothing, i.e. these layers are NOT bound by L2 -> shared-memory traffic
  // but by the tensor pipe's shared-memory operand feed at N = 64 (each M128 x N64 x K16 instruction re-reads its 4 KB A
  // tile + 2 KB B tile: 6 KB per 32 tensor-pipe cycles > 128 B/clk), a ceiling of ~2/3 of the N >= 128 rate.  Off by
  // default.
  const int zt = (d->block_n == 64 && d->d1 % 16 == 0 && variant == 3) ? 2 : 1;
  // two output x-planes per tile (N = 128 instructions for the two shared input planes): needs the re-laid weight copy
  static const int x2_mode = [] { const char* e = getenv("CTU_CONV_HALO_X2"); return e ? atoi(e) : 1; }();
  const bool x2 = x2_mode != 0 && d->block_n == 64 && d->w_x3 != nullptr && d->d3 % 2 == 0 && zt == 1 &&
                  !(d->a_c == 64 && d->a_c_live > 0 && d->a_c_live < 64);
  // Two z blocks per tile on ONE CTA per SM (TMEM 512 columns, 43 % less L2 -> shared-memory traffic) with TWO issuing warps,
  // one per z block: 64->64 @96^3 x 4 0.654 ms (1,198 TFLOP/s), 128->64 1.238 ms, against 0.695 / 1.285 for three single-slot
  // CTAs per SM (CTU_CONV_HALO_X2_ZT=1, also the shape for z extents that are not multiples of 16) and 0.764 / 1.435 for the
  // same tile with ONE issuing warp (CTU_CONV_HALO_X2_NI=1): a tcgen05.mma of N = 64 is shorter than a thread's issue interval.
  // Stand-alone the ZT = 2 shape is the fastest; inside the two-lane CUDA graph it is not (conv class 12.12 vs 11.89 ms per
  // training step, inference 26.20 vs 25.89 ms per 4 windows): a 218 KB CTA leaves no room for the other lane's kernels on the
  // SM, three 72 KB CTAs do.  Opt-in: CTU_CONV_HALO_X2_ZT=2.
  static const int x2_zt = [] { const char* e = getenv("CTU_CONV_HALO_X2_ZT"); return e ? atoi(e) : 1; }();
  const int zt_box = (x2 && x2_zt == 2 && d->d1 % 16 == 0) ? 2 : zt;   // z blocks per tile (halo box 18 instead of 10 deep)
  CUtensorMap tmA, tmB, tmC;
  {
    cuuint64_t dims[5] = {(cuuint64_t)d->a_c, (cuuint64_t)d->d1, (cuuint64_t)d->d2, (cuuint64_t)d->d3, (cuuint64_t)d->d4};
    cuuint64_t strides[4];
    strides[0] = (cuuint64_t)d->lda * 2;
    strides[1] = strides[0] * d->d1;
    strides[2] = strides[1] * d->d2;
    strides[3] = strides[2] * d->d3;
    cuuint32_t box[5] = {64, (cuuint32_t)halo_z(zt_box), HALO_Y, 1, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    if (tma_encoder()(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(d->a), dims, strides, box, es,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, l2p,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return CTU_E_DRIVER;
  }
  if (x2) {   // w_x3: [9 (y,z) taps][n_pad / 64 N tiles][3 x-taps (+1, 0, -1)][64 rows] x a_c columns
    cuuint64_t dims[2] = {(cuuint64_t)d->a_c, (cuuint64_t)(9 * (d->n_pad / 64) * 192)};
    cuuint64_t strides[1] = {(cuuint64_t)d->a_c * 2};
    cuuint32_t box[2] = {64, 64};
    cuuint32_t es[2] = {1, 1};
    if (tma_encoder()(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(d->w_x3), dims, strides, box, es,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, l2p,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return CTU_E_DRIVER;
  } else {
    cuuint64_t dims[2] = {(cuuint64_t)d->k_total, (cuuint64_t)d->n_pad};
    cuuint64_t strides[1] = {(cuuint64_t)d->k_total * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)d->block_n};
    cuuint32_t es[2] = {1, 1};
    if (tma_encoder()(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(d->w), dims, strides, box, es,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, l2p,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return CTU_E_DRIVER;
  }
  {
    cuuint64_t dims[5] = {(cuuint64_t)d->n_real, (cuuint64_t)d->d1, (cuuint64_t)d->d2, (cuuint64_t)d->d3, (cuuint64_t)d->d4};
    cuuint64_t strides[4];
    strides[0] = (cuuint64_t)d->ldc * 2;
    strides[1] = strides[0] * d->d1;
    strides[2] = strides[1] * d->d2;
    strides[3] = strides[2] * d->d3;
    cuuint32_t box[5] = {64, 8, 16, 1, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    void* base = reinterpret_cast<__nv_bfloat16*>(d->out) + d->out_col0;
    if (tma_encoder()(&tmC, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, l2p, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return CTU_E_DRIVER;
  }
  HaloParams p;
  p.T1 = d->d1 / (8 * zt);
  p.T2 = d->d2 / 16;
  p.d1 = d->d1; p.d2 = d->d2; p.d3 = d->d3; p.d4 = d->d4;
  p.n_tiles = d->n_pad / d->block_n;
  p.cblocks = d->a_c / 64;
  p.a_c = d->a_c;
  p.stats = d->stats; p.n_real = d->n_real; p.stats_ld = d->stats_ld;
  p.base_offset_mode = 0;
  // zero-padded channel rows (a_c == 64 with fewer live channels): the K steps over the padding are skipped
  p.ksteps = (d->a_c == 64 && d->a_c_live > 0 && d->a_c_live < 64) ? (d->a_c_live + 15) / 16 : 4;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(d->residual); p.ldr = d->ldr; p.res_col0 = d->out_col0;
  const long long tiles = (long long)p.T1 * p.T2 * d->d3 * d->d4 * p.n_tiles;
  if (tiles <= 0 || tiles > 0x7fffffffLL) return CTU_E_BADARG;
  p.total_tiles = (int)tiles;
  // several small CTAs per SM (one halo stage each) beat fewer CTAs with deeper rings: 3 x <64,1,4> reaches 1008
  // TFLOP/s where 2 x <64,2,5> reaches 895 and 1 x <64,3,8> 483
  if (x2) {
    p.d3 = d->d3 / 2;                                   // plane pairs
    p.total_tiles = (int)(tiles / 2);
    if (zt_box == 2) {
      p.T1 = d->d1 / 16;
      p.total_tiles /= 2;
      static const int zt_ni = [] { const char* e = getenv("CTU_CONV_HALO_X2_NI"); return e ? atoi(e) : 2; }();
      if (zt_ni == 2) return launch_halo_x2<4, 1, 2, 2, 2, 2>(tmA, tmB, tmC, p, stream);
      return launch_halo_x2<4, 1, 2, 2>(tmA, tmB, tmC, p, stream);
    }
    static const int x2_variant = [] { const char* e = getenv("CTU_CONV_HALO_X2_VARIANT"); return e ? atoi(e) : 0; }();
    // measured at 96^3 x 4, 64 -> 64 / 128 -> 64 (ms): three single-slot CTAs per SM 0.704 / 1.291; two double-slot CTAs
    // with a 3-stage weight ring 0.725 / 1.351, with a 2-stage ring 0.770 / 1.424; one CTA per SM 1.331 / 2.555
    if (x2_variant == 1) return launch_halo_x2<4, 1>(tmA, tmB, tmC, p, stream);
    if (x2_variant == 2) return launch_halo_x2<2, 2>(tmA, tmB, tmC, p, stream);
    if (x2_variant == 4) return launch_halo_x2<3, 2>(tmA, tmB, tmC, p, stream);
    return launch_halo_x2<2, 3, 1, 1, 1>(tmA, tmB, tmC, p, stream);
  }
  if (d->block_n == 64) {
    if (zt == 2) return launch_halo<64, 1, 4, 2, 2>(tmA, tmB, tmC, p, stream);
    if (variant == 2) return launch_halo<64, 2, 5, 2>(tmA, tmB, tmC, p, stream);
    return launch_halo<64, 1, 4, 3>(tmA, tmB, tmC, p, stream);
  }
  if (variant == 2) return launch_halo<128, 2, 4, 1>(tmA, tmB, tmC, p, stream);
  return launch_halo<128, 1, 3, 2>(tmA, tmB, tmC, p, stream);
}

}  // namespace ctu
