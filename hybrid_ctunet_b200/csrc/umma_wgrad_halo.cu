// Weight gradient of the stride-1 "same" 3x3x3 Conv3d with HALO REUSE of the activation operand (sm_100a) — the
// wgrad counterpart of umma_conv3_halo.cu, for layers with 64 or 128 input channels whose (z, y) extents are
// multiples of (8, 16).
//
//   dW[(tap, ci), co] += sum over voxels v of  X[v + tap - 1, ci] * dY[v, co]
//
// A voxel tile is 8 (z) x 16 (y) voxels of one x-plane.  For the x-tap t3 of a work item the producer loads ONE
// 10 x 18-voxel halo box of X per 64-channel block (23 KB) and the 128-voxel dY tile; every (y, z) tap of that
// plane is a shifted view of the halo tile through an MN-major SWIZZLE_128B descriptor: the 8-voxel z-run of
// y-line l starts at row (l + ty) * 10 + tz, so K groups are 1280 B apart (SBO) and a 16-voxel MMA step advances
// by two lines.  The 128 rows of one MMA are two 64-channel slabs LBO apart: two TAPS of the same halo tile when
// C_in = 64 (LBO = their row distance), the two channel blocks of one tap when C_in = 128 (LBO = halo tile pitch).
// Per voxel tile an item issues J x 8 tcgen05.mma for 39-78 KB of loads (the per-tap kernel: 8 MMAs per 32 KB
// activation stage).  Work item = (x-tap, group of <= J row tiles, output-channel block, voxel-tile chunk);
// epilogue: fp32 reductions into dW as in umma_wgrad.cu.
#include "common.cuh"
#include "../../include/ctunet_b200.h"
#include "host_util.h"
#include <stdlib.h>

namespace ctu {

constexpr int WH_Z = 10, WH_Y = 18;
constexpr int WH_TILE_BYTES = WH_Z * WH_Y * 128;  // 23040
constexpr int WH_TILE_PITCH = 23 * 1024;
constexpr int WH_SLAB_BYTES = 128 * 128;

struct WhParams {
  int T1, T2, d3, d4;
  int vox_tiles, splits;
  int cblocks;      // 1 or 2
  int MTP;          // row tiles per x-plane: 5 (C_in = 64: tap pairs) or 9 (C_in = 128: one tap each)
  int SG;           // row-tile groups per plane
  int NT;
  int total_items;
  float* dw;
  int ldw, n, cin;
};

__device__ __forceinline__ uint64_t wh_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__host__ __device__ constexpr uint32_t wh_idesc(int m, int n) {  // bf16 x bf16 -> fp32, both operands MN-major
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}

struct WhItem {
  int t3, m0, m1, nt, v0, v1;
};

__device__ __forceinline__ WhItem wh_decode(const WhParams& p, int item) {
  WhItem w;
  w.nt = item % p.NT;
  int t = item / p.NT;
  const int sg = t % p.SG; t /= p.SG;
  w.t3 = t % 3;
  const int s = t / 3;
  w.m0 = (p.MTP * sg) / p.SG;
  w.m1 = (p.MTP * (sg + 1)) / p.SG;
  w.v0 = (int)(((long long)p.vox_tiles * s) / p.splits);
  w.v1 = (int)(((long long)p.vox_tiles * (s + 1)) / p.splits);
  return w;
}

// row offset (in 128-byte lines) of (y, z) tap t21 = ty * 3 + tz inside the halo tile
__device__ __forceinline__ int wh_tap_off(int t21) { return (t21 / 3) * WH_Z + (t21 % 3); }

// NI issuing warps (1 or 2): a tcgen05.mma of N = 64 is shorter than one thread's issue interval, so one CTA per SM with a
// single issuer starves the tensor pipe; with NI = 2 the row tiles of an item alternate between two issuing warps (different
// accumulators, shared operand stages).
template <int BN, int J, int CB, int CTAS_PER_SM, int ST, int NI>
__global__ void __launch_bounds__(160 + 32 * NI, CTAS_PER_SM) wgrad_halo_kernel(const __grid_constant__ CUtensorMap tmX,
                                                                      const __grid_constant__ CUtensorMap tmY,
                                                                      const WhParams p) {
  constexpr int A_BYTES = CB * WH_TILE_PITCH;
  constexpr int B_BYTES = (BN / 64) * WH_SLAB_BYTES;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;   // ST stages: one (halo tiles + dY tile) set each
  constexpr int TMEM_COLS = (J * BN <= 128) ? 128 : (J * BN <= 256 ? 256 : 512);
  constexpr uint32_t IDESC = wh_idesc(128, BN);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem_a + A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ST * STAGE_BYTES);
  uint64_t* bar_full = bars;            // [ST] halo tile(s) + dY tile landed
  uint64_t* bar_empty = bars + ST;      // [ST] MMAs that read them completed
  uint64_t* bar_tfull = bars + 2 * ST;
  uint64_t* bar_tempty = bars + 2 * ST + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * ST + 2);

  pdl_trigger();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmY);
    for (int st = 0; st < ST; ++st) {
      mbar_init(smem_u32(&bar_full[st]), 1);
      mbar_init(smem_u32(&bar_empty[st]), NI);
    }
    mbar_init(smem_u32(bar_tfull), NI);
    mbar_init(smem_u32(bar_tempty), 4);
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
    // ------------------------------------------------------------------ TMA producer (one stage per CTA; several
    // CTAs per SM overlap each other's loads and MMAs)
    if (lane == 0) {
      uint32_t it = 0;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        const WhItem w = wh_decode(p, item);
        const int n0 = w.nt * BN;
        for (int vt = w.v0; vt < w.v1; ++vt, ++it) {
          int m = vt;
          const int z0 = (m % p.T1) * 8; m /= p.T1;
          const int y0 = (m % p.T2) * 16; m /= p.T2;
          const int x = m % p.d3;
          const int b = m / p.d3;
          const int st = it % ST;
          mbar_wait(smem_u32(&bar_empty[st]), ((it / ST) & 1) ^ 1);
          const uint32_t full = smem_u32(&bar_full[st]);
          mbar_expect_tx(full, CB * WH_TILE_BYTES + B_BYTES);
#pragma unroll
          for (int cb = 0; cb < CB; ++cb)
            tma_load_5d(smem_u32(smem_a + st * STAGE_BYTES + cb * WH_TILE_PITCH), &tmX, full, cb * 64, z0 - 1, y0 - 1,
                        x + w.t3 - 1, b);
#pragma unroll
          for (int sl = 0; sl < BN / 64; ++sl)
            tma_load_5d(smem_u32(smem_b + st * STAGE_BYTES + sl * WH_SLAB_BYTES), &tmY, full, n0 + sl * 64, z0, y0, x, b);
        }
      }
    }
  } else if (warp <= NI) {
    // ------------------------------------------------------------------ MMA issuer(s): warp 1 (and 2)
    if (lane == 0) {
      const int iw = warp - 1;
      uint32_t it = 0;
      int li = 0;
      const uint32_t a_base0 = smem_u32(smem_a);
      const uint32_t b_base0 = smem_u32(smem_b);
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++li) {
        const WhItem w = wh_decode(p, item);
        mbar_wait(smem_u32(bar_tempty), (li & 1) ^ 1);
        tc_fence_after();
        for (int vt = w.v0; vt < w.v1; ++vt, ++it) {
          const int st = it % ST;
          mbar_wait(smem_u32(&bar_full[st]), (it / ST) & 1);
          tc_fence_after();
          const uint32_t a_base = a_base0 + (uint32_t)(st * STAGE_BYTES);
          const uint64_t db0 = wh_desc(b_base0 + (uint32_t)(st * STAGE_BYTES), WH_SLAB_BYTES, 1024);
          for (int mt = w.m0 + iw; mt < w.m1; mt += NI) {
            // row tile -> start offset of its first slab and the byte distance to its second slab
            uint32_t off0, lbo;
            if (CB == 1) {
              const int ta = 2 * mt, tb = (2 * mt + 1 < 9) ? 2 * mt + 1 : 2 * mt;
              off0 = (uint32_t)wh_tap_off(ta) * 128u;
              lbo = (uint32_t)(wh_tap_off(tb) - wh_tap_off(ta)) * 128u;
              if (lbo == 0) lbo = 128u;  // dummy second slab of the odd tap: its rows are never written out
            } else {
              off0 = (uint32_t)wh_tap_off(mt) * 128u;
              lbo = (uint32_t)WH_TILE_PITCH;
            }
            const uint32_t acc = tmem_base + (uint32_t)((mt - w.m0) * BN);
            // eight K steps of 16 voxels = two y-lines each: 2 x 1280 B further in the halo tile, 2 x 1024 B in the dY tile
            // (one asm block: the issuing thread, not the tensor pipe, bounds this loop when its descriptors are rebuilt
            // and its predicate re-evaluated per instruction)
            umma_bf16_k8(acc, wh_desc(a_base + off0, lbo, WH_Z * 128), db0, (uint64_t)((2 * WH_Z * 128) >> 4),
                         (uint64_t)(2048 >> 4), IDESC, vt > w.v0 ? 1u : 0u);
          }
          umma_commit(smem_u32(&bar_empty[st]));
        }
        umma_commit(smem_u32(bar_tfull));
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (four warps after the issuers)
    const int q = warp & 3;
    const int r = q * 32 + lane;
    int li = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++li) {
      const WhItem w = wh_decode(p, item);
      const int n0 = w.nt * BN;
      mbar_wait(smem_u32(bar_tfull), li & 1);
      tc_fence_after();
      if (w.v1 > w.v0) {
        for (int mt = w.m0; mt < w.m1; ++mt) {
          int tap21, ci;
          bool row_ok = true;
          if (CB == 1) {
            tap21 = 2 * mt + (r >> 6);
            ci = r & 63;
            row_ok = tap21 < 9;
          } else {
            tap21 = mt;
            ci = r;
          }
          const int tap = w.t3 * 9 + tap21;
          float* drow = p.dw + ((long long)tap * p.cin + ci) * p.ldw + n0;
#pragma unroll 1
          for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t raw[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((mt - w.m0) * BN + c0), raw);
            tmem_ld_wait();
            if (row_ok) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                if (n0 + c0 + i < p.n)
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(drow + c0 + i),
                               "f"(__uint_as_float(raw[i])), "f"(__uint_as_float(raw[i + 1])),
                               "f"(__uint_as_float(raw[i + 2])), "f"(__uint_as_float(raw[i + 3]))
                               : "memory");
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(bar_tempty));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

static int wh_sm_count() {
  static int n = [] {
    int dev = 0, v = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v;
  }();
  return n;
}

template <int BN, int J, int CB, int CTAS_PER_SM, int ST = 1, int NI = 1>
static int launch_wh(const CUtensorMap& tmX, const CUtensorMap& tmY, WhParams p, int per_slot, cudaStream_t stream) {
  constexpr int smem = 1024 + ST * (CB * WH_TILE_PITCH + (BN / 64) * WH_SLAB_BYTES) + (2 * ST + 2) * 8 + 16;
  constexpr int tmem = (J * BN <= 128) ? 128 : (J * BN <= 256 ? 256 : 512);
  static_assert(CTAS_PER_SM * (smem + 1024) <= 228 * 1024, "shared memory budget");
  static_assert(CTAS_PER_SM * tmem <= 512, "TMEM budget");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_halo_kernel<BN, J, CB, CTAS_PER_SM, ST, NI>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  p.SG = (p.MTP + J - 1) / J;
  const int slots = persistent_sms(wh_sm_count()) * CTAS_PER_SM;
  const int base = 3 * p.SG * p.NT;
  static const int forced = [] { const char* e = getenv("CTU_WGRAD_ITEMS_PER_SLOT"); return e ? atoi(e) : 0; }();
  if (forced > 0) per_slot = forced;
  int splits = (per_slot * slots) / base;   // rounded DOWN: base * splits <= per_slot * slots, i.e. no CTA gets one item more
                                            // than the others (rounding up left e.g. 891 items on 148 CTAs: 7 vs 6)
  if (splits > p.vox_tiles) splits = p.vox_tiles;
  if (splits < 1) splits = 1;
  p.splits = splits;
  p.total_items = base * splits;
  const int grid = p.total_items < slots ? p.total_items : slots;
  const cudaError_t le = launch_pdl(wgrad_halo_kernel<BN, J, CB, CTAS_PER_SM, ST, NI>, dim3(grid), dim3(160 + 32 * NI), smem, stream, tmX, tmY, p);
  count_launch();
  return le != cudaSuccess ? (int)le : (int)cudaGetLastError();
}

// Returns CTU_E_UNSUPPORTED when the problem does not fit this kernel (the caller then uses umma_wgrad_kernel).
int wgrad_halo_dispatch(const ctu_wgrad_desc* d, cudaStream_t stream) {
  // CTU_WGRAD_HALO=0 switches this kernel off (A/B).  Measured on B200, batch 2 (profiles/r01_wgrad_halo_sweep.txt):
  // 64->64 @96^3 636 -> 854 TFLOP/s, 128->64 @96^3 672 -> 813, 128->128 @48x48x96 840 -> 950, 64->64 @48x48x96 600 -> 686.
  static const int mode = [] { const char* e = getenv("CTU_WGRAD_HALO"); return e ? atoi(e) : 1; }();
  if (mode == 0) return CTU_E_UNSUPPORTED;
  if (d->k1 != 3 || d->k2 != 3 || d->k3 != 3) return CTU_E_UNSUPPORTED;
  if (d->x_c != 64 && d->x_c != 128) return CTU_E_UNSUPPORTED;
  if (d->block_n != 64 && d->block_n != 128) return CTU_E_UNSUPPORTED;
  if (d->d1 % 8 != 0 || d->d2 % 16 != 0 || d->n % 64 != 0) return CTU_E_UNSUPPORTED;
  if (!tma_encoder()) return CTU_E_DRIVER;
  const CUtensorMapL2promotion l2p = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  CUtensorMap tmX, tmY;
  for (int which = 0; which < 2; ++which) {
    const int c = which == 0 ? d->x_c : d->n;
    const int ld = which == 0 ? d->ldx : d->ldy;
    const void* base = which == 0 ? d->x : d->dy;
    cuuint64_t dims[5] = {(cuuint64_t)c, (cuuint64_t)d->d1, (cuuint64_t)d->d2, (cuuint64_t)d->d3, (cuuint64_t)d->d4};
    cuuint64_t strides[4];
    strides[0] = (cuuint64_t)ld * 2;
    strides[1] = strides[0] * d->d1;
    strides[2] = strides[1] * d->d2;
    strides[3] = strides[2] * d->d3;
    cuuint32_t box[5] = {64, (cuuint32_t)(which == 0 ? WH_Z : 8), (cuuint32_t)(which == 0 ? WH_Y : 16), 1, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    if (tma_encoder()(which == 0 ? &tmX : &tmY, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides,
                      box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, l2p,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return CTU_E_DRIVER;
  }
  WhParams p;
  p.T1 = d->d1 / 8;
  p.T2 = d->d2 / 16;
  p.d3 = d->d3; p.d4 = d->d4;
  const long long vt = (long long)p.T1 * p.T2 * d->d3 * d->d4;
  if (vt <= 0 || vt > 0x7fffffffLL) return CTU_E_BADARG;
  p.vox_tiles = (int)vt;
  p.cblocks = d->x_c / 64;
  p.MTP = p.cblocks == 1 ? 5 : 9;
  p.NT = (d->n + d->block_n - 1) / d->block_n;
  p.dw = d->dw; p.ldw = d->ldw; p.n = d->n; p.cin = d->x_c;
  p.SG = 0; p.splits = 1; p.total_items = 0;
  static const int variant = [] { const char* e = getenv("CTU_WGRAD_HALO_VARIANT"); return e ? atoi(e) : 0; }();
  static const int ps_env = [] { const char* e = getenv("CTU_WGRAD_HALO_PER_SLOT"); return e ? atoi(e) : 0; }();
  const int ps = ps_env > 0 ? ps_env : ((p.vox_tiles >= 8000 || d->x_c >= 128) ? 8 : 4);
  // Variants 2 / 3: ALL (or half of) the row tiles of an x-tap in one item — every (halo tile, dY tile) set is then read by
  // 3 (or 6) items instead of 9-15: the J = 2 shapes move 9-10 TB/s from L2 to shared memory (ncu: 4.1 GB per launch for
  // 128 -> 128 @48x48x96, 18x the algorithmic bytes), i.e. they sit at the L2 throughput cap with the tensor pipe 43 % busy.
  // The wide accumulators leave room for one CTA per SM only, so the loads are pipelined inside the CTA (ST stages).
  const int ps1 = ps_env > 0 ? ps_env : 2;
  if (d->x_c == 64) {
    if (d->block_n == 64) {
      if (variant == 1) return launch_wh<64, 3, 1, 2>(tmX, tmY, p, ps, stream);
      if (variant == 2) return launch_wh<64, 5, 1, 1, 3>(tmX, tmY, p, ps1, stream);
      if (variant == 3) return launch_wh<64, 3, 1, 2, 2>(tmX, tmY, p, ps, stream);
      // default: all five row tiles of an x-tap in one item, ONE CTA per SM with a three-stage ring and TWO issuing warps:
      // a third of the L2 -> shared-memory traffic of the J = 2 shape.  64 -> 64 @96^3 x 2: 843 -> 1,027 TFLOP/s (items per
      // CTA 2: 736, 4: 857, 6: 1,027, 8: 899, 12: 929, 16: 854); with ONE issuing warp the same shape runs at 534.
      // Variant 6: the former four-CTA J = 2 shape.
      if (variant == 6) return launch_wh<64, 2, 1, 4>(tmX, tmY, p, ps, stream);
      return launch_wh<64, 5, 1, 1, 3, 2>(tmX, tmY, p, ps_env > 0 ? ps_env : 3, stream);   // after the split-rounding fix: 3: 1,078, 4: 1,061, 6: 1,030
    }
    if (variant == 2) return launch_wh<128, 3, 1, 1, 3>(tmX, tmY, p, ps1, stream);
    return launch_wh<128, 2, 1, 2>(tmX, tmY, p, ps, stream);
  }
  if (d->block_n == 64) {
    if (variant == 1) return launch_wh<64, 3, 2, 2>(tmX, tmY, p, ps, stream);
    if (variant == 2) return launch_wh<64, 5, 2, 1, 3>(tmX, tmY, p, ps1, stream);
    // 128 -> 64 @96^3 x 2: 805 -> 1,026 TFLOP/s with the same one-CTA, two-issuer shape (five + four row tiles per x-tap)
    if (variant == 6) return launch_wh<64, 2, 2, 3>(tmX, tmY, p, ps, stream);
    return launch_wh<64, 5, 2, 1, 3, 2>(tmX, tmY, p, ps_env > 0 ? ps_env : 4, stream);
  }
  // 128 -> 128: three row tiles per item on one CTA per SM with a two-stage ring (930 -> 1,129 TFLOP/s at 48x48x96 x 2);
  // variant 4: the former two-CTA J = 2 shape.  (The wide-J shapes LOSE for 64 output channels: 850 -> 534 TFLOP/s.)
  if (variant == 4) return launch_wh<128, 2, 2, 2>(tmX, tmY, p, ps, stream);
  return launch_wh<128, 3, 2, 1, 2>(tmX, tmY, p, ps_env > 0 ? ps_env : 3, stream);   // items per CTA (after the split-rounding fix) 3: 1,351, 4: 1,312, 5: 1,286, 6: 1,249, 8: 1,198 TFLOP/s
}

}  // namespace ctu
