// tcgen05 / TMEM / TMA weight-gradient kernel for sm_100a (see include/ctunet_b200.h, ctu_umma_wgrad).
//
//   dW[(tap, ci), co] += sum over voxels v of  X[v + tap - pad, ci] * dY[v, co]
//
// is a GEMM whose contraction axis is the VOXEL axis, so both tensor-core operands are "MN-major": the very
// same TMA boxes the forward kernel loads (128 voxels x 64 channels, 128-byte rows, SWIZZLE_128B) are consumed
// through MN-major shared-memory descriptors (64-channel chunks LBO apart, 8-voxel groups SBO = 1024 B apart),
// 16 voxels per tcgen05.mma.  One work item = (group of J row tiles of 128 (tap, ci) rows) x (one BN-wide block
// of output channels) x (one contiguous chunk of voxel tiles):
//   warp 0 (1 lane) : TMA producer — per voxel tile one dY box per 64 output channels (reused by the J row
//                     tiles) and, per row tile, two activation slabs (two filter taps of a 64-channel layer,
//                     or two 64-channel halves of one tap) shifted by the tap with hardware zero fill.
//   warp 1 (1 lane) : MMA issuer — 8 x tcgen05.mma (M128 x BN x K16, both operands MN-major) per row tile and
//                     voxel tile into J TMEM accumulators that stay resident for the whole voxel chunk.
//   warps 2..5      : epilogue — tcgen05.ld and vectorised fp32 reductions (red.global.add.v4.f32) into dW;
//                     split-K over voxel chunks is what fills the 148 SMs.
#include "common.cuh"
#include "../../include/ctunet_b200.h"
#include "host_util.h"
#include <stdlib.h>

namespace ctu {

constexpr int WG_SLAB_BYTES = 128 * 128;          // 128 voxels x 64 bf16
constexpr int WG_A_STAGE_BYTES = 2 * WG_SLAB_BYTES;

struct WgradParams {
  int b1, b2, b3;
  int T1, T2, T3;
  int vox_tiles;   // T1*T2*T3*d4
  int splits;      // voxel chunks
  int MT, MG, NT;  // row tiles, row-tile groups, output-channel blocks
  int nslabs, cblocks;
  int k1, k2, pad;
  int d4;
  int total_items;
  float* dw;
  int ldw, n;
};

// MN-major operand tile, SWIZZLE_128B: 64-element (128 B) rows along M/N, 8 K-rows per 1024-byte atom (SBO),
// further 64-element chunks of M/N `lbo_bytes` apart.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__host__ __device__ constexpr uint32_t umma_idesc_bf16_mn(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

struct WgItem {
  int m0, m1, nt, v0, v1;  // row tiles [m0, m1), output-channel block, voxel tiles [v0, v1)
};

__device__ __forceinline__ WgItem wg_decode(const WgradParams& p, int item) {
  WgItem w;
  w.nt = item % p.NT;
  int t = item / p.NT;
  const int mg = t % p.MG;
  const int s = t / p.MG;
  // row tiles are spread evenly over the MG groups (at most J per group)
  w.m0 = (p.MT * mg) / p.MG;
  w.m1 = (p.MT * (mg + 1)) / p.MG;
  w.v0 = (int)(((long long)p.vox_tiles * s) / p.splits);
  w.v1 = (int)(((long long)p.vox_tiles * (s + 1)) / p.splits);
  return w;
}

template <int BN, int J, int SA, int SB, int CTAS_PER_SM>
__global__ void __launch_bounds__(192, CTAS_PER_SM) umma_wgrad_kernel(const __grid_constant__ CUtensorMap tmX,
                                                                      const __grid_constant__ CUtensorMap tmY,
                                                                      const WgradParams p) {
  constexpr int B_STAGE_BYTES = (BN / 64) * WG_SLAB_BYTES;
  constexpr int TMEM_COLS = J * BN;  // 256 or 512
  constexpr uint32_t IDESC = umma_idesc_bf16_mn(128, BN);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem_a + SA * WG_A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + SB * B_STAGE_BYTES);
  uint64_t* full_a = bars;
  uint64_t* empty_a = bars + SA;
  uint64_t* full_b = bars + 2 * SA;
  uint64_t* empty_b = bars + 2 * SA + SB;
  uint64_t* bar_tfull = bars + 2 * SA + 2 * SB;
  uint64_t* bar_tempty = bar_tfull + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tempty + 1);

  pdl_trigger();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmY);
    for (int s = 0; s < SA; ++s) { mbar_init(smem_u32(&full_a[s]), 1); mbar_init(smem_u32(&empty_a[s]), 1); }
    for (int s = 0; s < SB; ++s) { mbar_init(smem_u32(&full_b[s]), 1); mbar_init(smem_u32(&empty_b[s]), 1); }
    mbar_init(smem_u32(bar_tfull), 1);
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
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t ia = 0, ib = 0;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        const WgItem w = wg_decode(p, item);
        const int n0 = w.nt * BN;
        for (int vt = w.v0; vt < w.v1; ++vt) {
          int m = vt;
          const int x1 = (m % p.T1) * p.b1; m /= p.T1;
          const int x2 = (m % p.T2) * p.b2; m /= p.T2;
          const int x3 = (m % p.T3) * p.b3;
          const int t4 = m / p.T3;
          {
            const int s = ib % SB;
            mbar_wait(smem_u32(&empty_b[s]), ((ib / SB) & 1) ^ 1);
            const uint32_t full = smem_u32(&full_b[s]);
            mbar_expect_tx(full, B_STAGE_BYTES);
#pragma unroll
            for (int sl = 0; sl < BN / 64; ++sl)
              tma_load_5d(smem_u32(smem_b + s * B_STAGE_BYTES + sl * WG_SLAB_BYTES), &tmY, full, n0 + sl * 64, x1, x2, x3, t4);
            ++ib;
          }
          for (int mt = w.m0; mt < w.m1; ++mt) {
            const int s = ia % SA;
            mbar_wait(smem_u32(&empty_a[s]), ((ia / SA) & 1) ^ 1);
            const uint32_t full = smem_u32(&full_a[s]);
            mbar_expect_tx(full, WG_A_STAGE_BYTES);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int slab = 2 * mt + h;
              const uint32_t dst = smem_u32(smem_a + s * WG_A_STAGE_BYTES + h * WG_SLAB_BYTES);
              if (slab < p.nslabs) {
                const int tap = slab / p.cblocks;
                const int cb = slab - tap * p.cblocks;
                const int f1 = tap % p.k1;
                const int f2 = (tap / p.k1) % p.k2;
                const int f3 = tap / (p.k1 * p.k2);
                tma_load_5d(dst, &tmX, full, cb * 64, x1 + f1 - p.pad, x2 + f2 - p.pad, x3 + f3 - p.pad, t4);
              } else {
                // odd slab count: the second half of the last row tile is a box entirely outside the tensor
                // (batch index d4), which the TMA unit fills with zeros
                tma_load_5d(dst, &tmX, full, 0, x1, x2, x3, p.d4);
              }
            }
            ++ia;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      uint32_t ia = 0, ib = 0;
      int li = 0;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++li) {
        const WgItem w = wg_decode(p, item);
        mbar_wait(smem_u32(bar_tempty), (li & 1) ^ 1);  // epilogue has drained the accumulators of the last item
        tc_fence_after();
        for (int vt = w.v0; vt < w.v1; ++vt) {
          const int sb = ib % SB;
          mbar_wait(smem_u32(&full_b[sb]), (ib / SB) & 1);
          tc_fence_after();
          const uint64_t db = umma_desc_mn_sw128(smem_u32(smem_b + sb * B_STAGE_BYTES), WG_SLAB_BYTES);
          for (int mt = w.m0; mt < w.m1; ++mt) {
            const int j = mt - w.m0;
            const int sa = ia % SA;
            mbar_wait(smem_u32(&full_a[sa]), (ia / SA) & 1);
            tc_fence_after();
            const uint64_t da = umma_desc_mn_sw128(smem_u32(smem_a + sa * WG_A_STAGE_BYTES), WG_SLAB_BYTES);
            const uint32_t acc = tmem_base + (uint32_t)(j * BN);
            // eight K steps of 16 voxels = two 8-row swizzle atoms = 2048 bytes each: +128 in the (addr >> 4) field
            umma_bf16_k8(acc, da, db, 128ull, 128ull, IDESC, vt > w.v0 ? 1u : 0u);
            umma_commit(smem_u32(&empty_a[sa]));
            ++ia;
          }
          umma_commit(smem_u32(&empty_b[sb]));
          ++ib;
        }
        umma_commit(smem_u32(bar_tfull));
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int q = warp & 3;
    const int r = q * 32 + lane;
    int li = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++li) {
      const WgItem w = wg_decode(p, item);
      const int n0 = w.nt * BN;
      mbar_wait(smem_u32(bar_tfull), li & 1);
      tc_fence_after();
      if (w.v1 > w.v0) {
        for (int mt = w.m0; mt < w.m1; ++mt) {
          const int j = mt - w.m0;
          const int slab = 2 * mt + (r >> 6);
          const bool row_ok = slab < p.nslabs;
          float* drow = p.dw + ((long long)slab * 64 + (r & 63)) * p.ldw + n0;
#pragma unroll 1
          for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t raw[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * BN + c0), raw);
            tmem_ld_wait();
            if (row_ok) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                if (n0 + c0 + i < p.n)
                  red_add_v4(drow + c0 + i, __uint_as_float(raw[i]), __uint_as_float(raw[i + 1]),
                             __uint_as_float(raw[i + 2]), __uint_as_float(raw[i + 3]));
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

static int wg_sm_count() {
  static int n = [] {
    int dev = 0, v = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v;
  }();
  return n;
}

template <int BN, int J, int SA, int SB, int CTAS_PER_SM>
static int launch_wgrad(const CUtensorMap& tmX, const CUtensorMap& tmY, WgradParams p, int per_slot, cudaStream_t stream,
                        int max_splits = 0) {
  constexpr int smem = 1024 + SA * WG_A_STAGE_BYTES + SB * (BN / 64) * WG_SLAB_BYTES + (2 * SA + 2 * SB + 2) * 8 + 16;
  static_assert(CTAS_PER_SM * (smem + 1024) <= 228 * 1024, "shared memory budget");
  static_assert(CTAS_PER_SM * J * BN <= 512, "TMEM budget");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(umma_wgrad_kernel<BN, J, SA, SB, CTAS_PER_SM>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  p.MG = (p.MT + J - 1) / J;
  const int slots = persistent_sms(wg_sm_count()) * CTAS_PER_SM;
  const int base = p.MG * p.NT;
  // split the voxel axis so that every CTA slot gets `per_slot` work items: more, shorter items balance the SMs
  // better; fewer, longer items mean fewer fp32 reductions in the epilogue (measured per shape class, see
  // profiles/r01_wgrad_sweep.txt)
  static const int forced = [] { const char* e = getenv("CTU_WGRAD_ITEMS_PER_SLOT"); return e ? atoi(e) : 0; }();
  if (forced > 0) per_slot = forced;
  int splits = (per_slot * slots) / base;   // rounded DOWN: base * splits <= per_slot * slots, i.e. no CTA gets one item more
                                            // than the others (rounding up left e.g. 891 items on 148 CTAs: 7 vs 6)
  static const int forced_splits = [] { const char* e = getenv("CTU_WGRAD_SPLITS"); return e ? atoi(e) : 0; }();
  if (forced_splits > 0) splits = forced_splits;
  else if (max_splits > 0 && splits > max_splits) splits = max_splits;
  if (splits > p.vox_tiles) splits = p.vox_tiles;
  if (splits < 1) splits = 1;
  p.splits = splits;
  const long long items = (long long)base * splits;
  if (items > 0x7fffffffLL) return CTU_E_BADARG;
  p.total_items = (int)items;
  const int grid = p.total_items < slots ? p.total_items : slots;
  const cudaError_t le = launch_pdl(umma_wgrad_kernel<BN, J, SA, SB, CTAS_PER_SM>, dim3(grid), dim3(192), smem, stream, tmX, tmY, p);
  count_launch();
  return le != cudaSuccess ? (int)le : (int)cudaGetLastError();
}

}  // namespace ctu

extern "C" int ctu_umma_wgrad(const ctu_wgrad_desc* d, void* stream_) {
  using namespace ctu;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!d || !d->x || !d->dy || !d->dw) return CTU_E_BADARG;
  if (d->b1 * d->b2 * d->b3 != 128 || d->b1 > 256 || d->b2 > 256 || d->b3 > 256) return CTU_E_BADARG;
  if (!((d->k1 == 1 && d->k2 == 1 && d->k3 == 1) || (d->k1 == 3 && d->k2 == 3 && d->k3 == 3))) return CTU_E_UNSUPPORTED;
  if (d->x_c <= 0 || d->x_c % 64 != 0) return CTU_E_UNSUPPORTED;
  if (d->n <= 0 || d->n % 4 != 0 || d->ldw < d->n || d->ldw % 4 != 0) return CTU_E_BADARG;
  if (d->ldx % 8 != 0 || d->ldy % 8 != 0 || d->ldx < d->x_c || d->ldy < d->n) return CTU_E_BADARG;
  if ((reinterpret_cast<uintptr_t>(d->x) & 15) || (reinterpret_cast<uintptr_t>(d->dy) & 15) ||
      (reinterpret_cast<uintptr_t>(d->dw) & 15))
    return CTU_E_BADARG;
  if (d->block_n != 64 && d->block_n != 128 && d->block_n != 256) return CTU_E_UNSUPPORTED;
  if (!tma_encoder()) return CTU_E_DRIVER;
  if (d->k1 == 3) {  // 64/128-input-channel layers on the (8,16) tile grid: halo-reuse kernel
    const int rc = wgrad_halo_dispatch(d, stream);
    if (rc != CTU_E_UNSUPPORTED) return rc;
  }

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
    cuuint32_t box[5] = {64, (cuuint32_t)d->b1, (cuuint32_t)d->b2, (cuuint32_t)d->b3, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = tma_encoder()(which == 0 ? &tmX : &tmY, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base),
                               dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, l2p,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return CTU_E_DRIVER;
  }

  WgradParams p;
  p.b1 = d->b1; p.b2 = d->b2; p.b3 = d->b3;
  p.T1 = (d->d1 + d->b1 - 1) / d->b1;
  p.T2 = (d->d2 + d->b2 - 1) / d->b2;
  p.T3 = (d->d3 + d->b3 - 1) / d->b3;
  const long long vt = (long long)p.T1 * p.T2 * p.T3 * d->d4;
  if (vt <= 0 || vt > 0x7fffffffLL) return CTU_E_BADARG;
  p.vox_tiles = (int)vt;
  p.cblocks = d->x_c / 64;
  p.nslabs = d->k1 * d->k2 * d->k3 * p.cblocks;
  p.MT = (p.nslabs + 1) / 2;
  p.NT = (d->n + d->block_n - 1) / d->block_n;
  p.k1 = d->k1; p.k2 = d->k2;
  p.pad = d->k1 == 3 ? 1 : 0;
  p.d4 = d->d4;
  p.dw = d->dw; p.ldw = d->ldw; p.n = d->n;
  p.MG = 0; p.splits = 1; p.total_items = 0;

  // Launch shapes (CUDA-event sweep on B200, profiles/r01_wgrad_sweep.txt): several small CTAs per SM beat one CTA
  // with a deep ring — like the forward kernel, one MMA-issuing thread cannot keep the tensor pipe busy with
  // N <= 128 instructions — and large problems want 6-8 work items per CTA slot.
  static const int variant = [] { const char* e = getenv("CTU_WGRAD_VARIANT"); return e ? atoi(e) : -1; }();
  const bool conv = d->k1 == 3;
  switch (d->block_n) {
    case 64:
      if (variant == 0 || (variant < 0 && p.vox_tiles < 8000)) return launch_wgrad<64, 4, 2, 2, 2>(tmX, tmY, p, 1, stream);
      if (variant == 1) return launch_wgrad<64, 8, 3, 2, 1>(tmX, tmY, p, 2, stream);
      return launch_wgrad<64, 2, 1, 1, 4>(tmX, tmY, p, 8, stream);
    case 128:
      if (variant == 0) return launch_wgrad<128, 4, 3, 2, 1>(tmX, tmY, p, 2, stream);
      if (variant == 1 || (variant < 0 && (p.vox_tiles < 2000 || (!conv && p.MT * p.NT < 3))))
        return launch_wgrad<128, 2, 2, 1, 2>(tmX, tmY, p, 1, stream);
      return launch_wgrad<128, 1, 1, 1, 3>(tmX, tmY, p, conv ? 6 : 2, stream);
    case 256:
      if (variant == 1) return launch_wgrad<256, 1, 1, 1, 2>(tmX, tmY, p, 2, stream);
      // ViT-sized GEMMs (7 voxel tiles, 4.7-9.4 MB of dW): the fp32 reductions of the epilogue dominate — at most 3 splits
      return launch_wgrad<256, 2, 2, 2, 1>(tmX, tmY, p, conv && p.vox_tiles >= 400 ? 4 : (conv ? 1 : 2), stream,
                                           (!conv && p.vox_tiles <= 8) ? 3 : 0);
    default: return CTU_E_UNSUPPORTED;
  }
}
