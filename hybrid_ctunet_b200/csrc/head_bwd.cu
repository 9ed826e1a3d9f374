// Backward of a logits head (UnetOutBlock 1x1x1 conv + bias, hybrid_CTUNet.py:781-783,810; DecoderLinear, :671-691)
// in ONE pass on the CUDA cores.  The head maps C (64 / 128 / 256) channels to n_cls = 14 logits, so its gradients are
// skinny contractions (K or N = 14) that cannot feed a tensor-core tile: run as GEMMs they cost four launches over
// 64-channel zero-padded copies of the logit gradient (layout conversion, bias column sum, weight gradient, input
// gradient).  Here each block stages a tile of the NCDHW fp32 logit gradient in shared memory and every thread owns
// one pair of input channels for a band of voxels:
//     da[v, c]  (+)= sum_k g[v, k] * W[k, c]          (input gradient, bf16 channels-last, optional in-place add)
//     dW[c, k]  +=   sum_v a[v, c] * g[v, k]          (fp32, block-reduced, one atomic per (c, k) per block)
//     db[k]     +=   sum_v g[v, k]
// HBM traffic: g (14 x 4 B) + a (2C B) + da (2C B, twice when accumulating) per voxel — the algorithmic minimum.
#include <cstdlib>

#include "attention_common.cuh"
#include "../../include/ctunet_b200.h"
#include "host_util.h"

namespace ctu {

constexpr int HB_TV = 128;    // voxels per tile
constexpr int HB_KP = 16;     // shared-memory row: classes padded to 16 (n_cls <= 16)

// NK: classes the arithmetic loops cover (14 for the reference's heads, else 16)
template <int C, int NK>
__global__ void __launch_bounds__(256, 2) head_bwd_kernel(const float* __restrict__ g, const __nv_bfloat16* __restrict__ a,
                                                          long long lda, const float* __restrict__ w,
                                                          __nv_bfloat16* __restrict__ da, long long ldda, int accumulate,
                                                          float* __restrict__ dw, int ldw, float* __restrict__ db, int B,
                                                          long long S, int ncls) {
  constexpr int PAIRS = C / 2;            // channel pairs per voxel row
  constexpr int LANES = 256 / PAIRS;      // voxel lanes per block (C = 64: 8, 128: 4, 256: 2)
  constexpr int VPL = HB_TV / LANES;      // voxels per lane per tile
  constexpr int CHV = 8;                  // voxels whose loads are in flight together
  constexpr int GJ = HB_TV * HB_KP / 256; // staged logit-gradient elements per thread per tile (8)
  static_assert(VPL % CHV == 0, "voxel chunking");
  __shared__ __align__(16) float gs[HB_TV][HB_KP];
  __shared__ float red[LANES][C];         // cross-lane reduction of one class at a time

  const int t = threadIdx.x;
  const int p = t % PAIRS, vl = t / PAIRS;
  float w0[NK], w1[NK], acc0[NK], acc1[NK];
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    w0[k] = k < ncls ? __ldg(w + (long long)k * C + 2 * p) : 0.f;
    w1[k] = k < ncls ? __ldg(w + (long long)k * C + 2 * p + 1) : 0.f;
    acc0[k] = acc1[k] = 0.f;
  }
  // the staging loop gives thread t the voxel t % 128 of classes t/128 + 2j: their running sums are the bias gradient
  const int gi = t % HB_TV, gk0 = t / HB_TV;
  float bsum[GJ], gq[GJ];
#pragma unroll
  for (int j = 0; j < GJ; ++j) bsum[j] = 0.f;

  const long long tiles_per_b = (S + HB_TV - 1) / HB_TV;
  const long long tiles = tiles_per_b * B;
  auto load_g = [&](long long tile) {   // next tile's logit gradients: in flight while the current tile is computed
    const int b = (int)(tile / tiles_per_b);
    const long long s = (tile - (long long)b * tiles_per_b) * HB_TV + gi;
    const float* gp = g + ((long long)b * ncls + gk0) * S + s;
#pragma unroll
    for (int j = 0; j < GJ; ++j) {
      const int k = gk0 + 2 * j;
      gq[j] = (k < ncls && s < S) ? __ldg(gp + (long long)(2 * j) * S) : 0.f;
    }
  };
  if ((long long)blockIdx.x < tiles) load_g(blockIdx.x);

  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int b = (int)(tile / tiles_per_b);
    const long long s0 = (tile - (long long)b * tiles_per_b) * HB_TV;
    const long long row0 = (long long)b * S + s0 + vl * VPL;
    const char* arow = reinterpret_cast<const char*>(a + row0 * lda + 2 * p);
    char* drow = reinterpret_cast<char*>(da + row0 * ldda + 2 * p);
    const long long astep = lda * 2, dstep = ldda * 2;
    const int nvalid = (int)((S - s0 - vl * VPL) < VPL ? (S - s0 - vl * VPL) : VPL);   // voxels of this lane inside the volume
    uint32_t araw[CHV], oraw[CHV];
    auto fetch = [&](int chunk) {
#pragma unroll
      for (int i = 0; i < CHV; ++i) {
        const int v = chunk * CHV + i;
        araw[i] = v < nvalid ? *reinterpret_cast<const uint32_t*>(arow + v * astep) : 0u;
        oraw[i] = (v < nvalid && accumulate) ? *reinterpret_cast<const uint32_t*>(drow + v * dstep) : 0u;
      }
    };
    fetch(0);
    __syncthreads();  // previous tile fully consumed
#pragma unroll
    for (int j = 0; j < GJ; ++j) {
      const int k = gk0 + 2 * j;
      gs[gi][(((k >> 2) ^ ((gi >> 1) & 3)) << 2) | (k & 3)] = gq[j];   // 16-byte groups swizzled by the voxel
      bsum[j] += gq[j];
    }
    __syncthreads();
    if (tile + gridDim.x < tiles) load_g(tile + gridDim.x);
#pragma unroll 1
    for (int chunk = 0; chunk < VPL / CHV; ++chunk) {
      if (chunk > 0) fetch(chunk);
#pragma unroll
      for (int i = 0; i < CHV; ++i) {
        const int v = chunk * CHV + i;
        if (v < nvalid) {
          const int vt = vl * VPL + v;   // voxel inside the tile
          const float2 av = unpack_bf16x2(araw[i]);
          float gv[HB_KP];
#pragma unroll
          for (int k4 = 0; k4 < HB_KP; k4 += 4) {
            const float4 q = *reinterpret_cast<const float4*>(&gs[vt][((k4 >> 2) ^ ((vt >> 1) & 3)) << 2]);
            gv[k4] = q.x; gv[k4 + 1] = q.y; gv[k4 + 2] = q.z; gv[k4 + 3] = q.w;
          }
          const float2 old = unpack_bf16x2(oraw[i]);
          float d0 = old.x, d1 = old.y;
#pragma unroll
          for (int k = 0; k < NK; ++k) {
            acc0[k] = fmaf(gv[k], av.x, acc0[k]);
            acc1[k] = fmaf(gv[k], av.y, acc1[k]);
            d0 = fmaf(gv[k], w0[k], d0);
            d1 = fmaf(gv[k], w1[k], d1);
          }
          *reinterpret_cast<uint32_t*>(drow + v * dstep) = pack_bf16x2(d0, d1);
        }
      }
    }
  }

  // dW[c][k]: reduce the voxel lanes through shared memory, one class at a time; one atomic per (c, k) per block
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    if (k < ncls) {   // (block-uniform)
      __syncthreads();
      red[vl][2 * p] = acc0[k];
      red[vl][2 * p + 1] = acc1[k];
      __syncthreads();
      for (int c = t; c < C; c += 256) {
        float s = 0.f;
#pragma unroll
        for (int l = 0; l < LANES; ++l) s += red[l][c];
        atomicAdd(dw + (long long)c * ldw + k, s);
      }
    }
  }
  // db[k]: the 128 threads with the same t / 128 share a class per j
#pragma unroll
  for (int j = 0; j < GJ; ++j) {
    float s = bsum[j];
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    const int k = gk0 + 2 * j;
    if ((t & 31) == 0 && k < ncls) atomicAdd(db + k, s);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Tensor-core variant for C = 64 (the three full-resolution heads: 0.9 of the 1.1 ms the class costs per step).  The CUDA-core
// kernel above is bound by its own instruction stream (ncu: 121 warp instructions per voxel and channel pair, 56 of them the
// FFMAs, issue slots 65 % busy at 4 warps per scheduler; profiles/r02_ncu_head_bwd.json).  Here one warp owns 16 voxels:
//     dA [16 vox x 64]   = G [16 vox x 16 cls] * W [16 cls x 64]        8 n-tiles of mma.m16n8k16, G as (hi + lo) bf16 pairs
//     dW^T [16 cls x 64] += G^T [16 cls x 16 vox] * A [16 vox x 64]     8 n-tiles, A fragments by ldmatrix.trans from shared memory
// G comes straight from the class-major fp32 gradient in both fragment layouts (the second read hits L1); splitting it into
// bf16 hi + lo keeps 16 mantissa bits (the weight gradient is checked to 1e-5).  W is rounded to bf16 — the forward GEMM's
// operand.  db in fp32 from the same registers.  The dA tile leaves through shared memory as 16-byte rows.
__device__ __forceinline__ void hb_ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

__device__ __forceinline__ void hb_split(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat16 xh = __float2bfloat16_rn(x), yh = __float2bfloat16_rn(y);
  hi = pack_bf16x2(__bfloat162float(xh), __bfloat162float(yh));
  lo = pack_bf16x2(x - __bfloat162float(xh), y - __bfloat162float(yh));
}

constexpr int HBM_WARPS = 4;
constexpr int HBM_LD = 72;   // bf16 elements per shared-memory row (144 B): conflict-free ldmatrix rows and fragment stores

__global__ void __launch_bounds__(HBM_WARPS * 32, 3) head_bwd_mma64_kernel(
    const float* __restrict__ g, const __nv_bfloat16* __restrict__ a, long long lda, const float* __restrict__ w,
    __nv_bfloat16* __restrict__ da, long long ldda, int accumulate, float* __restrict__ dw, int ldw, float* __restrict__ db,
    int B, long long S, int ncls) {
  constexpr int C = 64, NT = C / 8;
  __shared__ __align__(16) __nv_bfloat16 tile[HBM_WARPS][16 * HBM_LD];
  __shared__ float dws[C * 16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = lane >> 2, t = lane & 3;
  for (int i = threadIdx.x; i < C * 16; i += HBM_WARPS * 32) dws[i] = 0.f;
  __syncthreads();

  // W as B fragments of the dA product: (k = class 2t + {0,1} (+8), n = channel 8j + q)
  uint32_t wf[NT][2];
#pragma unroll
  for (int j = 0; j < NT; ++j) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int k = 2 * t + 8 * h;
      const float x = k < ncls ? __ldg(w + (long long)k * C + 8 * j + q) : 0.f;
      const float y = k + 1 < ncls ? __ldg(w + (long long)(k + 1) * C + 8 * j + q) : 0.f;
      wf[j][h] = pack_bf16x2(x, y);
    }
  }
  float e[NT][4];   // dW^T accumulators: (class q | q + 8, channel 8j + 2t + {0,1})
#pragma unroll
  for (int j = 0; j < NT; ++j) e[j][0] = e[j][1] = e[j][2] = e[j][3] = 0.f;
  float bs0 = 0.f, bs1 = 0.f;   // bias gradient of classes q and q + 8

  __nv_bfloat16* ts = tile[warp];
  const long long tiles_per_b = S / 16;   // S % 16 == 0 (checked by the launcher)
  const long long tiles = tiles_per_b * B;
  for (long long wt = (long long)blockIdx.x * HBM_WARPS + warp; wt < tiles; wt += (long long)gridDim.x * HBM_WARPS) {
    const int b = (int)(wt / tiles_per_b);
    const long long s0 = (wt - (long long)b * tiles_per_b) * 16;
    const long long row0 = (long long)b * S + s0;
    const float* gb = g + (long long)b * ncls * S + s0;
    // ---- global loads, all in flight together
    uint4 av[4];
#pragma unroll
    for (int it = 0; it < 4; ++it)
      av[it] = __ldg(reinterpret_cast<const uint4*>(a + (row0 + it * 4 + (lane >> 3)) * lda + (lane & 7) * 8));
    float ga[4][2];   // G fragment: [reg][pair along the class axis]
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int v = q + 8 * (r & 1), k = 2 * t + 8 * (r >> 1);
      ga[r][0] = k < ncls ? __ldg(gb + (long long)k * S + v) : 0.f;
      ga[r][1] = k + 1 < ncls ? __ldg(gb + (long long)(k + 1) * S + v) : 0.f;
    }
    float2 gt[4];     // G^T fragment: [reg] = pair along the voxel axis
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int k = q + 8 * (r & 1), v = 2 * t + 8 * (r >> 1);
      gt[r] = k < ncls ? __ldg(reinterpret_cast<const float2*>(gb + (long long)k * S + v)) : make_float2(0.f, 0.f);
    }
    float d[NT][4];   // dA accumulators: (voxel q | q + 8, channel 8j + 2t + {0,1})
    if (accumulate) {
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const float2 lo = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(da + (row0 + q) * ldda + 8 * j + 2 * t));
        const float2 hi = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(da + (row0 + q + 8) * ldda + 8 * j + 2 * t));
        d[j][0] = lo.x; d[j][1] = lo.y; d[j][2] = hi.x; d[j][3] = hi.y;
      }
    } else {
#pragma unroll
      for (int j = 0; j < NT; ++j) d[j][0] = d[j][1] = d[j][2] = d[j][3] = 0.f;
    }
    // ---- A tile to shared memory (row pitch 144 B)
#pragma unroll
    for (int it = 0; it < 4; ++it)
      *reinterpret_cast<uint4*>(ts + (it * 4 + (lane >> 3)) * HBM_LD + (lane & 7) * 8) = av[it];
    bs0 += (gt[0].x + gt[0].y) + (gt[2].x + gt[2].y);
    bs1 += (gt[1].x + gt[1].y) + (gt[3].x + gt[3].y);
    uint32_t gah[4], gal[4], gth[4], gtl[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      hb_split(ga[r][0], ga[r][1], gah[r], gal[r]);
      hb_split(gt[r].x, gt[r].y, gth[r], gtl[r]);
    }
    __syncwarp();
    // ---- dW^T += G^T A: B fragments of A by ldmatrix.trans (lane addresses row lane & 15, column block lane >> 4)
    const uint32_t abase = smem_u32(ts + (lane & 15) * HBM_LD + (lane >> 4) * 8);
#pragma unroll
    for (int j = 0; j < NT; j += 2) {
      uint32_t bf[4];
      hb_ldmatrix_x4_trans(bf, abase + j * 16);
      mma_bf16_16816(e[j], gth, bf[0], bf[1]);
      mma_bf16_16816(e[j], gtl, bf[0], bf[1]);
      mma_bf16_16816(e[j + 1], gth, bf[2], bf[3]);
      mma_bf16_16816(e[j + 1], gtl, bf[2], bf[3]);
    }
    // ---- dA (+)= G W
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      mma_bf16_16816(d[j], gah, wf[j][0], wf[j][1]);
      mma_bf16_16816(d[j], gal, wf[j][0], wf[j][1]);
    }
    __syncwarp();   // every lane is done reading the A tile
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      *reinterpret_cast<uint32_t*>(ts + q * HBM_LD + 8 * j + 2 * t) = pack_bf16x2(d[j][0], d[j][1]);
      *reinterpret_cast<uint32_t*>(ts + (q + 8) * HBM_LD + 8 * j + 2 * t) = pack_bf16x2(d[j][2], d[j][3]);
    }
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int r = it * 4 + (lane >> 3);
      *reinterpret_cast<uint4*>(da + (row0 + r) * ldda + (lane & 7) * 8) =
          *reinterpret_cast<const uint4*>(ts + r * HBM_LD + (lane & 7) * 8);
    }
    __syncwarp();   // the tile buffer is free for the next iteration
  }

  // ---- dW: warps meet in shared memory, then one global atomic per (channel, class) per block
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    const int c = 8 * j + 2 * t;
    if (q < ncls) { atomicAdd(&dws[c * 16 + q], e[j][0]); atomicAdd(&dws[(c + 1) * 16 + q], e[j][1]); }
    if (q + 8 < ncls) { atomicAdd(&dws[c * 16 + q + 8], e[j][2]); atomicAdd(&dws[(c + 1) * 16 + q + 8], e[j][3]); }
  }
  bs0 += __shfl_xor_sync(0xffffffffu, bs0, 1);
  bs0 += __shfl_xor_sync(0xffffffffu, bs0, 2);
  bs1 += __shfl_xor_sync(0xffffffffu, bs1, 1);
  bs1 += __shfl_xor_sync(0xffffffffu, bs1, 2);
  if (t == 0) {
    if (q < ncls) atomicAdd(db + q, bs0);
    if (q + 8 < ncls) atomicAdd(db + q + 8, bs1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * 16; i += HBM_WARPS * 32) {
    const int c = i >> 4, k = i & 15;
    if (k < ncls) atomicAdd(dw + (long long)c * ldw + k, dws[i]);
  }
}

template <int C>
static void launch_head_bwd(unsigned grid, cudaStream_t st, const float* g, const __nv_bfloat16* a, long long lda,
                            const float* w, __nv_bfloat16* da, long long ldda, int accumulate, float* dw, int ldw,
                            float* db, int B, long long S, int ncls) {
  if (ncls <= 14) head_bwd_kernel<C, 14><<<grid, 256, 0, st>>>(g, a, lda, w, da, ldda, accumulate, dw, ldw, db, B, S, ncls);
  else head_bwd_kernel<C, 16><<<grid, 256, 0, st>>>(g, a, lda, w, da, ldda, accumulate, dw, ldw, db, B, S, ncls);
}

}  // namespace ctu

extern "C" int ctu_head_bwd(const float* g, const void* a, long long lda, const float* w, void* da, long long ldda,
                            int accumulate, float* dw, int ldw, float* db, int B, long long S, int C, int ncls,
                            void* stream) {
  using namespace ctu;
  if (!g || !a || !w || !da || !dw || !db || B <= 0 || S <= 0) return CTU_E_BADARG;
  if (ncls <= 0 || ncls > HB_KP || ldw < ncls || lda < C || ldda < C || (lda & 1) || (ldda & 1)) return CTU_E_BADARG;
  if ((reinterpret_cast<uintptr_t>(a) & 3) || (reinterpret_cast<uintptr_t>(da) & 3)) return CTU_E_BADARG;
  const long long tiles = ((S + HB_TV - 1) / HB_TV) * B;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long grid = (long long)sms * 2;   // one wave of resident blocks: every block ends with C x n_cls atomics
  if (grid > tiles) grid = tiles;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const __nv_bfloat16* ap = reinterpret_cast<const __nv_bfloat16*>(a);
  __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(da);
  static const bool use_mma = [] { const char* e = getenv("CTU_HEAD_BWD_MMA"); return e ? atoi(e) != 0 : true; }();
  if (use_mma && C == 64 && S % 16 == 0 && lda % 8 == 0 && ldda % 8 == 0 && !(reinterpret_cast<uintptr_t>(a) & 15) &&
      !(reinterpret_cast<uintptr_t>(da) & 15) && !(reinterpret_cast<uintptr_t>(g) & 7)) {
    const long long wtiles = (S / 16) * B;
    long long mgrid = (long long)sms * 3;
    if (mgrid * HBM_WARPS > wtiles) mgrid = (wtiles + HBM_WARPS - 1) / HBM_WARPS;
    head_bwd_mma64_kernel<<<(unsigned)mgrid, HBM_WARPS * 32, 0, st>>>(g, ap, lda, w, dp, ldda, accumulate, dw, ldw, db, B, S, ncls);
    count_launch();
    return (int)cudaGetLastError();
  }
  switch (C) {
    case 64: launch_head_bwd<64>((unsigned)grid, st, g, ap, lda, w, dp, ldda, accumulate, dw, ldw, db, B, S, ncls); break;
    case 128: launch_head_bwd<128>((unsigned)grid, st, g, ap, lda, w, dp, ldda, accumulate, dw, ldw, db, B, S, ncls); break;
    case 256: launch_head_bwd<256>((unsigned)grid, st, g, ap, lda, w, dp, ldda, accumulate, dw, ldw, db, B, S, ncls); break;
    default: return CTU_E_UNSUPPORTED;
  }
  count_launch();
  return (int)cudaGetLastError();
}
