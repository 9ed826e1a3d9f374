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
#include "common.cuh"
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
  switch (C) {
    case 64: launch_head_bwd<64>((unsigned)grid, st, g, ap, lda, w, dp, ldda, accumulate, dw, ldw, db, B, S, ncls); break;
    case 128: launch_head_bwd<128>((unsigned)grid, st, g, ap, lda, w, dp, ldda, accumulate, dw, ldw, db, B, S, ncls); break;
    case 256: launch_head_bwd<256>((unsigned)grid, st, g, ap, lda, w, dp, ldda, accumulate, dw, ldw, db, B, S, ncls); break;
    default: return CTU_E_UNSUPPORTED;
  }
  count_launch();
  return (int)cudaGetLastError();
}
