// Sliding-window Gaussian blend (sm_100a, HBM-bound): the accumulation and normalisation that
// trainer_CTUNet.py:541-549 / trainer_CUNet.py:388-392 perform with four torch ops per window per head.
//   acc[c, x0+i, y0+j, z0+k] += imp[i,j,k] * logits[c,i,j,k]        (one launch per window, both heads)
//   cnt[x0+i, y0+j, z0+k]    += imp[i,j,k]                           (geometry-only, 1 channel instead of 14)
//   out = acc / cnt
// Products and sums are issued as separate IEEE roundings (no FMA contraction) in the reference's window
// order, so given the same logits the blended volume is bit-identical to the reference's.
#include "common.cuh"
#include "../../include/ctunet_b200.h"
#include "host_util.h"

namespace ctu {

template <bool VEC>
__global__ void __launch_bounds__(256) blend_accumulate_kernel(const float* __restrict__ l0, const float* __restrict__ l1,
                                                               const float* __restrict__ imp, float* __restrict__ a0,
                                                               float* __restrict__ a1, int C, int r3, int r2, int r1,
                                                               int X, int Y, int Z, int x0, int y0, int z0) {
  const float* logits = blockIdx.y == 0 ? l0 : l1;
  float* acc = blockIdx.y == 0 ? a0 : a1;
  const long long win_vox = (long long)r3 * r2 * r1;
  constexpr int V = VEC ? 4 : 1;
  const long long total = (long long)C * win_vox / V;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long e = i * V;
    const int c = (int)(e / win_vox);
    long long s = e - (long long)c * win_vox;
    const long long sp = s;
    const int k = (int)(s % r1); s /= r1;
    const int j = (int)(s % r2);
    const int ii = (int)(s / r2);
    float* ap = acc + (((long long)c * X + (x0 + ii)) * Y + (y0 + j)) * Z + (z0 + k);
    if constexpr (VEC) {
      const float4 lv = *reinterpret_cast<const float4*>(logits + e);
      const float4 w = *reinterpret_cast<const float4*>(imp + sp);
      float4 av = *reinterpret_cast<float4*>(ap);
      av.x = __fadd_rn(av.x, __fmul_rn(w.x, lv.x));
      av.y = __fadd_rn(av.y, __fmul_rn(w.y, lv.y));
      av.z = __fadd_rn(av.z, __fmul_rn(w.z, lv.z));
      av.w = __fadd_rn(av.w, __fmul_rn(w.w, lv.w));
      *reinterpret_cast<float4*>(ap) = av;
    } else {
      *ap = __fadd_rn(*ap, __fmul_rn(imp[sp], logits[e]));
    }
  }
}

__global__ void __launch_bounds__(256) blend_count_kernel(const float* __restrict__ imp, float* __restrict__ cnt, int r3,
                                                          int r2, int r1, int X, int Y, int Z, int x0, int y0, int z0) {
  const long long total = (long long)r3 * r2 * r1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long s = i;
    const int k = (int)(s % r1); s /= r1;
    const int j = (int)(s % r2);
    const int ii = (int)(s / r2);
    float* cp = cnt + ((long long)(x0 + ii) * Y + (y0 + j)) * Z + (z0 + k);
    *cp = __fadd_rn(*cp, imp[i]);
  }
}

__global__ void __launch_bounds__(256) blend_normalize_kernel(const float* __restrict__ acc, const float* __restrict__ cnt,
                                                              float* __restrict__ out, int C, long long vox) {
  const long long total = (long long)C * vox;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    out[i] = __fdiv_rn(acc[i], cnt[i % vox]);
  }
}

// 16-byte vectorised variant: a thread owns four consecutive voxels, loads their counts once and walks the channels
// (coalesced per channel; the count map is read once instead of C times)
__global__ void __launch_bounds__(256) blend_normalize_vec_kernel(const float4* __restrict__ acc, const float4* __restrict__ cnt,
                                                                  float4* __restrict__ out, int C, long long vox4) {
  constexpr int G = 7;  // channels in flight per thread: all G loads are issued before the first division
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < vox4; v += (long long)gridDim.x * blockDim.x) {
    const float4 cn = cnt[v];
    for (int c0 = 0; c0 < C; c0 += G) {
      float4 a[G];
#pragma unroll
      for (int i = 0; i < G; ++i)
        if (c0 + i < C) a[i] = __ldcs(acc + (long long)(c0 + i) * vox4 + v);
#pragma unroll
      for (int i = 0; i < G; ++i) {
        if (c0 + i < C) {
          float4 o;
          o.x = __fdiv_rn(a[i].x, cn.x); o.y = __fdiv_rn(a[i].y, cn.y); o.z = __fdiv_rn(a[i].z, cn.z);
          o.w = __fdiv_rn(a[i].w, cn.w);
          __stcs(out + (long long)(c0 + i) * vox4 + v, o);
        }
      }
    }
  }
}

}  // namespace ctu

using namespace ctu;

static int blend_grid(long long items) {
  long long b = (items + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  if (b < 1) b = 1;
  return (int)b;
}

// logits0/1: fp32 [C][r3][r2][r1] of ONE window (head 0 / head 1; logits1/acc1 may be NULL for the one-head
// variant, trainer_CUNet.py:376); imp: fp32 [r3][r2][r1]; acc: fp32 [C][X][Y][Z] of one batch item.
extern "C" int ctu_blend_accumulate(const float* logits0, const float* logits1, const float* imp, float* acc0,
                                    float* acc1, int C, int r3, int r2, int r1, int X, int Y, int Z, int x0, int y0,
                                    int z0, void* stream) {
  if (!logits0 || !imp || !acc0 || (logits1 && !acc1)) return CTU_E_BADARG;
  if (x0 < 0 || y0 < 0 || z0 < 0 || x0 + r3 > X || y0 + r2 > Y || z0 + r1 > Z) return CTU_E_BADARG;
  const int heads = logits1 ? 2 : 1;
  const bool vec = (r1 % 4 == 0) && (Z % 4 == 0) && (z0 % 4 == 0) &&
                   ((reinterpret_cast<uintptr_t>(logits0) | reinterpret_cast<uintptr_t>(logits1) |
                     reinterpret_cast<uintptr_t>(imp) | reinterpret_cast<uintptr_t>(acc0) |
                     reinterpret_cast<uintptr_t>(acc1)) % 16 == 0);
  const long long items = (long long)C * r3 * r2 * r1 / (vec ? 4 : 1);
  dim3 grid(blend_grid(items), heads);
  if (vec)
    blend_accumulate_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(logits0, logits1, imp, acc0, acc1, C, r3, r2, r1,
                                                                          X, Y, Z, x0, y0, z0);
  else
    blend_accumulate_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(logits0, logits1, imp, acc0, acc1, C, r3, r2,
                                                                           r1, X, Y, Z, x0, y0, z0);
  count_launch();
  return (int)cudaGetLastError();
}

extern "C" int ctu_blend_count(const float* imp, float* cnt, int r3, int r2, int r1, int X, int Y, int Z, int x0, int y0,
                               int z0, void* stream) {
  if (!imp || !cnt) return CTU_E_BADARG;
  if (x0 < 0 || y0 < 0 || z0 < 0 || x0 + r3 > X || y0 + r2 > Y || z0 + r1 > Z) return CTU_E_BADARG;
  blend_count_kernel<<<blend_grid((long long)r3 * r2 * r1), 256, 0, (cudaStream_t)stream>>>(imp, cnt, r3, r2, r1, X, Y, Z,
                                                                                           x0, y0, z0);
  count_launch();
  return (int)cudaGetLastError();
}

// out[c][v] = acc[c][v] / cnt[v]  (out may alias acc)
extern "C" int ctu_blend_normalize(const float* acc, const float* cnt, float* out, int C, long long vox, void* stream) {
  if (!acc || !cnt || !out) return CTU_E_BADARG;
  const bool vec = (vox % 4 == 0) && ((reinterpret_cast<uintptr_t>(acc) | reinterpret_cast<uintptr_t>(cnt) |
                                       reinterpret_cast<uintptr_t>(out)) % 16 == 0);
  if (vec)
    blend_normalize_vec_kernel<<<blend_grid(vox / 4), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(acc), reinterpret_cast<const float4*>(cnt), reinterpret_cast<float4*>(out), C, vox / 4);
  else
    blend_normalize_kernel<<<blend_grid((long long)C * vox), 256, 0, (cudaStream_t)stream>>>(acc, cnt, out, C, vox);
  count_launch();
  return (int)cudaGetLastError();
}
