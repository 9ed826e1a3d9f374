// HBM-bound kernels of the CTUNet path (sm_100a): InstanceNorm statistics / apply (+LeakyReLU, +residual),
// LayerNorm, ViT patchify+LayerNorm, the binary cross-weight fusion, strided sub-sampling.
// All are 16-byte vectorised, coalesced along the channel axis of channels-last bf16 tensors and sized as
// grid-stride loops over a multiple of the SM count.
#include "common.cuh"
#include "../../include/ctunet_b200.h"
#include "host_util.h"

namespace ctu {

static int num_sms() {
  static int n = [] {
    int dev = 0, v = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v;
  }();
  return n;
}

__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 t;
  t = unpack_bf16x2(u.x); f[0] = t.x; f[1] = t.y;
  t = unpack_bf16x2(u.y); f[2] = t.x; f[3] = t.y;
  t = unpack_bf16x2(u.z); f[4] = t.x; f[5] = t.y;
  t = unpack_bf16x2(u.w); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  float2 t;
  t = unpack_bf16x2(u.x); f[0] = t.x; f[1] = t.y;
  t = unpack_bf16x2(u.y); f[2] = t.x; f[3] = t.y;
  t = unpack_bf16x2(u.z); f[4] = t.x; f[5] = t.y;
  t = unpack_bf16x2(u.w); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&f)[8]) {
  *reinterpret_cast<uint4*>(p) =
      make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}

// ------------------------------------------------------------------------------------------------
// InstanceNorm statistics: stats[b][c] += (sum x, sum x^2) over the S voxels of instance b (fp64 atomics).
// Same accumulator format as the fused epilogue of the tensor-core kernel.
__global__ void __launch_bounds__(256) in_stats_kernel(const __nv_bfloat16* __restrict__ x, int ldx, long long S, int C,
                                                       double* __restrict__ stats, int stats_ld) {
  __shared__ float red[2][256][8];
  const int tpr = C / 8;            // threads per row
  const int rpb = 256 / tpr;        // rows per block iteration
  const int cv = threadIdx.x % tpr;
  const int rl = threadIdx.x / tpr;
  const int b = blockIdx.y;
  const __nv_bfloat16* xb = x + (long long)b * S * ldx + cv * 8;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  // four rows in flight per thread, every load issued before the first use
  constexpr int UNR = 4;
  const long long stride = (long long)gridDim.x * rpb;
  for (long long r0 = (long long)blockIdx.x * rpb + rl; r0 < S; r0 += stride * UNR) {
    uint4 v[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long long r = r0 + u * stride;
      v[u] = r < S ? *reinterpret_cast<const uint4*>(xb + r * ldx) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      float f[8];
      unpack8(v[u], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) { s[j] += f[j]; q[j] += f[j] * f[j]; }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { red[0][threadIdx.x][j] = s[j]; red[1][threadIdx.x][j] = q[j]; }
  __syncthreads();
  // one (sum | sum of squares, channel) output per thread over the rpb row slots, then ONE fp64 atomic each
  for (int o = threadIdx.x; o < 2 * C; o += 256) {
    const int a = o / C, c = o - a * C;
    const float* rp = &red[a][c >> 3][c & 7];
    float acc = 0.f;
    for (int i = 0; i < rpb; ++i) acc += rp[i * tpr * 8];
    atomicAdd(stats + ((long long)b * stats_ld + c) * 2 + a, (double)acc);
  }
}

// ------------------------------------------------------------------------------------------------
// y = act( IN(x) [+ IN(r) | + r] ), IN(v) = (v - mean) * rsqrt(var + eps) from fp64 (sum, sumsq) accumulators.
// Covers resnet.py:110-124 (gn + lrelu, gn + residual + lrelu) and hybrid_CTUNet.py:95-104.
__device__ __forceinline__ void in_coeffs(const double* st, double inv_n, float eps, float& scale, float& shift) {
  const double mean = st[0] * inv_n;
  double var = st[1] * inv_n - mean * mean;
  var = var < 0.0 ? 0.0 : var;
  const double rstd = rsqrt(var + (double)eps);
  scale = (float)rstd;
  shift = (float)(-mean * rstd);
}

// A thread owns 8 channels of a row and walks rows with a grid stride, IN_UNR rows per iteration with every load issued
// before the first use; the per-channel (scale, shift) pairs are computed once per block (one fp64 rsqrt per channel,
// spread over the threads) and staged in shared memory.  `rev` walks each instance from its last row to its first:
// the kernel runs right after the GEMM that wrote x front to back, so the tail of x is what the L2 still holds.
template <int RES, int IN_UNR>  // RES: 0 none, 1 raw residual, 2 normalised residual
__global__ void __launch_bounds__(256) in_apply_kernel(const __nv_bfloat16* __restrict__ x, int ldx,
                                                       const double* __restrict__ xstats, int xs_ld,
                                                       const __nv_bfloat16* __restrict__ res, int ldr,
                                                       const double* __restrict__ rstats, int rs_ld,
                                                       __nv_bfloat16* __restrict__ out, int ldo, long long S, int C,
                                                       float eps, int act, float slope, int rev) {
  extern __shared__ __align__(16) float in_sm[];  // [4][C]: scale, shift of x, then of the residual
  const int tpr = C / 8;
  const int rpb = 256 / tpr;
  const int cv = threadIdx.x % tpr;
  const int rl = threadIdx.x / tpr;
  const int b = blockIdx.y;
  const double inv_n = 1.0 / (double)S;
  for (int c = threadIdx.x; c < C; c += 256) {
    in_coeffs(xstats + ((long long)b * xs_ld + c) * 2, inv_n, eps, in_sm[c], in_sm[C + c]);
    if (RES == 2) in_coeffs(rstats + ((long long)b * rs_ld + c) * 2, inv_n, eps, in_sm[2 * C + c], in_sm[3 * C + c]);
  }
  __syncthreads();
  float sc[8], sh[8], rsc[8], rsh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = in_sm[cv * 8 + j]; sh[j] = in_sm[C + cv * 8 + j];
    if (RES == 2) { rsc[j] = in_sm[2 * C + cv * 8 + j]; rsh[j] = in_sm[3 * C + cv * 8 + j]; }
  }
  const __nv_bfloat16* xb = x + (long long)b * S * ldx + cv * 8;
  const __nv_bfloat16* rb = RES ? res + (long long)b * S * ldr + cv * 8 : nullptr;
  __nv_bfloat16* ob = out + (long long)b * S * ldo + cv * 8;
  const long long stride = (long long)gridDim.x * rpb;
  const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
  for (long long r0 = (long long)blockIdx.x * rpb + rl; r0 < S; r0 += stride * IN_UNR) {
    uint4 xq[IN_UNR], rq[IN_UNR];
    long long rows[IN_UNR];
#pragma unroll
    for (int u = 0; u < IN_UNR; ++u) {
      const long long r = r0 + u * stride;
      const bool ok = r < S;
      rows[u] = ok ? (rev ? S - 1 - r : r) : -1;
      xq[u] = ok ? *reinterpret_cast<const uint4*>(xb + rows[u] * ldx) : z4;
      if (RES) rq[u] = ok ? *reinterpret_cast<const uint4*>(rb + rows[u] * ldr) : z4;
    }
#pragma unroll
    for (int u = 0; u < IN_UNR; ++u) {
      if (rows[u] < 0) continue;
      float f[8], g[8];
      unpack8(xq[u], f);
      if (RES) unpack8(rq[u], g);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v = fmaf(f[j], sc[j], sh[j]);
        if (RES == 1) v += g[j];
        if (RES == 2) v += fmaf(g[j], rsc[j], rsh[j]);
        if (act) v = v > 0.f ? v : v * slope;
        f[j] = v;
      }
      store8(ob + rows[u] * ldo, f);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over the last dim (vit.py:35,55,116,118; hybrid_CTUNet.py:456,518,630-631): one warp per row,
// the row lives in registers, exact two-pass mean / variance in fp32.
template <int VPL, typename TIN, typename TOUT>
__global__ void __launch_bounds__(256) layernorm_kernel(const TIN* __restrict__ x, long long ldx,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        const float* __restrict__ add, long long add_rows,
                                                        TOUT* __restrict__ out, long long ldo, long long M, int C,
                                                        float eps) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  const int nvec = C / 8;
  float v[VPL][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vi = lane + 32 * i;
    if (vi < nvec) {
      if constexpr (sizeof(TIN) == 2) {
        load8(reinterpret_cast<const __nv_bfloat16*>(x) + row * ldx + vi * 8, v[i]);
      } else {
        const float4 a = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x) + row * ldx + vi * 8);
        const float4 b = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x) + row * ldx + vi * 8 + 4);
        v[i][0] = a.x; v[i][1] = a.y; v[i][2] = a.z; v[i][3] = a.w;
        v[i][4] = b.x; v[i][5] = b.y; v[i][6] = b.z; v[i][7] = b.w;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += v[i][j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[i][j] = 0.f;
    }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / (float)C;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    if (lane + 32 * i < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = v[i][j] - mean; sq += d * d; }
    }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / (float)C + eps);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vi = lane + 32 * i;
    if (vi < nvec) {
      float o8[8];
      const float4 g0 = *reinterpret_cast<const float4*>(gamma + vi * 8), g1 = *reinterpret_cast<const float4*>(gamma + vi * 8 + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(beta + vi * 8), b1 = *reinterpret_cast<const float4*>(beta + vi * 8 + 4);
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) o8[j] = (v[i][j] - mean) * rstd * gg[j] + bb[j];
      if (add != nullptr) {
        const float* ap = add + (row % add_rows) * C + vi * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) o8[j] += ap[j];
      }
      if constexpr (sizeof(TOUT) == 2) {
        store8(reinterpret_cast<__nv_bfloat16*>(out) + row * ldo + vi * 8, o8);
      } else {
        float* op = reinterpret_cast<float*>(out) + row * ldo + vi * 8;
        *reinterpret_cast<float4*>(op) = make_float4(o8[0], o8[1], o8[2], o8[3]);
        *reinterpret_cast<float4*>(op + 4) = make_float4(o8[4], o8[5], o8[6], o8[7]);
      }
    }
  }
}

// Narrow rows (C = 64 / 128 / 256, bf16 in and out — the per-voxel LayerNorms of the cross-weight fusion and the
// window-attention blocks at 48·48·96 and 24·24·48): LPR = C/8 lanes own a row (one 16-byte vector each), so a warp
// covers 32/LPR rows per pass with every lane loading; rows are walked with a grid stride, two rows in flight.
template <int LPR>
__global__ void __launch_bounds__(256) layernorm_narrow_kernel(const __nv_bfloat16* __restrict__ x, long long ldx,
                                                               const float* __restrict__ gamma,
                                                               const float* __restrict__ beta,
                                                               __nv_bfloat16* __restrict__ out, long long ldo,
                                                               long long M, float eps) {
  constexpr int C = LPR * 8;
  constexpr int RPB = 256 / LPR;
  constexpr int UNR = 2;
  const int sub = threadIdx.x % LPR, rl = threadIdx.x / LPR;
  float gg[8], bb[8];
  {
    const float4 g0 = *reinterpret_cast<const float4*>(gamma + sub * 8), g1 = *reinterpret_cast<const float4*>(gamma + sub * 8 + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(beta + sub * 8), b1 = *reinterpret_cast<const float4*>(beta + sub * 8 + 4);
    gg[0] = g0.x; gg[1] = g0.y; gg[2] = g0.z; gg[3] = g0.w; gg[4] = g1.x; gg[5] = g1.y; gg[6] = g1.z; gg[7] = g1.w;
    bb[0] = b0.x; bb[1] = b0.y; bb[2] = b0.z; bb[3] = b0.w; bb[4] = b1.x; bb[5] = b1.y; bb[6] = b1.z; bb[7] = b1.w;
  }
  const float invC = 1.f / (float)C;
  const long long stride = (long long)gridDim.x * RPB;
  // block-uniform trip count: every lane takes part in the shuffles, out-of-range rows are masked
  const long long iters = (M + stride * UNR - 1) / (stride * UNR);
  for (long long it = 0; it < iters; ++it) {
    uint4 q[UNR];
    long long rows[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long long r = ((it * UNR + u) * gridDim.x + blockIdx.x) * RPB + rl;
      rows[u] = r < M ? r : -1;
      q[u] = r < M ? *reinterpret_cast<const uint4*>(x + r * ldx + sub * 8) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      float v[8];
      unpack8(q[u], v);
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += v[j];
#pragma unroll
      for (int o = LPR / 2; o >= 1; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float mean = sum * invC;
      float sq = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = v[j] - mean; sq += d * d; }
#pragma unroll
      for (int o = LPR / 2; o >= 1; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
      const float rstd = rsqrtf(sq * invC + eps);
      if (rows[u] >= 0) {
        float o8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o8[j] = (v[j] - mean) * rstd * gg[j] + bb[j];
        store8(out + rows[u] * ldo + sub * 8, o8);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// ViT patch embedding front end (vit.py:115-116): 'b c (h p1) (w p2) (f pf) -> b (h w f) (p1 p2 pf c)', c = 1,
// fused with LayerNorm(p1*p2*pf).  One block per token, thread = (p1, p2), PF contiguous floats per thread.
template <int PF>
__global__ void __launch_bounds__(256) patchify_ln_kernel(const float* __restrict__ img, int X, int Y, int Z,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          __nv_bfloat16* __restrict__ out, float eps) {
  __shared__ float red[8];
  __shared__ float bc;
  const int nh = X / 16, nw = Y / 16, nf = Z / PF;
  int tok = blockIdx.x;
  const int f = tok % nf; tok /= nf;
  const int w = tok % nw; tok /= nw;
  const int h = tok % nh;
  const int b = tok / nh;
  const int p1 = threadIdx.x / 16, p2 = threadIdx.x % 16;
  const float* src = img + (((long long)b * X + (h * 16 + p1)) * Y + (w * 16 + p2)) * Z + f * PF;
  float v[PF];
#pragma unroll
  for (int j = 0; j < PF; j += 4) {
    const float4 t = *reinterpret_cast<const float4*>(src + j);
    v[j] = t.x; v[j + 1] = t.y; v[j + 2] = t.z; v[j + 3] = t.w;
  }
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const float n = 256.f * PF;
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < PF; ++j) s += v[j];
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) red[wp] = s;
  __syncthreads();
  if (threadIdx.x == 0) { float t = 0; for (int i = 0; i < 8; ++i) t += red[i]; bc = t / n; }
  __syncthreads();
  const float mean = bc;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < PF; ++j) { const float d = v[j] - mean; q += d * d; }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  __syncthreads();
  if (lane == 0) red[wp] = q;
  __syncthreads();
  if (threadIdx.x == 0) { float t = 0; for (int i = 0; i < 8; ++i) t += red[i]; bc = rsqrtf(t / n + eps); }
  __syncthreads();
  const float rstd = bc;
  const int e0 = threadIdx.x * PF;
  __nv_bfloat16* dst = out + (long long)blockIdx.x * (256 * PF) + e0;
#pragma unroll
  for (int j = 0; j < PF; j += 8) {
    float o8[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) o8[k] = (v[j + k] - mean) * rstd * gamma[e0 + j + k] + beta[e0 + j + k];
    store8(dst + j, o8);
  }
}

// ------------------------------------------------------------------------------------------------
// Binary cross-weight fusion, the per-token middle of pixelweight_attention (hybrid_CTUNet.py:658-665):
//   d1 = <q2,k1> * scale, d2 = <q1,k2> * scale per (token, head of 32); a = softmax([d1,d2]); out = a1*v1 + a2*v2.
// qkv rows are [q | k | v] of width 3C (output of the to_qkv GEMMs); 4 lanes cover one head (8 channels each).
__global__ void __launch_bounds__(256) pwa_fuse_kernel(const __nv_bfloat16* __restrict__ qkv1,
                                                       const __nv_bfloat16* __restrict__ qkv2,
                                                       __nv_bfloat16* __restrict__ out, long long T, int C, float scale) {
  const int tpr = C / 8;
  const long long total = T * tpr;
  // block-uniform loop bound: every lane takes part in the shuffles, out-of-range lanes recompute the last item
  for (long long base = (long long)blockIdx.x * blockDim.x; base < total; base += (long long)gridDim.x * blockDim.x) {
    long long i = base + threadIdx.x;
    const bool ok = i < total;
    if (!ok) i = total - 1;
    const long long t = i / tpr;
    const int cv = (int)(i - t * tpr);
    const __nv_bfloat16* r1 = qkv1 + t * 3 * C + cv * 8;
    const __nv_bfloat16* r2 = qkv2 + t * 3 * C + cv * 8;
    float q1[8], k1[8], v1[8], q2[8], k2[8], v2[8];
    load8(r1, q1); load8(r1 + C, k1); load8(r1 + 2 * C, v1);
    load8(r2, q2); load8(r2 + C, k2); load8(r2 + 2 * C, v2);
    float d1 = 0.f, d2 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { d1 += q2[j] * k1[j]; d2 += q1[j] * k2[j]; }
    // 4 consecutive lanes hold one head (tpr is a multiple of 4, so groups never straddle tokens)
    d1 += __shfl_xor_sync(0xffffffffu, d1, 1); d2 += __shfl_xor_sync(0xffffffffu, d2, 1);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 2); d2 += __shfl_xor_sync(0xffffffffu, d2, 2);
    d1 *= scale; d2 *= scale;
    const float m = fmaxf(d1, d2);
    const float e1 = __expf(d1 - m), e2 = __expf(d2 - m);
    const float inv = 1.f / (e1 + e2);
    const float a1 = e1 * inv, a2 = e2 * inv;
    float o8[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o8[j] = a1 * v1[j] + a2 * v2[j];
    if (!ok) continue;
    store8(out + t * C + cv * 8, o8);
  }
}

// ------------------------------------------------------------------------------------------------
// Strided sub-sampling of a channels-last tensor: out[b,x,y,z,:] = in[b, x*s3, y*s2, z*s1, :]
// (the input side of the stride-2 1x1x1 downsample convs, resnet.py:197, and of stride-2 3x3x3 convs computed at
// full resolution).
__global__ void __launch_bounds__(256) subsample_kernel(const __nv_bfloat16* __restrict__ in, int ldi, int I1, int I2,
                                                        int I3, __nv_bfloat16* __restrict__ out, int ldo, int O1,
                                                        int O2, int O3, int s1, int s2, int s3, int C, int B) {
  const int tpr = C / 8;
  const long long total = (long long)B * O3 * O2 * O1 * tpr;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long r = i / tpr;
    const int cv = (int)(i - r * tpr);
    const long long orow = r;
    const int z = (int)(r % O1); r /= O1;
    const int y = (int)(r % O2); r /= O2;
    const int x = (int)(r % O3);
    const int b = (int)(r / O3);
    const long long irow = (((long long)b * I3 + x * s3) * I2 + y * s2) * I1 + z * s1;
    *reinterpret_cast<uint4*>(out + orow * ldo + cv * 8) = *reinterpret_cast<const uint4*>(in + irow * ldi + cv * 8);
  }
}

}  // namespace ctu

using namespace ctu;

static inline int grid_for(long long work_items, int per_block) {
  long long blocks = (work_items + per_block - 1) / per_block;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

extern "C" int ctu_in_stats(const void* x, int ldx, int B, long long S, int C, double* stats, int stats_ld, void* stream) {
  if (!x || !stats || C % 8 || C > 2048 || (2048 % C) || ldx % 8) return CTU_E_BADARG;
  const int rpb = 256 / (C / 8);
  int gx = grid_for(S, rpb * 16);
  const int cap = (num_sms() * 8 + B - 1) / B;
  if (gx > cap) gx = cap;
  in_stats_kernel<<<dim3(gx, B), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, ldx, S, C, stats, stats_ld);
  count_launch();
  return (int)cudaGetLastError();
}

extern "C" int ctu_in_apply(const void* x, int ldx, const double* xstats, int xs_ld, const void* res, int ldr,
                            const double* rstats, int rs_ld, void* out, int ldo, int B, long long S, int C, float eps,
                            int act, float slope, void* stream) {
  if (!x || !xstats || !out || C % 8 || C > 2048 || (2048 % C) || ldx % 8 || ldo % 8) return CTU_E_BADARG;
  if (res && ldr % 8) return CTU_E_BADARG;
  const int rpb = 256 / (C / 8);
  int gx = grid_for(S, rpb * 8);
  const int cap = (num_sms() * 8 + B - 1) / B;
  if (gx > cap) gx = cap;
  dim3 grid(gx, B);
  cudaStream_t st = (cudaStream_t)stream;
  const __nv_bfloat16* xp = (const __nv_bfloat16*)x;
  const __nv_bfloat16* rp = (const __nv_bfloat16*)res;
  __nv_bfloat16* op = (__nv_bfloat16*)out;
  const size_t smem = (size_t)(rstats ? 4 : 2) * C * sizeof(float);
  if (!res)
    in_apply_kernel<0, 4><<<grid, 256, smem, st>>>(xp, ldx, xstats, xs_ld, nullptr, 0, nullptr, 0, op, ldo, S, C, eps, act, slope, 1);
  else if (!rstats)
    in_apply_kernel<1, 2><<<grid, 256, smem, st>>>(xp, ldx, xstats, xs_ld, rp, ldr, nullptr, 0, op, ldo, S, C, eps, act, slope, 1);
  else
    in_apply_kernel<2, 2><<<grid, 256, smem, st>>>(xp, ldx, xstats, xs_ld, rp, ldr, rstats, rs_ld, op, ldo, S, C, eps, act, slope, 1);
  count_launch();
  return (int)cudaGetLastError();
}

template <typename TIN, typename TOUT>
static int launch_ln(const void* x, long long ldx, const float* g, const float* b, const float* add, long long add_rows,
                     void* out, long long ldo, long long M, int C, float eps, cudaStream_t st) {
  const int nvec = C / 8;
  const int vpl = (nvec + 31) / 32;
  const unsigned grid = (unsigned)((M + 7) / 8);
#define CTU_LN(V) layernorm_kernel<V, TIN, TOUT><<<grid, 256, 0, st>>>((const TIN*)x, ldx, g, b, add, add_rows, (TOUT*)out, ldo, M, C, eps)
  if (vpl <= 1) CTU_LN(1);
  else if (vpl <= 2) CTU_LN(2);
  else if (vpl <= 4) CTU_LN(4);
  else if (vpl <= 8) CTU_LN(8);
  else return CTU_E_UNSUPPORTED;
#undef CTU_LN
  count_launch();
  return (int)cudaGetLastError();
}

extern "C" int ctu_layernorm(const void* x, int x_is_f32, long long ldx, const float* gamma, const float* beta,
                             const float* add, long long add_rows, void* out, int out_is_f32, long long ldo, long long M,
                             int C, float eps, void* stream) {
  if (!x || !gamma || !beta || !out || C % 8 || C > 2048 || ldx % 8 || ldo % 8 || M <= 0) return CTU_E_BADARG;
  if (add && add_rows <= 0) return CTU_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (!x_is_f32 && !out_is_f32 && !add && (C == 64 || C == 128 || C == 256)) {
    const int rpb = 256 / (C / 8);
    const int grid = grid_for(M, rpb * 4);
#define CTU_LNN(L)                                                                                                  \
  layernorm_narrow_kernel<L><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, ldx, gamma, beta, (__nv_bfloat16*)out, \
                                                   ldo, M, eps)
    if (C == 64) CTU_LNN(8);
    else if (C == 128) CTU_LNN(16);
    else CTU_LNN(32);
#undef CTU_LNN
    count_launch();
    return (int)cudaGetLastError();
  }
  if (x_is_f32 && out_is_f32) return launch_ln<float, float>(x, ldx, gamma, beta, add, add_rows, out, ldo, M, C, eps, st);
  if (x_is_f32) return launch_ln<float, __nv_bfloat16>(x, ldx, gamma, beta, add, add_rows, out, ldo, M, C, eps, st);
  if (out_is_f32) return launch_ln<__nv_bfloat16, float>(x, ldx, gamma, beta, add, add_rows, out, ldo, M, C, eps, st);
  return launch_ln<__nv_bfloat16, __nv_bfloat16>(x, ldx, gamma, beta, add, add_rows, out, ldo, M, C, eps, st);
}

extern "C" int ctu_patchify_ln(const float* img, int B, int X, int Y, int Z, int pf, const float* gamma,
                               const float* beta, void* out, float eps, void* stream) {
  if (!img || !gamma || !beta || !out || X % 16 || Y % 16 || Z % pf) return CTU_E_BADARG;
  const int tokens = B * (X / 16) * (Y / 16) * (Z / pf);
  cudaStream_t st = (cudaStream_t)stream;
  if (pf == 8) patchify_ln_kernel<8><<<tokens, 256, 0, st>>>(img, X, Y, Z, gamma, beta, (__nv_bfloat16*)out, eps);
  else if (pf == 16) patchify_ln_kernel<16><<<tokens, 256, 0, st>>>(img, X, Y, Z, gamma, beta, (__nv_bfloat16*)out, eps);
  else return CTU_E_UNSUPPORTED;
  count_launch();
  return (int)cudaGetLastError();
}

extern "C" int ctu_pwa_fuse(const void* qkv1, const void* qkv2, void* out, long long T, int C, int dim_head, void* stream) {
  if (!qkv1 || !qkv2 || !out || dim_head != 32 || C % 32 || T <= 0) return CTU_E_BADARG;
  const float scale = 1.0f / sqrtf((float)dim_head);
  const int grid = grid_for(T * (C / 8), 256);
  pwa_fuse_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)qkv1, (const __nv_bfloat16*)qkv2,
                                                          (__nv_bfloat16*)out, T, C, scale);
  count_launch();
  return (int)cudaGetLastError();
}

extern "C" int ctu_subsample(const void* in, int ldi, int i1, int i2, int i3, void* out, int ldo, int s1, int s2,
                             int s3, int C, int B, void* stream) {
  if (!in || !out || C % 8 || ldi % 8 || ldo % 8 || s1 < 1 || s2 < 1 || s3 < 1) return CTU_E_BADARG;
  const int o1 = (i1 + s1 - 1) / s1, o2 = (i2 + s2 - 1) / s2, o3 = (i3 + s3 - 1) / s3;
  const long long total = (long long)B * o1 * o2 * o3 * (C / 8);
  subsample_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)in, ldi, i1, i2, i3,
                                                                          (__nv_bfloat16*)out, ldo, o1, o2, o3, s1, s2,
                                                                          s3, C, B);
  count_launch();
  return (int)cudaGetLastError();
}
