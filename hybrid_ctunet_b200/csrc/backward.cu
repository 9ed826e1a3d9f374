// HBM-bound kernels of the BACKWARD pass of the CTUNet path (sm_100a): InstanceNorm(+LeakyReLU,+residual)
// backward, LayerNorm backward, GELU forward/backward, cross-weight fusion backward, column sums (bias
// gradients), layout conversions feeding the tensor-core dgrad / wgrad kernels (NCDHW fp32 -> channels-last
// bf16, space-to-depth, zero-stuffed up-sampling, single-channel im2col) and gradient accumulation.
// Same conventions as elementwise.cu: channels-last bf16, 16-byte vectors, grid-stride loops.
#include "common.cuh"
#include "../../include/ctunet_b200.h"
#include "host_util.h"

namespace ctu {

static int bw_num_sms() {
  static int n = [] {
    int dev = 0, v = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v;
  }();
  return n;
}

static inline int bw_grid(long long work_items, int per_block, int waves = 8) {
  long long blocks = (work_items + per_block - 1) / per_block;
  const long long cap = (long long)bw_num_sms() * waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

__device__ __forceinline__ void ld8(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 t;
  t = unpack_bf16x2(u.x); f[0] = t.x; f[1] = t.y;
  t = unpack_bf16x2(u.y); f[2] = t.x; f[3] = t.y;
  t = unpack_bf16x2(u.z); f[4] = t.x; f[5] = t.y;
  t = unpack_bf16x2(u.w); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const float (&f)[8]) {
  *reinterpret_cast<uint4*>(p) =
      make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}
__device__ __forceinline__ void ld8f(const float* p, float (&f)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void st8f(float* p, const float (&f)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}

__device__ __forceinline__ void in_coeffs_d(const double* st, double inv_n, float eps, double& rstd, double& shift) {
  const double mean = st[0] * inv_n;
  double var = st[1] * inv_n - mean * mean;
  var = var < 0.0 ? 0.0 : var;
  rstd = rsqrt(var + (double)eps);
  shift = -mean * rstd;
}

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  float2 t;
  t = unpack_bf16x2(u.x); f[0] = t.x; f[1] = t.y;
  t = unpack_bf16x2(u.y); f[2] = t.x; f[3] = t.y;
  t = unpack_bf16x2(u.z); f[4] = t.x; f[5] = t.y;
  t = unpack_bf16x2(u.w); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ uint4 ldg16(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }

// ------------------------------------------------------------------------------------------------
// Backward of  out = act( IN(x) [+ r | + IN(r)] )  (forward: in_apply_kernel).  With g = dout * act'(out):
//   dx = rstd_x * (g - mean(g) - xhat * mean(g * xhat)),   dr = g   or the same formula with r's statistics.
// Pass 1 accumulates the per-(instance, channel) sums (fp64 atomics), pass 2 applies them.
//
// Both passes: a thread owns 8 channels of a row (16-byte vectors, a warp covers 512 contiguous bytes) and walks rows
// with a grid stride, UNR rows per iteration with every load issued before the first use.  Per-channel coefficients
// are computed ONCE per block (one fp64 rsqrt per channel, spread over the threads) and staged in shared memory.
// `rev` walks the instance from its last row to its first: pass 1 runs right after the kernel that produced dout
// front to back, so the tail of dout is what the L2 still holds; pass 2 then walks forward over what pass 1 read last.

template <int RES, bool HASX, int IN_UNR>
__global__ void __launch_bounds__(256) in_bwd_stats_kernel(const __nv_bfloat16* __restrict__ dout, int ldd,
                                                           const __nv_bfloat16* __restrict__ out, int ldo,
                                                           const __nv_bfloat16* __restrict__ x, int ldx,
                                                           const double* __restrict__ xstats, int xs_ld,
                                                           const __nv_bfloat16* __restrict__ res, int ldr,
                                                           const double* __restrict__ rstats, int rs_ld, long long S,
                                                           int C, float eps, int act, float slope,
                                                           double* __restrict__ sums, int rev) {
  extern __shared__ __align__(16) float in_sm[];  // coefficients [4][C] first, then the block reduction [3][256][8]
  const int tpr = C / 8;
  const int rpb = 256 / tpr;
  const int cv = threadIdx.x % tpr;
  const int rl = threadIdx.x / tpr;
  const int b = blockIdx.y;
  const double inv_n = 1.0 / (double)S;
  const float inv_slope = 1.f / slope;
  constexpr bool has_x = HASX;
  if (has_x || RES == 2) {
    for (int c = threadIdx.x; c < C; c += 256) {
      double r_, s_;
      if (has_x) {
        in_coeffs_d(xstats + ((long long)b * xs_ld + c) * 2, inv_n, eps, r_, s_);
        in_sm[c] = (float)r_; in_sm[C + c] = (float)s_;
      }
      if (RES == 2) {
        in_coeffs_d(rstats + ((long long)b * rs_ld + c) * 2, inv_n, eps, r_, s_);
        in_sm[2 * C + c] = (float)r_; in_sm[3 * C + c] = (float)s_;
      }
    }
    __syncthreads();
  }
  float sc[8], sh[8], rsc[8], rsh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = has_x ? in_sm[cv * 8 + j] : 0.f;
    sh[j] = has_x ? in_sm[C + cv * 8 + j] : 0.f;
    rsc[j] = RES == 2 ? in_sm[2 * C + cv * 8 + j] : 0.f;
    rsh[j] = RES == 2 ? in_sm[3 * C + cv * 8 + j] : 0.f;
  }
  float sg[8], sgx[8], sgr[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sg[j] = sgx[j] = sgr[j] = 0.f;
  const long long base = (long long)b * S;
  const long long stride = (long long)gridDim.x * rpb;
  const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
  for (long long r0 = (long long)blockIdx.x * rpb + rl; r0 < S; r0 += stride * IN_UNR) {
    uint4 gq[IN_UNR], oq[IN_UNR], xq[IN_UNR], rq[IN_UNR];
#pragma unroll
    for (int u = 0; u < IN_UNR; ++u) {
      const long long r = r0 + u * stride;
      const bool ok = r < S;
      const long long row = base + (rev ? S - 1 - r : r);
      gq[u] = ok ? ldg16(dout + row * ldd + cv * 8) : z4;
      oq[u] = (ok && act) ? ldg16(out + row * ldo + cv * 8) : z4;
      xq[u] = (ok && has_x) ? ldg16(x + row * ldx + cv * 8) : z4;
      if (RES == 2) rq[u] = ok ? ldg16(res + row * ldr + cv * 8) : z4;
    }
#pragma unroll
    for (int u = 0; u < IN_UNR; ++u) {
      float g[8], o[8], xv[8], rv[8];
      unpack8(gq[u], g); unpack8(oq[u], o); unpack8(xq[u], xv);
      if (RES == 2) unpack8(rq[u], rv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float gj = g[j];
        if (act) gj = o[j] > 0.f ? gj : gj * slope;
        sg[j] += gj;
        // without a residual the normalised value is recovered from the output: xhat = lrelu^-1(out)
        const float xh = has_x ? fmaf(xv[j], sc[j], sh[j]) : (o[j] > 0.f ? o[j] : o[j] * inv_slope);
        sgx[j] += gj * xh;
        if (RES == 2) sgr[j] += gj * fmaf(rv[j], rsc[j], rsh[j]);
      }
    }
  }
  // block reduction over the rpb row slots: one output (array, channel) per thread, then ONE fp64 atomic each
  __syncthreads();  // everyone is done with the coefficient table
  float* red = in_sm;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[(0 * 256 + threadIdx.x) * 8 + j] = sg[j];
    red[(1 * 256 + threadIdx.x) * 8 + j] = sgx[j];
    if (RES == 2) red[(2 * 256 + threadIdx.x) * 8 + j] = sgr[j];
  }
  __syncthreads();
  constexpr int NA = RES == 2 ? 3 : 2;
  for (int o = threadIdx.x; o < NA * C; o += 256) {
    const int a = o / C, c = o - a * C;
    const float* rp = red + (a * 256 + (c >> 3)) * 8 + (c & 7);
    float acc = 0.f;
    for (int i = 0; i < rpb; ++i) acc += rp[i * tpr * 8];
    atomicAdd(sums + ((long long)b * C + c) * 4 + a, (double)acc);
  }
}

template <int RES, bool HASX, int IN_UNR>
__global__ void __launch_bounds__(256) in_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dout, int ldd,
                                                           const __nv_bfloat16* __restrict__ out, int ldo,
                                                           const __nv_bfloat16* __restrict__ x, int ldx,
                                                           const double* __restrict__ xstats, int xs_ld,
                                                           const __nv_bfloat16* __restrict__ res, int ldr,
                                                           const double* __restrict__ rstats, int rs_ld, long long S,
                                                           int C, float eps, int act, float slope,
                                                           const double* __restrict__ sums,
                                                           __nv_bfloat16* __restrict__ dx, int lddx,
                                                           __nv_bfloat16* __restrict__ dres, int lddr, int rev) {
  // dx = A g + B v + K with v = x (or xhat recovered from the output when x is not kept); likewise dr for RES == 2
  extern __shared__ __align__(16) float in_sm[];  // [6][C]
  const int tpr = C / 8;
  const int rpb = 256 / tpr;
  const int cv = threadIdx.x % tpr;
  const int rl = threadIdx.x / tpr;
  const int b = blockIdx.y;
  const double inv_n = 1.0 / (double)S;
  constexpr bool has_x = HASX;
  for (int c = threadIdx.x; c < C; c += 256) {
    double r_, s_;
    in_coeffs_d(xstats + ((long long)b * xs_ld + c) * 2, inv_n, eps, r_, s_);
    const double* sp = sums + ((long long)b * C + c) * 4;
    const double mg = sp[0] * inv_n, mgx = sp[1] * inv_n;
    in_sm[c] = (float)r_;
    in_sm[C + c] = (float)(has_x ? -r_ * r_ * mgx : -r_ * mgx);
    in_sm[2 * C + c] = (float)(has_x ? -r_ * mg - s_ * r_ * mgx : -r_ * mg);
    if (RES == 2) {
      const double mgr = sp[2] * inv_n;
      in_coeffs_d(rstats + ((long long)b * rs_ld + c) * 2, inv_n, eps, r_, s_);
      in_sm[3 * C + c] = (float)r_;
      in_sm[4 * C + c] = (float)(-r_ * r_ * mgr);
      in_sm[5 * C + c] = (float)(-r_ * mg - s_ * r_ * mgr);
    }
  }
  __syncthreads();
  float cA[8], cB[8], cK[8], rA[8], rB[8], rK[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    cA[j] = in_sm[cv * 8 + j]; cB[j] = in_sm[C + cv * 8 + j]; cK[j] = in_sm[2 * C + cv * 8 + j];
    if (RES == 2) { rA[j] = in_sm[3 * C + cv * 8 + j]; rB[j] = in_sm[4 * C + cv * 8 + j]; rK[j] = in_sm[5 * C + cv * 8 + j]; }
  }
  const long long base = (long long)b * S;
  const float inv_slope = 1.f / slope;
  const long long stride = (long long)gridDim.x * rpb;
  const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
  for (long long r0 = (long long)blockIdx.x * rpb + rl; r0 < S; r0 += stride * IN_UNR) {
    uint4 gq[IN_UNR], oq[IN_UNR], xq[IN_UNR], rq[IN_UNR];
    long long rows[IN_UNR];
#pragma unroll
    for (int u = 0; u < IN_UNR; ++u) {
      const long long r = r0 + u * stride;
      const bool ok = r < S;
      rows[u] = ok ? base + (rev ? S - 1 - r : r) : -1;
      gq[u] = ok ? ldg16(dout + rows[u] * ldd + cv * 8) : z4;
      oq[u] = (ok && act) ? ldg16(out + rows[u] * ldo + cv * 8) : z4;
      xq[u] = (ok && has_x) ? ldg16(x + rows[u] * ldx + cv * 8) : z4;
      if (RES == 2) rq[u] = ok ? ldg16(res + rows[u] * ldr + cv * 8) : z4;
    }
#pragma unroll
    for (int u = 0; u < IN_UNR; ++u) {
      if (rows[u] < 0) continue;
      float g[8], o[8], xv[8], rv[8], ox[8], orr[8];
      unpack8(gq[u], g); unpack8(oq[u], o); unpack8(xq[u], xv);
      if (RES == 2) unpack8(rq[u], rv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float gj = g[j];
        if (act) gj = o[j] > 0.f ? gj : gj * slope;
        const float v = has_x ? xv[j] : (o[j] > 0.f ? o[j] : o[j] * inv_slope);
        ox[j] = fmaf(gj, cA[j], fmaf(v, cB[j], cK[j]));
        if (RES == 1) orr[j] = gj;
        if (RES == 2) orr[j] = fmaf(gj, rA[j], fmaf(rv[j], rB[j], rK[j]));
      }
      st8(dx + rows[u] * lddx + cv * 8, ox);
      if (RES) st8(dres + rows[u] * lddr + cv * 8, orr);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward.  dx = rstd * (gamma*dy - mean(gamma*dy) - xhat * mean(gamma*dy*xhat)) [+ dx_in].
// A row is handled by LPR lanes (8 channels per lane and vector, VPL vectors per lane): LPR = C/8 for C <= 256 so a
// warp works on several rows at once, a whole warp with VPL = 2..4 vectors for wider rows.  Rows are strided over a
// persistent grid; the per-channel dgamma / dbeta partial sums live in registers, are combined per block in shared
// memory and flushed with ONE global atomic per (block, channel).
template <int LPR, int VPL, typename TIN>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const TIN* __restrict__ x, long long ldx,
                                                            const float* __restrict__ gamma,
                                                            const __nv_bfloat16* __restrict__ dy, long long ldd,
                                                            const void* __restrict__ dx_in, int dxin_f32, long long ld_in,
                                                            float* __restrict__ dx_f32, long long ld_f,
                                                            __nv_bfloat16* __restrict__ dx_bf, long long ld_b,
                                                            float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                            long long M, int C, float eps) {
  extern __shared__ float ln_red[];  // [2][C]
  constexpr int RPB = 256 / LPR;     // rows per block iteration
  const int sub = threadIdx.x % LPR;
  const int rl = threadIdx.x / LPR;
  const int nvec = C / 8;
  for (int i = threadIdx.x; i < 2 * C; i += 256) ln_red[i] = 0.f;
  float dg[VPL][8], db[VPL][8], gm[VPL][8];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vi = sub + LPR * i;
#pragma unroll
    for (int j = 0; j < 8; ++j) { dg[i][j] = 0.f; db[i][j] = 0.f; gm[i][j] = 0.f; }
    if (vi < nvec) ld8f(gamma + vi * 8, gm[i]);
  }
  const float invC = 1.f / (float)C;
  // block-uniform trip count: every lane takes part in the shuffles, out-of-range rows are masked
  const long long iters = (M + (long long)gridDim.x * RPB - 1) / ((long long)gridDim.x * RPB);
  for (long long it = 0; it < iters; ++it) {
    const long long row = (it * gridDim.x + blockIdx.x) * RPB + rl;
    const bool rok = row < M;
    float v[VPL][8], d[VPL][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = sub + LPR * i;
      if (rok && vi < nvec) {
        if constexpr (sizeof(TIN) == 2) ld8(reinterpret_cast<const __nv_bfloat16*>(x) + row * ldx + vi * 8, v[i]);
        else ld8f(reinterpret_cast<const float*>(x) + row * ldx + vi * 8, v[i]);
        ld8(dy + row * ldd + vi * 8, d[i]);
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += v[i][j];
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) { v[i][j] = 0.f; d[i][j] = 0.f; }
      }
    }
#pragma unroll
    for (int o = LPR / 2; o >= 1; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * invC;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      if (sub + LPR * i < nvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float t = v[i][j] - mean; sq += t * t; }
      }
    }
#pragma unroll
    for (int o = LPR / 2; o >= 1; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq * invC + eps);
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      if (rok && sub + LPR * i < nvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (v[i][j] - mean) * rstd;
          const float gd = gm[i][j] * d[i][j];
          v[i][j] = xh;
          m1 += gd;
          m2 += gd * xh;
          dg[i][j] += d[i][j] * xh;
          db[i][j] += d[i][j];
        }
      }
    }
#pragma unroll
    for (int o = LPR / 2; o >= 1; o >>= 1) {
      m1 += __shfl_xor_sync(0xffffffffu, m1, o);
      m2 += __shfl_xor_sync(0xffffffffu, m2, o);
    }
    m1 *= invC; m2 *= invC;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = sub + LPR * i;
      if (rok && vi < nvec) {
        float o8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o8[j] = rstd * (gm[i][j] * d[i][j] - m1 - v[i][j] * m2);
        if (dx_in != nullptr) {
          float a[8];
          if (dxin_f32) ld8f(reinterpret_cast<const float*>(dx_in) + row * ld_in + vi * 8, a);
          else ld8(reinterpret_cast<const __nv_bfloat16*>(dx_in) + row * ld_in + vi * 8, a);
#pragma unroll
          for (int j = 0; j < 8; ++j) o8[j] += a[j];
        }
        if (dx_f32 != nullptr) st8f(dx_f32 + row * ld_f + vi * 8, o8);
        if (dx_bf != nullptr) st8(dx_bf + row * ld_b + vi * 8, o8);
      }
    }
  }
  __syncthreads();  // ln_red zero-fill visible
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vi = sub + LPR * i;
    if (vi < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(&ln_red[vi * 8 + j], dg[i][j]);
        atomicAdd(&ln_red[C + vi * 8 + j], db[i][j]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += 256) {
    atomicAdd(dgamma + i, ln_red[i]);
    atomicAdd(dbeta + i, ln_red[C + i]);
  }
}

// ------------------------------------------------------------------------------------------------
// Exact-erf GELU (vit.py:37; hybrid_CTUNet.py:520) as a stand-alone pass for the training path, where the
// pre-activation has to be kept for the backward, and its derivative.
__global__ void __launch_bounds__(256) gelu_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                   long long nvec) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    float f[8];
    ld8(x + i * 8, f);
#pragma unroll
    for (int j = 0; j < 8; j += 2) gelu_pair<false>(f[j], f[j + 1], f[j], f[j + 1]);   // packed fp32, one MUFU per element
    st8(y + i * 8, f);
  }
}
__global__ void __launch_bounds__(256) gelu_bwd_kernel(const __nv_bfloat16* __restrict__ x,
                                                       const __nv_bfloat16* __restrict__ dy,
                                                       __nv_bfloat16* __restrict__ dx, long long nvec) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    float f[8], g[8];
    ld8(x + i * 8, f);
    ld8(dy + i * 8, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float cdf = 0.5f * (1.f + erff(f[j] * 0.70710678118654752440f));
      const float pdf = 0.3989422804014327f * __expf(-0.5f * f[j] * f[j]);
      g[j] *= cdf + f[j] * pdf;
    }
    st8(dx + i * 8, g);
  }
}

// ------------------------------------------------------------------------------------------------
// Backward of the binary cross-weight fusion (forward: pwa_fuse_kernel, hybrid_CTUNet.py:658-665).
__global__ void __launch_bounds__(256) pwa_fuse_bwd_kernel(const __nv_bfloat16* __restrict__ qkv1,
                                                           const __nv_bfloat16* __restrict__ qkv2,
                                                           const __nv_bfloat16* __restrict__ dout,
                                                           __nv_bfloat16* __restrict__ dqkv1,
                                                           __nv_bfloat16* __restrict__ dqkv2, long long T, int C,
                                                           float scale) {
  const int tpr = C / 8;
  const long long total = T * tpr;
  for (long long base = (long long)blockIdx.x * blockDim.x; base < total; base += (long long)gridDim.x * blockDim.x) {
    long long i = base + threadIdx.x;
    const bool ok = i < total;
    if (!ok) i = total - 1;
    const long long t = i / tpr;
    const int cv = (int)(i - t * tpr);
    const long long o3 = t * 3 * C + cv * 8;
    float q1[8], k1[8], v1[8], q2[8], k2[8], v2[8], go[8];
    ld8(qkv1 + o3, q1); ld8(qkv1 + o3 + C, k1); ld8(qkv1 + o3 + 2 * C, v1);
    ld8(qkv2 + o3, q2); ld8(qkv2 + o3 + C, k2); ld8(qkv2 + o3 + 2 * C, v2);
    ld8(dout + t * C + cv * 8, go);
    float d1 = 0.f, d2 = 0.f, da1 = 0.f, da2 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      d1 += q2[j] * k1[j]; d2 += q1[j] * k2[j];
      da1 += go[j] * v1[j]; da2 += go[j] * v2[j];
    }
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
      d1 += __shfl_xor_sync(0xffffffffu, d1, o); d2 += __shfl_xor_sync(0xffffffffu, d2, o);
      da1 += __shfl_xor_sync(0xffffffffu, da1, o); da2 += __shfl_xor_sync(0xffffffffu, da2, o);
    }
    d1 *= scale; d2 *= scale;
    const float m = fmaxf(d1, d2);
    const float e1 = __expf(d1 - m), e2 = __expf(d2 - m);
    const float inv = 1.f / (e1 + e2);
    const float a1 = e1 * inv, a2 = e2 * inv;
    const float dd1 = a1 * a2 * (da1 - da2) * scale;  // d loss / d <q2,k1>; d loss / d <q1,k2> = -dd1
    float o8[8];
    if (!ok) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) o8[j] = -dd1 * k2[j];
    st8(dqkv1 + o3, o8);                       // dq1
#pragma unroll
    for (int j = 0; j < 8; ++j) o8[j] = dd1 * q2[j];
    st8(dqkv1 + o3 + C, o8);                   // dk1
#pragma unroll
    for (int j = 0; j < 8; ++j) o8[j] = a1 * go[j];
    st8(dqkv1 + o3 + 2 * C, o8);               // dv1
#pragma unroll
    for (int j = 0; j < 8; ++j) o8[j] = dd1 * k1[j];
    st8(dqkv2 + o3, o8);                       // dq2
#pragma unroll
    for (int j = 0; j < 8; ++j) o8[j] = -dd1 * q1[j];
    st8(dqkv2 + o3 + C, o8);                   // dk2
#pragma unroll
    for (int j = 0; j < 8; ++j) o8[j] = a2 * go[j];
    st8(dqkv2 + o3 + 2 * C, o8);               // dv2
  }
}

// ------------------------------------------------------------------------------------------------
// out[c] += sum over rows of x[row][c]  (bias gradients, position-embedding gradient, relative-position bias
// gradient summed over windows).  A block covers CW column vectors x (256 / CW) row lanes — CW = 8 for the narrow
// matrices of the bias gradients (so that every lane loads), 32 otherwise; grid.y splits the rows.
template <typename T, int CW>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, long long ldx, long long M, long long N,
                                                     float* __restrict__ out) {
  constexpr int V = sizeof(T) == 2 ? 8 : 4;
  constexpr int RL = 256 / CW;
  __shared__ float red[RL][CW][V];
  const int cl = threadIdx.x % CW, rl = threadIdx.x / CW;
  const long long cvec = (long long)blockIdx.x * CW + cl;
  const long long nvec = N / V;
  float acc[V];
#pragma unroll
  for (int j = 0; j < V; ++j) acc[j] = 0.f;
  if (cvec < nvec) {
    // eight rows per iteration, every load issued before the first add
    constexpr int UNR = 8;
    const long long stride = (long long)gridDim.y * RL;
    for (long long r0 = (long long)blockIdx.y * RL + rl; r0 < M; r0 += stride * UNR) {
      uint4 q[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const long long r = r0 + u * stride;
        q[u] = r < M ? *reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(x) + (r * ldx + cvec * V) * sizeof(T))
                     : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        if constexpr (sizeof(T) == 2) {
          float f[8];
          unpack8(q[u], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += f[j];
        } else {
          acc[0] += __uint_as_float(q[u].x); acc[1] += __uint_as_float(q[u].y);
          acc[2] += __uint_as_float(q[u].z); acc[3] += __uint_as_float(q[u].w);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < V; ++j) red[rl][cl][j] = acc[j];
  __syncthreads();
  if (rl == 0 && cvec < nvec) {
#pragma unroll
    for (int j = 0; j < V; ++j) {
      float s = 0.f;
      for (int i = 0; i < RL; ++i) s += red[i][cl][j];
      atomicAdd(out + cvec * V + j, s);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// NCDHW fp32 [B][C][S] -> channels-last bf16 [B][S][ldd], channels C..cpad-1 zero-filled (logit gradients
// entering the head dgrad / wgrad GEMMs).  Tile transpose through shared memory: 32 voxels x cpad channels.
__global__ void __launch_bounds__(256) cf_to_cl_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                       int C, long long S, int ldd, int cpad) {
  __shared__ float tile[64][33];
  const int b = blockIdx.y;
  const long long s0 = (long long)blockIdx.x * 32;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int c = w; c < cpad; c += 8) {
    float v = 0.f;
    if (c < C && s0 + lane < S) v = src[((long long)b * C + c) * S + s0 + lane];
    tile[c][lane] = v;
  }
  __syncthreads();
  const int nv = cpad / 8;
  for (int i = threadIdx.x; i < 32 * nv; i += 256) {
    const int r = i / nv, cv = i % nv;
    if (s0 + r < S) {
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = tile[cv * 8 + j][r];
      st8(dst + ((long long)b * S + s0 + r) * ldd + cv * 8, f);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Space-to-depth: in [B][X*u3][Y*u2][Z*u1][ldi] (C channels) -> out [B][X][Y][Z][u3*u2*u1*C] with column
// (sub*C + c), sub = (a3*u2 + a2)*u1 + a1 — the gradient of a kernel==stride transposed convolution / pixel
// shuffle arranged as the plain-GEMM output it was in the forward.
// IDX = unsigned when every row / element count fits 32 bits (the 64-bit divisions otherwise dominate the copy), two
// 16-byte elements in flight per thread.
template <typename IDX>
__global__ void __launch_bounds__(256) space_to_depth_kernel(const __nv_bfloat16* __restrict__ in, int ldi,
                                                             __nv_bfloat16* __restrict__ out, int B, int X, int Y, int Z,
                                                             int u3, int u2, int u1, int C) {
  const IDX tpr = (IDX)(C / 8);
  const IDX k3 = (IDX)(u1 * u2 * u3);
  const IDX total = (IDX)B * X * Y * Z * k3 * tpr;
  const IDX step = (IDX)gridDim.x * blockDim.x;
  auto src_of = [&](IDX i, long long& dst) -> const uint4* {
    const IDX r = i / tpr;
    const int cv = (int)(i - r * tpr);
    const IDX orow = r / k3;
    const int sub = (int)(r - orow * k3);
    IDX t = orow;
    const IDX tz = t / (IDX)Z;
    const int z = (int)(t - tz * (IDX)Z); t = tz;
    const IDX ty = t / (IDX)Y;
    const int y = (int)(t - ty * (IDX)Y); t = ty;
    const IDX b = t / (IDX)X;
    const int x = (int)(t - b * (IDX)X);
    const int a1 = sub % u1, a2 = (sub / u1) % u2, a3 = sub / (u1 * u2);
    const long long irow = (((long long)b * X * u3 + (x * u3 + a3)) * (Y * u2) + (y * u2 + a2)) * (Z * u1) + (z * u1 + a1);
    dst = (long long)r * C + cv * 8;
    return reinterpret_cast<const uint4*>(in + irow * ldi + cv * 8);
  };
  for (IDX i = (IDX)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += 2 * step) {
    long long d0, d1 = 0;
    const uint4* s0 = src_of(i, d0);
    const bool two = i + step < total;
    const uint4* s1 = two ? src_of(i + step, d1) : s0;
    const uint4 v0 = *s0;
    const uint4 v1 = *s1;
    *reinterpret_cast<uint4*>(out + d0) = v0;
    if (two) *reinterpret_cast<uint4*>(out + d1) = v1;
  }
}

// ------------------------------------------------------------------------------------------------
// Backward of subsample_kernel: dfull[b, x*s3, y*s2, z*s1, :] (+)= dsub[b,x,y,z,:]; other positions are written
// as zero unless `accumulate`.
__global__ void __launch_bounds__(256) subsample_bwd_kernel(const __nv_bfloat16* __restrict__ dsub, int lds,
                                                            __nv_bfloat16* __restrict__ dfull, int ldf, int I1, int I2,
                                                            int I3, int O1, int O2, int O3, int s1, int s2, int s3,
                                                            int C, int B, int accumulate) {
  const int tpr = C / 8;
  const long long total = (long long)B * I3 * I2 * I1 * tpr;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long r = i / tpr;
    const int cv = (int)(i - r * tpr);
    const long long frow = r;
    const int z = (int)(r % I1); r /= I1;
    const int y = (int)(r % I2); r /= I2;
    const int x = (int)(r % I3);
    const int b = (int)(r / I3);
    const bool hit = (z % s1 == 0) && (y % s2 == 0) && (x % s3 == 0);
    __nv_bfloat16* dp = dfull + frow * ldf + cv * 8;
    if (hit) {
      const long long srow = (((long long)b * O3 + x / s3) * O2 + y / s2) * O1 + z / s1;
      float g[8];
      ld8(dsub + srow * lds + cv * 8, g);
      if (accumulate) {
        float a[8];
        ld8(dp, a);
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] += a[j];
      }
      st8(dp, g);
    } else if (!accumulate) {
      *reinterpret_cast<uint4*>(dp) = make_uint4(0, 0, 0, 0);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// im2col of a single-channel fp32 volume into channels-last bf16 rows [B][Xo][Yo][Zo][kpad] (tap-major columns,
// taps..kpad-1 zero): lets the C_in = 1 convolutions (ResNet stem forward, and the weight gradients of the stem and of
// vit_encoder0) run on the tensor-core kernels.
// One block = a run of IM_ZT output voxels (32 for wide rows, up to 128 for 64-column rows) along z of one (b, xo, yo): the input patch they read (kx x ky x zp floats)
// is staged in shared memory with coalesced loads, then consecutive threads write consecutive 16-byte tap groups of a
// voxel row (coalesced stores); `tapoff` maps a tap to its offset inside the patch (-1: zero padding column).
__global__ void __launch_bounds__(256) im2col_cin1_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out,
                                                          int B, int X, int Y, int Z, int Xo, int Yo, int Zo, int kx,
                                                          int ky, int kz, int sx, int sy, int sz, int px, int py, int pz,
                                                          int kpad, int IM_ZT) {
  extern __shared__ float im_smem[];
  const int zp = (IM_ZT - 1) * sz + kz;          // z extent of the patch
  const int taps = kx * ky * kz;
  float* patch = im_smem;                        // [kx * ky][zp]
  int* tapoff = reinterpret_cast<int*>(im_smem + ((kx * ky * zp + 3) & ~3));   // [kpad], 16-byte aligned
  const int zruns = (Zo + IM_ZT - 1) / IM_ZT;
  const int gpv = kpad / 8;
  for (int k = threadIdx.x; k < kpad; k += 256) {
    const int fz = k % kz, fy = (k / kz) % ky, fx = k / (kz * ky);
    tapoff[k] = k < taps ? (fx * ky + fy) * zp + fz : -1;
  }
  const long long blocks = (long long)B * Xo * Yo * zruns;
  for (long long blk = blockIdx.x; blk < blocks; blk += gridDim.x) {
    long long v = blk;
    const int zr = (int)(v % zruns); v /= zruns;
    const int yo = (int)(v % Yo); v /= Yo;
    const int xo = (int)(v % Xo);
    const int b = (int)(v / Xo);
    const int zo0 = zr * IM_ZT;
    const int x0 = xo * sx - px, y0 = yo * sy - py, z0 = zo0 * sz - pz;
    const float* ib = img + (long long)b * X * Y * Z;
    __syncthreads();   // the previous tile's patch has been consumed (and tapoff is written)
    for (int i = threadIdx.x; i < kx * ky * zp; i += 256) {
      const int zi = i % zp, line = i / zp;
      const int fy = line % ky, fx = line / ky;
      const int x = x0 + fx, y = y0 + fy, z = z0 + zi;
      patch[i] = (x >= 0 && x < X && y >= 0 && y < Y && z >= 0 && z < Z) ? __ldg(ib + ((long long)x * Y + y) * Z + z) : 0.f;
    }
    __syncthreads();
    const int nz = Zo - zo0 < IM_ZT ? Zo - zo0 : IM_ZT;
    const long long vox0 = (((long long)b * Xo + xo) * Yo + yo) * Zo + zo0;
    for (int i = threadIdx.x; i < nz * gpv; i += 256) {
      const int zi = i / gpv, grp = i - zi * gpv;
      const int4 o0 = *reinterpret_cast<const int4*>(tapoff + grp * 8);
      const int4 o1 = *reinterpret_cast<const int4*>(tapoff + grp * 8 + 4);
      const int off[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
      const int zs = zi * sz;
      float val[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) val[e] = off[e] >= 0 ? patch[off[e] + zs] : 0.f;
      st8(out + (vox0 + zi) * kpad + grp * 8, val);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// dst[row][0..C) += src[row][0..C)   (gradient accumulation where an activation has several consumers);
// src / dst are bf16 or fp32 independently (bf16 branch gradients joining an fp32 residual-stream gradient).
template <typename TS, typename TD>
__global__ void __launch_bounds__(256) accumulate_kernel(const TS* __restrict__ src, long long lds, TD* __restrict__ dst,
                                                         long long ldd, long long M, int C) {
  const int tpr = C / 8;
  const long long total = M * tpr;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / tpr;
    const int cv = (int)(i - r * tpr);
    float a[8], b[8];
    if constexpr (sizeof(TS) == 2) ld8(reinterpret_cast<const __nv_bfloat16*>(src) + r * lds + cv * 8, a);
    else ld8f(reinterpret_cast<const float*>(src) + r * lds + cv * 8, a);
    if constexpr (sizeof(TD) == 2) ld8(reinterpret_cast<const __nv_bfloat16*>(dst) + r * ldd + cv * 8, b);
    else ld8f(reinterpret_cast<const float*>(dst) + r * ldd + cv * 8, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] += b[j];
    if constexpr (sizeof(TD) == 2) st8(reinterpret_cast<__nv_bfloat16*>(dst) + r * ldd + cv * 8, a);
    else st8f(reinterpret_cast<float*>(dst) + r * ldd + cv * 8, a);
  }
}

// dst(bf16)[row][0..C) = src(fp32)[row][0..C)
__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ src, long long lds,
                                                            __nv_bfloat16* __restrict__ dst, long long ldd, long long M,
                                                            int C) {
  const int tpr = C / 8;
  const long long total = M * tpr;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / tpr;
    const int cv = (int)(i - r * tpr);
    float f[8];
    ld8f(src + r * lds + cv * 8, f);
    st8(dst + r * ldd + cv * 8, f);
  }
}

// ------------------------------------------------------------------------------------------------
// Parameter gradients of the LayerNorm fused into the ViT patchify (forward: patchify_ln_kernel, vit.py:115-116):
// dgamma[e] += sum_tokens dtok[e] * xhat[e], dbeta[e] += sum_tokens dtok[e].  The image needs no gradient.
// Block = 256 threads = (p1, p2); loops over tokens with a grid stride, partial sums in registers.
template <int PF>
__global__ void __launch_bounds__(256) patchify_ln_bwd_kernel(const float* __restrict__ img, int X, int Y, int Z,
                                                              int tokens, const __nv_bfloat16* __restrict__ dtok,
                                                              float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                              float eps) {
  __shared__ float red[8];
  __shared__ float bc;
  const int nh = X / 16, nw = Y / 16, nf = Z / PF;
  const int p1 = threadIdx.x / 16, p2 = threadIdx.x % 16;
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const float n = 256.f * PF;
  float dg[PF], db[PF];
#pragma unroll
  for (int j = 0; j < PF; ++j) dg[j] = db[j] = 0.f;
  for (int tk = blockIdx.x; tk < tokens; tk += gridDim.x) {
    int tok = tk;
    const int f = tok % nf; tok /= nf;
    const int w = tok % nw; tok /= nw;
    const int h = tok % nh;
    const int b = tok / nh;
    const float* src = img + (((long long)b * X + (h * 16 + p1)) * Y + (w * 16 + p2)) * Z + f * PF;
    float v[PF];
#pragma unroll
    for (int j = 0; j < PF; j += 4) {
      const float4 t = *reinterpret_cast<const float4*>(src + j);
      v[j] = t.x; v[j + 1] = t.y; v[j + 2] = t.z; v[j + 3] = t.w;
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < PF; ++j) s += v[j];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __syncthreads();
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
    const __nv_bfloat16* dp = dtok + (long long)tk * (256 * PF) + threadIdx.x * PF;
#pragma unroll
    for (int j = 0; j < PF; j += 8) {
      float g[8];
      ld8(dp + j, g);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        dg[j + k] += g[k] * (v[j + k] - mean) * rstd;
        db[j + k] += g[k];
      }
    }
  }
  const int e0 = threadIdx.x * PF;
#pragma unroll
  for (int j = 0; j < PF; ++j) {
    atomicAdd(dgamma + e0 + j, dg[j]);
    atomicAdd(dbeta + e0 + j, db[j]);
  }
}

}  // namespace ctu

using namespace ctu;
typedef __nv_bfloat16 bf16;

static inline void in_grid(long long S, int C, int B, int rows_per_thread, dim3& grid) {
  const int rpb = 256 / (C / 8);
  int gx = bw_grid(S, rpb * rows_per_thread);
  const int cap = (bw_num_sms() * 8 + B - 1) / B;
  if (gx > cap) gx = cap;
  grid = dim3(gx, B);
}

extern "C" int ctu_in_bwd_stats(const void* dout, int ldd, const void* out, int ldo, const void* x, int ldx,
                                const double* xstats, int xs_ld, const void* res, int ldr, const double* rstats,
                                int rs_ld, int B, long long S, int C, float eps, int act, float slope, double* sums,
                                void* stream) {
  if (!dout || !xstats || !sums || (act && !out) || C % 8 || C > 2048 || (2048 % C)) return CTU_E_BADARG;
  if (!x && (!act || res)) return CTU_E_BADARG;  /* xhat from the output needs the invertible activation, no residual */
  if (ldd % 8 || (x && ldx % 8) || (act && ldo % 8) || (rstats && (!res || ldr % 8))) return CTU_E_BADARG;
  dim3 grid;
  in_grid(S, C, B, 16, grid);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t red_bytes = (size_t)(rstats ? 3 : 2) * 256 * 8 * sizeof(float);
  const size_t coef_bytes = (size_t)4 * C * sizeof(float);
  const size_t smem = red_bytes > coef_bytes ? red_bytes : coef_bytes;
#define CTU_INS(R, HX, U)                                                                                              \
  in_bwd_stats_kernel<R, HX, U><<<grid, 256, smem, st>>>((const bf16*)dout, ldd, (const bf16*)out, ldo, (const bf16*)x,  \
                                                         ldx, xstats, xs_ld, (const bf16*)(R == 2 ? res : nullptr), ldr, \
                                                         R == 2 ? rstats : nullptr, rs_ld, S, C, eps, act, slope, sums, 1)
  if (rstats) {
    if (!x) return CTU_E_BADARG;
    CTU_INS(2, true, 2);
  } else if (x) {
    CTU_INS(0, true, 2);
  } else {
    CTU_INS(0, false, 4);
  }
#undef CTU_INS
  count_launch();
  return (int)cudaGetLastError();
}

extern "C" int ctu_in_bwd_apply(const void* dout, int ldd, const void* out, int ldo, const void* x, int ldx,
                                const double* xstats, int xs_ld, const void* res, int ldr, const double* rstats,
                                int rs_ld, int res_mode, int B, long long S, int C, float eps, int act, float slope,
                                const double* sums, void* dx, int lddx, void* dres, int lddr, void* stream) {
  if (!dout || !xstats || !sums || !dx || (act && !out) || C % 8 || C > 2048 || (2048 % C)) return CTU_E_BADARG;
  if (!x && (!act || res_mode != 0)) return CTU_E_BADARG;
  if (ldd % 8 || (x && ldx % 8) || lddx % 8 || (act && ldo % 8)) return CTU_E_BADARG;
  if (res_mode < 0 || res_mode > 2 || (res_mode && (!dres || lddr % 8)) || (res_mode == 2 && (!res || !rstats || ldr % 8)))
    return CTU_E_BADARG;
  if (res_mode && !x) return CTU_E_BADARG;
  dim3 grid;
  in_grid(S, C, B, 8, grid);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)(res_mode == 2 ? 6 : 3) * C * sizeof(float);
#define CTU_INB(R, HX, U)                                                                                              \
  in_bwd_apply_kernel<R, HX, U><<<grid, 256, smem, st>>>((const bf16*)dout, ldd, (const bf16*)out, ldo, (const bf16*)x,  \
                                                         ldx, xstats, xs_ld, (const bf16*)res, ldr, rstats, rs_ld, S, C, \
                                                         eps, act, slope, sums, (bf16*)dx, lddx, (bf16*)dres, lddr, 0)
  if (res_mode == 0 && !x) CTU_INB(0, false, 4);
  else if (res_mode == 0) CTU_INB(0, true, 2);
  else if (res_mode == 1) CTU_INB(1, true, 2);
  else CTU_INB(2, true, 2);
#undef CTU_INB
  count_launch();
  return (int)cudaGetLastError();
}

template <typename TIN>
static int launch_ln_bwd(const void* x, long long ldx, const float* gamma, const void* dy, long long ldd, const void* dx_in,
                         int dxin_f32, long long ld_in, float* dx_f32, long long ld_f, void* dx_bf, long long ld_b,
                         float* dgamma, float* dbeta, long long M, int C, float eps, cudaStream_t st) {
  const int nvec = C / 8;
  const size_t smem = 2 * (size_t)C * sizeof(float);
#define CTU_LNB(L, V)                                                                                                  \
  do {                                                                                                                 \
    long long blocks = (M + (256 / L) - 1) / (256 / L);                                                                \
    const long long cap = (long long)bw_num_sms() * 6;                                                                 \
    if (blocks > cap) blocks = cap;                                                                                    \
    layernorm_bwd_kernel<L, V, TIN><<<(unsigned)blocks, 256, smem, st>>>((const TIN*)x, ldx, gamma, (const bf16*)dy, ldd, \
                                                                         dx_in, dxin_f32, ld_in, dx_f32, ld_f,         \
                                                                         (bf16*)dx_bf, ld_b, dgamma, dbeta, M, C, eps); \
  } while (0)
  if (nvec <= 8) CTU_LNB(8, 1);
  else if (nvec <= 16) CTU_LNB(16, 1);
  else if (nvec <= 32) CTU_LNB(32, 1);
  else if (nvec <= 64) CTU_LNB(32, 2);
  else if (nvec <= 128) CTU_LNB(32, 4);
  else return CTU_E_UNSUPPORTED;
#undef CTU_LNB
  count_launch();
  return (int)cudaGetLastError();
}

extern "C" int ctu_layernorm_bwd(const void* x, int x_is_f32, long long ldx, const float* gamma, const void* dy,
                                 long long ldd, const void* dx_in, int dxin_is_f32, long long ld_in, float* dx_f32,
                                 long long ld_f, void* dx_bf16, long long ld_b, float* dgamma, float* dbeta, long long M,
                                 int C, float eps, void* stream) {
  if (!x || !gamma || !dy || !dgamma || !dbeta || (!dx_f32 && !dx_bf16) || C % 8 || C > 1024 || M <= 0) return CTU_E_BADARG;
  if (ldx % 8 || ldd % 8 || (dx_in && ld_in % 8) || (dx_f32 && ld_f % 8) || (dx_bf16 && ld_b % 8)) return CTU_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (x_is_f32)
    return launch_ln_bwd<float>(x, ldx, gamma, dy, ldd, dx_in, dxin_is_f32, ld_in, dx_f32, ld_f, dx_bf16, ld_b, dgamma,
                                dbeta, M, C, eps, st);
  return launch_ln_bwd<bf16>(x, ldx, gamma, dy, ldd, dx_in, dxin_is_f32, ld_in, dx_f32, ld_f, dx_bf16, ld_b, dgamma, dbeta,
                             M, C, eps, st);
}

extern "C" int ctu_gelu(const void* x, void* y, long long n, void* stream) {
  if (!x || !y || n <= 0 || n % 8) return CTU_E_BADARG;
  gelu_kernel<<<bw_grid(n / 8, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)y, n / 8);
  count_launch();
  return (int)cudaGetLastError();
}

extern "C" int ctu_gelu_bwd(const void* x, const void* dy, void* dx, long long n, void* stream) {
  if (!x || !dy || !dx || n <= 0 || n % 8) return CTU_E_BADARG;
  gelu_bwd_kernel<<<bw_grid(n / 8, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (const bf16*)dy, (bf16*)dx, n / 8);
  count_launch();
  return (int)cudaGetLastError();
}

extern "C" int ctu_pwa_fuse_bwd(const void* qkv1, const void* qkv2, const void* dout, void* dqkv1, void* dqkv2,
                                long long T, int C, int dim_head, void* stream) {
  if (!qkv1 || !qkv2 || !dout || !dqkv1 || !dqkv2 || dim_head != 32 || C % 32 || T <= 0) return CTU_E_BADARG;
  const float scale = 1.0f / sqrtf((float)dim_head);
  pwa_fuse_bwd_kernel<<<bw_grid(T * (C / 8), 256), 256, 0, (cudaStream_t)stream>>>(
      (const bf16*)qkv1, (const bf16*)qkv2, (const bf16*)dout, (bf16*)dqkv1, (bf16*)dqkv2, T, C, scale);
  count_launch();
  return (int)cudaGetLastError();
}

extern "C" int ctu_colsum(const void* x, int x_is_f32, long long ldx, long long M, long long N, float* out, void* stream) {
  const int V = x_is_f32 ? 4 : 8;
  if (!x || !out || M <= 0 || N <= 0 || N % V || ldx % V) return CTU_E_BADARG;
  const long long nvec = N / V;
  const int cw = nvec <= 2 ? 2 : (nvec <= 4 ? 4 : (nvec <= 8 ? 8 : (nvec <= 16 ? 16 : 32)));
  const long long gx = (nvec + cw - 1) / cw;
  if (gx > 0x7fffffffLL) return CTU_E_BADARG;
  const int rl = 256 / cw;
  long long gy = (M + rl * 32 - 1) / (rl * 32);
  const long long cap = ((long long)bw_num_sms() * 8 + gx - 1) / gx;   // ~8 blocks per SM: each ends with one atomic per column
  if (gy > cap) gy = cap;
  if (gy > 65535) gy = 65535;
  if (gy < 1) gy = 1;
  dim3 grid((unsigned)gx, (unsigned)gy);
  cudaStream_t st = (cudaStream_t)stream;
  if (x_is_f32) {
    if (cw == 2) colsum_kernel<float, 2><<<grid, 256, 0, st>>>((const float*)x, ldx, M, N, out);
    else if (cw == 4) colsum_kernel<float, 4><<<grid, 256, 0, st>>>((const float*)x, ldx, M, N, out);
    else if (cw == 8) colsum_kernel<float, 8><<<grid, 256, 0, st>>>((const float*)x, ldx, M, N, out);
    else if (cw == 16) colsum_kernel<float, 16><<<grid, 256, 0, st>>>((const float*)x, ldx, M, N, out);
    else colsum_kernel<float, 32><<<grid, 256, 0, st>>>((const float*)x, ldx, M, N, out);
  } else {
    if (cw == 2) colsum_kernel<bf16, 2><<<grid, 256, 0, st>>>((const bf16*)x, ldx, M, N, out);
    else if (cw == 4) colsum_kernel<bf16, 4><<<grid, 256, 0, st>>>((const bf16*)x, ldx, M, N, out);
    else if (cw == 8) colsum_kernel<bf16, 8><<<grid, 256, 0, st>>>((const bf16*)x, ldx, M, N, out);
    else if (cw == 16) colsum_kernel<bf16, 16><<<grid, 256, 0, st>>>((const bf16*)x, ldx, M, N, out);
    else colsum_kernel<bf16, 32><<<grid, 256, 0, st>>>((const bf16*)x, ldx, M, N, out);
  }
  count_launch();
  return (int)cudaGetLastError();
}

extern "C" int ctu_cf_to_cl(const float* src, void* dst, int B, int C, long long S, int ldd, int cpad, void* stream) {
  if (!src || !dst || B <= 0 || C <= 0 || C > cpad || cpad > 64 || cpad % 8 || ldd % 8 || ldd < cpad) return CTU_E_BADARG;
  dim3 grid((unsigned)((S + 31) / 32), B);
  cf_to_cl_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, (bf16*)dst, C, S, ldd, cpad);
  count_launch();
  return (int)cudaGetLastError();
}

extern "C" int ctu_space_to_depth(const void* in, int ldi, void* out, int B, int X, int Y, int Z, int u3, int u2, int u1,
                                  int C, void* stream) {
  if (!in || !out || C % 8 || ldi % 8 || u1 < 1 || u2 < 1 || u3 < 1) return CTU_E_BADARG;
  const long long total = (long long)B * X * Y * Z * u1 * u2 * u3 * (C / 8);
  if (total < 0x7fffffffLL)
    space_to_depth_kernel<unsigned><<<bw_grid(total, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)in, ldi, (bf16*)out, B, X,
                                                                                         Y, Z, u3, u2, u1, C);
  else
    space_to_depth_kernel<long long><<<bw_grid(total, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)in, ldi, (bf16*)out, B,
                                                                                          X, Y, Z, u3, u2, u1, C);
  count_launch();
  return (int)cudaGetLastError();
}

extern "C" int ctu_subsample_bwd(const void* dsub, int lds, void* dfull, int ldf, int i1, int i2, int i3, int s1, int s2,
                                 int s3, int C, int B, int accumulate, void* stream) {
  if (!dsub || !dfull || C % 8 || lds % 8 || ldf % 8 || s1 < 1 || s2 < 1 || s3 < 1) return CTU_E_BADARG;
  const int o1 = (i1 + s1 - 1) / s1, o2 = (i2 + s2 - 1) / s2, o3 = (i3 + s3 - 1) / s3;
  const long long total = (long long)B * i1 * i2 * i3 * (C / 8);
  subsample_bwd_kernel<<<bw_grid(total, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)dsub, lds, (bf16*)dfull, ldf, i1,
                                                                              i2, i3, o1, o2, o3, s1, s2, s3, C, B,
                                                                              accumulate);
  count_launch();
  return (int)cudaGetLastError();
}

extern "C" int ctu_im2col_cin1(const float* img, void* out, int B, int X, int Y, int Z, int kx, int ky, int kz, int sx,
                               int sy, int sz, int px, int py, int pz, int kpad, void* stream) {
  if (!img || !out || kpad % 8 || kpad < kx * ky * kz) return CTU_E_BADARG;
  const int Xo = (X + 2 * px - kx) / sx + 1, Yo = (Y + 2 * py - ky) / sy + 1, Zo = (Z + 2 * pz - kz) / sz + 1;
  const int IM_ZT = kpad <= 64 ? (Zo < 128 ? Zo : 128) : 32;   // enough 16-byte stores per block to amortise the staging
  const long long blocks = (long long)B * Xo * Yo * ((Zo + IM_ZT - 1) / IM_ZT);
  const int zp = (IM_ZT - 1) * sz + kz;
  const size_t smem = ((size_t)((kx * ky * zp + 3) & ~3) + kpad) * 4;
  if (smem > 48 * 1024) return CTU_E_UNSUPPORTED;
  const long long cap = (long long)bw_num_sms() * 16;
  im2col_cin1_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, smem, (cudaStream_t)stream>>>(
      img, (bf16*)out, B, X, Y, Z, Xo, Yo, Zo, kx, ky, kz, sx, sy, sz, px, py, pz, kpad, IM_ZT);
  count_launch();
  return (int)cudaGetLastError();
}

extern "C" int ctu_accumulate(const void* src, int src_is_f32, long long lds, void* dst, int dst_is_f32, long long ldd,
                              long long M, int C, void* stream) {
  if (!src || !dst || M <= 0 || C <= 0 || C % 8 || lds % (src_is_f32 ? 4 : 8) || ldd % (dst_is_f32 ? 4 : 8)) return CTU_E_BADARG;
  const int grid = bw_grid(M * (C / 8), 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (src_is_f32 && dst_is_f32) accumulate_kernel<float, float><<<grid, 256, 0, st>>>((const float*)src, lds, (float*)dst, ldd, M, C);
  else if (src_is_f32) accumulate_kernel<float, bf16><<<grid, 256, 0, st>>>((const float*)src, lds, (bf16*)dst, ldd, M, C);
  else if (dst_is_f32) accumulate_kernel<bf16, float><<<grid, 256, 0, st>>>((const bf16*)src, lds, (float*)dst, ldd, M, C);
  else accumulate_kernel<bf16, bf16><<<grid, 256, 0, st>>>((const bf16*)src, lds, (bf16*)dst, ldd, M, C);
  count_launch();
  return (int)cudaGetLastError();
}

extern "C" int ctu_cast_f32_bf16(const float* src, long long lds, void* dst, long long ldd, long long M, int C,
                                 void* stream) {
  if (!src || !dst || M <= 0 || C <= 0 || C % 8 || lds % 4 || ldd % 8) return CTU_E_BADARG;
  cast_f32_bf16_kernel<<<bw_grid(M * (C / 8), 256), 256, 0, (cudaStream_t)stream>>>(src, lds, (bf16*)dst, ldd, M, C);
  count_launch();
  return (int)cudaGetLastError();
}

extern "C" int ctu_patchify_ln_bwd(const float* img, int B, int X, int Y, int Z, int pf, const void* dtok, float* dgamma,
                                   float* dbeta, float eps, void* stream) {
  if (!img || !dtok || !dgamma || !dbeta || X % 16 || Y % 16 || Z % pf) return CTU_E_BADARG;
  const int tokens = B * (X / 16) * (Y / 16) * (Z / pf);
  int grid = tokens < bw_num_sms() * 2 ? tokens : bw_num_sms() * 2;
  cudaStream_t st = (cudaStream_t)stream;
  if (pf == 8) patchify_ln_bwd_kernel<8><<<grid, 256, 0, st>>>(img, X, Y, Z, tokens, (const bf16*)dtok, dgamma, dbeta, eps);
  else if (pf == 16) patchify_ln_bwd_kernel<16><<<grid, 256, 0, st>>>(img, X, Y, Z, tokens, (const bf16*)dtok, dgamma, dbeta, eps);
  else return CTU_E_UNSUPPORTED;
  count_launch();
  return (int)cudaGetLastError();
}
