// Fused softmax(Q K^T * scale + bias) V for the short sequences of the CTUNet path (sm_100a):
//   - ViT MHSA, 432 tokens, 12 heads x 64                       (vit.py:66-78)
//   - MultiAxisAttention, 216-token 6x6x6 windows, dim/32 heads x 32, block or grid partition, additive
//     relative-position bias                                     (hybrid_CTUNet.py:481-511, 559-567)
// These are <1 % of the path's FLOPs and far too short for a TMEM pipeline, so they run flash-style on
// mma.sync m16n8k16 (bf16 in, fp32 accumulate) with the whole K/V of one (window, head) resident in shared
// memory.  The window / grid partition of the reference's einops rearranges is folded into the row gather,
// so activations never leave their natural channels-last order.
#include "attention_common.cuh"
#include "../../include/ctunet_b200.h"
#include "host_util.h"

namespace ctu {

// One CTA = NW warps x 16 queries of one (window, head); K and V of the whole (window, head) are staged ONCE per CTA
// as row-major [key][D] tiles (16-byte vector stores only).  The 6^3 windows (n = 216 <= 224) use NW = 14 so that a
// single CTA owns every query of the (window, head) and nothing is staged twice; the ViT (n = 432, D = 64) uses four
// warps per CTA and seven CTAs per (batch, head).  The P.V product takes its B fragments from the row-major V tile with
// ldmatrix.trans (no transposed copy).  KC = keys per online-softmax step.
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

template <int D, int NW, int KC>
__global__ void __launch_bounds__(32 * NW) attention_kernel(const __nv_bfloat16* __restrict__ qkv, int ld_qkv, int C,
                                                            __nv_bfloat16* __restrict__ out, int ldo,
                                                            const float* __restrict__ bias, float scale, int n, int n_pad,
                                                            TokenMap map, float* __restrict__ lse) {
  constexpr int LDK = D + 8;   // row pitch 80 / 144 bytes: 16-byte aligned rows, conflict-free fragment and ldmatrix reads
  constexpr int NT = 32 * NW;
  extern __shared__ __align__(16) uint8_t smem_att[];
  __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(smem_att);  // [n_pad][LDK]
  __nv_bfloat16* Vs = Ks + (size_t)n_pad * LDK;                     // [n_pad][LDK]

  const int qt = blockIdx.x, h = blockIdx.y, win = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;

  // ---- stage K and V of this (window, head)
  constexpr int VPR = D / 8;
  for (int i = tid; i < n_pad * VPR; i += NT) {
    const int j = i / VPR, vi = i % VPR;
    uint4 kv = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
    if (j < n) {
      const __nv_bfloat16* rp = qkv + token_row(map, win, j, n) * ld_qkv + h * D + vi * 8;
      kv = *reinterpret_cast<const uint4*>(rp + C);
      vv = *reinterpret_cast<const uint4*>(rp + 2 * C);
    }
    *reinterpret_cast<uint4*>(Ks + (size_t)j * LDK + vi * 8) = kv;
    *reinterpret_cast<uint4*>(Vs + (size_t)j * LDK + vi * 8) = vv;
  }

  // ---- Q fragments (A operand) straight from global memory
  const int q0 = qt * (16 * NW) + warp * 16;
  const int qa = q0 + g, qb = q0 + g + 8;
  const bool va = qa < n, vb = qb < n;
  const long long ra = va ? token_row(map, win, qa, n) : 0;
  const long long rb = vb ? token_row(map, win, qb, n) : 0;
  uint32_t qf[D / 16][4];
#pragma unroll
  for (int kk = 0; kk < D / 16; ++kk) {
    const int c = h * D + kk * 16 + 2 * t;
    qf[kk][0] = va ? *reinterpret_cast<const uint32_t*>(qkv + ra * ld_qkv + c) : 0u;
    qf[kk][1] = vb ? *reinterpret_cast<const uint32_t*>(qkv + rb * ld_qkv + c) : 0u;
    qf[kk][2] = va ? *reinterpret_cast<const uint32_t*>(qkv + ra * ld_qkv + c + 8) : 0u;
    qf[kk][3] = vb ? *reinterpret_cast<const uint32_t*>(qkv + rb * ld_qkv + c + 8) : 0u;
  }
  __syncthreads();

  float o[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  float m_a = -INFINITY, m_b = -INFINITY, l_a = 0.f, l_b = 0.f;
  const float LOG2E = 1.4426950408889634f;
  const float sl = scale * LOG2E;
  const float* bias_a = bias ? bias + ((long long)h * n + (va ? qa : 0)) * n : nullptr;
  const float* bias_b = bias ? bias + ((long long)h * n + (vb ? qb : 0)) * n : nullptr;

  for (int k0 = 0; k0 < n_pad; k0 += KC) {
    float s[KC / 8][4];
#pragma unroll
    for (int j = 0; j < KC / 8; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
      const __nv_bfloat16* kp = Ks + (size_t)(k0 + j * 8 + g) * LDK + 2 * t;
#pragma unroll
      for (int kk = 0; kk < D / 16; ++kk) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(kp + kk * 16);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(kp + kk * 16 + 8);
        mma_bf16_16816(s[j], qf[kk], b0, b1);
      }
    }
    // scale (+bias), mask, running max
    float mx_a = -INFINITY, mx_b = -INFINITY;
#pragma unroll
    for (int j = 0; j < KC / 8; ++j) {
      const int key = k0 + j * 8 + 2 * t;
      float ba0 = 0.f, ba1 = 0.f, bb0 = 0.f, bb1 = 0.f;
      if (bias != nullptr && key < n) {
        const float2 x = *reinterpret_cast<const float2*>(bias_a + key);
        const float2 y = *reinterpret_cast<const float2*>(bias_b + key);
        ba0 = x.x * LOG2E; ba1 = x.y * LOG2E; bb0 = y.x * LOG2E; bb1 = y.y * LOG2E;
      }
      s[j][0] = key < n ? fmaf(s[j][0], sl, ba0) : -INFINITY;
      s[j][1] = key + 1 < n ? fmaf(s[j][1], sl, ba1) : -INFINITY;
      s[j][2] = key < n ? fmaf(s[j][2], sl, bb0) : -INFINITY;
      s[j][3] = key + 1 < n ? fmaf(s[j][3], sl, bb1) : -INFINITY;
      mx_a = fmaxf(mx_a, fmaxf(s[j][0], s[j][1]));
      mx_b = fmaxf(mx_b, fmaxf(s[j][2], s[j][3]));
    }
    mx_a = fmaxf(mx_a, __shfl_xor_sync(0xffffffffu, mx_a, 1));
    mx_a = fmaxf(mx_a, __shfl_xor_sync(0xffffffffu, mx_a, 2));
    mx_b = fmaxf(mx_b, __shfl_xor_sync(0xffffffffu, mx_b, 1));
    mx_b = fmaxf(mx_b, __shfl_xor_sync(0xffffffffu, mx_b, 2));
    const float mn_a = fmaxf(m_a, mx_a), mn_b = fmaxf(m_b, mx_b);  // finite: chunk 0 always holds valid keys
    const float ca = exp2f(m_a - mn_a), cb = exp2f(m_b - mn_b);
    m_a = mn_a; m_b = mn_b;
    l_a *= ca; l_b *= cb;
#pragma unroll
    for (int i = 0; i < D / 8; ++i) { o[i][0] *= ca; o[i][1] *= ca; o[i][2] *= cb; o[i][3] *= cb; }
    uint32_t pf[KC / 16][4];
#pragma unroll
    for (int j = 0; j < KC / 8; ++j) {
      const float p0 = exp2f(s[j][0] - mn_a), p1 = exp2f(s[j][1] - mn_a);
      const float p2 = exp2f(s[j][2] - mn_b), p3 = exp2f(s[j][3] - mn_b);
      l_a += p0 + p1; l_b += p2 + p3;
      // C fragments of key tiles (2kt, 2kt+1) are the A fragment of k-tile kt
      pf[j >> 1][(j & 1) * 2 + 0] = pack_bf16x2(p0, p1);
      pf[j >> 1][(j & 1) * 2 + 1] = pack_bf16x2(p2, p3);
    }
#pragma unroll
    for (int kt = 0; kt < KC / 16; ++kt) {
      // lane l addresses row (l & 15) of the 16-key tile, column block (l >> 4) of a pair of 8-wide d blocks
      const uint32_t vbase = smem_u32(Vs + (size_t)(k0 + kt * 16 + (lane & 15)) * LDK + (lane >> 4) * 8);
#pragma unroll
      for (int dn = 0; dn < D / 8; dn += 2) {
        uint32_t bfr[4];
        ldmatrix_x4_trans(bfr, vbase + dn * 16);
        mma_bf16_16816(o[dn], pf[kt], bfr[0], bfr[1]);
        mma_bf16_16816(o[dn + 1], pf[kt], bfr[2], bfr[3]);
      }
    }
  }
  l_a += __shfl_xor_sync(0xffffffffu, l_a, 1);
  l_a += __shfl_xor_sync(0xffffffffu, l_a, 2);
  l_b += __shfl_xor_sync(0xffffffffu, l_b, 1);
  l_b += __shfl_xor_sync(0xffffffffu, l_b, 2);
  const float ia = 1.f / l_a, ib = 1.f / l_b;
  if (lse != nullptr && t == 0) {
    // log2-domain log-sum-exp of the scaled (+biased) scores, kept for the backward pass
    const int heads = C / D;
    if (va) lse[ra * heads + h] = m_a + log2f(l_a);
    if (vb) lse[rb * heads + h] = m_b + log2f(l_b);
  }
#pragma unroll
  for (int dn = 0; dn < D / 8; ++dn) {
    const int c = h * D + dn * 8 + 2 * t;
    if (va) *reinterpret_cast<uint32_t*>(out + ra * ldo + c) = pack_bf16x2(o[dn][0] * ia, o[dn][1] * ia);
    if (vb) *reinterpret_cast<uint32_t*>(out + rb * ldo + c) = pack_bf16x2(o[dn][2] * ib, o[dn][3] * ib);
  }
}

}  // namespace ctu

using namespace ctu;

// qkv: bf16 [rows][ld_qkv] with q|k|v thirds of width C = heads*dim_head; out: bf16 [rows][ldo].
// mode 0: `windows` consecutive groups of n rows (ViT: windows = batch, n = 432).
// mode 1/2: block / grid partition of a [batch, X, Y, Z] token grid into w^3 windows (n = w^3).
// bias: fp32 [heads][n][n] added to the scaled scores, or NULL.
extern "C" int ctu_attention(const void* qkv, int ld_qkv, int C, int dim_head, void* out, int ldo, const float* bias,
                             int n, int windows, int mode, int batch, int X, int Y, int Z, int w, float* lse,
                             void* stream) {
  if (!qkv || !out || (dim_head != 32 && dim_head != 64) || C % dim_head || ld_qkv % 8 || ldo % 2) return CTU_E_BADARG;
  TokenMap m;
  m.mode = mode; m.X = X; m.Y = Y; m.Z = Z; m.w = w;
  m.nwx = m.nwy = m.nwz = 1;
  if (mode != 0) {
    if (w <= 0 || X % w || Y % w || Z % w || n != w * w * w) return CTU_E_BADARG;
    m.nwx = X / w; m.nwy = Y / w; m.nwz = Z / w;
    windows = batch * m.nwx * m.nwy * m.nwz;
  }
  if (bias && (n % 2)) return CTU_E_BADARG;
  const int heads = C / dim_head;
  const float scale = 1.0f / sqrtf((float)dim_head);
  cudaStream_t st = (cudaStream_t)stream;
  constexpr int kMaxSmem = 160 * 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attention_kernel<64, 4, 48>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(attention_kernel<32, 4, 48>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(attention_kernel<32, 14, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const bool one_cta = dim_head == 32 && n <= 224;   // 6^3 windows: one 14-warp CTA per (window, head)
  const int kc = one_cta ? 32 : 48;
  const int n_pad = (n + kc - 1) / kc * kc;
  const size_t smem = (size_t)2 * n_pad * (dim_head + 8) * 2;
  if (smem > (size_t)kMaxSmem) return CTU_E_UNSUPPORTED;
  if (one_cta) {
    dim3 grid(1, heads, windows);
    attention_kernel<32, 14, 32><<<grid, 32 * 14, smem, st>>>((const __nv_bfloat16*)qkv, ld_qkv, C, (__nv_bfloat16*)out, ldo,
                                                              bias, scale, n, n_pad, m, lse);
  } else {
    dim3 grid((n + 63) / 64, heads, windows);
    if (dim_head == 64)
      attention_kernel<64, 4, 48><<<grid, 128, smem, st>>>((const __nv_bfloat16*)qkv, ld_qkv, C, (__nv_bfloat16*)out, ldo,
                                                           bias, scale, n, n_pad, m, lse);
    else
      attention_kernel<32, 4, 48><<<grid, 128, smem, st>>>((const __nv_bfloat16*)qkv, ld_qkv, C, (__nv_bfloat16*)out, ldo,
                                                           bias, scale, n, n_pad, m, lse);
  }
  count_launch();
  return (int)cudaGetLastError();
}
