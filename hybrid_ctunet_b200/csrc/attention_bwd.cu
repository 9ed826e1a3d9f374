// Backward of the fused softmax(Q K^T * scale + bias) V of attention.cu (sm_100a), for ViT MHSA (vit.py:66-78)
// and MultiAxisAttention (hybrid_CTUNet.py:481-511).  Same mma.sync m16n8k16 machinery as the forward:
// a CTA owns 64 keys (16 per warp) of one (window, head) and streams the queries in chunks of 32 through
// shared memory, recomputing P^T = exp2(S^T - lse) from the log-sum-exp the forward saved:
//   dV  += P^T dO            dP^T = V dO^T          dS^T = P^T o (dP^T - delta),  delta = rowsum(dO o O)
//   dK  += dS^T Q * scale    dQ   += dS K * scale   dBias = dS (summed over windows by ctu_colsum afterwards)
// dK / dV rows are owned by exactly one warp and written as bf16; dQ gets contributions from every key, so the
// warps park dS of the chunk in shared memory as one [query][key] tile and the dQ product over the CTA's whole key
// range is split by OUTPUT tile over the warps (no shared-memory atomics: fp32 shared atomics are CAS loops).  When
// the CTA does not hold every key, its partial dQ is added to an fp32 buffer with vector reductions.
#include "attention_common.cuh"
#include "../../include/ctunet_b200.h"
#include "host_util.h"

namespace ctu {

constexpr int QC = 32;          // queries per chunk
constexpr int LDT = QC + 8;     // row pitch of the transposed chunk copies

// delta[row][h] = sum_d dO[row][h*D + d] * O[row][h*D + d]
__global__ void __launch_bounds__(256) attn_delta_kernel(const __nv_bfloat16* __restrict__ o, long long ldo,
                                                         const __nv_bfloat16* __restrict__ dout, long long ldd,
                                                         float* __restrict__ delta, long long rows, int C, int D) {
  const int tpr = C / 8;
  const int lph = D / 8;  // lanes per head: 4 or 8
  const long long total = rows * tpr;
  for (long long base = (long long)blockIdx.x * blockDim.x; base < total; base += (long long)gridDim.x * blockDim.x) {
    long long i = base + threadIdx.x;
    const bool ok = i < total;
    if (!ok) i = total - 1;
    const long long r = i / tpr;
    const int cv = (int)(i - r * tpr);
    const uint4 a = *reinterpret_cast<const uint4*>(o + r * ldo + cv * 8);
    const uint4 b = *reinterpret_cast<const uint4*>(dout + r * ldd + cv * 8);
    float2 x, y;
    float s = 0.f;
    x = unpack_bf16x2(a.x); y = unpack_bf16x2(b.x); s += x.x * y.x + x.y * y.y;
    x = unpack_bf16x2(a.y); y = unpack_bf16x2(b.y); s += x.x * y.x + x.y * y.y;
    x = unpack_bf16x2(a.z); y = unpack_bf16x2(b.z); s += x.x * y.x + x.y * y.y;
    x = unpack_bf16x2(a.w); y = unpack_bf16x2(b.w); s += x.x * y.x + x.y * y.y;
    for (int off = 1; off < lph; off <<= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (ok && (cv % lph) == 0) delta[r * (C / D) + cv / lph] = s;
  }
}

template <int D, int NW, bool DIRECT_DQ>
__global__ void __launch_bounds__(32 * NW) attention_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, int ld_qkv, int C,
                                                            const __nv_bfloat16* __restrict__ dout, int ldd,
                                                            const float* __restrict__ lse, const float* __restrict__ delta,
                                                            const float* __restrict__ biasT, float scale,
                                                            __nv_bfloat16* __restrict__ dqkv, int ld_dqkv,
                                                            float* __restrict__ dq_f32, __nv_bfloat16* __restrict__ ds_out,
                                                            int n, TokenMap map) {
  constexpr int LDQ = D + 8;
  __shared__ __align__(16) __nv_bfloat16 Qs[QC * LDQ];
  __shared__ __align__(16) __nv_bfloat16 dOs[QC * LDQ];
  __shared__ __align__(16) __nv_bfloat16 Qt[D * LDT];
  __shared__ __align__(16) __nv_bfloat16 dOt[D * LDT];
  constexpr int NT = 32 * NW;  // threads per CTA: NW warps, 16 keys each
  constexpr int LDS = NW * 16 + 8;  // pitch of the key-contiguous buffers (LDS/2 = 4 mod 8: conflict-free fragments)
  __shared__ __align__(16) __nv_bfloat16 Kt[D * LDS];    // K^T of the CTA's keys: [d][key]
  __shared__ __align__(16) __nv_bfloat16 dSs[QC * LDS];  // dS of the current chunk: [query][key]
  __shared__ float lse_s[QC], delta_s[QC];
  __shared__ long long qrow_s[QC];

  const int kb = blockIdx.x, h = blockIdx.y, win = blockIdx.z;
  const int heads = C / D;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const float LOG2E = 1.4426950408889634f;
  const float sl = scale * LOG2E;

  // ---- this warp's 16 keys: K and V as A fragments (rows = keys), K^T in shared memory for the dQ product
  const int key0 = kb * (16 * NW) + warp * 16;
  const int ka = key0 + g, kbk = key0 + g + 8;
  const bool va = ka < n, vb = kbk < n;
  const long long ra = va ? token_row(map, win, ka, n) : 0;
  const long long rb = vb ? token_row(map, win, kbk, n) : 0;
  uint32_t kf[D / 16][4], vf[D / 16][4];
#pragma unroll
  for (int kk = 0; kk < D / 16; ++kk) {
    const int c = h * D + kk * 16 + 2 * t;
    kf[kk][0] = va ? *reinterpret_cast<const uint32_t*>(qkv + ra * ld_qkv + C + c) : 0u;
    kf[kk][1] = vb ? *reinterpret_cast<const uint32_t*>(qkv + rb * ld_qkv + C + c) : 0u;
    kf[kk][2] = va ? *reinterpret_cast<const uint32_t*>(qkv + ra * ld_qkv + C + c + 8) : 0u;
    kf[kk][3] = vb ? *reinterpret_cast<const uint32_t*>(qkv + rb * ld_qkv + C + c + 8) : 0u;
    vf[kk][0] = va ? *reinterpret_cast<const uint32_t*>(qkv + ra * ld_qkv + 2 * C + c) : 0u;
    vf[kk][1] = vb ? *reinterpret_cast<const uint32_t*>(qkv + rb * ld_qkv + 2 * C + c) : 0u;
    vf[kk][2] = va ? *reinterpret_cast<const uint32_t*>(qkv + ra * ld_qkv + 2 * C + c + 8) : 0u;
    vf[kk][3] = vb ? *reinterpret_cast<const uint32_t*>(qkv + rb * ld_qkv + 2 * C + c + 8) : 0u;
  }
  for (int i = lane; i < 16 * (D / 8); i += 32) {
    const int j = i / (D / 8), vi = i % (D / 8);
    uint4 kv = make_uint4(0, 0, 0, 0);
    if (key0 + j < n) kv = *reinterpret_cast<const uint4*>(qkv + token_row(map, win, key0 + j, n) * ld_qkv + C + h * D + vi * 8);
    const __nv_bfloat16* ke = reinterpret_cast<const __nv_bfloat16*>(&kv);
#pragma unroll
    for (int e = 0; e < 8; ++e) Kt[(vi * 8 + e) * LDS + warp * 16 + j] = ke[e];
  }

  float dk[D / 8][4], dv[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i) { dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f; dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f; }

  const float* bT_a = biasT ? biasT + ((long long)h * n + (va ? ka : 0)) * n : nullptr;
  const float* bT_b = biasT ? biasT + ((long long)h * n + (vb ? kbk : 0)) * n : nullptr;
  __nv_bfloat16* ds_a = ds_out ? ds_out + (((long long)win * heads + h) * n + (va ? ka : 0)) * n : nullptr;
  __nv_bfloat16* ds_b = ds_out ? ds_out + (((long long)win * heads + h) * n + (vb ? kbk : 0)) * n : nullptr;

  const int n_chunks = (n + QC - 1) / QC;
  for (int ch = 0; ch < n_chunks; ++ch) {
    const int q0 = ch * QC;
    __syncthreads();  // the previous chunk's readers of the staging buffers are done
    // ---- stage Q, dO (row-major and transposed), lse, delta of this query chunk
    if (tid < QC) {
      const int q = q0 + tid;
      const bool ok = q < n;
      const long long row = ok ? token_row(map, win, q, n) : 0;
      qrow_s[tid] = ok ? row : -1;
      lse_s[tid] = ok ? lse[row * heads + h] : INFINITY;
      delta_s[tid] = ok ? delta[row * heads + h] : 0.f;
    }
    __syncthreads();
    for (int i = tid; i < QC * (D / 8); i += NT) {
      const int j = i / (D / 8), vi = i % (D / 8);
      const long long row = qrow_s[j];
      uint4 qv = make_uint4(0, 0, 0, 0), gv = make_uint4(0, 0, 0, 0);
      if (row >= 0) {
        qv = *reinterpret_cast<const uint4*>(qkv + row * ld_qkv + h * D + vi * 8);
        gv = *reinterpret_cast<const uint4*>(dout + row * ldd + h * D + vi * 8);
      }
      *reinterpret_cast<uint4*>(Qs + j * LDQ + vi * 8) = qv;
      *reinterpret_cast<uint4*>(dOs + j * LDQ + vi * 8) = gv;
      const __nv_bfloat16* qe = reinterpret_cast<const __nv_bfloat16*>(&qv);
      const __nv_bfloat16* ge = reinterpret_cast<const __nv_bfloat16*>(&gv);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        Qt[(vi * 8 + e) * LDT + j] = qe[e];
        dOt[(vi * 8 + e) * LDT + j] = ge[e];
      }
    }
    __syncthreads();

    // ---- S^T and dP^T for 16 keys x 32 queries (four 8-query sub-tiles)
    float s[QC / 8][4], dp[QC / 8][4];
#pragma unroll
    for (int qs = 0; qs < QC / 8; ++qs) {
      s[qs][0] = s[qs][1] = s[qs][2] = s[qs][3] = 0.f;
      dp[qs][0] = dp[qs][1] = dp[qs][2] = dp[qs][3] = 0.f;
      const __nv_bfloat16* qp = Qs + (qs * 8 + g) * LDQ + 2 * t;
      const __nv_bfloat16* gp = dOs + (qs * 8 + g) * LDQ + 2 * t;
#pragma unroll
      for (int kk = 0; kk < D / 16; ++kk) {
        mma_bf16_16816(s[qs], kf[kk], *reinterpret_cast<const uint32_t*>(qp + kk * 16),
                       *reinterpret_cast<const uint32_t*>(qp + kk * 16 + 8));
        mma_bf16_16816(dp[qs], vf[kk], *reinterpret_cast<const uint32_t*>(gp + kk * 16),
                       *reinterpret_cast<const uint32_t*>(gp + kk * 16 + 8));
      }
    }
    // ---- P^T and dS^T (rows: keys g / g+8, columns: queries 2t, 2t+1 of each sub-tile)
    uint32_t pA[QC / 16][4], dsA[QC / 16][4];
#pragma unroll
    for (int qs = 0; qs < QC / 8; ++qs) {
      const int ql = qs * 8 + 2 * t;
      const int q = q0 + ql;
      float b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
      if (biasT != nullptr && q < n) {
        const float2 x = *reinterpret_cast<const float2*>(bT_a + q);
        const float2 y = *reinterpret_cast<const float2*>(bT_b + q);
        b0 = x.x * LOG2E; b1 = x.y * LOG2E; b2 = y.x * LOG2E; b3 = y.y * LOG2E;
      }
      const float l0 = lse_s[ql], l1 = lse_s[ql + 1];
      const float d0 = delta_s[ql], d1 = delta_s[ql + 1];
      const float p0 = va ? exp2f(fmaf(s[qs][0], sl, b0) - l0) : 0.f;
      const float p1 = va ? exp2f(fmaf(s[qs][1], sl, b1) - l1) : 0.f;
      const float p2 = vb ? exp2f(fmaf(s[qs][2], sl, b2) - l0) : 0.f;
      const float p3 = vb ? exp2f(fmaf(s[qs][3], sl, b3) - l1) : 0.f;
      const float e0 = p0 * (dp[qs][0] - d0), e1 = p1 * (dp[qs][1] - d1);
      const float e2 = p2 * (dp[qs][2] - d0), e3 = p3 * (dp[qs][3] - d1);
      pA[qs >> 1][(qs & 1) * 2 + 0] = pack_bf16x2(p0, p1);
      pA[qs >> 1][(qs & 1) * 2 + 1] = pack_bf16x2(p2, p3);
      const uint32_t e01 = pack_bf16x2(e0, e1), e23 = pack_bf16x2(e2, e3);
      dsA[qs >> 1][(qs & 1) * 2 + 0] = e01;
      dsA[qs >> 1][(qs & 1) * 2 + 1] = e23;
      if (ds_out != nullptr && q < n) {
        if (va) *reinterpret_cast<uint32_t*>(ds_a + q) = e01;
        if (vb) *reinterpret_cast<uint32_t*>(ds_b + q) = e23;
      }
    }
    // ---- dV += P^T dO, dK += dS^T Q  (contraction over the 32 queries)
#pragma unroll
    for (int kt = 0; kt < QC / 16; ++kt) {
#pragma unroll
      for (int dn = 0; dn < D / 8; ++dn) {
        const __nv_bfloat16* gp = dOt + (dn * 8 + g) * LDT + kt * 16 + 2 * t;
        mma_bf16_16816(dv[dn], pA[kt], *reinterpret_cast<const uint32_t*>(gp), *reinterpret_cast<const uint32_t*>(gp + 8));
        const __nv_bfloat16* qp = Qt + (dn * 8 + g) * LDT + kt * 16 + 2 * t;
        mma_bf16_16816(dk[dn], dsA[kt], *reinterpret_cast<const uint32_t*>(qp), *reinterpret_cast<const uint32_t*>(qp + 8));
      }
    }
    // ---- dS of this warp's 16 keys into the CTA-wide [query][key] tile
#pragma unroll
    for (int qs = 0; qs < QC / 8; ++qs) {
      const uint32_t u01 = dsA[qs >> 1][(qs & 1) * 2 + 0], u23 = dsA[qs >> 1][(qs & 1) * 2 + 1];
      const __nv_bfloat162 v01 = *reinterpret_cast<const __nv_bfloat162*>(&u01);
      const __nv_bfloat162 v23 = *reinterpret_cast<const __nv_bfloat162*>(&u23);
      __nv_bfloat16* dp_ = dSs + (qs * 8 + 2 * t) * LDS + warp * 16 + g;
      dp_[0] = v01.x;
      dp_[LDS] = v01.y;
      dp_[8] = v23.x;
      dp_[LDS + 8] = v23.y;
    }
    __syncthreads();
    // ---- dQ of this chunk = dS (32 x keys) . K (keys x D): 2 x D/8 output tiles of 16 x 8 spread over the warps, the
    // whole key range of the CTA contracted by each (no cross-warp reduction).  With every key of the (window, head) in
    // this CTA (DIRECT_DQ) that is the final value, written as bf16; otherwise a partial sum added to the fp32 buffer.
    {
      constexpr int TILES = 2 * (D / 8);
      constexpr int TPW = TILES / NW > 0 ? TILES / NW : 1;  // tiles per warp (all in one 16-query row block)
      const int tile0 = warp * TPW;
      if (tile0 < TILES) {
        const int mt = tile0 / (D / 8), dn0 = tile0 % (D / 8);
        float acc[TPW][4];
#pragma unroll
        for (int i = 0; i < TPW; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
        const __nv_bfloat16* ap = dSs + (mt * 16 + g) * LDS + 2 * t;
#pragma unroll
        for (int ks = 0; ks < NW; ++ks) {
          uint32_t af[4];
          af[0] = *reinterpret_cast<const uint32_t*>(ap + ks * 16);
          af[1] = *reinterpret_cast<const uint32_t*>(ap + 8 * LDS + ks * 16);
          af[2] = *reinterpret_cast<const uint32_t*>(ap + ks * 16 + 8);
          af[3] = *reinterpret_cast<const uint32_t*>(ap + 8 * LDS + ks * 16 + 8);
#pragma unroll
          for (int i = 0; i < TPW; ++i) {
            const __nv_bfloat16* bp = Kt + ((dn0 + i) * 8 + g) * LDS + ks * 16 + 2 * t;
            mma_bf16_16816(acc[i], af, *reinterpret_cast<const uint32_t*>(bp), *reinterpret_cast<const uint32_t*>(bp + 8));
          }
        }
        const long long row_a = qrow_s[mt * 16 + g], row_b = qrow_s[mt * 16 + g + 8];
#pragma unroll
        for (int i = 0; i < TPW; ++i) {
          const int c = h * D + (dn0 + i) * 8 + 2 * t;
          if constexpr (DIRECT_DQ) {
            if (row_a >= 0) *reinterpret_cast<uint32_t*>(dqkv + row_a * ld_dqkv + c) = pack_bf16x2(acc[i][0] * scale, acc[i][1] * scale);
            if (row_b >= 0) *reinterpret_cast<uint32_t*>(dqkv + row_b * ld_dqkv + c) = pack_bf16x2(acc[i][2] * scale, acc[i][3] * scale);
          } else {
            if (row_a >= 0)
              asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(dq_f32 + row_a * C + c), "f"(acc[i][0] * scale),
                           "f"(acc[i][1] * scale) : "memory");
            if (row_b >= 0)
              asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(dq_f32 + row_b * C + c), "f"(acc[i][2] * scale),
                           "f"(acc[i][3] * scale) : "memory");
          }
        }
      }
    }
  }

  // ---- dK (scaled) and dV rows of this warp's keys
#pragma unroll
  for (int dn = 0; dn < D / 8; ++dn) {
    const int c = h * D + dn * 8 + 2 * t;
    if (va) {
      *reinterpret_cast<uint32_t*>(dqkv + ra * ld_dqkv + C + c) = pack_bf16x2(dk[dn][0] * scale, dk[dn][1] * scale);
      *reinterpret_cast<uint32_t*>(dqkv + ra * ld_dqkv + 2 * C + c) = pack_bf16x2(dv[dn][0], dv[dn][1]);
    }
    if (vb) {
      *reinterpret_cast<uint32_t*>(dqkv + rb * ld_dqkv + C + c) = pack_bf16x2(dk[dn][2] * scale, dk[dn][3] * scale);
      *reinterpret_cast<uint32_t*>(dqkv + rb * ld_dqkv + 2 * C + c) = pack_bf16x2(dv[dn][2], dv[dn][3]);
    }
  }
}


// ------------------------------------------------------------------------------------------------
// "Resident" variant: a CTA of NW warps owns 16 * NW keys of one (window, head) and keeps Q, dO (row-major), lse and
// delta of EVERY query of that (window, head) in shared memory, staged once; the query chunks then run without further
// global loads and with ONE block barrier each (the [query][key] dS tile is double buffered).  All transposed operands
// (dO and Q for dV / dK, K for dQ) are read from the row-major tiles with ldmatrix.trans.
//   6^3 windows (n <= 224, D = 32): NW = 14, one CTA holds every key too -> dQ is final, written as bf16 (DIRECT).
//   ViT (n = 432, D = 64): NW = 5 (80 keys, 6 CTAs per (batch, head) = 144 CTAs, one wave); partial dQ -> fp32 reductions.
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

template <int D, int NW>
__host__ __device__ constexpr size_t resident_smem_bytes(int n_pad) {
  return (size_t)(2 * n_pad * (D + 8) + 16 * NW * (D + 8) + 2 * QC * (16 * NW + 8)) * 2 + (size_t)n_pad * (4 + 4 + 8);
}

template <int D, int NW, bool DIRECT, int NPAD>
__global__ void __launch_bounds__(32 * NW) attention_bwd_resident_kernel(
    const __nv_bfloat16* __restrict__ qkv, int ld_qkv, int C, const __nv_bfloat16* __restrict__ dout, int ldd,
    const float* __restrict__ lse, const float* __restrict__ delta, const float* __restrict__ biasT, float scale,
    __nv_bfloat16* __restrict__ dqkv, int ld_dqkv, float* __restrict__ dq_f32, __nv_bfloat16* __restrict__ ds_out, int n,
    TokenMap map) {
  constexpr int n_pad = NPAD;        // rows of the resident query tiles (n <= NPAD, zero-filled beyond n)
  constexpr int LD = D + 8;          // pitch of the [token][d] tiles (80 / 144 bytes: 16-byte aligned, conflict-free)
  constexpr int KB = 16 * NW;        // keys of this CTA
  constexpr int LS = KB + 8;         // pitch of the [query][key] dS tile (LS/2 = 4 mod 8: conflict-free fragments)
  constexpr int NT = 32 * NW;
  constexpr int VPR = D / 8;
  extern __shared__ __align__(16) uint8_t smem_bw[];
  long long* qrow_s = reinterpret_cast<long long*>(smem_bw);       // [n_pad]
  float* lse_s = reinterpret_cast<float*>(qrow_s + n_pad);         // [n_pad]
  float* delta_s = lse_s + n_pad;                                  // [n_pad]
  __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(delta_s + n_pad);  // [n_pad][LD]
  __nv_bfloat16* dOs = Qs + (size_t)n_pad * LD;                    // [n_pad][LD]
  __nv_bfloat16* Ks = dOs + (size_t)n_pad * LD;                    // [KB][LD]
  __nv_bfloat16* dSs = Ks + KB * LD;                               // [2][QC][LS]

  const int kb = blockIdx.x, h = blockIdx.y, win = blockIdx.z;
  const int heads = C / D;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const float LOG2E = 1.4426950408889634f;
  const float sl = scale * LOG2E;

  // ---- stage every query of the (window, head) and this CTA's keys
  for (int i = tid; i < n_pad; i += NT) {
    const bool ok = i < n;
    const long long row = ok ? token_row(map, win, i, n) : 0;
    qrow_s[i] = ok ? row : -1;
    lse_s[i] = ok ? lse[row * heads + h] : INFINITY;
    delta_s[i] = ok ? delta[row * heads + h] : 0.f;
  }
  for (int i = tid; i < n_pad * VPR; i += NT) {
    const int j = i / VPR, vi = i % VPR;
    uint4 qv = make_uint4(0, 0, 0, 0), gv = qv;
    if (j < n) {
      const long long row = token_row(map, win, j, n);
      qv = *reinterpret_cast<const uint4*>(qkv + row * ld_qkv + h * D + vi * 8);
      gv = *reinterpret_cast<const uint4*>(dout + row * ldd + h * D + vi * 8);
    }
    *reinterpret_cast<uint4*>(Qs + (size_t)j * LD + vi * 8) = qv;
    *reinterpret_cast<uint4*>(dOs + (size_t)j * LD + vi * 8) = gv;
  }
  for (int i = tid; i < KB * VPR; i += NT) {
    const int j = i / VPR, vi = i % VPR;
    const int key = kb * KB + j;
    uint4 kv = make_uint4(0, 0, 0, 0);
    if (key < n) kv = *reinterpret_cast<const uint4*>(qkv + token_row(map, win, key, n) * ld_qkv + C + h * D + vi * 8);
    *reinterpret_cast<uint4*>(Ks + j * LD + vi * 8) = kv;
  }

  // ---- this warp's 16 keys: K and V as A fragments (rows = keys)
  const int key0 = kb * KB + warp * 16;
  const int ka = key0 + g, kbk = key0 + g + 8;
  const bool va = ka < n, vb = kbk < n;
  const long long ra = va ? token_row(map, win, ka, n) : 0;
  const long long rb = vb ? token_row(map, win, kbk, n) : 0;
  uint32_t kf[D / 16][4], vf[D / 16][4];
#pragma unroll
  for (int kk = 0; kk < D / 16; ++kk) {
    const int c = h * D + kk * 16 + 2 * t;
    kf[kk][0] = va ? *reinterpret_cast<const uint32_t*>(qkv + ra * ld_qkv + C + c) : 0u;
    kf[kk][1] = vb ? *reinterpret_cast<const uint32_t*>(qkv + rb * ld_qkv + C + c) : 0u;
    kf[kk][2] = va ? *reinterpret_cast<const uint32_t*>(qkv + ra * ld_qkv + C + c + 8) : 0u;
    kf[kk][3] = vb ? *reinterpret_cast<const uint32_t*>(qkv + rb * ld_qkv + C + c + 8) : 0u;
    vf[kk][0] = va ? *reinterpret_cast<const uint32_t*>(qkv + ra * ld_qkv + 2 * C + c) : 0u;
    vf[kk][1] = vb ? *reinterpret_cast<const uint32_t*>(qkv + rb * ld_qkv + 2 * C + c) : 0u;
    vf[kk][2] = va ? *reinterpret_cast<const uint32_t*>(qkv + ra * ld_qkv + 2 * C + c + 8) : 0u;
    vf[kk][3] = vb ? *reinterpret_cast<const uint32_t*>(qkv + rb * ld_qkv + 2 * C + c + 8) : 0u;
  }
  float dk[D / 8][4], dv[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i) { dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f; dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f; }
  const float* bT_a = biasT ? biasT + ((long long)h * n + (va ? ka : 0)) * n : nullptr;
  const float* bT_b = biasT ? biasT + ((long long)h * n + (vb ? kbk : 0)) * n : nullptr;
  __nv_bfloat16* ds_a = ds_out ? ds_out + (((long long)win * heads + h) * n + (va ? ka : 0)) * n : nullptr;
  __nv_bfloat16* ds_b = ds_out ? ds_out + (((long long)win * heads + h) * n + (vb ? kbk : 0)) * n : nullptr;
  __syncthreads();

  const int n_chunks = (n + QC - 1) / QC;
  for (int ch = 0; ch < n_chunks; ++ch) {
    const int q0 = ch * QC;
    __nv_bfloat16* dS = dSs + (ch & 1) * QC * LS;
    // ---- S^T and dP^T for 16 keys x 32 queries
    float s[QC / 8][4], dp[QC / 8][4];
#pragma unroll
    for (int qs = 0; qs < QC / 8; ++qs) {
      s[qs][0] = s[qs][1] = s[qs][2] = s[qs][3] = 0.f;
      dp[qs][0] = dp[qs][1] = dp[qs][2] = dp[qs][3] = 0.f;
      const __nv_bfloat16* qp = Qs + (size_t)(q0 + qs * 8 + g) * LD + 2 * t;
      const __nv_bfloat16* gp = dOs + (size_t)(q0 + qs * 8 + g) * LD + 2 * t;
#pragma unroll
      for (int kk = 0; kk < D / 16; ++kk) {
        mma_bf16_16816(s[qs], kf[kk], *reinterpret_cast<const uint32_t*>(qp + kk * 16),
                       *reinterpret_cast<const uint32_t*>(qp + kk * 16 + 8));
        mma_bf16_16816(dp[qs], vf[kk], *reinterpret_cast<const uint32_t*>(gp + kk * 16),
                       *reinterpret_cast<const uint32_t*>(gp + kk * 16 + 8));
      }
    }
    // ---- P^T and dS^T (rows: keys g / g+8, columns: queries 2t, 2t+1 of each sub-tile)
    uint32_t pA[QC / 16][4], dsA[QC / 16][4];
#pragma unroll
    for (int qs = 0; qs < QC / 8; ++qs) {
      const int ql = qs * 8 + 2 * t;
      const int q = q0 + ql;
      float b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
      if (biasT != nullptr && q < n) {
        const float2 x = *reinterpret_cast<const float2*>(bT_a + q);
        const float2 y = *reinterpret_cast<const float2*>(bT_b + q);
        b0 = x.x * LOG2E; b1 = x.y * LOG2E; b2 = y.x * LOG2E; b3 = y.y * LOG2E;
      }
      const float l0 = lse_s[q], l1 = lse_s[q + 1];
      const float d0 = delta_s[q], d1 = delta_s[q + 1];
      const float p0 = va ? exp2f(fmaf(s[qs][0], sl, b0) - l0) : 0.f;
      const float p1 = va ? exp2f(fmaf(s[qs][1], sl, b1) - l1) : 0.f;
      const float p2 = vb ? exp2f(fmaf(s[qs][2], sl, b2) - l0) : 0.f;
      const float p3 = vb ? exp2f(fmaf(s[qs][3], sl, b3) - l1) : 0.f;
      const float e0 = p0 * (dp[qs][0] - d0), e1 = p1 * (dp[qs][1] - d1);
      const float e2 = p2 * (dp[qs][2] - d0), e3 = p3 * (dp[qs][3] - d1);
      pA[qs >> 1][(qs & 1) * 2 + 0] = pack_bf16x2(p0, p1);
      pA[qs >> 1][(qs & 1) * 2 + 1] = pack_bf16x2(p2, p3);
      const uint32_t e01 = pack_bf16x2(e0, e1), e23 = pack_bf16x2(e2, e3);
      dsA[qs >> 1][(qs & 1) * 2 + 0] = e01;
      dsA[qs >> 1][(qs & 1) * 2 + 1] = e23;
      if (ds_out != nullptr && q < n) {
        if (va) *reinterpret_cast<uint32_t*>(ds_a + q) = e01;
        if (vb) *reinterpret_cast<uint32_t*>(ds_b + q) = e23;
      }
      // dS of this warp's keys into the CTA-wide [query][key] tile
      const __nv_bfloat162 v01 = *reinterpret_cast<const __nv_bfloat162*>(&e01);
      const __nv_bfloat162 v23 = *reinterpret_cast<const __nv_bfloat162*>(&e23);
      __nv_bfloat16* dp_ = dS + ql * LS + warp * 16 + g;
      dp_[0] = v01.x;
      dp_[LS] = v01.y;
      dp_[8] = v23.x;
      dp_[LS + 8] = v23.y;
    }
    // ---- dV += P^T dO, dK += dS^T Q (contraction over the 32 queries; B fragments by ldmatrix.trans)
#pragma unroll
    for (int kt = 0; kt < QC / 16; ++kt) {
      const size_t off = (size_t)(q0 + kt * 16 + (lane & 15)) * LD + (lane >> 4) * 8;
      const uint32_t ga = smem_u32(dOs + off), qa = smem_u32(Qs + off);
#pragma unroll
      for (int dn = 0; dn < D / 8; dn += 2) {
        uint32_t bg[4], bq[4];
        ldsm_x4_trans(bg, ga + dn * 16);
        ldsm_x4_trans(bq, qa + dn * 16);
        mma_bf16_16816(dv[dn], pA[kt], bg[0], bg[1]);
        mma_bf16_16816(dv[dn + 1], pA[kt], bg[2], bg[3]);
        mma_bf16_16816(dk[dn], dsA[kt], bq[0], bq[1]);
        mma_bf16_16816(dk[dn + 1], dsA[kt], bq[2], bq[3]);
      }
    }
    __syncthreads();
    // ---- dQ of this chunk = dS (32 x KB keys) . K (KB keys x D): 2 x D/8 output tiles of 16 x 8 spread over the warps
    {
      constexpr int TILES = 2 * (D / 8);
      constexpr int TPW = (TILES + NW - 1) / NW;  // tiles per warp (all in one 16-query row block)
      static_assert((D / 8) % TPW == 0, "a warp's tiles must share their row block");
      const int tile0 = warp * TPW;
      if (tile0 < TILES) {
        const int mt = tile0 / (D / 8), dn0 = tile0 % (D / 8);
        float acc[TPW][4];
#pragma unroll
        for (int i = 0; i < TPW; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
        const __nv_bfloat16* ap = dS + (mt * 16 + g) * LS + 2 * t;
        const uint32_t kbase = smem_u32(Ks + (lane & 15) * LD + dn0 * 8);
#pragma unroll
        for (int ks = 0; ks < NW; ++ks) {
          uint32_t af[4];
          af[0] = *reinterpret_cast<const uint32_t*>(ap + ks * 16);
          af[1] = *reinterpret_cast<const uint32_t*>(ap + 8 * LS + ks * 16);
          af[2] = *reinterpret_cast<const uint32_t*>(ap + ks * 16 + 8);
          af[3] = *reinterpret_cast<const uint32_t*>(ap + 8 * LS + ks * 16 + 8);
#pragma unroll
          for (int i = 0; i < TPW; ++i) {
            uint32_t b0, b1;
            asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];"
                         : "=r"(b0), "=r"(b1)
                         : "r"(kbase + (uint32_t)(ks * 16 * LD * 2 + i * 16)));
            mma_bf16_16816(acc[i], af, b0, b1);
          }
        }
        const long long row_a = qrow_s[q0 + mt * 16 + g], row_b = qrow_s[q0 + mt * 16 + g + 8];
#pragma unroll
        for (int i = 0; i < TPW; ++i) {
          const int c = h * D + (dn0 + i) * 8 + 2 * t;
          if constexpr (DIRECT) {
            if (row_a >= 0) *reinterpret_cast<uint32_t*>(dqkv + row_a * ld_dqkv + c) = pack_bf16x2(acc[i][0] * scale, acc[i][1] * scale);
            if (row_b >= 0) *reinterpret_cast<uint32_t*>(dqkv + row_b * ld_dqkv + c) = pack_bf16x2(acc[i][2] * scale, acc[i][3] * scale);
          } else {
            if (row_a >= 0)
              asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(dq_f32 + row_a * C + c), "f"(acc[i][0] * scale),
                           "f"(acc[i][1] * scale) : "memory");
            if (row_b >= 0)
              asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(dq_f32 + row_b * C + c), "f"(acc[i][2] * scale),
                           "f"(acc[i][3] * scale) : "memory");
          }
        }
      }
    }
  }

  // ---- dK (scaled) and dV rows of this warp's keys
#pragma unroll
  for (int dn = 0; dn < D / 8; ++dn) {
    const int c = h * D + dn * 8 + 2 * t;
    if (va) {
      *reinterpret_cast<uint32_t*>(dqkv + ra * ld_dqkv + C + c) = pack_bf16x2(dk[dn][0] * scale, dk[dn][1] * scale);
      *reinterpret_cast<uint32_t*>(dqkv + ra * ld_dqkv + 2 * C + c) = pack_bf16x2(dv[dn][0], dv[dn][1]);
    }
    if (vb) {
      *reinterpret_cast<uint32_t*>(dqkv + rb * ld_dqkv + C + c) = pack_bf16x2(dk[dn][2] * scale, dk[dn][3] * scale);
      *reinterpret_cast<uint32_t*>(dqkv + rb * ld_dqkv + 2 * C + c) = pack_bf16x2(dv[dn][2], dv[dn][3]);
    }
  }
}

}  // namespace ctu

using namespace ctu;

extern "C" int ctu_attention_delta(const void* o, long long ldo, const void* dout, long long ldd, float* delta,
                                   long long rows, int C, int dim_head, void* stream) {
  if (!o || !dout || !delta || rows <= 0 || (dim_head != 32 && dim_head != 64) || C % dim_head || ldo % 8 || ldd % 8)
    return CTU_E_BADARG;
  const long long total = rows * (C / 8);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  attn_delta_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)o, ldo,
                                                                        (const __nv_bfloat16*)dout, ldd, delta, rows, C,
                                                                        dim_head);
  count_launch();
  return (int)cudaGetLastError();
}

// qkv / dqkv: bf16 [rows][3C] (q|k|v); dqkv receives dK and dV, dQ is ACCUMULATED into dq_f32 (fp32 [rows][C],
// zeroed by the caller) — except for dim_head 32 with n <= 224 (the 6^3 windows), where dQ is written straight into
// dqkv[:, :C] and dq_f32 may be NULL.  biasT: fp32 [heads][n][n] indexed [key][query] or NULL; ds_out: bf16
// [windows][heads][n][n] indexed [key][query] or NULL.
extern "C" int ctu_attention_bwd(const void* qkv, int ld_qkv, int C, int dim_head, const void* dout, int ldd,
                                 const float* lse, const float* delta, const float* biasT, void* dqkv, int ld_dqkv,
                                 float* dq_f32, void* ds_out, int n, int windows, int mode, int batch, int X, int Y, int Z,
                                 int w, void* stream) {
  if (!qkv || !dout || !lse || !delta || !dqkv) return CTU_E_BADARG;
  if ((dim_head != 32 && dim_head != 64) || C % dim_head || ld_qkv % 8 || ldd % 8 || ld_dqkv % 8 || C % 4) return CTU_E_BADARG;
  if ((biasT || ds_out) && (n % 2)) return CTU_E_BADARG;
  TokenMap m;
  m.mode = mode; m.X = X; m.Y = Y; m.Z = Z; m.w = w;
  m.nwx = m.nwy = m.nwz = 1;
  if (mode != 0) {
    if (w <= 0 || X % w || Y % w || Z % w || n != w * w * w) return CTU_E_BADARG;
    m.nwx = X / w; m.nwy = Y / w; m.nwz = Z / w;
    windows = batch * m.nwx * m.nwy * m.nwz;
  }
  if (windows <= 0 || windows > 65535) return CTU_E_BADARG;
  const int heads = C / dim_head;
  const float scale = 1.0f / sqrtf((float)dim_head);
  cudaStream_t st = (cudaStream_t)stream;
  const __nv_bfloat16* q = (const __nv_bfloat16*)qkv;
  const __nv_bfloat16* go = (const __nv_bfloat16*)dout;
  __nv_bfloat16* dq = (__nv_bfloat16*)dqkv;
  __nv_bfloat16* ds = (__nv_bfloat16*)ds_out;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attention_bwd_resident_kernel<32, 14, true, 224>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)resident_smem_bytes<32, 14>(224));
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(attention_bwd_resident_kernel<64, 5, false, 448>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)resident_smem_bytes<64, 5>(448));
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  if (dim_head == 32 && n <= 224) {
    // 6x6x6 windows: one CTA of 14 warps owns every key and every query of a (window, head): everything is staged once
    // and dQ needs neither global reductions nor the fp32 buffer
    dim3 grid(1, heads, windows);
    attention_bwd_resident_kernel<32, 14, true, 224><<<grid, 32 * 14, resident_smem_bytes<32, 14>(224), st>>>(
        q, ld_qkv, C, go, ldd, lse, delta, biasT, scale, dq, ld_dqkv, nullptr, ds, n, m);
  } else if (dim_head == 64 && n <= 448) {
    // ViT: 80 keys per CTA, every query resident (144 CTAs at n = 432, batch 2: one wave)
    if (!dq_f32) return CTU_E_BADARG;
    dim3 grid((n + 79) / 80, heads, windows);
    attention_bwd_resident_kernel<64, 5, false, 448><<<grid, 32 * 5, resident_smem_bytes<64, 5>(448), st>>>(
        q, ld_qkv, C, go, ldd, lse, delta, biasT, scale, dq, ld_dqkv, dq_f32, ds, n, m);
  } else {
    if (!dq_f32) return CTU_E_BADARG;
    dim3 grid((n + 63) / 64, heads, windows);
    if (dim_head == 64)
      attention_bwd_kernel<64, 4, false><<<grid, 128, 0, st>>>(q, ld_qkv, C, go, ldd, lse, delta, biasT, scale, dq, ld_dqkv,
                                                               dq_f32, ds, n, m);
    else
      attention_bwd_kernel<32, 4, false><<<grid, 128, 0, st>>>(q, ld_qkv, C, go, ldd, lse, delta, biasT, scale, dq, ld_dqkv,
                                                               dq_f32, ds, n, m);
  }
  count_launch();
  return (int)cudaGetLastError();
}
