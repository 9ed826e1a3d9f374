// tcgen05 / TMEM / TMA contraction kernel for sm_100a: plain GEMM, 3x3x3 implicit-GEMM Conv3d and
// kernel==stride ConvTranspose3d share one warp-specialised kernel (see include/ctunet_b200.h).
//
//   warp 0 (1 lane) : TMA producer — per K block one 5-D box of activations (128 voxels x 64 channels, shifted
//                     by the filter tap, out-of-range voxels zero-filled by the TMA unit = conv padding) and
//                     one 2-D box of packed weights (BLOCK_N rows x 64).
//   warp 1 (1 lane) : MMA issuer — 4 x tcgen05.mma (M128 x BLOCK_N x K16) per K block into a TMEM accumulator,
//                     tcgen05.commit releases the smem stage / signals the epilogue.
//   warps 2..5      : epilogue — tcgen05.ld the accumulator (one voxel row per thread), bias / GELU / residual,
//                     InstanceNorm partial statistics (warp transpose-reduce + fp64 atomics), bf16/fp32 stores.
// One output tile (128 x BLOCK_N) per CTA; several CTAs are co-resident per SM so one CTA's epilogue overlaps
// another's main loop.
#include "common.cuh"
#include "../../include/ctunet_b200.h"
#include "host_util.h"

namespace ctu {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;

struct GemmParams {
  int b1, b2, b3;
  int T1, T2, T3;
  int d1, d2, d3;
  int n_tiles;
  int num_kb, cblocks, a_c;
  int k1, k2;
  int pad;
  // epilogue
  void* out;
  const float* bias;
  const void* residual;
  double* stats;
  int n_real, out_mode, ldc, act, res_mode, ldr;
  int convt_cout, u1, u2, u3;
  int stats_ld, out_col0;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// Column sums over the 32 lanes of a warp for 32 per-lane values: after the call, lane L holds the sum of
// v[L] over all lanes (31 shuffles instead of 160).
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16, n = 16; off >= 1; off >>= 1, n >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      const float send = upper ? v[i] : v[i + n];
      const float keep = upper ? v[i + n] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(192) umma_gemm_kernel(const __grid_constant__ CUtensorMap tmA,
                                                        const __grid_constant__ CUtensorMap tmB,
                                                        const GemmParams p) {
  constexpr int B_STAGE_BYTES = BN * BLOCK_K * 2;
  constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
  constexpr int CH = BN < 32 ? BN : 32;  // columns per tcgen05.ld
  constexpr uint32_t IDESC = umma_idesc_bf16(BLOCK_M, BN);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + STAGES * B_STAGE_BYTES);
  // bars[0..STAGES) full, [STAGES..2*STAGES) empty, [2*STAGES] accumulator ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);
  float2* stat_scratch = reinterpret_cast<float2*>(tmem_slot + 2);  // [4][BN]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&bars[s]), 1);
      mbar_init(smem_u32(&bars[STAGES + s]), 1);
    }
    mbar_init(smem_u32(&bars[2 * STAGES]), 1);
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

  // tile decode
  const int n_tile = blockIdx.x % p.n_tiles;
  int m_tile = blockIdx.x / p.n_tiles;
  const int t1 = m_tile % p.T1;
  m_tile /= p.T1;
  const int t2 = m_tile % p.T2;
  m_tile /= p.T2;
  const int t3 = m_tile % p.T3;
  const int t4 = m_tile / p.T3;
  const int x1 = t1 * p.b1, x2 = t2 * p.b2, x3 = t3 * p.b3;
  const int n0 = n_tile * BN;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < p.num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(smem_u32(&bars[STAGES + s]), ph ^ 1);
        const uint32_t full = smem_u32(&bars[s]);
        mbar_expect_tx(full, A_STAGE_BYTES + B_STAGE_BYTES);
        const int tap = kb / p.cblocks;
        const int cb = kb - tap * p.cblocks;
        const int f1 = tap % p.k1;
        const int f2 = (tap / p.k1) % p.k2;
        const int f3 = tap / (p.k1 * p.k2);
        tma_load_5d(smem_u32(smem_a + s * A_STAGE_BYTES), &tmA, full, cb * BLOCK_K, x1 + f1 - p.pad,
                    x2 + f2 - p.pad, x3 + f3 - p.pad, t4);
        tma_load_2d(smem_u32(smem_b + s * B_STAGE_BYTES), &tmB, full, tap * p.a_c + cb * BLOCK_K, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      for (int kb = 0; kb < p.num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(smem_u32(&bars[s]), ph);
        tc_fence_after();
        const uint64_t da = umma_desc_k_sw128(smem_u32(smem_a + s * A_STAGE_BYTES));
        const uint64_t db = umma_desc_k_sw128(smem_u32(smem_b + s * B_STAGE_BYTES));
#pragma unroll
        for (int k = 0; k < BLOCK_K / 16; ++k) {
          // advance 16 bf16 = 32 bytes inside the 128-byte swizzled row: +2 in the (addr >> 4) field
          umma_bf16(tmem_base, da + 2 * k, db + 2 * k, IDESC, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(smem_u32(&bars[STAGES + s]));
      }
      umma_commit(smem_u32(&bars[2 * STAGES]));
    }
  } else {
    // ------------------------------------------------------------------ epilogue
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int r = q * 32 + lane;
    const int i1 = r % p.b1;
    const int i2 = (r / p.b1) % p.b2;
    const int i3 = r / (p.b1 * p.b2);
    const int v1 = x1 + i1, v2 = x2 + i2, v3 = x3 + i3;
    const bool valid = (v1 < p.d1) && (v2 < p.d2) && (v3 < p.d3);

    int a1 = 0, a2 = 0, a3 = 0, colbase = n0;
    if (p.convt_cout > 0) {
      const int sub = n0 / p.convt_cout;
      colbase = n0 - sub * p.convt_cout;
      a1 = sub % p.u1;
      a2 = (sub / p.u1) % p.u2;
      a3 = sub / (p.u1 * p.u2);
    }
    const long long o1 = (long long)p.d1 * p.u1, o2 = (long long)p.d2 * p.u2, o3 = (long long)p.d3 * p.u3;
    const long long out_row = ((t4 * o3 + (long long)v3 * p.u3 + a3) * o2 + ((long long)v2 * p.u2 + a2)) * o1 +
                              ((long long)v1 * p.u1 + a1);

    mbar_wait(smem_u32(&bars[2 * STAGES]), 0);
    tc_fence_after();

#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += CH) {
      uint32_t raw[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
      if constexpr (CH == 32) tmem_ld32(taddr, raw);
      else tmem_ld16(taddr, raw);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int j = 0; j < CH; ++j) v[j] = __uint_as_float(raw[j]);
#pragma unroll
      for (int j = CH; j < 32; ++j) v[j] = 0.f;

      const int gcol = n0 + c0;  // column in the GEMM's N space (bias / stats / n_real)
      if (p.bias != nullptr) {
#pragma unroll
        for (int j = 0; j < CH; ++j)
          if (gcol + j < p.n_real) v[j] += __ldg(p.bias + gcol + j);
      }
      if (p.act == CTU_ACT_GELU) {
#pragma unroll
        for (int j = 0; j < CH; ++j) v[j] = gelu_erf(v[j]);
      }
      const int ocol = p.out_col0 + colbase + c0;  // column inside an output row
      if (p.res_mode == CTU_RES_F32 && valid) {
        const float* rp = reinterpret_cast<const float*>(p.residual) + out_row * p.ldr + ocol;
#pragma unroll
        for (int j = 0; j < CH; j += 4) {
          if (gcol + j < p.n_real) {
            const float4 rv = *reinterpret_cast<const float4*>(rp + j);
            v[j] += rv.x; v[j + 1] += rv.y; v[j + 2] += rv.z; v[j + 3] += rv.w;
          }
        }
      } else if (p.res_mode == CTU_RES_BF16 && valid) {
        const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(p.residual) + out_row * p.ldr + ocol;
#pragma unroll
        for (int j = 0; j < CH; j += 8) {
          if (gcol + j < p.n_real) {
            const uint4 rv = *reinterpret_cast<const uint4*>(rp + j);
            float2 f;
            f = unpack_bf16x2(rv.x); v[j] += f.x; v[j + 1] += f.y;
            f = unpack_bf16x2(rv.y); v[j + 2] += f.x; v[j + 3] += f.y;
            f = unpack_bf16x2(rv.z); v[j + 4] += f.x; v[j + 5] += f.y;
            f = unpack_bf16x2(rv.w); v[j + 6] += f.x; v[j + 7] += f.y;
          }
        }
      }

      if (p.out_mode == CTU_OUT_BF16_ROWS) {
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < CH / 2; ++j) pk[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
        if (valid) {
          __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + out_row * p.ldc + ocol;
#pragma unroll
          for (int j = 0; j < CH; j += 8) {
            if (gcol + j < p.n_real)
              *reinterpret_cast<uint4*>(op + j) = make_uint4(pk[j / 2], pk[j / 2 + 1], pk[j / 2 + 2], pk[j / 2 + 3]);
          }
        }
        if (p.stats != nullptr) {
          // statistics of the values as stored (bf16-rounded), masked rows contribute zero
#pragma unroll
          for (int j = 0; j < CH / 2; ++j) {
            const float2 f = unpack_bf16x2(pk[j]);
            v[2 * j] = valid ? f.x : 0.f;
            v[2 * j + 1] = valid ? f.y : 0.f;
          }
        }
      } else if (p.out_mode == CTU_OUT_F32_ROWS) {
        if (valid) {
          float* op = reinterpret_cast<float*>(p.out) + out_row * p.ldc + ocol;
#pragma unroll
          for (int j = 0; j < CH; j += 4) {
            if (gcol + j < p.n_real) *reinterpret_cast<float4*>(op + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
        }
        if (p.stats != nullptr) {
#pragma unroll
          for (int j = 0; j < CH; ++j) v[j] = valid ? v[j] : 0.f;
        }
      } else {  // CTU_OUT_F32_CF: lanes hold consecutive voxels, so each column is a coalesced 128-byte store
        if (valid) {
          const long long S = (long long)p.d1 * p.d2 * p.d3;
          const long long s_idx = ((long long)v3 * p.d2 + v2) * p.d1 + v1;
          float* op = reinterpret_cast<float*>(p.out) + ((long long)t4 * p.n_real) * S + s_idx;
#pragma unroll
          for (int j = 0; j < CH; ++j)
            if (gcol + j < p.n_real) op[(long long)(gcol + j) * S] = v[j];
        }
      }

      if (p.stats != nullptr) {
        float sq[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) sq[j] = v[j] * v[j];
        const float s_sum = warp_transpose_reduce(v, lane);
        const float s_sq = warp_transpose_reduce(sq, lane);
        if (lane < CH) stat_scratch[q * BN + c0 + lane] = make_float2(s_sum, s_sq);
      }
    }

    if (p.stats != nullptr) {
      named_bar_sync(1, 128);
      const int e = threadIdx.x - 64;  // 0..127
      for (int c = e; c < BN; c += 128) {
        if (n0 + c < p.n_real) {
          const float2 s0 = stat_scratch[c], s1 = stat_scratch[BN + c], s2 = stat_scratch[2 * BN + c],
                       s3 = stat_scratch[3 * BN + c];
          double* dst = p.stats + ((long long)t4 * p.stats_ld + n0 + c) * 2;
          atomicAdd(dst, (double)s0.x + (double)s1.x + (double)s2.x + (double)s3.x);
          atomicAdd(dst + 1, (double)s0.y + (double)s1.y + (double)s2.y + (double)s3.y);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int BN, int STAGES>
constexpr int gemm_smem_bytes() {
  return 1024 + STAGES * (A_STAGE_BYTES + BN * BLOCK_K * 2) + (2 * STAGES + 1) * 8 + 16 + 4 * BN * 8;
}

template <int BN, int STAGES>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, int grid,
                       cudaStream_t stream) {
  constexpr int smem = gemm_smem_bytes<BN, STAGES>();
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(umma_gemm_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  umma_gemm_kernel<BN, STAGES><<<grid, 192, smem, stream>>>(tmA, tmB, p);
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace ctu

extern "C" int ctu_umma_gemm(const ctu_gemm_desc* d, void* stream_) {
  using namespace ctu;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!d || !d->a || !d->w || !d->out) return CTU_E_BADARG;
  if (d->b1 * d->b2 * d->b3 != BLOCK_M) return CTU_E_BADARG;
  if (d->b1 > 256 || d->b2 > 256 || d->b3 > 256) return CTU_E_BADARG;
  const int taps = d->k1 * d->k2 * d->k3;
  if (!((d->k1 == 1 && d->k2 == 1 && d->k3 == 1) || (d->k1 == 3 && d->k2 == 3 && d->k3 == 3))) return CTU_E_UNSUPPORTED;
  if (taps > 1 && (d->a_c % BLOCK_K) != 0) return CTU_E_UNSUPPORTED;
  if (d->k_total != taps * d->a_c) return CTU_E_BADARG;
  if ((d->lda % 8) != 0 || (d->k_total % 8) != 0 || (d->a_c % 8) != 0) return CTU_E_BADARG;
  if (d->n_pad % d->block_n != 0 || d->n_real > d->n_pad || d->n_real <= 0) return CTU_E_BADARG;
  if ((reinterpret_cast<uintptr_t>(d->a) & 15) || (reinterpret_cast<uintptr_t>(d->w) & 15) ||
      (reinterpret_cast<uintptr_t>(d->out) & 15))
    return CTU_E_BADARG;
  if (d->out_mode != CTU_OUT_F32_CF) {
    if ((d->ldc % 8) != 0 || (d->out_col0 % 8) != 0 || (d->n_real % 8) != 0) return CTU_E_BADARG;
  }
  if (d->res_mode != CTU_RES_NONE && (!d->residual || (d->ldr % 8) != 0)) return CTU_E_BADARG;
  if (d->convt_cout > 0) {
    if (d->convt_cout % d->block_n != 0 || d->n_real != d->convt_cout * d->u1 * d->u2 * d->u3) return CTU_E_BADARG;
    if (d->stats != nullptr) return CTU_E_UNSUPPORTED;
  }
  if (d->stats != nullptr && d->out_mode == CTU_OUT_F32_CF) return CTU_E_UNSUPPORTED;
  if (!tma_encoder()) return CTU_E_DRIVER;

  // A: 5-D channels-last tensor map (C, d1, d2, d3, d4)
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[5] = {(cuuint64_t)d->a_c, (cuuint64_t)d->d1, (cuuint64_t)d->d2, (cuuint64_t)d->d3, (cuuint64_t)d->d4};
    cuuint64_t strides[4];
    strides[0] = (cuuint64_t)d->lda * 2;
    strides[1] = strides[0] * d->d1;
    strides[2] = strides[1] * d->d2;
    strides[3] = strides[2] * d->d3;
    cuuint32_t box[5] = {(cuuint32_t)BLOCK_K, (cuuint32_t)d->b1, (cuuint32_t)d->b2, (cuuint32_t)d->b3, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = tma_encoder()(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(d->a), dims, strides, box,
                               es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return CTU_E_DRIVER;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)d->k_total, (cuuint64_t)d->n_pad};
    cuuint64_t strides[1] = {(cuuint64_t)d->k_total * 2};
    cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)d->block_n};
    cuuint32_t es[2] = {1, 1};
    CUresult r = tma_encoder()(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(d->w), dims, strides, box,
                               es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return CTU_E_DRIVER;
  }

  GemmParams p;
  p.b1 = d->b1; p.b2 = d->b2; p.b3 = d->b3;
  p.T1 = (d->d1 + d->b1 - 1) / d->b1;
  p.T2 = (d->d2 + d->b2 - 1) / d->b2;
  p.T3 = (d->d3 + d->b3 - 1) / d->b3;
  p.d1 = d->d1; p.d2 = d->d2; p.d3 = d->d3;
  p.n_tiles = d->n_pad / d->block_n;
  p.cblocks = (d->a_c + BLOCK_K - 1) / BLOCK_K;
  p.num_kb = taps * p.cblocks;
  p.a_c = d->a_c;
  p.k1 = d->k1; p.k2 = d->k2;
  p.pad = (d->k1 == 3) ? 1 : 0;
  p.out = d->out; p.bias = d->bias; p.residual = d->residual; p.stats = d->stats;
  p.n_real = d->n_real; p.out_mode = d->out_mode; p.ldc = d->ldc; p.act = d->act;
  p.res_mode = d->res_mode; p.ldr = d->ldr;
  p.convt_cout = d->convt_cout;
  p.u1 = d->convt_cout > 0 ? d->u1 : 1;
  p.u2 = d->convt_cout > 0 ? d->u2 : 1;
  p.u3 = d->convt_cout > 0 ? d->u3 : 1;
  p.stats_ld = d->stats_ld; p.out_col0 = d->out_col0;
  const long long grid_ll = (long long)p.T1 * p.T2 * p.T3 * d->d4 * p.n_tiles;
  if (grid_ll <= 0 || grid_ll > 0x7fffffffLL) return CTU_E_BADARG;
  const int grid = (int)grid_ll;

  switch (d->block_n) {
    case 16: return launch_gemm<16, 4>(tmA, tmB, p, grid, stream);
    case 32: return launch_gemm<32, 4>(tmA, tmB, p, grid, stream);
    case 64: return launch_gemm<64, 4>(tmA, tmB, p, grid, stream);
    case 128: return launch_gemm<128, 3>(tmA, tmB, p, grid, stream);
    case 256: return launch_gemm<256, 4>(tmA, tmB, p, grid, stream);
    default: return CTU_E_UNSUPPORTED;
  }
}
