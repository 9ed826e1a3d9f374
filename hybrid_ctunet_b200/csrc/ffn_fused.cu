// Fused token FeedForward of the 128-channel window-attention stage for inference (hybrid_CTUNet.py:513-526 inside
// Residual, :434-440):   out = x + W2 · GELU(W1 · LN(x) + b1) + b2   with the [rows, 4C] hidden activation kept ON CHIP.
//
// The unfused path writes the hidden tensor (rows x 512 bf16 = 906 MB at 4 windows) and reads it back; here one
// persistent CTA per SM walks 128-row tiles:
//   warp 0 (1 lane) : TMA producer — the tile's A rows (LN(x), 128 x C bf16 as C/64 SWIZZLE_128B K blocks, double
//                     buffered) and a ring of 16 KB weight stages (128 rows x 64 K) in the order the MMA warp consumes
//                     them: W1 chunk g, then W2 chunk g-1, ...  (weights are 256 KB: L2-resident, re-streamed per tile).
//   warp 1 (1 lane) : MMA issuer — per 128-wide hidden chunk j:  acc1[j&1] = A · W1_jᵀ  (TMEM, double buffered), and one
//                     chunk later  acc2 += H_j · W2_jᵀ  where H_j is the GELU'd chunk the epilogue warps wrote to shared
//                     memory as a K-major SWIZZLE_128B operand.  GEMM1 of chunk j+1 is issued BEFORE GEMM2 of chunk j, so
//                     the tensor pipe always has work while the epilogue warps run the GELU of chunk j.
//   warps 2..       : epilogue (8 or 16 warps) — E1: tcgen05.ld acc1, + b1, exact-erf GELU (packed fp32 polynomial), bf16,
//                     st.shared into H[j&1];  E2 (once per tile, issued after E1 of the NEXT tile's first chunk so that it
//                     never waits for the last GEMM2): tcgen05.ld acc2, + b2, + x (residual rows from global), bf16 into a
//                     swizzled staging tile, one TMA store per 64-channel slab.
// TMEM: acc1 2 x 128 columns + acc2 2 x C columns = 512 (C = 128).  Shared memory: A 32 KB, W 6 x 16 KB, H 2 x 32 KB,
// staging 32 KB = 224 KB (the weight ring needs >= 96 KB in flight to cover the L2 latency at 64 GB/s per SM).  Bound: the GELU issue rate of the epilogue warps (~8 issue slots per hidden element) and the
// tensor pipe (33.5 MFLOP per tile) are of the same size; HBM traffic is rows x C x 2 x 3 bytes (A, residual, out).
#include "common.cuh"
#include "../../include/ctunet_b200.h"
#include "host_util.h"
#include <stdlib.h>

namespace ctu {
namespace ffn {

constexpr int BM = 128;                    // rows per tile
constexpr int HC = 128;                    // hidden columns per chunk
constexpr int KB_BYTES = 128 * 64 * 2;     // one operand K block: 128 rows x 64 bf16, SWIZZLE_128B
constexpr int W_STAGES = 6;

struct Params {
  const float* b1;
  const float* b2;
  const __nv_bfloat16* residual;
  long long ldr;
  long long M;
  int tiles;
  int nc;        // hidden / HC
  int dbg;       // timing experiments only (CTU_FFN_DBG): 1 = no GELU math, 4 = no residual loads, 8 = no chunk rotation
};

template <int C>
struct Smem {
  static constexpr int A_BYTES = (C / 64) * KB_BYTES;
  static constexpr int H_BYTES = (HC / 64) * KB_BYTES;
  static constexpr int CST_BYTES = (C / 64) * KB_BYTES;
  static constexpr int OFF_A = 0;
  static constexpr int OFF_W = OFF_A + A_BYTES;            // the A rows are single-buffered: the ring gets the space
  static constexpr int OFF_H = OFF_W + W_STAGES * KB_BYTES;
  static constexpr int OFF_CST = OFF_H + 2 * H_BYTES;
  static constexpr int OFF_BAR = OFF_CST + CST_BYTES;
  static constexpr int NUM_BARS = 2 + 2 + 2 * W_STAGES + 2 + 2 + 2 + 2 + 2 + 2;
  static constexpr int OFF_BIAS = OFF_BAR + 256;                       // barriers + TMEM slot fit in 256 bytes
  static constexpr int B1_SMEM = 512;                                  // hidden biases staged in shared memory (floats)
  static constexpr int TOTAL = OFF_BIAS + (B1_SMEM + C) * 4;           // no alignment slack: the base is 1024-aligned
};

template <int C, int EPI_WARPS>
__global__ void __launch_bounds__(64 + 32 * EPI_WARPS, 1) ffn_fused_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                           const __grid_constant__ CUtensorMap tmW1,
                                                                           const __grid_constant__ CUtensorMap tmW2,
                                                                           const __grid_constant__ CUtensorMap tmOut,
                                                                           const Params p) {
  using S = Smem<C>;
  static_assert(C == 128, "TMEM / shared-memory plan is for C = 128");
  static_assert(EPI_WARPS == 8 || EPI_WARPS == 16, "epilogue warps");
  constexpr int EPI_THREADS = 32 * EPI_WARPS;
  constexpr int COL_SPLIT = EPI_WARPS / 4;          // warps sharing one TMEM lane quarter
  constexpr int E1_COLS = HC / COL_SPLIT;           // hidden columns per warp in E1 (64 or 32)
  constexpr int E2_COLS = C / COL_SPLIT;            // output columns per warp in E2
  constexpr uint32_t IDESC = umma_idesc_bf16(BM, 128);
  constexpr uint32_t ACC2_COL = 2 * HC;             // TMEM column of acc2[0]

  static_assert(S::NUM_BARS * 8 + 16 <= 256, "barrier block");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) __trap();   // SWIZZLE_128B operand tiles need a 1024-byte aligned base
  uint8_t* smem_a = smem + S::OFF_A;
  uint8_t* smem_w = smem + S::OFF_W;
  uint8_t* smem_h = smem + S::OFF_H;
  uint8_t* smem_c = smem + S::OFF_CST;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint64_t* a_full = bars;                  // [2]  TMA -> MMA
  uint64_t* a_empty = bars + 2;             // [2]  MMA (commit) -> TMA
  uint64_t* w_full = bars + 4;              // [W_STAGES]
  uint64_t* w_empty = w_full + W_STAGES;    // [W_STAGES]
  uint64_t* acc1_full = w_empty + W_STAGES; // [2]  MMA -> epilogue
  uint64_t* acc1_empty = acc1_full + 2;     // [2]  epilogue -> MMA
  uint64_t* h_full = acc1_empty + 2;        // [2]  epilogue -> MMA
  uint64_t* h_empty = h_full + 2;           // [2]  MMA (commit) -> epilogue
  uint64_t* acc2_full = h_empty + 2;        // [2]
  uint64_t* acc2_empty = acc2_full + 2;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc2_empty + 2);
  float* bias1_s = reinterpret_cast<float*>(smem + S::OFF_BIAS);   // [B1_SMEM]
  float* bias2_s = bias1_s + S::B1_SMEM;                          // [C]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmOut);
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&a_full[s]), 1);
      mbar_init(smem_u32(&a_empty[s]), 1);
      mbar_init(smem_u32(&acc1_full[s]), 1);
      mbar_init(smem_u32(&acc1_empty[s]), EPI_WARPS);
      mbar_init(smem_u32(&h_full[s]), EPI_WARPS);
      mbar_init(smem_u32(&h_empty[s]), 1);
      mbar_init(smem_u32(&acc2_full[s]), 1);
      mbar_init(smem_u32(&acc2_empty[s]), EPI_WARPS);
    }
    for (int s = 0; s < W_STAGES; ++s) {
      mbar_init(smem_u32(&w_full[s]), 1);
      mbar_init(smem_u32(&w_empty[s]), 1);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  const int NC = p.nc;
  for (int i = threadIdx.x; i < NC * HC; i += blockDim.x) bias1_s[i] = __ldg(p.b1 + i);   // host: hidden <= B1_SMEM
  for (int i = threadIdx.x; i < C; i += blockDim.x) bias2_s[i] = __ldg(p.b2 + i);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // The hidden chunks may be summed in any order: CTA b starts at chunk b % NC, so that the 148 CTAs do not all stream the
  // same 64 KB of weights out of the same L2 lines at the same moment.
  const int rot = (p.dbg & 8) ? 0 : (int)(blockIdx.x % (unsigned)NC);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t wit = 0;
      auto load_w = [&](const CUtensorMap* tm, int k0, int n0) {
        const int s = wit % W_STAGES;
        const uint32_t ph = (wit / W_STAGES) & 1;
        mbar_wait(smem_u32(&w_empty[s]), ph ^ 1);
        const uint32_t full = smem_u32(&w_full[s]);
        mbar_expect_tx(full, KB_BYTES);
        tma_load_2d(smem_u32(smem_w + s * KB_BYTES), tm, full, k0, n0);
        ++wit;
      };
      auto load_a = [&](int lt, int tile) {   // single buffer: waits until GEMM1 of the previous tile's last chunk is complete
        mbar_wait(smem_u32(&a_empty[0]), (lt & 1) ^ 1);
        const uint32_t full = smem_u32(&a_full[0]);
        mbar_expect_tx(full, S::A_BYTES);
#pragma unroll
        for (int kb = 0; kb < C / 64; ++kb)
          tma_load_2d(smem_u32(smem_a + kb * KB_BYTES), &tmA, full, kb * 64, tile * BM);
      };
      int lt = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++lt) {
        if (lt == 0) load_a(0, tile);
        for (int j = 0; j < NC; ++j) {
#pragma unroll
          for (int kb = 0; kb < C / 64; ++kb) load_w(&tmW1, kb * 64, ((j + rot) % NC) * HC);
          if (lt > 0 || j > 0) {
            const int jp = (j + NC - 1 + rot) % NC;
#pragma unroll
            for (int kk = 0; kk < HC / 64; ++kk) load_w(&tmW2, jp * HC + kk * 64, 0);
          }
          // next tile's rows, behind the weight stages of this tile's last GEMM1 (the buffer frees when that GEMM1 completes)
          if (j == NC - 1 && tile + (int)gridDim.x < p.tiles) load_a(lt + 1, tile + gridDim.x);
        }
      }
      if (lt > 0) {
#pragma unroll
        for (int kk = 0; kk < HC / 64; ++kk) load_w(&tmW2, ((NC - 1 + rot) % NC) * HC + kk * 64, 0);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      uint32_t wit = 0;
      int g = 0;   // chunks issued by this CTA (GEMM1 count)
      auto gemm2 = [&](int gp) {
        const int ltp = gp / NC, jp = gp - ltp * NC;
        const int hs = gp & 1, u = ltp & 1;
        if (jp == 0) mbar_wait(smem_u32(&acc2_empty[u]), ((ltp >> 1) & 1) ^ 1);   // E2 of two tiles ago has drained acc2[u]
        mbar_wait(smem_u32(&h_full[hs]), (gp >> 1) & 1);
        tc_fence_after();
        const uint32_t acc2 = tmem_base + ACC2_COL + (uint32_t)(u * C);
#pragma unroll
        for (int kk = 0; kk < HC / 64; ++kk) {
          const int s = wit % W_STAGES;
          mbar_wait(smem_u32(&w_full[s]), (wit / W_STAGES) & 1);
          tc_fence_after();
          const uint64_t da = umma_desc_k_sw128(smem_u32(smem_h + hs * S::H_BYTES + kk * KB_BYTES));
          const uint64_t db = umma_desc_k_sw128(smem_u32(smem_w + s * KB_BYTES));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(acc2, da + 2 * k, db + 2 * k, IDESC, (jp | kk | k) != 0 ? 1u : 0u);
          umma_commit(smem_u32(&w_empty[s]));
          ++wit;
        }
        umma_commit(smem_u32(&h_empty[hs]));
        if (jp == NC - 1) umma_commit(smem_u32(&acc2_full[u]));
      };
      int lt = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++lt) {
        mbar_wait(smem_u32(&a_full[0]), lt & 1);
        tc_fence_after();
        for (int j = 0; j < NC; ++j, ++g) {
          const int as = g & 1;
          mbar_wait(smem_u32(&acc1_empty[as]), ((g >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t acc1 = tmem_base + (uint32_t)(as * HC);
#pragma unroll
          for (int kb = 0; kb < C / 64; ++kb) {
            const int s = wit % W_STAGES;
            mbar_wait(smem_u32(&w_full[s]), (wit / W_STAGES) & 1);
            tc_fence_after();
            const uint64_t da = umma_desc_k_sw128(smem_u32(smem_a + kb * KB_BYTES));
            const uint64_t db = umma_desc_k_sw128(smem_u32(smem_w + s * KB_BYTES));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(acc1, da + 2 * k, db + 2 * k, IDESC, (kb | k) != 0 ? 1u : 0u);
            umma_commit(smem_u32(&w_empty[s]));
            ++wit;
          }
          umma_commit(smem_u32(&acc1_full[as]));
          if (j == NC - 1) umma_commit(smem_u32(&a_empty[0]));   // the tile's rows are consumed
          if (g >= 1) gemm2(g - 1);
        }
      }
      if (g >= 1) gemm2(g - 1);
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int q = warp & 3;                    // TMEM lane quarter this warp may read
    const int r = q * 32 + lane;               // row of the tile
    const int e = threadIdx.x - 64;
    const int part = (warp - 2) >> 2;          // which share of the columns (warps w, w+4, ... share a lane quarter)

    // residual rows of the tile whose E2 comes next: loaded one GELU chunk ahead (their HBM latency hides behind E1)
    constexpr int RV = E2_COLS / 8;
    uint4 resv[RV];
    auto load_residual = [&](int tile) {
      const long long row = (long long)tile * BM + r;
      if (row < p.M && !(p.dbg & 4)) {
        const uint4* rp = reinterpret_cast<const uint4*>(p.residual + row * p.ldr + part * E2_COLS);
#pragma unroll
        for (int i = 0; i < RV; ++i) resv[i] = rp[i];
      } else {
#pragma unroll
        for (int i = 0; i < RV; ++i) resv[i] = make_uint4(0u, 0u, 0u, 0u);
      }
    };
    auto e2 = [&](int lt, int tile) {
      const int u = lt & 1;
      mbar_wait(smem_u32(&acc2_full[u]), (lt >> 1) & 1);
      tc_fence_after();
      if (e == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // previous store has read the staging tile
      named_bar_sync(1, EPI_THREADS);
#pragma unroll
      for (int b = 0; b < E2_COLS / 32; ++b) {
        const int c0 = part * E2_COLS + b * 32;
        uint32_t raw[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + ACC2_COL + (uint32_t)(u * C + c0), raw);
        tmem_ld_wait();
        const float4* bp = reinterpret_cast<const float4*>(bias2_s + c0);
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint4 rvi = resv[4 * b + i];
          const uint32_t uu[4] = {rvi.x, rvi.y, rvi.z, rvi.w};
          const float4 ba = bp[2 * i], bb = bp[2 * i + 1];
          const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const float2 f = unpack_bf16x2(uu[h]);
            const int j = 8 * i + 2 * h;
            pk[4 * i + h] = pack_bf16x2(__uint_as_float(raw[j]) + bv[2 * h] + f.x, __uint_as_float(raw[j + 1]) + bv[2 * h + 1] + f.y);
          }
        }
        const uint32_t base = smem_u32(smem_c) + (uint32_t)((c0 >> 6) * KB_BYTES + r * 128);
        const int cb = (c0 & 63) >> 3;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t addr = base + (uint32_t)(((cb + i) ^ (r & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * i]), "r"(pk[4 * i + 1]),
                       "r"(pk[4 * i + 2]), "r"(pk[4 * i + 3])
                       : "memory");
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&acc2_empty[u]));
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      named_bar_sync(2, EPI_THREADS);
      if (e == 0) {
#pragma unroll
        for (int sl = 0; sl < C / 64; ++sl) {
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tmOut),
                       "r"(smem_u32(smem_c + sl * KB_BYTES)), "r"(sl * 64), "r"(tile * BM)
                       : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    };

    int g = 0, lt = 0, prev_tile = -1;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++lt) {
      for (int j = 0; j < NC; ++j, ++g) {
        const int as = g & 1, hs = g & 1;
        const uint32_t ph = (g >> 1) & 1;
        if (j == 0 && prev_tile >= 0) load_residual(prev_tile);
        mbar_wait(smem_u32(&h_empty[hs]), ph ^ 1);     // GEMM2 of chunk g-2 has read H[hs]
        mbar_wait(smem_u32(&acc1_full[as]), ph);
        tc_fence_after();
        constexpr int NB = E1_COLS / 32;            // 32-column blocks per warp: all TMEM loads in flight before the math
        uint32_t raw[NB][32];
#pragma unroll
        for (int b = 0; b < NB; ++b)
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * HC + part * E1_COLS + b * 32), raw[b]);
        tmem_ld_wait();
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          const int c0 = part * E1_COLS + b * 32;
          const float4* bp = reinterpret_cast<const float4*>(bias1_s + ((j + rot) % NC) * HC + c0);
          uint32_t pk[16];
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            float v[16];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 bv = bp[4 * hh + i];
              v[4 * i] = __uint_as_float(raw[b][16 * hh + 4 * i]) + bv.x;
              v[4 * i + 1] = __uint_as_float(raw[b][16 * hh + 4 * i + 1]) + bv.y;
              v[4 * i + 2] = __uint_as_float(raw[b][16 * hh + 4 * i + 2]) + bv.z;
              v[4 * i + 3] = __uint_as_float(raw[b][16 * hh + 4 * i + 3]) + bv.w;
            }
            if (!(p.dbg & 1)) {
#pragma unroll
              for (int i = 0; i < 8; ++i) gelu_fwd_pair(v[2 * i], v[2 * i + 1], v[2 * i], v[2 * i + 1]);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) pk[8 * hh + i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
          }
          // K-major SWIZZLE_128B operand: K block (c0 / 64), row r at r * 128, 16-byte chunk c at (c ^ (r & 7))
          const uint32_t base = smem_u32(smem_h + hs * S::H_BYTES) + (uint32_t)((c0 >> 6) * KB_BYTES + r * 128);
          const int cb = (c0 & 63) >> 3;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t addr = base + (uint32_t)(((cb + i) ^ (r & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * i]), "r"(pk[4 * i + 1]),
                         "r"(pk[4 * i + 2]), "r"(pk[4 * i + 3])
                         : "memory");
          }
        }
        tc_fence_before();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // H visible to the tensor core's async proxy
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(smem_u32(&acc1_empty[as]));
          mbar_arrive(smem_u32(&h_full[hs]));
        }
        if (j == 0 && prev_tile >= 0) e2(lt - 1, prev_tile);   // the previous tile's output, after this tile's first GELU chunk
      }
      prev_tile = tile;
    }
    if (prev_tile >= 0) {
      load_residual(prev_tile);
      e2(lt - 1, prev_tile);
    }
    if (e == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

static int sm_count() {
  static int n = [] {
    int dev = 0, v = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v;
  }();
  return n;
}

static bool encode_2d(CUtensorMap* tm, const void* base, long long inner, long long rows, long long row_stride_elems) {
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)row_stride_elems * 2};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t es[2] = {1, 1};
  return tma_encoder()(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int C, int EPI_WARPS>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmW1, const CUtensorMap& tmW2, const CUtensorMap& tmOut,
                  const Params& p, cudaStream_t stream) {
  constexpr int smem = Smem<C>::TOTAL;
  static_assert(smem <= 232448, "shared memory budget (227 KB opt-in maximum)");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(ffn_fused_kernel<C, EPI_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const int cap = persistent_sms(sm_count());
  const int grid = p.tiles < cap ? p.tiles : cap;
  ffn_fused_kernel<C, EPI_WARPS><<<grid, 64 + 32 * EPI_WARPS, smem, stream>>>(tmA, tmW1, tmW2, tmOut, p);
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace ffn
}  // namespace ctu

extern "C" int ctu_ffn_fused(const void* a, long long lda, const void* w1, const float* b1, const void* w2, const float* b2,
                             const void* residual, long long ldr, void* out, long long ldc, long long M, int C, int hidden,
                             void* stream_) {
  using namespace ctu;
  using namespace ctu::ffn;
  if (!a || !w1 || !b1 || !w2 || !b2 || !residual || !out || M <= 0) return CTU_E_BADARG;
  if (C != 128 || hidden <= 0 || hidden % HC != 0 || hidden > Smem<128>::B1_SMEM) return CTU_E_UNSUPPORTED;
  if ((lda % 8) || (ldr % 8) || (ldc % 8) || lda < C || ldr < C || ldc < C) return CTU_E_BADARG;
  if ((reinterpret_cast<uintptr_t>(a) & 15) || (reinterpret_cast<uintptr_t>(w1) & 15) || (reinterpret_cast<uintptr_t>(w2) & 15) ||
      (reinterpret_cast<uintptr_t>(residual) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) ||
      (reinterpret_cast<uintptr_t>(b1) & 15) || (reinterpret_cast<uintptr_t>(b2) & 15))
    return CTU_E_BADARG;
  const long long tiles = (M + BM - 1) / BM;
  if (tiles > 0x7fffffffLL / BM) return CTU_E_BADARG;
  if (!tma_encoder()) return CTU_E_DRIVER;
  CUtensorMap tmA, tmW1, tmW2, tmOut;
  if (!encode_2d(&tmA, a, C, M, lda) || !encode_2d(&tmW1, w1, C, hidden, C) || !encode_2d(&tmW2, w2, hidden, C, hidden) ||
      !encode_2d(&tmOut, out, C, M, ldc))
    return CTU_E_DRIVER;
  Params p;
  p.b1 = b1; p.b2 = b2;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  p.ldr = ldr; p.M = M; p.tiles = (int)tiles; p.nc = hidden / HC;
  static const int dbg = [] { const char* e = getenv("CTU_FFN_DBG"); return e ? atoi(e) : 0; }();
  p.dbg = dbg;
  static const int epi16 = [] { const char* e = getenv("CTU_FFN_EPI16"); return e ? atoi(e) : 1; }();   // measured: 0.39 vs 0.42 ms
  if (epi16) return launch<128, 16>(tmA, tmW1, tmW2, tmOut, p, reinterpret_cast<cudaStream_t>(stream_));
  return launch<128, 8>(tmA, tmW1, tmW2, tmOut, p, reinterpret_cast<cudaStream_t>(stream_));
}
