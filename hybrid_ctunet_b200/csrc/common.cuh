// Device-side PTX wrappers shared by the sm_100a kernels (mbarrier, TMA, tcgen05/TMEM).
// Everything here is inline PTX for sm_100a; there is no fallback path for other architectures.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace ctu {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- programmatic dependent launch
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a protocol bug traps (visible as a CUDA error) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) {
      printf("ctunet_b200: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1,
                                            int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, "
      "%7}], [%2];" ::"r"(dst),
      "l"(m), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; bf16 inputs, fp32 accumulate, single-CTA group.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// The four K16 steps of one 64-wide K block (operands advance 32 bytes = +2 in the descriptors' address field): one asm
// block, one predicate.  `accumulate` = 0 zero-initialises the accumulator with the FIRST step only.
__device__ __forceinline__ void umma_bf16_k4(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      ".reg .b64 a, b;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "setp.eq.b32 q, 0, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "add.u64 a, %1, 2;\n"
      "add.u64 b, %2, 2;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, q;\n"
      "add.u64 a, %1, 4;\n"
      "add.u64 b, %2, 4;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, q;\n"
      "add.u64 a, %1, 6;\n"
      "add.u64 b, %2, 6;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, q;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Eight instructions whose operand descriptors advance by `step_a` / `step_b` (units of the descriptors' 16-byte address
// field) — the K loop of the weight-gradient kernels (16 voxels per instruction) as one asm block with one predicate.
__device__ __forceinline__ void umma_bf16_k8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint64_t step_a,
                                             uint64_t step_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      ".reg .b64 a, b;\n"
      "setp.ne.b32 p, %6, 0;\n"
      "setp.eq.b32 q, 0, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %5, p;\n"
      "add.u64 a, %1, %3;\n"
      "add.u64 b, %2, %4;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, q;\n"
      "add.u64 a, a, %3;\n"
      "add.u64 b, b, %4;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, q;\n"
      "add.u64 a, a, %3;\n"
      "add.u64 b, b, %4;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, q;\n"
      "add.u64 a, a, %3;\n"
      "add.u64 b, b, %4;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, q;\n"
      "add.u64 a, a, %3;\n"
      "add.u64 b, b, %4;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, q;\n"
      "add.u64 a, a, %3;\n"
      "add.u64 b, b, %4;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, q;\n"
      "add.u64 a, a, %3;\n"
      "add.u64 b, b, %4;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, q;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "l"(step_a), "l"(step_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32 lanes x 32 columns of fp32: thread i of the warp receives lane (base_lane + i), columns [col, col+32).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand tile in shared memory, 128-byte rows, SWIZZLE_128B (what a TMA box with a 64 x bf16 inner
// extent and CU_TENSOR_MAP_SWIZZLE_128B produces): 8-row groups are 1024 B apart (SBO); LBO is unused.
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);   // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                    // leading byte offset (ignored for swizzled K-major), bits [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                    // descriptor version (Blackwell), bits [46,48)
  d |= (uint64_t)2 << 61;                    // layout type SWIZZLE_128B, bits [61,64)
  return d;
}

// tcgen05 instruction descriptor: bf16 x bf16 -> fp32, both operands K-major, M x N tile.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n) {
  return (1u << 4)                   // c_format = F32
         | (1u << 7)                 // a_format = BF16
         | (1u << 10)                // b_format = BF16
         | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

// ---------------------------------------------------------------- packed fp32 (FFMA2 / FMUL2 on sm_100)
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Forward GELU through the EXPONENT:  Phi(-|x|) = 2^S(|x|),  S = log2 Phi(-a) as a degree-6 interpolant on a in [0, 5.2]
// (log Phi is smooth and nearly quadratic, so six Horner steps give |gelu - exact| <= 2.6e-6 over [-12, 12] in fp32 — the
// accuracy of the degree-10 erfcx form below at 60 % of its FMA-pipe cycles), then
//   gelu(x) = max(x, 0) - |x| * Phi(-|x|)
// with no compare / select.  Beyond |x| = 5.2 the clamped Phi(-5.2) = 1e-7 leaves |x| * 1e-7.  Packed: two values per chain.
__device__ __forceinline__ uint64_t gelu_log2phi(uint64_t a) {
  uint64_t s = f2_fma(f2_pack(2.868008041e-05f, 2.868008041e-05f), a, f2_pack(-6.992247615e-04f, -6.992247615e-04f));
  s = f2_fma(s, a, f2_pack(7.705668348e-03f, 7.705668348e-03f));
  s = f2_fma(s, a, f2_pack(-5.254218971e-02f, -5.254218971e-02f));
  s = f2_fma(s, a, f2_pack(-4.596425026e-01f, -4.596425026e-01f));
  s = f2_fma(s, a, f2_pack(-1.150893286e+00f, -1.150893286e+00f));
  s = f2_fma(s, a, f2_pack(-1.000011945e+00f, -1.000011945e+00f));
  return s;
}
__device__ __forceinline__ void gelu_fwd_pair(float xa, float xb, float& ya, float& yb) {
  constexpr float XMAX = 5.2f;
  float sa, sb;
  f2_unpack(gelu_log2phi(f2_pack(fminf(fabsf(xa), XMAX), fminf(fabsf(xb), XMAX))), sa, sb);
  ya = fmaf(-fabsf(xa), ex2_approx(sa), fmaxf(xa, 0.f));
  yb = fmaf(-fabsf(xb), ex2_approx(sb), fmaxf(xb, 0.f));
}

// Exact-form GELU 0.5 x (1 + erf(x / sqrt 2)) = x * Phi(x) for TWO values at once, without a reciprocal:
//   Phi(-|x|) = 0.5 erfc(|x| / sqrt 2) = exp(-x^2 / 2) * P(w),   w = |x| * sqrt(log2(e) / 2)   (so exp(-x^2/2) = 2^(-w^2)),
// P = degree-10 Chebyshev interpolant of 0.5 erfcx on w in [0, 4.9] (|x| <= 5.77; beyond, 2^(-w^2) < 1e-7 and P is clamped).
// |gelu - exact| <= 3.1e-6, |Phi - exact| <= 6.8e-6 over [-8, 8] (fp32 Horner; far below the bf16 rounding of the stored
// result).  One MUFU (ex2) and ~6 issue slots of packed fp32 arithmetic per element; `grad` also returns
// d/dx = Phi(x) + x phi(x) with the same exponential (phi(x) = 0.3989423 * 2^(-w^2)).
template <bool GRAD>
__device__ __forceinline__ void gelu_pair(float xa, float xb, float& ya, float& yb) {
  if (!GRAD) {   // the forward takes the cheaper exponent form above
    gelu_fwd_pair(xa, xb, ya, yb);
    return;
  }
  constexpr float K = 0.8493218002880191f, WMAX = 4.9f;
  const float wa = fabsf(xa) * K, wb = fabsf(xb) * K;
  const uint64_t w = f2_pack(fminf(wa, WMAX), fminf(wb, WMAX));
  const float ea = ex2_approx(-wa * wa), eb = ex2_approx(-wb * wb);
  uint64_t p = f2_pack(8.061951625e-07f, 8.061951625e-07f);
  p = f2_fma(p, w, f2_pack(-2.335833055e-05f, -2.335833055e-05f));
  p = f2_fma(p, w, f2_pack(2.997489816e-04f, 2.997489816e-04f));
  p = f2_fma(p, w, f2_pack(-2.258083635e-03f, -2.258083635e-03f));
  p = f2_fma(p, w, f2_pack(1.120215335e-02f, 1.120215335e-02f));
  p = f2_fma(p, w, f2_pack(-3.916574339e-02f, -3.916574339e-02f));
  p = f2_fma(p, w, f2_pack(1.018373904e-01f, 1.018373904e-01f));
  p = f2_fma(p, w, f2_pack(-2.071734658e-01f, -2.071734658e-01f));
  p = f2_fma(p, w, f2_pack(3.437036826e-01f, 3.437036826e-01f));
  p = f2_fma(p, w, f2_pack(-4.693814702e-01f, -4.693814702e-01f));
  p = f2_fma(p, w, f2_pack(4.999932302e-01f, 4.999932302e-01f));
  const uint64_t e = f2_pack(ea, eb);
  float ha, hb;
  f2_unpack(f2_mul(p, e), ha, hb);                 // Phi(-|x|)
  const float pa = xa >= 0.f ? 1.f - ha : ha, pb = xb >= 0.f ? 1.f - hb : hb;
  if (GRAD) {
    f2_unpack(f2_fma(f2_mul(f2_pack(xa, xb), f2_pack(0.3989422804014327f, 0.3989422804014327f)), e, f2_pack(pa, pb)), ya, yb);
  } else {
    ya = xa * pa;
    yb = xb * pb;
  }
}

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

}  // namespace ctu
