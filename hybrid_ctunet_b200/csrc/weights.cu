// Multi-tensor weight packing / gradient unpacking (sm_100a).
//
// The tensor-core kernels consume bf16 weights in their own layouts ([N_pad][taps*C_in] K-major for the forward
// contraction, transposed / tap-flipped for the input-gradient contraction) and produce weight gradients in the
// transposed-packed fp32 layout [(tap, c_in)][c_out].  A training step therefore re-packs ~400 parameters (twice) and
// unpacks ~400 gradients; done with torch ops that is ~2,600 tiny kernels per step.  Here each direction is ONE launch
// over a device-resident table of items: a block finds its item by binary search over the items' first work unit
// (256 thread-tasks per unit).  A thread-task moves the innermost run of its index map (8 consecutive K values, the 27
// taps of one (c_out, c_in) pair, the k^3 sub-voxels of one (c_in, c_out) pair ...) so that every 32-byte sector it
// touches on the strided side is fully used and the other side is coalesced across the warp.  The layouts that carry
// ~95 % of the bytes (Linear transposes, 3x3x3 convolutions) instead move one TILE per work unit through shared memory
// so that both the global reads and the global writes are long contiguous runs (16-byte vectors where aligned).
#include "common.cuh"
#include "../../include/ctunet_b200.h"
#include "host_util.h"

namespace ctu {

__device__ __forceinline__ int find_item(const ctu_pack_item* items, int n, long long unit) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (items[mid].unit0 <= unit) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// Number of thread-tasks of a pack item (host and device agree on this).
__host__ __device__ inline long long pack_tasks(int kind, int rows, int cols, int a, int b, int c) {
  switch (kind) {
    case CTU_PACK_LIN: return (long long)rows * (cols / 8);
    // tile kinds: one 256-thread unit per tile
    case CTU_PACK_LIN_T: return (long long)((rows + 63) / 64) * ((cols + 63) / 64) * 256;
    case CTU_PACK_CONV3: return (long long)rows * ((cols / 27 + 63) / 64) * 256;
    case CTU_PACK_CONV3_T: return (long long)((rows + 3) / 4) * ((cols / 27 + 63) / 64) * 256;
    case CTU_PACK_CONVT: return (long long)a * b;
    case CTU_PACK_CONVT_T: return (long long)a * b;
    default: return (long long)rows * cols;  // PS, PS_T, CIN1: one element per task
  }
}

__host__ __device__ inline long long unpack_tasks(int kind, int rows /*param elements*/, int a, int b, int c) {
  switch (kind) {
    case CTU_PACK_LIN: return (long long)((a + 63) / 64) * ((b + 63) / 64) * 256;   // tile kinds
    case CTU_PACK_CONV3: return (long long)((b + 3) / 4) * ((a + 63) / 64) * 256;
    case CTU_PACK_CONVT: return (long long)a * b;
    default: return rows;  // PS, PS_BIAS, CIN1, VEC: one element per task
  }
}


// ---------------------------------------------------------------------------------------------- tile movers
constexpr int TP = 65;    // pitch of a 64 x 64 fp32 tile
constexpr int CP = 109;   // pitch of a [64][4 * 27] fp32 tile
constexpr int TILE_SM_FLOATS = 64 * CP;

__device__ __forceinline__ uint4 pack8_bf16(const float (&v)[8]) {
  return make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}

// dst (bf16) [k][n] = src (fp32) [n][k]: 64 x 64 tile, rows of src read as 256-byte runs, rows of dst written as
// 128-byte runs
__device__ __forceinline__ void pack_lin_t_tile(const ctu_pack_item& it, long long tile, float* sm) {
  const float* src = reinterpret_cast<const float*>(it.src);
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(it.dst);
  const int tiles_n = (it.cols + 63) / 64;
  const int k0 = (int)(tile / tiles_n) * 64, n0 = (int)(tile % tiles_n) * 64;
  const int tid = threadIdx.x;
  const bool vec = (it.b % 4) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
  const bool vec_st = (it.cols % 8) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int nl = (tid >> 4) + 16 * p, kk = (tid & 15) * 4;
    const int n = n0 + nl, k = k0 + kk;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < it.a) {
      const float* sp = src + (long long)n * it.b + k;
      if (vec && k + 4 <= it.b) {
        v = *reinterpret_cast<const float4*>(sp);
      } else {
        if (k < it.b) v.x = sp[0];
        if (k + 1 < it.b) v.y = sp[1];
        if (k + 2 < it.b) v.z = sp[2];
        if (k + 3 < it.b) v.w = sp[3];
      }
    }
    float* d = sm + nl * TP + kk;
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
  __syncthreads();
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int kl = (tid >> 3) + 32 * p, n8 = (tid & 7) * 8;
    const int k = k0 + kl, n = n0 + n8;
    if (k < it.rows && n < it.cols) {
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = sm[(n8 + e) * TP + kl];
      __nv_bfloat16* dp = dst + (long long)k * it.cols + n;
      if (vec_st && n + 8 <= it.cols) {
        *reinterpret_cast<uint4*>(dp) = pack8_bf16(v);
      } else {
        for (int e = 0; e < 8 && n + e < it.cols; ++e) dp[e] = __float2bfloat16(v[e]);
      }
    }
  }
  __syncthreads();
}

// dst (bf16) [co][tap * cip + ci] = src (fp32) [co][ci][27]: one output row x 64 input channels per tile; the source
// is ONE contiguous run of 64 * 27 floats, the destination 27 runs of 128 bytes
__device__ __forceinline__ void pack_conv3_tile(const ctu_pack_item& it, long long tile, float* sm) {
  const float* src = reinterpret_cast<const float*>(it.src);
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(it.dst);
  const int cip = it.cols / 27;
  const int chunks = (cip + 63) / 64;
  const int r = (int)(tile / chunks), ci0 = (int)(tile % chunks) * 64;
  const int tid = threadIdx.x;
  int n_ci = it.b - ci0;
  n_ci = n_ci < 0 ? 0 : (n_ci > 64 ? 64 : n_ci);
  const int cnt = r < it.a ? n_ci * 27 : 0;
  const float* sp = src + ((long long)r * it.b + ci0) * 27;
  for (int i = tid; i < 64 * 27; i += 256) sm[i] = i < cnt ? sp[i] : 0.f;
  __syncthreads();
  for (int t = tid; t < 27 * 8; t += 256) {
    const int tap = t >> 3, c8 = (t & 7) * 8;
    if (ci0 + c8 < cip) {
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = sm[(c8 + e) * 27 + tap];
      *reinterpret_cast<uint4*>(dst + (long long)r * it.cols + (long long)tap * cip + ci0 + c8) = pack8_bf16(v);
    }
  }
  __syncthreads();
}

// dst (bf16) [ci][tap * cop + co] = src (fp32) [co][ci][26 - tap]: 4 input channels x 64 output channels per tile;
// source runs of 4 * 27 floats, destination runs of 128 bytes
__device__ __forceinline__ void pack_conv3_t_tile(const ctu_pack_item& it, long long tile, float* sm) {
  const float* src = reinterpret_cast<const float*>(it.src);
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(it.dst);
  const int cop = it.cols / 27;
  const int chunks = (cop + 63) / 64;
  const int ci0 = (int)(tile / chunks) * 4, o0 = (int)(tile % chunks) * 64;
  const int tid = threadIdx.x;
  int n_ci = it.b - ci0;
  n_ci = n_ci < 0 ? 0 : (n_ci > 4 ? 4 : n_ci);
  const int run = n_ci * 27;
  for (int i = tid; i < 64 * 108; i += 256) {
    const int o = i / 108, j = i - o * 108;
    sm[o * CP + j] = (o0 + o < it.a && j < run) ? src[((long long)(o0 + o) * it.b + ci0) * 27 + j] : 0.f;
  }
  __syncthreads();
  for (int t = tid; t < 4 * 27 * 8; t += 256) {
    const int o8 = (t & 7) * 8, rest = t >> 3;
    const int tap = rest % 27, cl = rest / 27;
    if (ci0 + cl < it.rows && o0 + o8 < cop) {
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = sm[(o8 + e) * CP + cl * 27 + 26 - tap];
      *reinterpret_cast<uint4*>(dst + (long long)(ci0 + cl) * it.cols + (long long)tap * cop + o0 + o8) = pack8_bf16(v);
    }
  }
  __syncthreads();
}

// g (fp32) [n][k] = buf (fp32) [k][ld] : 64 x 64 tile, 256-byte runs on both sides
__device__ __forceinline__ void unpack_lin_tile(const ctu_pack_item& it, long long tile, float* sm) {
  const float* buf = reinterpret_cast<const float*>(it.src);
  float* g = reinterpret_cast<float*>(it.dst);
  const int ld = it.cols;
  const int tiles_k = (it.b + 63) / 64;
  const int n0 = (int)(tile / tiles_k) * 64, k0 = (int)(tile % tiles_k) * 64;
  const int tid = threadIdx.x;
  const bool vec_in = (ld % 4) == 0 && (reinterpret_cast<uintptr_t>(buf) & 15) == 0;
  const bool vec_out = (it.b % 4) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int kl = (tid >> 4) + 16 * p, nn = (tid & 15) * 4;
    const int k = k0 + kl, n = n0 + nn;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k < it.b) {
      const float* sp = buf + (long long)k * ld + n;
      if (vec_in && n + 4 <= ld) {
        v = *reinterpret_cast<const float4*>(sp);
      } else {
        if (n < ld) v.x = sp[0];
        if (n + 1 < ld) v.y = sp[1];
        if (n + 2 < ld) v.z = sp[2];
        if (n + 3 < ld) v.w = sp[3];
      }
    }
    float* d = sm + kl * TP + nn;
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
  __syncthreads();
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int nl = (tid >> 4) + 16 * p, kk = (tid & 15) * 4;
    const int n = n0 + nl, k = k0 + kk;
    if (n < it.a && k < it.b) {
      const float v0 = sm[kk * TP + nl], v1 = sm[(kk + 1) * TP + nl], v2 = sm[(kk + 2) * TP + nl], v3 = sm[(kk + 3) * TP + nl];
      float* dp = g + (long long)n * it.b + k;
      if (vec_out && k + 4 <= it.b) {
        *reinterpret_cast<float4*>(dp) = make_float4(v0, v1, v2, v3);
      } else {
        dp[0] = v0;
        if (k + 1 < it.b) dp[1] = v1;
        if (k + 2 < it.b) dp[2] = v2;
        if (k + 3 < it.b) dp[3] = v3;
      }
    }
  }
  __syncthreads();
}

// g (fp32) [co][ci][27] = buf (fp32) [(tap * cip + ci)][ld]: 4 input channels x 64 output channels per tile; source
// runs of 256 bytes, destination runs of 4 * 27 floats
__device__ __forceinline__ void unpack_conv3_tile(const ctu_pack_item& it, long long tile, float* sm) {
  const float* buf = reinterpret_cast<const float*>(it.src);
  float* g = reinterpret_cast<float*>(it.dst);
  const int ld = it.cols, cip = it.c;
  const int chunks = (it.a + 63) / 64;
  const int ci0 = (int)(tile / chunks) * 4, o0 = (int)(tile % chunks) * 64;
  const int tid = threadIdx.x;
  int n_ci = it.b - ci0;
  n_ci = n_ci > 4 ? 4 : n_ci;
  for (int i = tid; i < 27 * 4 * 64; i += 256) {
    const int o = i & 63, rest = i >> 6;
    const int cl = rest & 3, tap = rest >> 2;
    sm[o * CP + cl * 27 + tap] =
        (cl < n_ci && o0 + o < ld) ? buf[((long long)tap * cip + ci0 + cl) * ld + o0 + o] : 0.f;
  }
  __syncthreads();
  const int run = n_ci * 27;
  for (int i = tid; i < 64 * 108; i += 256) {
    const int o = i / 108, j = i - o * 108;
    if (o0 + o < it.a && j < run) g[((long long)(o0 + o) * it.b + ci0) * 27 + j] = sm[o * CP + j];
  }
  __syncthreads();
}

// dst (bf16) [rows][cols] <- src (fp32 parameter)
__global__ void __launch_bounds__(256) pack_weights_kernel(const ctu_pack_item* __restrict__ items, int n_items,
                                                           long long total_units) {
  __shared__ __align__(16) float tile_sm[TILE_SM_FLOATS];
  for (long long unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
    const ctu_pack_item it = items[find_item(items, n_items, unit)];
    if (it.kind == CTU_PACK_LIN_T) { pack_lin_t_tile(it, unit - it.unit0, tile_sm); continue; }
    if (it.kind == CTU_PACK_CONV3) { pack_conv3_tile(it, unit - it.unit0, tile_sm); continue; }
    if (it.kind == CTU_PACK_CONV3_T) { pack_conv3_t_tile(it, unit - it.unit0, tile_sm); continue; }
    const long long t = (unit - it.unit0) * 256 + threadIdx.x;
    if (t >= pack_tasks(it.kind, it.rows, it.cols, it.a, it.b, it.c)) continue;
    const float* src = reinterpret_cast<const float*>(it.src);
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(it.dst);
    switch (it.kind) {
      case CTU_PACK_LIN: {  // a = N, b = K: 8 consecutive columns of one row
        const int cg = it.cols / 8;
        const int r = (int)(t / cg), c0 = (int)(t % cg) * 8;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (r < it.a && c0 + j < it.b) ? src[(long long)r * it.b + c0 + j] : 0.f;
        *reinterpret_cast<uint4*>(dst + (long long)r * it.cols + c0) =
            make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
        break;
      }
      case CTU_PACK_CONVT: {  // param [ci = a][co = b][k3 = c]; dst [k3 * co][ci]: (co, ci) pair, ci fastest
        const int ci = (int)(t % it.a), o = (int)(t / it.a);
        const float* sp = src + ((long long)ci * it.b + o) * it.c;
        for (int s = 0; s < it.c; ++s) dst[((long long)s * it.b + o) * it.cols + ci] = __float2bfloat16(sp[s]);
        break;
      }
      case CTU_PACK_CONVT_T: {  // dst [ci][k3 * co]: (ci, co) pair, co fastest
        const int o = (int)(t % it.b), ci = (int)(t / it.b);
        const float* sp = src + ((long long)ci * it.b + o) * it.c;
        for (int s = 0; s < it.c; ++s) dst[(long long)ci * it.cols + (long long)s * it.b + o] = __float2bfloat16(sp[s]);
        break;
      }
      case CTU_PACK_PS: {  // param [co = a][corg = b], k3 = c; rows = k3 * co, cols = corg * k3 (block diagonal in s)
        const int r = (int)(t / it.cols), c = (int)(t % it.cols);
        const int s = r / it.a, o = r % it.a, cc = c / it.c, s2 = c % it.c;
        dst[t] = __float2bfloat16(s == s2 ? src[(long long)o * it.b + cc] : 0.f);
        break;
      }
      case CTU_PACK_PS_T: {
        const int r = (int)(t / it.cols), c = (int)(t % it.cols);
        const int s = c / it.a, o = c % it.a, cc = r / it.c, s2 = r % it.c;
        dst[t] = __float2bfloat16(s == s2 ? src[(long long)o * it.b + cc] : 0.f);
        break;
      }
      case CTU_PACK_CIN1: {  // param [co = a][taps = b]; cols = taps padded
        const int r = (int)(t / it.cols), c = (int)(t % it.cols);
        dst[t] = __float2bfloat16((r < it.a && c < it.b) ? src[(long long)r * it.b + c] : 0.f);
        break;
      }
      // ---- "paired" layouts: two z-neighbouring voxels of a 32-channel tensor share one 64-channel row (slot s = z & 1),
      // so a 1x1x1 convolution becomes a block-diagonal GEMM and a 3x3x3 convolution a 3x3x3 convolution over pairs whose
      // z pair-tap pz and slots (s_in, s_out) select the real z tap dz = 2 (pz - 1) + s_in - s_out (zero if |dz| > 1).
      case CTU_PACK_PAIR_LIN: {  // param [co = a][ci = b]; dst [2 co][2 ci]: (s, o) x (s2, i), zero unless s == s2
        const int r = (int)(t / it.cols), c = (int)(t % it.cols);
        const int s = r / it.a, o = r - s * it.a, s2 = c / it.b, i = c - s2 * it.b;
        dst[t] = __float2bfloat16((s < 2 && s2 < 2 && s == s2) ? src[(long long)o * it.b + i] : 0.f);
        break;
      }
      case CTU_PACK_PAIR_LIN_T: {  // dst [2 ci][2 co]: (s2, i) x (s, o)
        const int r = (int)(t / it.cols), c = (int)(t % it.cols);
        const int s2 = r / it.b, i = r - s2 * it.b, s = c / it.a, o = c - s * it.a;
        dst[t] = __float2bfloat16((s < 2 && s2 < 2 && s == s2) ? src[(long long)o * it.b + i] : 0.f);
        break;
      }
      case CTU_PACK_PAIR_CONV3: {  // param [co = a][ci = b][27]; dst [2 co][27 * 2 ci]: (s, o) x (ptap, s2, i)
        const int r = (int)(t / it.cols), c = (int)(t % it.cols);
        const int s = r / it.a, o = r - s * it.a;
        const int ptap = c / (2 * it.b), rem = c - ptap * 2 * it.b, s2 = rem / it.b, i = rem - s2 * it.b;
        const int pz = ptap % 3, dz = 2 * (pz - 1) + s2 - s;
        float v = 0.f;
        if (s < 2 && ptap < 27 && dz >= -1 && dz <= 1) v = src[((long long)o * it.b + i) * 27 + (ptap - pz) + dz + 1];
        dst[t] = __float2bfloat16(v);
        break;
      }
      case CTU_PACK_PAIR_CONV3_T: {  // tap-flipped transpose for the input gradient: dst [2 ci][27 * 2 co]: (s2, i) x (26 - ptap, s, o)
        const int r = (int)(t / it.cols), c = (int)(t % it.cols);
        const int s2 = r / it.b, i = r - s2 * it.b;
        const int ft = c / (2 * it.a), rem = c - ft * 2 * it.a, s = rem / it.a, o = rem - s * it.a;
        const int ptap = 26 - ft, pz = ptap % 3, dz = 2 * (pz - 1) + s2 - s;
        float v = 0.f;
        if (s2 < 2 && ft < 27 && dz >= -1 && dz <= 1) v = src[((long long)o * it.b + i) * 27 + (ptap - pz) + dz + 1];
        dst[t] = __float2bfloat16(v);
        break;
      }
      case CTU_PACK_X3_FROM_PACKED: {  // src bf16 [a = n_pad][27 * b], b = a_c = cols; dst rows (t21, n tile, dxi, 64)
        const __nv_bfloat16* sb = reinterpret_cast<const __nv_bfloat16*>(it.src);
        const int r = (int)(t / it.cols), c = (int)(t % it.cols);
        const int per_tap = (it.a / 64) * 192;
        const int t21 = r / per_tap, rem = r - t21 * per_tap, nt = rem / 192, dxi = (rem % 192) / 64, col = rem % 64;
        dst[t] = sb[(long long)(nt * 64 + col) * 27 * it.b + (long long)((2 - dxi) * 9 + t21) * it.b + c];
        break;
      }
      default: break;
    }
  }
}

// dst (fp32, parameter layout) <- src (fp32 gradient accumulator in the transposed-packed layout, row pitch `cols`)
__global__ void __launch_bounds__(256) unpack_grads_kernel(const ctu_pack_item* __restrict__ items, int n_items,
                                                           long long total_units) {
  __shared__ __align__(16) float tile_sm[TILE_SM_FLOATS];
  for (long long unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
    const ctu_pack_item it = items[find_item(items, n_items, unit)];
    if (it.kind == CTU_PACK_LIN) { unpack_lin_tile(it, unit - it.unit0, tile_sm); continue; }
    if (it.kind == CTU_PACK_CONV3) { unpack_conv3_tile(it, unit - it.unit0, tile_sm); continue; }
    const long long t = (unit - it.unit0) * 256 + threadIdx.x;
    if (t >= unpack_tasks(it.kind, it.rows, it.a, it.b, it.c)) continue;
    const float* buf = reinterpret_cast<const float*>(it.src);
    float* g = reinterpret_cast<float*>(it.dst);
    const long long ld = it.cols;
    switch (it.kind) {
      case CTU_PACK_CONVT: {  // param [ci = a][co = b][k3 = c]; buf [ci][k3 * co]: (ci, co) pair, co fastest
        const int o = (int)(t % it.b), ci = (int)(t / it.b);
        float* gp = g + ((long long)ci * it.b + o) * it.c;
        for (int s = 0; s < it.c; ++s) gp[s] = buf[(long long)ci * ld + (long long)s * it.b + o];
        break;
      }
      case CTU_PACK_PS: {  // param [co = a][corg = b], k3 = c; buf [(cc * k3 + s)][(s * co + o)]
        const int cc = (int)(t % it.b);
        const long long o = t / it.b;
        float acc = 0.f;
        for (int s = 0; s < it.c; ++s) acc += buf[((long long)cc * it.c + s) * ld + (long long)s * it.a + o];
        g[t] = acc;
        break;
      }
      case CTU_PACK_PS_BIAS: {  // param [co = a], k3 = c; buf [k3 * co]
        float acc = 0.f;
        for (int s = 0; s < it.c; ++s) acc += buf[(long long)s * it.a + t];
        g[t] = acc;
        break;
      }
      case CTU_PACK_CIN1: {  // param [co = a][taps = b]; buf [taps'][ld]
        const long long tp = t % it.b, o = t / it.b;
        g[t] = buf[tp * ld + o];
        break;
      }
      case CTU_PACK_VEC:
        g[t] = buf[t];
        break;
      case CTU_PACK_PAIR_LIN: {  // param [co = a][ci = b]; buf [(s, i)][(s, o)]: the two diagonal blocks
        const int i = (int)(t % it.b), o = (int)(t / it.b);
        g[t] = buf[(long long)i * ld + o] + buf[(long long)(it.b + i) * ld + it.a + o];
        break;
      }
      case CTU_PACK_PAIR_CONV3: {  // param [co = a][ci = b][27]; buf [(ptap, s2, i)][(s, o)]: every real tap lives in two blocks
        const int tap = (int)(t % 27), i = (int)((t / 27) % it.b), o = (int)(t / (27LL * it.b));
        const int dz = tap % 3 - 1, t32 = tap - (tap % 3);   // t32 = (t3 * 3 + t2) * 3
        float acc = 0.f;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
#pragma unroll
          for (int s2 = 0; s2 < 2; ++s2) {
            const int num = dz - s2 + s;            // = 2 (pz - 1)
            if (num == -2 || num == 0 || num == 2) {
              const int pz = num / 2 + 1;
              acc += buf[((long long)(t32 + pz) * 2 * it.b + s2 * it.b + i) * ld + s * it.a + o];
            }
          }
        }
        g[t] = acc;
        break;
      }
      default: break;
    }
  }
}

// stats [B][ld][width] (fp64): columns c and c + half describe the same channel (paired rows): both become
// (v[c] + v[c + half]) * scale — scale 0.5 keeps "sum / rows" the channel mean for kernels that divide by the row count.
__global__ void stats_fold_kernel(double* __restrict__ stats, int B, int ld, int half, int width, double scale) {
  const int n = B * half * width;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const int w = t % width, c = (t / width) % half, b = t / (width * half);
    double* lo = stats + ((long long)b * ld + c) * width + w;
    double* hi = lo + (long long)half * width;
    const double v = (*lo + *hi) * scale;
    *lo = v;
    *hi = v;
  }
}

static int pack_grid(long long units) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long g = units < (long long)sms * 16 ? units : (long long)sms * 16;
  return (int)(g < 1 ? 1 : g);
}

}  // namespace ctu

extern "C" long long ctu_pack_item_tasks(int unpack, int kind, int rows, int cols, int a, int b, int c) {
  return unpack ? ctu::unpack_tasks(kind, rows, a, b, c) : ctu::pack_tasks(kind, rows, cols, a, b, c);
}

extern "C" int ctu_pack_weights(const ctu_pack_item* items_dev, int n_items, long long total_units, void* stream) {
  if (!items_dev || n_items <= 0 || total_units <= 0) return CTU_E_BADARG;
  ctu::pack_weights_kernel<<<ctu::pack_grid(total_units), 256, 0, (cudaStream_t)stream>>>(items_dev, n_items, total_units);
  ctu::count_launch();
  return (int)cudaGetLastError();
}

extern "C" int ctu_unpack_grads(const ctu_pack_item* items_dev, int n_items, long long total_units, void* stream) {
  if (!items_dev || n_items <= 0 || total_units <= 0) return CTU_E_BADARG;
  ctu::unpack_grads_kernel<<<ctu::pack_grid(total_units), 256, 0, (cudaStream_t)stream>>>(items_dev, n_items, total_units);
  ctu::count_launch();
  return (int)cudaGetLastError();
}

extern "C" int ctu_stats_fold(double* stats, int B, int ld, int half, int width, double scale, void* stream) {
  if (!stats || B <= 0 || half <= 0 || width <= 0 || ld < 2 * half) return CTU_E_BADARG;
  const int n = B * half * width;
  ctu::stats_fold_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(stats, B, ld, half, width, scale);
  ctu::count_launch();
  return (int)cudaGetLastError();
}
