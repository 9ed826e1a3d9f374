// Multi-tensor weight packing / gradient unpacking (sm_100a).
//
// The tensor-core kernels consume bf16 weights in their own layouts ([N_pad][taps*C_in] K-major for the forward
// contraction, transposed / tap-flipped for the input-gradient contraction) and produce weight gradients in the
// transposed-packed fp32 layout [(tap, c_in)][c_out].  A training step therefore re-packs ~400 parameters (twice) and
// unpacks ~400 gradients; done with torch ops that is ~2,600 tiny kernels per step.  Here each direction is ONE launch
// over a device-resident table of items: a block finds its item by binary search over the items' first work unit
// (256 thread-tasks per unit).  A thread-task moves the innermost run of its index map (8 consecutive K values, the 27
// taps of one (c_out, c_in) pair, the k^3 sub-voxels of one (c_in, c_out) pair ...) so that every 32-byte sector it
// touches on the strided side is fully used and the other side is coalesced across the warp.
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
    case CTU_PACK_LIN_T: return (long long)((rows + 7) / 8) * cols;
    case CTU_PACK_CONV3: return (long long)rows * (cols / 27);
    case CTU_PACK_CONV3_T: return (long long)rows * (cols / 27);
    case CTU_PACK_CONVT: return (long long)a * b;
    case CTU_PACK_CONVT_T: return (long long)a * b;
    default: return (long long)rows * cols;  // PS, PS_T, CIN1: one element per task
  }
}

__host__ __device__ inline long long unpack_tasks(int kind, int rows /*param elements*/, int a, int b, int c) {
  switch (kind) {
    case CTU_PACK_LIN: return (long long)a * ((b + 7) / 8);
    case CTU_PACK_CONV3: return (long long)a * b;
    case CTU_PACK_CONVT: return (long long)a * b;
    default: return rows;  // PS, PS_BIAS, CIN1, VEC: one element per task
  }
}

// dst (bf16) [rows][cols] <- src (fp32 parameter)
__global__ void __launch_bounds__(256) pack_weights_kernel(const ctu_pack_item* __restrict__ items, int n_items,
                                                           long long total_units) {
  for (long long unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
    const ctu_pack_item it = items[find_item(items, n_items, unit)];
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
      case CTU_PACK_LIN_T: {  // dst[k][n] = src[n][k]: 8 consecutive k of one n (n fastest across the warp)
        const int n = (int)(t % it.cols), k0 = (int)(t / it.cols) * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int k = k0 + j;
          if (k < it.rows) dst[(long long)k * it.cols + n] = __float2bfloat16((n < it.a && k < it.b) ? src[(long long)n * it.b + k] : 0.f);
        }
        break;
      }
      case CTU_PACK_CONV3: {  // a = co, b = ci: the 27 taps of one (co, ci) pair (ci fastest)
        const int cip = it.cols / 27;
        const int r = (int)(t / cip), ci = (int)(t % cip);
        const bool ok = r < it.a && ci < it.b;
        const float* sp = src + ((long long)r * it.b + ci) * 27;
#pragma unroll
        for (int tap = 0; tap < 27; ++tap)
          dst[(long long)r * it.cols + tap * cip + ci] = __float2bfloat16(ok ? sp[tap] : 0.f);
        break;
      }
      case CTU_PACK_CONV3_T: {  // rows = cip, cols = 27 * cop: taps flipped; (ci row, co) pair (co fastest)
        const int cop = it.cols / 27;
        const int r = (int)(t / cop), o = (int)(t % cop);
        const bool ok = r < it.b && o < it.a;
        const float* sp = src + ((long long)o * it.b + r) * 27;
#pragma unroll
        for (int tap = 0; tap < 27; ++tap)
          dst[(long long)r * it.cols + tap * cop + o] = __float2bfloat16(ok ? sp[26 - tap] : 0.f);
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
      default: break;
    }
  }
}

// dst (fp32, parameter layout) <- src (fp32 gradient accumulator in the transposed-packed layout, row pitch `cols`)
__global__ void __launch_bounds__(256) unpack_grads_kernel(const ctu_pack_item* __restrict__ items, int n_items,
                                                           long long total_units) {
  for (long long unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
    const ctu_pack_item it = items[find_item(items, n_items, unit)];
    const long long t = (unit - it.unit0) * 256 + threadIdx.x;
    if (t >= unpack_tasks(it.kind, it.rows, it.a, it.b, it.c)) continue;
    const float* buf = reinterpret_cast<const float*>(it.src);
    float* g = reinterpret_cast<float*>(it.dst);
    const long long ld = it.cols;
    switch (it.kind) {
      case CTU_PACK_LIN: {  // param [N = a][K = b]; buf [K'][ld]: 8 consecutive k of one n (n fastest)
        const int n = (int)(t % it.a), k0 = (int)(t / it.a) * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (k0 + j < it.b) g[(long long)n * it.b + k0 + j] = buf[(long long)(k0 + j) * ld + n];
        break;
      }
      case CTU_PACK_CONV3: {  // param [co = a][ci = b][27]; buf [(tap * cip + ci)][ld], cip = c: (ci, co) pair, co fastest
        const int o = (int)(t % it.a), ci = (int)(t / it.a);
        float* gp = g + ((long long)o * it.b + ci) * 27;
#pragma unroll
        for (int tap = 0; tap < 27; ++tap) gp[tap] = buf[((long long)tap * it.c + ci) * ld + o];
        break;
      }
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
      default: break;
    }
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
