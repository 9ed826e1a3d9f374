// Multi-tensor weight packing / gradient unpacking (sm_100a).
//
// The tensor-core kernels consume bf16 weights in their own layouts ([N_pad][taps*C_in] K-major for the forward
// contraction, transposed / tap-flipped for the input-gradient contraction) and produce weight gradients in the
// transposed-packed fp32 layout [(tap, c_in)][c_out].  A training step therefore re-packs ~400 parameters (twice) and
// unpacks ~400 gradients; done with torch ops that is ~2,600 tiny kernels and 7.5 ms per step.  Here each direction is
// ONE launch over a device-resident table of items: a block finds its item by binary search over the items' first
// work unit (256 elements per unit) and evaluates the item's index map.
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

// dst (bf16) [rows][cols] <- src (fp32 parameter)
__global__ void __launch_bounds__(256) pack_weights_kernel(const ctu_pack_item* __restrict__ items, int n_items,
                                                           long long total_units) {
  for (long long unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
    const ctu_pack_item it = items[find_item(items, n_items, unit)];
    const long long e = (unit - it.unit0) * 256 + threadIdx.x;
    const long long total = (long long)it.rows * it.cols;
    if (e >= total) continue;
    const int r = (int)(e / it.cols), c = (int)(e % it.cols);
    const float* src = reinterpret_cast<const float*>(it.src);
    float v = 0.f;
    switch (it.kind) {
      case CTU_PACK_LIN:  // a = N, b = K
        if (r < it.a && c < it.b) v = src[(long long)r * it.b + c];
        break;
      case CTU_PACK_LIN_T:
        if (r < it.b && c < it.a) v = src[(long long)c * it.b + r];
        break;
      case CTU_PACK_CONV3: {  // a = co, b = ci; cols = 27 * cip
        const int cip = it.cols / 27, tap = c / cip, ci = c % cip;
        if (r < it.a && ci < it.b) v = src[((long long)r * it.b + ci) * 27 + tap];
        break;
      }
      case CTU_PACK_CONV3_T: {  // rows = cip, cols = 27 * cop; taps flipped
        const int cop = it.cols / 27, tap = c / cop, o = c % cop;
        if (r < it.b && o < it.a) v = src[((long long)o * it.b + r) * 27 + (26 - tap)];
        break;
      }
      case CTU_PACK_CONVT: {  // param [ci = a][co = b][k3 = c]; rows = k3 * co, cols = ci
        const int s = r / it.b, o = r % it.b;
        v = src[((long long)c * it.b + o) * it.c + s];
        break;
      }
      case CTU_PACK_CONVT_T: {  // rows = ci, cols = k3 * co
        const int s = c / it.b, o = c % it.b;
        v = src[((long long)r * it.b + o) * it.c + s];
        break;
      }
      case CTU_PACK_PS: {  // param [co = a][corg = b], k3 = c; rows = k3 * co, cols = corg * k3 (block diagonal in s)
        const int s = r / it.a, o = r % it.a, cc = c / it.c, s2 = c % it.c;
        if (s == s2) v = src[(long long)o * it.b + cc];
        break;
      }
      case CTU_PACK_PS_T: {
        const int s = c / it.a, o = c % it.a, cc = r / it.c, s2 = r % it.c;
        if (s == s2) v = src[(long long)o * it.b + cc];
        break;
      }
      case CTU_PACK_CIN1:  // param [co = a][taps = b]; cols = taps padded
        if (r < it.a && c < it.b) v = src[(long long)r * it.b + c];
        break;
      default: break;
    }
    reinterpret_cast<__nv_bfloat16*>(it.dst)[e] = __float2bfloat16(v);
  }
}

// dst (fp32, parameter layout) <- src (fp32 gradient accumulator in the transposed-packed layout, row pitch `cols`)
__global__ void __launch_bounds__(256) unpack_grads_kernel(const ctu_pack_item* __restrict__ items, int n_items,
                                                           long long total_units) {
  for (long long unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
    const ctu_pack_item it = items[find_item(items, n_items, unit)];
    const long long e = (unit - it.unit0) * 256 + threadIdx.x;
    if (e >= (long long)it.rows) continue;  // rows = number of parameter elements
    const float* buf = reinterpret_cast<const float*>(it.src);
    const long long ld = it.cols;
    float g = 0.f;
    switch (it.kind) {
      case CTU_PACK_LIN: {  // param [N = a][K = b]; buf [K'][ld]
        const long long n = e / it.b, k = e % it.b;
        g = buf[k * ld + n];
        break;
      }
      case CTU_PACK_CONV3: {  // param [co = a][ci = b][27]; buf [(tap * cip + ci)][ld], cip = c
        const int tap = (int)(e % 27);
        const long long t = e / 27;
        const int ci = (int)(t % it.b);
        const long long o = t / it.b;
        g = buf[((long long)tap * it.c + ci) * ld + o];
        break;
      }
      case CTU_PACK_CONVT: {  // param [ci = a][co = b][k3 = c]; buf [ci][k3 * co]
        const int s = (int)(e % it.c);
        const long long t = e / it.c;
        const int o = (int)(t % it.b);
        const long long ci = t / it.b;
        g = buf[ci * ld + (long long)s * it.b + o];
        break;
      }
      case CTU_PACK_PS: {  // param [co = a][corg = b], k3 = c; buf [(cc * k3 + s)][(s * co + o)]
        const int cc = (int)(e % it.b);
        const long long o = e / it.b;
        for (int s = 0; s < it.c; ++s) g += buf[((long long)cc * it.c + s) * ld + (long long)s * it.a + o];
        break;
      }
      case CTU_PACK_PS_BIAS:  // param [co = a], k3 = c; buf [k3 * co]
        for (int s = 0; s < it.c; ++s) g += buf[(long long)s * it.a + e];
        break;
      case CTU_PACK_CIN1: {  // param [co = a][taps = b]; buf [taps'][ld]
        const long long t = e % it.b, o = e / it.b;
        g = buf[t * ld + o];
        break;
      }
      case CTU_PACK_VEC:
        g = buf[e];
        break;
      default: break;
    }
    reinterpret_cast<float*>(it.dst)[e] = g;
  }
}

static int pack_grid(long long units) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long g = units < (long long)sms * 16 ? units : (long long)sms * 16;
  return (int)(g < 1 ? 1 : g);
}

}  // namespace ctu

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
