// Multi-tensor AdamW (sm_100a, HBM-bound): the optimizer step main_CTUNet.py:190-193 hands to torch.optim.AdamW,
// as ONE launch over a device-resident item table (one item per parameter that has a gradient).  Decoupled weight
// decay, bias-corrected first / second moments, no amsgrad — the update of torch.optim.AdamW, term for term:
//   p  <- p * (1 - lr * wd)
//   m  <- b1 * m + (1 - b1) * g          v <- b2 * v + (1 - b2) * g * g
//   p  <- p - (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps),     bc1 = 1 - b1^t, bc2 = 1 - b2^t  (host, double)
// 28 bytes of traffic per parameter (g, p, m, v read; p, m, v written); a work unit is 1024 consecutive elements of
// one item (256 threads x 4), found by binary search over the items' first unit; 16-byte vectors when the four
// pointers of the item allow (gradients are slices of one flat buffer and may be only 4-byte aligned).
#include "common.cuh"
#include "../../include/ctunet_b200.h"
#include "host_util.h"

namespace ctu {

// the operations of torch's single-tensor AdamW, in its order: mul_(1 - lr wd); lerp_(g, 1 - b1); mul_(b2).addcmul_(g, g,
// 1 - b2); (sqrt / sqrt(bc2)).add_(eps); addcdiv_(m, denom, -lr / bc1)
__device__ __forceinline__ void adamw_one(float& p, float g, float& m, float& v, float decay, float w1, float b2, float w2,
                                          float step_size, float sqrt_bc2, float eps) {
  p *= decay;
  m = m + w1 * (g - m);
  v = v * b2 + (w2 * g) * g;
  const float denom = sqrtf(v) / sqrt_bc2 + eps;
  p = p - step_size * (m / denom);
}

__global__ void __launch_bounds__(256) adamw_kernel(const ctu_adamw_item* __restrict__ items, int n_items,
                                                    long long total_units, float decay, float w1, float b2, float w2,
                                                    float step_size, float sqrt_bc2, float eps) {
  for (long long unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
    int lo = 0, hi = n_items - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (items[mid].unit0 <= unit) lo = mid; else hi = mid - 1;
    }
    const ctu_adamw_item it = items[lo];
    const long long i0 = (unit - it.unit0) * 1024 + threadIdx.x * 4;
    if (i0 >= it.numel) continue;
    float* p = reinterpret_cast<float*>(it.param) + i0;
    const float* g = reinterpret_cast<const float*>(it.grad) + i0;
    float* m = reinterpret_cast<float*>(it.exp_avg) + i0;
    float* v = reinterpret_cast<float*>(it.exp_avg_sq) + i0;
    const bool vec = i0 + 4 <= it.numel &&
                     ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(v)) & 15) == 0;
    if (vec) {
      float4 pp = *reinterpret_cast<float4*>(p), mm = *reinterpret_cast<float4*>(m), vv = *reinterpret_cast<float4*>(v);
      const float4 gg = *reinterpret_cast<const float4*>(g);
      adamw_one(pp.x, gg.x, mm.x, vv.x, decay, w1, b2, w2, step_size, sqrt_bc2, eps);
      adamw_one(pp.y, gg.y, mm.y, vv.y, decay, w1, b2, w2, step_size, sqrt_bc2, eps);
      adamw_one(pp.z, gg.z, mm.z, vv.z, decay, w1, b2, w2, step_size, sqrt_bc2, eps);
      adamw_one(pp.w, gg.w, mm.w, vv.w, decay, w1, b2, w2, step_size, sqrt_bc2, eps);
      *reinterpret_cast<float4*>(p) = pp;
      *reinterpret_cast<float4*>(m) = mm;
      *reinterpret_cast<float4*>(v) = vv;
    } else {
      for (int j = 0; j < 4 && i0 + j < it.numel; ++j) {
        float pp = p[j], mm = m[j], vv = v[j];
        adamw_one(pp, g[j], mm, vv, decay, w1, b2, w2, step_size, sqrt_bc2, eps);
        p[j] = pp; m[j] = mm; v[j] = vv;
      }
    }
  }
}

}  // namespace ctu

extern "C" int ctu_adamw_step(const ctu_adamw_item* items_dev, int n_items, long long total_units, double lr, double beta1,
                              double beta2, double eps, double weight_decay, long long step, void* stream) {
  if (!items_dev || n_items <= 0 || total_units <= 0 || step < 1 || !(beta1 >= 0.0 && beta1 < 1.0) ||
      !(beta2 >= 0.0 && beta2 < 1.0))
    return CTU_E_BADARG;
  // hyper-parameters arrive as the Python doubles torch works with: 1 - beta2 formed in double (0.001, not 1 - 0.999f)
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  const float step_size = (float)(lr / bc1);
  const float sqrt_bc2 = (float)sqrt(bc2);
  const float w1 = (float)(1.0 - beta1), w2 = (float)(1.0 - beta2);
  const float decay = (float)(1.0 - lr * weight_decay);
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long grid = total_units < (long long)sms * 16 ? total_units : (long long)sms * 16;
  ctu::adamw_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(items_dev, n_items, total_units, decay, w1, (float)beta2,
                                                                     w2, step_size, sqrt_bc2, (float)eps);
  ctu::count_launch();
  return (int)cudaGetLastError();
}
