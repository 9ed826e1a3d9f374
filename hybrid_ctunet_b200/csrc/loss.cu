// Fused Dice-CE loss of the training step (trainer_CTUNet.py:92-103; monai.losses.DiceCELoss 0.7.0 with
// to_onehot_y, softmax, squared_pred): one pass over the fp32 NCDHW logits computes the per-voxel softmax and
// accumulates, per (batch item, class), sum p*y, sum p^2, sum y and the cross-entropy sum; the backward pass
// recomputes the softmax and writes dlogits in one more pass.  The reference runs ~15 torch kernels per head
// (softmax, one_hot, 3 reductions, log_softmax, nll_loss, ...) and their autograd counterparts.
// HBM-bound: algorithmic bytes = C*4 + 4 per voxel forward, 2*C*4 + 4 per voxel backward.
#include "common.cuh"
#include "../../include/ctunet_b200.h"
#include "host_util.h"

namespace ctu {

constexpr int LOSS_MAXC = 16;

// sums: double [B][C][3] = (sum p*y, sum p^2, sum y) followed by one double: sum over voxels of -log p[label]
template <int C>
__global__ void __launch_bounds__(256) dice_ce_fwd_kernel(const float* __restrict__ logits, const float* __restrict__ target,
                                                          long long S, double* __restrict__ sums, int B) {
  __shared__ float red[8][3 * C + 1];
  const int b = blockIdx.y;
  const float* lb = logits + (long long)b * C * S;
  const float* tb = target + (long long)b * S;
  float inter[C], psq[C], cnt[C];
#pragma unroll
  for (int c = 0; c < C; ++c) inter[c] = psq[c] = cnt[c] = 0.f;
  float ce = 0.f;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < S; v += (long long)gridDim.x * blockDim.x) {
    float x[C];
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < C; ++c) { x[c] = lb[(long long)c * S + v]; m = fmaxf(m, x[c]); }
    float den = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) { x[c] = __expf(x[c] - m); den += x[c]; }
    const float inv = 1.f / den;
    const int y = (int)tb[v];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float p = x[c] * inv;
      psq[c] += p * p;
      if (c == y) { inter[c] += p; cnt[c] += 1.f; ce -= __logf(fmaxf(p, 1e-38f)); }
    }
  }
  // block reduction: warp shuffles, then 8 warps through shared memory, then fp64 atomics
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < C; ++c) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      inter[c] += __shfl_xor_sync(0xffffffffu, inter[c], o);
      psq[c] += __shfl_xor_sync(0xffffffffu, psq[c], o);
      cnt[c] += __shfl_xor_sync(0xffffffffu, cnt[c], o);
    }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) ce += __shfl_xor_sync(0xffffffffu, ce, o);
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < C; ++c) { red[w][3 * c] = inter[c]; red[w][3 * c + 1] = psq[c]; red[w][3 * c + 2] = cnt[c]; }
    red[w][3 * C] = ce;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * C + 1; i += 256) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += (double)red[k][i];
    if (i < 3 * C) atomicAdd(sums + (long long)b * 3 * C + i, s);
    else atomicAdd(sums + (long long)B * 3 * C, s);
  }
}

// dlogit_c = p_c * (a_c - sum_k a_k p_k) + ce_scale * (p_c - y_c),  a_k = coef[b][k][0] * y_k + coef[b][k][1] * p_k
// (coef folds d loss / d dice_bk, the Dice denominators and the upstream gradient; built from `sums` on the device).
template <int C>
__global__ void __launch_bounds__(256) dice_ce_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ target,
                                                          long long S, const float* __restrict__ coef,
                                                          const float* __restrict__ ce_scale, float* __restrict__ dlogits) {
  const int b = blockIdx.y;
  const float* lb = logits + (long long)b * C * S;
  const float* tb = target + (long long)b * S;
  float* db = dlogits + (long long)b * C * S;
  float c0[C], c1[C];
#pragma unroll
  for (int c = 0; c < C; ++c) { c0[c] = coef[((long long)b * C + c) * 2]; c1[c] = coef[((long long)b * C + c) * 2 + 1]; }
  const float ces = *ce_scale;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < S; v += (long long)gridDim.x * blockDim.x) {
    float x[C];
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < C; ++c) { x[c] = lb[(long long)c * S + v]; m = fmaxf(m, x[c]); }
    float den = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) { x[c] = __expf(x[c] - m); den += x[c]; }
    const float inv = 1.f / den;
    const int y = (int)tb[v];
    float dot = 0.f;
    float a[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      x[c] *= inv;
      a[c] = (c == y ? c0[c] : 0.f) + c1[c] * x[c];
      dot += a[c] * x[c];
    }
#pragma unroll
    for (int c = 0; c < C; ++c) db[(long long)c * S + v] = x[c] * (a[c] - dot) + ces * (x[c] - (c == y ? 1.f : 0.f));
  }
}

static int loss_grid_x(long long S, int B) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long gx = (S + 255) / 256;
  const long long cap = ((long long)sms * 8 + B - 1) / B;
  if (gx > cap) gx = cap;
  return (int)(gx < 1 ? 1 : gx);
}

}  // namespace ctu

using namespace ctu;

extern "C" int ctu_dice_ce_fwd(const float* logits, const float* target, int B, int C, long long S, double* sums,
                               void* stream) {
  if (!logits || !target || !sums || B <= 0 || S <= 0) return CTU_E_BADARG;
  dim3 grid(loss_grid_x(S, B), B);
  cudaStream_t st = (cudaStream_t)stream;
  switch (C) {
    case 14: dice_ce_fwd_kernel<14><<<grid, 256, 0, st>>>(logits, target, S, sums, B); break;
    case 2: dice_ce_fwd_kernel<2><<<grid, 256, 0, st>>>(logits, target, S, sums, B); break;
    case 3: dice_ce_fwd_kernel<3><<<grid, 256, 0, st>>>(logits, target, S, sums, B); break;
    case 4: dice_ce_fwd_kernel<4><<<grid, 256, 0, st>>>(logits, target, S, sums, B); break;
    default: return CTU_E_UNSUPPORTED;
  }
  count_launch();
  return (int)cudaGetLastError();
}

extern "C" int ctu_dice_ce_bwd(const float* logits, const float* target, int B, int C, long long S, const float* coef,
                               const float* ce_scale, float* dlogits, void* stream) {
  if (!logits || !target || !coef || !ce_scale || !dlogits || B <= 0 || S <= 0) return CTU_E_BADARG;
  dim3 grid(loss_grid_x(S, B), B);
  cudaStream_t st = (cudaStream_t)stream;
  switch (C) {
    case 14: dice_ce_bwd_kernel<14><<<grid, 256, 0, st>>>(logits, target, S, coef, ce_scale, dlogits); break;
    case 2: dice_ce_bwd_kernel<2><<<grid, 256, 0, st>>>(logits, target, S, coef, ce_scale, dlogits); break;
    case 3: dice_ce_bwd_kernel<3><<<grid, 256, 0, st>>>(logits, target, S, coef, ce_scale, dlogits); break;
    case 4: dice_ce_bwd_kernel<4><<<grid, 256, 0, st>>>(logits, target, S, coef, ce_scale, dlogits); break;
    default: return CTU_E_UNSUPPORTED;
  }
  count_launch();
  return (int)cudaGetLastError();
}
