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

// One block: the scalar loss of up to CTU_LOSS_MAX_HEADS Dice-CE heads from their accumulated sums, plus everything the
// backward kernels need for a unit upstream gradient — replaces ~18 tiny torch kernels per head (slices, divisions, means,
// stacks) between the reduction pass and the gradient pass.  All arithmetic in double.
__global__ void __launch_bounds__(256) dice_ce_finalize_kernel(const ctu_loss_heads h, const double* __restrict__ sums,
                                                               float* __restrict__ loss, float* __restrict__ coef,
                                                               float* __restrict__ ce_scale) {
  __shared__ double part[256];
  double acc = 0.0;
  for (int hd = 0; hd < h.n_heads; ++hd) {
    const int B = h.B[hd], C = h.C[hd];
    const double* sh = sums + h.sums_off[hd];
    float* ch = coef + h.coef_off[hd];
    const double wd = h.weight[hd] * h.lambda_dice / (double)(B * C);
    for (int i = threadIdx.x; i < B * C; i += 256) {
      const double inter = sh[3 * i], den = sh[3 * i + 1] + sh[3 * i + 2] + h.smooth_dr;
      const double num = 2.0 * inter + h.smooth_nr;
      acc += wd * (1.0 - num / den);
      ch[2 * i] = (float)(wd * (-2.0 / den));
      ch[2 * i + 1] = (float)(wd * (2.0 * num / (den * den)));
    }
    if (threadIdx.x == 0) {
      const double wce = h.weight[hd] * h.lambda_ce / ((double)B * (double)h.S[hd]);
      acc += wce * sh[3 * B * C];
      ce_scale[hd] = (float)wce;
    }
  }
  part[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o >= 1; o >>= 1) {
    if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = (float)part[0];
}

// dst[b][xo][yo][zo] = src[b][ix[xo]][iy[yo]][iz[zo]] (0 where an index is negative): nearest-neighbour down-sampling of the label volume with the index
// tables scipy.ndimage.zoom(order=0) implies (trainer_CTUNet.py:93-94), one launch instead of one index_select per axis.
__global__ void __launch_bounds__(256) gather3d_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int X,
                                                       int Y, int Z, int Xo, int Yo, int Zo, const int* __restrict__ ix,
                                                       const int* __restrict__ iy, const int* __restrict__ iz) {
  const long long total = (long long)B * Xo * Yo * Zo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long v = i;
    const int zo = (int)(v % Zo); v /= Zo;
    const int yo = (int)(v % Yo); v /= Yo;
    const int xo = (int)(v % Xo);
    const int b = (int)(v / Xo);
    const int sx = ix[xo], sy = iy[yo], sz = iz[zo];      // a negative index = outside the volume (scipy's cval = 0)
    dst[i] = (sx | sy | sz) < 0 ? 0.f : src[(((long long)b * X + sx) * Y + sy) * Z + sz];
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

extern "C" int ctu_dice_ce_finalize(const ctu_loss_heads* heads, const double* sums, float* loss, float* coef,
                                    float* ce_scale, void* stream) {
  if (!heads || !sums || !loss || !coef || !ce_scale || heads->n_heads <= 0 || heads->n_heads > CTU_LOSS_MAX_HEADS)
    return CTU_E_BADARG;
  dice_ce_finalize_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(*heads, sums, loss, coef, ce_scale);
  count_launch();
  return (int)cudaGetLastError();
}

extern "C" int ctu_gather3d(const float* src, float* dst, int B, int X, int Y, int Z, int Xo, int Yo, int Zo, const int* ix,
                            const int* iy, const int* iz, void* stream) {
  if (!src || !dst || !ix || !iy || !iz || B <= 0 || Xo <= 0 || Yo <= 0 || Zo <= 0) return CTU_E_BADARG;
  const long long total = (long long)B * Xo * Yo * Zo;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long grid = (total + 255) / 256;
  if (grid > (long long)sms * 16) grid = (long long)sms * 16;
  gather3d_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(src, dst, B, X, Y, Z, Xo, Yo, Zo, ix, iy, iz);
  count_launch();
  return (int)cudaGetLastError();
}
