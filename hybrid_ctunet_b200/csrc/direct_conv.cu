// Direct Conv3d for single-channel inputs (sm_100a CUDA cores): the ResNet stem 1->64 k7x7x7 s(2,2,1) p3
// (resnet.py:150-155) and vit_encoder0's 1->64 k3 / k1 convs (hybrid_CTUNet.py:57-83 with in_channels=1).
// K = 343 / 27 / 1 with C_in = 1 is bandwidth/latency bound, so it stays off the tensor cores:
// one thread per output voxel, 64 fp32 accumulators, filter taps broadcast from shared memory, fp32 input read
// through L1 (neighbouring voxels share taps), 128-byte bf16 channels-last store per voxel.
#include "common.cuh"
#include "../../include/ctunet_b200.h"
#include "host_util.h"

namespace ctu {

constexpr int DC_COUT = 64;

__global__ void __launch_bounds__(128) conv_cin1_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        __nv_bfloat16* __restrict__ out, int ldo, int B, int X, int Y,
                                                        int Z, int Xo, int Yo, int Zo, int kx, int ky, int kz, int sx,
                                                        int sy, int sz, int px, int py, int pz) {
  extern __shared__ __align__(16) float wsm[];  // [taps][64]
  const int taps = kx * ky * kz;
  for (int i = threadIdx.x; i < taps * DC_COUT; i += blockDim.x) wsm[i] = w[i];
  __syncthreads();
  const long long total = (long long)B * Xo * Yo * Zo;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  long long r = idx;
  const int oz = (int)(r % Zo); r /= Zo;
  const int oy = (int)(r % Yo); r /= Yo;
  const int ox = (int)(r % Xo);
  const int b = (int)(r / Xo);
  float acc[DC_COUT];
#pragma unroll
  for (int c = 0; c < DC_COUT; ++c) acc[c] = 0.f;
  const float* xb = x + (long long)b * X * Y * Z;
  for (int tx = 0; tx < kx; ++tx) {
    const int ix = ox * sx + tx - px;
    if (ix < 0 || ix >= X) continue;
    for (int ty = 0; ty < ky; ++ty) {
      const int iy = oy * sy + ty - py;
      if (iy < 0 || iy >= Y) continue;
      const float* row = xb + ((long long)ix * Y + iy) * Z;
      const float* wrow = wsm + (size_t)((tx * ky + ty) * kz) * DC_COUT;
      for (int tz = 0; tz < kz; ++tz) {
        const int iz = oz * sz + tz - pz;
        if (iz < 0 || iz >= Z) continue;
        const float v = __ldg(row + iz);
        const float4* wp = reinterpret_cast<const float4*>(wrow + tz * DC_COUT);
#pragma unroll
        for (int c4 = 0; c4 < DC_COUT / 4; ++c4) {
          const float4 ww = wp[c4];
          acc[4 * c4 + 0] = fmaf(v, ww.x, acc[4 * c4 + 0]);
          acc[4 * c4 + 1] = fmaf(v, ww.y, acc[4 * c4 + 1]);
          acc[4 * c4 + 2] = fmaf(v, ww.z, acc[4 * c4 + 2]);
          acc[4 * c4 + 3] = fmaf(v, ww.w, acc[4 * c4 + 3]);
        }
      }
    }
  }
  __nv_bfloat16* op = out + idx * ldo;
#pragma unroll
  for (int c = 0; c < DC_COUT; c += 8) {
    *reinterpret_cast<uint4*>(op + c) = make_uint4(pack_bf16x2(acc[c], acc[c + 1]), pack_bf16x2(acc[c + 2], acc[c + 3]),
                                                   pack_bf16x2(acc[c + 4], acc[c + 5]), pack_bf16x2(acc[c + 6], acc[c + 7]));
  }
}

// First and second moment of the single-channel input per batch item: sums over S voxels into mom[b] = (sum x, sum x^2)
// (fp64 atomics; mom zeroed by the caller).
__global__ void __launch_bounds__(256) cin1_moments_kernel(const float* __restrict__ x, long long S, double* __restrict__ mom) {
  __shared__ double red[2][8];
  const float* xb = x + (long long)blockIdx.y * S;
  double s1 = 0.0, s2 = 0.0;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < S; v += (long long)gridDim.x * blockDim.x) {
    const double t = (double)__ldg(xb + v);
    s1 += t;
    s2 += t * t;
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s1; red[1][threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x < 2) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[threadIdx.x][k];
    atomicAdd(mom + 2 * blockIdx.y + threadIdx.x, t);
  }
}

// InstanceNorm sums of a 1 -> C pointwise convolution r[v][c] = x[v] * w[c] without reading r: sum r = w_c sum x,
// sum r^2 = w_c^2 sum x^2.
__global__ void cin1_k1_stats_kernel(const double* __restrict__ mom, const float* __restrict__ w, int B, int C,
                                     double* __restrict__ stats, int ld) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int b = i / C, c = i - b * C;
  const double wc = (double)w[c];
  stats[((long long)b * ld + c) * 2] = wc * mom[2 * b];
  stats[((long long)b * ld + c) * 2 + 1] = wc * wc * mom[2 * b + 1];
}

}  // namespace ctu

using namespace ctu;

extern "C" int ctu_cin1_k1_stats(const float* x, const float* w, int B, long long S, int C, double* mom, double* stats,
                                 int stats_ld, void* stream) {
  if (!x || !w || !mom || !stats || B <= 0 || S <= 0 || C <= 0 || stats_ld < C) return CTU_E_BADARG;
  long long nb = (S + 256 * 16 - 1) / (256 * 16);
  if (nb > 592) nb = 592;
  cin1_moments_kernel<<<dim3((unsigned)nb, (unsigned)B), 256, 0, (cudaStream_t)stream>>>(x, S, mom);
  cin1_k1_stats_kernel<<<(B * C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(mom, w, B, C, stats, stats_ld);
  count_launch(2);
  return (int)cudaGetLastError();
}

// x: fp32 [B][X][Y][Z] (the reference's [B,1,X,Y,Z]); w: fp32 [kx*ky*kz][64] (tap-major); out: bf16 channels-last
// [B][Xo][Yo][Zo][ldo], Xo = (X + 2*px - kx)/sx + 1 etc.
extern "C" int ctu_conv_cin1(const float* x, const float* w, void* out, int ldo, int cout, int B, int X, int Y, int Z,
                             int kx, int ky, int kz, int sx, int sy, int sz, int px, int py, int pz, void* stream) {
  if (!x || !w || !out || cout != DC_COUT || ldo % 8 || ldo < DC_COUT) return CTU_E_BADARG;
  const int Xo = (X + 2 * px - kx) / sx + 1, Yo = (Y + 2 * py - ky) / sy + 1, Zo = (Z + 2 * pz - kz) / sz + 1;
  if (Xo <= 0 || Yo <= 0 || Zo <= 0) return CTU_E_BADARG;
  const size_t smem = (size_t)kx * ky * kz * DC_COUT * sizeof(float);
  constexpr int kMaxSmem = 96 * 1024;
  if (smem > (size_t)kMaxSmem) return CTU_E_UNSUPPORTED;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_cin1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const long long total = (long long)B * Xo * Yo * Zo;
  const unsigned grid = (unsigned)((total + 127) / 128);
  conv_cin1_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(x, w, (__nv_bfloat16*)out, ldo, B, X, Y, Z, Xo, Yo, Zo, kx,
                                                             ky, kz, sx, sy, sz, px, py, pz);
  count_launch();
  return (int)cudaGetLastError();
}
