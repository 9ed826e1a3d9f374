// `Invertd` of the evaluation scripts (test_CTUNet.py:162-199, test_CTUNet_final.py:470-505) on the device: the blended
// logits of the cropped, resampled, RAS-oriented grid are carried back to the grid of the file.  The reference runs
// MONAI's inverse chain on the host, one full [14, X, Y, Z] float64 volume per step: CropForegroundd.inverse (zero pad),
// Spacingd.inverse (AffineTransform -> torch affine_grid + grid_sample, trilinear, border padding, align_corners=False,
// computed in float64 and stored as float32) and Orientationd.inverse (flips + transposes).  All three are index maps, so
// the host folds them into ONE 3x4 matrix (output voxel index -> fractional index in the padded grid) and one kernel
// gathers straight from the cropped prediction: out-of-crop corners read 0 (the pad), coordinates clamp to the padded grid
// (border).  The interpolation is done in float64 with grid_sample's corner order and weight products, then rounded to
// float32 — the reference's arithmetic.  ctu_invert_ensemble_argmax goes one step further and feeds the interpolated
// class scores of both models straight into the mask-complementation ensemble (ensemble.cu), so only uint8 masks (and
// the Dice counters) are written: the 2 x [14, X0, Y0, Z0] fp32 volumes of the reference never exist.
// HBM / L2-gather bound: 8 corner reads per class per output voxel (neighbouring voxels share them), 4 B (or 1-3 B) written.
#include "common.cuh"
#include "../../include/ctunet_b200.h"
#include "host_util.h"

namespace ctu {

struct InvertSample {
  long long base;     // offset of corner 0 inside one class plane of the cropped prediction (may lie outside: see valid)
  double w[8];        // grid_sample's weight of the corner
  unsigned valid;     // bit k: corner k lies inside the padded grid AND inside the crop (otherwise it contributes 0)
};

struct InvertGeomDev {
  double m[12];
  int out[3], pad[3], crop[3], pred[3];
  int mode;
};

// corner k = (kx, ky, kz) with kz the least significant bit: z is the fastest axis, which is grid_sample's "x", and
// grid_sample adds its corners in the order tnw, tne, tsw, tse, bnw, bne, bsw, bse = x fastest, then y, then z.
__device__ __forceinline__ void invert_sample(const InvertGeomDev& g, int o0, int o1, int o2, InvertSample& s) {
  double c[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    double v = g.m[a * 4 + 0] * o0 + g.m[a * 4 + 1] * o1 + g.m[a * 4 + 2] * o2 + g.m[a * 4 + 3];
    v = fmin(fmax(v, 0.0), (double)(g.pad[a] - 1));   // padding_mode = border: clip_coordinates
    c[a] = v;
  }
  s.valid = 0;
  if (g.mode == 0) {   // nearest: nearbyint (half to even), one corner
    long long off = 0;
    bool ok = true;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int i = (int)nearbyint(c[a]);
      const int p = i - g.crop[a];
      ok = ok && i >= 0 && i < g.pad[a] && p >= 0 && p < g.pred[a];
      off = off * g.pred[a] + p;
    }
    s.base = ok ? off : 0;
    s.w[0] = 1.0;
    s.valid = ok ? 1u : 0u;
    return;
  }
  int i0[3];
  double f1[3], f0[3];   // weight of the upper / lower corner along each axis: (x - x_low), (x_low + 1 - x)
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const double fl = floor(c[a]);
    i0[a] = (int)fl;
    f1[a] = c[a] - fl;
    f0[a] = (fl + 1.0) - c[a];
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int kx = (k >> 2) & 1, ky = (k >> 1) & 1, kz = k & 1;
    const int ix = i0[0] + kx, iy = i0[1] + ky, iz = i0[2] + kz;
    // grid_sample: (x weight * y weight) * z weight with x the fastest axis
    s.w[k] = ((kz ? f1[2] : f0[2]) * (ky ? f1[1] : f0[1])) * (kx ? f1[0] : f0[0]);
    const int px = ix - g.crop[0], py = iy - g.crop[1], pz = iz - g.crop[2];
    const bool ok = ix < g.pad[0] && iy < g.pad[1] && iz < g.pad[2] &&          // within_bounds_3d (lower corners are >= 0)
                    px >= 0 && px < g.pred[0] && py >= 0 && py < g.pred[1] && pz >= 0 && pz < g.pred[2];
    s.valid |= ok ? (1u << k) : 0u;
  }
  s.base = ((long long)(i0[0] - g.crop[0]) * g.pred[1] + (i0[1] - g.crop[1])) * g.pred[2] + (i0[2] - g.crop[2]);
}

// sx, sy: element strides of the x / y axes of the prediction (z is contiguous)
__device__ __forceinline__ float invert_value(const float* __restrict__ plane, const InvertSample& s, int mode, int sx, int sy) {
  const float* p = plane + s.base;
  if (mode == 0) return (s.valid & 1u) ? __ldg(p) : 0.f;
  float v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k)   // all eight loads in flight before the dependent float64 chain
    v[k] = (s.valid & (1u << k)) ? __ldg(p + ((k >> 2) & 1) * sx + ((k >> 1) & 1) * sy + (k & 1)) : 0.f;
  double acc = 0.0;
#pragma unroll
  for (int k = 0; k < 8; ++k)   // a corner that is out of bounds adds nothing in grid_sample; here it adds +0.0 * w
    acc = __dadd_rn(acc, __dmul_rn((double)v[k], s.w[k]));   // no FMA: the host kernel has none
  return (float)acc;
}

// the common case: trilinear, all eight corners inside the crop — four row pointers, immediate offsets, no predicates
__device__ __forceinline__ float invert_value_full(const float* __restrict__ p, const double (&w)[8], int sx, int sy) {
  const float* q = p + sy;
  const float* r = p + sx;
  const float* t = r + sy;
  const float v[8] = {__ldg(p), __ldg(p + 1), __ldg(q), __ldg(q + 1), __ldg(r), __ldg(r + 1), __ldg(t), __ldg(t + 1)};
  double acc = 0.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) acc = __dadd_rn(acc, __dmul_rn((double)v[k], w[k]));
  return (float)acc;
}

__global__ void __launch_bounds__(256, 4) invert_resample_kernel(const float* __restrict__ pred, int C, InvertGeomDev g,
                                                              float* __restrict__ out) {
  const unsigned V = (unsigned)g.out[0] * g.out[1] * g.out[2];   // < 2^31: checked by the caller
  const long long P = (long long)g.pred[0] * g.pred[1] * g.pred[2];
  const int sy = g.pred[2], sx = g.pred[1] * g.pred[2];
  for (unsigned v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) {
    const unsigned q = v / (unsigned)g.out[2];
    const int o2 = (int)(v - q * g.out[2]), o0 = (int)(q / (unsigned)g.out[1]), o1 = (int)(q - (unsigned)o0 * g.out[1]);
    InvertSample s;
    invert_sample(g, o0, o1, o2, s);
    float* o = out + v;
    if (g.mode == 1 && s.valid == 0xffu) {
      const float* p = pred + s.base;
#pragma unroll 2
      for (int c = 0; c < C; ++c, p += P, o += V) *o = invert_value_full(p, s.w, sx, sy);
    } else {
      for (int c = 0; c < C; ++c, o += V) *o = invert_value(pred + (long long)c * P, s, g.mode, sx, sy);
    }
  }
}

// One thread per output voxel.  The gather loop is the resample kernel's (two classes in flight, low register pressure) and
// parks the interpolated scores of both models in the thread's own shared-memory column; the softmax / mean / argmax stage
// then pulls them into registers — by then the sample's weights are dead.  From there on it is ensemble_kernel (ensemble.cu)
// statement for statement.
template <int C>
__global__ void __launch_bounds__(256, 4) invert_ensemble_kernel(const float* __restrict__ p1, const float* __restrict__ p2,
                                                              InvertGeomDev g, uint8_t* __restrict__ mask,
                                                              uint8_t* __restrict__ mask1, uint8_t* __restrict__ mask2,
                                                              const float* __restrict__ labels,
                                                              unsigned long long* __restrict__ counts) {
  __shared__ unsigned int sc[3 * C * 3];
  __shared__ float park[2][C][256];
  for (int i = threadIdx.x; i < 3 * C * 3; i += 256) sc[i] = 0;
  __syncthreads();
  const unsigned V = (unsigned)g.out[0] * g.out[1] * g.out[2];
  const long long P = (long long)g.pred[0] * g.pred[1] * g.pred[2];
  const int sy = g.pred[2], sx = g.pred[1] * g.pred[2];
  const int tid = threadIdx.x;
  for (unsigned v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) {
    const unsigned q = v / (unsigned)g.out[2];
    const int o2 = (int)(v - q * g.out[2]), o0 = (int)(q / (unsigned)g.out[1]), o1 = (int)(q - (unsigned)o0 * g.out[1]);
    {
      InvertSample s;
      invert_sample(g, o0, o1, o2, s);
      if (g.mode == 1 && s.valid == 0xffu) {
        const float* pa = p1 + s.base;
        const float* pb = p2 + s.base;
#pragma unroll 2
        for (int c = 0; c < C; ++c, pa += P, pb += P) {
          park[0][c][tid] = invert_value_full(pa, s.w, sx, sy);
          park[1][c][tid] = invert_value_full(pb, s.w, sx, sy);
        }
      } else {
        for (int c = 0; c < C; ++c) {
          park[0][c][tid] = invert_value(p1 + (long long)c * P, s, g.mode, sx, sy);
          park[1][c][tid] = invert_value(p2 + (long long)c * P, s, g.mode, sx, sy);
        }
      }
    }
    float a[C], b[C];
    float ma = -INFINITY, mb = -INFINITY;
    int ia = 0, ib = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      a[c] = park[0][c][tid];
      b[c] = park[1][c][tid];
      if (a[c] > ma) { ma = a[c]; ia = c; }   // first maximal index, like torch.argmax
      if (b[c] > mb) { mb = b[c]; ib = c; }
    }
    float sa = 0.f, sb = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) { a[c] = expf(a[c] - ma); sa += a[c]; b[c] = expf(b[c] - mb); sb += b[c]; }
    const float ra = 1.f / sa, rb = 1.f / sb;
    float best = -INFINITY;
    int ie = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float m = __fadd_rn(__fmul_rn(a[c], ra), __fmul_rn(b[c], rb)) / 2.0f;
      if (m > best) { best = m; ie = c; }
    }
    if (mask) mask[v] = (uint8_t)ie;
    if (mask1) mask1[v] = (uint8_t)ia;
    if (mask2) mask2[v] = (uint8_t)ib;
    if (labels != nullptr && counts != nullptr) {
      const int y = (int)labels[v];
      const int pr[3] = {ie, ia, ib};
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        atomicAdd(&sc[(k * C + pr[k]) * 3 + 1], 1u);
        if (pr[k] == y) atomicAdd(&sc[(k * C + y) * 3 + 0], 1u);
      }
      if (y >= 0 && y < C) {
#pragma unroll
        for (int k = 0; k < 3; ++k) atomicAdd(&sc[(k * C + y) * 3 + 2], 1u);
      }
    }
  }
  __syncthreads();
  if (labels != nullptr && counts != nullptr)
    for (int i = threadIdx.x; i < 3 * C * 3; i += 256)
      if (sc[i]) atomicAdd(counts + i, (unsigned long long)sc[i]);
}

static int invert_geom(const ctu_invert_geom* g, InvertGeomDev& d) {
  if (!g) return CTU_E_BADARG;
  for (int i = 0; i < 12; ++i) d.m[i] = g->m[i];
  for (int a = 0; a < 3; ++a) {
    d.out[a] = g->out_size[a];
    d.pad[a] = g->pad_size[a];
    d.crop[a] = g->crop_start[a];
    d.pred[a] = g->pred_size[a];
    if (d.out[a] <= 0 || d.pad[a] <= 0 || d.pred[a] <= 0) return CTU_E_BADARG;
  }
  d.mode = g->mode;
  if (d.mode != 0 && d.mode != 1) return CTU_E_UNSUPPORTED;
  if ((long long)d.out[0] * d.out[1] * d.out[2] > 0x7fffffffLL || (long long)d.pred[0] * d.pred[1] * d.pred[2] > 0x7fffffffLL)
    return CTU_E_UNSUPPORTED;   // 32-bit voxel indices inside one class plane
  return 0;
}

static int invert_grid(long long V, int per_sm) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long grid = (V + 255) / 256;
  if (grid > (long long)sms * per_sm) grid = (long long)sms * per_sm;
  return (int)grid;
}

}  // namespace ctu

extern "C" int ctu_invert_resample(const float* pred, int C, const ctu_invert_geom* geom, float* out, void* stream) {
  using namespace ctu;
  InvertGeomDev g;
  const int rc = invert_geom(geom, g);
  if (rc != 0) return rc;
  if (!pred || !out || C <= 0) return CTU_E_BADARG;
  const long long V = (long long)g.out[0] * g.out[1] * g.out[2];
  invert_resample_kernel<<<invert_grid(V, 16), 256, 0, (cudaStream_t)stream>>>(pred, C, g, out);
  count_launch();
  return (int)cudaGetLastError();
}

extern "C" int ctu_invert_ensemble_argmax(const float* p1, const float* p2, int C, const ctu_invert_geom* geom, uint8_t* mask,
                                          uint8_t* mask1, uint8_t* mask2, const float* labels, unsigned long long* counts,
                                          void* stream) {
  using namespace ctu;
  InvertGeomDev g;
  const int rc = invert_geom(geom, g);
  if (rc != 0) return rc;
  if (!p1 || !p2 || (!mask && !mask1 && !mask2 && !counts)) return CTU_E_BADARG;
  if (C != 14) return CTU_E_UNSUPPORTED;
  const long long V = (long long)g.out[0] * g.out[1] * g.out[2];
  invert_ensemble_kernel<14><<<invert_grid(V, 16), 256, 0, (cudaStream_t)stream>>>(p1, p2, g, mask, mask1, mask2, labels, counts);
  count_launch();
  return (int)cudaGetLastError();
}
