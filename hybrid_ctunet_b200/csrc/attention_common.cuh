// Token addressing and the mma.sync wrapper shared by the attention forward and backward kernels.
#pragma once
#include "common.cuh"

namespace ctu {

struct TokenMap {
  int mode;           // 0: rows = win*n + p; 1: block partition '(h h1)'; 2: grid partition '(h1 h)'
  int X, Y, Z;        // token grid of one batch item
  int nwx, nwy, nwz;  // windows per axis
  int w;              // window edge (6)
};

__device__ __forceinline__ long long token_row(const TokenMap& m, int win, int p, int n) {
  if (m.mode == 0) return (long long)win * n + p;
  const int wz = win % m.nwz;
  int t = win / m.nwz;
  const int wy = t % m.nwy;
  t /= m.nwy;
  const int wx = t % m.nwx;
  const int b = t / m.nwx;
  const int pz = p % m.w;
  const int py = (p / m.w) % m.w;
  const int px = p / (m.w * m.w);
  int x, y, z;
  if (m.mode == 1) {
    x = wx * m.w + px; y = wy * m.w + py; z = wz * m.w + pz;
  } else {
    x = px * m.nwx + wx; y = py * m.nwy + wy; z = pz * m.nwz + wz;
  }
  return (((long long)b * m.X + x) * m.Y + y) * m.Z + z;
}

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

}  // namespace ctu
