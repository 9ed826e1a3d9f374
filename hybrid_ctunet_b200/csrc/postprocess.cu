// Connected-component post-processing of a predicted label volume on the device (test_CTUNet_final.py:132-190,
// `remove_all_but_the_largest_connected_component`): the reference builds `mask = image == c` (or the union of a class
// group), labels it with scipy.ndimage.label (default structure: 6-connectivity), sizes every object with one full-volume
// pass per object and zeroes all but the largest one(s) on the host.  Here: a lock-free union-find over the voxels of the
// mask (root = smallest voxel index of the component), one size histogram keyed by root, one filter pass.  Integer work:
// the kept / removed voxel sets are identical to the reference's, whatever numbering scipy gives the objects.
// HBM-bound: ~26 B per voxel over the five passes (uint8 labels, int32 parents, int32 sizes).
#include "common.cuh"
#include "../../include/ctunet_b200.h"
#include "host_util.h"

namespace ctu {

__device__ __forceinline__ int cc_find(const int* __restrict__ parent, int x) {
  int p = parent[x];
  while (p != x) {
    x = p;
    p = parent[x];
  }
  return x;
}

// Lock-free union by index: the larger root is pointed at the smaller one with an atomicMin; a lost race retries from
// the value that won.
__device__ __forceinline__ void cc_union(int* parent, int a, int b) {
  while (true) {
    a = cc_find(parent, a);
    b = cc_find(parent, b);
    if (a == b) return;
    if (a < b) { const int t = a; a = b; b = t; }   // a > b: hang a under b
    const int old = atomicMin(&parent[a], b);
    if (old == a) return;
    a = old;                                         // someone re-rooted a in the meantime: merge that root with b
  }
}

__global__ void __launch_bounds__(256) cc_init_kernel(const uint8_t* __restrict__ image, const uint8_t* __restrict__ member,
                                                      long long V, int* __restrict__ parent) {
  __shared__ uint8_t m[256];
  m[threadIdx.x] = member[threadIdx.x];
  __syncthreads();
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (long long)gridDim.x * blockDim.x)
    parent[v] = m[image[v]] ? (int)v : -1;
}

// every voxel of the mask is merged with its three "backward" neighbours (x-1, y-1, z-1) that are in the mask
__global__ void __launch_bounds__(256) cc_merge_kernel(int* __restrict__ parent, int X, int Y, int Z) {
  const long long V = (long long)X * Y * Z;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (long long)gridDim.x * blockDim.x) {
    if (parent[v] < 0) continue;
    const int z = (int)(v % Z), y = (int)((v / Z) % Y), x = (int)(v / ((long long)Z * Y));
    if (z > 0 && parent[v - 1] >= 0) cc_union(parent, (int)v, (int)(v - 1));
    if (y > 0 && parent[v - Z] >= 0) cc_union(parent, (int)v, (int)(v - Z));
    if (x > 0 && parent[v - (long long)Z * Y] >= 0) cc_union(parent, (int)v, (int)(v - (long long)Z * Y));
  }
}

// parent[v] <- root of v; sizes[root] += 1
__global__ void __launch_bounds__(256) cc_flatten_kernel(int* __restrict__ parent, long long V, int* __restrict__ sizes) {
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (long long)gridDim.x * blockDim.x) {
    if (parent[v] < 0) continue;
    const int r = cc_find(parent, (int)v);
    parent[v] = r;
    atomicAdd(&sizes[r], 1);
  }
}

// summary[0] = number of objects, summary[1] = size of the largest one
__global__ void __launch_bounds__(256) cc_summary_kernel(const int* __restrict__ parent, const int* __restrict__ sizes,
                                                         long long V, int* __restrict__ summary) {
  int n = 0, mx = 0;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (long long)gridDim.x * blockDim.x) {
    if (parent[v] == (int)v) {
      ++n;
      mx = max(mx, sizes[v]);
    }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    n += __shfl_xor_sync(0xffffffffu, n, o);
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) {
    if (n) atomicAdd(&summary[0], n);
    if (mx) atomicMax(&summary[1], mx);
  }
}

// image[v] = 0 for every voxel of an object that is not (one of) the largest and — with a size threshold — smaller than it;
// summary[2] = size of the largest object removed.  The size comparison is the reference's float64 one: count * vpv.
__global__ void __launch_bounds__(256) cc_filter_kernel(uint8_t* __restrict__ image, const int* __restrict__ parent,
                                                        const int* __restrict__ sizes, long long V, double vpv,
                                                        double min_valid, int has_min, int* __restrict__ summary) {
  const int mx = summary[1];
  int removed = 0;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (long long)gridDim.x * blockDim.x) {
    const int r = parent[v];
    if (r < 0) continue;
    const int s = sizes[r];
    if (s == mx) continue;
    if (has_min && !((double)s * vpv < min_valid)) continue;
    image[v] = 0;
    removed = max(removed, s);
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) removed = max(removed, __shfl_xor_sync(0xffffffffu, removed, o));
  if ((threadIdx.x & 31) == 0 && removed) atomicMax(&summary[2], removed);
}

static int cc_grid(long long V) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long g = (V + 255) / 256;
  if (g > (long long)sms * 16) g = (long long)sms * 16;
  return (int)(g < 1 ? 1 : g);
}

}  // namespace ctu

extern "C" int ctu_cc_filter_largest(uint8_t* image, const uint8_t* member, int X, int Y, int Z, double volume_per_voxel,
                                     int has_min, double min_valid, int* parent, int* sizes, int* summary, void* stream_) {
  using namespace ctu;
  if (!image || !member || !parent || !sizes || !summary || X <= 0 || Y <= 0 || Z <= 0) return CTU_E_BADARG;
  const long long V = (long long)X * Y * Z;
  if (V > 0x7fffffffLL) return CTU_E_UNSUPPORTED;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int grid = cc_grid(V);
  cudaError_t e = cudaMemsetAsync(sizes, 0, (size_t)V * sizeof(int), stream);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemsetAsync(summary, 0, 4 * sizeof(int), stream);
  if (e != cudaSuccess) return (int)e;
  cc_init_kernel<<<grid, 256, 0, stream>>>(image, member, V, parent);
  cc_merge_kernel<<<grid, 256, 0, stream>>>(parent, X, Y, Z);
  cc_flatten_kernel<<<grid, 256, 0, stream>>>(parent, V, sizes);
  cc_summary_kernel<<<grid, 256, 0, stream>>>(parent, sizes, V, summary);
  cc_filter_kernel<<<grid, 256, 0, stream>>>(image, parent, sizes, V, volume_per_voxel, min_valid, has_min, summary);
  count_launch(5);
  return (int)cudaGetLastError();
}
