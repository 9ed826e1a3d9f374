// Host-side helpers shared by the launchers: driver entry point for TMA descriptor encoding, launch counter.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <atomic>
#include <stdint.h>

struct ctu_gemm_desc;
struct ctu_wgrad_desc;

namespace ctu {

typedef CUresult (*tma_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled, looked up through the runtime so the library does not link against libcuda.
tma_encode_fn tma_encoder();
void count_launch(int n = 1);
// SMs a persistent tensor-core kernel may occupy (ctu_set_persistent_sm_limit; the device's SM count by default)
int persistent_sms(int device_sms);

// 3x3x3 convolution with shared-memory halo reuse (umma_conv3_halo.cu); CTU_E_UNSUPPORTED when not applicable.
int conv3_halo_dispatch(const ::ctu_gemm_desc* d, cudaStream_t stream);
// its weight-gradient counterpart (umma_wgrad_halo.cu)
int wgrad_halo_dispatch(const ::ctu_wgrad_desc* d, cudaStream_t stream);

// Launch with programmatic dependent launch (PDL): the kernel may be scheduled while the previous kernel of the stream
// is still draining; it must execute pdl_wait() (griddepcontrol.wait) before its first access to global memory that a
// predecessor may have written, and calls pdl_trigger() early so that ITS successor can do the same.  Saves the
// launch + prologue latency (barrier init, TMEM allocation, descriptor prefetch) between kernels.  Measured on B200: no
// gain for this workload (the ~16 us floor of the small GEMMs is not launch latency), so it is OFF unless CTU_PDL=1.
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace ctu
